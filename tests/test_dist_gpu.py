"""Multi-GPU parity (needs >= 2 CUDA devices): the sharded LM of DESIGN.md §4, launched one process per GPU
with torchrun, against the single-GPU solve and the CPU oracle on the same graph."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("nodes,loops,port", [(1500, 300, 29541), (3000, 900, 29542)])
def test_two_gpu_solve_matches_single_gpu_and_oracle(tmp_path, nodes, loops, port):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs (run under `gpurun --gpus 2`); the 1-GPU path is covered by test_gpu_parity.py")
    out = tmp_path / "dist.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "dist_solve.py"), "--config", "3", "--nodes", str(nodes), "--loops", str(loops),
           "--oracle", "--out", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    d = json.load(open(out))
    assert d["ranks"][0]["n_border_nodes"] > 0 and d["ranks"][0]["n_collectives"] > 0
    for key in ("dist_vs_single", "dist_vs_oracle"):
        c = d[key]
        assert c["max_dt"] < 1e-5 and c["max_drot"] < 1e-4 and c["switch_states_equal"] and c["rel_cost"] < 1e-5, (key, c)
    assert d["dist_vs_single"]["same_trajectory"]
