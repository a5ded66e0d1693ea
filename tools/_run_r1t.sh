mkdir -p gpurun_out
cp solve_keyframe_pose_graph_b200/libpgs.so /tmp/libpgs_orig.so
for mb in 1 2 3 4; do cp solve_keyframe_pose_graph_b200/libpgs_mb$mb.so solve_keyframe_pose_graph_b200/libpgs.so; for rep in 1 2; do python bench.py --no-lm --no-cpu-baseline --steps 50 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('minb $mb', round(d['value']/1e9,3), round(d['ms_per_step']*1e3,2), round(d['roofline']['frac'],4), round(d['value_warm_l2']/1e9,3))"; done; done
cp /tmp/libpgs_orig.so solve_keyframe_pose_graph_b200/libpgs.so
