mkdir -p gpurun_out
for c in n2r2 n2r1 n1r1 n1r2; do echo "cfg $c"; ./tools/bin/upd_lab_$c 4000 600 2>&1 | grep -E "^N=" | tail -1; ./tools/bin/upd_lab_$c 5000 800 2>&1 | grep -E "^N=" | tail -1; done
