#!/bin/bash
# Round of GPU checks; every stage in its own process (a trapped kernel poisons the CUDA context) and under timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary.txt; }
[ -z "$SKIP_PROBE" ] && run probe python scripts/probe_tc.py
run t_elementwise python -m pytest tests/test_gpu_elementwise.py -m gpu -q --timeout 120
run t_tensorcore python -m pytest tests/test_gpu_tensorcore.py -m gpu -q --timeout 120
run t_model python -m pytest tests/test_gpu_model.py -m gpu -q -s --timeout 300
run smoke python __graft_entry__.py smoke
run bench python bench.py --steps 5 --warmup 3
for f in t_elementwise t_tensorcore t_model smoke bench; do echo "=== $f"; tail -25 gpurun_out/$f.log; done
exit 0
