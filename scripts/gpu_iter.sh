#!/bin/bash
# Development iteration on the GPU box: probe, tensor-core parity tests, model parity, layer microbench, bench.
# Each stage in its own process under timeout (a trapped kernel poisons its CUDA context only).
mkdir -p gpurun_out
: > gpurun_out/iter_summary.txt
run() { name=$1; shift; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/iter_summary.txt; }
for st in ${STAGES:-probe tc model layers bench}; do
  case $st in
    probe) run probe_shift python -m pytest tests/test_gpu_probe.py -m gpu -q ;;
    ew) run t_elementwise python -m pytest tests/test_gpu_elementwise.py -m gpu -q -x --timeout 120 ;;
    tc) run t_tensorcore python -m pytest tests/test_gpu_tensorcore.py -m gpu -q --timeout 120 ;;
    model) run t_model python -m pytest tests/test_gpu_model.py -m gpu -q -s --timeout 300 ;;
    all) run t_all python -m pytest tests -m gpu -q --timeout 300 ;;
    layers) ONLY="${ONLY:-fprop,dgrad}" run layers python scripts/bench_layers.py ;;
    layers_nores) B200UNET_NO_RES=1 ONLY="${ONLY:-fprop,dgrad}" run layers_nores python scripts/bench_layers.py ;;
    step) run profile_step python scripts/profile_step.py ;;
    bench) run bench python bench.py --steps 10 --warmup 3 ;;
    smoke) run smoke python __graft_entry__.py smoke ;;
  esac
done
for f in gpurun_out/probe_shift.log gpurun_out/t_tensorcore.log gpurun_out/t_model.log; do [ -f $f ] && { echo "=== $f"; tail -15 $f; }; done
cat gpurun_out/iter_summary.txt
exit 0
