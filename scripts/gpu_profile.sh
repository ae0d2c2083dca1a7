#!/bin/bash
# ncu evidence for one bench command: (1) launch list with device time per launch, (2) --set full capture of the
# kernels named in $KERNELS (regex). Each ncu pass only after the same command exited 0 without ncu.
# usage: TAG=r1a KERNELS='igemm_kernel|wgrad_kernel' bash scripts/gpu_profile.sh
mkdir -p gpurun_out
TAG=${TAG:-prof}
KERNELS=${KERNELS:-igemm_kernel}
SKIP=${SKIP:-4800}     # launches of the 3 warm-up steps + setup
COUNT=${COUNT:-1400}   # a little over one step
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $COUNT --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:$KERNELS" -s ${FULL_SKIP:-120} -c ${FULL_COUNT:-12} \
    -f -o gpurun_out/${TAG}_full $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -20
