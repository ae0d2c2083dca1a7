#!/bin/bash
# ncu evidence for the bench command (1 GPU): (1) launch list with the device time of every launch of ~2 steps,
# (2) --set full captures of the dominant kernels. Every ncu pass directly follows a plain run that exited 0.
# usage: TAG=r1 bash scripts/gpu_profile.sh        (outputs under gpurun_out/, summaries are made on the CPU box)
#        TAG=r2_attn CONFIG=attn512 CONVDRAM=0 FULL=gate COUNT=1200 bash scripts/gpu_profile.sh   (the attention-gated network)
mkdir -p gpurun_out
TAG=${TAG:-prof}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --config ${CONFIG:-train512}"
[ "${LAUNCHES:-1}" = 1 ] && $CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-900} -c ${COUNT:-700} --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
full() {  # name, kernel regex, skip (matching launches), count, extra flags
  $CMD > gpurun_out/${TAG}_plain_$1.log 2>&1 &&
  ncu --set full --clock-control none $5 -k "regex:$2" -s $3 -c $4 -f -o gpurun_out/${TAG}_$1 $CMD \
      > gpurun_out/${TAG}_ncu_$1.log 2>&1
  echo "full capture $1 rc=$?"
}
# per-launch DRAM bytes of every conv3x3 fprop/dgrad launch of one step. ncu matches -k against the function name without
# template arguments, and conv3_res_kernel also runs the first-layer and ConvTranspose2d GEMMs: 40 matching launches per
# step (34 conv3x3 + 6 others; scripts/conv_traffic.py keeps the 34); 4 steps precede (1 eager + 3 warm-up).
if [ "${CONVDRAM:-1}" = 1 ]; then
$CMD > gpurun_out/${TAG}_plain_convdram.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k "regex:conv3_pair_kernel|conv3_res2_kernel|conv3_res_kernel" -s 160 -c 40 --csv \
    --log-file gpurun_out/${TAG}_convdram.csv $CMD > gpurun_out/${TAG}_ncu_convdram.log 2>&1
fi
echo "conv dram rc=$?"
for what in ${FULL-pair res wgrad bn}; do   # FULL="" skips the full captures
  case $what in
    pair)  full pair  'conv3_pair_kernel' 80 7 "" ;;
    res)   full res   'conv3_res2_kernel|conv3_res_kernel' 96 8 "" ;;
    wgrad) full wgrad 'wgrad_kernel|wgrad_swap_kernel' 88 6 "" ;;
    bn)    full bn    'bn_bwd_apply_kernel|bn_bwd_reduce_kernel|bn_relu_fwd_kernel' 216 5 "" ;;
    gate)  full gate  'gate_psi_fwd_kernel|gate_apply_fwd_kernel|gate_apply_bwd_kernel|gate_bwd_reduce_kernel|gate_bwd_apply_kernel' 80 20 "" ;;  # CONFIG=attn512: the 20 gate launches of the timed step
  esac
done
# gpurun copies back at most 64 MiB: drop the largest reports until the directory fits
while [ $(du -sm gpurun_out | cut -f1) -gt 58 ]; do big=$(ls -S gpurun_out/*.ncu-rep | head -1); echo "dropping $big"; rm -f $big; done
ls -la gpurun_out | tail -20
