#!/bin/bash
# weak-scaling sweep on one box: bench.py at N = 1, 2, 4, 8 (whatever the box has), one JSON line each
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for n in ${NS:-1 2 4 8}; do
  [ $n -gt $NG ] && continue
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline > gpurun_out/scale_$n.log 2>&1
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
      bench.py --gpus $n --steps ${STEPS:-10} --warmup 3 > gpurun_out/scale_$n.log 2>&1
  fi
  echo "N=$n rc=$?"
  grep '^{' gpurun_out/scale_$n.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms', 'e2e', round(d['e2e']['value'],1), d['clocks'])" || tail -5 gpurun_out/scale_$n.log
done
