#!/bin/bash
# weak and strong scaling records on ONE box with up to 8 GPUs (run under `gpurun --gpus 8`):
#   weak:   per-GPU work fixed (bench.py defaults), N = 1, 2, 4, 8
#   strong: TOTAL conditions fixed (--scaling strong: 1184 Chen-shape / 10^4 Boehm-shape), N = 1, 2, 4, 8
# Output: gpurun_out/r02_scale_<workload>_<weak|strong>_<N>.json
mkdir -p gpurun_out
NMAX=${1:-8}
STEPS=${2:-3}
for W in chen boehm; do
  for MODE in strong weak; do
    for N in 1 2 4 8; do
      [ $N -gt $NMAX ] && continue
      out=gpurun_out/r02_scale_${W}_${MODE}_${N}.json
      if [ $N -eq 1 ]; then
        timeout 900 python bench.py --gpus 1 --steps $STEPS --warmup 3 --workload $W --scaling $MODE --no-extra --no-cpu-baseline > $out 2> ${out%.json}.err
      else
        timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
          bench.py --gpus $N --steps $STEPS --warmup 3 --workload $W --scaling $MODE --no-extra --no-cpu-baseline > $out 2> ${out%.json}.err
      fi
      echo "$W $MODE N=$N rc=$? $(python -c "import json,sys; d=json.loads(open('$out').read().strip().splitlines()[-1]); print('value %.1f e2e %.1f ms/step %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step']))" 2>/dev/null)"
    done
  done
done
