#!/bin/bash
# weak and strong scaling records at N GPUs of ONE box (run under `gpurun --gpus N`; one call per N keeps the
# charged GPU-minutes proportional to the work):
#   weak:   per-GPU work fixed (bench.py defaults: 148 Chen-shape / 18 944 Boehm-shape conditions per GPU)
#   strong: TOTAL conditions fixed (--scaling strong: 1184 Chen-shape / 10^4 Boehm-shape)
# usage: scaling.sh N [STEPS]      output: gpurun_out/r02_scale_<workload>_<weak|strong>_<N>.json
mkdir -p gpurun_out
N=${1:-1}
STEPS=${2:-2}
for W in chen boehm; do
  for MODE in strong weak; do
    out=gpurun_out/r02_scale_${W}_${MODE}_${N}.json
    ARGS="--gpus $N --steps $STEPS --warmup 3 --workload $W --scaling $MODE --no-extra --no-cpu-baseline --no-cost-only"
    if [ $N -eq 1 ]; then
      timeout 900 python bench.py $ARGS > $out 2> ${out%.json}.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py $ARGS > $out 2> ${out%.json}.err
    fi
    echo "$W $MODE N=$N rc=$? $(python -c "import json,sys; d=json.loads(open('$out').read().strip().splitlines()[-1]); print('value %.1f e2e %.1f ms/step %.1f conditions %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['conditions_total']))" 2>/dev/null)"
  done
done
