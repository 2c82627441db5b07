timeout 600 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -3
run() { PB200_KERNEL_VARIANT=$1 PB200_BENCH_CONDITIONS=$3 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --lanes $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('variant=$1 lanes=$2 C=$3 rep=$PB200_REPLICATE_CONDITIONS solves/s',round(d['value']),'ms',round(d['ms_per_step'],1),'grid',d['config']['grid'],'block',d['config']['block'], d['nllh'])"; }
run 2 8 16384
run 2 4 16384
run 0 8 16384
export PB200_REPLICATE_CONDITIONS=4; run 2 8 16384
