#!/bin/bash
# round-2 visit W (1 GPU): own x-pass of the forward transform with prepareFTsource fused into its load (xpass.cu)
TAG=${1:-r2w}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "own_xpass" > $OUT/pytest_xpass.log 2>&1; echo "pytest xpass exit $?"; tail -15 $OUT/pytest_xpass.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e --no-regimes --ablate fft_xpass=1:0:1:0 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -c 600 $OUT/bench.err
grep -h ablate $OUT/bench.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); m=d['ms']; print(d['ablate'], d['value'], {k:v for k,v in m.items() if 'fft' in k or 'prepare' in k or 'Poisson' in k})
"
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("ms_per_step", d["ms_per_step"], {k:round(v["ms_per_step"],2) for k,v in d["kernels"].items() if v["ms_per_step"]>0.4}, d["invariants"])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_xpass' -s 6 -c 3 -o $OUT/xpass python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-regimes --no-e2e > $OUT/ncu.log 2>&1; echo "ncu exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-regimes --no-e2e > $OUT/bench_under_ncu.log 2>&1; echo "ncu list exit $?"
