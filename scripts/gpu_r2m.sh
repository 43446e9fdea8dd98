#!/bin/bash
# round-2 visit M (1 GPU): tile flush by one TMA tensor reduction per component (deposit 18) against the row-wise bulk flush (10)
TAG=${1:-r2m}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
for v in 18 10; do
GEVB_DEPOSIT_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or checker or extreme or time_loop or N128 or empty or full_size or config1" > $OUT/pytest_gpu_deposit$v.log 2>&1; echo "pytest deposit variant $v exit $?"; tail -3 $OUT/pytest_gpu_deposit$v.log | cut -c1-400
done
timeout 900 python bench.py --steps 10 --warmup 15 --no-cpu-baseline --no-e2e --no-regimes --ablate deposit_variant=10:18:4:10:18:4 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -c 300 $OUT/bench.err
grep -h ablate $OUT/bench.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); m=d['ms']; print(d['ablate'], d['value'], 'deposit', m.get('projection_T00_Tij_project'), 'kick', m.get('kick_drift'), 'rebin', m.get('rebin_sort'))
"
GEVB_DEPOSIT_VARIANT=18 timeout 900 python bench.py --steps 10 --warmup 15 --no-cpu-baseline --no-e2e > $OUT/bench_d18.json 2> $OUT/bench_d18.err; echo "bench d18 exit $?"; tail -c 300 $OUT/bench_d18.err
python - <<PY
import json
for f in ("bench.json","bench_d18.json"):
    d=json.load(open("$OUT/"+f))
    print(f, "ms_per_step", d["ms_per_step"], {k:round(v["ms_per_step"],2) for k,v in d["kernels"].items() if v["ms_per_step"]>0.4})
    r=d["config"].get("regimes") or d.get("regimes")
    if r: print("  regimes", {k:{kk:vv for kk,vv in v.items() if kk.endswith('_ms')} for k,v in r.items() if isinstance(v,dict)})
PY
GEVB_DEPOSIT_VARIANT=18 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_deposit' -s 20 -c 1 -o $OUT/deposit_d18 python bench.py --steps 1 --warmup 20 --no-cpu-baseline --no-regimes --no-e2e > $OUT/ncu.log 2>&1; echo "ncu exit $?"
