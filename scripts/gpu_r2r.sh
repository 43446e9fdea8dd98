#!/bin/bash
# round-2 visit R (1 GPU): field tiles of the kick/drift kernel fetched by TMA tensor loads (geodesic_tma) against LDGSTS staging
TAG=${1:-r2r}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_gpu.log | cut -c1-300
timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "time_loop_one" > $OUT/racecheck.log 2>&1; echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" $OUT/racecheck.log | tail -3
timeout 1200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --ablate geodesic_tma=1:0:1:0:1:0 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -c 300 $OUT/bench.err
grep -h ablate $OUT/bench.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); m=d['ms']; print(d['ablate'], d['value'], 'deposit', m.get('projection_T00_Tij_project'), 'kick', m.get('kick_drift'), 'rebin', m.get('rebin_sort'))
"
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("ms_per_step", d["ms_per_step"], {k:round(v["ms_per_step"],2) for k,v in d["kernels"].items() if v["ms_per_step"]>0.4})
r=d["config"].get("regimes") or d.get("regimes")
print("  regimes", {k:{kk:vv for kk,vv in v.items() if kk.endswith('_ms')} for k,v in r.items() if isinstance(v,dict)})
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_geodesic' -s 20 -c 1 -o $OUT/geodesic_tma python bench.py --steps 1 --warmup 20 --no-cpu-baseline --no-regimes --no-e2e > $OUT/ncu.log 2>&1; echo "ncu exit $?"
