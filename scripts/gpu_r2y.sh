#!/bin/bash
# round-2 visit Y (1 GPU): spare cells in front of the component tiles (TMA box 18 x 9 x 5 instead of 18 x 11 x 5); kick/drift block shapes with TMA tiles
TAG=${1:-r2y}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
for v in 18 10 8; do
GEVB_DEPOSIT_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or checker or extreme or time_loop or N128 or empty or config1" > $OUT/pytest_gpu_deposit$v.log 2>&1; echo "pytest deposit variant $v exit $?"; tail -2 $OUT/pytest_gpu_deposit$v.log | cut -c1-300
done
timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "checker and 16" > $OUT/racecheck.log 2>&1; echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" $OUT/racecheck.log | tail -3
timeout 1200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --ablate deposit_variant=18:10:18:10,geodesic_variant=6:7:5:6:7:5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -c 300 $OUT/bench.err
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
