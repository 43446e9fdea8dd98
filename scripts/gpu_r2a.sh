#!/bin/bash
# round-2 visit A: parity tests, racecheck of the smoke cycle, bench with the deposit ablation
TAG=${1:-r2a}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log | cut -c1-400
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/racecheck.log 2>&1; echo "racecheck exit $?"; tail -6 $OUT/racecheck.log | cut -c1-300
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/memcheck.log 2>&1; echo "memcheck exit $?"; tail -4 $OUT/memcheck.log | cut -c1-300
timeout 900 python bench.py --no-cpu-baseline --steps 10 --warmup 5 --ablate deposit_variant=1:0 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -c 1500 $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
for k,v in sorted(d["kernels"].items(), key=lambda kv:-kv[1]["ms_per_step"]):
    print(f"{k:32s} {v['ms_per_step']:8.3f} ms/step calls {v['calls_per_step']:.0f} frac {v.get('frac',float('nan')):.3f}")
print(json.dumps(d["regimes"], indent=1)[:1500])
PY
