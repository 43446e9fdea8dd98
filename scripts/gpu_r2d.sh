#!/bin/bash
# round-2 visit D: thread-per-cell deposit (variant 2): parity tests under the variant, racecheck, ablation, ncu
TAG=${1:-r2d}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export GEVB_DEPOSIT_VARIANT=2
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log | cut -c1-400
timeout 300 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/racecheck.log 2>&1; echo "racecheck exit $?"; tail -3 $OUT/racecheck.log | cut -c1-300
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/memcheck.log 2>&1; echo "memcheck exit $?"; tail -3 $OUT/memcheck.log | cut -c1-300
timeout 900 python bench.py --no-cpu-baseline --steps 10 --warmup 5 --ablate deposit_variant=2:0 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -c 1200 $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
for k,v in sorted(d["kernels"].items(), key=lambda kv:-kv[1]["ms_per_step"])[:6]:
    print(f"{k:32s} {v['ms_per_step']:8.3f} ms/step calls {v['calls_per_step']:.0f} frac {v.get('frac',float('nan')):.3f}")
for r in ("lattice","clustered"): print(r, {k:v for k,v in d["regimes"][r].items() if k.endswith("_ms")})
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_deposit' -s 10 -c 1 -o $OUT/dep_v2 python bench.py --steps 1 --warmup 10 --no-cpu-baseline --no-regimes > $OUT/ncu_v2.log 2>&1; echo "ncu v2 exit $?"
ls -la $OUT
