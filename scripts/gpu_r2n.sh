#!/bin/bash
# round-2 visit N (1 GPU): candidate final state (deposit 18 = spare cells + run-time shuffle loop + solo tail batch + TMA tensor
# flush; windowed re-bin): full parity suite, smoke, driver-shaped bench, reference arm, launch list, timed-regime ncu captures
TAG=${1:-r2n}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/smoke.log
timeout 1200 python bench.py --steps 20 --warmup 5 --ablate deposit_variant=18:10:4:0,geodesic_variant=5:6,rebin_variant=2:0 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -c 300 $OUT/bench.err
grep -h ablate $OUT/bench.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); m=d['ms']; print(d['ablate'], d['value'], 'deposit', m.get('projection_T00_Tij_project'), 'kick', m.get('kick_drift'), 'rebin', m.get('rebin_sort'))
"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "roofline", d["roofline"])
for k,v in sorted(d["kernels"].items(), key=lambda kv:-kv[1]["ms_per_step"]):
    print(f"{k:32s} {v['ms_per_step']:8.3f} ms/step calls {v['calls_per_step']:.0f} frac {v.get('frac',float('nan')):.3f}")
r=d["config"].get("regimes") or d.get("regimes")
print("  regimes", {k:{kk:vv for kk,vv in v.items() if kk.endswith('_ms')} for k,v in r.items() if isinstance(v,dict)})
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-regimes --no-e2e > $OUT/bench_under_ncu.log 2>&1; echo "ncu list exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_deposit|k_geodesic|k_scatter' -s 60 -c 3 -o $OUT/particles_timed_regime python bench.py --steps 1 --warmup 21 --no-cpu-baseline --no-regimes --no-e2e > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"
