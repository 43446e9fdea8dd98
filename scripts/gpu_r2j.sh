#!/bin/bash
# round-2 visit J (1 GPU): new defaults (bulk-reduce tile flush, windowed re-bin move with prefetch): full parity suite,
# driver-shaped bench line, ablation of both knobs, ncu of the re-bin move in the timed regime
TAG=${1:-r2j}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/smoke.log
timeout 1200 python bench.py --steps 20 --warmup 5 --ablate deposit_variant=4:0:4:0,rebin_variant=2:0:2:0 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -c 300 $OUT/bench.err
grep -h ablate $OUT/bench.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); m=d['ms']; print(d['ablate'], d['value'], 'deposit', m.get('projection_T00_Tij_project'), 'kick', m.get('kick_drift'), 'rebin', m.get('rebin_sort'))
"
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "roofline", d["roofline"])
for k,v in sorted(d["kernels"].items(), key=lambda kv:-kv[1]["ms_per_step"]):
    print(f"{k:32s} {v['ms_per_step']:8.3f} ms/step calls {v['calls_per_step']:.0f} frac {v.get('frac',float('nan')):.3f}")
print(d["cpu_baseline"]); print(d["config"].get("e2e_parts_rank0"), d["config"].get("host_binding_rank0"))
r=d["config"].get("regimes") or d.get("regimes"); print(json.dumps(r)[:1500])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_scatter|k_deposit' -s 40 -c 2 -o $OUT/scatter_deposit python bench.py --steps 1 --warmup 20 --no-cpu-baseline --no-regimes --no-e2e > $OUT/ncu.log 2>&1; echo "ncu exit $?"
