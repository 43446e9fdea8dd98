#!/bin/bash
# round-2 visit G (1 GPU): full parity suite, the driver-shaped bench line, reference arm, ncu launch list of the same
# command, ncu --set full of the three particle kernels in the TIMED regime (cycle >= 20 of the bench's synthetic run)
TAG=${1:-r2g}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/smoke.log
timeout 1200 python bench.py --steps 20 --warmup 5 --ablate deposit_variant=0:2 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -c 400 $OUT/bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-regimes --no-e2e > $OUT/bench_under_ncu.log 2>&1; echo "ncu list exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_deposit|k_geodesic|k_scatter' -s 60 -c 3 -o $OUT/particles_timed_regime python bench.py --steps 1 --warmup 21 --no-cpu-baseline --no-regimes --no-e2e > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "roofline", d["roofline"])
for k,v in sorted(d["kernels"].items(), key=lambda kv:-kv[1]["ms_per_step"]):
    print(f"{k:32s} {v['ms_per_step']:8.3f} ms/step calls {v['calls_per_step']:.0f} frac {v.get('frac',float('nan')):.3f}")
print(d["cpu_baseline"])
PY
ls -la $OUT
GEVB_DEPOSIT_VARIANT=2 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or checker or extreme or time_loop_one or N128 or config1 or empty" > $OUT/pytest_gpu_variant2.log 2>&1; echo "pytest variant 2 exit $?"; tail -2 $OUT/pytest_gpu_variant2.log | cut -c1-200
GEVB_DEPOSIT_VARIANT=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_deposit' -s 20 -c 1 -o $OUT/dep_v2b python bench.py --steps 1 --warmup 20 --no-cpu-baseline --no-regimes --no-e2e > $OUT/ncu_v2b.log 2>&1; echo "ncu v2b exit $?"
grep -h ablate $OUT/bench.err | cut -c1-200
