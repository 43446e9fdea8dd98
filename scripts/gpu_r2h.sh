#!/bin/bash
# round-2 visit H (1 GPU): k_deposit with bulk-reduce flush (variant 3), split phase barrier (4), both (5):
# parity subset per variant, ablation in the timed regime, regimes for the combined variant
TAG=${1:-r2h}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
for v in 5 3 4; do
GEVB_DEPOSIT_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or checker or extreme or time_loop_one or N128 or empty" > $OUT/pytest_gpu_variant$v.log 2>&1; echo "pytest variant $v exit $?"; tail -2 $OUT/pytest_gpu_variant$v.log | cut -c1-300
done
timeout 900 python bench.py --steps 10 --warmup 15 --no-cpu-baseline --no-regimes --no-e2e --ablate deposit_variant=0:3:4:5:0:5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -c 300 $OUT/bench.err
grep -h ablate $OUT/bench.err | cut -c1-260
GEVB_DEPOSIT_VARIANT=5 timeout 900 python bench.py --steps 10 --warmup 15 --no-cpu-baseline --no-e2e > $OUT/bench_v5.json 2> $OUT/bench_v5.err; echo "bench v5 exit $?"
python - <<PY
import json
for f in ("bench.json","bench_v5.json"):
    d=json.load(open("$OUT/"+f))
    print(f, "ms_per_step", d["ms_per_step"], d["roofline"])
    print("  regimes", d["config"].get("regimes") or d.get("regimes"))
PY
GEVB_DEPOSIT_VARIANT=5 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_deposit' -s 20 -c 1 -o $OUT/dep_v5 python bench.py --steps 1 --warmup 20 --no-cpu-baseline --no-regimes --no-e2e > $OUT/ncu_v5.log 2>&1; echo "ncu v5 exit $?"
