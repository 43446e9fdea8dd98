#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list, ncu full captures of the top kernels.
# usage (from the repo root, under gpurun): bash scripts/gpu_round.sh <tag> [steps]
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/smoke.log
echo "== bench"; timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -c 600 $OUT/bench.err; head -c 1500 $OUT/bench.json
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref exit $?"; cat $OUT/bench_reference.json | head -c 1200
echo "== ncu launch list (same command, bounded)"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-regimes > $OUT/bench_under_ncu.log 2>&1; echo "ncu list exit $?"
echo "== ncu full (512^3, the step's own top kernels: one launch each of the 4th cycle)"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_deposit|k_geodesic|k_scatter|k_prepare_tensor' -s 13 -c 4 -o $OUT/prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-regimes > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la $OUT
