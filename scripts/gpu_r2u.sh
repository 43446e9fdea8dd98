#!/bin/bash
# round-2 visit U (1 GPU): L2 promotion of the tensor maps (kick/drift tile loads, deposit tile reductions)
TAG=${1:-r2u}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 1200 python bench.py --steps 10 --warmup 15 --no-cpu-baseline --no-e2e --no-regimes --ablate tma_l2_promotion=0:2:3:1:0:2:3:1 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -c 300 $OUT/bench.err
grep -h ablate $OUT/bench.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); m=d['ms']; print(d['ablate'], d['value'], 'deposit', m.get('projection_T00_Tij_project'), 'kick', m.get('kick_drift'), 'rebin', m.get('rebin_sort'))
"
