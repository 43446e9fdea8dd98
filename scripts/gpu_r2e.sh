#!/bin/bash
# round-2 visit E (multi-GPU): slab parity tests, bench line with the in-bench parity check, small runs of configs 4 and 5
TAG=${1:-r2e}; NG=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 tests/multigpu_worker.py > $OUT/multigpu_parity_$NG.log 2>&1; echo "parity exit $?"; grep -E "BAD|MULTIGPU|Error|error" $OUT/multigpu_parity_$NG.log | head -10
timeout 900 $TR --master-port 29512 bench.py --gpus $NG --no-cpu-baseline > $OUT/bench_$NG.json 2> $OUT/bench_$NG.err; echo "bench exit $?"; tail -c 600 $OUT/bench_$NG.err
for c in 4 5; do
timeout 900 $TR --master-port 2951$c bench.py --gpus $NG --config $c --ngrid 128 --steps 4 --warmup 3 --no-cpu-baseline --no-parity > $OUT/config${c}_small_$NG.json 2> $OUT/config${c}_small_$NG.err; echo "config $c small exit $?"; tail -c 600 $OUT/config${c}_small_$NG.err
done
python - <<PY
import json
for f in ("bench_$NG","config4_small_$NG","config5_small_$NG"):
    try:
        d=json.loads(open("$OUT/"+f+".json").read().strip().splitlines()[-1])
        print(f, "ms_per_step", round(d["ms_per_step"],3), "value %.3e" % d["value"], "e2e", d["e2e"]["value"], "parity", d.get("parity_check"), "inv", d["invariants"], d["config"].get("ncdm_substeps"), d["config"].get("hij_spectrum_call_ms"))
    except Exception as e:
        print(f, "no line:", e)
PY
