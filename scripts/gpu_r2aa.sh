#!/bin/bash
# 4 GPUs: slab parity worker + config-3 bench line with the in-bench parity check (the one world size not visited yet this round)
TAG=${1:-r2aa}; NG=${2:-4}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/multigpu_worker.py > $OUT/multigpu_parity_$NG.log 2>&1; echo "parity exit $?"; grep -E "BAD|MULTIGPU|Error|error" $OUT/multigpu_parity_$NG.log | head -10
timeout 600 $TR --master-port 29512 bench.py --gpus $NG --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_$NG.json 2> $OUT/bench_$NG.err; echo "bench exit $?"; tail -c 300 $OUT/bench_$NG.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_$NG.json").read().strip().splitlines()[-1])
print("ms_per_step", round(d["ms_per_step"],3), "value %.3e" % d["value"], "e2e", d["e2e"]["value"], "parity failed", d["parity_check"]["failed"], d["invariants"])
for k,v in sorted(d["kernels"].items(), key=lambda kv:-kv[1]["ms_per_step"])[:10]:
    print(f"   {k:32s} {v['ms_per_step']:8.3f} ms/step")
PY
