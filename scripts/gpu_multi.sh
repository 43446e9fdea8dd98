#!/bin/bash
# multi-GPU visit (gpurun --gpus N): slab-decomposition parity + scaling bench lines.  usage: bash scripts/gpu_multi.sh <tag> <ngpus> [bench N list]
TAG=${1:-m}; NG=${2:-2}; shift; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_worker.py > $OUT/multigpu_parity_$NG.log 2>&1; echo "parity exit $?"; grep -E "BAD|MULTIGPU|Error|error" $OUT/multigpu_parity_$NG.log | head -20
for n in ${@:-$NG}; do
  if [ "$n" = "1" ]; then timeout 900 python bench.py --no-cpu-baseline > $OUT/bench_$n.json 2> $OUT/bench_$n.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --no-cpu-baseline $GEVB_BENCH_EXTRA > $OUT/bench_$n.json 2> $OUT/bench_$n.err; fi
  echo "bench $n exit $?"; tail -c 300 $OUT/bench_$n.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$n.json").read().strip().splitlines()[-1])
    print("gpus $n ms_per_step", d["ms_per_step"], "value %.3e" % d["value"], "e2e", d["e2e"]["value"])
    for k,v in sorted(d["kernels"].items(), key=lambda kv:-kv[1]["ms_per_step"])[:9]:
        print(f"   {k:32s} {v['ms_per_step']:8.3f} ms/step")
except Exception as e:
    print("no bench line:", e)
PY
done
