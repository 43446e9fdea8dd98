#!/bin/bash
# round-2 visit P (8 GPUs): config 3 with the final kernels (driver-shaped run, parity check inside), fft_overlap ablation
TAG=${1:-r2p}; NG=${2:-8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29512 bench.py --gpus $NG --steps 20 --warmup 5 --no-cpu-baseline --ablate fft_overlap=2:1:2:1 > $OUT/bench_$NG.json 2> $OUT/bench_$NG.err; echo "bench exit $?"; tail -c 400 $OUT/bench_$NG.err
grep -h ablate $OUT/bench_$NG.err | cut -c1-400
python - <<PY
import json
d=json.loads(open("$OUT/bench_$NG.json").read().strip().splitlines()[-1])
print("ms_per_step", round(d["ms_per_step"],3), "value %.3e" % d["value"], "e2e", d["e2e"]["value"], "parity", d.get("parity_check",{}).get("failed"), d["invariants"])
for k,v in sorted(d["kernels"].items(), key=lambda kv:-kv[1]["ms_per_step"])[:14]:
    print(f"   {k:32s} {v['ms_per_step']:8.3f} ms/step")
print(d["config"].get("e2e_parts_rank0"), d["config"].get("host_binding_rank0")); print(d.get("nvlink"))
PY
