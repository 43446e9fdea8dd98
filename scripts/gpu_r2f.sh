#!/bin/bash
# round-2 visit F (8 GPUs): BASELINE configs 4 and 5 at size, and the config-3 bench line with its in-bench parity check
TAG=${1:-r2f}; NG=${2:-8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt; free -g | head -2 > $OUT/host_mem.txt; nproc >> $OUT/host_mem.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29515 bench.py --gpus $NG --config 5 --steps 10 --warmup 5 --no-cpu-baseline --no-e2e > $OUT/config5_$NG.json 2> $OUT/config5_$NG.err; echo "config 5 exit $?"; grep -v "^\*\|OMP_NUM\|^$" $OUT/config5_$NG.err | tail -5 | cut -c1-300
timeout 900 $TR --master-port 29514 bench.py --gpus $NG --config 4 --steps 10 --warmup 5 --no-cpu-baseline --no-parity > $OUT/config4_$NG.json 2> $OUT/config4_$NG.err; echo "config 4 exit $?"; grep -v "^\*\|OMP_NUM\|^$" $OUT/config4_$NG.err | tail -5 | cut -c1-300
timeout 900 $TR --master-port 29513 bench.py --gpus $NG --steps 20 --warmup 5 --no-cpu-baseline --no-parity > $OUT/bench_$NG.json 2> $OUT/bench_$NG.err; echo "bench exit $?"; grep -v "^\*\|OMP_NUM\|^$" $OUT/bench_$NG.err | tail -5 | cut -c1-300
python - <<PY
import json
for f in ("config5_$NG","config4_$NG","bench_$NG"):
    try:
        d=json.loads(open("$OUT/"+f+".json").read().strip().splitlines()[-1])
        print(f, "ms_per_step", round(d["ms_per_step"],3), "value %.3e" % d["value"], "e2e", d["e2e"]["value"], "mem", d["config"].get("device_memory_used_gb_rank0"), "parity", d.get("parity_check"), "inv", d["invariants"], d["config"].get("ncdm_substeps"), d["config"].get("hij_spectrum_call_ms"))
        for k,v in sorted(d["kernels"].items(), key=lambda kv:-kv[1]["ms_per_step"])[:8]:
            print(f"   {k:32s} {v['ms_per_step']:8.3f} ms/step")
        print("   nvlink", d.get("nvlink"))
    except Exception as e:
        print(f, "no line:", e)
PY
