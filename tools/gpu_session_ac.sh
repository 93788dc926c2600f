#!/bin/bash
# 2 GPUs: the data-parallel step inside the CUDA graph (NCCL all-reduces captured) vs eager; teardown must not hang
mkdir -p gpurun_out
O=gpurun_out
for mode in "--graph-dp" ""; do
tag=$( [ -n "$mode" ] && echo graph || echo eager )
t0=$SECONDS; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-micro --no-parity --no-also $mode > $O/ac_bench_2gpu_$tag.json 2> $O/ac_bench_2gpu_$tag.log
echo "$tag rc $? wall $((SECONDS-t0)) s"; grep -E "CUDA graph|wall|Error|error" $O/ac_bench_2gpu_$tag.log | cut -c1-300 | tail -5
python - <<PY
import json
lines=[l for l in open("gpurun_out/ac_bench_2gpu_$tag.json").read().splitlines() if l.strip().startswith("{")]
if lines:
    d=json.loads(lines[-1]); print("$tag", {k:d.get(k) for k in ("value","n_gpus","ms_per_step","cuda_graph","loss")}, d.get("e2e"))
PY
done
