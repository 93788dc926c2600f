#!/bin/bash
# 2 GPUs: cap NCCL's CTA count so that its kernels fit next to the 128 co-resident CTAs of the LSTM kernels (148 SMs)
mkdir -p gpurun_out
O=gpurun_out
for ctas in 16 8 default; do
if [ "$ctas" = default ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$ctas; fi
t0=$SECONDS; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-micro --no-parity --no-also --no-e2e > $O/ad_bench_2gpu_$ctas.json 2> $O/ad_bench_2gpu_$ctas.log
echo "ctas=$ctas rc $? wall $((SECONDS-t0)) s"
python - <<PY
import json
lines=[l for l in open("gpurun_out/ad_bench_2gpu_$ctas.json").read().splitlines() if l.strip().startswith("{")]
if lines:
    d=json.loads(lines[-1]); print("ctas=$ctas", {k:d.get(k) for k in ("value","n_gpus","ms_per_step","cuda_graph")})
PY
done
