#!/bin/bash
# final 2-GPU check: DP parity test + the driver's torchrun lines (own arm and reference arm)
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_zz_dp.py -m gpu -q -p no:cacheprovider > $O/u_dp.log 2>&1; tail -3 $O/u_dp.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/u_bench_2gpu.json 2> $O/u_bench_2gpu.log
echo "bench rc $?"; tail -2 $O/u_bench_2gpu.log | cut -c1-300
python - <<'PY'
import json
lines=[l for l in open("gpurun_out/u_bench_2gpu.json").read().splitlines() if l.strip().startswith("{")]
d=json.loads(lines[-1])
print({k:d.get(k) for k in ("value","n_gpus","ms_per_step","cuda_graph")}, d.get("e2e"))
print("also", d.get("also"))
print("parity", {k: {kk: vv for kk, vv in v.items() if "rel_err" in kk or "equal" in kk} for k, v in (d.get("parity") or {}).items()})
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/u_ref_2gpu.json 2> $O/u_ref_2gpu.log; echo "ref rc $?"; cut -c1-200 $O/u_ref_2gpu.json
