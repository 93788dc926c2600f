#!/bin/bash
# 2-GPU: the torchrun line the driver runs at N=2 (+ reference arm), DP parity test, attention tests after the template split
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_zz_dp.py -m gpu -q -p no:cacheprovider -k "attention or data_parallel or dp" > $O/o_tests.log 2>&1; tail -3 $O/o_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/o_bench_2gpu.json 2> $O/o_bench_2gpu.log
tail -2 $O/o_bench_2gpu.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/o_bench_2gpu.json"))
print({k:d.get(k) for k in ("value","n_gpus","ms_per_step","cuda_graph")}, d.get("e2e"))
print("also", d.get("also"))
print("parity", {k: {kk: vv for kk, vv in v.items() if "rel_err" in kk or "equal" in kk} for k, v in (d.get("parity") or {}).items()})
PY
timeout 600 python bench.py --workload cfgD --no-cpu-baseline --no-micro --no-parity > $O/o_bench_cfgD.json 2> $O/o_bench_cfgD.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/o_bench_cfgD.json"))
print("D", {k: d.get(k) for k in ("value", "ms_per_step")}, {k: (round(v["ms_per_step"], 3), v["launches_per_step"]) for k, v in d["kernels"].items() if "locattn" in k})
PY
