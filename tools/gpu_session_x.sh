#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "gemm" > $O/x_kernels.log 2>&1; tail -2 $O/x_kernels.log
timeout 300 python tools/time_gemm.py > $O/x_time_gemm.log 2>&1; grep "^nt" $O/x_time_gemm.log
timeout 600 python bench.py --no-cpu-baseline --no-micro --no-parity > $O/x_bench_cfgB.json 2> $O/x_bench_cfgB.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/x_bench_cfgB.json"))
print("B", {k: d.get(k) for k in ("value", "ms_per_step")}, {k: (round(v["ms_per_step"], 3), v["launches_per_step"]) for k, v in d["kernels"].items() if "gemm" in k})
PY
