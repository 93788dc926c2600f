#!/bin/bash
# nt GEMM with only the wide operand (x / layer output) pre-split: numerics, isolated timing, step A/B
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm3x_nt" -p no:cacheprovider > $O/ae_pytest.log 2>&1; tail -3 $O/ae_pytest.log
QUICK=1 timeout 300 python tools/time_gemm.py > $O/ae_time_gemm.log 2>&1; grep "nt " $O/ae_time_gemm.log
for ps in x 0; do
B200ASR_GEMM_PRESPLIT=$ps timeout 600 python bench.py --no-cpu-baseline --no-micro --no-parity --no-e2e > $O/ae_bench_ps$ps.json 2> $O/ae_bench_ps$ps.log
python - <<PY
import json
d=json.load(open("gpurun_out/ae_bench_ps$ps.json"))
print("presplit=$ps", {k:d[k] for k in ("value","ms_per_step","own_kernel_ms_per_step","library_ms_per_step","gpu_launches")})
print({k:(round(v["ms_per_step"],3),v["launches_per_step"]) for k,v in d["kernels"].items() if "gemm" in k or "resid" in k})
PY
done
