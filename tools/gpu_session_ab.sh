#!/bin/bash
# GEMM forms with both operands pre-split (no in-kernel split pass): numerics, isolated timing, step A/B
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm3x or bilstm or linear or conv1d" -p no:cacheprovider > $O/ab_pytest.log 2>&1; tail -5 $O/ab_pytest.log
QUICK=1 timeout 300 python tools/time_gemm.py > $O/ab_time_gemm.log 2>&1; cat $O/ab_time_gemm.log
for ps in 1 0; do
B200ASR_GEMM_PRESPLIT=$ps timeout 600 python bench.py --no-cpu-baseline --no-micro --no-parity > $O/ab_bench_ps$ps.json 2> $O/ab_bench_ps$ps.log
python - <<PY
import json
d=json.load(open("gpurun_out/ab_bench_ps$ps.json"))
print("presplit=$ps", {k:d[k] for k in ("value","ms_per_step","own_kernel_ms_per_step","library_ms_per_step","gpu_launches")}, d["e2e"]["value"])
print({k:(round(v["ms_per_step"],3),v["launches_per_step"]) for k,v in d["kernels"].items()})
PY
done
