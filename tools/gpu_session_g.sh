#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 120 tools/micro/xfer_probe > $O/g_xfer_probe.log 2>&1; cat $O/g_xfer_probe.log
for t in test_gemm3x_umma_is_fp32_class test_gemm3x_nn_is_fp32_class test_gemm3x_nt_is_fp32_class; do
  timeout 600 python -m pytest "tests/test_gpu_kernels.py::$t" -m gpu -q -p no:cacheprovider > $O/g_$t.log 2>&1
  echo "== $t"; grep -E "^E  +(Assertion|assert)|Error|passed|failed" $O/g_$t.log | head -8
done
timeout 600 python tools/time_gemm.py > $O/g_time_gemm.log 2>&1; cat $O/g_time_gemm.log
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -p no:cacheprovider > $O/g_tests_model.log 2>&1; tail -4 $O/g_tests_model.log
timeout 600 python bench.py --no-cpu-baseline --no-micro --parity-workloads cfgB > $O/g_bench_cfgB.json 2> $O/g_bench_cfgB.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/g_bench_cfgB.json"))
print({k:d[k] for k in ("value","ms_per_step","own_kernel_ms_per_step","library_ms_per_step")})
print({k:(round(v["ms_per_step"],3),v["launches_per_step"]) for k,v in d["kernels"].items()})
print({kk: vv for kk, vv in d["parity"]["cfgB"].items() if "rel_err" in kk or "equal" in kk})
PY
