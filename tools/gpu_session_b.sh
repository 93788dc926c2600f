#!/bin/bash
# GPU session B: own GEMM forms (tn with chunked accumulation, nn, nt) - accuracy, then speed, then the model-level tests.
mkdir -p gpurun_out
O=gpurun_out
for t in test_gemm3x_umma_is_fp32_class test_gemm3x_nn_is_fp32_class test_gemm3x_nt_is_fp32_class test_conv1d_k4s2_through_the_gemm_kernel test_linear_and_lstm_projection_through_the_gemm_kernel; do
  timeout 600 python -m pytest "tests/test_gpu_kernels.py::$t" -m gpu -q -p no:cacheprovider > $O/b_$t.log 2>&1
  echo "exit $?" >> $O/b_$t.log
  echo "== $t"; grep -E "^E  +(Assertion|assert)|passed|failed|exit" $O/b_$t.log | head -20
done
timeout 600 python tools/time_gemm.py > $O/b_time_gemm.log 2>&1; cat $O/b_time_gemm.log
B200ASR_GEMM=umma timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "not full_size" > $O/b_tests_umma.log 2>&1; tail -5 $O/b_tests_umma.log
B200ASR_GEMM=umma timeout 600 python bench.py --no-cpu-baseline --no-micro --parity-workloads cfgB > $O/b_bench_cfgB_umma.json 2> $O/b_bench_cfgB_umma.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/b_bench_cfgB_umma.json"))
print({k:d[k] for k in ("value","ms_per_step","own_kernel_ms_per_step","library_ms_per_step")})
print({k:(round(v["ms_per_step"],3),v["launches_per_step"]) for k,v in d["kernels"].items()})
print(d["parity"])
PY
