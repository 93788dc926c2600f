#!/bin/bash
# GPU session C: MN-major GEMM forms with the 32-byte-base swizzle; polling exchange of the LSTM forward kernel.
mkdir -p gpurun_out
O=gpurun_out
for t in test_gemm3x_nn_is_fp32_class test_gemm3x_nt_is_fp32_class test_conv1d_k4s2_through_the_gemm_kernel test_linear_and_lstm_projection_through_the_gemm_kernel test_bilstm_exchange_protocol_toggle; do
  timeout 600 python -m pytest "tests/test_gpu_kernels.py::$t" -m gpu -q -p no:cacheprovider > $O/c_$t.log 2>&1
  echo "exit $?" >> $O/c_$t.log
  echo "== $t"; grep -E "^E  +(Assertion|assert)|Error|passed|failed|exit" $O/c_$t.log | head -20
done
B=64 H=512 T=300 REPS=2 timeout 300 python tools/time_lstm.py > $O/c_time_lstm_512.log 2>&1; cat $O/c_time_lstm_512.log
B=32 H=640 T=299 I=640 REPS=2 timeout 300 python tools/time_lstm.py > $O/c_time_lstm_640.log 2>&1; cat $O/c_time_lstm_640.log
MODE=0 timeout 120 python tools/trace_lstm.py > $O/c_trace_fwd_flag.log 2>&1; cat $O/c_trace_fwd_flag.log
MODE=1024 timeout 120 python tools/trace_lstm.py > $O/c_trace_fwd_poll.log 2>&1; cat $O/c_trace_fwd_poll.log
B200ASR_GEMM=umma timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "not full_size" > $O/c_tests_umma.log 2>&1; tail -5 $O/c_tests_umma.log
B200ASR_GEMM=umma timeout 600 python bench.py --no-cpu-baseline --no-micro --parity-workloads cfgB,cfgD > $O/c_bench_cfgB_umma.json 2> $O/c_bench_cfgB_umma.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/c_bench_cfgB_umma.json"))
print({k:d[k] for k in ("value","ms_per_step","own_kernel_ms_per_step","library_ms_per_step")})
print({k:(round(v["ms_per_step"],3),v["launches_per_step"]) for k,v in d["kernels"].items()})
print(d["parity"])
PY
