#!/bin/bash
# Round-2 GPU session A (run under gpurun, one GPU): validated suite, then every not-yet-validated path in its own
# process (a trap in one must not poison the others), then the default bench and A/B timings.
mkdir -p gpurun_out
O=gpurun_out
UNVAL="tests/test_gpu_kernels.py::test_bilstm_backward_generation_toggle tests/test_gpu_kernels.py::test_gemm3x_umma_is_fp32_class tests/test_gpu_kernels.py::test_conv1d_k4s2_through_the_gemm_kernel tests/test_gpu_kernels.py::test_linear_and_lstm_projection_through_the_gemm_kernel"
DESEL=""
for t in $UNVAL; do DESEL="$DESEL --deselect $t"; done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider $DESEL > $O/a_tests_validated.log 2>&1
echo "validated suite exit $?" >> $O/a_tests_validated.log
for t in $UNVAL; do
  n=$(echo $t | sed 's/.*:://')
  timeout 600 python -m pytest "$t" -m gpu -q -p no:cacheprovider > $O/a_unval_$n.log 2>&1
  echo "exit $?" >> $O/a_unval_$n.log
done
B=64 H=512 T=300 REPS=2 timeout 300 python tools/time_lstm.py > $O/a_time_lstm_512.log 2>&1
B=32 H=640 T=299 I=640 REPS=2 timeout 300 python tools/time_lstm.py > $O/a_time_lstm_640.log 2>&1
timeout 900 python bench.py > $O/a_bench_cfgB.json 2> $O/a_bench_cfgB.log
B200ASR_GEMM=umma timeout 600 python bench.py --no-cpu-baseline --no-micro --parity-workloads cfgB > $O/a_bench_cfgB_umma.json 2> $O/a_bench_cfgB_umma.log
tail -3 $O/a_tests_validated.log
for t in $UNVAL; do n=$(echo $t | sed 's/.*:://'); tail -2 $O/a_unval_$n.log; done
cat $O/a_time_lstm_512.log $O/a_time_lstm_640.log
cat $O/a_bench_cfgB.json $O/a_bench_cfgB_umma.json
