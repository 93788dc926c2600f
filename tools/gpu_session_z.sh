#!/bin/bash
# 128-column n-blocks in the tcgen05 backward: numerics, per-step time (H = 512 / 640), timeline
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "lstm" > gpurun_out/z_pytest_lstm.log 2>&1; tail -5 gpurun_out/z_pytest_lstm.log
ONLY_MODES=0,3 timeout 200 python tools/time_lstm.py > gpurun_out/z_time_lstm_h512.log 2>&1; cat gpurun_out/z_time_lstm_h512.log
H=640 B=32 I=1280 ONLY_MODES=0,3 timeout 200 python tools/time_lstm.py > gpurun_out/z_time_lstm_h640.log 2>&1; cat gpurun_out/z_time_lstm_h640.log
MODE=0 timeout 120 python tools/trace_lstm_bwd.py > gpurun_out/z_trace_bwd_poll.log 2>&1; cat gpurun_out/z_trace_bwd_poll.log
timeout 200 python tools/check_bwd_protocols.py > gpurun_out/z_check_bwd.log 2>&1; tail -30 gpurun_out/z_check_bwd.log
