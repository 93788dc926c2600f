#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/debug_vgg.py > $O/d_debug_vgg.log 2>&1; cat $O/d_debug_vgg.log
timeout 600 python -m pytest "tests/test_gpu_kernels.py::test_bilstm_exchange_protocol_toggle" -m gpu -q -p no:cacheprovider > $O/d_poll.log 2>&1; tail -3 $O/d_poll.log
ONLY_MODES=0,1024 B=64 H=512 T=300 REPS=2 timeout 300 python tools/time_lstm.py > $O/d_time_lstm_512.log 2>&1; cat $O/d_time_lstm_512.log
