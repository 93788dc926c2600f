#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest "tests/test_gpu_kernels.py::test_bilstm_exchange_protocol_toggle" -m gpu -q -p no:cacheprovider > $O/p_poll.log 2>&1; grep -E "^E  +(Assertion|assert)|Error" $O/p_poll.log | head -5; tail -3 $O/p_poll.log
sed -i 's/(3072, "tcgen05+exchange-toggle")/(1024, "tcgen05+fwd-exchange-toggle")/' tools/time_lstm.py
ONLY_MODES=0,1024 B=64 H=512 T=300 REPS=3 timeout 300 python tools/time_lstm.py > $O/p_time_lstm_512.log 2>&1; cat $O/p_time_lstm_512.log
ONLY_MODES=0,1024 B=32 H=640 T=299 I=640 REPS=2 timeout 300 python tools/time_lstm.py > $O/p_time_lstm_640.log 2>&1; cat $O/p_time_lstm_640.log
MODE=1024 timeout 120 python tools/trace_lstm.py > $O/p_trace_fwd_poll.log 2>&1; cat $O/p_trace_fwd_poll.log
