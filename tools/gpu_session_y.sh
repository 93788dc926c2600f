#!/bin/bash
mkdir -p gpurun_out
MODE=0 timeout 120 python tools/trace_lstm_bwd.py > gpurun_out/y_trace_bwd_poll.log 2>&1; cat gpurun_out/y_trace_bwd_poll.log
MODE=2048 timeout 120 python tools/trace_lstm_bwd.py > gpurun_out/y_trace_bwd_flag.log 2>&1; cat gpurun_out/y_trace_bwd_flag.log
