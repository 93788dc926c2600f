#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -p no:cacheprovider -k "fbank or two_train_steps or solver or graph" > $O/t_front.log 2>&1; grep -E "^E  +(Assertion|assert)|Error" $O/t_front.log | head; tail -3 $O/t_front.log
timeout 600 python tools/microbench.py > $O/t_micro.jsonl 2> $O/t_micro.log; tail -c 1500 $O/t_micro.jsonl
