#!/bin/bash
# Round-2 profiling recipe (run under gpurun, one GPU): launch list of an eager cfg-B step + `ncu --set full` captures
# of the shipped kernels.  Summaries: python tools/ncu_summary.py gpurun_out/r02_*.ncu-rep  (works without a GPU).
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-e2e --no-parity --no-micro \
    > gpurun_out/r02_launch_bench.log 2>&1
T=128 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'bilstm_(fwd|bwd)_umma_kernel|gemm3x_kernel' -s 16 -c 8 -f -o gpurun_out/r02_lstm_gemm \
    python tools/profile_lstm.py > gpurun_out/r02_ncu_lstm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'ctc_grad|ctc_alpha|fbank_kernel|delta_|log_softmax_fwd|locattn|attn_dvalue|ce_fwd' -s 11 -c 12 -f -o gpurun_out/r02_misc \
    python tools/profile_kernels.py > gpurun_out/r02_ncu_misc.log 2>&1
ls -la gpurun_out/*.ncu-rep
