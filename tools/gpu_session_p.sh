#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -k "exchange_protocol_toggle or attention" -m gpu -q -p no:cacheprovider > $O/p_poll.log 2>&1; grep -E "^E  +(Assertion|assert)|Error" $O/p_poll.log | head -5; tail -3 $O/p_poll.log
sed -i 's/(3072, "tcgen05+exchange-toggle")/(1024, "tcgen05+fwd-exchange-toggle")/' tools/time_lstm.py
ONLY_MODES=0,1024 B=64 H=512 T=300 REPS=3 timeout 300 python tools/time_lstm.py > $O/p_time_lstm_512.log 2>&1; cat $O/p_time_lstm_512.log
ONLY_MODES=0,1024 B=32 H=640 T=299 I=640 REPS=2 timeout 300 python tools/time_lstm.py > $O/p_time_lstm_640.log 2>&1; cat $O/p_time_lstm_640.log
MODE=1024 timeout 120 python tools/trace_lstm.py > $O/p_trace_fwd_poll.log 2>&1; cat $O/p_trace_fwd_poll.log
timeout 600 python bench.py --workload cfgD --no-cpu-baseline --no-micro --no-parity > $O/p_bench_cfgD.json 2> $O/p_bench_cfgD.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/p_bench_cfgD.json"))
print("D", {k: d.get(k) for k in ("value", "ms_per_step")}, {k: (round(v["ms_per_step"], 3), v["launches_per_step"]) for k, v in d["kernels"].items() if "locattn" in k})
PY
