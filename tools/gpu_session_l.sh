#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "bilstm" > $O/l_bilstm.log 2>&1; grep -E "^E  +(Assertion|assert)|Error" $O/l_bilstm.log | head; tail -3 $O/l_bilstm.log
ONLY_MODES=0,512 B=32 H=640 T=299 I=640 REPS=2 timeout 300 python tools/time_lstm.py > $O/l_time_lstm_640.log 2>&1; cat $O/l_time_lstm_640.log
ONLY_MODES=0,512 B=64 H=512 T=300 REPS=2 timeout 300 python tools/time_lstm.py > $O/l_time_lstm_512.log 2>&1; cat $O/l_time_lstm_512.log
timeout 600 python bench.py --workload cfgD --no-cpu-baseline --no-micro > $O/l_bench_cfgD.json 2> $O/l_bench_cfgD.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/l_bench_cfgD.json"))
print({k: d.get(k) for k in ("value", "ms_per_step", "own_kernel_ms_per_step", "library_ms_per_step")})
print("  ", {k: (round(v["ms_per_step"], 3), v["launches_per_step"]) for k, v in d["kernels"].items()})
if d.get("parity"): print("  parity", {k: {kk: vv for kk, vv in v.items() if "rel_err" in kk or "equal" in kk} for k, v in d["parity"].items()})
PY
