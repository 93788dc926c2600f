#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
MODE=0 timeout 120 python tools/trace_lstm.py > $O/h_trace_fwd.log 2>&1; cat $O/h_trace_fwd.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "ctc" > $O/h_ctc.log 2>&1; tail -6 $O/h_ctc.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_zz_plumbing.py tests/test_gpu_zz_decode.py -m gpu -q -p no:cacheprovider > $O/h_model.log 2>&1; grep -E "^E  +(Assertion|assert)|Error" $O/h_model.log | head -10; tail -6 $O/h_model.log
timeout 900 python bench.py --no-cpu-baseline --parity-workloads cfgB,cfgC > $O/h_bench_cfgB.json 2> $O/h_bench_cfgB.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/h_bench_cfgB.json"))
print({k:d[k] for k in ("value","ms_per_step","own_kernel_ms_per_step","library_ms_per_step")})
print({k:(round(v["ms_per_step"],3),v["launches_per_step"]) for k,v in d["kernels"].items()})
for w,p in d["parity"].items(): print(w, {kk: vv for kk, vv in p.items() if "rel_err" in kk or "equal" in kk})
print({k:(round(v["ms_per_1000_utt"],3), round(v["frac_hbm"],4)) for k,v in d["micro"].items()})
PY
