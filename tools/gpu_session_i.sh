#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/i_tests.log 2>&1; grep -E "^E  +(Assertion|assert)|Error" $O/i_tests.log | head -10; tail -5 $O/i_tests.log
timeout 900 python bench.py --no-cpu-baseline --no-parity > $O/i_bench_cfgB.json 2> $O/i_bench_cfgB.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/i_bench_cfgB.json"))
print({k:d[k] for k in ("value","ms_per_step","own_kernel_ms_per_step","library_ms_per_step")})
print({k:(round(v["ms_per_1000_utt"],3), round(v["frac_hbm"],4)) for k,v in d["micro"].items()})
PY
