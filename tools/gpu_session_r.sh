#!/bin/bash
# "final state" session: full suite, the driver's default bench line, profiles
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r_tests.log 2>&1; grep -E "^E  +(Assertion|assert)|Error" $O/r_tests.log | head; tail -4 $O/r_tests.log
timeout 900 python bench.py > $O/r_bench_cfgB.json 2> $O/r_bench_cfgB.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/r_bench_cfgB.json"))
print({k:d[k] for k in ("value","ms_per_step","own_kernel_ms_per_step","library_ms_per_step","gpu_launches")}, d["e2e"], d["clocks"])
print({k:(round(v["ms_per_step"],3),v["launches_per_step"]) for k,v in d["kernels"].items()})
for w,p in d["parity"].items(): print(w, {kk: vv for kk, vv in p.items() if "rel_err" in kk or "equal" in kk})
print({k:(round(v["ms_per_1000_utt"],3), round(v["frac_hbm"],4)) for k,v in d["micro"].items()})
print(d["roofline"]); print(d["cpu_baseline"])
PY
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r_smoke.log 2>&1; tail -2 $O/r_smoke.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r_bench_ref.json 2> $O/r_bench_ref.log; cut -c1-300 $O/r_bench_ref.json
bash tools/profile_r02.sh
