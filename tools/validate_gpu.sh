#!/bin/bash
# final-state session: full suite, the driver's default bench line, smoke, reference arm, profiles
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/v_tests.log 2>&1; grep -E "^E  +(Assertion|assert)|Error" $O/v_tests.log | head; tail -4 $O/v_tests.log
timeout 900 python bench.py > $O/v_bench_cfgB.json 2> $O/v_bench_cfgB.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/v_bench_cfgB.json"))
print({k:d[k] for k in ("value","ms_per_step","own_kernel_ms_per_step","library_ms_per_step","gpu_launches")}, d["e2e"], d["clocks"])
print({k:(round(v["ms_per_step"],3),v["launches_per_step"]) for k,v in d["kernels"].items()})
for w,p in d["parity"].items(): print(w, {kk: vv for kk, vv in p.items() if "rel_err" in kk or "equal" in kk})
print({k:(round(v["ms_per_1000_utt"],3), round(v["frac_hbm"],4)) for k,v in d["micro"].items()})
print(d["roofline"]); print(d["cpu_baseline"])
PY
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/v_smoke.log 2>&1; tail -1 $O/v_smoke.log
if [ -n "$QUICK" ]; then exit 0; fi        # QUICK=1: suite + the driver's default bench line + smoke only
for w in cfgC cfgD; do
timeout 600 python bench.py --workload $w --no-cpu-baseline --no-micro --parity-workloads $w > $O/v_bench_$w.json 2> $O/v_bench_$w.log
done
python - <<'PY'
import json
for w in ("C", "D"):
    d = json.load(open("gpurun_out/v_bench_cfg%s.json" % w))
    print(w, {k: d.get(k) for k in ("value", "ms_per_step", "own_kernel_ms_per_step", "library_ms_per_step")}, d["e2e"]["value"])
    print("  ", {k: (round(v["ms_per_step"], 3), v["launches_per_step"]) for k, v in d["kernels"].items()})
PY
bash tools/profile_r02.sh
