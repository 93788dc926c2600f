#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/n_tests.log 2>&1; grep -E "^E  +(Assertion|assert)|Error" $O/n_tests.log | head; tail -4 $O/n_tests.log
for w in cfgC cfgD; do
timeout 600 python bench.py --workload $w --no-cpu-baseline --no-micro > $O/n_bench_$w.json 2> $O/n_bench_$w.log
done
python - <<'PY'
import json
for w in ("C", "D"):
    try: d = json.load(open("gpurun_out/n_bench_cfg%s.json" % w))
    except Exception as e: print(w, "failed", e); continue
    print(w, {k: d.get(k) for k in ("value", "ms_per_step", "own_kernel_ms_per_step", "library_ms_per_step")})
    print("  ", {k: (round(v["ms_per_step"], 3), v["launches_per_step"]) for k, v in d["kernels"].items()})
    if d.get("parity"): print("  parity", {k: {kk: vv for kk, vv in v.items() if "rel_err" in kk or "equal" in kk} for k, v in d["parity"].items() if k == "cfg" + w})
PY
