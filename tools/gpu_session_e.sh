#!/bin/bash
# GPU session E: full suite with the round-2 defaults, bench records (cfg B full, cfg C, cfg D), ncu profiles.
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/debug_vgg2.py > $O/e_debug_vgg2.log 2>&1; cat $O/e_debug_vgg2.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/e_tests.log 2>&1; tail -15 $O/e_tests.log
timeout 900 python bench.py > $O/e_bench_cfgB.json 2> $O/e_bench_cfgB.log
timeout 600 python bench.py --workload cfgC --no-cpu-baseline --no-micro --no-parity > $O/e_bench_cfgC.json 2> $O/e_bench_cfgC.log
timeout 600 python bench.py --workload cfgD --no-cpu-baseline --no-micro --no-parity > $O/e_bench_cfgD.json 2> $O/e_bench_cfgD.log
python - <<'PY'
import json
for w in ("B", "C", "D"):
    try:
        d = json.load(open("gpurun_out/e_bench_cfg%s.json" % w))
    except Exception as e:
        print(w, "failed", e); continue
    print(w, {k: d.get(k) for k in ("value", "ms_per_step", "own_kernel_ms_per_step", "library_ms_per_step")}, d.get("e2e"))
    print("  ", {k: (round(v["ms_per_step"], 3), v["launches_per_step"]) for k, v in d["kernels"].items()})
    print("  roofline", d.get("roofline"))
    if d.get("parity"): print("  parity", {k: {kk: vv for kk, vv in v.items() if "rel_err" in kk or "equal" in kk} for k, v in d["parity"].items()})
PY
bash tools/profile_r02.sh
