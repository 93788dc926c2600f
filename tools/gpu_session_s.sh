#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "gemm or decoder or conv1d or linear or bilstm_fwd_bwd" > $O/s_kernels.log 2>&1; grep -E "^E  +(Assertion|assert)|Error" $O/s_kernels.log | head; tail -3 $O/s_kernels.log
QUICK=1 timeout 300 python tools/time_gemm.py > $O/s_time_gemm.log 2>&1; cat $O/s_time_gemm.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_zz_plumbing.py -m gpu -q -p no:cacheprovider > $O/s_model.log 2>&1; grep -E "^E  +(Assertion|assert)|Error" $O/s_model.log | head; tail -3 $O/s_model.log
for w in cfgB cfgC cfgD; do
timeout 600 python bench.py --workload $w --no-cpu-baseline --no-micro --parity-workloads $w > $O/s_bench_$w.json 2> $O/s_bench_$w.log
done
python - <<'PY'
import json
for w in ("B", "C", "D"):
    try: d = json.load(open("gpurun_out/s_bench_cfg%s.json" % w))
    except Exception as e: print(w, "failed", e); continue
    print(w, {k: d.get(k) for k in ("value", "ms_per_step", "own_kernel_ms_per_step", "library_ms_per_step")})
    print("  ", {k: (round(v["ms_per_step"], 3), v["launches_per_step"]) for k, v in d["kernels"].items() if "gemm" in k or "resid" in k})
    if d.get("parity"): print("  parity", {k: {kk: vv for kk, vv in v.items() if "rel_err" in kk or "equal" in kk} for k, v in d["parity"].items()})
PY
