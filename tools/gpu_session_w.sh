#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "gemm or decoder or conv1d or linear or full_size" > $O/w_kernels.log 2>&1; grep -E "^E  +(Assertion|assert)|Error" $O/w_kernels.log | head; tail -3 $O/w_kernels.log
timeout 300 python tools/time_gemm.py > $O/w_time_gemm.log 2>&1; grep "^nt" $O/w_time_gemm.log
for w in cfgB cfgD; do
timeout 600 python bench.py --workload $w --no-cpu-baseline --no-micro --parity-workloads $w > $O/w_bench_$w.json 2> $O/w_bench_$w.log
done
python - <<'PY'
import json
for w in ("B", "D"):
    d = json.load(open("gpurun_out/w_bench_cfg%s.json" % w))
    print(w, {k: d.get(k) for k in ("value", "ms_per_step")}, {k: (round(v["ms_per_step"], 3), v["launches_per_step"]) for k, v in d["kernels"].items() if "gemm" in k})
    print("  parity", {k: {kk: vv for kk, vv in v.items() if "rel_err" in kk or "equal" in kk} for k, v in d["parity"].items()})
PY
