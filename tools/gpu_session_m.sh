#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "attention" > $O/m_attn.log 2>&1; grep -E "^E  +(Assertion|assert)|Error" $O/m_attn.log | head; tail -3 $O/m_attn.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_zz_decode.py -m gpu -q -p no:cacheprovider > $O/m_model.log 2>&1; grep -E "^E  +(Assertion|assert)|Error" $O/m_model.log | head; tail -3 $O/m_model.log
for w in cfgC cfgD; do
timeout 600 python bench.py --workload $w --no-cpu-baseline --no-micro --no-parity > $O/m_bench_$w.json 2> $O/m_bench_$w.log
done
python - <<'PY'
import json
for w in ("C", "D"):
    try: d = json.load(open("gpurun_out/m_bench_cfg%s.json" % w))
    except Exception as e: print(w, "failed", e); continue
    print(w, {k: d.get(k) for k in ("value", "ms_per_step", "own_kernel_ms_per_step", "library_ms_per_step")})
    print("  ", {k: (round(v["ms_per_step"], 3), v["launches_per_step"]) for k, v in d["kernels"].items()})
PY
