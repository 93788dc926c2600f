#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "attention" > $O/q_attn.log 2>&1; tail -2 $O/q_attn.log
for w in cfgC cfgD; do
timeout 600 python bench.py --workload $w --no-cpu-baseline --no-micro --no-parity > $O/q_bench_$w.json 2> $O/q_bench_$w.log
done
python - <<'PY'
import json
for w in ("C", "D"):
    d = json.load(open("gpurun_out/q_bench_cfg%s.json" % w))
    print(w, {k: d.get(k) for k in ("value", "ms_per_step")}, {k: (round(v["ms_per_step"], 3), v["launches_per_step"]) for k, v in d["kernels"].items() if "locattn" in k})
PY
QUICK=1 timeout 300 python tools/time_gemm.py > $O/q_time_gemm_base.log 2>&1; cat $O/q_time_gemm_base.log
cp end-to-end-asr-pytorch_b200/libb200asr.so /tmp/libb200asr_base.so
B200ASR_NVCC_EXTRA="-DB200ASR_GEMM_HH_FIRST=1" timeout 600 python -c "import importlib; b = importlib.import_module('end-to-end-asr-pytorch_b200._build'); print(b.build(force=True))" > $O/q_rebuild.log 2>&1; tail -1 $O/q_rebuild.log
QUICK=1 timeout 300 python tools/time_gemm.py > $O/q_time_gemm_hh.log 2>&1; cat $O/q_time_gemm_hh.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "gemm3x" > $O/q_gemm_hh.log 2>&1; tail -2 $O/q_gemm_hh.log
timeout 600 python bench.py --no-cpu-baseline --no-micro --no-parity > $O/q_bench_cfgB_hh.json 2> $O/q_bench_cfgB_hh.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/q_bench_cfgB_hh.json"))
print("B hh-first", {k: d.get(k) for k in ("value", "ms_per_step")}, {k: (round(v["ms_per_step"], 3), v["launches_per_step"]) for k, v in d["kernels"].items() if "gemm" in k})
PY
cp /tmp/libb200asr_base.so end-to-end-asr-pytorch_b200/libb200asr.so
timeout 600 python bench.py --no-cpu-baseline --no-micro --no-parity > $O/q_bench_cfgB_base.json 2> $O/q_bench_cfgB_base.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/q_bench_cfgB_base.json"))
print("B base", {k: d.get(k) for k in ("value", "ms_per_step")}, {k: (round(v["ms_per_step"], 3), v["launches_per_step"]) for k, v in d["kernels"].items() if "gemm" in k})
PY
