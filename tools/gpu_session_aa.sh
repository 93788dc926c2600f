#!/bin/bash
# CTC lattice prefetch ring (raw loads) + gradient stream with 4 loads in flight: numerics, micro
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "ctc" -p no:cacheprovider > $O/aa_pytest_ctc.log 2>&1; tail -5 $O/aa_pytest_ctc.log
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -m gpu -p no:cacheprovider > $O/aa_pytest_model.log 2>&1; tail -3 $O/aa_pytest_model.log
timeout 900 python bench.py --no-cpu-baseline --no-parity > $O/aa_bench_cfgB.json 2> $O/aa_bench_cfgB.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/aa_bench_cfgB.json"))
print({k:d[k] for k in ("value","ms_per_step","own_kernel_ms_per_step","library_ms_per_step","gpu_launches")}, d["e2e"], d["clocks"])
print({k:(round(v["ms_per_step"],3),v["launches_per_step"]) for k,v in d["kernels"].items()})
print({k:(round(v["ms_per_1000_utt"],3), round(v["frac_hbm"],4)) for k,v in d["micro"].items()})
PY
