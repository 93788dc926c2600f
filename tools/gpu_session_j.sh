#!/bin/bash
# 2-GPU session: data-parallel parity test + the torchrun bench line the driver will run at N=2
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "ctc" > $O/j_ctc.log 2>&1; tail -3 $O/j_ctc.log
timeout 900 python -m pytest tests/test_gpu_zz_dp.py -m gpu -q -p no:cacheprovider > $O/j_dp.log 2>&1; tail -5 $O/j_dp.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/j_bench_2gpu.json 2> $O/j_bench_2gpu.log
tail -3 $O/j_bench_2gpu.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/j_bench_2gpu.json"))
print({k:d.get(k) for k in ("value","n_gpus","ms_per_step","cuda_graph")}, d.get("e2e"))
print("also", d.get("also"))
print({k:(round(v["ms_per_1000_utt"],3), round(v["frac_hbm"],4)) for k,v in (d.get("micro") or {}).items()})
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/j_ref_2gpu.json 2> $O/j_ref_2gpu.log; cat $O/j_ref_2gpu.json | cut -c1-400
