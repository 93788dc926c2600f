#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "attention" > $O/k_attn.log 2>&1; grep -E "^E  +(Assertion|assert)|Error" $O/k_attn.log | head; tail -3 $O/k_attn.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_zz_decode.py tests/test_gpu_zz_plumbing.py -m gpu -q -p no:cacheprovider > $O/k_model.log 2>&1; grep -E "^E  +(Assertion|assert)|Error" $O/k_model.log | head; tail -3 $O/k_model.log
for w in cfgC cfgD; do
timeout 600 python bench.py --workload $w --no-cpu-baseline --no-micro > $O/k_bench_$w.json 2> $O/k_bench_$w.log
done
python - <<'PY'
import json
for w in ("C", "D"):
    try: d = json.load(open("gpurun_out/k_bench_cfg%s.json" % w))
    except Exception as e: print(w, "failed", e); continue
    print(w, {k: d.get(k) for k in ("value", "ms_per_step", "own_kernel_ms_per_step", "library_ms_per_step")})
    print("  ", {k: (round(v["ms_per_step"], 3), v["launches_per_step"]) for k, v in d["kernels"].items()})
    if d.get("parity"): print("  parity", {k: {kk: vv for kk, vv in v.items() if "rel_err" in kk or "equal" in kk} for k, v in d["parity"].items()})
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/k_launches_cfgC.csv \
    python bench.py --workload cfgC --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-e2e --no-parity --no-micro > $O/k_launch_bench.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open("gpurun_out/k_launches_cfgC.csv")))
for i,r in enumerate(rows):
    if "Kernel Name" in r: hdr=r; start=i; break
ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
t=collections.defaultdict(float); n=collections.Counter()
for r in rows[start+1:]:
    if len(r)<=vi: continue
    try: v=float(r[vi].replace(",",""))
    except: continue
    name=r[ki].split("(")[0][-70:]
    t[name]+=v; n[name]+=1
tot=sum(t.values())
print("cfgC total %.1f ms in %d launches" % (tot/1e6, sum(n.values())))
for k,v in sorted(t.items(), key=lambda x:-x[1])[:45]:
    print("%6.2f%% %10.1f us %5d  %s"%(100*v/tot, v/1e3, n[k], k))
PY
