"""Share of the step's kernel time per kernel from an `ncu --metrics gpu__time_duration.sum --csv` launch list
(per-launch times under ncu are serialised and cold-cache: the SHARES are the evidence, not the absolute values).
usage: python tools/launch_shares.py gpurun_out/r02_launches.csv "<title line>" > profiles/rNN_launches_cfgB.txt"""
import csv, re, sys
from collections import defaultdict

path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rows[1:]:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1e-3)
    name = re.sub(r"\(.*", "", r[ik])
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
own = sum(v for k, v in tot.items() if "b200asr" in k)
print(title)
print("per-launch times under ncu are serialised and cold-cache: the SHARES are the evidence, not the absolute values")
print("total kernel time over the captured launches: %.1f ms in %d launches" % (total / 1e3, sum(cnt.values())))
print("share of this library's kernels: %.2f%%   (ATen / cuBLAS glue: %.2f%%)" % (100 * own / total, 100 - 100 * own / total))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:40]:
    print("%6.2f%% %10.1f us %5d launches  %s" % (100 * v / total, v, cnt[k], k[-70:]))
