"""BASELINE.json configs[4] outside bench.py: mel-fbank (+delta+CMVN) and CTC micro-benchmark on 1000 synthetic
utterances of 2-30 s (the same function bench.py's `micro` section runs), plus the attention step at the cfg-C shape.
Prints one JSON line per kernel."""
import importlib, json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("end-to-end-asr-pytorch_b200")
dev = torch.device("cuda", 0)
peak, _ = bench.peaks()
for k, v in bench.micro_bench(pkg, dev, peak).items():
    print(json.dumps(dict(kernel=k, **v)))
B, T, D, E, K, R = 64, 149, 300, 2048, 10, 100
q, key, val = torch.randn(B, D, device=dev), torch.randn(B, T, D, device=dev), torch.randn(B, T, E, device=dev)
prev = torch.full((B, T), 1.0 / T, device=dev); ln = torch.full((B,), T, device=dev)
cw, pw, ew, eb = torch.randn(K, 1, 2 * R + 1, device=dev) * .1, torch.randn(D, K, device=dev) * .3, torch.randn(1, D, device=dev) * .1, torch.zeros(1, device=dev)
fn = lambda: pkg.ops.loc_attention_step(q, key, val, prev, ln, cw, pw, ew, eb, 0.5)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3): fn()
ms = []
for _ in range(10):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
m = sorted(ms)[5]; by = 4 * B * T * (D + E)
print(json.dumps(dict(kernel="locattn_fwd cfgC step", ms=m, algorithmic_gbs=by / m / 1e6, frac_hbm=by / m / 1e6 / peak)))
