"""BASELINE.json configs[4]: mel-fbank (+delta+CMVN) and CTC alpha/beta micro-benchmark on 1000 synthetic utterances
of 2-30 s; reports achieved algorithmic GB/s against the measured HBM peak.  Also times the attention step kernels at
the cfg-C shape.  Prints one JSON line per kernel."""
import importlib, json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("end-to-end-asr-pytorch_b200")
from oracle.make_golden import AUDIO_CFG
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
dev = "cuda"


def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ms = []
    for _ in range(iters):
        flush.zero_()                                   # L2 flush (256 MB > 126 MB L2)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    ms.sort()
    return ms[len(ms) // 2]


def report(name, ms, nbytes, extra=None):
    gbs = nbytes / (ms * 1e-3) / 1e9
    d = {"kernel": name, "ms": round(ms, 4), "algorithmic_bytes": int(nbytes), "achieved_gbs": round(gbs, 1),
         "peak_gbs": PEAK, "frac": round(gbs / PEAK, 4)}
    d.update(extra or {})
    print(json.dumps(d))


g = torch.Generator().manual_seed(0)
# ---- front end: 1000 utterances, 2..30 s, processed in batches of 100 (zero padded to the batch max)
tr, _ = pkg.create_transform(dict(AUDIO_CFG), device=dev)
fe = tr.frontend
lens = torch.randint(32000, 480001, (1000,), generator=g).sort(descending=True)[0]
tot_ms_fb = tot_ms_dc = 0.0; by_fb = by_dc = 0
for i in range(0, 1000, 100):
    l = lens[i:i + 100]
    wave = torch.zeros(100, int(l[0]), device=dev)
    for b in range(100):
        wave[b, :int(l[b])] = 0.05 * torch.randn(int(l[b]), device=dev)
    pkg.lib.TIMER.enabled = True
    def run():
        pkg.lib.TIMER.reset(); fe(wave, l)
    run(); run(); torch.cuda.synchronize(); run(); torch.cuda.synchronize()
    s = pkg.lib.TIMER.summary()
    tot_ms_fb += s["fbank_fwd"]["ms"]; tot_ms_dc += s["delta_cmvn_fwd"]["ms"]
    m = torch.clamp((l - 400) // 160 + 1, min=0)
    by_fb += int(4 * l.sum() + 160 * m.sum()); by_dc += int(m.sum()) * 4 * (40 + 120)
pkg.lib.TIMER.enabled = False
report("fbank_fwd (1000 utts 2-30s, per-utt bytes 4N+160m)", tot_ms_fb, by_fb)
report("delta_cmvn_fwd (1000 utts, per-utt bytes 4m(40+120))", tot_ms_dc, by_dc)
# ---- CTC: V in {31, 5000}, T' = m//4, 1000 utts in batches of 100
for V, Lr in ((31, (20, 130)), (5000, (6, 45))):
    tot = 0.0; nb = 0
    for i in range(0, 1000, 100):
        l = lens[i:i + 100]
        T = ((l - 400) // 160 + 1) // 4
        Tm = int(T.max())
        lp = torch.randn(100, Tm, V, device=dev).log_softmax(-1)
        tl = torch.minimum(torch.randint(Lr[0], Lr[1], (100,), generator=g), (T // 3).clamp(min=1))
        txt = torch.zeros(100, int(tl.max()), dtype=torch.long)
        for b in range(100):
            txt[b, :tl[b]] = torch.randint(1, V, (int(tl[b]),), generator=g)
        txt = txt.to(dev); Td = T.to(dev); tld = tl.to(dev)
        crit = pkg.CTCLoss(blank=0)
        fn = lambda: crit(lp.transpose(0, 1), txt, Td, tld)
        tot += timeit(fn, iters=5, warm=2)
        nb += int((2 * 4 * T * V).sum()) + 400
    report("ctc_fwd_bwd V=%d (1000 utts, per-utt bytes 2*4*T'*V)" % V, tot, nb)
# ---- attention step at cfg C
B, T, D, E, K, R = 64, 149, 300, 2048, 10, 100
q, key, val = torch.randn(B, D, device=dev), torch.randn(B, T, D, device=dev), torch.randn(B, T, E, device=dev)
prev = torch.full((B, T), 1.0 / T, device=dev); ln = torch.full((B,), T, device=dev)
cw, pw, ew, eb = torch.randn(K, 1, 2 * R + 1, device=dev) * .1, torch.randn(D, K, device=dev) * .3, torch.randn(1, D, device=dev) * .1, torch.zeros(1, device=dev)
report("locattn_fwd cfgC step (bytes 4BT(D+E))", timeit(lambda: pkg.ops.loc_attention_step(q, key, val, prev, ln, cw, pw, ew, eb, 0.5)), 4 * B * T * (D + E))
