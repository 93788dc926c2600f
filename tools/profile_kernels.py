"""Driver for ncu: runs every non-LSTM hand-written kernel once (after one warm-up) at the cfg-C shapes."""
import importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("end-to-end-asr-pytorch_b200")
from oracle.make_golden import AUDIO_CFG
dev = "cuda"
g = torch.Generator().manual_seed(0)
B, N, T, V, D, E, K, R, L = 64, 192000, 149, 5000, 300, 2048, 10, 100, 46
tr, _ = pkg.create_transform(dict(AUDIO_CFG), device=dev)
wave = (0.05 * torch.randn(B, N, generator=g)).to(dev)
logits = torch.randn(B, T, V, generator=g).to(dev).requires_grad_(True)
txt = torch.zeros(B, L, dtype=torch.long)
tl = torch.randint(25, 45, (B,), generator=g)
for b in range(B): txt[b, :tl[b]] = torch.randint(3, V, (int(tl[b]),), generator=g)
txt = txt.to(dev); il = torch.full((B,), T, device=dev); tld = tl.to(dev)
q, key, val = [torch.randn(*s, generator=g).to(dev).requires_grad_(True) for s in ((B, D), (B, T, D), (B, T, E))]
prev = torch.full((B, T), 1.0 / T, device=dev, requires_grad=True); ln = torch.full((B,), T, device=dev)
cw, pw, ew, eb = [(torch.randn(*s, generator=g) * 0.2).to(dev).requires_grad_(True) for s in ((K, 1, 2 * R + 1), (D, K), (1, D), (1,))]
ce_logits = torch.randn(B * 40, V, generator=g).to(dev).requires_grad_(True)
ce_tgt = torch.randint(0, V, (B * 40,), generator=g).to(dev)
for it in range(2):
    tr.batch(wave, [N] * B)
    head = pkg.ops.ctc_head(logits)                      # the train step's fused CTC head (row lse + arg-max only)
    loss = pkg.CTCLoss(blank=0)(head.transpose(0, 1), txt, il, tld)
    loss.backward()
    mem, mk, mv, mcw, mpw, mew, meb, token = pkg.ops.attention_memory(key, val, cw, pw, ew, eb)   # decode-loop form
    c, a = pkg.ops.loc_attention_mem_step(mem, token, q, mk, mv, prev, ln, mcw, mpw, mew, meb, 0.5)
    c2, a2 = pkg.ops.loc_attention_mem_step(mem, token, q, mk, mv, a, ln, mcw, mpw, mew, meb, 0.5)
    (c.sum() + c2.sum() + (a2 * a2).sum()).backward()
    pkg.ops.cross_entropy(ce_logits, ce_tgt).backward()
    torch.cuda.synchronize()
print("done")
