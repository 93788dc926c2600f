"""How well conditioned are the cfg-B parameter gradients at the benchmark's size?  CPU only (no GPU, no kernels):
the reference-equivalent CPU path (oracle/ref_port.py) run twice on the parity block's inputs - first 8 utterances of
the bench batch, T = 1198 frames, same initial weights - once in fp32 (what bench.py's `parity` compares against) and
once in fp64.  Prints, per parameter tensor, max|g32 - g64| / max|g64| and the same for the gradient norm: the floor
below which a difference between two fp32 implementations says nothing about either."""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("end-to-end-asr-pytorch_b200")
from oracle import ref_port

W = os.environ.get("WORKLOAD", "cfgB")
N = int(os.environ.get("N_REF", 8))
cfg = pkg.synthetic.load_config(W)
vocab = cfg["data"]["corpus"]["vocab_size"]
torch.manual_seed(0)
_, feat_dim = pkg.create_transform(dict(cfg["data"]["audio"]), device="cpu")
model = pkg.ASR(feat_dim, vocab, cfg["hparas"]["optimizer"] == "Adadelta", **cfg["model"])
P = {k: v.detach().clone() for k, v in model.state_dict().items()}
waves, lens, txt = pkg.synthetic.make_batch(vocab, cfg["data"]["corpus"]["batch_size"], 192000, seed=1000)
wl = [waves[b:b + 1, :int(lens[b])] for b in range(N)]
tl = [[int(x) for x in txt[b] if int(x) != 0] for b in range(N)]
feat, flen, t_ref, _ = ref_port.collate(wl, cfg["data"]["audio"], tl)
torch.set_num_threads(os.cpu_count() or 1)
grads = {}
for dt in (torch.float32, torch.float64):
    Pd = {k: (v.to(dt) if v.is_floating_point() else v).clone().requires_grad_(v.is_floating_point()) for k, v in P.items()}
    res = ref_port.forward_losses(Pd, cfg["model"], feat.to(dt), flen, t_ref)
    res["total_loss"].backward()
    grads[dt] = {k: v.grad.double() for k, v in Pd.items() if v.requires_grad and v.grad is not None}
    print("%s loss %.9f" % (dt, float(res["total_loss"])), flush=True)
g32, g64 = grads[torch.float32], grads[torch.float64]
n32 = sum((g ** 2).sum() for g in g32.values()).sqrt()
n64 = sum((g ** 2).sum() for g in g64.values()).sqrt()
print("grad-norm fp32 %.9f fp64 %.9f rel err %.3e" % (n32, n64, abs(n32 - n64) / n64))
rows = sorted(((float((g32[k] - g64[k]).abs().max() / g64[k].abs().max()), k) for k in g64), reverse=True)
for e, k in rows[:12]:
    print("  %.3e  %s" % (e, k))
