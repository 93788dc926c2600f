import importlib, sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("end-to-end-asr-pytorch_b200")
from oracle.make_golden import AUDIO_CFG
from oracle import ref_port, oracle_np as onp
torch.manual_seed(0)
B, N = 8, 192000
x = torch.clamp(0.05 * torch.randn(B, N), -1, 1)
raw, _ = pkg.create_transform(dict(AUDIO_CFG, delta_order=0, apply_cmvn=False))
full, _ = pkg.create_transform(dict(AUDIO_CFG))
xd = x.cuda()
fb1, n = raw.batch(xd, [N]*B)
fb2, _ = raw.batch(0.5*xd, [N]*B)
print("raw lin err", float((fb2-fb1-2*np.log(0.5)).abs().max()))
f1,_ = full.batch(xd, [N]*B); f2,_ = full.batch(0.5*xd, [N]*B)
d=(f1-f2).abs()
print("cmvn diff max", float(d.max()), "argmax", np.unravel_index(int(d.argmax()), d.shape))
print("per-order max", [float(d[:,:,40*o:40*(o+1)].max()) for o in range(3)])
print("per-batch max", [float(d[b].max()) for b in range(B)])
# vs torchaudio on CPU for utterance 0
ref = ref_port.frontend(x[:1], AUDIO_CFG)
print("vs kaldi cpu utt0: max abs", float((f1[0].cpu()-ref).abs().max()), "per order", [float((f1[0].cpu()-ref)[:,40*o:40*(o+1)].abs().max()) for o in range(3)])
refraw = ref_port.frontend(x[:1], dict(AUDIO_CFG, delta_order=0, apply_cmvn=False))
print("raw vs kaldi cpu utt0: max abs", float((fb1[0].cpu()-refraw).abs().max()))
# repeatability
f1b,_ = full.batch(xd, [N]*B)
print("repeat diff", float((f1-f1b).abs().max()))
t = d.amax(dim=(0,2)); print("worst frames", torch.topk(t,5))
