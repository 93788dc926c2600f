"""Per-parameter gradient error of the vgg golden model in both GEMM modes (debug aid)."""
import importlib, os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("end-to-end-asr-pytorch_b200")
from oracle.make_golden import tiny_model_cfg
DEV = torch.device("cuda:0")
kind = os.environ.get("KIND", "vgg")
g = dict(np.load(os.path.join(ROOT, "tests", "golden", "model_%s.npz" % kind)))
for mode in ("tf32x3", "umma"):
    pkg.ops.GEMM_MODE = mode
    cfg = tiny_model_cfg(kind)
    V = g["sd.ctc_layer.weight"].shape[0] if "sd.ctc_layer.weight" in g else g["sd.pre_embed.weight"].shape[0]
    model = pkg.ASR(int(g["feat"].shape[-1]), int(V), True, **cfg)
    sd = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd.")}
    model.load_state_dict(sd)
    model = model.to(DEV).train()
    feat = torch.from_numpy(g["feat"]).to(DEV)
    flen = torch.from_numpy(g["feat_len"]).to(DEV)
    txt = torch.from_numpy(g["txt"]).to(DEV)
    txt_len = (txt != 0).sum(-1)
    ctc_out, enc_len, att_out, att_seq, _ = model(feat, flen, int(txt_len.max()), tf_rate=1.0, teacher=txt)
    total = 0
    if ctc_out is not None:
        total = total + pkg.CTCLoss(blank=0)(ctc_out.transpose(0, 1), txt, enc_len, txt_len) * model.ctc_weight
    if att_out is not None:
        b, t, _ = att_out.shape
        total = total + pkg.ops.cross_entropy(att_out.view(b * t, -1), txt[:, :t].reshape(-1), ignore_index=0) * (1 - model.ctc_weight)
        print(mode, "att_out max abs err %.3e (max |ref| %.3e)" % (float(np.abs(att_out.detach().cpu().numpy() - g["att_output"]).max()), float(np.abs(g["att_output"]).max())))
    total.backward()
    print(mode, "loss", float(total), "ref", float(g["total_loss"]))
    for k, p in model.named_parameters():
        ref = g.get("grad." + k)
        if ref is None:
            continue
        err = float(np.abs(p.grad.cpu().numpy() - ref).max())
        print("  %-55s shape %-18s max|ref| %.3e  err %.3e  rel %.2e" % (k, tuple(ref.shape), float(np.abs(ref).max()), err, err / max(float(np.abs(ref).max()), 1e-12)))
