"""Model-level parity on the GPU: the reference's state_dict loaded into the B200 model must reproduce the
reference's outputs, losses, gradients and greedy ids (golden vectors from oracle/make_golden.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_err
from oracle.make_golden import tiny_model_cfg

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _build(pkg, g, kind):
    cfg = tiny_model_cfg(kind)
    D = g["feat"].shape[-1]
    V = g["sd.ctc_layer.weight"].shape[0] if "sd.ctc_layer.weight" in g else g["sd.pre_embed.weight"].shape[0]
    model = pkg.ASR(D, V, True, **cfg)
    sd = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd.")}
    assert set(sd.keys()) == set(model.state_dict().keys())              # identical state_dict contract
    for k, v in model.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    model.load_state_dict(sd)
    return model.to(DEV), cfg


@pytest.mark.parametrize("kind", ["ctc", "hybrid", "cnn", "att"])
def test_train_step_matches_reference(pkg, kind):
    g = load_golden("model_%s.npz" % kind)
    model, cfg = _build(pkg, g, kind)
    model.train()
    feat = torch.from_numpy(g["feat"]).to(DEV)
    flen = torch.from_numpy(g["feat_len"]).to(DEV)
    txt = torch.from_numpy(g["txt"]).to(DEV)
    txt_len = (txt != 0).sum(-1)
    ctc_out, enc_len, att_out, att_seq, _ = model(feat, flen, int(txt_len.max()), tf_rate=1.0, teacher=txt)
    assert np.array_equal(enc_len.cpu().numpy(), g["encode_len"])
    total = 0
    if ctc_out is not None:
        assert rel_err(ctc_out.detach().cpu().numpy(), g["ctc_output"]) < 1e-4
        assert np.array_equal(ctc_out.argmax(-1).cpu().numpy(), g["ctc_argmax"])          # bit exact ids
        assert np.array_equal(model.last_ctc_argmax.cpu().numpy(), g["ctc_argmax"])
        ctc = pkg.CTCLoss(blank=0)(ctc_out.transpose(0, 1), txt, enc_len, txt_len)
        assert abs(ctc.item() - float(g["ctc_loss"])) < 1e-4 * abs(float(g["ctc_loss"]))
        total = total + ctc * model.ctc_weight
    if att_out is not None:
        assert rel_err(att_out.detach().cpu().numpy(), g["att_output"]) < 1e-4
        assert rel_err(att_seq.detach().cpu().numpy(), g["att_seq"], floor=1e-4) < 1e-4
        assert np.array_equal(att_out.argmax(-1).cpu().numpy(), g["att_argmax"])
        b, t, _ = att_out.shape
        ce = pkg.ops.cross_entropy(att_out.view(b * t, -1), txt[:, :t].reshape(-1), ignore_index=0) \
            if hasattr(pkg.ops, "cross_entropy") else F.cross_entropy(att_out.view(b * t, -1),
                                                                      txt[:, :t].reshape(-1), ignore_index=0)
        assert abs(ce.item() - float(g["att_loss"])) < 1e-4 * abs(float(g["att_loss"]))
        total = total + ce * (1 - model.ctc_weight)
    assert abs(total.item() - float(g["total_loss"])) < 1e-4 * abs(float(g["total_loss"]))
    total.backward()
    sq = 0.0
    for k, p in model.named_parameters():
        key = "grad." + k
        if key in g:
            ref = g[key]
            scale = max(float(np.abs(ref).max()), 1e-4)
            assert float(np.abs(p.grad.cpu().numpy() - ref).max()) < 2e-4 * scale, k
            sq += float((p.grad.double() ** 2).sum())
    assert abs(np.sqrt(sq) - float(g["grad_norm"])) < 1e-4 * float(g["grad_norm"])


@pytest.mark.parametrize("kind", ["hybrid", "att"])
def test_greedy_inference_ids_bit_exact(pkg, kind):
    g = load_golden("model_%s.npz" % kind)
    model, cfg = _build(pkg, g, kind)
    model.eval()
    feat = torch.from_numpy(g["feat"]).to(DEV)
    flen = torch.from_numpy(g["feat_len"]).to(DEV)
    txt_len = (torch.from_numpy(g["txt"]) != 0).sum(-1)
    with torch.no_grad():
        _, _, out, _, _ = model(feat, flen, int(txt_len.max()) + 2)
    assert np.array_equal(out.argmax(-1).cpu().numpy(), g["greedy_argmax"])
    assert rel_err(out.cpu().numpy(), g["greedy_output"]) < 1e-4
