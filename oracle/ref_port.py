"""CPU ORACLE (test infrastructure, NOT a product path): functional torch-CPU port of the reference train step.

The reference's hot path is Python glue over ATen / torchaudio kernels.  This module restates that glue as pure
functions over a `{state_dict key: tensor}` dictionary and calls the SAME third-party CPU kernels the reference
calls (torch._VF.lstm == nn.LSTM's forward, F.conv1d, F.ctc_loss, F.cross_entropy, torchaudio.compliance.kaldi.fbank),
so it is numerically the reference's `--cpu` path and is what bench.py times as the CPU baseline on the GPU box
(where /root/reference does not exist).  Pinned against the real reference by tests/golden/*.npz.

Citations: src/audio.py:25-27,51-54,85-89,104-108; src/module.py:75-88,129-156,179-195,234-258; src/asr.py:72-155,
207-221,277-313,363-366; bin/train_asr.py:47-49,115-132; src/solver.py:83-85.
"""
import math

import torch
import torch.nn.functional as F

from .oracle_np import delta_filters


# ------------------------------------------------------------------------------------------------ front end
def frontend(wave, audio_cfg):
    """wave [1,N] fp32 -> [m, D] features, like create_transform(cfg)(path) after loading (src/audio.py:115-133)."""
    from torchaudio.compliance import kaldi
    cfg = dict(audio_cfg)
    assert cfg.pop("feat_type") == "fbank"
    n_mel = cfg.pop("feat_dim")
    order = cfg.pop("delta_order", 0)
    win = cfg.pop("delta_window_size", 2)
    cmvn = cfg.pop("apply_cmvn")
    sr = cfg.pop("sample_frequency", 16000.0)
    y = kaldi.fbank(wave, num_mel_bins=n_mel, channel=-1, sample_frequency=sr, **cfg)    # [m, n_mel]
    x = y.t().unsqueeze(0)                                                                # [1, n_mel, m]
    if order >= 1:
        filt = torch.tensor(delta_filters(order, win), dtype=torch.float32).unsqueeze(1).unsqueeze(1)
        x = F.conv2d(x.unsqueeze(0), weight=filt, padding=(0, (filt.shape[-1] - 1) // 2))[0]
    if cmvn:
        x = (x - x.mean(2, keepdim=True)) / (1e-10 + x.std(2, keepdim=True))
    return x.permute(2, 0, 1).reshape(x.shape[2], -1)


def collate(waves, audio_cfg, texts):
    """Per-utterance front end, sort by length (desc), zero-pad (src/data.py:14-43)."""
    feats = [frontend(w, audio_cfg) for w in waves]
    order = sorted(range(len(feats)), key=lambda i: feats[i].shape[0], reverse=True)
    feats = [feats[i] for i in order]
    texts = [torch.as_tensor(texts[i], dtype=torch.long) for i in order]
    flen = torch.tensor([f.shape[0] for f in feats], dtype=torch.long)
    feat = torch.nn.utils.rnn.pad_sequence(feats, batch_first=True)
    txt = torch.nn.utils.rnn.pad_sequence(texts, batch_first=True)
    return feat, flen, txt, order


# ------------------------------------------------------------------------------------------------ model
def _lstm(P, prefix, x, bidir, hx=None, num_layers=1):
    names = []
    for l in range(num_layers):
        for sfx in ([""] + (["_reverse"] if bidir else [])):
            names += ["%sweight_ih_l%d%s" % (prefix, l, sfx), "%sweight_hh_l%d%s" % (prefix, l, sfx),
                      "%sbias_ih_l%d%s" % (prefix, l, sfx), "%sbias_hh_l%d%s" % (prefix, l, sfx)]
    flat = [P[n] for n in names]
    H = P[prefix + "weight_hh_l0"].shape[1]
    B = x.shape[0]
    if hx is None:
        z = x.new_zeros(num_layers * (2 if bidir else 1), B, H)
        hx = (z, z)
    out, h, c = torch._VF.lstm(x, hx, flat, True, num_layers, 0.0, False, bidir, True)
    return out, (h, c)


def encoder(P, enc_cfg, x, x_len, collect=None):
    """Prenet + stacked BiLSTM layers over the PADDED frames (src/asr.py:363-366, src/module.py:129-156)."""
    li = 0
    if enc_cfg["prenet"] == "cnn":
        pre = "encoder.layers.0.extractor."
        y = x.transpose(1, 2)
        y = F.conv1d(y, P[pre + "0.weight"], P[pre + "0.bias"], stride=2, padding=1)
        y = F.conv1d(y, P[pre + "1.weight"], P[pre + "1.bias"], stride=2, padding=1)
        x, x_len = y.transpose(1, 2), x_len // 4
        li = 1
    elif enc_cfg["prenet"] == "vgg":
        pre = "encoder.layers.0.extractor."
        x_len = x_len // 4
        if x.shape[1] % 4:
            x = x[:, :-(x.shape[1] % 4)]
        B, T, D = x.shape
        ch = D // 40 if D % 40 == 0 else D // 13
        y = x.reshape(B, T, ch, D // ch).transpose(1, 2)
        for idx in (0, 2, 5, 7):
            y = F.relu(F.conv2d(y, P[pre + "%d.weight" % idx], P[pre + "%d.bias" % idx], padding=1))
            if idx in (2, 7):
                y = F.max_pool2d(y, 2, stride=2)
        y = y.transpose(1, 2)
        x = y.reshape(y.shape[0], y.shape[1], -1)
        li = 1
    bidir = enc_cfg["bidirection"]
    for l in range(len(enc_cfg["dim"])):
        pre = "encoder.layers.%d." % (l + li)
        x, _ = _lstm(P, pre + "layer.", x, bidir)
        if enc_cfg["layer_norm"][l]:
            x = F.layer_norm(x, x.shape[-1:], P[pre + "ln.weight"], P[pre + "ln.bias"])
        r = enc_cfg["sample_rate"][l]
        if r > 1:
            x_len = x_len // r
            if enc_cfg["sample_style"] == "drop":
                x = x[:, ::r].contiguous()
            else:
                T = x.shape[1]
                if T % r:
                    x = x[:, :-(T % r)]
                x = x.contiguous().view(x.shape[0], T // r, x.shape[2] * r)
        if enc_cfg["proj"][l]:
            x = torch.tanh(F.linear(x, P[pre + "pj.weight"], P[pre + "pj.bias"]))
        if collect is not None:
            collect.append(x)
    return x, x_len


def loc_attention_decode(P, att_cfg, dec_cfg, enc, enc_len, teacher_ids, steps):
    """Teacher-forced (tf_rate=1) or greedy attention decoding with single-head location-aware attention
    (src/asr.py:101-151, src/module.py:234-258).  Returns (att_output [B,L,V], att_seq [B,1,L,T])."""
    assert att_cfg["mode"] == "loc" and att_cfg["num_head"] == 1 and not att_cfg["v_proj"]
    B, T, _ = enc.shape
    dim = dec_cfg["dim"]
    nl = dec_cfg["layer"]
    dev = enc.device
    temp = att_cfg["temperature"]
    key = torch.tanh(F.linear(enc, P["attention.proj_k.weight"], P["attention.proj_k.bias"]))
    pad = torch.arange(T, device=dev)[None, :] >= enc_len.to(dev)[:, None]
    prev = (~pad).float() / enc_len.to(dev).float()[:, None]
    h = enc.new_zeros(nl, B, dim)
    c = enc.new_zeros(nl, B, dim)
    emb = P["pre_embed.weight"]
    last = emb[torch.zeros(B, dtype=torch.long, device=dev)]
    teach = emb[teacher_ids] if teacher_ids is not None else None
    outs, atts = [], []
    r = att_cfg["loc_kernel_size"]
    for t in range(steps):
        q = torch.tanh(F.linear(h.transpose(0, 1).reshape(B, -1), P["attention.proj_q.weight"],
                                P["attention.proj_q.bias"]))
        conv = F.conv1d(prev.unsqueeze(1), P["attention.att_layer.loc_conv.weight"], padding=r)       # [B,K,T]
        loc = torch.tanh(F.linear(conv.transpose(1, 2), P["attention.att_layer.loc_proj.weight"]))
        e = F.linear(torch.tanh(key + q.unsqueeze(1) + loc), P["attention.att_layer.gen_energy.weight"],
                     P["attention.att_layer.gen_energy.bias"]).squeeze(2)
        a = torch.softmax((e / temp).masked_fill(pad, float("-inf")), dim=-1)
        ctx = torch.bmm(a.unsqueeze(1), enc).squeeze(1)
        prev = a
        x = torch.cat([last, ctx], dim=-1).unsqueeze(1)
        y, (h, c) = _lstm(P, "decoder.layers.", x, False, (h, c), nl)
        y = y.squeeze(1)
        logit = F.linear(y, P["decoder.char_trans.weight"], P["decoder.char_trans.bias"])
        last = teach[:, t] if teach is not None else emb[logit.argmax(-1)]
        outs.append(logit)
        atts.append(a.unsqueeze(1))
    return torch.stack(outs, 1), torch.stack(atts, 2)


def forward_losses(P, model_cfg, feat, feat_len, txt, teacher_forcing=True, collect=None):
    """ASR.forward + both losses exactly as bin/train_asr.py:104-132 assembles them (tf_rate = 1)."""
    lam = model_cfg["ctc_weight"]
    enc, enc_len = encoder(P, model_cfg["encoder"], feat, feat_len, collect)
    txt_len = (txt != 0).sum(-1)
    res = {"encode_len": enc_len, "enc": enc}
    total = 0.0
    if lam > 0:
        lp = F.log_softmax(F.linear(enc, P["ctc_layer.weight"], P["ctc_layer.bias"]), dim=-1)
        ctc = F.ctc_loss(lp.transpose(0, 1), txt, enc_len, txt_len, blank=0, reduction="mean", zero_infinity=False)
        res.update(ctc_output=lp, ctc_loss=ctc)
        total = total + ctc * lam
    if lam != 1:
        L = int(txt_len.max())
        att, seq = loc_attention_decode(P, model_cfg["attention"], model_cfg["decoder"], enc, enc_len,
                                        txt if teacher_forcing else None, L)
        ce = F.cross_entropy(att.reshape(-1, att.shape[-1]), txt[:, :L].reshape(-1), ignore_index=0)
        res.update(att_output=att, att_seq=seq, att_loss=ce)
        total = total + ce * (1 - lam)
    res["total_loss"] = total
    return res


def grad_norm_clip(grads, max_norm=5.0):
    """torch.nn.utils.clip_grad_norm_ (src/solver.py:84-85): returns (total_norm, clip coefficient)."""
    tot = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    return tot, torch.clamp(max_norm / (tot + 1e-6), max=1.0)


def adadelta_update(p, g, sq, acc, lr=1.0, rho=0.9, eps=1e-8):
    """torch.optim.Adadelta single-tensor update (src/optim.py:52 -> torch)."""
    sq.mul_(rho).addcmul_(g, g, value=1 - rho)
    std = sq.add(eps).sqrt_()
    delta = acc.add(eps).sqrt_().div_(std).mul_(g)
    acc.mul_(rho).addcmul_(delta, delta, value=1 - rho)
    p.add_(delta, alpha=-lr)


class CpuTrainer:
    """Whole reference-equivalent CPU train step on a parameter dict: front end per utterance, forward, losses,
    backward, clip, Adadelta.  This is the `cpu_baseline` / `--impl reference` leg of bench.py."""

    def __init__(self, P, model_cfg, audio_cfg, lr=1.0, eps=1e-8):
        self.P = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
        self.model_cfg, self.audio_cfg = model_cfg, audio_cfg
        self.sq = {k: torch.zeros_like(v) for k, v in self.P.items()}
        self.acc = {k: torch.zeros_like(v) for k, v in self.P.items()}
        self.lr, self.eps = lr, eps

    def step(self, waves, texts):
        feat, flen, txt, _ = collate(waves, self.audio_cfg, texts)
        for v in self.P.values():
            v.grad = None
        res = forward_losses(self.P, self.model_cfg, feat, flen, txt)
        res["total_loss"].backward()
        used = [k for k, v in self.P.items() if v.grad is not None]
        norm, coef = grad_norm_clip([self.P[k].grad for k in used])
        if not math.isnan(float(norm)):
            with torch.no_grad():
                for k in used:
                    adadelta_update(self.P[k], self.P[k].grad * coef, self.sq[k], self.acc[k], self.lr, 0.9, self.eps)
        return float(res["total_loss"]), float(norm)
