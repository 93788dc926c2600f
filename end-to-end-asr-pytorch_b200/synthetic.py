"""Seeded synthetic workloads of SURVEY.md 8(d): LibriSpeech-shaped waveforms + token targets, and the yaml configs
under config/b200/ that define BASELINE.json's model configurations (the reference ships no such yamls)."""
import os

import torch
import yaml

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKLOADS = {"cfgB": "cfgB_ctc_char", "cfgC": "cfgC_hybrid_subword", "cfgD": "cfgD_cnn_blstm5x640"}


def load_config(name):
    path = os.path.join(_ROOT, "config", "b200", WORKLOADS.get(name, name) + ".yaml")
    with open(path) as f:
        return yaml.safe_load(f)


def make_batch(vocab_size, batch, n_samples=192000, seed=0, ragged=False):
    """waveform x = clamp(0.05*randn(N), -1, 1) (+ a few sinusoids so mel bands differ), 16 kHz; char targets
    L ~ U[80,140], subword targets L ~ U[25,45], ids in [3, V) + <eos>=1, zero padded (<pad>=0 is the CTC blank).
    Returns CPU tensors (waves [B,N] fp32, wave_len [B] int64, txt [B,Lmax] int64), sorted by length (desc)."""
    g = torch.Generator().manual_seed(seed)
    if ragged:
        lens = torch.randint(int(n_samples * 2 / 3), n_samples + 1, (batch,), generator=g)
        lens = torch.sort(lens, descending=True)[0]
        lens[0] = n_samples
    else:
        lens = torch.full((batch,), n_samples, dtype=torch.long)
    t = torch.arange(n_samples, dtype=torch.float32) / 16000.0
    waves = torch.zeros(batch, n_samples)
    for b in range(batch):
        x = 0.05 * torch.randn(n_samples, generator=g)
        f = 100.0 + 3000.0 * torch.rand(3, generator=g)
        for k in range(3):
            x += 0.02 * torch.sin(2 * 3.14159265 * f[k] * t)
        x = torch.clamp(x, -1, 1)
        n = int(lens[b])
        waves[b, :n] = x[:n]
    char = vocab_size <= 64
    lo, hi = (80, 141) if char else (25, 46)
    scale = n_samples / 192000.0
    lo, hi = max(2, int(lo * scale)), max(3, int(hi * scale))
    tl = torch.randint(lo, hi, (batch,), generator=g)
    txt = torch.zeros(batch, int(tl.max()) + 1, dtype=torch.long)
    for b in range(batch):
        L = int(tl[b])
        txt[b, :L] = torch.randint(3, vocab_size, (L,), generator=g)
        txt[b, L] = 1
    return waves, lens, txt
