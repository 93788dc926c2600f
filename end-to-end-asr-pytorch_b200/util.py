"""Helpers with the reference's names (/root/reference/src/util.py): initialisation, timer, error rate."""
import math
import time

import numpy as np
import torch
from torch import nn


class Timer:
    """Wall-clock split of a train step into read / forward / backward (src/util.py:13-42).  Unlike the
    reference, `sync=True` synchronises the device first so GPU numbers are not skewed."""

    def __init__(self, sync=False):
        self.sync = sync
        self.prev_t = time.time()
        self.clear()

    def _now(self):
        if self.sync and torch.cuda.is_available():
            torch.cuda.synchronize()
        return time.time()

    def set(self):
        self.prev_t = self._now()

    def cnt(self, mode):
        self.time_table[mode] += self._now() - self.prev_t
        self.set()
        if mode == "bw":
            self.click += 1

    def show(self):
        total = sum(self.time_table.values())
        tt = self.time_table
        avg = total / max(self.click, 1)
        msg = "{:.3f} sec/step (rd {:.1f}% | fw {:.1f}% | bw {:.1f}%)".format(
            avg, 100 * tt["rd"] / total, 100 * tt["fw"] / total, 100 * tt["bw"] / total)
        self.clear()
        return msg

    def clear(self):
        self.time_table = {"rd": 0, "fw": 0, "bw": 0}
        self.click = 0


def init_weights(module):
    """ESPnet-style init the reference applies with `self.apply` (src/util.py:47-71): embeddings N(0,1), biases 0,
    linear weights N(0, 1/sqrt(fan_in)), conv weights N(0, 1/sqrt(fan_in*k))."""
    if type(module) == nn.Embedding:
        module.weight.data.normal_(0, 1)
        return
    for p in module.parameters():
        data = p.data
        if data.dim() == 1:
            data.zero_()
        elif data.dim() == 2:
            data.normal_(0, 1.0 / math.sqrt(data.size(1)))
        elif data.dim() in (3, 4):
            n = data.size(1)
            for k in data.size()[2:]:
                n *= k
            data.normal_(0, 1.0 / math.sqrt(n))
        else:
            raise NotImplementedError


def init_gate(bias):
    """Forget-gate bias = 1 (src/util.py:74-77)."""
    n = bias.size(0)
    bias.data[n // 4:n // 2].fill_(1.0)
    return bias


def human_format(num):
    magnitude = 0
    while num >= 1000:
        magnitude += 1
        num /= 1000.0
    return "{:3.1f}{}".format(num, [" ", "K", "M", "G", "T", "P"][magnitude])


def edit_distance(a, b):
    """Levenshtein distance between two sequences (replaces the `editdistance` dependency)."""
    if len(a) < len(b):
        a, b = b, a
    prev = list(range(len(b) + 1))
    for i, x in enumerate(a, 1):
        cur = [i]
        for j, y in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (x != y)))
        prev = cur
    return prev[-1]


def cal_er(tokenizer, pred, truth, mode="wer", ctc=False):
    """Batch error rate from logits / ids (src/util.py:113-127)."""
    if pred is None:
        return np.nan
    if len(pred.shape) >= 3:
        pred = pred.argmax(dim=-1)
    er = []
    for p, t in zip(pred.tolist(), truth.tolist()):
        p = tokenizer.decode(p, ignore_repeat=ctc)
        t = tokenizer.decode(t)
        if mode == "wer":
            p = p.split(" ")
            t = t.split(" ")
        er.append(float(edit_distance(p, t)) / len(t))
    return sum(er) / len(er)
