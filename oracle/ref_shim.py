"""Import the UNMODIFIED reference from /root/reference (only possible in the build container - the path does
not exist on the GPU box).  Three shims, none of which touch the reference's arithmetic:
  1. `editdistance` and `matplotlib` are imported at module top by src/util.py:1,7,9-10 but are not installed ->
     stub modules (a 10-line Levenshtein; no-op pyplot);
  2. `torchaudio.load` needs TorchCodec here -> scipy.io.wavfile, int16/32768 -> fp32 [1,N] (the historic default);
  3. nothing else: torchaudio.compliance.kaldi.fbank, torch.nn.LSTM, CTCLoss ... are the real installed ones.
Used by oracle/make_golden.py to produce tests/golden/*.npz.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("B200ASR_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "src"))


def _levenshtein(a, b):
    prev = list(range(len(b) + 1))
    for i, x in enumerate(a, 1):
        cur = [i]
        for j, y in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (x != y)))
        prev = cur
    return prev[-1]


def install():
    """Make `import src.asr`, `import src.audio`, `import bin.train_asr` resolve to the reference."""
    if not available():
        raise RuntimeError("reference not present at %s" % REF_ROOT)
    if "editdistance" not in sys.modules:
        m = types.ModuleType("editdistance")
        m.eval = _levenshtein
        sys.modules["editdistance"] = m
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import torchaudio
    import numpy as np
    import torch
    from scipy.io import wavfile

    def _load(path, *a, **k):
        sr, data = wavfile.read(path)
        x = data.astype(np.float32) / (32768.0 if data.dtype == np.int16 else 1.0)
        x = x[None, :] if x.ndim == 1 else x.T
        return torch.from_numpy(np.ascontiguousarray(x)), sr

    torchaudio.load = _load
    return REF_ROOT
