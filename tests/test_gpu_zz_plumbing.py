"""BASELINE.json configs[0]: the reference's example experiment (VGG + BiLSTM 5x512 + location-aware attention,
config/libri/asr_example.yaml) run as a 2-step plumbing case on its sample wav - here through THIS package's Solver on
the GPU (main.py's sequence: Solver(config, paras, mode).load_data().set_model().exec()) against the losses and
grad-norms the reference's own Solver produced with --cpu on the same files, seed and batch order
(tests/golden/plumbing.npz, written by oracle/make_golden.golden_plumbing)."""
import argparse
import os

import pytest
import torch
import yaml

from conftest import load_golden
from oracle.make_golden import plumbing_config, plumbing_tree

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_configs0_example_yaml_plumbing_run_matches_reference(pkg, tmp_path):
    g = load_golden("plumbing.npz")
    pcm = load_golden("frontend.npz")["sample_pcm"]
    path = plumbing_tree(str(tmp_path), pcm, "wav")          # 16-bit PCM: stays int16 until the fbank kernel
    cfg = yaml.load(open(os.path.join(ROOT, "config", "b200", "cfgA_example_vgg.yaml")), Loader=yaml.FullLoader)
    cfg = plumbing_config(cfg, path, os.path.join(ROOT, "tests", "golden", "character.vocab"))
    paras = argparse.Namespace(config="cfgA_example_vgg.yaml", name="plumbing", logdir=str(tmp_path / "log"),
                               ckpdir=str(tmp_path / "ck"), outdir=str(tmp_path / "out"), load=None, seed=0, njobs=0,
                               gpu=True, pin_memory=False, verbose=False, amp=False)
    s = pkg.train_asr.Solver(cfg, paras, "train")
    s.load_data()
    torch.manual_seed(0)                                     # same seed -> the reference constructor's weights
    s.set_model()
    assert sum(p.numel() for p in s.model.parameters()) == int(g["n_params"])
    losses, norms, names = [], [], []
    real_backward, real_fetch = s.backward, s.fetch_data

    def backward(loss):
        losses.append(float(loss))
        n = real_backward(loss)
        norms.append(float(n))
        return n

    def fetch(data):
        names.append(",".join(data[0]))
        return real_fetch(data)

    s.backward, s.fetch_data = backward, fetch
    s.exec()
    assert s.step == 2 and len(losses) == 2
    assert names[:2] == [str(n) for n in g["names"][:2]]     # curriculum epoch: the same length-sorted batches
    for i in range(2):
        assert abs(losses[i] - float(g["loss"][i])) < 1e-4 * abs(float(g["loss"][i])), (i, losses, g["loss"])
        assert abs(norms[i] - float(g["grad_norm"][i])) < 5e-4 * float(g["grad_norm"][i]), (i, norms, g["grad_norm"])
    assert os.path.exists(str(tmp_path / "ck" / "plumbing" / "latest.pth"))
