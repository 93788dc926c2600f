"""greedy decoding through the reference's `--test` Solver sequence on the GPU path."""
import argparse

import pytest
import torch
import yaml

from oracle.make_golden import AUDIO_CFG, tiny_model_cfg

pytestmark = pytest.mark.gpu


def _train_cfg(kind):
    return {"data": {"audio": dict(AUDIO_CFG),
                     "corpus": {"name": "Synthetic", "path": "", "train_split": ["syn"], "dev_split": ["syn"],
                                "bucketing": False, "batch_size": 3, "n_samples": 8000, "vocab_size": 12,
                                "n_batches": 4}, "text": {"mode": "character", "vocab_file": ""}},
            "hparas": {"valid_step": 1000, "max_step": 2, "tf_start": 1.0, "tf_end": 1.0, "tf_step": 10,
                       "optimizer": "Adadelta", "lr": 1.0, "eps": 1e-8, "lr_scheduler": "fixed", "curriculum": 0},
            "model": tiny_model_cfg(kind)}


@pytest.mark.parametrize("kind", ["hybrid", "ctc"])
def test_greedy_decode_solver_writes_reference_csv(pkg, tmp_path, kind):
    cfg = _train_cfg(kind)
    paras = argparse.Namespace(config="tiny.yaml", name="t", logdir=str(tmp_path / "log"), ckpdir=str(tmp_path / "ck"),
                               outdir=str(tmp_path / "out"), load=None, seed=0, njobs=0, gpu=True, pin_memory=False,
                               verbose=False, amp=False)
    if kind == "hybrid":                                        # checkpoint written by the training Solver itself
        tr = pkg.train_asr.Solver(cfg, paras, "train")
        tr.load_data()
        tr.set_model()
        tr.exec()                                               # two steps + validation, saves latest.pth
        trained = tr.model
    else:                                                       # a checkpoint in the reference's dict layout
        torch.manual_seed(1)
        trained = pkg.ASR(120, 12, True, **cfg["model"]).to("cuda")
        (tmp_path / "ck" / "t").mkdir(parents=True)
        torch.save({"model": trained.state_dict(), "optimizer": {}, "global_step": 2, "wer": 0.5},
                   str(tmp_path / "ck" / "t" / "latest.pth"))
    src_yaml = tmp_path / "train.yaml"
    src_yaml.write_text(yaml.safe_dump(cfg))
    dec_cfg = {"src": {"ckpt": str(tmp_path / "ck" / "t" / "latest.pth"), "config": str(src_yaml)},
               "data": {"corpus": {"name": "Synthetic", "dev_split": ["syn"], "test_split": ["syn"], "batch_size": 3,
                                   "n_samples": 8000, "vocab_size": 12}},
               "decode": {"beam_size": 1, "min_len_ratio": 0.01, "max_len_ratio": 0.2}}
    paras.load = None
    te = pkg.test_asr.Solver(dec_cfg, paras, "test")
    te.load_data()
    te.set_model()
    assert not te.model.training
    ck = torch.load(dec_cfg["src"]["ckpt"], map_location="cpu")
    assert set(ck["model"].keys()) == set(te.model.state_dict().keys()) == set(trained.state_dict().keys())
    for k, v in te.model.state_dict().items():
        assert torch.equal(v.cpu(), ck["model"][k]), k          # the checkpointed weights are what decodes
    te.exec()
    for split in ("dev", "test"):
        lines = (tmp_path / "out" / ("t_%s_output.csv" % split)).read_text().splitlines()
        assert lines[0] == "idx\thyp\ttruth"
        assert len(lines) == 1 + 2 * 3                          # two synthetic batches of three utterances
        for n, line in enumerate(lines[1:]):
            idx, hyp, truth = line.split("\t")
            assert idx == str(n) and len(hyp) >= 1 and len(truth) >= 1
    # the csv rows are exactly the arg-max ids of the same forward pass, decoded by the tokenizer
    data = next(iter(te.dv_set))
    feat, feat_len, txt, _ = te.fetch_data(data)
    with torch.no_grad():
        ctc_out, _, att_out, _, _ = te.model(feat, feat_len, int(float(feat_len.max()) * 0.2))
    ids = (att_out if att_out is not None else ctc_out).argmax(-1).cpu().tolist()
    first = (tmp_path / "out" / "t_dev_output.csv").read_text().splitlines()[1].split("\t")
    assert first[1] == (te.tokenizer.decode(ids[0]) or " ")
    # beam search is not on this path: refused with a pointer to the reference
    dec_cfg["decode"]["beam_size"] = 5
    with pytest.raises(NotImplementedError):
        pkg.test_asr.Solver(dec_cfg, paras, "test")
