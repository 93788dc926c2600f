"""torchrun worker of tests/test_gpu_zz_dp.py: a data-parallel step of the PRODUCT TrainStep on N GPUs must equal the
single-GPU step on the concatenated (global) batch - loss, gradient norm and parameters after two updates
(shards padded to the global T_max, CTC / CE normalised by the global batch / token counts, one SUM
all-reduce of the flat gradient).  Every rank checks itself against its own single-GPU replica of the global step."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    pkg = importlib.import_module("end-to-end-asr-pytorch_b200")
    from oracle.make_golden import AUDIO_CFG, tiny_model_cfg
    kind = sys.argv[1] if len(sys.argv) > 1 else "hybrid"
    dp = pkg.dist.DataParallel()
    assert dp.enabled and dp.backend == "nccl", "run under torchrun with >= 2 GPUs"
    dev = torch.device("cuda", dp.local_rank)
    cfg = {"data": {"audio": dict(AUDIO_CFG)},
           "hparas": {"valid_step": 1000, "max_step": 2, "tf_start": 1.0, "tf_end": 1.0, "tf_step": 10,
                      "optimizer": "Adadelta", "lr": 1.0, "eps": 1e-8, "lr_scheduler": "fixed", "curriculum": 0},
           "model": tiny_model_cfg(kind)}
    V = 12
    G = 4 * dp.world                                         # global batch, sorted by length (descending)
    g = torch.Generator().manual_seed(21)
    lens = sorted([int(v) for v in torch.randint(6000, 12001, (G,), generator=g)], reverse=True)
    wave = torch.zeros(G, lens[0])
    for i, n in enumerate(lens):
        wave[i, :n] = torch.clamp(0.05 * torch.randn(n, generator=g), -1, 1)
    tl = torch.randint(2, 6, (G,), generator=g)
    txt = torch.zeros(G, int(tl.max()) + 1, dtype=torch.long)
    for i in range(G):
        txt[i, :int(tl[i])] = torch.randint(3, V, (int(tl[i]),), generator=g)
        txt[i, int(tl[i])] = 1
    lens_t = torch.tensor(lens)
    ntok = float((txt != 0).sum())

    dp_step = pkg.TrainStep(cfg, V, device=dev, dp=dp, seed=7)
    torch.distributed.broadcast(dp_step.optimizer.buf.flat, 0)
    one = pkg.dist.DataParallel.single()                    # a disabled communicator: the 1-GPU reference
    ref_step = pkg.TrainStep(cfg, V, device=dev, dp=one, seed=7)
    ref_step.optimizer.buf.flat.copy_(dp_step.optimizer.buf.flat)

    w, l, t = dp.shard(wave, lens_t, txt)                     # rows rank::world of the globally padded batch
    for it in range(2):
        loss_part = dp_step(w.to(dev), l, t.to(dev), global_batch=G, global_tokens=ntok)
        loss = loss_part.detach().clone()
        torch.distributed.all_reduce(loss)                    # each rank holds its share of the globally normalised loss
        ref_loss = ref_step(wave.to(dev), lens_t, txt.to(dev), max_len=int(txt.shape[1]))
        assert abs(float(loss) - float(ref_loss)) < 1e-5 * abs(float(ref_loss)), (it, float(loss), float(ref_loss))
        gn, gr = float(dp_step.last["grad_norm"]), float(ref_step.last["grad_norm"])
        assert abs(gn - gr) < 1e-4 * gr, (it, gn, gr)
    a, b = dp_step.optimizer.buf.flat, ref_step.optimizer.buf.flat
    err = float((a - b).abs().max()) / float(b.abs().max())
    assert err < 1e-5, err
    # all ranks hold identical parameters after the updates
    chk = a.double().sum().reshape(1).clone()
    lo, hi = chk.clone(), chk.clone()
    torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN)
    torch.distributed.all_reduce(hi, op=torch.distributed.ReduceOp.MAX)
    assert float(lo) == float(hi)
    if dp.rank == 0:
        print("DP_OK kind=%s world=%d loss=%.6f param_rel_err=%.2e" % (kind, dp.world, float(ref_loss), err))
    dp.barrier()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
