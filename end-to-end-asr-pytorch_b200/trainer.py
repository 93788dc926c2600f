"""The train step (the hot path) as one object: waveforms -> fused front end -> ASR.forward -> CTC (+CE) ->
backward -> [grad all-reduce] -> fused norm/clip/update.  Used by the Solver mirror (train_asr.py) and bench.py.
Mirrors bin/train_asr.py:95-137 + src/solver.py:76-91 of the reference."""
import os

import torch
import torch.nn.functional as F

from . import ops
from .asr import ASR
from .audio import create_transform
from .dist import DataParallel
from .optim import Optimizer


class TrainStep:
    def __init__(self, config, vocab_size, device="cuda", dp=None, seed=0):
        self.device = torch.device(device)
        self.dp = dp or DataParallel()
        self.config = config
        torch.manual_seed(seed)
        self.transform, self.feat_dim = create_transform(dict(config["data"]["audio"]), device=self.device)
        init_adadelta = config["hparas"]["optimizer"] == "Adadelta"
        self.model = ASR(self.feat_dim, vocab_size, init_adadelta, **config["model"]).to(self.device)
        self.model.train()
        # the train step reads ctc_output only through CTCLoss and arg-max: fuse the log-softmax into the CTC kernels
        self.model.fuse_ctc_head = os.environ.get("B200ASR_FUSE_CTC", "1") == "1"
        self.ctc_loss = ops.CTCLoss(blank=0, zero_infinity=False)
        self.optimizer = Optimizer([{"params": self.model.parameters()}], **config["hparas"])
        self.dp.attach(self.optimizer, self.model)
        self.step_id = 0
        self.last = {}
        self.graph = None
        self.graph_error = None

    def front_end(self, wave, wave_len):
        return self.transform.batch(wave, wave_len)

    def __call__(self, wave, wave_len, txt, global_batch=None, global_tokens=None, max_len=None):
        """wave [B,N] fp32 (device), wave_len [B], txt [B,L] int64 (device); returns the total loss (device scalar).
        Under data parallelism `wave`/`txt` are this rank's shard, padded to the GLOBAL maxima."""
        if self.graph is not None:
            # the captured step is valid for the captured shapes and loss normalisers only; the utterance lengths
            # are read on the device, so they are refreshed like the waveforms (anything else: eager step)
            if (tuple(wave.shape) == tuple(self._g_wave.shape) and tuple(txt.shape) == tuple(self._g_txt.shape)
                    and (global_batch, global_tokens) == self._g_norm and max_len in (None, int(txt.shape[1]))):
                self._g_wave.copy_(wave, non_blocking=True)
                self._g_txt.copy_(txt, non_blocking=True)
                self._g_len.copy_(torch.as_tensor(wave_len), non_blocking=True)
                self.graph.replay()
                self.step_id += 1
                return self._g_loss
        return self._eager(wave, wave_len, txt, global_batch, global_tokens, max_len)

    def _eager(self, wave, wave_len, txt, global_batch, global_tokens, max_len):
        model = self.model
        tf_rate = self.optimizer.pre_step(self.step_id)
        feat, feat_len = self.front_end(wave, wave_len)
        txt_len = torch.sum(txt != 0, dim=-1)
        if max_len is None:
            max_len = txt.shape[1] if global_batch is not None else int(txt_len.max())
        ctc_output, encode_len, att_output, att_align, _ = model(feat, feat_len, max_len, tf_rate=tf_rate, teacher=txt)
        total = 0
        ctc = att = None
        if ctc_output is not None:
            self.ctc_loss.global_batch = global_batch
            ctc = self.ctc_loss(ctc_output.transpose(0, 1), txt, encode_len, txt_len)
            total = total + ctc * model.ctc_weight
        if att_output is not None:
            b, t, _ = att_output.shape
            tgt = txt[:, :t].reshape(-1)
            if global_tokens is None:
                att = ops.cross_entropy(att_output.reshape(b * t, -1), tgt, ignore_index=0)
            else:
                att = ops.cross_entropy(att_output.reshape(b * t, -1), tgt, ignore_index=0, reduction="sum") / global_tokens
            total = total + att * (1 - model.ctc_weight)
        total.backward()
        if model.enable_att:     # drop the step's cached keys / alignments / decoder state (they pin the autograd graph)
            model.attention.reset_mem()
            model.decoder.hidden_state = None
            model.decoder._dw = None
        grad_norm = self.optimizer.step()
        self.step_id += 1
        det = lambda t: t.detach() if t is not None else None     # keep no reference to the autograd graph
        self.last = {"ctc": det(ctc), "att": det(att), "total": total.detach(), "grad_norm": grad_norm,
                     "ctc_output": det(ctc_output), "att_output": det(att_output), "encode_len": encode_len}
        return total.detach()

    def release_graph(self):
        """Drop the captured step.  Under data parallelism call this BEFORE the process group is destroyed: the graph
        holds NCCL kernels of that communicator."""
        if self.graph is not None:
            torch.cuda.synchronize()
            self.graph.reset()
            self.graph = None
            self._g_loss = None

    def capture(self, wave, wave_len, txt, global_batch=None, global_tokens=None, warmup=3):
        """Capture the WHOLE train step (front end, forward, losses, backward, all-reduce, clip + update) into one
        CUDA graph for fixed shapes: ~130 (cfg B) to ~2000 (cfg C: the decode loop) launches become one replay, which
        removes the host launch gaps between the small kernels.  Valid for a fixed learning rate, tf_rate = 1 and
        Adadelta (the update kernel's scalars are frozen into the graph).  Returns True when the graph is in use."""
        h = self.config["hparas"]
        if h["lr_scheduler"] not in ("fixed", None) or h["optimizer"] != "Adadelta" or self.optimizer.tf_rate(0) != 1:
            self.graph_error = "schedule-dependent scalars: eager mode"
            return False
        try:
            import gc
            self.last = {}
            gc.collect()                 # drop autograd nodes created on another stream by earlier eager steps
            self._g_wave, self._g_txt = wave.clone(), txt.clone()
            self._g_len = torch.as_tensor(wave_len).to(wave.device).clone()
            self._g_norm = (global_batch, global_tokens)
            max_len = int(txt.shape[1])
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    self._eager(self._g_wave, self._g_len, self._g_txt, global_batch, global_tokens, max_len)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            # data parallel: the bucketed NCCL all-reduces are captured with the step (forks / joins of the capture
            # stream).  NCCL's own threads may call the CUDA API meanwhile, so only THIS thread's calls are policed.
            mode = "thread_local" if self.dp.enabled else "global"
            with torch.cuda.graph(graph, capture_error_mode=mode):
                self._g_loss = self._eager(self._g_wave, self._g_len, self._g_txt, global_batch, global_tokens, max_len)
            self.graph = graph
            return True
        except Exception as e:  # stay on the eager path, say why
            self.graph = None
            self.graph_error = "%s: %s" % (type(e).__name__, e)
            torch.cuda.synchronize()
            return False
