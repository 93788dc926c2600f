"""clock64 timeline of one CTA of the tcgen05 BiLSTM forward kernel (csrc/lstm_umma.cu), averaged over the steps.
   env: B (64), H (512), T (300), I (120).  Slots: 0 epilogue: step start | 3 epilogue: accumulator complete |
   5 epilogue: h published | 10 control lane 0: producers of K atom 0 done | 11 its bulk copy issued |
   1 MMA lane: atom 0 landed | 6 MMA lane: last atom landed | 2 MMA lane: commit issued."""
import importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("end-to-end-asr-pytorch_b200")
B, H, T, I = (int(os.environ.get(k, d)) for k, d in (("B", 64), ("H", 512), ("T", 300), ("I", 120)))
torch.manual_seed(0)
ref = torch.nn.LSTM(I, H, bidirectional=True, batch_first=True)
params = [p.detach().cuda() for p in ref.parameters()]
x = torch.randn(B, T, I, device="cuda")
lib = pkg.lib.load()
lib.b200asr_debug_set_lstm_mode(int(os.environ.get("MODE", 0)))
with torch.no_grad():
    for _ in range(2):
        pkg.ops.bilstm(x, params, 2)
    tr = torch.zeros(T, 16, dtype=torch.int64, device="cuda")
    lib.b200asr_debug_set_lstm_trace(pkg.lib.ptr(tr))
    pkg.ops.bilstm(x, params, 2)
    torch.cuda.synchronize()
    lib.b200asr_debug_set_lstm_trace(None)
t = tr.cpu().double()
lo, hi = 5, T - 5
period = (t[lo + 1:hi + 1, 0] - t[lo:hi, 0]).mean()
print("B=%d H=%d T=%d  step period %.0f cycles" % (B, H, T, period))
# within step s (s >= 1): reference point = epilogue step start t0[s]
names = [(3, "accumulator complete (epilogue sees mma_done)"), (4, "TMEM loads done, d_free arrived"),
         (7, "pointwise done"), (5, "h stored, arrived on pub")]
for slot, name in names:
    print("  t%-2d - t0   %7.0f   %s" % (slot, (t[lo:hi, slot] - t[lo:hi, 0]).mean(), name))
# exchange of h_s -> MMAs of step s+1, relative to the publish of step s (slot 5)
for slot, name in [(8, "publisher: all epilogue warps arrived"), (9, "publisher: fence + red issued"),
                   (10, "control lane 0: atom 0 producers done"), (11, "control lane 0: bulk copy issued")]:
    print("  t%-2d - t5   %7.0f   %s" % (slot, (t[lo:hi, slot] - t[lo:hi, 5]).mean(), name))
for slot, name in [(15, "control lane NA-1: last atom producers done")]:
    print("  t%-2d - t5   %7.0f   %s" % (slot, (t[lo:hi, slot] - t[lo:hi, 5]).mean(), name))
for slot, name in [(1, "MMA lane: passed atom 0"), (14, "MMA lane: passed atom NA/2"), (6, "MMA lane: passed last atom"),
                   (2, "MMA lane: commit issued"),
                   (3, "epilogue: accumulator complete"), (0, "next step start")]:
    print("  t%-2d[s+1] - t5[s] %7.0f   %s" % (slot, (t[lo + 1:hi + 1, slot] - t[lo:hi, 5]).mean(), name))
