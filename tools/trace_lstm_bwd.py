"""clock64 timeline of one CTA of the tcgen05 BiLSTM BACKWARD kernel (csrc/lstm_umma.cu), averaged over the steps.
   env: B (64), H (512), T (300), I (120).  Slots: 0 step start | 3 inbox landed | 4 pointwise done | 5 A tiles
   released | 1 MMA lane: A ready | 2 MMA lane: commits issued | 6 drain: n-block 0 complete | 7 drain stored |
   8 publisher woke | 9 fence + red issued | 10 control: all sources published | 11 control: inbox copy issued."""
import importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("end-to-end-asr-pytorch_b200")
B, H, T, I = (int(os.environ.get(k, d)) for k, d in (("B", 64), ("H", 512), ("T", 300), ("I", 120)))
torch.manual_seed(0)
ref = torch.nn.LSTM(I, H, bidirectional=True, batch_first=True)
params = [p.detach().cuda().requires_grad_(True) for p in ref.parameters()]
x = torch.randn(B, T, I, device="cuda", requires_grad=True)
gy = torch.randn(B, T, 2 * H, device="cuda")
lib = pkg.lib.load()
lib.b200asr_debug_set_lstm_mode(int(os.environ.get("MODE", 0)) | 128)    # flag bit 3: trace the backward kernel
for _ in range(2):
    pkg.ops.bilstm(x, params, 2).backward(gy)
tr = torch.zeros(T, 16, dtype=torch.int64, device="cuda")
y = pkg.ops.bilstm(x, params, 2)
lib.b200asr_debug_set_lstm_trace(pkg.lib.ptr(tr))
y.backward(gy)
torch.cuda.synchronize()
lib.b200asr_debug_set_lstm_trace(None)
lib.b200asr_debug_set_lstm_mode(0)
t = tr.cpu().double()
lo, hi = 5, T - 6
print("B=%d H=%d T=%d  step period %.0f cycles" % (B, H, T, (t[lo + 1:hi + 1, 0] - t[lo:hi, 0]).mean()))
for slot, name in [(3, "inbox landed"), (4, "sum + pointwise done"), (5, "A tiles written, released"),
                   (1, "MMA lane: A ready"), (2, "MMA lane: all commits issued"), (6, "drain: n-block 0 complete"),
                   (7, "drain: stores issued, arrived on pub"), (8, "publisher woke"), (9, "publisher: fence + red issued"),
                   (10, "control 0: all sources published"), (11, "control 0: inbox copy issued")]:
    print("  t%-2d - t0   %7.0f   %s" % (slot, (t[lo:hi, slot] - t[lo:hi, 0]).mean(), name))
print("  t0[s+1]-t0  %6.0f   next step start" % (t[lo + 1:hi + 1, 0] - t[lo:hi, 0]).mean())
print("  t3[s+1]-t0  %6.0f   next inbox landed" % (t[lo + 1:hi + 1, 3] - t[lo:hi, 0]).mean())
