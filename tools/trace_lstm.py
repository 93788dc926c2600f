"""Per-step timeline of the persistent BiLSTM forward kernel (CTA 0, group 0), from in-kernel clock64 stamps."""
import importlib, sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("end-to-end-asr-pytorch_b200")
L = pkg.lib; lib = L.load()
B, T, I, H = int(os.environ.get("B", 64)), 200, 1024, 512
torch.manual_seed(0)
ref = torch.nn.LSTM(I, H, bidirectional=True, batch_first=True)
params = [p.detach().cuda() for p in ref.parameters()]
x = torch.randn(B, T, I, device="cuda")
y = pkg.ops.bilstm(x, params, 2); torch.cuda.synchronize()
tr = torch.zeros(T, 16, dtype=torch.int64, device="cuda")
lib.b200asr_debug_set_lstm_trace(L.ptr(tr))
y = pkg.ops.bilstm(x, params, 2); torch.cuda.synchronize()
lib.b200asr_debug_set_lstm_trace(None)
t = tr.cpu().numpy().astype(np.float64)
s = slice(20, T - 2)
def d(a, b, sa=0, sb=0):
    x = t[s, b][sb:] if sb else t[s, b]
    return x
names = {0: "step start", 1: "chunk A arrived", 2: "chunk B arrived", 3: "k-loop done", 4: "red barrier done",
         5: "stores issued + done-arrive", 8: "ctl: done seen", 9: "ctl: after fence", 10: "ctl: peers' counter reached",
         11: "ctl: copies issued"}
base = t[s, 0]
print("cycles relative to the step start (mean over steps %d..%d), SM clock ~1.9 GHz" % (s.start, T - 2))
for k in (1, 2, 3, 4, 5, 8, 9, 10, 11):
    print("  %-32s %8.0f" % (names[k], float(np.mean(t[s, k] - base))))
print("  step period                      %8.0f" % float(np.mean(np.diff(t[s, 0]))))
print("  next step's chunk A arrives at   %8.0f" % float(np.mean(t[s.start + 1:T - 1, 1] - base)))
