"""Per-step timeline of the persistent BiLSTM step kernels (CTA 0, group 0), from in-kernel clock64 stamps.
   env: B (64), H (512), MODE (debug_set_lstm_mode value: 0 default kernels, 1 fp32-FMA kernels, 64 first-generation MMA loops), BWD=1 traces the backward kernel."""
import importlib, sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("end-to-end-asr-pytorch_b200")
L = pkg.lib; lib = L.load()
B, T, I, H = int(os.environ.get("B", 64)), 200, 1024, int(os.environ.get("H", 512))
MODE, BWD = int(os.environ.get("MODE", 0)), int(os.environ.get("BWD", 0))
torch.manual_seed(0)
ref = torch.nn.LSTM(I, H, bidirectional=True, batch_first=True)
params = [p.detach().cuda().requires_grad_(True) for p in ref.parameters()]
x = torch.randn(B, T, I, device="cuda", requires_grad=True)
gy = torch.randn(B, T, 2 * H, device="cuda")
lib.b200asr_debug_set_lstm_mode(MODE | (128 if BWD else 0))     # flag bit 3 (mode bit 7): trace the backward kernel
def run():
    y = pkg.ops.bilstm(x, params, 2)
    if BWD:
        y.backward(gy)
    torch.cuda.synchronize()
run()
tr = torch.zeros(T, 16, dtype=torch.int64, device="cuda")
lib.b200asr_debug_set_lstm_trace(L.ptr(tr))
run()
lib.b200asr_debug_set_lstm_trace(None)
lib.b200asr_debug_set_lstm_mode(0)
t = tr.cpu().numpy().astype(np.float64)
s = slice(20, T - 2)
if BWD:
    names = {1: "inbox arrived", 2: "dG tile done + turn acquired", 3: "GEMM + scatter done", 5: "done-arrive"}
else:
    names = {1: "chunks arrived", 2: "turn acquired (MMA kernels)", 3: "k-loop done",
             4: "K-chunk reduction done (FMA kernels)", 5: "stores issued + done-arrive"}
names.update({8: "ctl: done seen", 9: "ctl: after fence", 10: "ctl: peers' counter reached", 11: "ctl: copies issued"})
base = t[s, 0]
print("%s kernel, mode %d: cycles relative to the step start (mean over steps %d..%d), SM clock ~1.9 GHz" % (
    "backward" if BWD else "forward", MODE, s.start, T - 2))
for k in sorted(names):
    if np.all(t[s, k] == 0):
        continue
    print("  %-38s %8.0f" % (names[k], float(np.mean(t[s, k] - base))))
print("  %-38s %8.0f" % ("step period", float(np.mean(np.diff(t[s, 0])))))
print("  %-38s %8.0f" % ("next step's block arrives at", float(np.mean(t[s.start + 1:T - 1, 1] - base))))
