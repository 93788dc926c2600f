"""Small driver for ncu: one BiLSTM layer forward+backward at the cfg-B/C shape (B=64, H=512) with a short T."""
import importlib, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("end-to-end-asr-pytorch_b200")
B, T, I, H = 64, int(os.environ.get("T", "64")), 1024, int(os.environ.get("H", "512"))
B = int(os.environ.get("B", B))
torch.manual_seed(0)
ref = torch.nn.LSTM(I, H, bidirectional=True, batch_first=True)
params = [p.detach().cuda().requires_grad_(True) for p in ref.parameters()]
x = torch.randn(B, T, I, device="cuda", requires_grad=True)
pkg.lib.load().b200asr_debug_set_lstm_mode(int(os.environ.get("MODE", "0")))
for it in range(3):
    y = pkg.ops.bilstm(x, params, 2)
    y.backward(torch.ones_like(y))
torch.cuda.synchronize()
print("done")
