import importlib, sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("end-to-end-asr-pytorch_b200")
def se(a,b): 
    a=a.double().cpu(); b=b.double().cpu(); return float((a-b).abs().max()/b.abs().max())
for (B,T,I,H) in [(64,10,120,512),(64,40,1024,512)]:
    torch.manual_seed(3); ref=torch.nn.LSTM(I,H,bidirectional=True,batch_first=True)
    torch.manual_seed(4); x=torch.randn(B,T,I)
    r64=torch.nn.LSTM(I,H,bidirectional=True,batch_first=True).double(); r64.load_state_dict({k:v.double() for k,v in ref.state_dict().items()})
    x64=x.double().requires_grad_(True); y64,_=r64(x64); gy=torch.randn(B,T,2*H); y64.backward(gy.double())
    x32=x.clone().requires_grad_(True); y32,_=ref(x32); y32.backward(gy)
    print("shape",(B,T,I,H))
    print(" aten-fp32-cpu vs fp64: y %.2e dx %.2e dWih %.2e dWhh %.2e"%(se(y32,y64),se(x32.grad,x64.grad),se(ref.weight_ih_l0.grad,r64.weight_ih_l0.grad),se(ref.weight_hh_l0.grad,r64.weight_hh_l0.grad)))
    for mode in ("fp32","tf32x3"):
        pkg.ops.GEMM_MODE=mode
        params=[p.detach().clone().cuda().requires_grad_(True) for p in ref.parameters()]
        xg=x.cuda().requires_grad_(True); y=pkg.ops.bilstm(xg,params,2); y.backward(gy.cuda())
        g64=[p.grad for p in r64.parameters()]
        print(" %-7s vs fp64: y %.2e dx %.2e dWih %.2e dWhh %.2e db %.2e | vs aten32: y %.2e dx %.2e"%(mode,se(y,y64),se(xg.grad,x64.grad),se(params[0].grad,g64[0]),se(params[1].grad,g64[1]),se(params[2].grad,g64[2]),se(y,y32),se(xg.grad,x32.grad)))
