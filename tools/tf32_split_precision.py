"""Effective precision of the 3xTF32 product scheme of csrc/gemm.cu, emulated on the CPU (no GPU): relative error of
K = 4096 dot products against fp64, in units of the sum of |terms|, for IEEE fp32 products, the shipped scheme (raw tile =
truncated hi, residual truncated to TF32 by the tensor core, lo.lo dropped) and the two precision knobs not taken
(residual rounded to nearest by the splitter: cvt.rna.tf32; a fourth product).  Exact accumulation isolates the
product error.  usage: python tools/tf32_split_precision.py"""
import numpy as np
rng=np.random.default_rng(0)
def trunc_tf32(x):
    return (x.view(np.uint32) & np.uint32(0xffffe000)).view(np.float32)
def rna_tf32(x):
    u=x.view(np.uint32).astype(np.uint64)+np.uint64(0x1000)
    return (u.astype(np.uint32) & np.uint32(0xffffe000)).view(np.float32)
K=4096; N=2000
a=rng.standard_normal((N,K)).astype(np.float32); b=rng.standard_normal((N,K)).astype(np.float32)
exact=(a.astype(np.float64)*b.astype(np.float64)).sum(1)
scale=np.abs(a.astype(np.float64)*b.astype(np.float64)).sum(1)      # sum of |terms| (no cancellation) as the yardstick
def three(lo_fn):
    ah,bh=trunc_tf32(a),trunc_tf32(b)
    al,bl=lo_fn(a-ah),lo_fn(b-bh)
    p=ah.astype(np.float64)*bh+al.astype(np.float64)*bh+ah.astype(np.float64)*bl
    return p.sum(1)
def four(lo_fn):
    ah,bh=trunc_tf32(a),trunc_tf32(b)
    al,bl=lo_fn(a-ah),lo_fn(b-bh)
    p=ah.astype(np.float64)*bh+al.astype(np.float64)*bh+ah.astype(np.float64)*bl+al.astype(np.float64)*bl
    return p.sum(1)
f32=(a*b).astype(np.float32).astype(np.float64).sum(1)            # IEEE-rounded products, exact accumulation
for name,v in (("fp32 products (RN)",f32),("3xTF32, lo truncated by the tensor core",three(trunc_tf32)),("3xTF32, lo rounded (cvt.rna)",three(rna_tf32)),
               ("4 products, lo truncated",four(trunc_tf32)),("4 products, lo rounded",four(rna_tf32))):
    e=(v-exact)/scale
    print("%-42s mean %+.2e  rms %.2e   (in units of 2^-24: mean %+.2f rms %.2f)"%(name,e.mean(),np.sqrt((e**2).mean()),e.mean()*2**24,np.sqrt((e**2).mean())*2**24))
