import math, sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
from asva_b200 import ops
be = ops.backend()
def rnd(shape, seed, scale=1.0, dtype=torch.bfloat16):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).cuda()
for (M,K,N,bn,sp) in [(384,3840,1280,256,8),(384,3840,1280,128,4),(384,1280,1280,64,2),(384,23040//9*9,1280,128,4)]:
    x, w = rnd((M,K),1), rnd((N,K),2,1/math.sqrt(K))
    b = rnd((N,),3,dtype=torch.float32); r = rnd((M,N),4)
    out = torch.zeros(M,N,dtype=torch.bfloat16,device="cuda")
    spec = ops.spec_linear(x,w,out,bias=b,res0=r)
    spec.block_n, spec.split_k, spec.cta_group, spec.epilogue = bn, sp, 1, 4
    print("plan", be.gemm_plan(spec))
    be.gemm(spec); torch.cuda.synchronize()
    ref = out.clone()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            for _ in range(3):
                be.gemm(spec)
        g.replay(); torch.cuda.synchronize()
        print("capture ok", M,K,N,bn,sp, float((out.float()-ref.float()).abs().max()))
    except Exception as e:
        print("CAPTURE FAILED", M,K,N,bn,sp, str(e)[:300])
