import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from f2nerf_b200._lib import call, stream
DEV="cuda"
def eq(a,b): return int((a.view(torch.int32)!=b.view(torch.int32)).sum())
gen=torch.Generator(device=DEV).manual_seed(3)
n=1<<20
p0=(torch.rand(n,device=DEV,generator=gen)*2-1)*1e-2
pa,pf=p0.clone(),p0.clone()
ma,va,mf,vf=(torch.zeros(n,device=DEV) for _ in range(4))
for step in range(1,4):
    g=torch.randn(n,device=DEV,generator=gen)*1e-4
    g[torch.rand(n,device=DEV,generator=gen)<0.7]=0.0
    lr=1e-2*(0.5+0.1*step)
    b1,b2,eps=0.9,0.99,1e-15
    bc1,bc2=1-b1**step,1-b2**step
    ma.mul_(b1).add_(g,alpha=1-b1)
    va.mul_(b2).addcmul_(g,g,value=1-b2)
    sq=va.sqrt()
    d1=sq/(bc2**0.5)
    den=d1.clone().add_(eps)
    q=ma/den
    pa_before=pa.clone()
    pa.addcdiv_(ma,den,value=-(lr/bc1))
    call("f2b_adam_step", pf, g, mf, vf, n, n, lr, b1, b2, eps, 0.0, step, None, stream())
    bad=(pa.view(torch.int32)!=pf.view(torch.int32))
    print("step",step,"m",eq(ma,mf),"v",eq(va,vf),"p",int(bad.sum()))
    if bad.any():
        i=int(bad.nonzero()[0])
        import math
        print(" idx",i,"g",float(g[i]),"m",float(ma[i]),"v",float(va[i]),"sq",float(sq[i]),"den",float(den[i]),"q",float(q[i]),"p_before",float(pa_before[i]),"pa",float(pa[i]),"pf",float(pf[i]))
        print(" inv py", (torch.tensor(1.0)/torch.tensor(bc2**0.5,dtype=torch.float32)).item(), "inv c", 1.0/float(torch.tensor(math.sqrt(bc2),dtype=torch.float32)), "negstep", -(lr/bc1), float(torch.tensor(-(lr/bc1),dtype=torch.float32)))
        # emulate kernel stages in torch with float scalars
        inv=(torch.tensor(1.0,device=DEV)/torch.tensor(math.sqrt(bc2),dtype=torch.float32,device=DEV))
        den2=sq*inv+torch.tensor(eps,dtype=torch.float32,device=DEV)
        print(" den vs den2 mism", eq(den,den2), " zero-den count", int((den==0).sum()), "tiny den", int((den<1e-12).sum()))
