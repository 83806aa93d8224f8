import torch
DEV="cuda"
gen=torch.Generator(device=DEV).manual_seed(3)
n=1<<20
m=(torch.rand(n,device=DEV,generator=gen)*2-1)*1e-3
v=torch.rand(n,device=DEV,generator=gen)*1e-6
p=(torch.rand(n,device=DEV,generator=gen)*2-1)*1e-2
g=torch.randn(n,device=DEV,generator=gen)*1e-3
f32=lambda x: x.to(torch.float32)
def fma(a,b,c): return (a.double()*b.double()+c.double()).float()
def eq(a,b): return float((a.view(torch.int32)!=b.view(torch.int32)).float().mean())
b1,b2=0.9,0.99
# stage 1
m1=m.clone().mul_(b1)
print("mul_ == m*float(b1):", eq(m1, m*torch.tensor(b1,dtype=torch.float32,device=DEV)))
m2=m1.clone().add_(g,alpha=1-b1)
a=torch.tensor(1-b1,dtype=torch.float32,device=DEV)
print("add_ alpha: fma?", eq(m2, fma(a.expand(n),g,m1)), " separate?", eq(m2, m1+(g*a)))
# stage 2
v1=v.clone().mul_(b2)
c=torch.tensor(1-b2,dtype=torch.float32,device=DEV)
v2=v1.clone().addcmul_(g,g,value=1-b2)
gg=g*g
print("addcmul: fma(c, g*g, v1)?", eq(v2, fma(c.expand(n),gg,v1)), " (c*g)*g + v1 fma?", eq(v2, fma(c*g,g,v1)), " separate?", eq(v2, v1+c*gg), " sep2", eq(v2, v1+(c*g)*g))
# stage 3
bc2=1-b2**3
sq=v2.sqrt()
d1=sq/(bc2**0.5)
inv=torch.tensor(1.0,dtype=torch.float32,device=DEV)/torch.tensor(bc2**0.5,dtype=torch.float32,device=DEV)
print("div scalar: mul by inv?", eq(d1, sq*inv), " true div?", eq(d1, sq/torch.tensor(bc2**0.5,dtype=torch.float32,device=DEV)))
d2=d1.clone().add_(1e-15)
print("add eps:", eq(d2, d1+torch.tensor(1e-15,dtype=torch.float32,device=DEV)))
# stage 4
bc1=1-b1**3
alpha=-(1e-2/bc1)
p2=p.clone().addcdiv_(m2,d2,value=alpha)
al=torch.tensor(alpha,dtype=torch.float32,device=DEV)
q=m2/d2
print("addcdiv: fma(al, m/d, p)?", eq(p2, fma(al.expand(n),q,p)), " separate?", eq(p2, p+al*q), " (al*m)/d + p ?", eq(p2, p+(al*m2)/d2), " fma((al*m),1/d..)", eq(p2, fma(al*m2, 1/d2, p)))
