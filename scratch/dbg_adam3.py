import torch, math
DEV="cuda"
gen=torch.Generator(device=DEV).manual_seed(1)
x=torch.rand(1<<18,device=DEV,generator=gen)*1e-3
def eq(a,b): return int((a.view(torch.int32)!=b.view(torch.int32)).sum())
res={"A_double_inv":0,"B_float_inv":0,"C_true_div":0}
for step in range(1,60):
    s=math.sqrt(1-0.99**step)
    y=x/s
    invA=torch.tensor(1.0/s,dtype=torch.float32,device=DEV)
    invB=torch.tensor(1.0,dtype=torch.float32,device=DEV)/torch.tensor(s,dtype=torch.float32,device=DEV)
    res["A_double_inv"]+=eq(y,x*invA); res["B_float_inv"]+=eq(y,x*invB); res["C_true_div"]+=eq(y,x/torch.tensor(s,dtype=torch.float32,device=DEV))
print(res)
