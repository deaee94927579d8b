import os, sys, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motiondiffusion_moe_b200 import ops
from oracle import motion_oracle as mo
print("allow_tf32 matmul", torch.backends.cuda.matmul.allow_tf32, "cudnn", torch.backends.cudnn.allow_tf32, "prec", torch.get_float32_matmul_precision())
DEV="cuda"
def rel(a,b): return ((a.double()-b.double()).norm()/b.double().norm()).item()
B,H,T,hd=3,4,196,128; D=H*hd
g=torch.Generator().manual_seed(1)
pm=torch.randn(hd,256,generator=g); q_,_=torch.linalg.qr(pm,mode="reduced")
P=(F.normalize(q_,dim=0)*hd**-0.25); nw=(1+0.1*torch.randn(hd,generator=g)); nb=0.05*torch.randn(hd,generator=g)
qkv=torch.randn(B*T,3*D,generator=torch.Generator().manual_seed(2))*2
length=torch.tensor([T,T//3,2])
def oracle(dev, dt):
    p={"fa.projection_matrix":P.to(dev,dt),"fa.norm.weight":nw.to(dev,dt),"fa.norm.bias":nb.to(dev,dt)}
    q,k,v=(t.view(B,T,H,hd).permute(0,2,1,3)*0.1 for t in qkv.to(dev,dt).view(B,T,3,D).unbind(2))
    return mo.fast_attention(p,"fa",q,k,v,mo.src_mask(T,length.to(dev)).to(dt)).permute(0,2,1,3).reshape(B*T,D)
r64=oracle("cpu",torch.float64)
rc=oracle("cpu",torch.float32); rg=oracle("cuda",torch.float32).cpu()
out=torch.empty(B*T,D,device=DEV)
ops.fastattn(qkv.to(DEV),P.to(DEV),nw.to(DEV),nb.to(DEV),length.to(DEV),0,B,H,T,hd,out)
print("oracle cpu f32 vs f64", rel(rc,r64)); print("oracle cuda f32 vs f64", rel(rg,r64)); print("kernel f32 vs f64", rel(out.cpu(),r64))
ob=torch.empty(B*T,D,device=DEV,dtype=torch.bfloat16)
ops.fastattn(qkv.to(DEV).bfloat16(),P.to(DEV),nw.to(DEV),nb.to(DEV),length.to(DEV),0,B,H,T,hd,ob)
qb=qkv.bfloat16().float()
def oracle_in(qkv_,dev,dt):
    p={"fa.projection_matrix":P.to(dev,dt),"fa.norm.weight":nw.to(dev,dt),"fa.norm.bias":nb.to(dev,dt)}
    q,k,v=(t.view(B,T,H,hd).permute(0,2,1,3)*0.1 for t in qkv_.to(dev,dt).view(B,T,3,D).unbind(2))
    return mo.fast_attention(p,"fa",q,k,v,mo.src_mask(T,length.to(dev)).to(dt)).permute(0,2,1,3).reshape(B*T,D)
rb=oracle_in(qb,"cpu",torch.float64)
print("TC kernel bf16 vs f64(oracle on bf16-rounded inputs)", rel(ob.float().cpu(),rb))
for bi in range(B):
    sl=slice(bi*T,(bi+1)*T)
    print(" seq",bi,"len",int(length[bi]),"rel", rel(ob.float().cpu()[sl],rb[sl]), "live rows", rel(ob.float().cpu()[bi*T:bi*T+int(length[bi])], rb[bi*T:bi*T+int(length[bi])]))
# the per-token (non-common) part
m=rb.view(B,T,D).mean(1,keepdim=True); 
print("signal fraction (token-dependent norm / total)", ((rb.view(B,T,D)-m).norm()/rb.norm()).item())
print("TC err on token-dependent part", (((ob.float().cpu().view(B,T,D)-ob.float().cpu().view(B,T,D).mean(1,keepdim=True))-(rb.view(B,T,D)-m)).norm()/(rb.view(B,T,D)-m).norm()).item())
