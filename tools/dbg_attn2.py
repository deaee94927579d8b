import os, sys, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motiondiffusion_moe_b200 import ops
from oracle import motion_oracle as mo
DEV="cuda"
def rel(a,b): return ((a.double()-b.double()).norm()/b.double().norm()).item()
for (B,H,T,hd,scale) in [(2,4,196,128,2.0),(2,4,196,64,2.0),(2,4,8,128,2.0),(2,4,196,128,0.2),(2,4,196,128,20.0),(2,4,4,128,2.0)]:
    D=H*hd
    g=torch.Generator().manual_seed(1)
    pm=torch.randn(hd,256,generator=g); q_,_=torch.linalg.qr(pm,mode="reduced")
    P=(F.normalize(q_,dim=0)*hd**-0.25); nw=(1+0.1*torch.randn(hd,generator=g)); nb=0.05*torch.randn(hd,generator=g)
    qkv=torch.randn(B*T,3*D,generator=torch.Generator().manual_seed(2))*scale
    length=torch.tensor([T,max(1,T//3)])
    p={"fa.projection_matrix":P.double(),"fa.norm.weight":nw.double(),"fa.norm.bias":nb.double()}
    q,k,v=(t.view(B,T,H,hd).permute(0,2,1,3)*0.1 for t in qkv.double().view(B,T,3,D).unbind(2))
    r64=mo.fast_attention(p,"fa",q,k,v,mo.src_mask(T,length).double()).permute(0,2,1,3).reshape(B*T,D)
    out=torch.empty(B*T,D,device=DEV)
    ops.fastattn(qkv.to(DEV),P.to(DEV),nw.to(DEV),nb.to(DEV),length.to(DEV),0,B,H,T,hd,out)
    o=out.cpu()
    print((B,H,T,hd,scale),"rel",rel(o,r64)," seq0",rel(o[:T],r64[:T])," head0",rel(o[:, :hd],r64[:, :hd]), " maxabs",(o.double()-r64).abs().max().item())
