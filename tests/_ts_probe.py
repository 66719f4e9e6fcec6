import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
import deepsc_gan_b200
from deepsc_gan_b200 import _lib as L
dev=torch.device('cuda:0')
g=torch.Generator().manual_seed(0)
M,K,N=300,128,256
x=torch.randn(M,K,generator=g); w=torch.randn(K,N,generator=g)/np.sqrt(K); b=torch.randn(N,generator=g)
ref=(x.double()@w.double()+b.double())
for prec in (16+2, 32+16+2, 16+1, 32+16+1):
    y=torch.zeros(M,N,device=dev)
    L.linear(x.to(dev), w.to(dev), b.to(dev), out=y, prec=prec)
    torch.cuda.synchronize()
    err=float((y.cpu().double()-ref).abs().max()/ref.abs().max())
    print('prec',prec,'swap',bool(prec&32),'rel err',err, flush=True)
