import sys; sys.path.insert(0,"/root/repo")
import torch, deepsc_gan_b200
from deepsc_gan_b200 import _lib as L
dev=torch.device("cuda:0")
for M in (40011, 128*296, 128*300):
  for N in (256,):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(M, 128, generator=g).to(dev); w=(torch.randn(128, N, generator=g)*0.1).to(dev); b=torch.randn(N, generator=g).to(dev)
    y = torch.full((M, N), float("nan"), device=dev); yt=torch.empty((M,N),device=dev)
    L.linear(x, w, b, act=1, n=N, prec=1, out=y); L.linear(x, w, b, act=1, n=N, prec=65, out=yt)
    torch.cuda.synchronize()
    ref=(x.double()@w.double()+b.double()).clamp_min(0)
    bad=(y!=yt)
    print(M,N,"nan",int(torch.isnan(y).sum()),"mismatch elems",int(bad.sum()),"rows",int(bad.any(1).sum()), "err vs ref", float((y.double()-ref).abs().max()), "tiled err", float((yt.double()-ref).abs().max()))
    if bad.any():
        r=bad.any(1).nonzero().flatten(); print(" first bad rows", r[:10].tolist(), "last", r[-5:].tolist(), "tiles", sorted(set((r//128).tolist()))[:20])
        c=bad.any(0).nonzero().flatten(); print(" bad cols", c[:8].tolist(), len(c))
