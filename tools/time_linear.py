import sys; sys.path.insert(0,"/root/repo")
import torch, deepsc_gan_b200
from deepsc_gan_b200 import _lib as L
dev=torch.device("cuda:0"); torch.manual_seed(0)
def timed(fn,reps=10,inner=20):
    """median over `reps` of (time of `inner` back-to-back launches / inner): the launch latency of a single call is hidden"""
    for _ in range(3): fn()
    ts=[]
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(inner): fn()
        b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b)*1e3/inner)
    ts.sort(); return ts[len(ts)//2]
for M,K,N in ((73408,128,256),(293632,128,256),(73408,128,512),(73408,512,128),(73408,256,16),(2368,128,384)):
    x=torch.randn(M,K,device=dev); w=torch.randn(K,(N+3)//4*4,device=dev)*0.1; b=torch.randn(N,device=dev); y=torch.empty(M,N,device=dev)
    t0=timed(lambda: L.linear(x,w,b,act=1,n=N,prec=1|64,out=y))
    t=timed(lambda: L.linear(x,w,b,act=1,n=N,prec=1,out=y))
    print(f"   tiled kernel {t0:7.1f} us  {(M*K+M*N)*4/t0/1e3:7.1f} GB/s;  default path:")
    print(f"M={M} K={K} N={N}: {t:7.1f} us  {(M*K+M*N)*4/t/1e3:7.1f} GB/s  {2*M*K*N*3/t/1e6:6.0f} TF/s bf16")
