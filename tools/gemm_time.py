import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from bcnf_b200 import train, _cabi
dev='cuda:0'
def t(fn, reps=200):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps*1e3
B=256
for (K,N) in [(526,526),(1370,526),(526,18)]:
    x=torch.randn(B,K,device=dev); w=torch.randn(N,K,device=dev)/30; b=torch.randn(N,device=dev)
    out=torch.empty(B,N,device=dev); pre=torch.empty(B,N,device=dev)
    d=torch.randn(B,N,device=dev); din=torch.empty(B,K,device=dev); dw=torch.empty(N,K,device=dev); savedK=torch.randn(B,K,device=dev)
    f=lambda: train._gemm(x,(K,1),w,(1,K),out,B,N,K,epi=_cabi.EPI_BIAS_GELU_DROP,bias=b,save=pre,seed=1,uid=1,p=0.4)
    g=lambda: train._gemm(d,(N,1),w,(K,1),din,B,K,N,epi=_cabi.EPI_DGELU_DROP,saved=savedK,seed=1,uid=1,p=0.4)
    h=lambda: train._gemm(d,(1,N),x,(K,1),dw,N,K,B)
    tf=lambda: torch.nn.functional.gelu(torch.addmm(b,x,w.t()))
    print(f"K={K} N={N}: fwd {t(f):.1f} us  dgrad {t(g):.1f} us  wgrad {t(h):.1f} us | torch addmm+gelu {t(tf):.1f} us  torch dgrad mm {t(lambda: d@w):.1f} us  torch wgrad mm {t(lambda: d.t()@x):.1f} us")
