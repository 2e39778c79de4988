import torch, numpy as np
dev='cuda'
x=torch.empty(3072*256*78, dtype=torch.float32, device=dev)   # 245 MB
m=torch.randint(0,2,(3072*256*78,),dtype=torch.uint8,device=dev)
big=torch.empty(256<<20,dtype=torch.uint8,device=dev)
def t(fn,n=10):
    fn(); torch.cuda.synchronize(); ts=[]
    for _ in range(n):
        big.zero_()
        s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return np.median(ts)
a=t(lambda: x.zero_()); print('zero 245MB: %.1f us -> %.0f GB/s'%(a*1e3, x.numel()*4/a/1e6))
y=torch.empty_like(x)
a=t(lambda: y.copy_(x)); print('copy 245MB: %.1f us -> %.0f GB/s (r+w)'%(a*1e3, 2*x.numel()*4/a/1e6))
a=t(lambda: m.view(torch.int64).sum()); print('read 61MB: %.1f us -> %.0f GB/s'%(a*1e3, m.numel()/a/1e6))
