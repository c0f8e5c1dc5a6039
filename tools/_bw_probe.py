import torch, time
dev='cuda'
def timeit(fn,n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
N=1<<30
x=torch.empty(N,dtype=torch.bfloat16,device=dev).normal_()
ms=timeit(lambda: x.sum()); print('sum bf16 read-only', N*2/ms/1e6,'GB/s')
xf=x.view(torch.float32)
ms=timeit(lambda: xf.sum()); print('sum f32 read-only', N*2/ms/1e6,'GB/s')
ms=timeit(lambda: xf.max()); print('max f32 read-only', N*2/ms/1e6,'GB/s')
rows=x.view(-1,64)  # 128B rows
idx=torch.randperm(rows.shape[0],device=dev)
out=torch.empty_like(rows)
ms=timeit(lambda: torch.index_select(rows,0,idx,out=out)); print('index_select random 128B rows (r+w)', 2*N*2/ms/1e6,'GB/s')
idx2=torch.arange(rows.shape[0],device=dev)
ms=timeit(lambda: torch.index_select(rows,0,idx2,out=out)); print('index_select sequential rows (r+w)', 2*N*2/ms/1e6,'GB/s')
y=torch.empty_like(x); ms=timeit(lambda: y.copy_(x)); print('copy r+w', 2*N*2/ms/1e6)
