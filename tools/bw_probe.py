import torch, time
dev=torch.device("cuda")
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
for gb in (1.3, 4.0, 16.0):
    n=int(gb*1e9/4)
    x=torch.empty(n,dtype=torch.float32,device=dev); y=torch.empty(n,dtype=torch.float32,device=dev)
    ms=t(lambda: x.fill_(1.0)); print(f"fill {gb} GB: {ms:.3f} ms  {gb*1e9/ms/1e6:.0f} GB/s")
    ms=t(lambda: x.zero_()); print(f"zero(memset) {gb} GB: {ms:.3f} ms  {gb*1e9/ms/1e6:.0f} GB/s")
    ms=t(lambda: y.copy_(x)); print(f"copy {gb} GB: {ms:.3f} ms  {2*gb*1e9/ms/1e6:.0f} GB/s (r+w)")
    ms=t(lambda: x.sum()); print(f"read(sum) {gb} GB: {ms:.3f} ms  {gb*1e9/ms/1e6:.0f} GB/s")
    del x,y
