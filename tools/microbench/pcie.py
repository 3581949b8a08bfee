import torch, time
dev='cuda:0'
h_in=torch.empty(45774336//4).pin_memory(); h_out=torch.empty(88510464//4).pin_memory()
d_in=torch.empty_like(h_in,device=dev); d_out=torch.empty_like(h_out,device=dev)
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
def t(fn,n=10):
    fn(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e3
print('H2D 45.8MB ms', t(lambda: d_in.copy_(h_in,non_blocking=True)))
print('D2H 88.5MB ms', t(lambda: h_out.copy_(d_out,non_blocking=True)))
def both():
    with torch.cuda.stream(s1): d_in.copy_(h_in,non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_out,non_blocking=True)
print('both ms', t(both))
