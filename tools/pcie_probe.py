import torch, time
n=400*1024*1024
d=torch.empty(n,dtype=torch.uint8,device="cuda"); h=torch.empty(n,dtype=torch.uint8).pin_memory()
d2=torch.empty(n//4,dtype=torch.uint8,device="cuda"); h2=torch.empty(n//4,dtype=torch.uint8).pin_memory()
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
def t(fn, reps=8):
    fn(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/reps
dt=t(lambda: h.copy_(d,non_blocking=True)); print("D2H alone GB/s", n/dt/1e9)
dt=t(lambda: d.copy_(h,non_blocking=True)); print("H2D alone GB/s", n/dt/1e9)
def both():
    with torch.cuda.stream(s1): h.copy_(d,non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2,non_blocking=True)
dt=t(both); print("D2H with concurrent H2D(1/4): D2H GB/s", n/dt/1e9)
