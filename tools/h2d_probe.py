import torch, time, numpy as np
n = 16*100*2048*2048
host = torch.empty(n, dtype=torch.uint16, pin_memory=True)
host.zero_()
dev = torch.empty(n, dtype=torch.uint16, device="cuda")
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter()-t0)/reps
one = t(lambda: dev.copy_(host, non_blocking=True))
print("single copy GB/s", n*2/one/1e9)
hv = host.view(16,-1); dv = dev.view(16,-1)
per = t(lambda: [dv[b].copy_(hv[b], non_blocking=True) for b in range(16)])
print("16 copies GB/s", n*2/per/1e9)
hn = host.numpy().reshape(16,-1)
per2 = t(lambda: [dv[b].copy_(torch.from_numpy(hn[b]), non_blocking=True) for b in range(16)])
print("16 copies via from_numpy GB/s", n*2/per2/1e9, torch.from_numpy(hn[0]).is_pinned())
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def two():
    for b in range(16):
        with torch.cuda.stream(s1 if b % 2 else s2):
            dv[b].copy_(hv[b], non_blocking=True)
per3 = t(two)
print("2 streams GB/s", n*2/per3/1e9)
