"""Is write-combined pinned memory faster for host -> device DMA on this host?"""
import ctypes as C
import time

import torch

torch.cuda.init()
rt = C.CDLL("libcudart.so.12")
n = 4 << 30
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream().cuda_stream


def bw(ptr):
    rt.cudaMemcpyAsync(C.c_void_p(dev.data_ptr()), C.c_void_p(ptr), C.c_size_t(n), 1, C.c_void_p(stream))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        rt.cudaMemcpyAsync(C.c_void_p(dev.data_ptr()), C.c_void_p(ptr), C.c_size_t(n), 1, C.c_void_p(stream))
    torch.cuda.synchronize()
    return 3 * n / (time.perf_counter() - t0) / 1e9


for name, flags in (("default pinned", 0), ("write-combined", 4), ("portable", 1)):
    p = C.c_void_p()
    rc = rt.cudaHostAlloc(C.byref(p), C.c_size_t(n), C.c_uint(flags))
    assert rc == 0, rc
    C.memset(p, 1, n)
    print(name, round(bw(p.value), 2), "GB/s", flush=True)
    rt.cudaFreeHost(p)
