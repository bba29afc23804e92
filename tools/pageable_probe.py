"""H2D rate from PAGEABLE host memory (what a datastore returning plain NumPy arrays gives the loader):
direct cudaMemcpy vs a double-buffered pinned staging ring filled by a few host threads."""
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

n = 100 * 2048 * 2048  # one bit volume, uint16
nb = 8
src = [np.random.default_rng(i).integers(0, 4000, n, dtype=np.uint16) for i in range(nb)]
dev = torch.empty((nb, n), dtype=torch.uint16, device="cuda")


def direct():
    for b in range(nb):
        dev[b].copy_(torch.from_numpy(src[b]), non_blocking=True)
    torch.cuda.synchronize()


def staged(chunk_mb=32, slots=4, workers=4):
    chunk = chunk_mb * (1 << 20) // 2
    ring = [torch.empty(chunk, dtype=torch.uint16, pin_memory=True) for _ in range(slots)]
    ring_np = [r.numpy() for r in ring]
    free_evt = [torch.cuda.Event() for _ in range(slots)]
    stream = torch.cuda.current_stream()
    jobs = [(b, o, min(chunk, n - o)) for b in range(nb) for o in range(0, n, chunk)]
    with ThreadPoolExecutor(workers) as ex:
        def fill(slot, job):
            b, o, m = job
            free_evt[slot].synchronize()
            np.copyto(ring_np[slot][:m], src[b][o:o + m])
            return slot, job
        futs = {}
        nxt = 0
        for s in range(min(slots, len(jobs))):
            futs[s] = ex.submit(fill, s, jobs[nxt]); nxt += 1
        done = 0
        slot = 0
        while done < len(jobs):
            s, (b, o, m) = futs[slot].result()
            dev[b, o:o + m].copy_(ring[s][:m], non_blocking=True)
            free_evt[s].record(stream)
            done += 1
            if nxt < len(jobs):
                futs[slot] = ex.submit(fill, slot, jobs[nxt]); nxt += 1
            slot = (slot + 1) % slots
    torch.cuda.synchronize()


def t(fn, reps=2):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


gb = nb * n * 2 / 1e9
print("direct pageable  GB/s", gb / t(direct), flush=True)
for w in (2, 4, 8):
    for c in (16, 64):
        print(f"staged workers={w} chunk={c}MB GB/s", gb / t(lambda: staged(c, 6, w)), flush=True)
ref = torch.from_numpy(np.stack(src)).cuda()
print("staged result identical:", bool(torch.equal(ref.view(torch.int16), dev.view(torch.int16))))
