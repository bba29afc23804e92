"""PCIe topology + concurrent H2D bandwidth per GPU (one process per GPU, like bench.py --gpus N)."""
import os
import subprocess
import sys
import time

import torch
import torch.multiprocessing as mp


def worker(rank, world, nbytes, barrier, q, affinity):
    if affinity is not None:
        os.sched_setaffinity(0, affinity[rank])
    torch.cuda.set_device(rank)
    host = torch.empty(nbytes // 2, dtype=torch.uint16, pin_memory=True)
    host.zero_()
    dev = torch.empty(nbytes // 2, dtype=torch.uint16, device="cuda")
    dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(3):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    q.put((rank, nbytes / dt / 1e9))


def run(world, affinity=None, label=""):
    ctx = mp.get_context("spawn")
    barrier = ctx.Barrier(world)
    q = ctx.Queue()
    ps = [ctx.Process(target=worker, args=(r, world, 4 << 30, barrier, q, affinity)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get() for _ in ps)
    for p in ps:
        p.join()
    print(label, "world", world, " ".join(f"gpu{r}:{g:.1f}GB/s" for r, g in res), "sum", round(sum(g for _, g in res), 1),
          flush=True)


def gpu_cpu_affinity():
    """{gpu index: set of cpus} from `nvidia-smi topo -m` (CPU Affinity column, e.g. 0-15,32-47)."""
    out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout
    aff = {}
    for line in out.splitlines():
        line = line.replace("\x1b[4m", "").replace("\x1b[0m", "")
        parts = line.split()
        if not parts or not parts[0].startswith("GPU") or not parts[0][3:].isdigit():
            continue
        idx = int(parts[0][3:])
        for tok in parts[1:]:
            if tok[0].isdigit() and all(c.isdigit() or c in "-," for c in tok) and ("-" in tok or "," in tok):
                cpus = set()
                for rng in tok.split(","):
                    a, _, b = rng.partition("-")
                    cpus.update(range(int(a), int(b or a) + 1))
                aff[idx] = cpus
                break
    return aff


if __name__ == "__main__":
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout, flush=True)
    print(subprocess.run(["bash", "-c", "lscpu | grep -i -E 'numa|model name|^CPU\\(s\\)'; nproc; free -g | head -2"],
                         capture_output=True, text=True).stdout, flush=True)
    n = torch.cuda.device_count()
    run(1, label="alone")
    if n >= 2:
        run(n, label="concurrent")
        aff = gpu_cpu_affinity()
        print("affinity", {k: (min(v), max(v), len(v)) for k, v in aff.items()}, flush=True)
        if len(aff) >= n:
            run(n, affinity=[aff[r] for r in range(n)], label="concurrent, NUMA-bound")
