"""Host-side bandwidth diagnosis for the e2e path (never a bench number): what the box gives for
pinned H2D / D2H, pageable H2D, and host memcpy with 1..16 threads; plus the CPU allowance."""
import os
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch


def cat(path):
    try:
        return open(path).read().strip()
    except OSError as e:
        return f'<{e.__class__.__name__}>'


print('nproc', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))
print('cgroup cpu.max:', cat('/sys/fs/cgroup/cpu.max'))
print('cgroup cpu.stat:', cat('/sys/fs/cgroup/cpu.stat').replace('\n', ' | '))
print('THP:', cat('/sys/kernel/mm/transparent_hugepage/enabled'))
print('numa nodes:', cat('/sys/devices/system/node/online'))

N = 256 << 20
dev = torch.device('cuda', 0)
d = torch.empty(N, dtype=torch.uint8, device=dev)
pin = torch.empty(N, dtype=torch.uint8).pin_memory()
pin.fill_(1)
page = torch.empty(N, dtype=torch.uint8)
page.fill_(2)
torch.cuda.synchronize()


def timed(fn, reps=4):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


print(f'pinned   H2D {N / timed(lambda: d.copy_(pin, non_blocking=True)) / 1e9:6.1f} GB/s')
print(f'pinned   D2H {N / timed(lambda: pin.copy_(d, non_blocking=True)) / 1e9:6.1f} GB/s')
print(f'pageable H2D {N / timed(lambda: d.copy_(page)) / 1e9:6.1f} GB/s')
print(f'pageable D2H {N / timed(lambda: page.copy_(d)) / 1e9:6.1f} GB/s')

src = np.ones(N, np.uint8)
dst = np.zeros(N, np.uint8)
pin_np = pin.numpy()
for nt in (1, 2, 4, 8, 16):
    sl = N // nt
    with ThreadPoolExecutor(nt) as ex:
        def job(i, a=dst, b=src):
            a[i * sl:(i + 1) * sl] = b[i * sl:(i + 1) * sl]
        best = 1e9
        for _ in range(4):
            t0 = time.perf_counter()
            list(ex.map(job, range(nt)))
            best = min(best, time.perf_counter() - t0)
        print(f'host memcpy pageable->pageable {nt:2d} threads {N / best / 1e9:6.1f} GB/s')
        best = 1e9
        for _ in range(4):
            t0 = time.perf_counter()
            list(ex.map(lambda i: job(i, pin_np, src), range(nt)))
            best = min(best, time.perf_counter() - t0)
        print(f'host memcpy pageable->pinned   {nt:2d} threads {N / best / 1e9:6.1f} GB/s')
print('cgroup cpu.stat after:', cat('/sys/fs/cgroup/cpu.stat').replace('\n', ' | '))
