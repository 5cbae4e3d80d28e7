"""Developer probe: device-to-host copies into PAGEABLE memory from 1, 2, 4, 8 host threads (one stream each), aggregate GB/s --
is the blocking part of rmx_rollout's pageable path (what a MATLAB mxArray caller gets) worth spreading over threads?"""
import threading
import time

import numpy as np
import torch

if __name__ == '__main__':
    n = 210 * 1024 * 1024 // 8
    src = torch.empty(n, dtype=torch.float64, device='cuda').normal_()
    dst = torch.from_numpy(np.empty(n, dtype=np.float64))   # pageable
    dst.fill_(0)
    for nt in (1, 2, 4, 8):
        streams = [torch.cuda.Stream() for _ in range(nt)]
        best = 1e9
        for rep in range(3):
            torch.cuda.synchronize()
            def work(i):
                lo, hi = n * i // nt, n * (i + 1) // nt
                with torch.cuda.stream(streams[i]):
                    dst[lo:hi].copy_(src[lo:hi], non_blocking=False)
                streams[i].synchronize()
            t0 = time.perf_counter()
            th = [threading.Thread(target=work, args=(i,)) for i in range(nt)]
            [t.start() for t in th]
            [t.join() for t in th]
            best = min(best, time.perf_counter() - t0)
        print('%d thread(s): %.1f ms for 210 MiB  = %.1f GB/s' % (nt, best * 1e3, n * 8 / best / 1e9), flush=True)
    pin = torch.empty(n, dtype=torch.float64).pin_memory()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); pin.copy_(src, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    print('page-locked destination: %.1f ms = %.1f GB/s' % ((t1 - t0) * 1e3, n * 8 / (t1 - t0) / 1e9))
    for nt in (1, 4, 8, 16):
        best = 1e9
        a = pin.numpy(); b = dst.numpy()
        for rep in range(3):
            def cp(i):
                lo, hi = n * i // nt, n * (i + 1) // nt
                np.copyto(b[lo:hi], a[lo:hi])
            t0 = time.perf_counter()
            th = [threading.Thread(target=cp, args=(i,)) for i in range(nt)]
            [t.start() for t in th]
            [t.join() for t in th]
            best = min(best, time.perf_counter() - t0)
        print('host memcpy page-locked -> pageable, %2d thread(s): %.1f ms = %.1f GB/s' % (nt, best * 1e3, n * 8 / best / 1e9), flush=True)
