"""Developer probe: LU versus Krylov (projected block-Jacobi BiCGStab) Newton linear solve, device-timed."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import redmax_b200 as rb  # noqa: E402
from redmax_b200 import _ffi  # noqa: E402


def run(n, B, nsteps, h, lin, tol=1e-6):
    sc = rb.chain_scene(n, h=h, nsteps=nsteps)
    sc.init()
    q0, qd0 = rb.synthetic_inputs(sc, B, seed=20260003)
    dq0, dqd0 = torch.from_numpy(q0).cuda(), torch.from_numpy(qd0).cuda()
    qo = torch.empty((B, nsteps, sc.nr), dtype=torch.float64, device='cuda')
    qdo = torch.empty_like(qo)
    st = torch.empty(B, dtype=torch.int32, device='cuda')
    it = torch.empty((B, 2), dtype=torch.int32, device='cuda')
    best = 1e30
    for r in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sc.rollout_dev(dq0, dqd0, qo, qdo, st, it, scheme=1, stream=torch.cuda.current_stream(), linsolve=lin, pcg_tol=tol)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    itc = it.cpu().numpy()
    kry = sc.linsolve_stats() if lin else 0
    print('n=%d B=%d linsolve=%s tol=%g: %.2f ms %.3e steps/s newton/step %.2f krylov/newton %.2f status|=%d'
          % (n, B, 'PCG' if lin else 'LU', tol, best, B * nsteps / (best * 1e-3), itc[:, 0].mean() / nsteps,
             kry / max(1, itc[:, 0].sum()), int(st.max())), flush=True)


if __name__ == '__main__':
    for n, B, ns, h in ((10, 1024, 100, 1e-3), (32, 4096, 100, 1e-3), (64, 4096, 20, 2e-4)):
        run(n, B, ns, h, 0)
        run(n, B, ns, h, _ffi.RMX_LINSOLVE_PCG, 1e-6)
        run(n, B, ns, h, _ffi.RMX_LINSOLVE_PCG, 1e-10)
