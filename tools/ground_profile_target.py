"""Developer helper: one device-resident launch of the ground-contact forward kernel (for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import redmax_b200 as rb  # noqa: E402

if __name__ == '__main__':
    gz = float(sys.argv[1]) if len(sys.argv) > 1 else -40.0
    h = float(sys.argv[2]) if len(sys.argv) > 2 else 2e-4
    B, nsteps = 4096, 100
    sc = rb.chain_scene(32, ground=True, h=h, nsteps=nsteps, ground_z=gz)
    sc.init()
    q0, qd0 = rb.synthetic_inputs(sc, B, seed=20260003)
    dq0, dqd0 = torch.from_numpy(q0).cuda(), torch.from_numpy(qd0).cuda()
    qo = torch.empty((B, nsteps, sc.nr), dtype=torch.float64, device='cuda')
    qdo = torch.empty_like(qo)
    st = torch.empty(B, dtype=torch.int32, device='cuda')
    it = torch.empty((B, 2), dtype=torch.int32, device='cuda')
    for rep in range(3):
        sc.rollout_dev(dq0, dqd0, qo, qdo, st, it, scheme=2, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    print('done', int(st.max()), float(it[:, 0].double().mean()) / nsteps)
