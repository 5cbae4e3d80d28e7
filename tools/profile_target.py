"""Developer tool: the smallest program that launches one workload's kernels, for ncu captures (never a bench number).
    ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <skip> -c 1 -o gpurun_out/prof \
        python tools/profile_target.py <workload> [launches]
Runs `launches` (default 2) device-resident passes of bench.py's workload: the first warms up, the later ones are for -s/-c."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import redmax_b200 as rb  # noqa: E402

if __name__ == '__main__':
    name = sys.argv[1]
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    w = bench.WORKLOADS[name]
    B, ns = w['B'], w['nsteps']
    sc = bench.make_scene(name)
    nr = sc.nr
    dev = torch.device('cuda', 0)
    st = torch.empty(B, dtype=torch.int32, device=dev)
    if w['kind'] == 'adjoint':
        p, xt = bench.adjoint_inputs(sc, B, bench.SEED)
        dq0 = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(sc.qInit, (B, nr)))).to(dev)
        dqd0 = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(sc.qdotInit, (B, nr)))).to(dev)
        dp, dxt = torch.from_numpy(p).to(dev), torch.from_numpy(xt).to(dev)
        dP = torch.empty(B, dtype=torch.float64, device=dev)
        dG = torch.empty((B, nr), dtype=torch.float64, device=dev)
        for _ in range(reps):
            sc.rollout_adjoint_dev(dq0, dqd0, dp, dxt, dP, dG, st)
    else:
        q0, qd0 = rb.synthetic_inputs(sc, B, seed=bench.SEED)
        dq0, dqd0 = torch.from_numpy(q0).to(dev), torch.from_numpy(qd0).to(dev)
        qo = torch.empty((B, ns, nr), dtype=torch.float64, device=dev)
        qdo = torch.empty_like(qo)
        it = torch.empty((B, 2), dtype=torch.int32, device=dev)
        for _ in range(reps):
            sc.rollout_dev(dq0, dqd0, qo, qdo, st, it, scheme=w['scheme'])
    torch.cuda.synchronize()
    print('profile_target', name, 'status!=0:', float((st != 0).float().mean()))
