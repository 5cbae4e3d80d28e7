"""Developer probe: Newton / line-search counts of the sweep kernels (RMX_IMPL=1, or n > 64) against the composite kernels
and the C oracle on the same rollouts, step by step (rmx_rollout_resume pieces), to locate where they stall."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'oracle'))
import oracle_c as oc  # noqa: E402
import redmax_oracle as oracle  # noqa: E402
import redmax_b200 as rb  # noqa: E402

if __name__ == '__main__':
    n = int(sys.argv[1])
    ns = 6
    h = 2e-4
    so = rb.chain_scene(n, h=h, api=oracle)
    so.init()
    sg = rb.chain_scene(n, h=h)
    sg.init()
    B = 4
    q0, qd0 = rb.synthetic_inputs(sg, B, seed=20260006)
    q, qd, st = oc.run_forward_batch(so, 1, q0, qd0, nsteps=ns, threads=4)
    print('n', n, 'RMX_IMPL', os.environ.get('RMX_IMPL'), 'oracle iters', st.tolist())
    for k in range(1, ns + 1):
        out = sg.rollout(q0, qd0, scheme=1, nsteps=k, iterMaxFactor=1)
        print(' nsteps', k, 'gpu iters', out['iters'].tolist(), 'status', out['status'].tolist(),
              'rel err q %.2e' % (np.abs(out['q'] - q[:, :k]).max() / np.abs(q).max()), flush=True)
