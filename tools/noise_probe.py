"""Developer probe: absolute error of the GPU residual g near a converged point (where |g| ~ 1e-10 and its terms are O(1)),
for the composite kernels (impl 2) and the sweep kernels (RMX_IMPL=1), against the dense C oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'oracle'))
import oracle_c as oc  # noqa: E402
import redmax_oracle as oracle  # noqa: E402
import redmax_b200 as rb  # noqa: E402

if __name__ == '__main__':
    impl = os.environ.get('RMX_IMPL', '2')
    for n in (32, 40, 64, 72, 100):
        if n > 64 and impl != '1':
            pass
        h = 2e-4
        so = rb.chain_scene(n, h=h, api=oracle)
        so.init()
        sg = rb.chain_scene(n, h=h)
        sg.init()
        q0, qd0 = rb.synthetic_inputs(so, 2, seed=20260006)
        for b in range(2):
            a0, ad0 = q0[b], qd0[b]

            def ev(q):
                r = oc.eval_direct(so, q, (q - a0) / h, q - a0 - h * ad0, h, h * h)
                return r['g'], r['H']
            q = a0 + h * ad0
            for it in range(4):
                g, H = ev(q)
                q = q - np.linalg.solve(H, g)
            g, H = ev(q)
            out = sg.eval(q, (q - a0) / h, q - a0 - h * ad0, h * h, 1.0 / h)
            print('impl %s n %3d b %d: |g_oracle| %.2e |g_gpu| %.2e |g_gpu - g_oracle| %.2e ; rel err H %.1e'
                  % (impl, n, b, np.linalg.norm(g), np.linalg.norm(out['g']), np.linalg.norm(out['g'] - g),
                     np.abs(out['H'] - H).max() / np.abs(H).max()), flush=True)
