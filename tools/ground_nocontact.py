"""Developer probe: what the external-force kernel variant costs when no force acts -- the C3 workload with the ground plane
far below the chain (no corner ever touches it), against the plain kernel on the same dynamics."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import redmax_b200 as rb  # noqa: E402
from explore_r2 import run  # noqa: E402

if __name__ == '__main__':
    print(torch.cuda.get_device_name(0), flush=True)
    for label, kw in (('plain kernel', dict()), ('ground kernel, plane at z = -1e6 (no contact)', dict(ground=True, ground_z=-1e6)),
                      ('ground kernel, plane at z = -40 (C3)', dict(ground=True, ground_z=-40.0))):
        sc = rb.chain_scene(32, h=2e-4, nsteps=100, **kw)
        sc.init()
        ms, it, st = run(sc, 4096, 2, 100)
        print('%-50s %8.2f ms  newton/step %.2f  ls/step %.2f  status!=0 %.2f%%' % (label, ms, it[:, 0].mean() / 100, it[:, 1].mean() / 100,
                                                                                  100.0 * (st != 0).mean()), flush=True)
