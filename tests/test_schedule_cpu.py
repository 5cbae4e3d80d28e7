"""CPU: the load-balancing plan of the forward launches (rmx_debug_schedule, host logic only).  Invariants of McNaughton's
wrap-around rule as the kernel relies on them: every step of every rollout exactly once and in order, at most one cut per
rollout, the waiting part last in its block, the signalling part first in the next block, no two parts of a rollout overlapping
in time, and all blocks within one quota of each other."""
import numpy as np
import pytest


def plan(B, nsteps, slots):
    from redmax_b200 import _ffi
    seg = np.zeros((B + slots, 4), dtype=np.int32)
    off = np.zeros(slots + 1, dtype=np.int32)
    n = _ffi.lib().rmx_debug_schedule(B, nsteps, slots, len(seg), _ffi.ptr(seg), _ffi.ptr(off))
    assert n > 0, _ffi.lib().rmx_last_error()
    return seg[:n], off


@pytest.mark.parametrize('B,nsteps,slots', [(4096, 100, 1184), (3001, 9, 1184), (1185, 2, 1184), (2367, 5, 1184), (5000, 3, 37),
                                            (10, 4, 3), (7, 100, 6), (65536 // 8, 100, 740)])
def test_wraparound_plan_invariants(rb, B, nsteps, slots):
    seg, off = plan(B, nsteps, slots)
    T = -(-B * nsteps // slots)
    assert off[0] == 0 and off[-1] == len(seg) and np.all(np.diff(off) >= 0)
    parts = {}
    start, end, load = {}, {}, []
    for m in range(slots):
        t = 0
        for i in range(off[m], off[m + 1]):
            b, k0, k1, fl = (int(v) for v in seg[i])
            assert 0 <= k0 < k1 <= nsteps
            parts.setdefault(b, []).append((k0, k1, fl))
            start[(b, k0)], end[(b, k1)] = t, t + (k1 - k0)
            t += k1 - k0
            if fl == 1:
                assert i == off[m + 1] - 1          # the waiting part closes its block's list
            if fl == 2:
                assert i == off[m]                  # the signalling part opens its block's list
        load.append(t)
    assert max(load) <= T and sorted(parts) == list(range(B))
    for b, l in parts.items():
        l.sort()
        assert l[0][0] == 0 and l[-1][1] == nsteps and len(l) <= 2
        if len(l) == 2:
            assert l[0][1] == l[1][0] and l[0][2] == 2 and l[1][2] == 1
            assert start[(b, l[1][0])] >= end[(b, l[1][0])]   # the second part starts after the first has finished
        else:
            assert l[0][2] == 0


def test_no_plan_when_every_rollout_has_a_block(rb):
    from redmax_b200 import _ffi
    seg = np.zeros((8, 4), dtype=np.int32)
    off = np.zeros(5, dtype=np.int32)
    assert _ffi.lib().rmx_debug_schedule(4, 10, 4, 8, _ffi.ptr(seg), _ffi.ptr(off)) < 0
