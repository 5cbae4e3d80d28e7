"""CPU: the N>1 host logic (contiguous batch shards, trajectory gather, objective reduction) with two real ranks over
gloo.  The kernels are not involved: each rank fabricates the 'trajectory' of global rollout b as a deterministic
function of b, so the gathered tensor can be compared with the unsharded one bit for bit (N-GPU
output == 1-GPU output)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_traj(lo, hi, nsteps, nr):
    import torch
    b = torch.arange(lo, hi, dtype=torch.float64).reshape(-1, 1, 1)
    k = torch.arange(nsteps, dtype=torch.float64).reshape(1, -1, 1)
    i = torch.arange(nr, dtype=torch.float64).reshape(1, 1, -1)
    return torch.sin(0.37 * b + 0.11 * k) + 1e-3 * i * b


def _worker(rank, world, port, B, q):
    try:
        sys.path.insert(0, ROOT)
        import torch
        import torch.distributed as dist
        from redmax_b200 import shard
        dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
        nsteps, nr = 5, 3
        lo, hi = shard.shard_bounds(B, world, rank)
        local = _fake_traj(lo, hi, nsteps, nr)
        full = shard.gather_trajectories(local, B=B)
        ok = torch.equal(full, _fake_traj(0, B, nsteps, nr))
        full2 = shard.gather_trajectories(local)  # B inferred with an all-reduce
        ok = ok and torch.equal(full2, full)
        # objective reduction with shared parameters
        Pb = torch.arange(lo, hi, dtype=torch.float64) * 0.5
        Gb = torch.arange(lo, hi, dtype=torch.float64).reshape(-1, 1) * torch.ones(1, nr, dtype=torch.float64)
        P, G = shard.reduce_objective(Pb, Gb)
        tot = float(sum(range(B)))
        ok = ok and abs(float(P) - 0.5 * tot) < 1e-12 and bool(torch.all(G == tot))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, bool(ok), ''))
    except Exception as e:  # pragma: no cover
        q.put((rank, False, repr(e)))


@pytest.mark.parametrize('B', [8, 7])  # equal and ragged shards
def test_two_rank_gather_matches_unsharded(B):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=180) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res


def test_shard_bounds_partition():
    from redmax_b200 import shard
    for B in (0, 1, 5, 4096, 65536 + 3):
        for world in (1, 2, 3, 4, 8):
            cuts = [shard.shard_bounds(B, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == B
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
            sz = [hi - lo for lo, hi in cuts]
            assert max(sz) - min(sz) <= 1
            assert sz == shard.shard_sizes(B, world)
    a = np.arange(10)
    assert shard.take_shard(a, 2, 1).tolist() == [5, 6, 7, 8, 9]
    with pytest.raises(ValueError):
        shard.shard_bounds(4, 2, 2)
