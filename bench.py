#!/usr/bin/env python
"""bench.py -- rollout-steps/s of the batched RedMax stepper on B200 (BASELINE.json metric).

Default workload (`config.workload`): 32-link serial chain, BDF1, h = 1e-3, 100 time steps, 4096 rollouts per GPU with
seeded per-rollout initial states (SURVEY.md section 8(d), C3 shape at the north-star scheme).  One bench "step" is one
pass of the hot path over that batch: one launch of the persistent rollout kernel = 4096 x 100 rollout-steps.  The other
BASELINE.json configurations are selectable with --workload (C2 chain10, C3 chain32 + ground friction BDF2, C4 hand-tree
adjoint = tape-writing forward rollout + backward sweep, C5 chain64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--scaling weak|strong]

N > 1 is launched by torchrun (one rank per GPU); the batch is sharded per rank with no data-path collective.  Weak scaling
(default): the workload's batch per GPU.  Strong scaling: the workload's batch is the GLOBAL batch, split over the ranks.
The NCCL all-gather of the trajectories the north star asks for is timed separately and reported as `gather_ms`.

value : whole-job rollout-steps/s, inputs resident in HBM, device-timed (CUDA events on the launch stream), max over ranks
e2e   : the same through the host-pointer C ABI call (rmx_rollout / rmx_rollout_adjoint): page-locked host inputs -> device,
        kernels, results -> page-locked host, all inside the timed region; `e2e_pageable` is the same call with ordinary
        (pageable) host arrays, which is what a MATLAB mxArray caller has
roofline : algorithmic HBM bytes per rollout-step (forward: q, qdot out = 16*nr, + 16*nr/nsteps for q0, qdot0 -- the workload
           has no per-step control input, so SURVEY 8(d)'s 8*nr "tau in" term is not moved and not counted; adjoint: the tape
           written once and read once) / kernel time vs the measured HBM peak, plus the FP64 view against the FP64 rate measured
           in this run (rmx_fp64_probe); this path is FP64 / latency bound, see DESIGN.md
cpu_baseline : the reference algorithm (oracle) timed on the host cores on a bounded sample of the same workload
--impl reference : only the CPU reference arm (oracle, all host threads); under torchrun rank 0 alone runs it
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'rollout-steps/sec'
UNIT = 'rollout-steps/s'

WORKLOADS = {
    # name: n links, scheme, h, nsteps, rollouts per GPU, ground ; kind 'fwd' (rmx_rollout) or 'adjoint' (rmx_rollout_adjoint)
    # h = 1e-3: at the reference's default h = 1e-2 its absolute Newton tolerance (1e-9, driverRedMaxBDF1.m:95) sits below
    # the round-off floor of g for a 32-link chain and newton() runs to iterMax in most steps
    'chain32-bdf1-b4096': dict(kind='fwd', n=32, scheme=1, h=1e-3, nsteps=100, B=4096, ground=False),
    'chain32-ground-bdf2-b4096': dict(kind='fwd', n=32, scheme=2, h=2e-4, nsteps=100, B=4096, ground=True),
    'chain10-bdf1-b1024': dict(kind='fwd', n=10, scheme=1, h=1e-3, nsteps=100, B=1024, ground=False),
    'chain64-bdf1-b8192': dict(kind='fwd', n=64, scheme=1, h=2e-4, nsteps=100, B=8192, ground=False),
    # C4: hand tree (fixed palm + 5 fingers x 4 revolute), TaskBDF1PointPos, objective + gradient; B = one GPU's share of 8192 on 4
    'hand20-adjoint-bdf1-b2048': dict(kind='adjoint', n=20, scheme=1, h=1e-2, nsteps=100, B=2048, ground=False),
}
SEED = 20260003


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs (burst copy, kernel timed alone)', d
        except Exception:
            pass
    return 6500.0, 'fallback from B200_PROFILING.md (MEASURED_PEAKS.json absent)', {}


def make_scene(name, api=None):
    """The scene of a workload, built through the host mirror of the +redmax API (or the oracle's, for the CPU arm)."""
    import redmax_b200.scenes as scenes
    w = WORKLOADS[name]
    kw = {} if api is None else dict(api=api)
    if w['kind'] == 'adjoint':
        sc = scenes.hand_scene(h=w['h'], nsteps=w['nsteps'], scheme=w['scheme'], **kw)
    else:
        sc = scenes.chain_scene(w['n'], ground=w['ground'], h=w['h'], nsteps=w['nsteps'], **kw)
    sc.init()
    return sc


def adjoint_inputs(sc, B, seed):
    """C4 inputs of SURVEY.md 8(d): p = 0.01 U(-1,1)^nr, xtarget = the task's target + U(-2,2)^3, per rollout."""
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(seed))
    p = 0.01 * rng.uniform(-1.0, 1.0, (B, sc.nr))
    xt = np.asarray(sc.task.xtarget, dtype=float)[None, :] + rng.uniform(-2.0, 2.0, (B, 3))
    return np.ascontiguousarray(p), np.ascontiguousarray(xt)


# ---------------------------------------------------------------------------------------------------------
# CPU reference arm (test infrastructure: executes oracle/)
# ---------------------------------------------------------------------------------------------------------
def _np_oracle_worker(args):
    name, seed, b, nsteps = args
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import redmax_oracle as oracle
    import redmax_b200.scenes as scenes
    w = WORKLOADS[name]
    so = make_scene(name, api=oracle)
    if w['kind'] == 'adjoint':
        p, xt = adjoint_inputs(so, b + 1, seed)
        so.task.setTarget(xt[b])
        oracle.task_objective(p[b], so, w['scheme'])
        return w['nsteps']
    q0, qd0 = scenes.synthetic_inputs(so, b + 1, seed=seed)
    oracle.run_forward(so, w['scheme'], q0[b], qd0[b], nsteps=nsteps)
    return nsteps


def cpu_reference(name, rollouts, nsteps, threads):
    """Time the oracle (the reference's dense algorithm) on `rollouts` rollouts x `nsteps` time steps of workload `name`
    using `threads` host threads.  Returns (rollout-steps/s, kind, description)."""
    w = WORKLOADS[name]
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    have_c = False
    if w['kind'] == 'fwd':
        try:
            import oracle_c  # compiled twin of the numpy oracle (oracle/redmax_oracle_c.c): forward rollouts only
            have_c = oracle_c.available()
        except Exception:
            have_c = False
    if have_c:
        import redmax_b200.scenes as scenes
        import redmax_oracle as oracle
        so = make_scene(name, api=oracle)
        q0, qd0 = scenes.synthetic_inputs(so, rollouts, seed=SEED)
        t0 = time.perf_counter()
        oracle_c.run_forward_batch(so, w['scheme'], q0, qd0, nsteps=nsteps, threads=threads)
        dt = time.perf_counter() - t0
        return rollouts * nsteps / dt, 'port', ('C restatement of matlab-diff (dense algorithm), %d rollouts x %d steps, '
                                                 'OpenMP over rollouts' % (rollouts, nsteps))
    import multiprocessing as mp
    ctx = mp.get_context('spawn')
    if w['kind'] == 'adjoint':
        nsteps = w['nsteps']  # objective + gradient need the whole rollout and the backward sweep
    jobs = [(name, SEED, b, nsteps) for b in range(rollouts)]
    with ctx.Pool(threads) as pool:
        pool.map(_np_oracle_worker, [(name, SEED, 0, 1 if w['kind'] == 'fwd' else nsteps)] * threads)  # import + warm
        t0 = time.perf_counter()
        pool.map(_np_oracle_worker, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    what = 'taskObjective (forward + tape + backward sweep)' if w['kind'] == 'adjoint' else 'forward rollouts'
    return rollouts * nsteps / dt, 'port', ('NumPy restatement of matlab-diff (dense algorithm), %s, %d rollouts x %d steps, '
                                             'one process per rollout' % (what, rollouts, nsteps))


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_sample(name, cores, budget_s):
    """A bounded sample of workload `name` for the CPU arm: (rollouts, nsteps) worth about budget_s seconds on `cores`."""
    w = WORKLOADS[name]
    if w['kind'] == 'adjoint':
        return cores, w['nsteps']  # one whole objective + gradient per core (the NumPy oracle needs seconds for each)
    probe_v, _, _ = cpu_reference(name, cores, 1, cores)
    units = max(cores, int(probe_v * budget_s))
    nsteps = max(1, min(w['nsteps'], units // cores))
    rollouts = cores * max(1, units // (cores * nsteps))
    return rollouts, nsteps


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    name = args.workload
    cores = host_cores()
    # bounded sample per step, sized from a probe so that (steps + warmup) samples finish within a few minutes
    rollouts, nsteps = cpu_sample(name, cores, 150.0 / max(1, args.steps + args.warmup))
    steps, warmup = args.steps, args.warmup
    if WORKLOADS[name]['kind'] == 'adjoint':
        steps, warmup = min(steps, 3), min(warmup, 1)  # seconds per sample: keep the arm within minutes (stated in `steps`)
    for _ in range(warmup):
        cpu_reference(name, rollouts, nsteps, cores)
    t0 = time.perf_counter()
    for _ in range(steps):
        v, kind, desc = cpu_reference(name, rollouts, nsteps, cores)
    dt = time.perf_counter() - t0
    value = steps * rollouts * nsteps / dt
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
        'warmup': warmup, 'ms_per_step': 1e3 * dt / steps, 'higher_is_better': True, 'scaling': args.scaling,
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(name, args.gpus, args.scaling),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': kind,
                         'sample': '%s; per bench step: %d rollouts x first %d time steps of the workload' % (desc, rollouts, nsteps)},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def per_gpu_batch(name, gpus, scaling):
    B = WORKLOADS[name]['B']
    return B if scaling == 'weak' else max(1, B // gpus)


def workload_config(name, gpus, scaling='weak'):
    w = WORKLOADS[name]
    B = per_gpu_batch(name, gpus, scaling)
    tree = ('hand: fixed palm + 5 fingers x 4 revolute (20 DOF), TaskBDF1PointPos' if w['kind'] == 'adjoint'
            else '%d-link serial chain, all revolute' % w['n'])
    return {'workload': name, 'tree': tree, 'nr': w['n'],
            'scheme': ('BDF1' if w['scheme'] == 1 else 'SDIRK2+BDF2') + (' + adjoint (objective and gradient)' if w['kind'] == 'adjoint' else ''),
            'h': w['h'], 'nsteps': w['nsteps'],
            'ground_friction': w['ground'], 'rollouts_per_gpu': B, 'global_rollouts': B * gpus,
            'parallelism': 'batch sharded over %d GPU(s), no data-path collective; on each GPU the rollouts are load-balanced '
                           'over the co-resident blocks, one persistent launch' % gpus,
            'l2': 'flushed between timed iterations (256 MiB write); outputs per step exceed L2 as well',
            'seed': SEED}


# ---------------------------------------------------------------------------------------------------------
# clocks sampler (pynvml; the recipe's clocks line)
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4): 'sw_power_cap',
            getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8): 'hw_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonHwPowerBrakeSlowdown', 0x80): 'hw_power_brake_slowdown',
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        s = sorted(self.samples)
        return {'sm_mhz': (s[len(s) // 2] if s else None), 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(s)}


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import redmax_b200 as rb
    from redmax_b200 import _ffi, shard

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (redmax_b200 has no CPU fallback)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    name = args.workload
    w = WORKLOADS[name]
    adjoint = w['kind'] == 'adjoint'
    scheme, nsteps = w['scheme'], w['nsteps']
    B = per_gpu_batch(name, world, args.scaling)
    sc = make_scene(name)
    nr, n = sc.nr, len(sc.joints)
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- inputs: per-rank shard of the global seeded batch, rollouts [rank*B, (rank+1)*B) -----------------------------
    if adjoint:
        p_all, xt_all = adjoint_inputs(sc, B * world, SEED)
        p_h, xt_h = shard.take_shard(p_all, world, rank), shard.take_shard(xt_all, world, rank)
        q0 = np.ascontiguousarray(np.broadcast_to(sc.qInit, (B, nr)))
        qd0 = np.ascontiguousarray(np.broadcast_to(sc.qdotInit, (B, nr)))
        dp, dxt = torch.from_numpy(np.ascontiguousarray(p_h)).to(dev), torch.from_numpy(np.ascontiguousarray(xt_h)).to(dev)
        dP = torch.empty(B, dtype=torch.float64, device=dev)
        dG = torch.empty((B, nr), dtype=torch.float64, device=dev)
    else:
        q0_all, qd0_all = rb.synthetic_inputs(sc, B * world, seed=SEED)
        q0 = np.ascontiguousarray(shard.take_shard(q0_all, world, rank))
        qd0 = np.ascontiguousarray(shard.take_shard(qd0_all, world, rank))
        qo = torch.empty((B, nsteps, nr), dtype=torch.float64, device=dev)
        qdo = torch.empty_like(qo)
        it = torch.empty((B, 2), dtype=torch.int32, device=dev)
    dq0, dqd0 = torch.from_numpy(q0).to(dev), torch.from_numpy(qd0).to(dev)
    st = torch.empty(B, dtype=torch.int32, device=dev)

    if adjoint:
        def one_step():
            sc.rollout_adjoint_dev(dq0, dqd0, dp, dxt, dP, dG, st, stream=stream)
        launches_per_step = 2  # tape-writing forward kernel + backward sweep
    else:
        def one_step():
            sc.rollout_dev(dq0, dqd0, qo, qdo, st, it, scheme=scheme, stream=stream)
        launches_per_step = 1

    for _ in range(max(args.warmup, 3)):
        one_step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    evs = []
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        one_step()
        e1.record(stream)
        evs.append((e0, e1))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = float(sum(ms))
    stc = st.cpu().numpy()
    sc.check_status(stc)
    if adjoint:
        finite = bool(torch.isfinite(dP).all().item() and torch.isfinite(dG).all().item())
        newton = ls = None
    else:
        itc = it.cpu().numpy()
        finite = bool(torch.isfinite(qo).all().item())
        newton = float(itc[:, 0].mean()) / nsteps
        ls = float(itc[:, 1].mean()) / nsteps

    # ---- e2e: host-pointer C ABI call, page-locked and pageable host buffers ------------------------------------------
    e2e_steps = max(1, min(args.steps, 5))
    if adjoint:
        def host_buffers(make):
            return dict(q0=make(q0), qd0=make(qd0), p=make(p_h), xt=make(xt_h), P=make(np.empty(B)), G=make(np.empty((B, nr))),
                        st=make(np.empty(B, dtype=np.int32)))

        def e2e_call(hb):
            sc.rollout_adjoint_into(hb['q0'], hb['qd0'], hb['p'], hb['xt'], hb['P'], hb['G'], hb['st'])
        h2d = (3 * B * nr + 3 * B) * 8
        d2h = B * 8 + B * nr * 8 + B * 4
        api = 'rmx_rollout_adjoint (host pointers): q0, qdot0, p, xtarget -> device, tape-writing rollout + backward sweep, P, dP/dp, status -> host'
    else:
        def host_buffers(make):
            return dict(q0=make(q0), qd0=make(qd0), q=make(np.empty((B, nsteps, nr))), qd=make(np.empty((B, nsteps, nr))))

        def e2e_call(hb):
            return sc.rollout_into(hb['q0'], hb['qd0'], hb['q'], hb['qd'], scheme=scheme)
        h2d = 2 * B * nr * 8
        d2h = 2 * B * nsteps * nr * 8 + B * 4 + B * 8
        api = ('rmx_rollout (host pointers; with page-locked buffers q(t), qdot(t) are stored to the mapped host buffers by the '
               'kernel as it runs, status/iters copied after; pageable buffers go through page-locked staging written the same way and '
               'are copied on by host threads, sub-batch by sub-batch)')

    def time_e2e(hb):
        for _ in range(2):
            e2e_call(hb)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            e2e_call(hb)
        barrier()
        return time.perf_counter() - t0

    keep = []

    def pinned(a):
        t = pin(a)
        keep.append(t)
        return t.numpy()
    hb_pin = host_buffers(pinned)
    e2e_s = time_e2e(hb_pin)
    hb_page = host_buffers(lambda a: np.array(a, copy=True))
    e2e_page_s = time_e2e(hb_page)
    # untimed check call into NaN-filled buffers: what the call delivers to the host equals the device path bit for bit
    if adjoint:
        hb_pin['P'].fill(np.nan)
        hb_pin['G'].fill(np.nan)
        e2e_call(hb_pin)
        same = bool(np.array_equal(hb_pin['P'], dP.cpu().numpy()) and np.array_equal(hb_pin['G'], dG.cpu().numpy()))
        e2e_call(hb_page)
        same = same and bool(np.array_equal(hb_page['G'], hb_pin['G']))
    else:
        for hb in (hb_pin, hb_page):
            hb['q'].fill(np.nan)
            hb['qd'].fill(np.nan)
            e2e_call(hb)
        same = bool(np.array_equal(hb_pin['q'], qo.cpu().numpy()) and np.array_equal(hb_pin['qd'], qdo.cpu().numpy())
                    and np.array_equal(hb_page['q'], hb_pin['q']) and np.array_equal(hb_page['qd'], hb_pin['qd']))

    # ---- trajectory gather the north star asks for (timed separately) -----------------------------------------
    gather_ms = None
    if world > 1 and not adjoint:
        gathered = torch.empty((world * B, nsteps, nr), dtype=torch.float64, device=dev)
        for _ in range(2):  # warm-up: communicator set-up and buffer registration are not the collective
            shard.gather_trajectories(qo, B=world * B, out=gathered)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        shard.gather_trajectories(qo, B=world * B, out=gathered)
        g1.record()
        torch.cuda.synchronize()
        gather_ms = g0.elapsed_time(g1)
        assert torch.equal(gathered[rank * B:(rank + 1) * B], qo)

    # ---- measured FP64 rate of this device (MEASURED_PEAKS.json holds none) -----------------------------------
    import ctypes as C
    dfma, dmma = C.c_double(0.0), C.c_double(0.0)
    _ffi.check(_ffi.lib().rmx_fp64_probe(C.byref(dfma), C.byref(dmma)), 'rmx_fp64_probe')

    # ---- max over ranks ------------------------------------------------------------------------------------
    red = torch.tensor([total_ms, e2e_s, gather_ms or 0.0, e2e_page_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
    total_ms, e2e_s, gather_max, e2e_page_s = [float(x) for x in red.cpu()]
    units = float(B) * nsteps * world
    value = units * args.steps / (total_ms * 1e-3)
    e2e_value = units * e2e_steps / e2e_s
    e2e_page_value = units * e2e_steps / e2e_page_s

    line = None
    if rank == 0:
        peak, peak_src, peaks = measured_peaks()
        kern_ms = total_ms / args.steps  # per step per rank
        if adjoint:
            # the tape (LU(H) + perm + dP/dq_k, M, D per rollout-step) is written once by the forward kernel and read once by
            # the backward sweep; q0, qdot0, p, xtarget in and P, dP/dp out are amortised over the steps
            tape = sc.adjoint_tape_bytes(B) / float(B * nsteps)
            balg = 2.0 * tape + (3 * nr + 3 + nr + 1) * 8.0 / nsteps
            bytes_note = ('adjoint tape written once + read once: 2 x 8 x (nr*(nr|1) + 1.5 nr + 2 nr^2) = %.0f B per rollout-step '
                          '(SURVEY 8(d) counts 48 nr^2 + 48 nr = %d)' % (2.0 * tape, 48 * nr * nr + 48 * nr))
            f_step = None
        else:
            balg = 16.0 * nr + 16.0 * nr / nsteps
            bytes_note = ('q, qdot out = 16 nr per rollout-step + q0, qdot0 in amortised; the workload has no per-step control, so '
                          "SURVEY 8(d)'s 8 nr tau-in term is not moved and not counted")
            # algorithmic FP64 flops of the lean formulation: per Newton iteration
            #   pair sweep ~ 330 flop x n(n+1)/2 pairs, LU 2/3 nr^3 + 2 nr^2, per-joint phases ~ 600 flop x n;
            #   per residual-only evaluation ~ 600 flop x n
            f_iter = 330.0 * n * (n + 1) / 2 + (2.0 / 3.0) * nr ** 3 + 2.0 * nr * nr + 600.0 * n
            f_step = newton * f_iter + ls * 600.0 * n
        achieved = balg * B * nsteps / (kern_ms * 1e-3) / 1e9
        ncu = ncu_record(name)
        roof = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': ncu.get('dram_bytes_per_launch'), 'peak_source': peak_src, 'alg_bytes_per_rollout_step': balg,
                'alg_bytes_note': bytes_note,
                'fp64': {'peak_measured_tflops': {'dfma': dfma.value, 'dmma_8x8x4': dmma.value},
                         'peak_source': 'rmx_fp64_probe in this run: independent DFMA / DMMA.8x8x4 chains on every SM, 64 warps per SM',
                         'ncu_pipe_fp64_cycles_active_pct': ncu.get('sm__pipe_fp64_cycles_active_pct'),
                         'ncu_note': ncu.get('note')}}
        if f_step is not None:
            roof['note'] = ('FP64 / dependency-latency bound by construction (arithmetic intensity ~%.0f flop/B)' % (f_step / balg))
            roof['fp64'].update({'alg_flop_per_rollout_step_model': f_step,
                                 'achieved_tflops_model': value / world * f_step / 1e12,
                                 'model_note': 'flop count of the scalar composite formulation x measured Newton iterations: a model '
                                               'kept for comparison across kernel generations; the pipe counter is the measurement'})
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': kern_ms, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'config': workload_config(name, world, args.scaling),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'steps': e2e_steps, 'api': api, 'host_buffers': 'page-locked', 'bitwise_equal_to_device_path': same},
            'e2e_pageable': {'value': e2e_page_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                             'steps': e2e_steps, 'host_buffers': 'pageable (what a MATLAB mxArray caller has)'},
            'gpu_launches': args.steps * launches_per_step,
            'kernel': ('rmx::rollout_fwd_kernel<ADJ> + rmx::adjoint_bwd_kernel (two launches per bench step: %d time steps of %d rollouts '
                       'forward with tape, then the backward sweep)' % (nsteps, B)) if adjoint else
                      ('rmx::rollout_fwd_kernel (one persistent launch per bench step: all %d time steps of %d rollouts)' % (nsteps, B)),
            'clocks': clocks, 'roofline': roof,
            'newton_iters_per_step': newton, 'linesearch_evals_per_step': ls,
            'status_nonzero_frac': float((stc != 0).mean()), 'finite': finite,
            'wall_s_timed_region': t_wall, 'gather_ms': (gather_max if gather_ms is not None else None),
            'host_cores': host_cores(),
        }
    if world > 1:
        dist.barrier()
    if rank == 0:
        if world == 1 and not args.no_cpu:
            cores = host_cores()
            ro_cpu, ns_cpu = cpu_sample(name, cores, 15.0)
            v, kind, desc = cpu_reference(name, ro_cpu, ns_cpu, cores)
            line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': kind,
                                    'sample': '%s (first %d of %d time steps)' % (desc, ns_cpu, nsteps)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def ncu_record(name):
    """What the committed ncu capture of this workload's dominant kernel says (profiles/ncu_traffic.json): DRAM bytes per
    launch (dram__bytes_read.sum + dram__bytes_write.sum) and the FP64 pipe utilisation.  {} if there is none."""
    p = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    try:
        v = json.load(open(p)).get(name)
    except Exception:
        return {}
    if isinstance(v, dict):
        return v
    return {} if v is None else {'dram_bytes_per_launch': v}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='chain32-bdf1-b4096', choices=sorted(WORKLOADS))
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help="weak: the workload's batch per GPU; strong: the workload's batch split over the GPUs")
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    return run_ours(args)


if __name__ == '__main__':
    sys.exit(main())
