#!/usr/bin/env python
"""bench.py -- rollout-steps/s of the batched RedMax stepper on B200 (BASELINE.json metric).

Workload (`config.workload`): 32-link serial chain, BDF1, h = 1e-3, 100 time steps, 4096 rollouts per GPU with
seeded per-rollout initial states (SURVEY.md section 8(d), C3 shape at the north-star scheme).  One bench "step" is one
pass of the hot path over that batch: one launch of the persistent rollout kernel = 4096 x 100 rollout-steps.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

N > 1 is launched by torchrun (one rank per GPU); the batch is sharded per rank with no data-path collective
(weak scaling, 4096 rollouts per GPU); the NCCL all-gather of the trajectories the north star asks for is timed
separately and reported as `gather_ms`.

value : whole-job rollout-steps/s, inputs resident in HBM, device-timed (CUDA events on the launch stream), max over ranks
e2e   : the same through the host-pointer C ABI call (rmx_rollout): pinned host q0/qdot0 -> device, kernel, q(t)/qdot(t)
        -> pinned host, all inside the timed region
roofline : algorithmic HBM bytes (24*nr per rollout-step + 16*nr/nsteps) / kernel time vs measured HBM peak,
           plus the FP64 view (this path is FP64/latency bound)
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
    # name: (n links, scheme, h, nsteps, rollouts per GPU, ground)
    # h = 1e-3: at the reference's default h = 1e-2 its absolute Newton tolerance (1e-9, driverRedMaxBDF1.m:95) sits below
    # the round-off floor of g for a 32-link chain and newton() runs to iterMax in most steps
    'chain32-bdf1-b4096': dict(n=32, scheme=1, h=1e-3, nsteps=100, B=4096, ground=False),
    'chain32-ground-bdf2-b4096': dict(n=32, scheme=2, h=5e-4, nsteps=100, B=4096, ground=True),
    'chain10-bdf1-b1024': dict(n=10, scheme=1, h=1e-3, nsteps=100, B=1024, ground=False),
    'chain64-bdf1-b8192': dict(n=64, scheme=1, h=2e-4, nsteps=100, B=8192, ground=False),
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
    so = scenes.chain_scene(w['n'], ground=w['ground'], h=w['h'], nsteps=w['nsteps'], api=oracle)
    so.init()
    q0, qd0 = scenes.synthetic_inputs(so, b + 1, seed=seed)
    oracle.run_forward(so, w['scheme'], q0[b], qd0[b], nsteps=nsteps)
    return nsteps


def cpu_reference(name, rollouts, nsteps, threads):
    """Time the oracle (the reference's dense algorithm) on `rollouts` rollouts x `nsteps` time steps of workload `name`
    using `threads` host threads.  Returns (rollout-steps/s, kind, description)."""
    w = WORKLOADS[name]
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    try:
        import oracle_c  # compiled twin of the numpy oracle (oracle/redmax_oracle_c.c)
        have_c = oracle_c.available()
    except Exception:
        have_c = False
    if have_c:
        import redmax_b200.scenes as scenes
        import redmax_oracle as oracle
        so = scenes.chain_scene(w['n'], ground=w['ground'], h=w['h'], nsteps=w['nsteps'], api=oracle)
        so.init()
        q0, qd0 = scenes.synthetic_inputs(so, rollouts, seed=SEED)
        t0 = time.perf_counter()
        oracle_c.run_forward_batch(so, w['scheme'], q0, qd0, nsteps=nsteps, threads=threads)
        dt = time.perf_counter() - t0
        return rollouts * nsteps / dt, 'port', ('C restatement of matlab-diff (dense algorithm), %d rollouts x %d steps, '
                                                 'OpenMP over rollouts' % (rollouts, nsteps))
    import multiprocessing as mp
    ctx = mp.get_context('spawn')
    jobs = [(name, SEED, b, nsteps) for b in range(rollouts)]
    with ctx.Pool(threads) as pool:
        pool.map(_np_oracle_worker, [(name, SEED, 0, 1)] * threads)  # import + warm
        t0 = time.perf_counter()
        pool.map(_np_oracle_worker, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    return rollouts * nsteps / dt, 'port', ('NumPy restatement of matlab-diff (dense algorithm), %d rollouts x %d steps, '
                                             'one process per rollout' % (rollouts, nsteps))


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    name = args.workload
    w = WORKLOADS[name]
    cores = host_cores()
    # bounded sample per step, sized from a probe so that (steps + warmup) samples finish within a few minutes
    probe_v, kind, desc = cpu_reference(name, cores, 1, cores)
    budget_s = 150.0 / max(1, args.steps + args.warmup)
    units = max(cores, int(probe_v * budget_s))
    nsteps = max(1, min(w['nsteps'], units // cores))
    rollouts = cores * max(1, units // (cores * nsteps))
    for _ in range(args.warmup):
        cpu_reference(name, rollouts, nsteps, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, kind, desc = cpu_reference(name, rollouts, nsteps, cores)
    dt = time.perf_counter() - t0
    value = args.steps * rollouts * nsteps / dt
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(name, args.gpus),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': kind,
                         'sample': '%s; per bench step: %d rollouts x first %d time steps of the workload' % (desc, rollouts, nsteps)},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(name, gpus):
    w = WORKLOADS[name]
    return {'workload': name, 'tree': '%d-link serial chain, all revolute' % w['n'], 'nr': w['n'],
            'scheme': 'BDF1' if w['scheme'] == 1 else 'SDIRK2+BDF2', 'h': w['h'], 'nsteps': w['nsteps'],
            'ground_friction': w['ground'], 'rollouts_per_gpu': w['B'], 'global_rollouts': w['B'] * gpus,
            'parallelism': 'batch sharded over %d GPU(s), no data-path collective; on each GPU the rollouts are load-balanced '
                           'over the co-resident blocks (McNaughton schedule), one persistent launch' % gpus,
            'l2': 'flushed between timed iterations (256 MiB write); outputs per step (%.0f MB) exceed L2 as well'
                  % (2 * 8 * w['n'] * w['nsteps'] * w['B'] / 1e6),
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
    from redmax_b200 import shard

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
    n, scheme, nsteps, B = w['n'], w['scheme'], w['nsteps'], w['B']
    sc = rb.chain_scene(n, ground=w['ground'], h=w['h'], nsteps=nsteps)
    sc.init()
    nr = sc.nr
    # per-rank shard of the global seeded batch: rollouts [rank*B, (rank+1)*B)
    q0_all, qd0_all = rb.synthetic_inputs(sc, B * world, seed=SEED)
    q0 = np.ascontiguousarray(shard.take_shard(q0_all, world, rank))
    qd0 = np.ascontiguousarray(shard.take_shard(qd0_all, world, rank))
    dq0 = torch.from_numpy(q0).to(dev)
    dqd0 = torch.from_numpy(qd0).to(dev)
    qo = torch.empty((B, nsteps, nr), dtype=torch.float64, device=dev)
    qdo = torch.empty_like(qo)
    st = torch.empty(B, dtype=torch.int32, device=dev)
    it = torch.empty((B, 2), dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        sc.rollout_dev(dq0, dqd0, qo, qdo, st, it, scheme=scheme, stream=stream)

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
    itc = it.cpu().numpy()
    stc = st.cpu().numpy()
    finite = bool(torch.isfinite(qo).all().item())

    # ---- e2e: host-pointer C ABI call with pinned host buffers ------------------------------------------------
    hq0 = torch.from_numpy(q0).pin_memory()
    hqd0 = torch.from_numpy(qd0).pin_memory()
    hq = torch.empty((B, nsteps, nr), dtype=torch.float64).pin_memory()
    hqd = torch.empty((B, nsteps, nr), dtype=torch.float64).pin_memory()
    e2e_steps = max(1, min(args.steps, 5))
    out = None
    for _ in range(2):
        out = sc.rollout_into(hq0.numpy(), hqd0.numpy(), hq.numpy(), hqd.numpy(), scheme=scheme)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        out = sc.rollout_into(hq0.numpy(), hqd0.numpy(), hq.numpy(), hqd.numpy(), scheme=scheme)
    barrier()
    e2e_s = time.perf_counter() - t0
    # untimed check call into NaN-filled buffers: what the call delivers to the host equals the device path bit for bit
    hq.fill_(float('nan'))
    hqd.fill_(float('nan'))
    sc.rollout_into(hq0.numpy(), hqd0.numpy(), hq.numpy(), hqd.numpy(), scheme=scheme)
    same = bool(np.array_equal(hq.numpy(), qo.cpu().numpy()) and np.array_equal(hqd.numpy(), qdo.cpu().numpy()))
    h2d = 2 * B * nr * 8
    d2h = 2 * B * nsteps * nr * 8 + B * 4 + B * 8

    # ---- trajectory gather the north star asks for (timed separately) -----------------------------------------
    gather_ms = None
    if world > 1:
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

    # ---- max over ranks ------------------------------------------------------------------------------------
    red = torch.tensor([total_ms, e2e_s, gather_ms or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
    total_ms, e2e_s, gather_max = [float(x) for x in red.cpu()]
    units = float(B) * nsteps * world
    value = units * args.steps / (total_ms * 1e-3)
    e2e_value = units * e2e_steps / e2e_s

    line = None
    if rank == 0:
        peak, peak_src, peaks = measured_peaks()
        balg = 24.0 * nr + 16.0 * nr / nsteps  # bytes per rollout-step
        kern_ms = total_ms / args.steps          # one launch per step per rank
        achieved = balg * B * nsteps / (kern_ms * 1e-3) / 1e9
        newton = float(itc[:, 0].mean()) / nsteps
        ls = float(itc[:, 1].mean()) / nsteps
        # algorithmic FP64 flops of the lean formulation: per Newton iteration
        #   pair sweep ~ 330 flop x n(n+1)/2 pairs, LU 2/3 nr^3 + 2 nr^2, per-joint phases ~ 600 flop x n;
        #   per residual-only evaluation ~ 600 flop x n
        f_iter = 330.0 * n * (n + 1) / 2 + (2.0 / 3.0) * nr ** 3 + 2.0 * nr * nr + 600.0 * n
        f_step = newton * f_iter + ls * 600.0 * n
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': kern_ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'config': workload_config(name, world),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'steps': e2e_steps, 'api': 'rmx_rollout (host pointers, pinned buffers; q(t), qdot(t) stored to the mapped host buffers by the kernel as it runs, status/iters copied after)', 'bitwise_equal_to_device_path': same},
            'gpu_launches': args.steps,
            'kernel': 'rmx::rollout_fwd_kernel (one persistent launch per bench step: all %d time steps of %d rollouts)' % (nsteps, B),
            'clocks': clocks,
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': ncu_traffic(name), 'peak_source': peak_src,
                         'alg_bytes_per_rollout_step': balg,
                         'note': 'FP64 / dependency-latency bound by construction (arithmetic intensity ~%.0f flop/B); '
                                 'fp64 view below' % (f_step / balg),
                         'fp64': {'alg_flop_per_rollout_step': f_step, 'achieved_tflops': value / world * f_step / 1e12,
                                  'peak_tflops_nominal': 37.0}},
            'newton_iters_per_step': newton, 'linesearch_evals_per_step': ls,
            'status_nonzero_frac': float((stc != 0).mean()), 'finite': finite,
            'wall_s_timed_region': t_wall, 'gather_ms': (gather_max if world > 1 else None),
            'host_cores': host_cores(),
        }
    if world > 1:
        dist.barrier()
    if rank == 0:
        if world == 1 and not args.no_cpu:
            cores = host_cores()
            probe_v, kind, desc = cpu_reference(name, cores, 1, cores)
            units_cpu = max(cores, int(probe_v * 15.0))
            ns_cpu = max(1, min(nsteps, units_cpu // cores))
            ro_cpu = cores * max(1, units_cpu // (cores * ns_cpu))
            v, kind, desc = cpu_reference(name, ro_cpu, ns_cpu, cores)
            line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': kind,
                                    'sample': '%s (first %d of %d time steps)' % (desc, ns_cpu, nsteps)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def ncu_traffic(name):
    """DRAM bytes per launch of the rollout kernel from the committed ncu capture (profiles/ncu_traffic.json), or None."""
    p = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    try:
        return json.load(open(p)).get(name)
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='chain32-bdf1-b4096', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    return run_ours(args)


if __name__ == '__main__':
    sys.exit(main())
