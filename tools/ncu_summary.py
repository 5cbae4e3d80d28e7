"""Developer tool: the JSON summary of one `ncu --set full` report that is committed under profiles/ (the .ncu-rep itself
stays in gpurun_out/, which is scratch).
    python tools/ncu_summary.py gpurun_out/x.ncu-rep --workload '...' --command '...' [--note '...'] > profiles/r02_ncu_x_summary.json"""
import argparse
import csv
import io
import json
import subprocess

KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static', 'launch__occupancy_limit_registers',
    'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'launch__waves_per_multiprocessor',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
    # instruction fetch: SM-level instruction cache and the GPC-level cache behind it
    'sm__icc_requests.sum', 'sm__icc_request_hit_rate.pct', 'gcc__cache_requests_type_instruction.sum',
    'gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed',
]


def to_bytes(v, unit):
    f = float(v)
    return f * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(unit, 1)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('report')
    ap.add_argument('--workload', default='')
    ap.add_argument('--command', default='')
    ap.add_argument('--note', default='')
    a = ap.parse_args()
    raw = subprocess.run(['ncu', '-i', a.report, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        m = {k: {'value': d[k], 'unit': u[k]} for k in KEYS if k in d}
        stalls = {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''): round(float(d[k]), 3)
                  for k in hdr if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio')
                  and float(d[k]) >= 0.005}
        rec = {'kernel': d.get('Kernel Name'), 'grid': d.get('Grid Size'), 'block': d.get('Block Size'), 'workload': a.workload,
               'command': a.command, 'note': a.note, 'metrics': m, 'warp_stalls_per_issue': stalls}
        if 'dram__bytes_read.sum' in d:
            rec['dram_bytes_per_launch'] = to_bytes(d['dram__bytes_read.sum'], u['dram__bytes_read.sum']) + \
                to_bytes(d['dram__bytes_write.sum'], u['dram__bytes_write.sum'])
        out.append(rec)
    print(json.dumps(out[0] if len(out) == 1 else out, indent=1))
