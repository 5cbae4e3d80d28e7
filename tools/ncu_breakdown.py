"""Per-source-region breakdown of one kernel of an `ncu --set full --import-source on` report.

ncu's CLI prints the source page per SASS instruction only; this tool joins it with the line table of the same kernel in
the in-tree library (cuobjdump -xelf + nvdisasm -g: `//## File "...", line N` markers, same instruction order) and sums
warp-stall samples and executed warp-instructions per source file, per named region (line ranges below) and per line.

    python tools/ncu_breakdown.py gpurun_out/r02g_rollout_end.ncu-rep \
        --kernel _ZN3rmx18rollout_fwd_kernelILi1ELb0ELb0ELi2ELi0EEEvNS_11RolloutArgsE --units 409600 > profiles/...txt

--units: rollout-steps per launch (instructions are also printed per unit).  The library must be the build the report
was taken with (the SASS instruction count is checked).
"""
import argparse
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# named regions: (file suffix, first line, last line, name); first match wins, anything else goes by file name
REGIONS = [
    ('rmx_tc.cuh', 1, 45, 'tc:dmma wrapper, pivot reciprocal'),
    ('rmx_tc.cuh', 46, 191, 'tc:columns (per-joint vectors, tiles)'),
    ('rmx_tc.cuh', 192, 361, 'tc:LU panel (pivot search, broadcast, rank-1 updates)'),
    ('rmx_tc.cuh', 362, 392, 'tc:LU remaining rows + U12'),
    ('rmx_tc.cuh', 393, 472, 'tc:LU trailing update'),
    ('rmx_tc.cuh', 473, 560, 'tc:LU back substitution'),
    ('rmx_rollout.cuh', 1, 10000, 'rollout (newton, line search, time loop, schedule, tape stores)'),
    ('rmx_fast.cuh', 312, 345, 'fast: xtmx_store (X^T K X, X^T D X of the contact blocks)'),
    ('rmx_fast.cuh', 1, 10000, 'fast: composite base evaluation'),
    ('rmx_device.cuh', 195, 222, 'device: group_barrier (lockstep groups: waiting for the slowest warp of the block)'),
    ('rmx_device.cuh', 414, 590, 'device: ground_body (ForceGroundCuboid)'),
    ('rmx_device.cuh', 1, 10000, 'device helpers (se3, reductions)'),
]


def sh(cmd):
    return subprocess.run(cmd, check=True, capture_output=True, text=True).stdout


def line_table(lib, kernel):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(['cuobjdump', '-xelf', 'all', lib], cwd=d, check=True, capture_output=True)
        cubins = [os.path.join(d, f) for f in os.listdir(d) if f.endswith('.cubin')]
        for cb in cubins:
            txt = sh(['nvdisasm', '-g', '-c', cb])
            m = re.search(r'^\s*\.section\s+\.text\.%s,' % re.escape(kernel), txt, re.M)
            if not m:
                continue
            body = txt[m.end():]
            nxt = re.search(r'^\s*\.section\s', body, re.M)
            body = body[:nxt.start()] if nxt else body
            cur, out = ('?', 0), []
            for ln in body.split('\n'):
                f = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
                if f:
                    cur = (f.group(1), int(f.group(2)))
                    continue
                if re.match(r'\s*/\*[0-9a-f]{4,}\*/', ln):  # an instruction: /*0000*/  OPCODE ...
                    out.append(cur)
            return out
    raise SystemExit('kernel %s not found in %s' % (kernel, lib))


def region_of(path, line):
    for suf, a, b, name in REGIONS:
        if path.endswith(suf) and a <= line <= b:
            return name
    return os.path.basename(path)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('report')
    ap.add_argument('--kernel', required=True, help='mangled kernel name as in the library')
    ap.add_argument('--lib', default=os.path.join(ROOT, 'redmax_b200', 'lib', 'libredmax_b200.so'))
    ap.add_argument('--units', type=float, default=0.0)
    ap.add_argument('--top', type=int, default=25)
    ap.add_argument('--reason', default='# Samples',
                    help="source-page column to use as the sample count, e.g. stall_no_inst, stall_long_sb (default: all samples)")
    a = ap.parse_args()
    src = sh(['ncu', '-i', a.report, '--page', 'source', '--csv'])
    rows = list(csv.reader(io.StringIO(src)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    names = rows[hdr]
    ci, cs = names.index('Instructions Executed'), names.index(a.reason)
    sass = rows[hdr + 1:]
    lt = line_table(a.lib, a.kernel)
    if len(lt) != len(sass):
        raise SystemExit('SASS length differs: report %d, library %d instructions -- not the build the report was taken with'
                         % (len(sass), len(lt)))
    tot_i = sum(float(r[ci]) for r in sass)
    tot_s = sum(float(r[cs]) for r in sass)
    by_region, by_line = defaultdict(lambda: [0.0, 0.0]), defaultdict(lambda: [0.0, 0.0])
    for (path, line), r in zip(lt, sass):
        for d, k in ((by_region, region_of(path, line)), (by_line, (os.path.basename(path), line))):
            d[k][0] += float(r[cs])
            d[k][1] += float(r[ci])
    print('report %s' % os.path.basename(a.report))
    print('kernel %s' % a.kernel)
    if a.reason != '# Samples':
        print('samples counted: %s only' % a.reason)
    print('SASS instructions %d, warp-instructions executed %.0f%s, stall samples %.0f'
          % (len(sass), tot_i, (' (%.2fk per unit)' % (tot_i / a.units / 1e3)) if a.units else '', tot_s))
    print('--- regions (share of samples, share of executed warp-instructions%s)' % (', k instructions per unit' if a.units else ''))
    for k, (s, i) in sorted(by_region.items(), key=lambda kv: -kv[1][0]):
        per = ('  %6.2fk' % (i / a.units / 1e3)) if a.units else ''
        print('%-58s %6.2f%%  %6.2f%%%s' % (k, 100 * s / tot_s, 100 * i / tot_i, per))
    print('--- top lines (file, line, share of samples, share of instructions)')
    for k, (s, i) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:a.top]:
        print('%-22s %5d  %6.2f%%  %6.2f%%' % (k[0], k[1], 100 * s / tot_s, 100 * i / tot_i))


if __name__ == '__main__':
    sys.exit(main())
