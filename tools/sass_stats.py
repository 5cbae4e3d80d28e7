"""Developer tool: static SASS statistics of the kernels in one object / library: instruction count and the mnemonics that
matter here (DMMA, DFMA/DMUL/DADD, LDS/STS, LDL/STL = local-memory traffic, SHFL, REDUX, BAR, MUFU)."""
import collections
import re
import subprocess
import sys

if __name__ == '__main__':
    path = sys.argv[1]
    filt = sys.argv[2] if len(sys.argv) > 2 else ''
    out = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True).stdout
    name, stats = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r'\s+Function : (\S+)', line)
        if m:
            name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace('rmx::', '').replace('(RolloutArgs)', '').replace('void ', '')
            stats[name] = collections.Counter()
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m and name:
            op = m.group(1).split('.')[0]
            stats[name]['total'] += 1
            stats[name][op] += 1
    keys = ['total', 'DMMA', 'DFMA', 'DMUL', 'DADD', 'LDS', 'STS', 'LDL', 'STL', 'SHFL', 'REDUX', 'BAR', 'MUFU', 'LDG', 'STG']
    print('%-60s' % 'kernel' + ''.join('%7s' % k for k in keys))
    for nme, c in stats.items():
        if filt in nme:
            print('%-60s' % nme[:60] + ''.join('%7d' % c[k] for k in keys))
