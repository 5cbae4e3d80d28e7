"""Registers / stack / spills of every kernel from the per-unit ptxas logs of the last build (build/obj/*.log)."""
import glob
import os
import re
import subprocess

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
PAT = re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'\nptxas info\s+: Function properties for \S+\n\s+(\d+) bytes stack frame, "
                 r"(\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers")

if __name__ == '__main__':
    for f in sorted(glob.glob(os.path.join(ROOT, 'build', 'obj', '*.log'))):
        for m in PAT.finditer(open(f).read()):
            name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace('rmx::', '').replace('(RolloutArgs)', '').replace('void ', '')
            print('%-26s %-52s stack %5s  spill st %5s ld %5s  regs %s'
                  % (os.path.basename(f)[:-4], name[:52], m.group(2), m.group(3), m.group(4), m.group(5)))
