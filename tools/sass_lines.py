"""Developer tool: static SASS size of one kernel by source line (nvdisasm -g line table of an object file), to see where the
code bytes are -- the external-force kernels are bound by instruction-cache misses (ncu: stall_no_instruction 5 per issue).
    python tools/sass_lines.py build/obj/rmx_k_fwd_i2_w1_g1.o 'rollout_fwd_kernelILi1ELb1ELb0' [top]"""
import collections
import os
import re
import subprocess
import sys
import tempfile

if __name__ == '__main__':
    obj, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=d, check=True, capture_output=True)
        for f in os.listdir(d):
            txt = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(d, f)], capture_output=True, text=True).stdout
            for m in re.finditer(r'^\s*\.section\s+\.text\.(\S+?),', txt, re.M):
                if pat not in m.group(1):
                    continue
                body = txt[m.end():]
                nxt = re.search(r'^\s*\.section\s', body, re.M)
                body = body[:nxt.start()] if nxt else body
                cur = ('?', 0)
                per_line, per_file, inl = collections.Counter(), collections.Counter(), collections.Counter()
                total = 0
                for ln in body.split('\n'):
                    fm = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
                    if fm:
                        cur = (os.path.basename(fm.group(1)), int(fm.group(2)))
                        continue
                    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', ln):
                        total += 1
                        per_line[cur] += 1
                        per_file[cur[0]] += 1
                print(m.group(1), 'total SASS instructions', total)
                for k, v in per_file.most_common():
                    print('   %-22s %6d' % (k, v))
                print('   top lines:')
                for (fn, l), v in per_line.most_common(top):
                    print('   %-22s line %5d  %6d' % (fn, l, v))
