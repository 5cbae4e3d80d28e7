set -x
O=gpurun_out/r2s; mkdir -p $O
run() { # workload group carve
  RMX_DEBUG_CARVEOUT=$3 RMX_GROUP=$2 timeout 300 python bench.py --workload $1 --steps 6 --warmup 3 --no-cpu > $O/$1.g$2.c$3.log 2>&1
  python - $O/$1.g$2.c$3.log $1 $2 $3 <<'PY'
import json,sys
ok=False
for l in open(sys.argv[1]):
    if l.startswith('{'):
        ok=True
        d=json.loads(l); print('GROUP %-28s G=%s carveout %s %%  %8.3f ms  %.3f M rollout-steps/s' % (sys.argv[2], sys.argv[3], sys.argv[4], d['ms_per_step'], d['value']/1e6))
if not ok: print('GROUP', sys.argv[2], sys.argv[3], 'FAILED'); print(open(sys.argv[1]).read()[-1500:])
PY
}
run chain32-ground-bdf2-b4096 5 0
run chain32-ground-bdf2-b4096 5 91
run chain32-ground-bdf2-b4096 4 0
run chain32-ground-bdf2-b4096 4 74
run chain32-ground-bdf2-b4096 4 80
run chain32-bdf1-b4096 1 0
run chain32-bdf1-b4096 1 85
