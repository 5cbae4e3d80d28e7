set -x
O=gpurun_out/r2s; mkdir -p $O
run() { # workload group
  RMX_GROUP=$2 timeout 300 python bench.py --workload $1 --steps 6 --warmup 3 --no-cpu > $O/$1.g$2.log 2>&1
  python - $O/$1.g$2.log $1 $2 <<'PY'
import json,sys
ok=False
for l in open(sys.argv[1]):
    if l.startswith('{'):
        ok=True
        d=json.loads(l); print('GROUP %-28s G=%s  %8.3f ms  %.3f M rollout-steps/s  status!=0 %.4f finite %s' % (sys.argv[2], sys.argv[3], d['ms_per_step'], d['value']/1e6, d['status_nonzero_frac'], d['finite']))
if not ok: print('GROUP', sys.argv[2], sys.argv[3], 'FAILED'); print(open(sys.argv[1]).read()[-1500:])
PY
}
for g in 1 2 3 5; do run chain32-ground-bdf2-b4096 $g; done
for g in 2 4 8; do run chain32-bdf1-b4096 $g; done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_long_chains.py tests/test_gpu_trees.py -q -m gpu -x -k "ground or shortcut or sched or tree" > $O/tests.log 2>&1; tail -4 $O/tests.log
