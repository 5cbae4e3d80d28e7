set -x
O=gpurun_out/r2u; mkdir -p $O
timeout 300 python bench.py --workload hand20-adjoint-bdf1-b2048 --steps 10 --warmup 3 --no-cpu > $O/adj_default.log 2>&1
RMX_LIB=$PWD/build/lib_adjni/libredmax_b200.so timeout 300 python bench.py --workload hand20-adjoint-bdf1-b2048 --steps 10 --warmup 3 --no-cpu > $O/adj_noinline.log 2>&1
for f in adj_default adj_noinline; do python - $O/$f.log $f <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('AB %-14s %8.3f ms  %.3f M rollout-steps/s' % (sys.argv[2], d['ms_per_step'], d['value']/1e6))
PY
done
