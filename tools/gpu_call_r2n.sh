set -x
O=gpurun_out/r2n; mkdir -p $O
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
ls -la $O
timeout 900 python -m pytest tests/test_gpu_trees.py -q -m gpu -x > $O/trees.log 2>&1; tail -15 $O/trees.log
