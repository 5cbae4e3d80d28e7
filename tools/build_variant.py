"""Developer tool: build a tagged variant of the CUDA library with extra nvcc flags for A/B measurements, e.g.
    python tools/build_variant.py exact -DRMX_PIVOT_EXACT     ->  build/lib_exact/libredmax_b200.so
and run anything against it with RMX_LIB=build/lib_exact/libredmax_b200.so.  The product library is the untagged build."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import __graft_entry__ as ge  # noqa: E402

if __name__ == '__main__':
    tag, flags = sys.argv[1], sys.argv[2:]
    print(ge.build_cuda(extra_flags=flags, tag=tag))
