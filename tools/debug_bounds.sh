#!/bin/bash
# the debug build with in-kernel bounds checks (stands in for compute-sanitizer, which the pool does not offer):
#   tools/debug_bounds.sh            build exp/lib_bounds.so
#   tools/debug_bounds.sh run        on the GPU box: the whole -m gpu suite against it (a violated check raises GtfError)
cd "$(dirname "$0")/.." && mkdir -p exp
if [ "$1" = run ]; then
  GTF_LIB=$PWD/exp/lib_bounds.so python -m pytest tests -m gpu -x -q 2>&1 | tail -4
else
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -diag-suppress 550 -DGTF_DEBUG_BOUNDS \
    -o exp/lib_bounds.so gnn-track-finding_b200/csrc/gtf_b200.cu && ls -la exp/lib_bounds.so
fi
