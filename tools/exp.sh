#!/bin/bash
# build experiment variants of the library: tools/exp.sh name "-DFLAG ..." [name "-DFLAG" ...]; outputs exp/lib_<name>.so
cd "$(dirname "$0")/.." && mkdir -p exp
while [ $# -ge 2 ]; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -diag-suppress 550 $2 \
    -o exp/lib_$1.so gnn-track-finding_b200/csrc/gtf_b200.cu &
  shift 2
done
wait
ls exp/
