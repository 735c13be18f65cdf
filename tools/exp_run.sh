#!/bin/bash
# on the GPU box: time the iteration kernels with every exp/lib_*.so   (tools/exp_run.sh [events])
cd "$(dirname "$0")/.."
for f in exp/lib_*.so; do
  echo "== $f"; GTF_LIB=$PWD/$f python tools/prof_iter.py ${1:-128} 2>&1 | tail -2
done
