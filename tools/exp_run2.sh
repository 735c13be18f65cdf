#!/bin/bash
# on the GPU box: tools/exp_run.sh for every value of GTF_L2_FETCH given   (tools/exp_run2.sh [events] [granularity ...])
cd "$(dirname "$0")/.."
ev=${1:-128}; shift
for g in "$@"; do
  for f in exp/lib_*.so; do
    echo "== $f GTF_L2_FETCH=$g"; GTF_L2_FETCH=$g GTF_LIB=$PWD/$f python tools/prof_iter.py $ev 2>&1 | tail -2
  done
done
