#!/bin/bash
# rebuild libgtf_b200.so with ptxas statistics of the iteration kernels
cd "$(dirname "$0")/.." && python -c "
import gtf_b200
from gtf_b200 import lib; lib.build(force=True, verbose=True)" 2>&1 | grep -E "error|warning|k_send|k_exec|k_node2|k_hv|k_big" -A2 | grep -E "error|warning|Compiling|Used|spill" | sed 's/ptxas info    : //; s/Compiling entry function//; s/for .sm_100a.//' | cut -c1-160
