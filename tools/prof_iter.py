import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
import gtf_b200, bench
ne = int(sys.argv[1]) if len(sys.argv)>1 else 16
hb = bench.build_batch(ne, 1000, 3000, 16)
b = gtf_b200.EventBatch(hb)
b.seed(); b.cluster("track_state_estimates", 1.0, 2.0)
for _ in range(3): b.iterate_dry()
b.sync()
# timing of sub-programs through the stage entry points on copies of state is destructive; just time the fused one
b.set_timing(True)
for _ in range(5): b.iterate_dry()
print("prefix_ms, tile_ms, heavy_ms, n:", b.timing())
print("kernels:", b.timing_kernels())

import time
b.set_timing(False)
for _ in range(5): b.iterate_dry()
b.sync()
t0 = time.perf_counter()
n = 300
for _ in range(n): b.iterate_dry()
b.sync()
print("wall ms per uncommitted iteration (no timing events):", (time.perf_counter() - t0) / n * 1e3)
