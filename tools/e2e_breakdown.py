"""Wall-clock breakdown of the end-to-end step (host events -> candidate table) on one chunk: python tools/e2e_breakdown.py [events]"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
import gtf_b200, bench
ne = int(sys.argv[1]) if len(sys.argv) > 1 else 128
pool = bench.event_pool(1000, 16)
hb = bench.concat_events(pool, list(range(ne)))
c = bench.E2EChunk(hb, 0, torch)
b = c.b
def t(fn):
    b.sync(); t0 = time.perf_counter(); r = fn(); b.sync(); return (time.perf_counter() - t0) * 1e3, r
for rep in range(3):
    rows = []
    rows.append(("load_events", t(c.load)[0]))
    rows.append(("seed_cluster", t(lambda: b.seed_cluster(1.0, 2.0))[0]))
    ms, st = t(lambda: b.iterate(max_iter=10, stop_when_converged=True))
    rows.append(("iterate x%d" % len(st), ms))
    rows.append(("extract", t(lambda: b.extract(want_arrays=False))[0]))
    rows.append(("candidates", t(lambda: b.candidates_into(c.rows))[0]))
print("  ".join("%s %.2f" % r for r in rows), " total %.2f ms for %d events" % (sum(r[1] for r in rows), ne))
