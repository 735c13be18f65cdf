"""Per-iteration kernel times of the committed loop (128-event batch): where do the later iterations spend their time?
  python tools/loop_profile.py [events] [iterations]"""
import sys, os, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import gtf_b200, bench
ne = int(sys.argv[1]) if len(sys.argv) > 1 else 128
nit = int(sys.argv[2]) if len(sys.argv) > 2 else 10
hb = bench.build_batch(ne, 1000, 3000, 16)
b = gtf_b200.EventBatch(hb)
for rep in range(2):           # second pass: warm
    b.seed_cluster(1.0, 2.0)
    rows = []
    for it in range(nit):
        b.set_timing(True)
        st = b.iterate(max_iter=1, stop_when_converged=False)[0]
        kt = b.timing_kernels()
        rows.append({"iteration": it + 1, "ms": {k: round(kt[k], 4) for k in ("k_send", "k_exec", "k_node2", "k_hv")},
                     "total_ms": round(sum(kt[k] for k in ("k_send", "k_exec", "k_node2", "k_hv")), 4),
                     "edges_sent": st["edges_sent"], "active_edges": st["active_edges"], "active_changed": st["active_changed"],
                     "nodes_merged": st["nodes_merged"]})
    b.set_timing(False)
for r in rows:
    print(json.dumps(r))
print(json.dumps({"loop_ms": round(sum(r["total_ms"] for r in rows), 3)}))

# the loop as a user runs it: one gtf_iterate call (device-side loop, sparse send from the second iteration on), wall clock
import time
for rep in range(3):
    b.seed_cluster(1.0, 2.0)
    b.sync()
    t0 = time.perf_counter()
    st = b.iterate(max_iter=nit, stop_when_converged=False)
    b.sync()
    ms = (time.perf_counter() - t0) * 1e3
print(json.dumps({"gtf_iterate_wall_ms": round(ms, 3), "iterations": len(st), "launches": b.iteration_launches()}))
