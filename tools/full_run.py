"""wall-clock of a complete reconstruction of a batch of cfg2 events on one GPU: upload, seed, cluster on the seeds,
iterate until converged, components + candidate extraction, candidate table.  Usage: python tools/full_run.py [events]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench, gtf_b200
ne = int(sys.argv[1]) if len(sys.argv) > 1 else 128
hb = bench.build_batch(ne, 1000, 3000, 16)
for rep in range(3):
    t = [time.perf_counter()]
    b = gtf_b200.EventBatch(hb); b.sync(); t.append(time.perf_counter())
    b.seed(); b.sync(); t.append(time.perf_counter())
    b.cluster("track_state_estimates", 1.0, 2.0); t.append(time.perf_counter())
    st = b.iterate(max_iter=10); t.append(time.perf_counter())
    n_acc = b.extract(want_arrays=False)[0]; t.append(time.perf_counter())
    rows = b.candidates(); t.append(time.perf_counter())
    names = ["create+upload", "seed", "cluster(seeds)", "iterate x%d" % len(st), "extract", "candidate table"]
    d = [(t[i + 1] - t[i]) * 1e3 for i in range(len(names))]
    dev = sum(d[1:])
    print("events %d  " % ne + "  ".join("%s %.2f ms" % (n, v) for n, v in zip(names, d)) +
          "  | device pipeline %.2f ms = %.0f events/s, %d accepted nodes, active per iteration %s" %
          (dev, ne / dev * 1e3, n_acc, [s["active_edges"] for s in st]))
    b.close()
