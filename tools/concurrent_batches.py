"""Experiment: is one big batch faster or slower than the same events split over K batches iterated concurrently (one CUDA
stream each)?  python tools/concurrent_batches.py [events] [K ...]"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import gtf_b200, bench
ne = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ks = [int(v) for v in sys.argv[2:]] or [1, 2, 4]
pool = bench.event_pool(1000, 16)
for K in ks:
    bs = []
    for k in range(K):
        hb = bench.concat_events(pool, list(range(k, ne, K)))
        b = gtf_b200.EventBatch.with_capacity(len(hb["x"]), len(hb["in_src"]), len(hb["sub_event"]))
        b.load_events(hb)
        b.seed()
        b.cluster("track_state_estimates", 1.0, 2.0)
        bs.append(b)
    for _ in range(5):
        for b in bs:
            b.iterate_dry()
    for b in bs:
        b.sync()
    n = 200
    t0 = time.perf_counter()
    for _ in range(n):
        for b in bs:
            b.iterate_dry()
    for b in bs:
        b.sync()
    print("K=%d batches of %d events: %.4f ms per iteration of all %d events" % (K, ne // K, (time.perf_counter() - t0) / n * 1e3, ne))
    for b in bs:
        b.close()
