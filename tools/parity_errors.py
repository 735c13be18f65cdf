"""Per-stage, per-field maximum relative error of the chained schedule (seeded once, every stage fed by the previous one)
against the reference fixtures: the oracle on the CPU, or the GPU path with --gpu.   python tools/parity_errors.py [--gpu] [fixture ...]
Decisions (activation flags, merged-state existence, candidate sets) are asserted bit-exact; the table shows how far the
fp64 VALUES drift (two different but equally valid operation orders through 2x2 / 3x3 inverses with cond up to 7e6)."""
import os, sys
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np
import golden_util as gu, oracle_lib as ol
from test_oracle_golden import blank_seed

gpu = "--gpu" in sys.argv
names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["barrel25_deg6", "barrel40_eta1", "barrel60_deg16", "barrel100_cfg1", "barrel1000_cfg2", "shipped_vol79"]
FIELDS = ("m_a", "m_b", "m_c", "m_p00", "m_p01", "m_p11", "m_p22", "m_prior")


def errs(got, want):
    live = gu.inplay_nodes(want) & (want["has_merged"] > 0) & (got["has_merged"] > 0)
    assert np.array_equal(got["active"][gu.edge_exists(want)], want["active"][gu.edge_exists(want)])
    assert np.array_equal(got["has_merged"][gu.inplay_nodes(want)], want["has_merged"][gu.inplay_nodes(want)])
    out = {}
    for f in FIELDS:
        out[f] = (gu.rel_err(got[f][live], want[f][live]), gu.rel_err(got[f][live], want[f][live], gu.field_floor(want[f][live])))
    return out


for name in names:
    fx = gu.load(name)
    hb = blank_seed(gu.stage_batch(fx, "seed"))
    if gpu:
        import gtf_b200
        b = gtf_b200.EventBatch(hb)
        state = lambda: b.download()          # noqa: E731
        steps = [("c1", lambda: b.cluster(0, 1.0, 2.0)), ("x1", lambda: b.extract()), ("e2", lambda: b.extrapolate_stage(2.0)),
                 ("x2", lambda: b.extract()), ("m2", lambda: b.remove_state_metadata()), ("c3", lambda: b.cluster(1, 1000.0, 100.0))]
        b.seed()
    else:
        ob = ol.OracleBatch(hb)
        state = lambda: ob.hb                 # noqa: E731
        steps = [("c1", lambda: ob.cluster(0, 1.0, 2.0)), ("x1", lambda: ob.extract()), ("e2", lambda: ob.extrapolate_stage(2.0)),
                 ("x2", lambda: ob.extract()), ("m2", lambda: ob.remove_state_metadata()), ("c3", lambda: ob.cluster(1, 1000.0, 100.0))]
        ob.seed()
    print("%s (%s vs reference): max relative error of the merged states, element-wise / against the field's median magnitude" % (
        name, "GPU" if gpu else "oracle"))
    for stage, fn in steps:
        fn()
        if stage.startswith("x") or stage == "m2":
            continue
        e = errs(state(), gu.stage_batch(fx, stage))
        print("  %-3s " % stage + "  ".join("%s %.1e/%.1e" % (f[2:], a, b_) for f, (a, b_) in e.items()))
