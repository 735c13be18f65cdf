"""Helpers to load golden fixtures (tests/golden/*.npz) into host batches and compare states."""
import os
import sys
import numpy as np

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, REPO)
import gtf_b200  # noqa: E402,F401
from gtf_b200 import fields as F  # noqa: E402

GOLDEN = os.path.join(REPO, "tests", "golden")
STAGES = ("seed", "c1", "x1", "e2", "x2", "m2", "c3", "x3")
RTOL = 1e-9      # north_star: states / covariances / KL within 1e-9 relative (fp64)


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def stage_batch(fx, stage):
    """Host batch (dict) = static topology + the mutable arrays recorded after `stage`."""
    hb = {}
    for k in fx.files:
        if k.startswith("topo_"):
            hb[k[5:]] = fx[k]
        elif k.startswith(stage + "/"):
            hb[k[len(stage) + 1:]] = fx[k]
    for k in ("truth", "orig_id", "in_key", "accepted", "cand_label", "pvals"):
        hb.pop(k, None)
    return F.complete_host_batch(hb)


def rel_err(a, b, floor=None):
    """max relative error over entries where both are finite; inf if NaN patterns differ.
    `floor`: optional magnitude below which the error is measured against the floor instead of the value
    (a component that crosses zero has no meaningful element-wise relative error)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return np.inf
    m = ~na
    if not m.any():
        return 0.0
    d = np.abs(a[m] - b[m])
    s = np.maximum(np.abs(a[m]), np.abs(b[m]))
    if floor is not None:
        s = np.maximum(s, floor)
    s[s == 0] = 1.0
    return float(np.max(d / s))


def field_floor(x):
    """typical magnitude of a field: median |x| over its finite entries"""
    x = np.abs(np.asarray(x, np.float64))
    x = x[np.isfinite(x)]
    return float(np.median(x)) if x.size else None


def edge_exists(hb):
    src = hb["in_src"]
    return (hb["alive"][np.maximum(src, 0)] > 0) & (src >= 0) & (hb["alive"][hb["slot_dst"]] > 0)


def inplay_nodes(hb):
    return (hb["alive"] > 0) & (hb["sub_state"][hb["sub"]] == 0)


def dict_order(hb, key="uts"):
    """per-node tuple of slots in dict order, for order comparisons"""
    out = []
    for i in range(len(hb["x"])):
        s0, s1 = hb["in_off"][i], hb["in_off"][i + 1]
        sl = [s for s in range(s0, s1) if hb[key + "_present"][s]]
        if key == "uts":
            sl.sort(key=lambda s: hb["uts_rank"][s])
        out.append(tuple(sl))
    return out


def compare_states(got, want, what, rtol=RTOL, chained=None):
    """Compare two complete host batches on the reference-visible part of the state.
    Returns list of mismatch strings (empty = parity).  Per-stage comparisons (identical inputs) are
    strictly element-wise relative; chained comparisons (rtol > RTOL) measure components smaller than the
    field's typical magnitude against that magnitude."""
    if chained is None:
        chained = rtol > RTOL
    bad = []
    ex = edge_exists(want)
    live = inplay_nodes(want)
    live_slot = live[want["slot_dst"]]
    if "alive" in what:
        for f in ("alive", "sub_state"):
            if not np.array_equal(got[f], want[f]):
                bad.append("%s differs at %d places" % (f, int((got[f] != want[f]).sum())))
    if "active" in what:
        m = ex & live_slot
        if not np.array_equal(got["active"][m], want["active"][m]):
            bad.append("active differs on %d existing edges" % int((got["active"][m] != want["active"][m]).sum()))
    if "merged" in what:
        if not np.array_equal(got["has_merged"][live], want["has_merged"][live]):
            bad.append("has_merged differs at %d nodes" % int((got["has_merged"][live] != want["has_merged"][live]).sum()))
        m = live & (want["has_merged"] > 0) & (got["has_merged"] > 0)
        for f in ("m_a", "m_b", "m_c", "m_p00", "m_p01", "m_p11", "m_p22", "m_prior"):
            e = rel_err(got[f][m], want[f][m], field_floor(want[f][m]) if chained else None)
            if not e <= rtol:
                bad.append("%s rel err %.3g" % (f, e))
    for key in ("tse", "uts"):
        if key not in what:
            continue
        pm = live_slot
        if not np.array_equal(got[key + "_present"][pm], want[key + "_present"][pm]):
            bad.append("%s_present differs at %d slots" % (key, int((got[key + "_present"][pm] != want[key + "_present"][pm]).sum())))
        m = pm & (want[key + "_present"] > 0) & (got[key + "_present"] > 0)
        names = ["a", "b", "c", "tau", "p00", "p01", "p11", "p22", "prior", "w"]
        if key == "uts":
            names += ["lik", "lrn"]
        for f in names:
            e = rel_err(got["%s_%s" % (key, f)][m], want["%s_%s" % (key, f)][m],
                        field_floor(want["%s_%s" % (key, f)][m]) if chained else None)
            if not e <= rtol:
                bad.append("%s_%s rel err %.3g" % (key, f, e))
        if key == "uts":
            if not np.array_equal(got["uts_side"][m], want["uts_side"][m]):
                bad.append("uts_side differs")
            if not np.array_equal(got["has_uts"][live], want["has_uts"][live]):
                bad.append("has_uts differs")
            if dict_order(got) != dict_order(want):
                bad.append("uts dict order differs")
    if "degree" in what:
        if not np.array_equal(got["degree"][live], want["degree"][live]):
            bad.append("degree differs at %d nodes" % int((got["degree"][live] != want["degree"][live]).sum()))
    if "edge_w" in what:
        m = ex & live_slot
        e = rel_err(got["edge_w"][m], want["edge_w"][m])
        if not e <= rtol:
            bad.append("edge_w rel err %.3g" % e)
    bad += compare_frozen(got, want, what, rtol)
    return bad


def compare_frozen(got, want, what, rtol=RTOL):
    """Nodes that are NOT in play (extracted, or in a sub-graph that left the list as a fragment / empty:
    extract_track_candidates.py:460-467) are never touched again by the reference: their rows -- and the dict entries
    stored at them -- must still hold what they held when they left play."""
    bad = []
    dead = ~inplay_nodes(want)
    dead_slot = dead[want["slot_dst"]]
    if not dead.any():
        return bad
    node_f, slot_f = [], []
    if "merged" in what:
        node_f += ["has_merged", "m_a", "m_b", "m_c", "m_p00", "m_p01", "m_p11", "m_p22", "m_prior"]
    if "degree" in what:
        node_f += ["degree"]
    if "active" in what:
        slot_f += ["active"]
    if "edge_w" in what:
        slot_f += ["edge_w"]
    for key in ("tse", "uts"):
        if key in what:
            slot_f += [key + "_present"] + ["%s_%s" % (key, f) for f in ("a", "b", "c", "tau", "p00", "p01", "p11", "p22", "prior", "w")]
    if "uts" in what:
        slot_f += ["uts_lik", "uts_lrn", "uts_side"]
        node_f += ["has_uts"]
    for f in node_f:
        m = dead & (want["has_merged"] > 0) & (got["has_merged"] > 0) if f.startswith("m_") else dead
        if not _same(got[f][m], want[f][m], rtol):
            bad.append("out-of-play nodes: %s changed" % f)
    for f in slot_f:
        m = dead_slot
        if f.startswith(("tse_", "uts_")) and not f.endswith("_present"):
            pk = f[:3] + "_present"
            m = m & (want[pk] > 0) & (got[pk] > 0)
        if f in ("active", "edge_w"):
            m = m & edge_exists_at_exit(want)
        if not _same(got[f][m], want[f][m], rtol):
            bad.append("out-of-play slots: %s changed" % f)
    return bad


def edge_exists_at_exit(hb):
    """slots whose source is a real node (not a ghost row)"""
    return hb["in_src"] >= 0


def _same(a, b, rtol):
    if a.dtype.kind == "f":
        return rel_err(a, b) <= rtol
    return np.array_equal(a, b)
