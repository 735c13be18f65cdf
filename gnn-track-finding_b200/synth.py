"""Seeded synthetic events.

Two generators, both deterministic in `seed` (numpy `default_rng`):

* `barrel_event`  -- TrackML-shaped 10-layer barrel event: helical tracks from a
  smeared beam-spot, doublet edges l->l+1 and l->l+2 inside a d-phi window and a
  z0-extrapolation window.  This is the cfg2/cfg3/cfg4 workload of BASELINE.json.
* `toy_event`     -- restatement of the reference's 2-D toy simulator
  (/root/reference/src/toyMC_model/track_simulation_xy.py:36-160): 88 straight
  radial tracks x 10 hits, Gaussian smear 0.05, edges by layer gap / distance /
  intercept cuts.  z = r = 0 there, so it is only an edge-topology generator.

An event is a plain dict of numpy arrays:
  x, y, z, r        f64[N]   hit coordinates (r = sqrt(x^2 + y^2))
  layer             i32[N]   in_volume_layer_id  (helper.py:18-19 semantics)
  volume            i32[N]   volume_id
  truth             i64[N]   truth particle label
  edge_a, edge_b    i32[M]   undirected doublets in list order; the graph gets
                             a->b then b->a per row (helper.py:517-518)
"""
import numpy as np

BARREL_RADII = np.array([32., 72., 116., 172., 260., 360., 500., 660., 820., 1020.])


def _window_pairs(phi_a, phi_b, w):
    """All (i, j) with |wrap(phi_b[j] - phi_a[i])| < w; vectorised via sort + searchsorted."""
    order = np.argsort(phi_b, kind="stable")
    pb = phi_b[order]
    # pad periodic images so a plain window search handles the wrap-around
    pb_ext = np.concatenate([pb - 2 * np.pi, pb, pb + 2 * np.pi])
    idx_ext = np.concatenate([order, order, order])
    lo = np.searchsorted(pb_ext, phi_a - w, side="right")
    hi = np.searchsorted(pb_ext, phi_a + w, side="left")
    cnt = np.maximum(hi - lo, 0)
    tot = int(cnt.sum())
    if tot == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    ia = np.repeat(np.arange(len(phi_a)), cnt)
    start = np.repeat(lo, cnt)
    within = np.arange(tot) - np.repeat(np.cumsum(cnt) - cnt, cnt)
    jb = idx_ext[start + within]
    return ia, jb


def barrel_event(n_tracks=1000, seed=0, eta_max=0.5, dphi_window=None, z0_window=200.0,
                 target_degree=10.0, n_layers=10):
    """One synthetic barrel event.

    True doublets (same track, layer gap 1 or 2) are always present; fake doublets are
    drawn from a d-phi/z0 window and subsampled so the mean in-degree is `target_degree`.
    """
    rng = np.random.default_rng(seed)
    radii = BARREL_RADII[:n_layers]
    pt = rng.uniform(1.0, 10.0, n_tracks)                 # GeV
    q = rng.choice(np.array([-1.0, 1.0]), n_tracks)
    R = pt / 0.6 * 1000.0                                 # mm, B = 2 T
    phi0 = rng.uniform(-np.pi, np.pi, n_tracks)
    eta = rng.uniform(-eta_max, eta_max, n_tracks)
    z0 = rng.normal(0.0, 30.0, n_tracks)

    rl = radii[None, :]                                   # (1, L)
    theta = 2.0 * np.arcsin(rl / (2.0 * R[:, None]))      # turning angle to reach radius r
    phi_hit = phi0[:, None] - q[:, None] * 0.5 * theta
    x = rl * np.cos(phi_hit)
    y = rl * np.sin(phi_hit)
    z = z0[:, None] + R[:, None] * theta * np.sinh(eta)[:, None]
    x = x + rng.normal(0.0, 0.05, x.shape)
    y = y + rng.normal(0.0, 0.05, y.shape)
    z = z + rng.normal(0.0, 0.05, z.shape)

    # node ids: layer-major so that ids of one layer are contiguous
    x = x.T.reshape(-1)
    y = y.T.reshape(-1)
    z = z.T.reshape(-1)
    r = np.sqrt(x * x + y * y)
    layer_idx = np.repeat(np.arange(n_layers), n_tracks)
    truth = np.tile(np.arange(n_tracks, dtype=np.int64), n_layers)
    phi = np.arctan2(y, x)

    # true doublets: same track, layer gap 1 or 2 (always kept)
    # fake doublets: random pairs inside a d-phi window and a z0 window, subsampled so that
    # the mean in-degree hits `target_degree` ("windows tuned so mean in-degree = target")
    n_nodes = n_tracks * n_layers
    n_true = n_tracks * ((n_layers - 1) + (n_layers - 2))
    n_fake_wanted = max(int(round(target_degree * n_nodes / 2.0)) - n_true, 0)
    n_pairs_layers = (n_layers - 1) + (n_layers - 2)
    if dphi_window is None:
        # candidates per layer pair ~ n_tracks^2 * w / pi * (z acceptance ~ 0.4); ask for ~3x the need
        per_pair = 3.0 * n_fake_wanted / max(n_pairs_layers, 1)
        dphi_window = min(max(per_pair * np.pi / (0.4 * max(n_tracks, 1) ** 2), 1e-4), 0.8)

    ea, eb, fa, fb = [], [], [], []
    for l in range(n_layers):
        a_ids = np.nonzero(layer_idx == l)[0]
        for gap in (1, 2):
            l2 = l + gap
            if l2 >= n_layers:
                continue
            b_ids = np.nonzero(layer_idx == l2)[0]
            ea.append(a_ids)                              # track t on layer l  -> track t on layer l2
            eb.append(b_ids)
            if n_fake_wanted == 0:
                continue
            ia, jb = _window_pairs(phi[a_ids], phi[b_ids], dphi_window)
            a = a_ids[ia]
            b = b_ids[jb]
            zz = z[a] - r[a] * (z[b] - z[a]) / (r[b] - r[a])   # doublet extrapolated to r = 0
            keep = (np.abs(zz) < z0_window) & (truth[a] != truth[b])
            fa.append(a[keep])
            fb.append(b[keep])
    if n_fake_wanted > 0:
        fa = np.concatenate(fa)
        fb = np.concatenate(fb)
        if len(fa) > n_fake_wanted:
            sel = np.sort(rng.choice(len(fa), n_fake_wanted, replace=False))
            fa, fb = fa[sel], fb[sel]
        ea.append(fa)
        eb.append(fb)
    ea = np.concatenate(ea)
    eb = np.concatenate(eb)
    order = np.lexsort((eb, ea))                          # list order: by first node, then second
    ea, eb = [ea[order]], [eb[order]]
    edge_a = np.concatenate(ea).astype(np.int32)
    edge_b = np.concatenate(eb).astype(np.int32)
    return {
        "x": x, "y": y, "z": z, "r": r,
        "layer": (2 * (layer_idx + 1)).astype(np.int32),
        "volume": np.full(x.shape, 8, np.int32),
        "truth": truth,
        "edge_a": edge_a, "edge_b": edge_b,
    }


def toy_event(seed=0, sigma0=0.05):
    """2-D toy event (track_simulation_xy.py:36-160), RNG seeded instead of global."""
    rng = np.random.default_rng(seed)
    num_hits, radius = 10, 10.0
    ends = []
    for i in [0.5, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9.5]:
        yy = np.sqrt(radius ** 2 - i ** 2)
        ends += [(i, yy), (yy, i), (-i, -yy), (-yy, -i), (i, -yy), (-yy, i), (-i, yy), (yy, -i)]
    ends = np.array(ends)
    n_tracks = len(ends)
    xs, ys, layers, truth = [], [], [], []
    for n in range(n_tracks):
        grad = ends[n, 1] / ends[n, 0]
        xx = np.linspace(0.0, ends[n, 0], num_hits)
        yy = grad * xx + sigma0 * rng.normal(0.0, 1.0, num_hits)
        xs.append(xx)
        ys.append(yy)
        layers.append(np.arange(num_hits))
        truth.append(np.full(num_hits, n))
    x = np.concatenate(xs)
    y = np.concatenate(ys)
    layer = np.concatenate(layers)
    truth = np.concatenate(truth).astype(np.int64)
    N = len(x)
    dx = x[None, :] - x[:, None]
    dy = y[None, :] - y[:, None]
    diff = np.abs(layer[None, :] - layer[:, None])
    with np.errstate(divide="ignore", invalid="ignore"):
        m = dy / dx
        c = y[None, :] - m * x[None, :]
        xint = -c / m
        ok = (diff > 0) & (diff <= 2) & (np.sqrt(dx ** 2 + dy ** 2) <= 3) \
            & (np.abs(c) <= 1.8) & (np.abs(xint) <= 1.8)
    ok &= ~np.eye(N, dtype=bool)
    alive = np.ones(N, bool)
    alive[np.arange(n_tracks) * num_hits] = False        # collision point removed (:134-135)
    ok &= alive[:, None] & alive[None, :]
    ii, jj = np.nonzero(ok & (np.abs(dx) > 0.75))        # nodes on long-dx edges dropped (:155-159)
    alive[ii] = False
    alive[jj] = False
    ok &= alive[:, None] & alive[None, :]
    ii, jj = np.nonzero(np.triu(ok | ok.T))
    remap = -np.ones(N, np.int64)
    remap[alive] = np.arange(int(alive.sum()))
    z = np.zeros(int(alive.sum()))
    return {
        "x": x[alive], "y": y[alive], "z": z, "r": z.copy(),
        "layer": layer[alive].astype(np.int32),
        "volume": np.zeros(int(alive.sum()), np.int32),
        "truth": truth[alive],
        "edge_a": remap[ii].astype(np.int32), "edge_b": remap[jj].astype(np.int32),
    }


def degree_stats(ev):
    """(n_nodes, n_directed_edges, mean in-degree, fraction of nodes with 3 <= d <= 15)."""
    n = len(ev["x"])
    deg = np.bincount(ev["edge_a"], minlength=n) + np.bincount(ev["edge_b"], minlength=n)
    return n, 2 * len(ev["edge_a"]), float(deg.mean()), float(((deg >= 3) & (deg <= 15)).mean())


# ------------------------------------------------------------------------------------------------
# native (networkx-free) construction of the flat layout for synthetic events


def reorder_slots_pyset(hb, node_ids):
    """Re-order every node's in-slots (= the key order of its seeded state dict) the way the reference does:
    helper.py:280 iterates `set(nx.all_neighbors(G, node))`, i.e. the order is CPython's set-iteration order of the
    neighbour IDS, filled from the predecessors (ascending node order inside the sub-graph copy, event_conversion.py:84)
    followed by the successors (adjacency insertion order), and the list is then REVERSED (helper.py:350-351) before the
    entries are inserted.  This runs under the same interpreter, so Python's own `set`
    gives exactly that order (host-side ingest logic; no arithmetic).  `node_ids[i]` = graph id of node row i."""
    N, E = len(hb["x"]), len(hb["in_src"])
    in_off, in_src = hb["in_off"], hb["in_src"]
    out_off, out_slot, slot_dst = hb["out_off"], hb["out_slot"], hb["slot_dst"]
    ids = [int(v) for v in node_ids]
    new_of_old = np.empty(E, np.int64)
    for v in range(N):
        s0, s1 = int(in_off[v]), int(in_off[v + 1])
        if s1 - s0 <= 1:
            new_of_old[s0:s1] = np.arange(s0, s1)
            continue
        srcs = in_src[s0:s1].tolist()
        slot_of = {ids[u]: s0 + k for k, u in enumerate(srcs)}
        succ = [ids[int(slot_dst[out_slot[o]])] for o in range(int(out_off[v]), int(out_off[v + 1]))]
        order = [k for k in set([ids[u] for u in sorted(srcs)] + succ) if k in slot_of][::-1]   # `keys.reverse()`, helper.py:351
        for k, key in enumerate(order):
            new_of_old[slot_of[key]] = s0 + k
    out = dict(hb)
    perm = np.empty(E, np.int64)          # perm[new] = old
    perm[new_of_old] = np.arange(E)
    out["in_src"] = in_src[perm]
    out["slot_dst"] = slot_dst[perm]
    out["out_slot"] = new_of_old[out_slot].astype(np.int32)
    rs = hb["rev_slot"][perm]
    out["rev_slot"] = np.where(rs >= 0, new_of_old[np.maximum(rs, 0)], -1).astype(np.int32)
    return out


def networkx_subgraph_order(n, src, dst, ids):
    """Node order of the sub-graph list `[G.subgraph(c).copy() for c in nx.weakly_connected_components(G)]`
    (event_conversion.py:84) without networkx: components in the order of their first node; inside a component networkx
    (3.x: `_plain_bfs` returns its `seen` set, `subgraph` filters with `set(nodes)` and iterates THAT set when it is smaller
    than half the graph) yields the nodes in CPython set-iteration order of the ids, inserted in BFS order (successors, then
    predecessors, adjacency insertion order).  Host-side ingest logic under the same interpreter: Python's own `set` is used.
    src / dst: directed edges in insertion order as node rows 0..n-1; ids[row] = graph id.  Returns (order, sub_of_row)."""
    succ = [[] for _ in range(n)]
    pred = [[] for _ in range(n)]
    for a, b in zip(src.tolist(), dst.tolist()):
        succ[a].append(b)
        pred[b].append(a)
    ids = [int(v) for v in ids]
    row_of = {k: r for r, k in enumerate(ids)}
    seen_all = np.zeros(n, bool)
    order, sub_of = [], np.zeros(n, np.int64)
    n_sub = 0
    for v in range(n):
        if seen_all[v]:
            continue
        seen = {ids[v]}
        nextlevel = [v]
        while nextlevel:
            thislevel, nextlevel = nextlevel, []
            for u in thislevel:
                for w in succ[u] + pred[u]:
                    if ids[w] not in seen:
                        seen.add(ids[w])
                        nextlevel.append(w)
        nodes = set(k for k in seen)                        # nx.filters.show_nodes: set(nbunch_iter(c)) -- filled one by one
                                                            # from a generator (set(seen) would copy the table layout)
        if 2 * len(nodes) < n:
            rows = [row_of[k] for k in nodes]
        else:
            rows = sorted(row_of[k] for k in nodes)         # FilterAtlas falls back to the graph's own node order
        seen_all[rows] = True
        sub_of[rows] = n_sub
        order += rows
        n_sub += 1
    return np.array(order, np.int64), sub_of


def event_to_host(ev, event_id=0, dict_order="insertion"):
    """Flat host batch (topology + hit arrays) of one synthetic event, no networkx involved.
    dict_order="pyset": the orders the REFERENCE's ingest produces for this event (networkx_subgraph_order for the nodes,
    reorder_slots_pyset for the state dicts); needs ev["node_idx"] when the graph ids are not 0..n-1.

    Orders are *defined* here (they are inputs of the algorithm, SURVEY.md §7 "Order as data"):
      * sub-graphs = weakly connected components, ordered by their smallest hit id; nodes inside a
        sub-graph in ascending hit id
      * successor order of u = order in which u's out-edges are first added when the doublet list is
        walked adding a->b then b->a (helper.py:517-518)
      * slot (state-dict) order at v = order in which v's in-edges are added by the same walk
    """
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    n = len(ev["x"])
    a = ev["edge_a"].astype(np.int64)
    b = ev["edge_b"].astype(np.int64)
    m = len(a)
    # directed edges in insertion order: (a0,b0),(b0,a0),(a1,b1),...
    src = np.empty(2 * m, np.int64)
    dst = np.empty(2 * m, np.int64)
    src[0::2], dst[0::2] = a, b
    src[1::2], dst[1::2] = b, a
    # drop duplicate directed edges, keeping the first insertion
    key = src * n + dst
    _, first = np.unique(key, return_index=True)
    keep = np.sort(first)
    src, dst = src[keep], dst[keep]
    ncomp, comp = connected_components(coo_matrix((np.ones(len(src), np.int8), (src, dst)), shape=(n, n)),
                                       directed=True, connection="weak")
    comp_min = np.full(ncomp, n, np.int64)
    np.minimum.at(comp_min, comp, np.arange(n))
    comp_rank = np.argsort(np.argsort(comp_min))          # sub-graph index by smallest member
    sub_of = comp_rank[comp]
    order = np.lexsort((np.arange(n), sub_of))            # new position -> old id
    if dict_order == "pyset":
        order, sub_of = networkx_subgraph_order(n, src, dst, ev["node_idx"] if "node_idx" in ev else np.arange(n))
    newpos = np.empty(n, np.int64)
    newpos[order] = np.arange(n)
    S = ncomp
    E = len(src)
    s_new, d_new = newpos[src], newpos[dst]
    # in-CSR: group by destination, stable in insertion order
    o_in = np.argsort(d_new, kind="stable")
    in_src = s_new[o_in]
    slot_dst = d_new[o_in]
    in_off = np.zeros(n + 1, np.int64)
    np.add.at(in_off, d_new + 1, 1)
    in_off = np.cumsum(in_off)
    slot_of_edge = np.empty(E, np.int64)
    slot_of_edge[o_in] = np.arange(E)
    # out-CSR: group by source, stable in insertion order
    o_out = np.argsort(s_new, kind="stable")
    out_slot = slot_of_edge[o_out]
    out_off = np.zeros(n + 1, np.int64)
    np.add.at(out_off, s_new + 1, 1)
    out_off = np.cumsum(out_off)
    # reverse slot: slot of (dst -> src)
    fwd_key = s_new * n + d_new
    rev_key = d_new * n + s_new
    sorter = np.argsort(fwd_key)
    pos = np.searchsorted(fwd_key[sorter], rev_key)
    pos = np.clip(pos, 0, E - 1) if E else pos
    found = (fwd_key[sorter][pos] == rev_key) if E else np.zeros(0, bool)
    rev_edge = np.where(found, sorter[pos], -1)
    rev_slot = np.full(E, -1, np.int64)
    rev_slot[slot_of_edge] = np.where(rev_edge >= 0, slot_of_edge[np.maximum(rev_edge, 0)], -1)
    sub_sorted = sub_of[order]
    sub_off = np.zeros(S + 1, np.int64)
    np.add.at(sub_off, sub_sorted + 1, 1)
    sub_off = np.cumsum(sub_off)
    hb = {
        "x": ev["x"][order], "y": ev["y"][order], "z": ev["z"][order], "r": ev["r"][order],
        "layer": ev["layer"][order].astype(np.int32), "volume": ev["volume"][order].astype(np.int32),
        "truth": ev["truth"][order], "orig_id": order.astype(np.int64),
        "sub": sub_sorted.astype(np.int32), "alive": np.ones(n, np.uint8),
        "sub_off": sub_off.astype(np.int32), "sub_state": np.zeros(S, np.uint8),
        "sub_event": np.full(S, event_id, np.int32),
        "in_off": in_off.astype(np.int32), "in_src": in_src.astype(np.int32), "slot_dst": slot_dst.astype(np.int32),
        "out_off": out_off.astype(np.int32), "out_slot": out_slot.astype(np.int32),
        "rev_slot": rev_slot.astype(np.int32),
    }
    if dict_order == "pyset":
        ids = ev["node_idx"] if "node_idx" in ev else np.arange(n)
        hb = reorder_slots_pyset(hb, np.asarray(ids)[order])
    return hb


_NODE_IDX = ("in_src", "slot_dst")
_SLOT_IDX = ("out_slot", "rev_slot")


def concat_host_batches(hbs):
    """Concatenate per-event host batches into one batch (events stay independent sub-graph sets)."""
    out = {}
    n_off = e_off = s_off = 0
    parts = {}
    for hb in hbs:
        N, E, S = len(hb["x"]), len(hb["in_src"]), len(hb["sub_off"]) - 1
        for k, v in hb.items():
            if k in ("in_off", "out_off"):
                v = v[:-1].astype(np.int64) + e_off
            elif k == "sub_off":
                v = v[:-1].astype(np.int64) + n_off
            elif k in _NODE_IDX:
                v = np.where(v >= 0, v.astype(np.int64) + n_off, -1)
            elif k in _SLOT_IDX:
                v = np.where(v >= 0, v.astype(np.int64) + e_off, -1)
            elif k == "sub":
                v = v.astype(np.int64) + s_off
            parts.setdefault(k, []).append(v)
        n_off, e_off, s_off = n_off + N, e_off + E, s_off + S
    for k, vs in parts.items():
        arr = np.concatenate(vs)
        if k in ("in_off", "out_off"):
            arr = np.concatenate([arr, [e_off]])
        elif k == "sub_off":
            arr = np.concatenate([arr, [n_off]])
        if k in ("in_off", "out_off", "sub_off", "sub", "in_src", "slot_dst", "out_slot", "rev_slot"):
            arr = arr.astype(np.int32)
        out[k] = arr
    return out
