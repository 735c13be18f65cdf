"""tests/golden/helpers.npz: inputs and outputs of the reference's stand-alone helper functions
(clustering/clustering.py:11-124, extrapolate/extrapolate_merged_states.py:26), called UNMODIFIED.  Build-container only."""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
warnings.filterwarnings("ignore")
import ref_harness as rh  # noqa: E402

rh.setup_reference()
import golden_util as gu  # noqa: E402
from gtf_b200 import nxio  # noqa: E402


def spd3(rng, block):
    a = rng.normal(size=(3, 3))
    c = a @ a.T + 0.1 * np.eye(3)
    if block:
        c[0:2, 2] = 0.0
        c[2, 0:2] = 0.0
    return c * rng.uniform(1e-4, 1.0)


def main():
    from clustering import clustering as cl
    from extrapolate import extrapolate_merged_states as em
    rng = np.random.default_rng(77)
    data = {}
    n = 7
    svs = rng.normal(size=(n, 3)) * [1e-3, 1e-1, 1.0]
    covs = np.array([spd3(rng, True) for _ in range(n)])
    node = np.array([30.0, -12.0, 15.0, 32.3])
    nbrs = np.c_[rng.uniform(500, 600, n) * rng.choice([-1, 1], n), rng.normal(size=n) * 50, rng.normal(size=n) * 300, rng.uniform(60, 900, n)]
    nbrs[::2, 0] = rng.normal(size=len(nbrs[::2])) * 40            # both sides of abs(x) >= endcap_boundary (clustering.py:49-57)
    with rh.quiet("/tmp"):
        data["chi2_matrix"] = cl.calc_pairwise_distances_chi2(n, svs, covs, node, nbrs, 0.4, 0.6, 550.0)
        data["chi2_pair"] = np.array(cl.mahalanobis_distance(svs[2], covs[2], svs[5], covs[5], node, nbrs[2], nbrs[5], 0.4, 0.6, 550.0))
    data.update(svs=svs, covs=covs, node=node, nbrs=nbrs)
    gm = rng.normal(size=(n + 1, 3))
    gc = np.array([spd3(rng, False) for _ in range(n + 1)])
    mm, mc = cl.merge_states(gm[0], gc[0], gm[1], gc[1])
    data.update(gm=gm, gc=gc, merged_mean=mm, merged_cov=mc, kl=np.array(cl.KLDistance(gm[0], gc[0], gm[1], gc[1])),
                kl_to_merged=np.array(cl.calc_dist_to_merged_state(n - 1, gm[2:], gc[2:], mm, mc)))
    # extrapolate_validate on edges of the barrel40_eta1 fixture after the first extraction (merged states present)
    fx = gu.load("barrel40_eta1")
    hb = gu.stage_batch(fx, "x1")
    graphs = nxio.host_to_graphs(hb, orig_id=fx["topo_orig_id"], truth=fx["topo_truth"])
    rows = []
    with rh.quiet("/tmp"):
        for g in graphs:
            for u, attr in g.nodes(data=True):
                if "merged_state" not in attr:
                    continue
                for v in g.neighbors(u):
                    if g[u][v]["activated"] != 1 or len(rows) >= 60:
                        continue
                    cov = np.array(attr["merged_cov"], copy=True)
                    cov_in = cov.copy()
                    gg = g.copy()
                    cut = 2.0 if len(rows) % 2 == 0 else 0.02
                    out = em.extrapolate_validate(gg, u, gg.nodes[u], v, gg.nodes[v], cut, np.array(attr["merged_state"]), cov,
                                                  0.3, 0.4, 0.6, 550.0)
                    assert (d_ := out[0]) is not None or gg[u][v]["activated"] == 0
                    d = out[0]
                    a, b_ = attr["GNN_Measurement"], g.nodes[v]["GNN_Measurement"]
                    row = [a.x, a.y, a.z, a.r, b_.x, b_.y, b_.z, b_.r] + list(attr["merged_state"]) + list(cov_in.reshape(-1)) + \
                          [cov[1, 1], out[1], 0.0 if d is None else 1.0]
                    if d is not None:
                        c2 = d["joint_vector_covariance"]
                        row += list(d["edge_state_vector"]) + [d["joint_vector"][2], c2[0, 0], c2[0, 1], c2[1, 1], c2[2, 2], d["likelihood"],
                                                               d["mixture_weight"]]
                    else:
                        row += [np.nan] * 10
                    rows.append(row + [cut])
    data["edges"] = np.array(rows)
    path = os.path.join(HERE, "helpers.npz")
    np.savez_compressed(path, **data)
    e = data["edges"]
    print("helpers.npz: %d edges (%d pass, %d gated) -> %.1f KB" % (len(e), int(e[:, 22].sum()), int((e[:, 22] == 0).sum()), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
