// gtf_hostmath.cpp -- host build of gtf_math.cuh so the kernels' algebra can be unit-tested on a CPU-only
// machine against the oracle (tests/test_hostmath.py).  Not used by the product path: the CUDA library
// (libgtf_b200.so) compiles the same header for sm_100a.
#include "gtf_math.cuh"
#include <stdint.h>

extern "C" {
void gtfh_seed_entry(const double *node, const double *key, double tau, double var_tau_sq, const double *geom,
                     double *out8)
{
    GtfGeom g{geom[0], geom[1], geom[2], geom[3]};
    GtfState s;
    gtf_seed_entry(node[0], node[1], node[2], node[3], key[0], key[1], key[2], key[3], tau, var_tau_sq, g, s);
    out8[0] = s.a; out8[1] = s.b; out8[2] = s.c; out8[3] = s.tau;
    out8[4] = s.p00; out8[5] = s.p01; out8[6] = s.p11; out8[7] = s.p22;
}
double gtfh_var_ms(double a, double b, double xk, double dr, double dz, double zside, double endcap)
{
    return gtf_var_ms(a, b, xk, dr, dz, zside, endcap);
}
// merged7 = a b c p00 p01 p11_eff p22 ; out = chi2 lik pass a b c tau p00 p01 p11 p22
void gtfh_extrapolate(const double *u, const double *v, const double *merged7, double var_ms, double chi2_cut,
                      const double *geom, double *out11)
{
    GtfGeom g{geom[0], geom[1], geom[2], geom[3]};
    GtfExtrapOut o;
    o.lik = 0; o.s = GtfState{0, 0, 0, 0, 0, 0, 0, 0};
    gtf_extrapolate(u[0], u[1], u[2], u[3], v[0], v[1], v[2], v[3], merged7[0], merged7[1], merged7[2], merged7[3],
                    merged7[4], merged7[5], merged7[6], var_ms, chi2_cut, g, o);
    out11[0] = o.chi2; out11[1] = o.lik; out11[2] = o.pass;
    out11[3] = o.s.a; out11[4] = o.s.b; out11[5] = o.s.c; out11[6] = o.s.tau;
    out11[7] = o.s.p00; out11[8] = o.s.p01; out11[9] = o.s.p11; out11[10] = o.s.p22;
}
static GtfState st(const double *p) { return GtfState{p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7]}; }
double gtfh_pair_chi2(const double *si, const double *sj, const double *node, const double *nbi, const double *nbj,
                      const double *geom)
{
    GtfGeom g{geom[0], geom[1], geom[2], geom[3]};
    return gtf_pair_chi2(st(si), st(sj), node[0], node[2], node[3], nbi[0], nbi[2], nbi[3], nbj[0], nbj[2], nbj[3], g);
}
void gtfh_merge(const double *s1, const double *s2, double *out8)
{
    GtfState m;
    gtf_merge(st(s1), st(s2), m);
    out8[0] = m.a; out8[1] = m.b; out8[2] = m.c; out8[3] = m.tau;
    out8[4] = m.p00; out8[5] = m.p01; out8[6] = m.p11; out8[7] = m.p22;
}
double gtfh_kl(const double *s1, const double *s2) { return gtf_kl(st(s1), st(s2)); }
void gtfh_seed_parabolic(const double *node_xy, const double *nbr_xy, double s0, double sA, double sB, double *sv3, double *cov9)
{
    gtf_seed_parabolic(node_xy[0], node_xy[1], nbr_xy[0], nbr_xy[1], s0, sA, sB, sv3, cov9);
}
double gtfh_kl_general(const double *m1, const double *c1, const double *m2, const double *c2) { return gtf_kl_general(m1, c1, m2, c2); }
}

extern "C" void gtfh_track_fit(double *co4, int n, double sigma0xy, double sigma0rz, double endcap, double sep3d,
                               double *pv2)
{
    gtf_track_fit((double (*)[4])co4, n, sigma0xy, sigma0rz, endcap, sep3d, pv2[0], pv2[1]);
}

// greedy clustering of one node in information form (mirrors node_cluster's loop in gtf_tile.cuh):
// states8[n][8], prior[n], coords of node and neighbours; returns merged8, merged prior, remaining mask, or -1
extern "C" int gtfh_cluster_node(const double *states8, const double *prior, int n, const double *node,
                                 const double *nbc4, double chi2_thr, double kl_thr, const double *geom,
                                 double *merged8, double *mprior, unsigned *rem_out)
{
    GtfGeom g{geom[0], geom[1], geom[2], geom[3]};
    double best = INFINITY;
    int bi = -1, bj = -1;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < i; j++) {
            double v = gtf_pair_chi2(st(states8 + 8 * i), st(states8 + 8 * j), node[0], node[2], node[3], nbc4[4 * i],
                                     nbc4[4 * i + 2], nbc4[4 * i + 3], nbc4[4 * j], nbc4[4 * j + 2], nbc4[4 * j + 3], g);
            if (v != 0.0 && v < best) { best = v; bi = i; bj = j; }
        }
    if (bi < 0 || !(best < chi2_thr)) return 0;
    GtfInfo M, t;
    gtf_to_info(st(states8 + 8 * bi), M);
    gtf_to_info(st(states8 + 8 * bj), t);
    gtf_info_add(M, t);
    GtfState m;
    gtf_from_info(M, m);
    double mp = prior[bi] + prior[bj];
    unsigned rem = ((1u << n) - 1u) & ~((1u << bi) | (1u << bj));
    while (rem) {
        double bv = INFINITY;
        int bk = -1;
        for (int k = 0; k < n; k++)
            if ((rem >> k) & 1u) {
                GtfInfo ei;
                gtf_to_info(st(states8 + 8 * k), ei);
                double kl = gtf_kl_info(st(states8 + 8 * k), ei, m, M);
                if (kl < bv) { bv = kl; bk = k; }
            }
        if (bk < 0 || !(bv < kl_thr)) break;
        gtf_to_info(st(states8 + 8 * bk), t);
        gtf_info_add(M, t);
        gtf_from_info(M, m);
        mp = prior[bk] + mp;
        rem &= ~(1u << bk);
    }
    merged8[0] = m.a; merged8[1] = m.b; merged8[2] = m.c; merged8[3] = m.tau;
    merged8[4] = m.p00; merged8[5] = m.p01; merged8[6] = m.p11; merged8[7] = m.p22;
    *mprior = mp;
    *rem_out = rem;
    return 1;
}
