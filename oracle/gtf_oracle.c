/* gtf_oracle.c -- CPU ORACLE (test infrastructure; see gtf_oracle.h).
 *
 * Literal restatement of the reference's Python hot path.  Every function cites the reference
 * file:line (relative to /root/reference/src) it follows.  Arithmetic is written in the reference's
 * operation order with libm transcendentals and general (pivoted) matrix inverses, i.e. it is
 * deliberately NOT the CUDA kernels' algebra: the two implementations are independent.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off).  Single-threaded like the reference; callers that
 * want all host cores run independent event batches on Python threads (ctypes drops the GIL).
 */
#include "gtf_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAXD 15 /* cluster() only handles 3..15 components (clustering.py:207) */

/* ---------------------------------------------------------------- small dense algebra */

/* general n x n inverse (n <= 3) by Gauss-Jordan with partial pivoting -- stands in for
 * np.linalg.inv (LAPACK getrf/getri); same pivoting rule, results equal up to rounding */
static void inv_n(const double *a, double *out, int n)
{
    double m[3][6];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            m[i][j] = a[i * n + j];
            m[i][n + j] = (i == j) ? 1.0 : 0.0;
        }
    for (int c = 0; c < n; c++) {
        int p = c;
        for (int i = c + 1; i < n; i++)
            if (fabs(m[i][c]) > fabs(m[p][c])) p = i;
        if (p != c)
            for (int j = 0; j < 2 * n; j++) {
                double t = m[c][j];
                m[c][j] = m[p][j];
                m[p][j] = t;
            }
        double piv = m[c][c];
        for (int j = 0; j < 2 * n; j++) m[c][j] /= piv;
        for (int i = 0; i < n; i++)
            if (i != c) {
                double f = m[i][c];
                if (f != 0.0)
                    for (int j = 0; j < 2 * n; j++) m[i][j] -= f * m[c][j];
            }
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) out[i * n + j] = m[i][n + j];
}

static void mat3_mul(const double *a, const double *b, double *c)
{
    double t[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s += a[i * 3 + k] * b[k * 3 + j];
            t[i * 3 + j] = s;
        }
    memcpy(c, t, sizeof t);
}
static void mat3_mul_bt(const double *a, const double *b, double *c) /* a * b^T */
{
    double t[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s += a[i * 3 + k] * b[j * 3 + k];
            t[i * 3 + j] = s;
        }
    memcpy(c, t, sizeof t);
}
static void mat3_vec(const double *a, const double *v, double *o)
{
    double t[3];
    for (int i = 0; i < 3; i++) t[i] = a[i * 3] * v[0] + a[i * 3 + 1] * v[1] + a[i * 3 + 2] * v[2];
    o[0] = t[0];
    o[1] = t[1];
    o[2] = t[2];
}
static void cov_from4(double p00, double p01, double p11, double p22, double *c)
{
    c[0] = p00; c[1] = p01; c[2] = 0.0;
    c[3] = p01; c[4] = p11; c[5] = 0.0;
    c[6] = 0.0; c[7] = 0.0; c[8] = p22;
}

/* ---------------------------------------------------------------- clustering.py:11-124 helpers */

/* clustering.py:90-94 KLDistance.  NOTE `(cov1 - cov2) * (inv2 - inv1)` is an ELEMENT-WISE product
 * (numpy ndarray `*`); the trace therefore only sees the diagonals.  Pinned by the shipped golden CSV. */
double gtfo_kl_distance(const double m1[3], const double c1[9], const double m2[3], const double c2[9])
{
    double i1[9], i2[9];
    inv_n(c1, i1, 3);
    inv_n(c2, i2, 3);
    double tr = 0.0;
    for (int k = 0; k < 3; k++) tr += (c1[k * 4] - c2[k * 4]) * (i2[k * 4] - i1[k * 4]);
    double d[3] = {m1[0] - m2[0], m1[1] - m2[1], m1[2] - m2[2]};
    double s[9], t[3];
    for (int k = 0; k < 9; k++) s[k] = i1[k] + i2[k];
    /* (mean1 - mean2).T.dot(inv1 + inv2).dot(mean1 - mean2): row-vector times matrix first */
    for (int j = 0; j < 3; j++) t[j] = d[0] * s[j] + d[1] * s[3 + j] + d[2] * s[6 + j];
    return tr + (t[0] * d[0] + t[1] * d[1] + t[2] * d[2]);
}

/* clustering.py:97-105 merge_states (inverse-variance weighting) */
void gtfo_merge_states(const double m1[3], const double c1[9], const double m2[3], const double c2[9],
                       double mm[3], double mc[9])
{
    double i1[9], i2[9], s[9], a[3], b[3], v[3];
    inv_n(c1, i1, 3);
    inv_n(c2, i2, 3);
    for (int k = 0; k < 9; k++) s[k] = i1[k] + i2[k];
    inv_n(s, mc, 3);
    mat3_vec(i1, m1, a);
    mat3_vec(i2, m2, b);
    v[0] = a[0] + b[0];
    v[1] = a[1] + b[1];
    v[2] = a[2] + b[2];
    mat3_vec(mc, v, mm);
}

/* clustering.py:11-78 mahalanobis_distance.  node = a, neighbour1 = b, neighbour2 = c.
 * The endcap test looks at abs(x) (clustering.py:49-57), not abs(z): reproduced. */
double gtfo_mahalanobis(const double m1[3], const double c1[9], const double m2[3], const double c2[9],
                        const double node[4], const double nb1[4], const double nb2[4], double sigma0rz,
                        double sigma0rz2, double endcap)
{
    double res[2] = {m1[0] - m2[0], m1[1] - m2[1]};
    double cd[4] = {c1[0] + c2[0], c1[1] + c2[1], c1[3] + c2[3], c1[4] + c2[4]};
    double ic[4];
    inv_n(cd, ic, 2);
    double t0 = res[0] * ic[0] + res[1] * ic[2];
    double t1 = res[0] * ic[1] + res[1] * ic[3];
    double distance1 = t0 * res[0] + t1 * res[1];

    double x_a = node[0], x_b = nb1[0], x_c = nb2[0];
    double z_a = node[2], r_a = node[3];
    double z_b = nb1[2], r_b = nb1[3];
    double z_c = nb2[2], r_c = nb2[3];
    double j2 = 1 / (r_b - r_a);
    double j3 = -1 / (r_c - r_a);
    double j1 = -j3 - j2;
    double j5 = -(z_b - z_a) / ((r_b - r_a) * (r_b - r_a));
    double j6 = (z_c - z_a) / ((r_c - r_a) * (r_c - r_a));
    double j4 = -j5 - j6;
    double J[6] = {j1, j2, j3, j4, j5, j6};
    double sza = sigma0rz2, szb = sigma0rz2, szc = sigma0rz2;
    double sra = sigma0rz, srb = sigma0rz, src = sigma0rz;
    if (fabs(x_a) >= endcap) { sza = sigma0rz; sra = sigma0rz2; }
    if (fabs(x_b) >= endcap) { szb = sigma0rz; srb = sigma0rz2; }
    if (fabs(x_c) >= endcap) { szc = sigma0rz; src = sigma0rz2; }
    double Sd[6] = {sza * sza, szb * szb, szc * szc, sra * sra, srb * srb, src * src};
    double cov_delta_tau = 0.0;
    for (int k = 0; k < 6; k++) cov_delta_tau += (J[k] * Sd[k]) * J[k];
    double inv_cov = 1 / cov_delta_tau;
    double tau1 = (z_b - z_a) / (r_b - r_a);
    double tau2 = (z_c - z_a) / (r_c - r_a);
    double r2 = tau1 - tau2;
    double distance2 = (r2 * r2) * inv_cov;
    return distance1 + distance2;
}

/* ---------------------------------------------------------------- graph helpers */

static inline int edge_exists(const gtfo_arrays *A, int s)
{
    int u = A->in_src[s], v = A->slot_dst[s];
    return u >= 0 && A->alive[u] && A->alive[v];
}
static int sub_alive_count(const gtfo_arrays *A, int g)
{
    int c = 0;
    for (int i = A->sub_off[g]; i < A->sub_off[g + 1]; i++) c += A->alive[i];
    return c;
}

/* dict order of a node's state dict: TSE = slot order; UTS = ascending uts_rank. Returns count. */
static int dict_slots(const gtfo_arrays *A, int node, int key, int *slots, int cap)
{
    int n = 0, s0 = A->in_off[node], s1 = A->in_off[node + 1];
    if (key == GTF_KEY_TSE) {
        for (int s = s0; s < s1; s++)
            if (A->tse_present[s]) {
                if (n < cap) slots[n] = s;
                n++;
            }
        return n;
    }
    for (int s = s0; s < s1; s++)
        if (A->uts_present[s]) {
            if (n < cap) {
                int k = n; /* insertion sort by rank */
                while (k > 0 && A->uts_rank[slots[k - 1]] > A->uts_rank[s]) {
                    slots[k] = slots[k - 1];
                    k--;
                }
                slots[k] = s;
            }
            n++;
        }
    return n;
}
static int has_dict(const gtfo_arrays *A, int node, int key)
{
    return key == GTF_KEY_TSE ? 1 : A->has_uts[node];
}

/* ---------------------------------------------------------------- helper.py:24-94 */

int gtfo_initialize_edge_activation(gtfo_arrays *A) /* helper.py:24-25 */
{
    for (int s = 0; s < A->E; s++) A->active[s] = 1;
    return 0;
}

int gtfo_compute_prior_probabilities(gtfo_arrays *A, int key) /* helper.py:30-63 */
{
    double *prior = key == GTF_KEY_TSE ? A->tse_prior : A->uts_prior;
    const uint8_t *present = key == GTF_KEY_TSE ? A->tse_present : A->uts_present;
    for (int g = 0; g < A->S; g++) {
        if (A->sub_state[g] != GTF_SUB_INPLAY) continue;
        if (sub_alive_count(A, g) == 1) continue; /* :33 */
        for (int i = A->sub_off[g]; i < A->sub_off[g + 1]; i++) {
            if (!A->alive[i] || !has_dict(A, i, key)) continue;
            int s0 = A->in_off[i], s1 = A->in_off[i + 1];
            for (int s = s0; s < s1; s++) {
                if (!present[s] || !edge_exists(A, s) || A->active[s] != 1) continue;
                int lay = A->layer[A->in_src[s]], cnt = 0;
                for (int t = s0; t < s1; t++)
                    if (present[t] && edge_exists(A, t) && A->active[t] == 1 && A->layer[A->in_src[t]] == lay)
                        cnt++;
                prior[s] = 1.0 / cnt; /* :61 */
            }
        }
    }
    return 0;
}

int gtfo_compute_mixture_weights(gtfo_arrays *A, int key) /* helper.py:76-94 */
{
    int err = 0;
    double *w = key == GTF_KEY_TSE ? A->tse_w : A->uts_w;
    const uint8_t *present = key == GTF_KEY_TSE ? A->tse_present : A->uts_present;
    for (int g = 0; g < A->S; g++) {
        if (A->sub_state[g] != GTF_SUB_INPLAY) continue;
        if (sub_alive_count(A, g) == 1) continue; /* :79 */
        for (int i = A->sub_off[g]; i < A->sub_off[g + 1]; i++) {
            if (!A->alive[i] || !has_dict(A, i, key)) continue;
            int n = 0;
            for (int s = A->in_off[i]; s < A->in_off[i + 1]; s++) n += present[s];
            if (n == 0) { err |= GTFO_ERR_ZERO_DIV; continue; } /* 1/len({}) :90 */
            double mw = 1.0 / n;
            for (int s = A->in_off[i]; s < A->in_off[i + 1]; s++)
                if (present[s]) w[s] = mw;
        }
    }
    return err;
}

int gtfo_query_node_degree(gtfo_arrays *A) /* helper.py:67-73 applied to every node (clustering.py:324-327) */
{
    for (int g = 0; g < A->S; g++) {
        if (A->sub_state[g] != GTF_SUB_INPLAY) continue;
        for (int i = A->sub_off[g]; i < A->sub_off[g + 1]; i++) {
            if (!A->alive[i]) continue;
            int d = 0;
            for (int s = A->in_off[i]; s < A->in_off[i + 1]; s++)
                if (edge_exists(A, s) && A->active[s] == 1) d++;
            A->degree[i] = d;
        }
    }
    return 0;
}

/* ---------------------------------------------------------------- helper.py:238-452 seeding */

static double tau_cov(double z1, double r1, double z2, double r2, double sz, double szn, double sr, double srn,
                      double prefix)
{
    /* helper.py:317-330 (prefix = 1) and :339-345 (theta variant) */
    double j1 = prefix / (r1 - r2);
    double j2 = -prefix / (r1 - r2);
    double j3 = (-prefix * (z1 - z2)) / ((r1 - r2) * (r1 - r2));
    double j4 = (prefix * (z1 - z2)) / ((r1 - r2) * (r1 - r2));
    double J[4] = {j1, j2, j3, j4};
    double S2[4] = {sz * sz, szn * szn, sr * sr, srn * srn};
    double c = 0.0;
    for (int k = 0; k < 4; k++) c += (J[k] * S2[k]) * J[k];
    return c;
}

int gtfo_seed(gtfo_arrays *A, const gtfo_geom *g) /* helper.py:238-452 */
{
    const double sigmaO = 4.0, sigmaA = g->sigma0xy, sigmaB = g->sigma0xy; /* :243-245 */
    const double Sd[3] = {sigmaO * sigmaO, sigmaA * sigmaA, sigmaB * sigmaB};
    for (int i = 0; i < A->N; i++) {
        int s0 = A->in_off[i], d = A->in_off[i + 1] - s0;
        double xA = A->x[i], yA = A->y[i], zA = A->z[i], rA = A->r[i];
        double sigma_r = g->sigma0rz, sigma_z = g->sigma0rz2;
        if (fabs(zA) >= g->endcap_boundary) { sigma_z = g->sigma0rz; sigma_r = g->sigma0rz2; }
        double az = atan2(yA, xA), ca = cos(az), sa = sin(az);
        double x_0 = (0.0 - xA) * ca + (0.0 - yA) * sa; /* origin in the node frame (:363,371) */
        /* gradients over neighbours in set-iteration order N[i'] = key of slot d-1-i' */
        double gsum = 0.0;
        for (int q = 0; q < d; q++) {
            int nb = A->in_src[s0 + d - 1 - q];
            gsum += (A->y[nb] - yA) / (A->x[nb] - xA);
        }
        double gmean = d ? gsum / d : NAN, gvar = 0.0;
        for (int q = 0; q < d; q++) {
            int nb = A->in_src[s0 + d - 1 - q];
            double t = (A->y[nb] - yA) / (A->x[nb] - xA) - gmean;
            gvar += t * t;
        }
        A->emp_var[i] = d ? gvar / d : NAN; /* np.var, :446 */
        for (int k = 0; k < d; k++) {
            int s = s0 + k, key = A->in_src[s];
            /* quirk 5: tau / var(tau) stored under this key belong to the neighbour of slot d-1-k */
            int other = A->in_src[s0 + d - 1 - k];
            double z2 = A->z[other], r2 = A->r[other];
            double tau = (z2 - zA) / (r2 - rA); /* :302 */
            double szn = g->sigma0rz2, srn = g->sigma0rz;
            if (fabs(z2) >= g->endcap_boundary) { szn = g->sigma0rz; srn = g->sigma0rz2; }
            double cov_tau = tau_cov(zA, rA, z2, r2, sigma_z, szn, sigma_r, srn, 1.0);
            double variance_tau = cov_tau * cov_tau; /* :421 squares it */
            /* parabola through origin, node, key (:375-389) */
            double xk = A->x[key], yk = A->y[key], zk = A->z[key], rk = A->r[key];
            double x_B = (xk - xA) * ca + (yk - yA) * sa;
            double m_B = -(xk - xA) * sa + (yk - yA) * ca;
            double H[9] = {0.5 * (x_0 * x_0), x_0, 1, 0.0, 0.0, 1, 0.5 * (x_B * x_B), x_B, 1};
            double Hi[9];
            inv_n(H, Hi, 3);
            double meas[3] = {0.0, 0.0, m_B}, sv[3];
            mat3_vec(Hi, meas, sv);
            double a = sv[0], b = sv[1];
            double dr = rA - rk, dz = zA - zk; /* :402-403 */
            double hyp = sqrt(dr * dr + dz * dz);
            double sin_t = fabs(dr) / hyp;
            double kappa = (2 * a) / pow(1 + ((2 * a * xk) + b) * ((2 * a * xk) + b), 1.5);
            double q = (13.6 * 1e-3 * sqrt(0.02) * kappa) / 0.3;
            double var_ms = sin_t * (q * q);
            if (fabs(zA) >= g->endcap_boundary) var_ms = var_ms * fabs(dr / dz); /* :412-415 */
            double HS[9], cov[9];
            for (int r_ = 0; r_ < 3; r_++)
                for (int c_ = 0; c_ < 3; c_++) HS[r_ * 3 + c_] = Hi[r_ * 3 + c_] * Sd[c_];
            mat3_mul_bt(HS, Hi, cov);
            cov[4] += var_ms; /* :418 */
            A->tse_present[s] = 1;
            A->tse_a[s] = sv[0];
            A->tse_b[s] = sv[1];
            A->tse_c[s] = sv[2];
            A->tse_tau[s] = tau;
            A->tse_p00[s] = cov[0];
            A->tse_p01[s] = cov[1];
            A->tse_p11[s] = cov[4];
            A->tse_p22[s] = variance_tau + var_ms; /* :425 */
        }
    }
    return 0;
}

/* ---------------------------------------------------------------- clustering.py:149-376 cluster() */

static void load_entry(const gtfo_arrays *A, int key, int s, double par[3], double joint[3], double cov[9],
                       double *prior)
{
    if (key == GTF_KEY_TSE) {
        par[0] = A->tse_a[s]; par[1] = A->tse_b[s]; par[2] = A->tse_c[s];
        joint[0] = par[0]; joint[1] = par[1]; joint[2] = A->tse_tau[s];
        cov_from4(A->tse_p00[s], A->tse_p01[s], A->tse_p11[s], A->tse_p22[s], cov);
        *prior = A->tse_prior[s];
    } else {
        par[0] = A->uts_a[s]; par[1] = A->uts_b[s]; par[2] = A->uts_c[s];
        joint[0] = par[0]; joint[1] = par[1]; joint[2] = A->uts_tau[s];
        cov_from4(A->uts_p00[s], A->uts_p01[s], A->uts_p11[s], A->uts_p22[s], cov);
        *prior = A->uts_prior[s];
    }
}

/* one node of the loop at clustering.py:193-307.  Returns error bits; writes merged state and marks
 * deactivate[slot]=1 for the un-absorbed components. */
static int cluster_node(gtfo_arrays *A, int key, int node, double chi2_thr, double kl_thr, const gtfo_geom *g,
                        uint8_t *deactivate, gtfo_stats *st)
{
    int slots[MAXD + 1];
    int n = dict_slots(A, node, key, slots, MAXD + 1);
    if (n <= 2 || n >= 16) return 0; /* :207 */
    double par[MAXD][3], joint[MAXD][3], cov[MAXD][9], prior[MAXD], nbc[MAXD][4];
    int keyslot[MAXD];
    double nc[4] = {A->x[node], A->y[node], A->z[node], A->r[node]};
    for (int k = 0; k < n; k++) {
        load_entry(A, key, slots[k], par[k], joint[k], cov[k], &prior[k]);
        int nb = A->in_src[slots[k]];
        nbc[k][0] = A->x[nb]; nbc[k][1] = A->y[nb]; nbc[k][2] = A->z[nb]; nbc[k][3] = A->r[nb];
        keyslot[k] = slots[k];
    }
    /* :80-86 lower-triangular chi2 matrix, zeros elsewhere */
    double D[MAXD][MAXD];
    memset(D, 0, sizeof D);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < i; j++)
            D[i][j] = gtfo_mahalanobis(joint[i], cov[i], joint[j], cov[j], nc, nbc[i], nbc[j], g->sigma0rz,
                                       g->sigma0rz2, g->endcap_boundary);
    /* :119-123 min over non-zero entries, np.where(== min) in row-major order */
    int have = 0, isnan_ = 0;
    double smallest = 0.0;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            double v = D[i][j];
            if (v == 0.0) continue; /* np.nonzero drops +-0; NaN is kept */
            if (v != v) isnan_ = 1;
            if (!have) { smallest = v; have = 1; }
            else if (v < smallest) smallest = v;
        }
    if (!have) return GTFO_ERR_EMPTY_MIN;
    if (isnan_) return 0; /* np.min -> nan; `nan < thr` is False -> "No clusters found" (:304) */
    int rows[MAXD * MAXD], cols[MAXD * MAXD], nm = 0;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++)
            if (D[i][j] == smallest) { rows[nm] = i; cols[nm] = j; nm++; }
    if (!(smallest < chi2_thr)) return 0; /* :228 */
    int idx0 = rows[0], idx1 = nm > 1 ? rows[1] : cols[0]; /* idx = concat(rows, cols); idx[0], idx[1] */
    uint8_t gone[MAXD];
    memset(gone, 0, sizeof gone);
    for (int k = 0; k < nm; k++) { gone[rows[k]] = 1; gone[cols[k]] = 1; } /* np.delete(.., idx) */

    double pm[3], pc[9], jm[3], jc[9];
    gtfo_merge_states(par[idx0], cov[idx0], par[idx1], cov[idx1], pm, pc);     /* :231 */
    gtfo_merge_states(joint[idx0], cov[idx0], joint[idx1], cov[idx1], jm, jc); /* :233 */
    double merged_prior = prior[idx0] + prior[idx1];                           /* :234 */

    int rem[MAXD], nr = 0;
    for (int k = 0; k < n; k++)
        if (!gone[k]) rem[nr++] = k;
    int err = 0;
    if (nr == 0) {
        err = GTFO_ERR_EMPTY_MIN; /* np.min([]) at :252 */
    } else {
        for (;;) {
            /* :107-112 + :114-117: list form, np.min then list.index (first occurrence) */
            double best = 0.0;
            int bi = -1, nanseen = 0;
            for (int k = 0; k < nr; k++) {
                double dkl = gtfo_kl_distance(joint[rem[k]], cov[rem[k]], jm, jc);
                if (dkl != dkl) nanseen = 1;
                if (bi < 0 || dkl < best) { best = dkl; bi = k; }
            }
            if (nanseen) { err = GTFO_ERR_NAN_INDEX; break; }
            double thr = kl_thr;
            if (!(best < thr)) break; /* :261 */
            int e = rem[bi];
            double npm[3], npc[9], njm[3], njc[9];
            gtfo_merge_states(par[e], cov[e], pm, pc, npm, npc);     /* :263 */
            gtfo_merge_states(joint[e], cov[e], jm, jc, njm, njc);   /* :265 */
            memcpy(pm, npm, sizeof pm); memcpy(pc, npc, sizeof pc);
            memcpy(jm, njm, sizeof jm); memcpy(jc, njc, sizeof jc);
            merged_prior = prior[e] + merged_prior;                  /* :266 */
            for (int k = bi; k + 1 < nr; k++) rem[k] = rem[k + 1];
            nr--;
            if (nr == 0) break; /* :283 */
        }
    }
    if (err) return err;
    A->has_merged[node] = 1; /* :291-293 */
    A->m_a[node] = pm[0]; A->m_b[node] = pm[1]; A->m_c[node] = pm[2];
    A->m_p00[node] = pc[0]; A->m_p01[node] = pc[1]; A->m_p11[node] = pc[4]; A->m_p22[node] = pc[8];
    A->m_prior[node] = merged_prior;
    if (st) st->nodes_merged++;
    for (int k = 0; k < nr; k++) deactivate[keyslot[rem[k]]] = 1; /* :297-302 */
    return 0;
}

int gtfo_cluster(gtfo_arrays *A, int key, double chi2_thr, double kl_thr, const double *kl_lut,
                 const gtfo_geom *g, gtfo_stats *st)
{
    int err = 0;
    uint8_t *deact = (uint8_t *)calloc(A->E ? A->E : 1, 1);
    for (int gph = 0; gph < A->S; gph++) {
        if (A->sub_state[gph] != GTF_SUB_INPLAY) continue;
        for (int i = A->sub_off[gph]; i < A->sub_off[gph + 1]; i++) {
            if (!A->alive[i] || !has_dict(A, i, key)) continue;
            if (st) st->nodes_with_state++;
            double thr = kl_thr;
            if (kl_lut) { /* LUT mode (SURVEY.md 8c last row): per-node threshold from emp_var bin */
                int bin = (int)floor(A->emp_var[i] / 0.05);
                if (bin < 0) bin = 0;
                if (bin > 27 || A->emp_var[i] != A->emp_var[i]) bin = 27;
                thr = kl_lut[bin];
            }
            err |= cluster_node(A, key, i, chi2_thr, thr, g, deact, st);
        }
    }
    /* :311-321 simultaneous deactivation */
    for (int s = 0; s < A->E; s++)
        if (deact[s] && edge_exists(A, s)) {
            A->active[s] = 0;
            if (st) st->edges_deactivated++;
        }
    free(deact);
    gtfo_query_node_degree(A);                     /* :324-327 */
    err |= gtfo_compute_mixture_weights(A, key);   /* :372 */
    gtfo_compute_prior_probabilities(A, key);      /* :373 */
    return err;
}

/* ---------------------------------------------------------------- extrapolate_merged_states.py */

/* extrapolate_validate (:26-402) for the edge node(u) -> neighbour(v) stored in slot s */
static int extrapolate_validate(gtfo_arrays *A, int u, int v, int s, double chi2_cut, const gtfo_geom *g,
                                gtfo_stats *st)
{
    double node_x = A->x[u], node_y = A->y[u], node_z = A->z[u], node_r = A->r[u];
    double nb_x = A->x[v], nb_y = A->y[v], nb_z = A->z[v], nb_r = A->r[v];
    double ang = atan2(node_y, node_x); /* :41 */
    double x_A = (nb_x - node_x) * cos(ang) + (nb_y - node_y) * sin(ang);  /* :52 */
    double y_A = -(nb_x - node_x) * sin(ang) + (nb_y - node_y) * cos(ang); /* :53 */
    double a = A->m_a[u], b = A->m_b[u], c = A->m_c[u];
    double phi = atan2((node_x * nb_y) - (node_y * nb_x), (node_x * nb_x) + (node_y * nb_y)); /* :59 */
    double sp = sin(phi), cp = cos(phi);
    double x_prime = x_A + (c * sp);
    double Vx_prime = cp + (b * sp);
    double Ax_prime = a * sp;
    double s_star = (-x_prime * ((2 * (Vx_prime * Vx_prime)) + (Ax_prime * x_prime))) / (2 * pow(Vx_prime, 3)); /* :68 */
    (void)y_A; /* y', Vy', Ay', a_c, b_c, y_c (:71-79) never reach an output */
    double numer = x_A + c * sp, denom = cp + b * sp; /* :82-86 */
    double ds_da = -(sp * (numer * numer)) / pow(denom, 3);
    double ds_db = ((sp * numer) * (1 + ((3 * a * sp * numer) / (denom * denom)))) / (denom * denom);
    double ds_dc = -sp * (1 + ((2 * a * sp * numer) / (denom * denom))) / denom;
    denom = cp + ((2 * a + b) * sp); /* :89-92 */
    double da_da = (1 / pow(denom, 3)) * (1 - ((6 * a * sp) * (s_star + a * ds_da) / denom));
    double da_db = (-3 * a * sp * ((2 * a * ds_db) + 1)) / pow(denom, 4);
    double da_dc = (-6 * sp * ds_dc * (a * a)) / pow(denom, 4);
    denom = cp + ((2 * a * s_star + b) * sp); /* :95-99 */
    double bracket = cp - ((sp * (-sp + ((2 * a * s_star + b) * cp))) / denom);
    double db_da = (2 * (s_star + a * ds_da) * bracket) / denom;
    double db_db = ((1 + (2 * a * ds_da)) * bracket) / denom;
    double db_dc = (2 * a * ds_dc * bracket) / denom;
    bracket = (cp * (2 * a + b)) - sp; /* :102-105 */
    double dc_da = (ds_da * bracket) + ((s_star * s_star) * cp);
    double dc_db = (ds_db * bracket) + (s_star * cp);
    double dc_dc = (ds_dc * bracket) + cp;
    double F[9] = {da_da, da_db, da_dc, db_da, db_db, db_dc, dc_da, dc_db, dc_dc};

    double dr = nb_r - node_r, dz = nb_z - node_z; /* :114-124 */
    double hyp = sqrt(dr * dr + dz * dz);
    double sin_t = fabs(dr) / hyp;
    double kb = (2 * a * nb_x) + b;
    double kappa = (2 * a) / pow(1 + kb * kb, 1.5);
    double q = (13.6 * 1e-3 * sqrt(0.02) * kappa) / 0.3;
    double var_ms = sin_t * (q * q);
    if (fabs(node_z) >= g->endcap_boundary) var_ms = var_ms * (fabs(dr) / fabs(dz));

    A->m_p11[u] += var_ms; /* :127-128: aliases the node attribute, accumulates over successors */
    double P[9];
    cov_from4(A->m_p00[u], A->m_p01[u], A->m_p11[u], A->m_p22[u], P);
    double xs[3] = {a, b, c}, xe[3], FP[9], Pe[9];
    mat3_vec(F, xs, xe);     /* :129 */
    mat3_mul(F, P, FP);      /* :130 */
    mat3_mul_bt(FP, F, Pe);
    double residual = 0.0 - xe[2];                               /* :137 */
    double S = Pe[8] + g->sigma0xy * g->sigma0xy;                /* :138 */
    double inv_S = 1 / S;
    double chi2 = (residual * inv_S) * residual;                 /* :140 */
    A->uts_chi2[s] = chi2;
    if (st) st->edges_sent++;
    if (!(chi2 <= chi2_cut)) { /* :298, :393 */
        A->active[s] = 0;
        if (st) st->edges_gated++;
        return 0;
    }
    double factor = 2 * M_PI * fabs(S); /* :302-304 */
    double likelihood = pow(factor, -0.5) * exp(-0.5 * chi2);
    /* filterpy KalmanFilter: predict() then update(0.0) (:307-323) */
    double R = g->sigma0xy * g->sigma0xy;
    double xp[3], Pp[9];
    mat3_vec(F, xe, xp);
    mat3_mul(F, Pe, FP);
    mat3_mul_bt(FP, F, Pp);
    Pp[4] += var_ms; /* + Q, only Q[1][1] is non-zero (adding 0.0 elsewhere is exact) */
    double yres = 0.0 - xp[2];
    double PHT[3] = {Pp[2], Pp[5], Pp[8]};
    double Sk = PHT[2] + R, SI = 1 / Sk;
    double K[3] = {PHT[0] * SI, PHT[1] * SI, PHT[2] * SI};
    double xu[3] = {xp[0] + K[0] * yres, xp[1] + K[1] * yres, xp[2] + K[2] * yres};
    double IKH[9] = {1, 0, 0 - K[0], 0, 1, 0 - K[1], 0, 0, 1 - K[2]};
    double T[9], Pu[9];
    mat3_mul(IKH, Pp, T);
    mat3_mul_bt(T, IKH, Pu);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Pu[i * 3 + j] += (K[i] * R) * K[j];
    /* :326-358 tau and its variance */
    double tau = dz / dr;
    double sigma_r = g->sigma0rz, sigma_z = g->sigma0rz2;
    if (fabs(node_z) >= g->endcap_boundary) { sigma_z = g->sigma0rz; sigma_r = g->sigma0rz2; }
    double srn = g->sigma0rz, szn = g->sigma0rz2;
    if (fabs(nb_z) >= g->endcap_boundary) { szn = g->sigma0rz; srn = g->sigma0rz2; }
    double J[4] = {1 / dr, -1 / dr, -dz / (dr * dr), dz / (dr * dr)};
    double S2[4] = {sigma_z * sigma_z, szn * szn, sigma_r * sigma_r, srn * srn};
    double variance_tau = 0.0;
    for (int k = 0; k < 4; k++) variance_tau += (J[k] * S2[k]) * J[k];
    /* :361-385 store at the receiver */
    int rs = A->rev_slot[s];
    if (rs < 0 || !A->tse_present[rs]) return GTFO_ERR_NO_TSE;
    if (!A->uts_present[s]) { /* :443-447 dict insertion */
        A->uts_present[s] = 1;
        A->uts_rank[s] = A->uts_next[v]++;
        A->has_uts[v] = 1;
    }
    A->uts_a[s] = xu[0]; A->uts_b[s] = xu[1]; A->uts_c[s] = xu[2]; A->uts_tau[s] = tau;
    A->uts_p00[s] = Pu[0]; A->uts_p01[s] = Pu[1]; A->uts_p11[s] = Pu[4];
    A->uts_p22[s] = variance_tau + var_ms;
    A->uts_lik[s] = likelihood;
    A->uts_w[s] = A->tse_w[rs]; /* :384 */
    A->uts_prior[s] = NAN;      /* a fresh dict has no 'prior' / 'lr_layer_norm' / 'side' yet */
    A->uts_lrn[s] = NAN;
    A->uts_side[s] = 0;
    return 0;
}

int gtfo_message_passing(gtfo_arrays *A, double chi2_cut, const gtfo_geom *g, gtfo_stats *st) /* :406-447 */
{
    int err = 0;
    for (int gph = 0; gph < A->S; gph++) {
        if (A->sub_state[gph] != GTF_SUB_INPLAY) continue;
        if (sub_alive_count(A, gph) == 1) continue; /* :416 */
        for (int u = A->sub_off[gph]; u < A->sub_off[gph + 1]; u++) {
            if (!A->alive[u] || !A->has_merged[u]) continue; /* :425 */
            for (int o = A->out_off[u]; o < A->out_off[u + 1]; o++) { /* :430 successors in adjacency order */
                int s = A->out_slot[o], v = A->slot_dst[s];
                if (!A->alive[v]) continue;       /* removed nodes are gone from the adjacency */
                if (A->active[s] != 1) continue;  /* :431 */
                err |= extrapolate_validate(A, u, v, s, chi2_cut, g, st);
            }
        }
    }
    return err;
}

/* ---------------------------------------------------------------- helper.py:99-200 reweight */

int gtfo_reweight(gtfo_arrays *A, int key, double threshold, gtfo_stats *st)
{
    int err = 0;
    if (key != GTF_KEY_UTS) return 0; /* 'likelihood' only exists on updated_track_states entries */
    for (int gph = 0; gph < A->S; gph++) {
        if (A->sub_state[gph] != GTF_SUB_INPLAY) continue;
        if (sub_alive_count(A, gph) == 1) continue; /* :154 */
        for (int i = A->sub_off[gph]; i < A->sub_off[gph + 1]; i++) {
            if (!A->alive[i] || !A->has_uts[i]) continue; /* :160 */
            int deg = A->in_off[i + 1] - A->in_off[i];
            int *slots = (int *)malloc(sizeof(int) * (deg + 1));
            int n = dict_slots(A, i, key, slots, deg + 1);
            /* calculate_side_norm_factor (:99-139) */
            double node_x = A->x[i];
            int nl = 0, nrr = 0;
            int *left = (int *)malloc(sizeof(int) * (n + 1)), *right = (int *)malloc(sizeof(int) * (n + 1));
            for (int k = 0; k < n; k++) {
                int s = slots[k];
                if (edge_exists(A, s) && A->active[s] == 1) {
                    if (A->x[A->in_src[s]] < node_x) left[nl++] = s; else right[nrr++] = s;
                }
            }
            int left_norm = 0, right_norm = 0; /* len(set(coords)): distinct x values */
            for (int p = 0; p < nl; p++) {
                int dup = 0;
                for (int q2 = 0; q2 < p; q2++)
                    if (A->x[A->in_src[left[q2]]] == A->x[A->in_src[left[p]]]) dup = 1;
                left_norm += !dup;
            }
            for (int p = 0; p < nrr; p++) {
                int dup = 0;
                for (int q2 = 0; q2 < p; q2++)
                    if (A->x[A->in_src[right[q2]]] == A->x[A->in_src[right[p]]]) dup = 1;
                right_norm += !dup;
            }
            if (nl + nrr > 0) {
                int last = slots[n - 1]; /* stale `neighbour_num` = last key iterated (:131,138) */
                int last_active = 0;
                if (!edge_exists(A, last)) err |= GTFO_ERR_KEY; else last_active = (A->active[last] == 1);
                for (int p = 0; p < nl; p++) {
                    A->uts_side[left[p]] = 1;
                    A->uts_lrn[left[p]] = last_active ? left_norm : 1;
                }
                for (int p = 0; p < nrr; p++) {
                    A->uts_side[right[p]] = 2;
                    A->uts_lrn[right[p]] = last_active ? right_norm : 1;
                }
            }
            /* :165-169 */
            double denom = 0;
            for (int k = 0; k < n; k++) {
                int s = slots[k];
                if (edge_exists(A, s) && A->active[s] == 1) denom += (A->uts_w[s] * A->uts_lik[s]);
            }
            /* :172-195 */
            for (int k = 0; k < n; k++) {
                int s = slots[k];
                if (!(edge_exists(A, s) && A->active[s] == 1)) continue;
                double rw = (A->uts_w[s] * A->uts_lik[s] * A->uts_prior[s]) / denom;
                rw /= A->uts_lrn[s];
                A->uts_w[s] = rw;
                A->edge_w[s] = rw;
                if (rw < threshold) {
                    A->active[s] = 0;
                    if (st) st->edges_reweight_off++;
                } else
                    A->active[s] = 1;
            }
            free(slots); free(left); free(right);
        }
    }
    return err;
}

/* ---------------------------------------------------------------- update/remove_state_metadata.py:31-53 */

int gtfo_remove_state_metadata(gtfo_arrays *A, gtfo_stats *st)
{
    for (int gph = 0; gph < A->S; gph++) {
        if (A->sub_state[gph] != GTF_SUB_INPLAY) continue;
        for (int i = A->sub_off[gph]; i < A->sub_off[gph + 1]; i++) {
            if (!A->alive[i]) continue;
            int key = A->has_uts[i] ? GTF_KEY_UTS : GTF_KEY_TSE; /* :35-38 */
            for (int s = A->in_off[i]; s < A->in_off[i + 1]; s++) {
                /* key `sn` survives iff it is still a successor of the node (:42-44); edges are
                 * bidirectional by construction (helper.py:517-518) so that is "sn alive" + out-edge */
                int sn = A->in_src[s];
                int is_succ = 0;
                if (sn >= 0 && A->alive[sn])
                    for (int o = A->out_off[i]; o < A->out_off[i + 1]; o++)
                        if (A->slot_dst[A->out_slot[o]] == sn) { is_succ = 1; break; }
                if (is_succ) continue;
                if (key == GTF_KEY_UTS) A->uts_present[s] = 0; else A->tse_present[s] = 0;
            }
        }
    }
    gtfo_compute_prior_probabilities(A, GTF_KEY_TSE); /* :51-53 */
    gtfo_compute_prior_probabilities(A, GTF_KEY_UTS);
    return gtfo_reweight(A, GTF_KEY_UTS, 0.1, st);
}

/* ---------------------------------------------------------------- extract_track_candidates.py */

static int uf_find(int32_t *p, int i)
{
    while (p[i] != i) { p[i] = p[p[i]]; i = p[i]; }
    return i;
}

int gtfo_cca(gtfo_arrays *A) /* extract...py:332-346 */
{
    int32_t *p = A->label;
    for (int i = 0; i < A->N; i++) p[i] = A->alive[i] ? i : -1;
    for (int gph = 0; gph < A->S; gph++) {
        if (A->sub_state[gph] != GTF_SUB_INPLAY) continue;
        int b = A->sub_off[gph], e = A->sub_off[gph + 1];
        int any_inactive = 0;
        for (int i = b; i < e && !any_inactive; i++)
            for (int s = A->in_off[i]; s < A->in_off[i + 1]; s++)
                if (edge_exists(A, s) && A->active[s] == 0) { any_inactive = 1; break; }
        if (!any_inactive) { /* :343-344 the whole (possibly disconnected) graph is one candidate */
            int first = -1;
            for (int i = b; i < e; i++)
                if (A->alive[i]) { if (first < 0) first = i; p[i] = first; }
            continue;
        }
        for (int i = b; i < e; i++)
            for (int s = A->in_off[i]; s < A->in_off[i + 1]; s++)
                if (edge_exists(A, s) && A->active[s] != 0) {
                    int ra = uf_find(p, A->in_src[s]), rb = uf_find(p, i);
                    if (ra < rb) p[rb] = ra; else if (rb < ra) p[ra] = rb;
                }
        for (int i = b; i < e; i++)
            if (A->alive[i]) p[i] = uf_find(p, i);
    }
    return 0;
}

/* regularised upper incomplete gamma Q(a, x): series below a+1, Lentz continued fraction above.
 * Stands in for scipy.stats.distributions.chi2.sf (extract...py:321,325): sf(x, k) = Q(k/2, x/2) */
static double gammq(double a, double x)
{
    if (x != x || a != a) return NAN;
    if (x <= 0.0) return 1.0;
    if (isinf(x)) return 0.0;
    double gln = lgamma(a);
    if (x < a + 1.0) {
        double ap = a, sum = 1.0 / a, del = sum;
        for (int n = 0; n < 10000; n++) {
            ap += 1.0;
            del *= x / ap;
            sum += del;
            if (fabs(del) < fabs(sum) * 1e-17) break;
        }
        return 1.0 - sum * exp(-x + a * log(x) - gln);
    }
    double tiny = 1e-300, b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
    for (int i = 1; i < 10000; i++) {
        double an = -i * (i - a);
        b += 2.0;
        d = an * d + b;
        if (fabs(d) < tiny) d = tiny;
        c = b + an / c;
        if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < 1e-16) break;
    }
    return exp(-x + a * log(x) - gln) * h;
}
double gtfo_chi2_sf(double x, double k) { return gammq(0.5 * k, 0.5 * x); }

/* KF_track_fit_moliere (extract...py:209-328); coords ordered outermost -> innermost, already rotated */
static void kf_track_fit(const double (*co)[4], int n, double sigma0xy, double sigma0rz, double endcap,
                         double *pval, double *pval_zr)
{
    double fx[3] = {co[0][1], 0., 0.};
    double fP[9] = {sigma0xy * sigma0xy, 0, 0, 0, 1., 0, 0, 0, 1.};
    double fR = sigma0xy * sigma0xy;
    double gx[2] = {co[0][3], 0.};
    double gP[4] = {sigma0rz * sigma0rz, 0., 0., 1000.};
    double gR = sigma0rz * sigma0rz;
    double chi2_xy = 0.0, chi2_zr = 0.0; /* sum() starts at int 0 and adds left to right */
    for (int i = 0; i < n - 1; i++) {
        double x1 = .0, y1 = .0, x2 = co[i][0], y2 = co[i][1], x3 = co[i + 1][0], y3 = co[i + 1][1];
        double den = (x1 - x2) * (x1 - x3) * (x2 - x3); /* :202-204 */
        double a = ((x3 * (y2 - y1)) + (x2 * (y1 - y3)) + (x1 * (y3 - y2))) / den;
        double b = (((x3 * x3) * (y1 - y2)) + ((x2 * x2) * (y3 - y1)) + ((x1 * x1) * (y2 - y3))) / den;
        double z2 = co[i][2], r2 = co[i][3], z3 = co[i + 1][2], r3 = co[i + 1][3];
        double dr = r3 - r2, dz = z3 - z2;
        double hyp = sqrt(dr * dr + dz * dz), sin_t = fabs(dr) / hyp;
        double kb = (2 * a * x3) + b;
        double kappa = (2 * a) / pow(1 + kb * kb, 1.5);
        double q = (13.6 * 1e-3 * sqrt(0.02) * kappa) / 0.3;
        double var_ms = sin_t * (q * q);
        if (fabs(z3) >= endcap) var_ms = var_ms * fabs(dr / dz);
        double dx = x3 - x2, alpha = 0.1;
        double e1 = exp(-fabs(dx) * alpha), f1 = (1.0 - e1) / alpha, g1 = (fabs(dx) - f1) / alpha;
        double sigma_ou = 0.00001, sw2 = sigma_ou * sigma_ou, st2 = var_ms, dx2 = dx * dx, dxw2 = dx2 * sw2;
        double Q02 = 0.5 * dxw2, Q01 = dx * (st2 + Q02), Q12 = dx * sw2;
        double F[9] = {1., dx, g1, 0., 1., f1, 0., 0., e1};
        double Q[9] = {dx2 * (st2 + 0.25 * dxw2), Q01, Q02, Q01, st2 + dxw2, Q12, Q02, Q12, sw2};
        /* predict + update (xy) */
        double xp[3], FP[9], Pp[9];
        mat3_vec(F, fx, xp);
        mat3_mul(F, fP, FP);
        mat3_mul_bt(FP, F, Pp);
        for (int k = 0; k < 9; k++) Pp[k] += Q[k];
        double meas = co[i + 1][1];
        double yr = meas - xp[0];
        double PHT[3] = {Pp[0], Pp[3], Pp[6]};
        double S = PHT[0] + fR, SI = 1 / S;
        double K[3] = {PHT[0] * SI, PHT[1] * SI, PHT[2] * SI};
        for (int k = 0; k < 3; k++) fx[k] = xp[k] + K[k] * yr;
        double IKH[9] = {1 - K[0], 0, 0, 0 - K[1], 1, 0, 0 - K[2], 0, 1};
        double T[9];
        mat3_mul(IKH, Pp, T);
        mat3_mul_bt(T, IKH, fP);
        for (int r_ = 0; r_ < 3; r_++)
            for (int c_ = 0; c_ < 3; c_++) fP[r_ * 3 + c_] += (K[r_] * fR) * K[c_];
        double res = meas - fx[0]; /* :292-296 post-fit residual */
        double S2 = fP[0] + fR;
        chi2_xy += (res * (1 / S2)) * res;
        /* zr filter (:299-316); g.Q is a scalar -> broadcast onto all four entries */
        double gxp[2] = {gx[0] + dz * gx[1], gx[1]};
        double GF[4] = {1., dz, 0., 1.};
        double GFP[4] = {GF[0] * gP[0] + GF[1] * gP[2], GF[0] * gP[1] + GF[1] * gP[3],
                         GF[2] * gP[0] + GF[3] * gP[2], GF[2] * gP[1] + GF[3] * gP[3]};
        double gPp[4] = {GFP[0] * GF[0] + GFP[1] * GF[1] + var_ms, GFP[0] * GF[2] + GFP[1] * GF[3] + var_ms,
                         GFP[2] * GF[0] + GFP[3] * GF[1] + var_ms, GFP[2] * GF[2] + GFP[3] * GF[3] + var_ms};
        double gm = co[i + 1][3];
        double gy = gm - gxp[0];
        double gPHT[2] = {gPp[0], gPp[2]};
        double gS = gPHT[0] + gR, gSI = 1 / gS;
        double gK[2] = {gPHT[0] * gSI, gPHT[1] * gSI};
        gx[0] = gxp[0] + gK[0] * gy;
        gx[1] = gxp[1] + gK[1] * gy;
        double gI[4] = {1 - gK[0], 0, 0 - gK[1], 1};
        double gT[4] = {gI[0] * gPp[0] + gI[1] * gPp[2], gI[0] * gPp[1] + gI[1] * gPp[3],
                        gI[2] * gPp[0] + gI[3] * gPp[2], gI[2] * gPp[1] + gI[3] * gPp[3]};
        gP[0] = gT[0] * gI[0] + gT[1] * gI[1] + (gK[0] * gR) * gK[0];
        gP[1] = gT[0] * gI[2] + gT[1] * gI[3] + (gK[0] * gR) * gK[1];
        gP[2] = gT[2] * gI[0] + gT[3] * gI[1] + (gK[1] * gR) * gK[0];
        gP[3] = gT[2] * gI[2] + gT[3] * gI[3] + (gK[1] * gR) * gK[1];
        double gres = gm - gx[0];
        double gS2 = gP[0] + gR;
        chi2_zr += (gres * (1 / gS2)) * gres;
    }
    int dof = n - 2; /* :320 */
    *pval = gtfo_chi2_sf(chi2_xy, dof);
    *pval_zr = gtfo_chi2_sf(chi2_zr, dof);
}

int gtfo_extract(gtfo_arrays *A, const gtfo_geom *g, double pval_cut, int numhits, double sep3d,
                 double merge_dist, uint8_t *accepted, double *pval_xy, double *pval_zr)
{
    int n_acc = 0;
    gtfo_cca(A);
    int32_t *lab = A->label;
    for (int i = 0; i < A->N; i++) {
        if (accepted) accepted[i] = 0;
        if (pval_xy) pval_xy[i] = NAN;
        if (pval_zr) pval_zr[i] = NAN;
    }
    int *members = (int *)malloc(sizeof(int) * (A->N + 1));
    uint8_t *remove = (uint8_t *)calloc(A->N + 1, 1);
    for (int gph = 0; gph < A->S; gph++) {
        if (A->sub_state[gph] != GTF_SUB_INPLAY) continue;
        int b = A->sub_off[gph], e = A->sub_off[gph + 1];
        for (int root = b; root < e; root++) {
            if (!A->alive[root] || lab[root] != root) continue;
            int n = 0;
            for (int i = root; i < e; i++)
                if (A->alive[i] && lab[i] == root) members[n++] = i;
            if (n < numhits) continue; /* :415 */
            /* check_close_proximity_nodes (:58-151) */
            int n2 = 0, nother_bad = 0;
            for (int p = 0; p < n; p++) {
                int first = 1, cnt = 0;
                for (int q2 = 0; q2 < n; q2++)
                    if (A->volume[members[q2]] == A->volume[members[p]] && A->layer[members[q2]] == A->layer[members[p]]) {
                        if (q2 < p) first = 0;
                        cnt++;
                    }
                if (!first) continue;
                if (cnt == 2) n2++; else if (cnt != 1) nother_bad = 1;
            }
            double (*co)[4] = (double (*)[4])malloc(sizeof(double) * 4 * n);
            uint8_t *drop = (uint8_t *)calloc(n, 1);
            for (int p = 0; p < n; p++) {
                int m = members[p];
                co[p][0] = A->x[m]; co[p][1] = A->y[m]; co[p][2] = A->z[m]; co[p][3] = A->r[m];
            }
            int use_merged = 0;
            if (n2 > 0 && n2 <= 2 && !nother_bad) {
                use_merged = 1;
                for (int p = 0; p < n && use_merged; p++)
                    for (int q2 = p + 1; q2 < n; q2++) {
                        int mp = members[p], mq = members[q2];
                        if (A->volume[mp] != A->volume[mq] || A->layer[mp] != A->layer[mq]) continue;
                        double ddx = A->x[mp] - A->x[mq], ddy = A->y[mp] - A->y[mq], ddz = A->z[mp] - A->z[mq];
                        double dist = sqrt(ddx * ddx + ddy * ddy + ddz * ddz);
                        if (dist <= merge_dist) { /* :106-132 midpoint replaces node1, node2 dropped */
                            double xm = (A->x[mp] + A->x[mq]) / 2, ym = (A->y[mp] + A->y[mq]) / 2;
                            double zm = (A->z[mp] + A->z[mq]) / 2;
                            co[p][0] = xm; co[p][1] = ym; co[p][2] = zm; co[p][3] = sqrt(xm * xm + ym * ym);
                            drop[q2] = 1;
                        } else {
                            use_merged = 0; /* :136-139 copied_subgraph = None */
                        }
                        break;
                    }
            }
            int nk = 0; /* nodes of candidate_to_assess */
            int ok = 1;
            double (*cs)[4] = (double (*)[4])malloc(sizeof(double) * 4 * n);
            int *lay = (int *)malloc(sizeof(int) * 2 * n);
            for (int p = 0; p < n; p++) {
                if (use_merged && drop[p]) continue;
                int m = members[p];
                if (use_merged) memcpy(cs[nk], co[p], sizeof(double) * 4);
                else { cs[nk][0] = A->x[m]; cs[nk][1] = A->y[m]; cs[nk][2] = A->z[m]; cs[nk][3] = A->r[m]; }
                lay[2 * nk] = A->volume[m]; lay[2 * nk + 1] = A->layer[m];
                nk++;
            }
            for (int p = 0; p < nk && ok; p++) /* :427-429 one hit per layer */
                for (int q2 = 0; q2 < p; q2++)
                    if (lay[2 * p] == lay[2 * q2] && lay[2 * p + 1] == lay[2 * q2 + 1]) { ok = 0; break; }
            if (ok && nk >= numhits) {
                /* :434-436 stable sort by r, largest first */
                for (int p = 1; p < nk; p++) {
                    double t[4];
                    memcpy(t, cs[p], sizeof t);
                    int q2 = p;
                    while (q2 > 0 && cs[q2 - 1][3] < t[3]) { memcpy(cs[q2], cs[q2 - 1], sizeof t); q2--; }
                    memcpy(cs[q2], t, sizeof t);
                }
                /* rotate_track (:172-193) */
                const double *p1 = cs[nk - 1], *p2 = cs[nk - 2];
                double d3 = sqrt((p1[0] - p2[0]) * (p1[0] - p2[0]) + (p1[1] - p2[1]) * (p1[1] - p2[1]) +
                                 (p1[2] - p2[2]) * (p1[2] - p2[2]));
                if (d3 < sep3d) p2 = cs[nk - 3];
                double axy = atan2(p2[1] - p1[1], p2[0] - p1[0]);
                double azr = atan2(p2[2] - p1[2], p2[3] - p1[3]);
                for (int p = 0; p < nk; p++) {
                    double x = cs[p][0], y = cs[p][1], z = cs[p][2], r = cs[p][3];
                    cs[p][0] = x * cos(axy) + y * sin(axy);
                    cs[p][1] = -x * sin(axy) + y * cos(axy);
                    cs[p][3] = r * cos(azr) + r * sin(azr);
                    cs[p][2] = -z * sin(azr) + z * cos(azr);
                }
                double pv, pvz;
                kf_track_fit((const double (*)[4])cs, nk, g->sigma0xy, g->sigma0rz, g->endcap_boundary, &pv, &pvz);
                if (pval_xy) pval_xy[root] = pv;
                if (pval_zr) pval_zr[root] = pvz;
                if (pv >= pval_cut && pvz >= pval_cut) { /* :442 */
                    n_acc++;
                    for (int p = 0; p < n; p++) {
                        remove[members[p]] = 1;
                        if (accepted) accepted[members[p]] = 1;
                    }
                }
            }
            free(co); free(drop); free(cs); free(lay);
        }
        /* :460-467 */
        int left = 0;
        for (int i = b; i < e; i++) {
            if (remove[i]) A->alive[i] = 0;
            left += A->alive[i];
        }
        if (left == 0) A->sub_state[gph] = GTF_SUB_EMPTY;
        else if (left < numhits) A->sub_state[gph] = GTF_SUB_FRAGMENT;
    }
    free(members); free(remove);
    return n_acc;
}

/* ---------------------------------------------------------------- tag_propagation/tag_propagation.py:64-164 */

int gtfo_tag_propagation(gtfo_arrays *A, double threshold, int32_t *tags, int max_sweeps)
{
    /* work list: nodes with at least one successor of radius <= own (:99-110); isolated nodes have
     * no successors and never enter.  Jacobi sweeps of tag = max(own, kept successors) (:137-164). */
    int N = A->N;
    uint8_t *work = (uint8_t *)calloc(N + 1, 1);
    int nwork = 0;
    for (int i = 0; i < N; i++) {
        if (!A->alive[i] || A->sub_state[A->sub[i]] != GTF_SUB_INPLAY) continue;
        for (int o = A->out_off[i]; o < A->out_off[i + 1]; o++) {
            int v = A->slot_dst[A->out_slot[o]];
            if (A->alive[v] && !(A->r[v] > A->r[i])) { work[i] = 1; break; }
        }
        nwork += work[i];
    }
    int32_t *next = (int32_t *)malloc(sizeof(int32_t) * (N + 1));
    int sweeps = 0;
    double frac = 1.0;
    while (frac > threshold && nwork > 0 && sweeps < max_sweeps) {
        int flipped = 0;
        memcpy(next, tags, sizeof(int32_t) * N);
        for (int i = 0; i < N; i++) {
            if (!work[i]) continue;
            int32_t t = tags[i];
            for (int o = A->out_off[i]; o < A->out_off[i + 1]; o++) {
                int v = A->slot_dst[A->out_slot[o]];
                if (A->alive[v] && !(A->r[v] > A->r[i]) && tags[v] > t) t = tags[v];
            }
            next[i] = t;
            if (t != tags[i]) flipped++;
        }
        memcpy(tags, next, sizeof(int32_t) * N);
        frac = (double)flipped / nwork;
        sweeps++;
    }
    free(work); free(next);
    return sweeps;
}

/* learn_KL_parabolic_model/src/generate_training_data/utils.py:221-299 compute_track_state_estimates for ONE
 * (node, neighbour) pair, with :197-218 rotate_track (coords = [neighbours..., node, (0, 0)]: p1 = origin, p2 = node):
 * literal -- rotation of every point by 2 pi - atan2, translation, H built as written, np.linalg.inv -> inv_n,
 * H_inv . m and H_inv . S . H_inv^T as matrix products.  sv[3], cov[9] row-major. */
void gtfo_seed_parabolic(const double node_xy[2], const double nbr_xy[2], double sigma0, double sigmaA, double sigmaB,
                         double sv[3], double cov[9])
{
    const double pi = 3.14159265358979323846;
    const double p1[2] = {0.0, 0.0};
    double a = atan2(node_xy[1] - p1[1], node_xy[0] - p1[0]);    /* :191-194 */
    while (a < 0.0) a += pi * 2;                                  /* :184-187 */
    const double angle = 2 * pi - a;                              /* :209 */
    const double pts[3][2] = {{nbr_xy[0], nbr_xy[1]}, {node_xy[0], node_xy[1]}, {0.0, 0.0}};
    double rot[3][2];
    for (int k = 0; k < 3; k++) {                                 /* :213-217 */
        rot[k][0] = pts[k][0] * cos(angle) - pts[k][1] * sin(angle);
        rot[k][1] = pts[k][0] * sin(angle) + pts[k][1] * cos(angle);
    }
    const double x_trans = rot[1][0], y_trans = rot[1][1];        /* :261-262 rotated_coords[-2] = the node */
    double tr[3][2];
    for (int k = 0; k < 3; k++) { tr[k][0] = rot[k][0] - x_trans; tr[k][1] = rot[k][1] - y_trans; }
    const double x_0 = tr[2][0], x_B = tr[0][0], m_B = tr[0][1];  /* :273, :277-279 */
    const double meas[3] = {0.0, 0.0, m_B};
    const double H[9] = {x_0 * x_0, x_0, 1, 0, 0, 1, x_B * x_B, x_B, 1};   /* :281-283 */
    double Hi[9];
    inv_n(H, Hi, 3);                                              /* :286 */
    const double S[9] = {sigma0 * sigma0, 0, 0, 0, sigmaA * sigmaA, 0, 0, 0, sigmaB * sigmaB};
    for (int i = 0; i < 3; i++) {
        sv[i] = 0.0;
        for (int k = 0; k < 3; k++) sv[i] += Hi[3 * i + k] * meas[k];      /* :287 */
    }
    double HS[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            HS[3 * i + j] = 0.0;
            for (int k = 0; k < 3; k++) HS[3 * i + j] += Hi[3 * i + k] * S[3 * k + j];
        }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            cov[3 * i + j] = 0.0;
            for (int k = 0; k < 3; k++) cov[3 * i + j] += HS[3 * i + k] * Hi[3 * j + k];   /* :288 */
        }
}
