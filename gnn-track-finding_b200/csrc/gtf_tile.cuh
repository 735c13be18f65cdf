// gtf_tile.cuh -- the per-stage kernel on the SoA fields: one launch = one reference stage (or a short chain of them).
// The fused iteration runs on the packed layout instead (gtf_iter.cuh).
//
// One CTA owns a *tile*: a contiguous range of nodes (hits) and therefore a contiguous range of in-slots
// (incoming edges / mixture components).  Phases:
//   1. load     thread-per-slot, coalesced loads of the slot SoA into shared memory; per-slot gathers of the
//               source hit's coordinates / layer (L2-resident: sources live in the same event)
//   2. OP_E     thread-per-slot: extrapolate the source's merged state along the edge, chi2 gate, Kalman
//               update (extrapolate_merged_states.py:26-402) -- full-lane fp64 work, results stay in smem
//   3. node     warp-per-node over the node's slots in smem: dict-order bookkeeping, priors (helper.py:30-63),
//               side-norm + reweight + prune (helper.py:99-200), pairwise chi2 + greedy KL clustering
//               (clustering.py:193-307), degree / mixture weights (helper.py:67-94)
//   4. store    thread-per-slot coalesced write-back of what the program changed
// The program (list of OP_*) selects which of these run, so gtf_cluster / gtf_reweight / gtf_message_passing / ...
// are the same kernel with a one- to six-op program.
#pragma once
#include "gtf_dev.cuh"

#define F_EX 1u      // edge exists (both end nodes alive)
#define F_ACT 2u     // G[src][dst]['activated'] == 1
#define F_PRES 4u    // entry present in the working dict
#define F_NEW 8u     // inserted into the dict by this launch
#define F_FRESH 16u  // state (re)written by this launch
#define F_RW 32u     // weight rewritten by reweight in this launch
#define F_ORIG 64u   // activated flag as loaded
#define F_TMP 128u   // scratch (reweight eligibility)

#define SD_ORIGPRES 0x80u // (side array) entry was in the dict when the tile was loaded

#define NF_OK 1u     // node alive and its sub-graph still in play
#define NF_MULTI 2u  // sub-graph has != 1 nodes
#define NF_DICT 4u   // node has the working dict
#define NF_HASUTS 8u
#define NF_CLUSTERED 16u

struct TileSmem {
    double st[8][GTF_TILE_SLOTS]; // a b c tau p00 p01 p11 p22 of the working dict entry
    double prior[GTF_TILE_SLOTS], w[GTF_TILE_SLOTS], lik[GTF_TILE_SLOTS];
    double srcx[GTF_TILE_SLOTS];
    int32_t src[GTF_TILE_SLOTS], rank[GTF_TILE_SLOTS], layer[GTF_TILE_SLOTS];
    uint16_t ordl[GTF_TILE_SLOTS], dstl[GTF_TILE_SLOTS];
    uint8_t flags[GTF_TILE_SLOTS], side[GTF_TILE_SLOTS];
    uint16_t nbeg[GTF_TILE_NODES + 1];
    uint8_t nflags[GTF_TILE_NODES];
    double D[GTF_TILE_THREADS / 32][GTF_MAXD * (GTF_MAXD - 1) / 2 + 1];
    unsigned int cnt[GTF_NCOUNTERS];
    int next_node, elist_n;
};

template <class SM>
__device__ __forceinline__ GtfState tile_state(const SM &sm, int ls)
{
    GtfState s;
    s.a = sm.st[0][ls]; s.b = sm.st[1][ls]; s.c = sm.st[2][ls]; s.tau = sm.st[3][ls];
    s.p00 = sm.st[4][ls]; s.p01 = sm.st[5][ls]; s.p11 = sm.st[6][ls]; s.p22 = sm.st[7][ls];
    return s;
}
template <class SM>
__device__ __forceinline__ void tile_put_state(SM &sm, int ls, const GtfState &s)
{
    sm.st[0][ls] = s.a; sm.st[1][ls] = s.b; sm.st[2][ls] = s.c; sm.st[3][ls] = s.tau;
    sm.st[4][ls] = s.p00; sm.st[5][ls] = s.p01; sm.st[6][ls] = s.p11; sm.st[7][ls] = s.p22;
}
__device__ __forceinline__ double warp_min(double v)
{
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_min_i(int v)
{
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ unsigned warp_or(unsigned v)
{
    for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// pair p of the strictly lower triangle in row-major order: p = i (i - 1) / 2 + j, j < i <= 14
__constant__ uint8_t c_pair_i[GTF_MAXD * (GTF_MAXD - 1) / 2] = {
    1, 2, 2, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 5, 6, 6, 6, 6, 6, 6, 7, 7, 7, 7, 7, 7, 7, 8, 8, 8, 8, 8, 8, 8, 8,
    9, 9, 9, 9, 9, 9, 9, 9, 9, 10, 10, 10, 10, 10, 10, 10, 10, 10, 10, 11, 11, 11, 11, 11, 11, 11, 11, 11, 11, 11,
    12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 13, 13, 13, 13, 13, 13, 13, 13, 13, 13, 13, 13, 13,
    14, 14, 14, 14, 14, 14, 14, 14, 14, 14, 14, 14, 14, 14};
__device__ __forceinline__ void pair_decode(int p, int &i, int &j)
{
    i = c_pair_i[p];
    j = p - i * (i - 1) / 2;
}
// 1/k for small positive integer k, exactly the correctly rounded quotient (helper.py:61,90)
__device__ __forceinline__ double recip_small(int k)
{
    return k == 1 ? 1.0 : k == 2 ? 0.5 : k == 4 ? 0.25 : 1.0 / (double)k;
}

// dict order: ordl[b0 + k] = local slot of the k-th entry (ascending rank); returns the entry count
__device__ __forceinline__ int node_build_order(TileSmem &sm, int b0, int b1, int lane)
{
    int n = 0;
    for (int base = b0; base < b1; base += 32) {
        int ls = base + lane;
        bool pres = ls < b1 && (sm.flags[ls] & F_PRES);
        if (pres) {
            int r = sm.rank[ls], pos = 0;
            for (int t = b0; t < b1; t++)
                if ((sm.flags[t] & F_PRES) && sm.rank[t] < r) pos++;
            sm.ordl[b0 + pos] = (uint16_t)ls;
        }
        n += __popc(__ballot_sync(0xffffffffu, pres));
    }
    __syncwarp();
    return n;
}

// helper.py:30-63 compute_prior_probabilities for one node
__device__ __forceinline__ void node_prior(TileSmem &sm, int b0, int b1, int lane)
{
    const unsigned m = F_PRES | F_EX | F_ACT;
    for (int base = b0; base < b1; base += 32) {
        int ls = base + lane;
        if (ls < b1 && (sm.flags[ls] & m) == m) {
            int lay = sm.layer[ls], cnt = 0;
            for (int t = b0; t < b1; t++)
                if ((sm.flags[t] & m) == m && sm.layer[t] == lay) cnt++;
            sm.prior[ls] = 1.0 / cnt;
        }
    }
    __syncwarp();
}

// helper.py:99-200 calculate_side_norm_factor + reweight for one node
__device__ __forceinline__ void node_reweight(TileSmem &sm, int b0, int b1, int n, double nodex, double thr, int lane,
                                              double *edge_w_tile, double *lrn_tile, const DevBatch &B, int s0)
{
    const unsigned m = F_PRES | F_EX | F_ACT;
    int nl = 0, nr = 0, normL = 0, normR = 0;
    for (int base = b0; base < b1; base += 32) {
        int ls = base + lane;
        bool el = ls < b1 && (sm.flags[ls] & m) == m;
        bool left = el && sm.srcx[ls] < nodex, right = el && !left;
        bool first = el;
        if (el) {
            double xs = sm.srcx[ls];
            for (int t = b0; t < ls; t++)
                if ((sm.flags[t] & m) == m && (sm.srcx[t] < nodex) == left && sm.srcx[t] == xs) { first = false; break; }
            sm.flags[ls] |= F_TMP;
            sm.side[ls] = (uint8_t)((sm.side[ls] & SD_ORIGPRES) | (left ? 1 : 2));
        } else if (ls < b1)
            sm.flags[ls] &= ~F_TMP;
        nl += __popc(__ballot_sync(0xffffffffu, left));
        nr += __popc(__ballot_sync(0xffffffffu, right));
        normL += __popc(__ballot_sync(0xffffffffu, left && first));
        normR += __popc(__ballot_sync(0xffffffffu, right && first));
    }
    __syncwarp();
    if (nl + nr == 0) return;
    // stale `neighbour_num`: the LAST key of the dict decides whether the norms apply (helper.py:131,138)
    unsigned lf = sm.flags[sm.ordl[b0 + n - 1]];
    if (!(lf & F_EX) && lane == 0) atomicOr(&sm.cnt[CNT_REFERR], (unsigned)GTF_REF_KEY);
    bool last_active = (lf & (F_EX | F_ACT)) == (F_EX | F_ACT);
    double denom = 0.0; // helper.py:165-169, dict order
    for (int k = 0; k < n; k++) {
        int t = sm.ordl[b0 + k];
        if (sm.flags[t] & F_TMP) denom += sm.w[t] * sm.lik[t];
    }
    int off = 0;
    for (int base = b0; base < b1; base += 32) {
        int ls = base + lane;
        if (ls < b1 && (sm.flags[ls] & F_TMP)) {
            double norm = last_active ? (double)((sm.side[ls] & 3) == 1 ? normL : normR) : 1.0;
            double rw = (sm.w[ls] * sm.lik[ls] * sm.prior[ls]) / denom;
            rw = rw / norm;
            lrn_tile[ls] = norm;
            sm.w[ls] = rw;
            edge_w_tile[ls] = rw; // helper.py:180 edge attribute (coalesced: consecutive lanes, consecutive slots)
            unsigned f = sm.flags[ls] | F_RW;
            near_note(B, GTF_NEAR_REWEIGHT, s0 + ls, rw, thr);
            if (rw < thr) { f &= ~F_ACT; off++; } else f |= F_ACT;
            sm.flags[ls] = (uint8_t)f;
        }
    }
    if (off) atomicAdd(&sm.cnt[CNT_RWOFF], (unsigned)off);
    __syncwarp();
}

// order-preserving map double -> uint64 (no NaNs): lets REDUX (32-bit integer warp reductions) do fp64 arg-mins
__device__ __forceinline__ unsigned long long dbl_key(double v)
{
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_dbl(unsigned long long k)
{
    unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
// warp minimum of v over lanes with `have` (three REDUX instead of ten 64-bit shuffle steps); first lane on ties
__device__ __forceinline__ int warp_argmin(double v, bool have, double &vmin)
{
    const unsigned FULL = 0xffffffffu;
    unsigned long long k = have ? dbl_key(v) : ~0ull;
    unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    unsigned mh = __reduce_min_sync(FULL, hi);
    bool c1 = have && hi == mh;
    unsigned ml = __reduce_min_sync(FULL, c1 ? lo : 0xffffffffu);
    unsigned m = __ballot_sync(FULL, c1 && lo == ml);
    vmin = key_dbl(((unsigned long long)mh << 32) | ml);
    return m ? __ffs(m) - 1 : -1;
}
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ void shfl_info(const GtfInfo &in, int src, GtfInfo &out)
{
    out.s00 = shfl_d(in.s00, src); out.s01 = shfl_d(in.s01, src); out.s11 = shfl_d(in.s11, src); out.sq = shfl_d(in.sq, src);
    out.v0 = shfl_d(in.v0, src); out.v1 = shfl_d(in.v1, src); out.vc = shfl_d(in.vc, src); out.vt = shfl_d(in.vt, src);
}

// clustering.py:193-307 for one node.  Returns true and the merged state when a cluster was formed.
// Lane k < n owns dict entry k.  Pairwise chi2: lane-per-pair into Dw.  Greedy loop in information form
// (gtf_math.cuh): per round one KL per lane without any inverse, a REDUX arg-min, and one 2x2 inverse.
template <class SM>
__device__ __forceinline__ bool node_cluster(SM &sm, double *Dw, int b0, int n, double nx, double nz, double nr_,
                                             double chi2_thr, double kl_thr, const GtfGeom &g, int lane,
                                             GtfState &merged, double &mprior, const double *gz, const double *gr,
                                             const DevBatch &B, int node)
{
    if (n < 3 || n > GTF_MAXD) return false; // clustering.py:207
    const unsigned FULL = 0xffffffffu;
    int npairs = n * (n - 1) / 2;
    double lbest = INFINITY;
    bool nz_any = false, nan_any = false;
    for (int p = lane; p < npairs; p += 32) {
        int i, j;
        pair_decode(p, i, j);
        int ei = sm.ordl[b0 + i], ej = sm.ordl[b0 + j];
        int ui = sm.src[ei], uj = sm.src[ej];
        double v = gtf_pair_chi2(tile_state(sm, ei), tile_state(sm, ej), nx, nz, nr_, sm.srcx[ei], gz[ui], gr[ui],
                                 sm.srcx[ej], gz[uj], gr[uj], g);
        Dw[p] = v;
        if (v != 0.0) {          // np.nonzero keeps NaN, drops +-0 (clustering.py:119)
            nz_any = true;
            if (v != v) nan_any = true; else lbest = fmin(lbest, v);
        }
    }
    __syncwarp();
    nz_any = __any_sync(FULL, nz_any);
    nan_any = __any_sync(FULL, nan_any);
    if (!nz_any) { // np.min([]) -> ValueError
        if (lane == 0) atomicOr(&sm.cnt[CNT_REFERR], (unsigned)GTF_REF_EMPTY_MIN);
        return false;
    }
    if (nan_any) return false; // np.min -> nan, `nan < thr` False
    double best;
    warp_argmin(lbest, true, best);
    if (lane == 0) near_note(B, GTF_NEAR_CLUSTER_CHI2, node, best, chi2_thr);
    if (!(best < chi2_thr)) return false; // clustering.py:228
    // np.where(distances == smallest): all tied positions in row-major order (clustering.py:122-123)
    int p1 = 1 << 30, nm = 0;
    unsigned gone = 0;
    for (int p = lane; p < npairs; p += 32)
        if (Dw[p] == best) {
            int i, j;
            pair_decode(p, i, j);
            p1 = min(p1, p);
            nm++;
            gone |= (1u << i) | (1u << j);
        }
    int pfirst = (int)__reduce_min_sync(FULL, (unsigned)p1);
    nm = (int)__reduce_add_sync(FULL, (unsigned)nm);
    gone = __reduce_or_sync(FULL, gone);
    int idx0, idx1;
    pair_decode(pfirst, idx0, idx1); // unique minimum: idx = [row, col]
    if (nm > 1) {                    // ties: idx = [rows..., cols...] -> idx[1] is the SECOND ROW
        int p2 = 1 << 30;
        for (int p = lane; p < npairs; p += 32)
            if (Dw[p] == best && p > pfirst) p2 = min(p2, p);
        p2 = (int)__reduce_min_sync(FULL, (unsigned)p2);
        int jj;
        pair_decode(p2, idx1, jj);
    }
    unsigned rem = ((1u << n) - 1u) & ~gone;
    if (rem == 0) { // np.min([]) at :252
        if (lane == 0) atomicOr(&sm.cnt[CNT_REFERR], (unsigned)GTF_REF_EMPTY_MIN);
        return false;
    }
    GtfState mine;
    GtfInfo mine_i, M, t;
    double myprior = 0.0;
    {
        int e = sm.ordl[b0 + min(lane, n - 1)];
        mine = tile_state(sm, e);
        myprior = sm.prior[e];
        gtf_to_info(mine, mine_i);
    }
    shfl_info(mine_i, idx0, M);                                // clustering.py:231-233: Sigma^-1 = S_i + S_j
    shfl_info(mine_i, idx1, t);
    gtf_info_add(M, t);
    gtf_from_info(M, merged);
    mprior = shfl_d(myprior, idx0) + shfl_d(myprior, idx1);    // :234
    for (;;) {
        bool have = lane < n && ((rem >> lane) & 1u);
        double kl = have ? gtf_kl_info(mine, mine_i, merged, M) : INFINITY; // clustering.py:107-112
        if (__any_sync(FULL, have && kl != kl)) {                           // list.index(nan) -> ValueError
            if (lane == 0) atomicOr(&sm.cnt[CNT_REFERR], (unsigned)GTF_REF_NAN_INDEX);
            return false;
        }
        double bv;
        int bk = warp_argmin(kl, have, bv);                  // list.index: first occurrence
        if (bk >= 0 && lane == 0) near_note(B, GTF_NEAR_CLUSTER_KL, node, bv, kl_thr);
        if (bk < 0 || !(bv < kl_thr)) break;                 // clustering.py:261
        shfl_info(mine_i, bk, t);                            // :263-265 merge_states(entry, merged)
        gtf_info_add(M, t);
        gtf_from_info(M, merged);
        mprior = shfl_d(myprior, bk) + mprior;               // :266
        rem &= ~(1u << bk);
        if (rem == 0) break;                                 // :283
    }
    // un-absorbed components: their in-edge is deactivated (clustering.py:297-321)
    if (lane < n && ((rem >> lane) & 1u)) {
        int e = sm.ordl[b0 + lane];
        if (sm.flags[e] & F_EX) {
            sm.flags[e] &= ~F_ACT;
            atomicAdd(&sm.cnt[CNT_DEACT], 1u);
        }
    }
    __syncwarp();
    return true;
}

// ------------------------------------------------------------------------------------------------
// generic per-node program (any in-degree; state in shared memory).  Only used for nodes with more than 32
// in-slots -- everything else runs the register-resident fast path below.
__device__ __noinline__ void node_program_generic(TileSmem &sm, const DevBatch B, const Prog P, const GtfGeom g, int i,
                                                  int ln, int s0, int warp, int lane, bool uts, uint8_t *hm_out,
                                                  double *const *mo, double *lrn_base, double *ew_base)
{
    const int b0 = sm.nbeg[ln], b1 = sm.nbeg[ln + 1];
    unsigned nf = sm.nflags[ln];
    int n = -1;
    bool clustered = false;
    GtfState merged;
    double mprior = 0.0;
    for (int k = 0; k < 12 && P.ops[k] != OP_END; k++) {
        const int op = P.ops[k];
        if (op == OP_E) {
            int nnew = 0;
            for (int base = b0; base < b1; base += 32) {
                int ls = base + lane;
                nnew += __popc(__ballot_sync(0xffffffffu, ls < b1 && (sm.flags[ls] & F_NEW)));
            }
            if (nnew) {
                int nxt = B.uts_next[i];
                for (int base = b0; base < b1; base += 32) {
                    int ls = base + lane;
                    if (ls < b1 && (sm.flags[ls] & F_NEW)) {
                        int before = 0, me = sm.src[ls];
                        for (int t = b0; t < b1; t++)
                            if ((sm.flags[t] & F_NEW) && sm.src[t] < me) before++;
                        sm.rank[ls] = nxt + before;
                    }
                }
                __syncwarp();
                if (lane == 0) { B.uts_next[i] = nxt + nnew; B.has_uts[i] = 1; }
                nf |= NF_DICT | NF_HASUTS;
                n = -1;
                __syncwarp();
            }
        } else if (op == OP_POP) {
            bool mine = (nf & NF_OK) && (uts ? (nf & NF_HASUTS) != 0 : (nf & NF_HASUTS) == 0);
            if (mine) {
                for (int base = b0; base < b1; base += 32) {
                    int ls = base + lane;
                    if (ls < b1 && (sm.flags[ls] & F_PRES)) {
                        bool succ = (sm.flags[ls] & F_EX) && B.rev_slot[s0 + ls] >= 0;
                        if (!succ) sm.flags[ls] &= ~F_PRES;
                    }
                }
                n = -1;
                __syncwarp();
            }
        } else if (op == OP_PRIOR) {
            if ((nf & (NF_OK | NF_MULTI | NF_DICT)) == (NF_OK | NF_MULTI | NF_DICT)) node_prior(sm, b0, b1, lane);
        } else if (op == OP_RW) {
            if (uts && (nf & (NF_OK | NF_MULTI | NF_HASUTS)) == (NF_OK | NF_MULTI | NF_HASUTS)) {
                if (n < 0) n = node_build_order(sm, b0, b1, lane);
                node_reweight(sm, b0, b1, n, B.x[i], P.rw_thr, lane, ew_base, lrn_base, B, s0);
            }
        } else if (op == OP_CLUSTER) {
            if ((nf & (NF_OK | NF_DICT)) == (NF_OK | NF_DICT)) {
                if (n < 0) n = node_build_order(sm, b0, b1, lane);
                double thr = P.cl_kl;
                if (P.use_lut) {
                    double ev = B.emp_var[i];
                    int bin = (ev == ev) ? (int)floor(ev / 0.05) : 27;
                    thr = P.lut[max(0, min(27, bin))];
                }
                clustered = node_cluster(sm, sm.D[warp], b0, n, B.x[i], B.z[i], B.r[i], P.cl_chi2, thr, g, lane, merged, mprior, B.z, B.r, B, i);
            }
        } else if (op == OP_DEGREE) {
            int deg = 0;
            for (int base = b0; base < b1; base += 32) {
                int ls = base + lane;
                deg += __popc(__ballot_sync(0xffffffffu, ls < b1 && (sm.flags[ls] & (F_EX | F_ACT)) == (F_EX | F_ACT)));
            }
            if ((nf & NF_OK) && lane == 0) B.degree[i] = deg;
        } else if (op == OP_WEIGHTS) {
            if ((nf & (NF_OK | NF_MULTI | NF_DICT)) == (NF_OK | NF_MULTI | NF_DICT)) {
                if (n < 0) n = node_build_order(sm, b0, b1, lane);
                if (n == 0) {
                    if (lane == 0) atomicOr(&sm.cnt[CNT_REFERR], (unsigned)GTF_REF_ZERO_DIV);
                } else {
                    double mw = 1.0 / n;
                    for (int base = b0; base < b1; base += 32) {
                        int ls = base + lane;
                        if (ls < b1 && (sm.flags[ls] & F_PRES)) sm.w[ls] = mw;
                    }
                    __syncwarp();
                }
            }
        }
    }
    if (clustered && lane == 0) {
        hm_out[i] = 1;
        mo[0][i] = merged.a; mo[1][i] = merged.b; mo[2][i] = merged.c; mo[3][i] = merged.p00;
        mo[4][i] = merged.p01; mo[5][i] = merged.p11; mo[6][i] = merged.p22; mo[7][i] = mprior;
        sm.nflags[ln] = (uint8_t)(nf | NF_CLUSTERED);
        atomicAdd(&sm.cnt[CNT_MERGED], 1u);
    }
}

// register-resident per-node program for nodes with <= 32 in-slots: lane l owns slot b0 + l; counts,
// same-layer / same-x groupings and dict positions come from ballots, MATCH.ANY and shuffles instead of
// O(d^2) shared-memory loops.
template <class SM>
__device__ __forceinline__ void node_program_fast(SM &sm, const DevBatch &B, const Prog &P, const GtfGeom &g, int i,
                                                  int ln, int s0, int warp, int lane, bool uts, uint8_t *hm_out,
                                                  double *const *mo)
{
    const unsigned FULL = 0xffffffffu;
    const int b0 = sm.nbeg[ln], d = sm.nbeg[ln + 1] - b0;
    const int ls = b0 + lane;
    const bool valid = lane < d;
    unsigned nf = sm.nflags[ln];
    unsigned f = valid ? sm.flags[ls] : 0u;
    int lay = valid ? sm.layer[ls] : (-100 - lane);
    int src = valid ? sm.src[ls] : 0, rank = valid ? sm.rank[ls] : 0x7fffffff;
    double sx = valid ? sm.srcx[ls] + 0.0 : 0.0;
    double w = 0.0, lik = 0.0, prior = 0.0;
    int sidev = 0;
    if (f & F_PRES) { w = sm.w[ls]; lik = sm.lik[ls]; prior = sm.prior[ls]; }
    const double nodex = B.x[i];
    const unsigned lt = (1u << lane) - 1u;
    int n = -1, pos = 0;
    bool clustered = false;
    GtfState merged;
    double mprior = 0.0;
    double *scratch = sm.D[warp];

#pragma unroll
    for (int k = 0; k < 12; k++) {
        const int op = P.ops[k];
        if (op == OP_END) break;
        bool need_order = (op == OP_RW || op == OP_CLUSTER || op == OP_WEIGHTS);
        if (op == OP_E) {
            // dict insertion order of new entries = ascending source node index (extrapolate...py:419-447)
            unsigned newmask = __ballot_sync(FULL, (f & F_NEW) != 0);
            if (newmask) {
                int nxt = B.uts_next[i], before = 0;
                for (unsigned m = newmask; m; m &= m - 1) {
                    int sk = __shfl_sync(FULL, src, __ffs(m) - 1);
                    before += sk < src;
                }
                if (f & F_NEW) rank = nxt + before;
                __syncwarp();
                if (lane == 0) { B.uts_next[i] = nxt + __popc(newmask); B.has_uts[i] = 1; }
                nf |= NF_DICT | NF_HASUTS;
                n = -1;
            }
        } else if (op == OP_POP) {
            bool mine = (nf & NF_OK) && (uts ? (nf & NF_HASUTS) != 0 : (nf & NF_HASUTS) == 0);
            if (mine) {
                if (f & F_PRES) {
                    bool succ = (f & F_EX) && B.rev_slot[s0 + ls] >= 0;
                    if (!succ) f &= ~F_PRES;
                }
                n = -1;
            }
        }
        if (need_order && n < 0) {
            unsigned pm = __ballot_sync(FULL, (f & F_PRES) != 0);
            n = __popc(pm);
            pos = 0;
            if (!uts) pos = __popc(pm & lt); // seed dict: slot order
            else
                for (unsigned m = pm; m; m &= m - 1) {
                    int rk = __shfl_sync(FULL, rank, __ffs(m) - 1);
                    pos += rk < rank;
                }
        }
        if (op == OP_PRIOR) {
            if ((nf & (NF_OK | NF_MULTI | NF_DICT)) == (NF_OK | NF_MULTI | NF_DICT)) {
                const unsigned m3 = F_PRES | F_EX | F_ACT;
                bool el = (f & m3) == m3;
                unsigned elmask = __ballot_sync(FULL, el);
                unsigned same = __match_any_sync(FULL, lay);
                if (el) prior = recip_small(__popc(same & elmask)); // helper.py:61
            }
        } else if (op == OP_RW) {
            if (uts && (nf & (NF_OK | NF_MULTI | NF_HASUTS)) == (NF_OK | NF_MULTI | NF_HASUTS)) {
                const unsigned m3 = F_PRES | F_EX | F_ACT;
                bool el = (f & m3) == m3;
                bool left = el && sx < nodex;
                unsigned elmask = __ballot_sync(FULL, el), leftmask = __ballot_sync(FULL, left);
                if (elmask) {
                    unsigned samex = __match_any_sync(FULL, __double_as_longlong(sx));
                    unsigned grp = samex & (left ? leftmask : (elmask & ~leftmask));
                    bool first = el && (__ffs(grp) - 1 == lane);      // distinct x per side: len(set(coords))
                    int normL = __popc(__ballot_sync(FULL, first && left));
                    int normR = __popc(__ballot_sync(FULL, first && !left));
                    // stale `neighbour_num`: the LAST dict key gates the norms (helper.py:131,138)
                    unsigned lastm = __ballot_sync(FULL, (f & F_PRES) && pos == n - 1);
                    unsigned lf = __shfl_sync(FULL, f, __ffs(lastm) - 1);
                    if (!(lf & F_EX) && lane == 0) atomicOr(&sm.cnt[CNT_REFERR], (unsigned)GTF_REF_KEY);
                    bool last_active = (lf & (F_EX | F_ACT)) == (F_EX | F_ACT);
                    // denominator in dict order (helper.py:165-169); adding 0.0 for the others is exact
                    if (f & F_PRES) scratch[pos] = el ? w * lik : 0.0;
                    __syncwarp();
                    double denom = 0.0;
                    for (int q = 0; q < n; q++) denom += scratch[q];
                    __syncwarp();
                    if (el) {
                        double norm = last_active ? (double)(left ? normL : normR) : 1.0;
                        double rw = (w * lik * prior) / denom;
                        if (norm != 1.0) rw = rw / norm;
                        B.uts_lrn[s0 + ls] = norm;
                        sidev = left ? 1 : 2;
                        w = rw;
                        B.edge_w[s0 + ls] = rw; // helper.py:180
                        f |= F_RW;
                        near_note(B, GTF_NEAR_REWEIGHT, s0 + ls, rw, P.rw_thr);
                        if (rw < P.rw_thr) f &= ~F_ACT; else f |= F_ACT;
                    }
                    int off = __popc(__ballot_sync(FULL, el && !(f & F_ACT)));
                    if (off && lane == 0) atomicAdd(&sm.cnt[CNT_RWOFF], (unsigned)off);
                }
            }
        } else if (op == OP_CLUSTER) {
            if ((nf & (NF_OK | NF_DICT)) == (NF_OK | NF_DICT) && n >= 3 && n <= GTF_MAXD) {
                if (valid) { sm.flags[ls] = (uint8_t)f; sm.prior[ls] = prior; }
                if (f & F_PRES) sm.ordl[b0 + pos] = (uint16_t)ls;
                __syncwarp();
                double thr = P.cl_kl;
                if (P.use_lut) {
                    double ev = B.emp_var[i];
                    int bin = (ev == ev) ? (int)floor(ev / 0.05) : 27;
                    thr = P.lut[max(0, min(27, bin))];
                }
                clustered = node_cluster(sm, scratch, b0, n, nodex, B.z[i], B.r[i], P.cl_chi2, thr, g, lane, merged, mprior, B.z, B.r, B, i);
                if (valid) f = sm.flags[ls];
            }
        } else if (op == OP_DEGREE) {
            int deg = __popc(__ballot_sync(FULL, (f & (F_EX | F_ACT)) == (F_EX | F_ACT)));
            if ((nf & NF_OK) && lane == 0) B.degree[i] = deg;
        } else if (op == OP_WEIGHTS) {
            if ((nf & (NF_OK | NF_MULTI | NF_DICT)) == (NF_OK | NF_MULTI | NF_DICT)) {
                if (n == 0) {
                    if (lane == 0) atomicOr(&sm.cnt[CNT_REFERR], (unsigned)GTF_REF_ZERO_DIV);
                } else if (f & F_PRES)
                    w = recip_small(n);
            }
        }
    }
    if (valid) {
        sm.flags[ls] = (uint8_t)f;
        sm.rank[ls] = rank;
        if (f & F_PRES) { sm.w[ls] = w; sm.prior[ls] = prior; }
        if (f & F_RW) sm.side[ls] = (uint8_t)((sm.side[ls] & SD_ORIGPRES) | sidev);
    }
    if (clustered && lane == 0) {
        hm_out[i] = 1;
        mo[0][i] = merged.a; mo[1][i] = merged.b; mo[2][i] = merged.c; mo[3][i] = merged.p00;
        mo[4][i] = merged.p01; mo[5][i] = merged.p11; mo[6][i] = merged.p22; mo[7][i] = mprior;
        sm.nflags[ln] = (uint8_t)(nf | NF_CLUSTERED);
        atomicAdd(&sm.cnt[CNT_MERGED], 1u);
    }
}


__global__ void __launch_bounds__(GTF_TILE_THREADS, GTF_TILE_MINB) k_tile(DevBatch B, Prog P, GtfGeom g)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileSmem &sm = *reinterpret_cast<TileSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n0 = B.tile_begin[blockIdx.x], n1 = B.tile_begin[blockIdx.x + 1];
    const int nn = n1 - n0;
    const int s0 = B.in_off[n0], ns = B.in_off[n1] - s0;
    const bool uts = P.key == GTF_KEY_UTS;
    const int wb = P.wb;
    bool has_E = false;
    for (int k = 0; k < 12 && P.ops[k] != OP_END; k++) has_E |= P.ops[k] == OP_E;

    if (tid < GTF_NCOUNTERS) sm.cnt[tid] = 0;
    if (tid == 0) { sm.next_node = GTF_TILE_THREADS / 32; sm.elist_n = 0; }
    __syncthreads();
    // ---------------------------------------------------------------- node table
    for (int ln = tid; ln <= nn; ln += GTF_TILE_THREADS) {
        sm.nbeg[ln] = (uint16_t)(B.in_off[n0 + ln] - s0);
        if (ln < nn) {
            int i = n0 + ln;
            unsigned f = B.node_ok[i]; // NF_OK | NF_MULTI, derived per node
            bool hu = B.has_uts[i] != 0;
            if (hu) f |= NF_HASUTS;
            if (!uts || hu) f |= NF_DICT;
            sm.nflags[ln] = (uint8_t)f;
        }
    }
    // ---------------------------------------------------------------- load (thread per slot)
    const uint8_t *present_in = uts ? B.uts_present : B.tse_present;
    // stage 1: the coalesced per-slot arrays of all of this thread's slots (independent loads in flight together)
    constexpr int NIT = (GTF_TILE_SLOTS + GTF_TILE_THREADS - 1) / GTF_TILE_THREADS;
    int r_src[NIT], r_dst[NIT];
    unsigned r_f[NIT];
#pragma unroll
    for (int it = 0; it < NIT; it++) {
        const int ls = it * GTF_TILE_THREADS + tid;
        r_src[it] = -1; r_dst[it] = n0; r_f[it] = 0;
        if (ls < ns) {
            const int s = s0 + ls;
            r_src[it] = B.in_src[s];
            r_dst[it] = B.slot_dst[s];
            unsigned f = 0;
            if (B.active[s] == 1) f |= F_ACT | F_ORIG;
            if (present_in[s]) f |= F_PRES;
            r_f[it] = f;
        }
    }
    // stage 2: gathers through the source index, shared-memory staging, message list
#pragma unroll
    for (int it = 0; it < NIT; it++) {
        const int base = it * GTF_TILE_THREADS;
        if (base >= ns) break;
        const int ls = base + tid;
        bool send = false;
        if (ls < ns) {
        int s = s0 + ls;
        int src = r_src[it], dst = r_dst[it];
        sm.src[ls] = src;
        sm.dstl[ls] = (uint16_t)(dst - n0);
        unsigned f = r_f[it];
        if (src >= 0 && (B.all_alive || (B.alive[src] && B.alive[dst]))) f |= F_EX;
        // E pass 1 folded in: does this slot carry a message this iteration? (extrapolate...py:416,425,431)
        if (has_E && (f & (F_EX | F_ACT)) == (F_EX | F_ACT))
            send = (B.node_ok[dst] & (NF_OK | NF_MULTI)) == (NF_OK | NF_MULTI) && B.has_merged[src] != 0;
        // source x / layer are only ever used for dict entries (priors, side norm, pairwise chi2) and messages
        double sx = 0;
        int lay = -1;
        if (src >= 0 && ((f & F_PRES) || send)) { sx = B.x[src]; lay = B.layer[src]; }
        sm.srcx[ls] = sx; sm.layer[ls] = lay;
        unsigned sd = (f & F_PRES) ? SD_ORIGPRES : 0u;
        if (f & F_PRES) {
            if (uts) {
                sm.st[0][ls] = B.uts_a[s]; sm.st[1][ls] = B.uts_b[s]; sm.st[2][ls] = B.uts_c[s]; sm.st[3][ls] = B.uts_tau[s];
                sm.st[4][ls] = B.uts_p00[s]; sm.st[5][ls] = B.uts_p01[s]; sm.st[6][ls] = B.uts_p11[s]; sm.st[7][ls] = B.uts_p22[s];
                sm.prior[ls] = B.uts_prior[s]; sm.w[ls] = B.uts_w[s]; sm.lik[ls] = B.uts_lik[s];
                sd |= (unsigned)B.uts_side[s] & 3u;
                sm.rank[ls] = B.uts_rank[s];
            } else {
                sm.st[0][ls] = B.tse_a[s]; sm.st[1][ls] = B.tse_b[s]; sm.st[2][ls] = B.tse_c[s]; sm.st[3][ls] = B.tse_tau[s];
                sm.st[4][ls] = B.tse_p00[s]; sm.st[5][ls] = B.tse_p01[s]; sm.st[6][ls] = B.tse_p11[s]; sm.st[7][ls] = B.tse_p22[s];
                sm.prior[ls] = B.tse_prior[s]; sm.w[ls] = B.tse_w[s];
                sm.rank[ls] = ls;
            }
        } else
            sm.rank[ls] = uts ? 0x7fffffff : ls;
        sm.flags[ls] = (uint8_t)f;
        sm.side[ls] = (uint8_t)sd;
        }
        if (has_E) { // dense list of message slots, so the extrapolation runs with full warps
            unsigned m = __ballot_sync(0xffffffffu, send);
            if (m) {
                int basepos = 0;
                if (lane == 0) basepos = atomicAdd(&sm.elist_n, __popc(m));
                basepos = __shfl_sync(0xffffffffu, basepos, 0);
                if (send) sm.ordl[basepos + __popc(m & ((1u << lane) - 1u))] = (uint16_t)ls;
            }
        }
    }
    __syncthreads();

    // ---------------------------------------------------------------- OP_E
    if (has_E) {
        const int nsend = sm.elist_n;
        unsigned gated = 0;
        // pass 2: thread per message: extrapolate, chi2 gate, Kalman update (extrapolate...py:26-402)
        for (int q = tid; q < nsend; q += GTF_TILE_THREADS) {
            int ls = sm.ordl[q];
            unsigned f = sm.flags[ls];
            int u = sm.src[ls];
            int s = s0 + ls, v = n0 + sm.dstl[ls];
            GtfExtrapOut o;
            gtf_extrapolate(sm.srcx[ls], B.y[u], B.z[u], B.r[u], B.x[v], B.y[v], B.z[v], B.r[v], B.m_a[u],
                            B.m_b[u], B.m_c[u], B.m_p00[u], B.m_p01[u], B.slot_p11[s], B.m_p22[u], B.slot_vms[s],
                            P.chi2_cut, g, o);
            near_note(B, GTF_NEAR_GATE, s, o.chi2, P.chi2_cut);
            B.uts_chi2[s] = o.chi2;
            if (o.pass) {
                int rs = B.rev_slot[s];
                double wv = NAN;
                if (rs >= 0 && B.tse_present[rs]) wv = B.tse_w[rs]; // :384
                else atomicOr(&sm.cnt[CNT_REFERR], (unsigned)GTF_REF_NO_TSE);
                tile_put_state(sm, ls, o.s);
                sm.lik[ls] = o.lik;
                sm.w[ls] = wv;
                sm.prior[ls] = NAN; // a fresh dict entry has no prior / lr_layer_norm / side yet
                sm.side[ls] &= SD_ORIGPRES;
                f |= F_FRESH;
                if (!(f & F_PRES)) f |= F_PRES | F_NEW;
            } else {
                f &= ~F_ACT; // :393
                gated++;
            }
            sm.flags[ls] = (uint8_t)f;
        }
        if (tid == 0 && nsend) atomicAdd(&sm.cnt[CNT_SENT], (unsigned)nsend);
        if (gated) atomicAdd(&sm.cnt[CNT_GATED], gated);
        __syncthreads();
    }

    // ---------------------------------------------------------------- node programs (warp per node, dynamic)
    uint8_t *hm_out = B.has_merged;
    double *const mo[8] = {B.m_a, B.m_b, B.m_c, B.m_p00, B.m_p01, B.m_p11, B.m_p22, B.m_prior};
    for (int ln = warp; ln < nn;) {
        const int i = n0 + ln;
        if (sm.nbeg[ln + 1] - sm.nbeg[ln] <= 32) node_program_fast<TileSmem>(sm, B, P, g, i, ln, s0, warp, lane, uts, hm_out, mo);
        else node_program_generic(sm, B, P, g, i, ln, s0, warp, lane, uts, hm_out, mo, B.uts_lrn + s0, B.edge_w + s0);
        int nxt = 0;
        if (lane == 0) nxt = atomicAdd(&sm.next_node, 1);
        ln = __shfl_sync(0xffffffffu, nxt, 0);
    }
    __syncthreads();

    // ---------------------------------------------------------------- store
    uint8_t *act_out = B.active;
    unsigned n_act = 0, n_chg = 0;
    for (int ls = tid; ls < ns; ls += GTF_TILE_THREADS) {
        int s = s0 + ls;
        unsigned f = sm.flags[ls];
        bool a = f & F_ACT, a0 = f & F_ORIG;
        if (f & F_EX) {
            n_act += a;
            n_chg += a != a0;
        }
        if ((wb & WB_ACTIVE) && a != a0) act_out[s] = a ? 1 : 0;
        if (uts) {
            if (wb & WB_PRESENT) {
                if (f & F_NEW) B.uts_present[s] = 1;
                else if (!(f & F_PRES) && (sm.side[ls] & SD_ORIGPRES)) B.uts_present[s] = 0;
            }
            if ((wb & WB_STATE) && (f & F_FRESH)) {
                B.uts_a[s] = sm.st[0][ls]; B.uts_b[s] = sm.st[1][ls]; B.uts_c[s] = sm.st[2][ls]; B.uts_tau[s] = sm.st[3][ls];
                B.uts_p00[s] = sm.st[4][ls]; B.uts_p01[s] = sm.st[5][ls]; B.uts_p11[s] = sm.st[6][ls]; B.uts_p22[s] = sm.st[7][ls];
                B.uts_lik[s] = sm.lik[ls];
                if (f & F_NEW) B.uts_rank[s] = sm.rank[ls];
            }
            if (f & F_PRES) {
                if (wb & WB_PRIOR) B.uts_prior[s] = sm.prior[ls];
                if (wb & WB_W) B.uts_w[s] = sm.w[ls];
                if ((wb & WB_UTSX) && (f & (F_RW | F_FRESH))) {
                    B.uts_side[s] = (int8_t)(sm.side[ls] & 3);
                    if (!(f & F_RW)) B.uts_lrn[s] = NAN; // fresh entry never reweighted: no lr_layer_norm yet
                }
            }
        } else {
            if ((wb & WB_PRESENT) && !(f & F_PRES) && (sm.side[ls] & SD_ORIGPRES)) B.tse_present[s] = 0;
            if (f & F_PRES) {
                if (wb & WB_PRIOR) B.tse_prior[s] = sm.prior[ls];
                if (wb & WB_W) B.tse_w[s] = sm.w[ls];
            }
        }
    }
    if (wb & WB_COUNT_ACTIVE) {
        if (n_act) atomicAdd(&sm.cnt[CNT_ACTIVE], n_act);
        if (n_chg) atomicAdd(&sm.cnt[CNT_CHANGED], n_chg);
    }
    __syncthreads();
    if (tid < GTF_NCOUNTERS && sm.cnt[tid]) {
        if (tid == CNT_REFERR) atomicOr(&B.counters[tid], (unsigned long long)sm.cnt[tid]);
        else atomicAdd(&B.counters[tid], (unsigned long long)sm.cnt[tid]);
    }
}

// per-source sequential multiple-scattering prefix (extrapolate_merged_states.py:114-128, quirk 2):
// the k-th ACTIVE successor of u sees merged_cov[1,1] + sum_{j<=k} var_ms_j, summed left to right in
// adjacency order; the total stays on the node.
__global__ void k_prefix(DevBatch B, GtfGeom g)
{
    int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= B.N) return;
    double p = B.m_p11[u];
    bool ok = B.has_merged[u] && (B.node_ok[u] & (NF_OK | NF_MULTI)) == (NF_OK | NF_MULTI);
    if (ok) {
        double a = B.m_a[u], b = B.m_b[u], ur = B.r[u], uz = B.z[u];
        for (int o = B.out_off[u]; o < B.out_off[u + 1]; o++) {
            int s = B.out_slot[o], v = B.slot_dst[s];
            if (B.active[s] != 1 || !(B.all_alive || B.alive[v])) continue;
            double vms = gtf_var_ms(a, b, B.x[v], B.r[v] - ur, B.z[v] - uz, uz, g.endcap);
            p += vms;
            B.slot_p11[s] = p;
            B.slot_vms[s] = vms;
        }
    }
    B.node_p11tot[u] = p;
}

