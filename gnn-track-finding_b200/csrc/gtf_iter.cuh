// gtf_iter.cuh -- one message-passing iteration on the packed layout (DevPack, gtf_dev.cuh):
//   k_begin    thread per word/node: next activation bitmap := current, presence snapshot, counters, carried p11
//   k_send     CTA per source tile : which out-edges carry a message (extrapolate...py:416,425,431), Highland term from
//                                   the per-out-edge static records, the per-source running sum of merged_cov[1,1]
//                                   (quirk 2, summed in successor order) -> compact source-major message list
//   k_exec     thread per message : extrapolate, chi2 gate, Kalman update (extrapolate_merged_states.py:26-402);
//                                   writes the 64 B state record and the 32 B weight record of the receiving dict
//                                   entry; software-pipelined (cp.async descriptors, gathers between the two halves)
//   k_node2    thread per node    : scans the node's bits; <= 2 dict entries -> closed-form priors / side norms /
//                                   re-weighting / pruning right here (helper.py:30-200); >= 3 -> binned lists
//   k_hv<G>    G lanes per node   : cooperative nodes (3..32 entries), G = 4, 8, 16, 32 lanes per node, entries in
//                                   registers in dict order: priors, re-weighting, pairwise chi2 + greedy KL merge
//                                   (clustering.py:193-307), degree, mixture weights, priors
//   k_big      CTA per node       : more than 32 entries (generic shared-memory program of gtf_tile.cuh)
//   k_pack_* / k_unpack_*         : SoA fields <-> packed records and bitmaps
// Activation / presence flags are bitmaps (1.6 MB per 12.8 M slots: L2 resident, so the scattered tests of k_send
// and the per-node scans cost no DRAM traffic); state, weights, geometry + tag and merged states are
// array-of-records so that an entry is read and written as whole 32 B sectors (a scattered 8 B store is a DRAM
// read-modify-write).
#pragma once

#define H_EX 1u
#define H_ACT 2u
#define H_PRES 4u
#define H_NEW 8u
#define H_RW 32u
#define H_ORIG 64u   // activated flag of the committed state
#define H_ACT0 128u  // activated flag as loaded (after the extrapolation gate)

__device__ __forceinline__ bool bm_get(const uint32_t *bm, int s) { return (bm[s >> 5] >> (s & 31)) & 1u; }
// 32 bits starting at bit p
__device__ __forceinline__ uint32_t bm_win(const uint32_t *bm, int p)
{
    const uint32_t lo = bm[p >> 5], hi = bm[(p >> 5) + 1];
    return __funnelshift_r(lo, hi, p & 31);
}
__device__ __forceinline__ void bm_clear(uint32_t *bm, int s) { atomicAnd(&bm[s >> 5], ~(1u << (s & 31))); }
__device__ __forceinline__ void bm_set(uint32_t *bm, int s) { atomicOr(&bm[s >> 5], 1u << (s & 31)); }

// per-slot auxiliary record: static geometry of the source hit (16 B) and the rarely written tag (8 B) share one sector
__device__ __forceinline__ GeoRec *geo_p(const DevPack &K, int s) { return &K.aux[s].g; }
__device__ __forceinline__ TagRec *tag_p(const DevPack &K, int s) { return &K.aux[s].t; }
// 16 B geometry record, streaming (read once per iteration)
__device__ __forceinline__ GeoRec ld_geo(const GeoRec *p)
{
    const int4 v = __ldcs(reinterpret_cast<const int4 *>(p));
    GeoRec g;
    g.sx = __hiloint2double(v.y, v.x); g.lay = v.z; g.src = v.w;
    return g;
}

__device__ __forceinline__ void flush_counters(unsigned int *s_cnt, unsigned long long *counters, int tid)
{
    if (tid < GTF_NCOUNTERS && s_cnt[tid]) {
        if (tid == CNT_REFERR) atomicOr(&counters[tid], (unsigned long long)s_cnt[tid]);
        else atomicAdd(&counters[tid], (unsigned long long)s_cnt[tid]);
    }
}

// ------------------------------------------------------------------------------------------------ pack / unpack
__global__ void k_pack_slots(DevBatch B, DevPack K, int do_static, int do_act, int do_pres, int do_rec)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    const bool in = s < B.E;
    int src = -1, dst = 0;
    if (in) { src = B.in_src[s]; dst = B.slot_dst[s]; }
    const bool ex = in && src >= 0 && B.alive[src] && B.alive[dst];
    const unsigned mex = __ballot_sync(0xffffffffu, ex);
    const unsigned mact = __ballot_sync(0xffffffffu, in && B.active[s] == 1);
    const unsigned mpres = __ballot_sync(0xffffffffu, in && B.uts_present[s] != 0);
    const unsigned min_ = __ballot_sync(0xffffffffu, in);
    if (lane == 0 && min_) {
        if (do_act) { K.act[s >> 5] = mact; K.exists[s >> 5] = mex; }
        if (do_pres) K.pres[s >> 5] = mpres;
        if (do_act && mex != min_) atomicAdd(&K.counts[PK_MISSING], __popc(min_ & ~mex));
    }
    if (!in) return;
    if (do_static) {
        GeoRec gr;
        gr.sx = src >= 0 ? B.x[src] : 0.0;
        gr.lay = src >= 0 ? B.layer[src] : -1;
        gr.src = src;
        (*geo_p(K, s)) = gr;
    }
    if (do_rec) {
        double2 *st = reinterpret_cast<double2 *>(K.state + 8 * (size_t)s);
        st[0] = make_double2(B.uts_a[s], B.uts_b[s]);
        st[1] = make_double2(B.uts_c[s], B.uts_tau[s]);
        st[2] = make_double2(B.uts_p00[s], B.uts_p01[s]);
        st[3] = make_double2(B.uts_p11[s], B.uts_p22[s]);
        MetaRec m;
        m.w = B.uts_w[s]; m.lik = B.uts_lik[s]; m.prior = B.uts_prior[s]; m.ew = B.edge_w[s];
        K.meta[s] = m;
        TagRec t;
        t.rank = B.uts_rank[s]; t.side = B.uts_side[s]; t.pad = 0; t.lrn = -1;
        (*tag_p(K, s)) = t;
    }
}
// activation / presence bytes -> bitmaps only (the existing-edge bitmap and everything else is current): the per-step
// path of a host that uploads the flags every iteration
__global__ void k_pack_bits(DevBatch B, DevPack K, int do_act, int do_pres)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = s < B.E;
    const unsigned mact = __ballot_sync(0xffffffffu, do_act && in && B.active[s] == 1);
    const unsigned mpres = __ballot_sync(0xffffffffu, do_pres && in && B.uts_present[s] != 0);
    if ((threadIdx.x & 31) == 0 && in) {
        if (do_act) K.act[s >> 5] = mact;
        if (do_pres) K.pres[s >> 5] = mpres;
    }
}
__global__ void k_pack_out(DevBatch B, DevPack K)
{
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= B.E) return;
    const int s = B.out_slot[o];
    const int u = B.in_src[s], v = B.slot_dst[s], rs = B.rev_slot[s];
    K.out_dst[o] = v;
    OutRec r;
    gtf_var_ms_geo(B.r[v] - B.r[u], B.z[v] - B.z[u], r.sin_t, r.rdz);
    r.xk = B.x[v];
    const bool has = rs >= 0 && B.tse_present[rs];          // extrapolate...py:384: u's seed entry for this neighbour
    r.w = has ? B.tse_w[rs] : __longlong_as_double(GTF_NO_TSE_BITS);
    K.orec[o] = r;
}
__global__ void k_pack_nodes(DevBatch B, DevPack K, int do_static, int do_merged)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (do_static && i <= B.N) {
        SrcRec r;
        r.z = i < B.N ? B.z[i] : 0.0; r.off = B.out_off[i]; r.pad = 0;
        K.srec[i] = r;
    }
    if (i >= B.N) return;
    if (do_static) {
        NodeXYZR v;
        v.x = B.x[i]; v.y = B.y[i]; v.z = B.z[i]; v.r = B.r[i];
        K.xyzr[i] = v;
    }
    if (do_merged) {
        MergedRec m;
        m.a = B.m_a[i]; m.b = B.m_b[i]; m.c = B.m_c[i]; m.p00 = B.m_p00[i]; m.p01 = B.m_p01[i]; m.p22 = B.m_p22[i];
        m.prior = B.m_prior[i]; m.cl_p11 = NAN;
        K.mrec[i] = m;
        K.mab[i] = make_double2(m.a, m.b);
    }
}
__global__ void k_unpack_nodes(DevBatch B, DevPack K)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B.N) return;
    const MergedRec m = K.mrec[i];
    B.m_a[i] = m.a; B.m_b[i] = m.b; B.m_c[i] = m.c; B.m_p00[i] = m.p00; B.m_p01[i] = m.p01; B.m_p22[i] = m.p22;
    B.m_prior[i] = m.prior;
}
__global__ void k_unpack_slots(DevBatch B, DevPack K, int do_act, int do_pres, int do_rec)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= B.E) return;
    if (do_act) B.active[s] = bm_get(K.act, s) ? 1 : 0;
    if (do_pres) B.uts_present[s] = bm_get(K.pres, s) ? 1 : 0;
    if (do_rec) {
        const MetaRec m = K.meta[s];
        B.edge_w[s] = m.ew;
        if (bm_get(K.pres, s)) {         // (records of absent entries are scratch: cluster() on the seed dict borrows them)
            const double2 *st = reinterpret_cast<const double2 *>(K.state + 8 * (size_t)s);
            double2 v0 = st[0], v1 = st[1], v2 = st[2], v3 = st[3];
            B.uts_a[s] = v0.x; B.uts_b[s] = v0.y; B.uts_c[s] = v1.x; B.uts_tau[s] = v1.y;
            B.uts_p00[s] = v2.x; B.uts_p01[s] = v2.y; B.uts_p11[s] = v3.x; B.uts_p22[s] = v3.y;
            B.uts_w[s] = m.w; B.uts_lik[s] = m.lik; B.uts_prior[s] = m.prior;
            const TagRec t = (*tag_p(K, s));
            B.uts_rank[s] = t.rank; B.uts_side[s] = t.side;
            if (t.lrn >= 0) B.uts_lrn[s] = t.lrn == 0 ? NAN : (double)t.lrn;
        }
    }
}

// the SEED dict (track_state_estimates) in the packed layout, so that cluster() on the seeds runs on the fast node kernels:
// every in-slot's entry in slot order (= dict order, stamp = slot index), bitmaps, geometry
__global__ void k_pack_tse(DevBatch B, DevPack K)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    const bool in = s < B.E;
    int src = -1, dst = 0;
    if (in) { src = B.in_src[s]; dst = B.slot_dst[s]; }
    const bool ex = in && src >= 0 && B.alive[src] && B.alive[dst];
    const bool pres = in && B.tse_present[s] != 0;
    const unsigned mex = __ballot_sync(0xffffffffu, ex);
    const unsigned mact = __ballot_sync(0xffffffffu, in && B.active[s] == 1);
    const unsigned mpres = __ballot_sync(0xffffffffu, pres);
    const unsigned min_ = __ballot_sync(0xffffffffu, in);
    if (lane == 0 && min_) {
        K.act[s >> 5] = mact; K.exists[s >> 5] = mex; K.pres[s >> 5] = mpres;
        if (mex != min_) atomicAdd(&K.counts[PK_MISSING], __popc(min_ & ~mex));
    }
    if (!in) return;
    GeoRec gr;
    gr.sx = src >= 0 ? B.x[src] : 0.0;
    gr.lay = src >= 0 ? B.layer[src] : -1;
    gr.src = src;
    (*geo_p(K, s)) = gr;
    MetaRec m;                          // (the edge attribute `mixture_weight` rides in the record of every slot)
    m.w = 0.0; m.lik = 0.0; m.prior = 0.0; m.ew = B.edge_w[s];
    if (pres) {
        double2 *st = reinterpret_cast<double2 *>(K.state + 8 * (size_t)s);
        st[0] = make_double2(B.tse_a[s], B.tse_b[s]);
        st[1] = make_double2(B.tse_c[s], B.tse_tau[s]);
        st[2] = make_double2(B.tse_p00[s], B.tse_p01[s]);
        st[3] = make_double2(B.tse_p11[s], B.tse_p22[s]);
        m.w = B.tse_w[s]; m.prior = B.tse_prior[s];
    }
    K.meta[s] = m;
    TagRec t;
    t.rank = s; t.side = 0; t.pad = 0; t.lrn = -1;
    (*tag_p(K, s)) = t;
}
__global__ void k_unpack_tse(DevBatch B, DevPack K)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= B.E) return;
    B.active[s] = bm_get(K.act, s) ? 1 : 0;
    if (B.tse_present[s]) {
        const MetaRec m = K.meta[s];
        B.tse_w[s] = m.w; B.tse_prior[s] = m.prior;
    }
}

// ------------------------------------------------------------------------------------------------ k_begin
// start of an iteration: next activation bitmap := current, snapshot of the presence bitmap, list counters := 0,
// accumulated p11 of the nodes carried over (k_send overwrites the ones that send; quirk 2)
__global__ void k_begin(DevBatch B, DevPack K, int words, int reset_counters)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t <= PK_BIG) K.counts[t] = 0;
    if (K.counts[PK_STOP]) return;       // a queued iteration after the loop converged: empty lists, nothing else happens
    if (t < words) { K.act_nx[t] = K.act[t]; K.pres0[t] = K.pres[t]; }
    if (t < B.N) B.m_p11_nx[t] = B.m_p11[t];
    if (reset_counters && t < GTF_NCOUNTERS_ALL) B.counters[t] = 0;
}
// end of a queued iteration of gtf_iterate: its counters go to the loop's table; converged = no activation flag changed
__global__ void k_iter_end(const unsigned long long *counters, unsigned long long *table, int *counts, int stop_when_converged)
{
    const int t = threadIdx.x;
    if (counts[PK_STOP]) return;
    const int it = counts[PK_DONE];
    if (t < GTF_NCOUNTERS_ALL) table[(size_t)it * GTF_NCOUNTERS_ALL + t] = counters[t];
    __syncwarp();
    if (t == 0) {
        counts[PK_DONE] = it + 1;
        if (stop_when_converged && counters[CNT_CHANGED] == 0) counts[PK_STOP] = 1;
    }
}

// ------------------------------------------------------------------------------------------------ k_send
// Persistent CTAs, one tile of whole sources (<= GTF_SEND_SRCS sources, <= GTF_SEND_EDGES out-edges; descriptor table built
// by build_tiles) per loop trip, software-pipelined one tile ahead:
//   prefetch   everything a tile reads that is CONTIGUOUS in memory travels global -> shared with cp.async.bulk (TMA unit,
//              completion on an mbarrier), issued by one thread while the CTA still works on the previous tile: the
//              out-CSR range (out_slot, out_dst), the per-source static records (z, out_off), has_merged / node_ok flags
//              and the accumulated merged_cov[1,1]; the sources' (a, b) are gathered with 16 B cp.async.  What is left on
//              the dependent chain of a tile: the activation-bit tests (L2) and the per-message record gather.
//   phase 0  thread per source : "sends at all" flag (merged state, sub-graph in play), source of every edge
//   phase 1  thread per edge   : does the edge carry a message?  (edge active and existing: extrapolate...py:416,425,431)
//                                -> ordered compaction into shared memory
//   phase 2  thread per message: Highland term var_ms (extrapolate...py:114-124), carried mixture weight (:384)
//   phase 3  thread per source : merged_cov[1,1] as each edge sees it = node value + the terms of the source's earlier
//                                active successors, summed left to right (quirk 2); the total stays on the node
//   phase 4  thread per message: coalesced append to the global source-major list k_exec consumes densely
#ifndef GTF_SEND_THREADS
#define GTF_SEND_THREADS 128
#endif
#ifndef GTF_SEND_EPT
#define GTF_SEND_EPT 3                                   // out-edges per thread
#endif
#ifndef GTF_SEND_MINB
#define GTF_SEND_MINB 8
#endif
#define GTF_SEND_EDGES (GTF_SEND_THREADS * GTF_SEND_EPT)
#define GTF_SEND_SRCS (GTF_SEND_THREADS - 1)             // (+1 offsets: one per thread; <= 255: uint8 source index)
#define GTF_SEND_MPT ((GTF_SEND_EDGES + GTF_SEND_THREADS - 1) / GTF_SEND_THREADS)

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
// global -> shared bulk copy (16 B aligned, size a multiple of 16), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct SendStage {                                       // what one tile reads, as it lies in global memory
    int32_t slot[GTF_SEND_EDGES + 8];                    // out_slot range, from the 16 B boundary below the tile's first edge
    int32_t dst[GTF_SEND_EDGES + 8];                     // out_dst, same range
    SrcRec srec[GTF_SEND_SRCS + 1];                      // (z, out_off) of the tile's sources + the one after
    double2 ab[GTF_SEND_SRCS + 1];                       // merged (a, b) of the sources
    double p11[GTF_SEND_SRCS + 3];                       // merged_cov[1,1], from the 16 B boundary below the first source
    uint8_t hm[GTF_SEND_SRCS + 33], nok[GTF_SEND_SRCS + 33]; // has_merged / node flags, from the 16 B boundary below
};
struct SendWork {
    double m_vms[GTF_SEND_EDGES], m_p11[GTF_SEND_EDGES];
    uint16_t m_le[GTF_SEND_EDGES];                       // local out-edge of every message
    uint16_t first[GTF_SEND_SRCS + 1], last[GTF_SEND_SRCS + 1]; // a source's message range in the tile list
    uint8_t esrc[GTF_SEND_EDGES];                        // local source of every out-edge of the tile
    uint8_t m_src[GTF_SEND_EDGES];
    uint8_t ok[GTF_SEND_SRCS + 1];
    int wsum[GTF_SEND_THREADS / 32];
    int base;
};
struct __align__(16) SendSmem {
    SendStage st[2];
    SendWork w;
    uint64_t full[2];
};
static_assert(offsetof(SendStage, dst) % 16 == 0 && offsetof(SendStage, srec) % 16 == 0 && offsetof(SendStage, ab) % 16 == 0 &&
              offsetof(SendStage, p11) % 16 == 0 && offsetof(SendStage, hm) % 16 == 0 && offsetof(SendStage, nok) % 16 == 0 &&
              sizeof(SendStage) % 16 == 0, "bulk-copy destinations are 16 B aligned");

// issue the loads of tile `d` = (first source, sources, first out-edge, out-edges) into stage `st`
__device__ __forceinline__ void send_prefetch(const DevBatch &B, const DevPack &K, SendStage &st, uint64_t *bar, const int4 d, int tid)
{
    const int u0 = d.x, ns = d.y, o0 = d.z, ne = d.w;
    if (tid == 0) {
        // the stage was last read through the generic proxy (previous tile, finished at a CTA barrier); order those reads
        // before the asynchronous-proxy writes
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const int oa = o0 & ~3, on = ((o0 + ne + 3) & ~3) - oa;
        const int ua = u0 & ~15, un = ((u0 + ns + 15) & ~15) - ua;
        const int pa = u0 & ~1, pn = ((u0 + ns + 1) & ~1) - pa;
        const unsigned b_edges = 4u * on, b_src = 16u * (ns + 1), b_fl = (unsigned)un, b_p = 8u * pn;
#ifndef GTF_SEND_GATHER_AB
        mbar_expect_tx(bar, 2 * b_edges + b_src + 2 * b_fl + b_p + 16u * ns);
        bulk_g2s(st.ab, K.mab + u0, 16u * ns, bar);      // merged (a, b) of the tile's sources: compact array
#else
        mbar_expect_tx(bar, 2 * b_edges + b_src + 2 * b_fl + b_p);
#endif
        if (on) {
            bulk_g2s(st.slot, B.out_slot + oa, b_edges, bar);
            bulk_g2s(st.dst, K.out_dst + oa, b_edges, bar);
        }
        bulk_g2s(st.srec, K.srec + u0, b_src, bar);
        bulk_g2s(st.hm, B.has_merged + ua, b_fl, bar);
        bulk_g2s(st.nok, B.node_ok + ua, b_fl, bar);
        bulk_g2s(st.p11, B.m_p11 + pa, b_p, bar);
    }
#ifdef GTF_SEND_GATHER_AB
    if (tid < ns)                                        // merged (a, b): first 16 B of the 64 B node record
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&st.ab[tid])), "l"(K.mrec + u0 + tid) : "memory");
#endif
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// the tile loop of k_send
__device__ __forceinline__ void send_tiles(const DevBatch &B, const DevPack &K, SendSmem &S, const int4 *__restrict__ tdesc, int n_tiles,
                                           const GtfGeom &g, const int tid)
{
    SendWork &sm = S.w;
    const int lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x;
    int t = blockIdx.x;
    const int4 zero4 = make_int4(0, 0, 0, 0);
    int4 d_cur = t < n_tiles ? tdesc[t] : zero4;
    int4 d_nxt = tdesc[min(t + G, n_tiles - 1)];
    if (t < n_tiles) send_prefetch(B, K, S.st[0], &S.full[0], d_cur, tid);
    for (int it = 0; t < n_tiles; t += G, it++) {
        const int s = it & 1;
        SendStage &st = S.st[s];
        // the next tile travels while this one is processed (its stage was released by the barrier that ended the previous trip)
        const int tn = t + G;
        if (tn < n_tiles) send_prefetch(B, K, S.st[s ^ 1], &S.full[s ^ 1], d_nxt, tid);
        else asm volatile("cp.async.commit_group;" ::: "memory");
        const int4 d_nn = tdesc[min(tn + G, n_tiles - 1)];             // descriptor after next: a register prefetch, first
                                                                       // touched when the trip ends (no select on it here)
        asm volatile("cp.async.wait_group 1;" ::: "memory");          // this tile's (a, b) gathers (issued one trip ago)
        mbar_wait(&S.full[s], (unsigned)(it >> 1) & 1u);
        const int u0 = d_cur.x, ns = d_cur.y, o_base = d_cur.z, ne = d_cur.w;
        const int epad = o_base & 3, upad = u0 & 15, ppad = u0 & 1;
        GTF_BOUND(B, ns >= 1 && ns <= GTF_SEND_SRCS && ne >= 0 && ne <= GTF_SEND_EDGES && u0 >= 0 && u0 + ns <= B.N && o_base >= 0 && o_base + ne <= B.E);
        // ---- phase 0
        if (tid < ns) {
            const int my_off = st.srec[tid].off - o_base, my_end = st.srec[tid + 1].off - o_base;
            GTF_BOUND(B, my_off >= 0 && my_off <= my_end && my_end <= ne);
            sm.ok[tid] = st.hm[upad + tid] && (st.nok[upad + tid] & (NF_OK | NF_MULTI)) == (NF_OK | NF_MULTI);
            sm.first[tid] = 0xffff;
            for (int o = my_off; o < my_end; o++) sm.esrc[o] = (uint8_t)tid;
        }
        __syncthreads();
        // ---- phase 1: GTF_SEND_EPT consecutive edges per thread keep the successor order
        int cnt = 0;
        unsigned mymask = 0;
        const int e0 = tid * GTF_SEND_EPT;
#pragma unroll
        for (int j = 0; j < GTF_SEND_EPT; j++) {
            if (e0 + j < ne) {
                const int sl = st.slot[epad + e0 + j];
                GTF_BOUND(B, sl >= 0 && sl < B.E && sm.esrc[e0 + j] < ns);
                if (sm.ok[sm.esrc[e0 + j]] && bm_get(K.act, sl) && (K.all_exist || bm_get(K.exists, sl))) { mymask |= 1u << j; cnt++; }
            }
        }
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) sm.wsum[warp] = incl;
        __syncthreads();
        int woff = 0, M = 0;
#pragma unroll
        for (int w = 0; w < GTF_SEND_THREADS / 32; w++) {
            if (w < warp) woff += sm.wsum[w];
            M += sm.wsum[w];
        }
        if (M) {                                                     // (uniform over the CTA)
            if (tid == 0) sm.base = atomicAdd(&K.counts[PK_MSG], M); // one global atomic per tile
            {
                int pos = woff + incl - cnt;
#pragma unroll
                for (int j = 0; j < GTF_SEND_EPT; j++)
                    if ((mymask >> j) & 1u) {
                        GTF_BOUND(B, pos >= 0 && pos < M && M <= ne);
                        sm.m_le[pos] = (uint16_t)(e0 + j); sm.m_src[pos] = sm.esrc[e0 + j];
                        pos++;
                    }
            }
            __syncthreads();
            // ---- phase 2
            double wq[GTF_SEND_MPT];
#pragma unroll
            for (int k = 0; k < GTF_SEND_MPT; k++) {
                const int q = tid + k * GTF_SEND_THREADS;
                wq[k] = 0.0;
                if (q < M) {
                    const int sl = sm.m_src[q];
                    const double2 *rp = reinterpret_cast<const double2 *>(K.orec + o_base + sm.m_le[q]);   // one sector, in successor order
                    const double2 r0 = rp[0], r1 = rp[1];
                    const double2 ab = st.ab[sl];
                    wq[k] = r1.y;
                    sm.m_vms[q] = gtf_var_ms_pre(ab.x, ab.y, r0.y, r0.x, r1.x, st.srec[sl].z, g.endcap);
                    if (q == 0 || sm.m_src[q - 1] != sl) sm.first[sl] = (uint16_t)q;
                    if (q == M - 1 || sm.m_src[q + 1] != sl) sm.last[sl] = (uint16_t)q;
                }
            }
            __syncthreads();
            // ---- phase 3
            if (tid < ns && sm.first[tid] != 0xffff) {
                const int q1 = sm.last[tid];
                double p = st.p11[ppad + tid];
                for (int q = sm.first[tid]; q <= q1; q++) {
                    p += sm.m_vms[q];
                    sm.m_p11[q] = p;
                }
                B.m_p11_nx[u0 + tid] = p;
            }
            __syncthreads();
            // ---- phase 4
            const int base = sm.base;
#pragma unroll
            for (int k = 0; k < GTF_SEND_MPT; k++) {
                const int q = tid + k * GTF_SEND_THREADS;
                if (q < M) {
                    const int gq = base + q, le = sm.m_le[q];
                    GTF_BOUND(B, gq >= 0 && gq < B.E && le < ne && st.dst[epad + le] >= 0 && st.dst[epad + le] < B.N);
                    const bool has = __double_as_longlong(wq[k]) != GTF_NO_TSE_BITS;
                    K.msg_desc[gq] = make_int4(st.slot[epad + le] | (has ? 0 : (int)0x80000000), u0 + sm.m_src[q], st.dst[epad + le], 0);
                    K.msg_w[gq] = has ? wq[k] : NAN;
                    K.msg_p11[gq] = sm.m_p11[q];
                    K.msg_vms[gq] = sm.m_vms[q];
                }
            }
        }
        __syncthreads();       // every read of this stage and of the work arrays is done: the next trip may overwrite them
        d_cur = d_nxt;
        d_nxt = d_nn;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
}

__global__ void __launch_bounds__(GTF_SEND_THREADS, GTF_SEND_MINB) k_send(DevBatch B, DevPack Kin, const int4 *__restrict__ tdesc, int n_tiles,
                                                                         GtfGeom g)
{
    if (Kin.counts[PK_STOP] | Kin.counts[PK_SPARSE]) return;
    DevPack K = Kin;
    K.all_exist = Kin.counts[PK_MISSING] == 0; // every slot is an existing edge (counted when the bitmaps were packed)
    extern __shared__ __align__(16) unsigned char send_raw[];
    SendSmem &S = *reinterpret_cast<SendSmem *>(send_raw);
    if (threadIdx.x == 0) {
        mbar_init(&S.full[0], 1);
        mbar_init(&S.full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    send_tiles(B, K, S, tdesc, n_tiles, g, threadIdx.x);
}


// ------------------------------------------------------------------------------------------------ k_send_sparse
// Late iterations of a committed loop: a few percent of the edges are still active, yet the tiled k_send scans every
// out-edge (0.12 ms per 12.8 M edges whatever they carry).  Inside one loop an activation flag only ever falls (the kernels
// clear bits, nothing sets one), so the out-edges active after the loop's first iteration are a superset of everything later
// iterations can send: k_compact_out lists them once per source (successor order kept), k_send_sparse walks the lists --
// thread per source: the running sum of quirk 2 is sequential per source anyway.  Chosen on the device (PK_SPARSE): the
// lists are only built when fewer than a quarter of the slots are active.
__global__ void k_compact_out(DevBatch B, DevPack Kin)
{
    DevPack K = Kin;
    if (K.counts[PK_STOP]) return;
    const unsigned long long active = B.counters[CNT_ACTIVE];      // of the iteration that just ended
    if (active * 4ull > (unsigned long long)B.E) return;           // (uniform over the grid)
    K.all_exist = Kin.counts[PK_MISSING] == 0;
    const int u = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    if (u == 0) K.counts[PK_SPARSE] = 1;
    int n = 0, o0 = 0, o1 = 0;
    if (u < B.N) {
        o0 = K.srec[u].off; o1 = K.srec[u + 1].off;
        for (int o = o0; o < o1; o++) {
            const int sl = B.out_slot[o];
            n += bm_get(K.act, sl) && (K.all_exist || bm_get(K.exists, sl));
        }
    }
    int incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    int base = 0;
    const int tot = __shfl_sync(0xffffffffu, incl, 31);
    if (lane == 31 && tot) base = atomicAdd(&K.counts[PK_CLIST], tot);
    base = __shfl_sync(0xffffffffu, base, 31) + incl - n;
    if (u < B.N) {
        K.c_rng[u] = make_int2(base, n);
        if (n)
            for (int o = o0; o < o1; o++) {
                const int sl = B.out_slot[o];
                if (bm_get(K.act, sl) && (K.all_exist || bm_get(K.exists, sl))) K.c_edge[base++] = make_int4(sl, K.out_dst[o], o, u);
            }
    }
}
// thread per source walking its list.  (A thread per LISTED edge -- one coalesced 16 B read each, the earlier active successors'
// terms recomputed -- was measured too: 78 vs 90 us per iteration with 0.9 M listed edges, but 0.06 ms SLOWER per iteration on the
// bench schedule, where 2.6 M edges are listed and a source keeps several active successors.)
__global__ void __launch_bounds__(256) k_send_sparse(DevBatch B, DevPack K, GtfGeom g)
{
    if (K.counts[PK_STOP] | !K.counts[PK_SPARSE]) return;
    const int u = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    int n = 0;
    int2 rng = make_int2(0, 0);
    unsigned live = 0;                                   // which of the source's first 32 listed edges are still active (later
                                                         // ones are tested again in the second pass: the bitmap does not change)
    if (u < B.N && B.has_merged[u] && (B.node_ok[u] & (NF_OK | NF_MULTI)) == (NF_OK | NF_MULTI)) {
        rng = K.c_rng[u];
        for (int k = 0; k < rng.y; k++) {
            const bool a = bm_get(K.act, K.c_edge[rng.x + k].x);
            if (a && k < 32) live |= 1u << k;
            n += a;
        }
    }
    int incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    int base = 0;
    const int tot = __shfl_sync(0xffffffffu, incl, 31);
    if (lane == 31 && tot) base = atomicAdd(&K.counts[PK_MSG], tot);   // one global atomic per warp
    int gq = __shfl_sync(0xffffffffu, base, 31) + incl - n;
    if (!n) return;
    const double2 ab = K.mab[u];
    const double z = K.srec[u].z;
    double p = B.m_p11[u];
    for (int k = 0; k < rng.y; k++) {
        const int4 e = K.c_edge[rng.x + k];
        if (k < 32 ? !((live >> k) & 1u) : !bm_get(K.act, e.x)) continue;
        const double2 *rp = reinterpret_cast<const double2 *>(K.orec + e.z);
        const double2 r0 = rp[0], r1 = rp[1];            // (sin_t, xk), (rdz, w)
        const double vms = gtf_var_ms_pre(ab.x, ab.y, r0.y, r0.x, r1.x, z, g.endcap);
        p += vms;                                        // quirk 2: summed in successor order
        const bool has = __double_as_longlong(r1.y) != GTF_NO_TSE_BITS;
        GTF_BOUND(B, gq >= 0 && gq < B.E && e.y >= 0 && e.y < B.N);
        K.msg_desc[gq] = make_int4(e.x | (has ? 0 : (int)0x80000000), u, e.y, 0);
        K.msg_w[gq] = has ? r1.y : NAN;
        K.msg_p11[gq] = p;
        K.msg_vms[gq] = vms;
        gq++;
    }
    B.m_p11_nx[u] = p;
}

// ------------------------------------------------------------------------------------------------ k_exec
// thread per message: extrapolate, chi2 gate, Kalman update (extrapolate_merged_states.py:26-402); writes the state
// record and the weight record of the receiving dict entry, clears the activation bit of a gated edge.
#ifndef GTF_EXEC_THREADS
#define GTF_EXEC_THREADS 128
#endif
#ifndef GTF_EXEC_MINB
#define GTF_EXEC_MINB 4
#endif
__global__ void __launch_bounds__(GTF_EXEC_THREADS, GTF_EXEC_MINB) k_exec(DevBatch B, DevPack K, double chi2_cut, GtfGeom g, int record_chi2)
{
    __shared__ unsigned int s_cnt[GTF_NCOUNTERS];
    const int tid = threadIdx.x;
    if (tid < GTF_NCOUNTERS) s_cnt[tid] = 0;
    __syncthreads();
    const int count = K.counts[PK_MSG];
    const int stride = gridDim.x * GTF_EXEC_THREADS;
    unsigned gated = 0, sent = 0;
    // software pipeline, two stages deep: the DESCRIPTOR (slot, source, destination) of the message after next travels
    // to shared memory with cp.async (no register held), the GATHERS it indexes for the next message are issued between
    // the two halves of the extrapolation, so both latencies hide behind the algebra (only 16 warps per SM fit the
    // 128 registers)
    struct In {
        int sraw;
        double ux, uy, uz, ur, vx, vy, vz, vr, a, b, c, p00, p01, p22, w, p, vms;
    };
    __shared__ int4 s_desc[GTF_EXEC_THREADS];
    const unsigned s_addr = (unsigned)__cvta_generic_to_shared(&s_desc[tid]);
    auto fetch_desc = [&](int q) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n cp.async.commit_group;" ::"r"(s_addr), "l"(K.msg_desc + q) : "memory");
    };
    auto load = [&](int q, const int4 d, In &x) {
        const int u = d.y, v = d.z;
        GTF_BOUND(B, u >= 0 && u < B.N && v >= 0 && v < B.N);
        const NodeXYZR U = K.xyzr[u], V = K.xyzr[v];
        x.sraw = d.x;
        x.ux = U.x; x.uy = U.y; x.uz = U.z; x.ur = U.r; x.vx = V.x; x.vy = V.y; x.vz = V.z; x.vr = V.r;
        const double2 *mr = reinterpret_cast<const double2 *>(K.mrec + u);
        const double2 r0 = mr[0], r1 = mr[1], r2 = mr[2];
        x.a = r0.x; x.b = r0.y; x.c = r1.x; x.p00 = r1.y; x.p01 = r2.x; x.p22 = r2.y;
        x.w = __ldcs(K.msg_w + q); x.p = __ldcs(K.msg_p11 + q); x.vms = __ldcs(K.msg_vms + q);
    };
    int q = blockIdx.x * GTF_EXEC_THREADS + tid;
    In cur;
    if (q < count) {
        load(q, __ldcs(K.msg_desc + q), cur);
        if (q + stride < count) fetch_desc(q + stride);
    }
    while (q < count) {
        const int s = cur.sraw & 0x7fffffff;
        const bool notse = cur.sraw < 0;
        GTF_BOUND(B, s >= 0 && s < B.E);
        const double w = cur.w, p = cur.p, vms = cur.vms, p00 = cur.p00, p01 = cur.p01, p22 = cur.p22;
        const double dr = cur.vr - cur.ur, dz = cur.vz - cur.uz, uz = cur.uz, vz = cur.vz;
        GtfJac J;
        gtf_extrap_jac(cur.ux, cur.uy, cur.vx, cur.vy, cur.a, cur.b, cur.c, J);
        const int qn = q + stride;
        if (qn < count) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            const int4 d = s_desc[tid];                 // descriptor of the next message: fetched one iteration ago
            if (qn + stride < count) fetch_desc(qn + stride);
            load(qn, d, cur);
        }
        GtfExtrapOut o;
        gtf_extrap_update(J, dr, dz, uz, vz, p00, p01, p, p22, vms, chi2_cut, g, o);
        if (record_chi2) B.uts_chi2[s] = o.chi2; // diagnostic only (the reference appends it to a CSV): a partial-sector write
        sent++;
        if (o.pass) {
            if (notse) atomicOr(&s_cnt[CNT_REFERR], (unsigned)GTF_REF_NO_TSE);
            double2 *st = reinterpret_cast<double2 *>(K.state + 8 * (size_t)s);
            __stcs(st + 0, make_double2(o.s.a, o.s.b));
            __stcs(st + 1, make_double2(o.s.c, o.s.tau));
            __stcs(st + 2, make_double2(o.s.p00, o.s.p01));
            __stcs(st + 3, make_double2(o.s.p11, o.s.p22));
            // a fresh dict entry has no prior / lr_layer_norm / side yet (prior = NaN is what marks it for the node
            // kernels, which reset side / lrn in the tag record); its edge weight is set by the re-weighting that always
            // follows in this iteration.  Whole-sector write.
            double2 *m = reinterpret_cast<double2 *>(K.meta + s);
            __stcs(m + 0, make_double2(w, o.lik));
            __stcs(m + 1, make_double2(NAN, NAN));
            bm_set(K.pres, s); // fire and forget; 'inserted this pass' = present now and not in the snapshot k_begin took
        } else {
            bm_clear(K.act_nx, s); // :393
            gated++;
        }
        near_note(B, GTF_NEAR_GATE, s, o.chi2, chi2_cut);
        q = qn;
    }
    if (sent) atomicAdd(&s_cnt[CNT_SENT], sent);
    if (gated) atomicAdd(&s_cnt[CNT_GATED], gated);
    __syncthreads();
    flush_counters(s_cnt, B.counters, tid);
}

// ------------------------------------------------------------------------------------------------ k_sx
// k_send and k_exec as ONE warp-specialised kernel: the message list never goes to global memory (40 B written and 40 B read
// back per message: a fifth of the iteration's DRAM traffic) and the memory-bound scan overlaps the fp64-bound update on the
// same SM.  A CTA = one SCANNING warp group (warps 0-3, 40 registers after setmaxnreg.dec) + one EXECUTING warp group
// (warps 4-7: the k_exec message program, 120 registers after setmaxnreg.inc -- the pool is the CTA's own 256 x 80);
// three CTAs per SM.
//   scan     twelve scanning warps per SM cannot hide a tile's dependent global loads the way k_send's 32 do, so the tile loop
//            is a two-stage software pipeline: trip i runs stage A of tile i+1 (flags -> ordered compaction -> descriptors
//            into the ring, per-message OutRec gathers STARTED with cp.async straight into the ring slots) around stage B
//            of tile i (gathers landed a trip ago: Highland terms, per-source running sums, publish); the bitmap words A
//            tests are loaded before B and used after it, the tile inputs arrive by cp.async.bulk two trips ahead.
//   ring     shared memory, GTF_RING entries of 48 B: (slot | no-seed-entry bit, source, destination) + the OutRec gathered
//            in place, rewritten by stage B to (var_ms, merged_cov[1,1] as the edge sees it, -, carried weight).  The scanners
//            publish `tail`; executing warp w takes the chunks of 32 consecutive positions c = w, w + 4, ...: dense warps
//            whatever the tiles' message counts are.
//   execute  software-pipelined like k_exec: the next chunk's ring entries and the gathers they index are fetched between
//            the Jacobian half and the covariance half of the current one when the chunk is already there, otherwise at the
//            top of the next trip.
#ifndef GTF_SX_CTAS
#define GTF_SX_CTAS 3
#endif
#ifndef GTF_SX_SCAN_REGS
#define GTF_SX_SCAN_REGS 40
#endif
#ifndef GTF_SX_EXEC_REGS
#define GTF_SX_EXEC_REGS 120
#endif
#define GTF_STR2(x) #x
#define GTF_STR(x) GTF_STR2(x)
#define GTF_RING 800                                     // >= 2 GTF_SEND_EDGES (a tile not yet published + the next one's bound) + 31
                                                         // (an incomplete chunk nobody can take yet): the scanners never wait for
                                                         // space that only they could free
struct __align__(32) SxRec { double vms, p11, rdz, w; }; // as gathered: OutRec (sin_t, xk, rdz, w)
struct SxRing {
    SxRec rec[GTF_RING];
    int4 desc[GTF_RING];
    int tail, done;                                      // messages published so far; no more will come
    int cons_next[4];                                    // per executing warp: first ring position it has not released yet
};
struct SxWork {                                          // per tile in flight (stage A of tile i+1 / stage B of tile i)
    uint16_t first[GTF_SEND_SRCS + 1], last[GTF_SEND_SRCS + 1];
    uint8_t m_src[GTF_SEND_EDGES];
};
struct __align__(32) SxSmem {
    SxRing ring;
    SendStage st[4];
    SxWork w[2];
    int4 tdesc[8];                                       // descriptors of tiles i .. i+4 (cp.async, two trips before their first use)
    uint64_t full[4];
    uint8_t esrc[GTF_SEND_EDGES], ok[GTF_SEND_SRCS + 1];
    int wsum[4];
    unsigned int cnt[GTF_NCOUNTERS];
};
static_assert((sizeof(SxSmem) + 1024) * GTF_SX_CTAS <= 228 * 1024, "k_sx: shared memory of the resident CTAs (228 KB per SM, 1 KB reserved per CTA)");
static_assert(sizeof(OutRec) == sizeof(SxRec) && GTF_RING >= 2 * GTF_SEND_EDGES + 31 && GTF_RING % 32 == 0, "ring");
__device__ __forceinline__ int ld_acq_s(const int *p)
{
    int v;
    asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_rel_s(int *p, int v)
{
    asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void sx_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the scanning warp group only

// the scanning warp group of k_sx.  Trip i: compaction of tile i+1 (its activation words were loaded a trip ago), the words of
// tile i+2, then the messages of tile i (their record gathers were started a trip ago); tile i+3's inputs are on their way.
__device__ __forceinline__ void sx_scan(const DevBatch &B, const DevPack &K, SxSmem &S, const int4 *__restrict__ tdesc, int n_tiles,
                                        const GtfGeom &g, const int tid)
{
    SxRing &R = S.ring;
    const int lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x;
    const int n_my = (n_tiles - (int)blockIdx.x + G - 1) / G;          // tiles of this CTA: blockIdx.x + j G
    if (tid < 5 && tid < n_my) S.tdesc[tid] = tdesc[blockIdx.x + tid * G];
    sx_sync();
    // the bulk copies complete on their stage's mbarrier; no cp.async group belongs to them
    auto prefetch = [&](int j) {
        if (tid == 0) {
            const int4 d = S.tdesc[j & 7];
            SendStage &st = S.st[j & 3];
            uint64_t *bar = &S.full[j & 3];
            const int u0 = d.x, ns = d.y, o0 = d.z, ne = d.w;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const int oa = o0 & ~3, on = ((o0 + ne + 3) & ~3) - oa;
            const int ua = u0 & ~15, un = ((u0 + ns + 15) & ~15) - ua;
            const int pa = u0 & ~1, pn = ((u0 + ns + 1) & ~1) - pa;
            const unsigned b_edges = 4u * on, b_src = 16u * (ns + 1), b_fl = (unsigned)un, b_p = 8u * pn;
            mbar_expect_tx(bar, 2 * b_edges + b_src + 2 * b_fl + b_p + 16u * ns);
            bulk_g2s(st.ab, K.mab + u0, 16u * ns, bar);
            if (on) {
                bulk_g2s(st.slot, B.out_slot + oa, b_edges, bar);
                bulk_g2s(st.dst, K.out_dst + oa, b_edges, bar);
            }
            bulk_g2s(st.srec, K.srec + u0, b_src, bar);
            bulk_g2s(st.hm, B.has_merged + ua, b_fl, bar);
            bulk_g2s(st.nok, B.node_ok + ua, b_fl, bar);
            bulk_g2s(st.p11, B.m_p11 + pa, b_p, bar);
        }
    };
    const int e0 = tid * GTF_SEND_EPT;
    unsigned wa[GTF_SEND_EPT];                           // activation words of this thread's out-edges of the tile in stage A
    // every thread observes the stage's mbarrier itself (visibility of the asynchronous-proxy writes), then reads the words
    auto act_words = [&](int j) {
        const int4 d = S.tdesc[j & 7];
        mbar_wait(&S.full[j & 3], (unsigned)(j >> 2) & 1u);
        const SendStage &st = S.st[j & 3];
        const int epad = d.z & 3, ne = d.w;
#pragma unroll
        for (int k = 0; k < GTF_SEND_EPT; k++) {
            wa[k] = 0;
            if (e0 + k < ne) {
                const int sl = st.slot[epad + e0 + k];
                GTF_BOUND(B, sl >= 0 && sl < B.E);
                wa[k] = K.act[sl >> 5];
            }
        }
    };
    for (int j = 0; j < 3 && j < n_my; j++) prefetch(j);
    act_words(0);
    int tail = 0;                                        // ring position after the last tile that went through stage A
    int M_b = 0, base_b = 0;                             // messages of the tile in stage B and their first ring position
    for (int i = -1; i < n_my; i++) {
        if (tid == 0 && i + 5 < n_my && i >= 0)           // descriptor of tile i+5: part of this trip's cp.async group, complete
                                                         // before stage B of the next trip, first read two trips from now
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&S.tdesc[(i + 5) & 7])), "l"(tdesc + blockIdx.x + (i + 5) * G) : "memory");
        if (i + 3 < n_my && i >= 0) prefetch(i + 3);     // (its stage was last read by tile i-1: done before the previous trip's last barrier)
        // ---- stage A, tile i+1: which sources send at all, ordered compaction, descriptors and record gathers into the ring
        const bool has_a = i + 1 < n_my;
        int M_a = 0;
        if (has_a) {
            const int4 d = S.tdesc[(i + 1) & 7];
            const int u0a = d.x, nsa = d.y, oba = d.z, nea = d.w, epad = d.z & 3;
            GTF_BOUND(B, nsa >= 1 && nsa <= GTF_SEND_SRCS && nea >= 0 && nea <= GTF_SEND_EDGES && u0a >= 0 && u0a + nsa <= B.N && oba >= 0 && oba + nea <= B.E);
            SendStage &sa = S.st[(i + 1) & 3];
            SxWork &Wa = S.w[(i + 1) & 1];
            if (tid < nsa) {
                const int upad = u0a & 15;
                const int my_off = sa.srec[tid].off - oba, my_end = sa.srec[tid + 1].off - oba;
                GTF_BOUND(B, my_off >= 0 && my_off <= my_end && my_end <= nea);
                S.ok[tid] = sa.hm[upad + tid] && (sa.nok[upad + tid] & (NF_OK | NF_MULTI)) == (NF_OK | NF_MULTI);
                Wa.first[tid] = 0xffff;
                for (int o = my_off; o < my_end; o++) S.esrc[o] = (uint8_t)tid;
            }
            if (tid == 0)                                                // room for the tile's messages (at most nea)?
                for (;;) {
                    const int head = min(min(ld_acq_s(&R.cons_next[0]), ld_acq_s(&R.cons_next[1])),
                                         min(ld_acq_s(&R.cons_next[2]), ld_acq_s(&R.cons_next[3])));
                    if (tail + nea - head <= GTF_RING) break;
                    __nanosleep(100);
                }
            sx_sync();
            int cnt = 0;
            unsigned mymask = 0;
#pragma unroll
            for (int j = 0; j < GTF_SEND_EPT; j++) {
                if (e0 + j < nea && S.ok[S.esrc[e0 + j]]) {
                    const int sl = sa.slot[epad + e0 + j];
                    if (((wa[j] >> (sl & 31)) & 1u) && (K.all_exist || bm_get(K.exists, sl))) { mymask |= 1u << j; cnt++; }
                }
            }
            int incl = cnt;
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, dd);
                if (lane >= dd) incl += v;
            }
            if (lane == 31) S.wsum[warp] = incl;
            sx_sync();
            int woff = 0;
#pragma unroll
            for (int w = 0; w < 4; w++) {
                if (w < warp) woff += S.wsum[w];
                M_a += S.wsum[w];
            }
            int pos = woff + incl - cnt;
#pragma unroll
            for (int j = 0; j < GTF_SEND_EPT; j++)
                if ((mymask >> j) & 1u) {
                    GTF_BOUND(B, pos >= 0 && pos < M_a && M_a <= nea);
                    const int rp = (tail + pos) % GTF_RING, le = e0 + j, src = S.esrc[le];
                    GTF_BOUND(B, sa.dst[epad + le] >= 0 && sa.dst[epad + le] < B.N);
                    Wa.m_src[pos] = (uint8_t)src;
                    R.desc[rp] = make_int4(sa.slot[epad + le], u0a + src, sa.dst[epad + le], 0);
                    const unsigned dsts = smem_u32(&R.rec[rp]);
                    const OutRec *gp = K.orec + oba + le;            // one sector, in successor order
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n cp.async.cg.shared.global [%2], [%3], 16;"
                                 ::"r"(dsts), "l"(gp), "r"(dsts + 16), "l"(reinterpret_cast<const char *>(gp) + 16) : "memory");
                    pos++;
                }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        // ---- the activation words of tile i+2 travel while stage B runs
        if (i + 2 < n_my) act_words(i + 2);
        // ---- stage B, tile i: its record gathers are the group before the one just committed
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        sx_sync();       // tile i's gathers, tile i+1's compaction and work arrays are complete
        if (i >= 0 && M_b) {
            SendStage &sb = S.st[i & 3];
            SxWork &Wb = S.w[i & 1];
            const int4 d = S.tdesc[i & 7];
            const int u0 = d.x, ns = d.y, ppad = d.x & 1;
#pragma unroll
            for (int k = 0; k < GTF_SEND_MPT; k++) {
                const int q = tid + k * GTF_SEND_THREADS;
                if (q < M_b) {
                    const int rp = (base_b + q) % GTF_RING;
                    const int sl = Wb.m_src[q];
                    const SxRec r = R.rec[rp];       // (sin_t, xk, rdz, w)
                    const double2 ab = sb.ab[sl];
                    R.rec[rp].vms = gtf_var_ms_pre(ab.x, ab.y, r.p11, r.vms, r.rdz, sb.srec[sl].z, g.endcap);
                    if (__double_as_longlong(r.w) == GTF_NO_TSE_BITS) { R.rec[rp].w = NAN; R.desc[rp].x |= (int)0x80000000; }
                    if (q == 0 || Wb.m_src[q - 1] != sl) Wb.first[sl] = (uint16_t)q;
                    if (q == M_b - 1 || Wb.m_src[q + 1] != sl) Wb.last[sl] = (uint16_t)q;
                }
            }
            sx_sync();
            if (tid < ns && Wb.first[tid] != 0xffff) {               // quirk 2: summed in successor order
                const int q1 = Wb.last[tid];
                double p = sb.p11[ppad + tid];
                int rp = (base_b + Wb.first[tid]) % GTF_RING;
                for (int q = Wb.first[tid]; q <= q1; q++) {
                    p += R.rec[rp].vms;
                    R.rec[rp].p11 = p;
                    rp = rp + 1 == GTF_RING ? 0 : rp + 1;
                }
                B.m_p11_nx[u0 + tid] = p;
            }
            sx_sync();   // (also ends the trip: every read of tile i's stage and work arrays is done)
            if (tid == 0) st_rel_s(&R.tail, base_b + M_b);          // the tile's messages are complete: publish them
        }
        base_b = tail;
        M_b = M_a;
        tail += M_a;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (tid == 0) st_rel_s(&R.done, 1);
}

// entries of chunk c that can be taken: 32, fewer for the last chunk, 0 = not there yet (only when !block), -1 = none will come
__device__ __forceinline__ int sx_chunk(SxRing &R, int c, bool block, int lane)
{
    int n = 0;
    if (lane == 0) {
        unsigned ns = 128;
        for (;;) {
            const int done = ld_acq_s(&R.done);
            const int tail = ld_acq_s(&R.tail);
            if (tail >= 32 * (c + 1)) { n = 32; break; }
            if (done) { n = tail > 32 * c ? tail - 32 * c : -1; break; }
            if (!block) break;
            __nanosleep(ns);                             // (a starved warp must not take issue slots from the scanners)
            if (ns < 2048) ns *= 2;
        }
    }
    n = __shfl_sync(0xffffffffu, n, 0);
    __syncwarp();                                        // (orders the other lanes' ring reads after lane 0's acquire)
    return n;
}
__global__ void __launch_bounds__(256, GTF_SX_CTAS) k_sx(DevBatch B, DevPack Kin, const int4 *__restrict__ tdesc, int n_tiles, double chi2_cut,
                                                         GtfGeom g, int record_chi2)
{
    static_assert(GTF_SEND_THREADS == 128, "one warp group scans");
    if (Kin.counts[PK_STOP]) return;
    extern __shared__ __align__(32) unsigned char sx_raw[];
    SxSmem &S = *reinterpret_cast<SxSmem *>(sx_raw);
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&S.full[0], 1);
        mbar_init(&S.full[1], 1);
        mbar_init(&S.full[2], 1);
        mbar_init(&S.full[3], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        S.ring.tail = 0; S.ring.done = 0;
    }
    if (tid < 4) S.ring.cons_next[tid] = 32 * tid;
    if (tid < GTF_NCOUNTERS) S.cnt[tid] = 0;
    __syncthreads();
    if (tid < 128) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 " GTF_STR(GTF_SX_SCAN_REGS) ";");
        DevPack K = Kin;
        K.all_exist = Kin.counts[PK_MISSING] == 0;
        sx_scan(B, K, S, tdesc, n_tiles, g, tid);
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 " GTF_STR(GTF_SX_EXEC_REGS) ";");
        const DevPack &K = Kin;
        SxRing &R = S.ring;
        const int lane = tid & 31, w = (tid >> 5) - 4;
        unsigned gated = 0, sent = 0;
        struct In {
            int sraw;
            double ux, uy, uz, ur, vx, vy, vz, vr, a, b, c, p00, p01, p22, w, p, vms;
        };
        auto load = [&](int pos, In &x) {
            const int rp = pos % GTF_RING;
            const int4 d = R.desc[rp];
            const int u = d.y, v = d.z;
            GTF_BOUND(B, u >= 0 && u < B.N && v >= 0 && v < B.N);
            const NodeXYZR U = K.xyzr[u], V = K.xyzr[v];
            x.sraw = d.x;
            x.ux = U.x; x.uy = U.y; x.uz = U.z; x.ur = U.r; x.vx = V.x; x.vy = V.y; x.vz = V.z; x.vr = V.r;
            const double2 *mr = reinterpret_cast<const double2 *>(K.mrec + u);
            const double2 r0 = mr[0], r1 = mr[1], r2 = mr[2];
            x.a = r0.x; x.b = r0.y; x.c = r1.x; x.p00 = r1.y; x.p01 = r2.x; x.p22 = r2.y;
            const SxRec r = R.rec[rp];
            x.w = r.w; x.p = r.p11; x.vms = r.vms;
        };
        In cur;
        int c = w;                                       // this warp's chunk
        int n_cur = 0;                                   // its entries, 0: not loaded yet
        for (;;) {
            if (n_cur == 0) {
                n_cur = sx_chunk(R, c, true, lane);
                if (n_cur < 0) break;
                if (lane < n_cur) load(32 * c + lane, cur);
            }
            const bool mine = lane < n_cur;
            const int s = cur.sraw & 0x7fffffff;
            const bool notse = cur.sraw < 0;
            GTF_BOUND(B, !mine || (s >= 0 && s < B.E));
            const double cw = cur.w, p = cur.p, vms = cur.vms, p00 = cur.p00, p01 = cur.p01, p22 = cur.p22;
            const double dr = cur.vr - cur.ur, dz = cur.vz - cur.uz, uz = cur.uz, vz = cur.vz;
            GtfJac J;
            if (mine) gtf_extrap_jac(cur.ux, cur.uy, cur.vx, cur.vy, cur.a, cur.b, cur.c, J);
            // the ring entries of chunk c are in registers: release them, then look for the next chunk without waiting
            __syncwarp();
            if (lane == 0) st_rel_s(&R.cons_next[w], 32 * (c + 4));
            const int n_nxt = sx_chunk(R, c + 4, false, lane);
            if (lane < n_nxt) load(32 * (c + 4) + lane, cur);
#ifdef GTF_SX_NOEXEC                                     // ablation build (profiles/r02_k_sx.txt): scan + ring + gathers only
            if (mine && J.f00 == 1.2345e-300) {
#else
            if (mine) {
#endif
                GtfExtrapOut o;
                gtf_extrap_update(J, dr, dz, uz, vz, p00, p01, p, p22, vms, chi2_cut, g, o);
                if (record_chi2) B.uts_chi2[s] = o.chi2;
                sent++;
                if (o.pass) {
                    if (notse) atomicOr(&S.cnt[CNT_REFERR], (unsigned)GTF_REF_NO_TSE);
                    double2 *st = reinterpret_cast<double2 *>(K.state + 8 * (size_t)s);
                    __stcs(st + 0, make_double2(o.s.a, o.s.b));
                    __stcs(st + 1, make_double2(o.s.c, o.s.tau));
                    __stcs(st + 2, make_double2(o.s.p00, o.s.p01));
                    __stcs(st + 3, make_double2(o.s.p11, o.s.p22));
                    double2 *m = reinterpret_cast<double2 *>(K.meta + s);     // (see k_exec)
                    __stcs(m + 0, make_double2(cw, o.lik));
                    __stcs(m + 1, make_double2(NAN, NAN));
                    bm_set(K.pres, s);
                } else {
                    bm_clear(K.act_nx, s); // :393
                    gated++;
                }
                near_note(B, GTF_NEAR_GATE, s, o.chi2, chi2_cut);
            }
            c += 4;
            if (n_nxt < 0) break;
            n_cur = n_nxt;
        }
        if (sent) atomicAdd(&S.cnt[CNT_SENT], sent);
        if (gated) atomicAdd(&S.cnt[CNT_GATED], gated);
    }
    __syncthreads();
    flush_counters(S.cnt, B.counters, tid);
}

// ------------------------------------------------------------------------------------------------ k_node
struct LEnt {          // one dict entry of a light node
    int s;             // slot
    unsigned f;        // H_*
    int lay, rank, side, lrn, rank0, tag0;
    double sx, w, lik, prior, ew;
};
// second word of the tag record: side in byte 0, lr_layer_norm code in the upper half
__device__ __forceinline__ int tag_pack(int side, int lrn) { return (side & 0xff) | (lrn << 16); }
// `pf`: the entry's flag bits when the caller already holds the bitmap windows they live in (H_EX | H_ACT | H_ACT0 | H_ORIG |
// H_NEW), or ~0u: test the bitmaps here
__device__ __forceinline__ void lent_load(const DevBatch &B, const DevPack &K, int s, LEnt &e, unsigned pf = ~0u)
{
    const double2 m0 = __ldcs(reinterpret_cast<const double2 *>(K.meta + s));
    const double2 m1 = __ldcs(reinterpret_cast<const double2 *>(K.meta + s) + 1);
    const int2 tg = __ldcs(reinterpret_cast<const int2 *>(tag_p(K, s)));
    const GeoRec gr = ld_geo(geo_p(K, s));
    e.s = s;
    e.w = m0.x; e.lik = m0.y; e.prior = m1.x; e.ew = m1.y;
    e.rank = tg.x; e.tag0 = tg.y;
    e.side = (int)(int8_t)(tg.y & 0xff);
    e.lrn = tg.y >> 16;
    e.rank0 = tg.x;
    e.sx = gr.sx + 0.0; e.lay = gr.lay;
    unsigned f = H_PRES;
    if (pf != ~0u)
        f |= pf;
    else {
        if (K.all_exist || bm_get(K.exists, s)) f |= H_EX;
        if (bm_get(K.act_nx, s)) f |= H_ACT | H_ACT0;
        if (bm_get(K.act, s)) f |= H_ORIG;
        if (!bm_get(K.pres0, s)) f |= H_NEW;
    }
    if (e.prior != e.prior) { e.side = 0; e.lrn = 0; }   // entry just written by the extrapolation: no side / lr_layer_norm yet
    e.f = f;
}
// flag bits of the entry at bit `k` of the bitmap windows (existing, next activation, activation, presence snapshot)
__device__ __forceinline__ unsigned win_flags(unsigned exw, unsigned anw, unsigned a0w, unsigned p0w, int k)
{
    unsigned f = 0;
    if ((exw >> k) & 1u) f |= H_EX;
    if ((anw >> k) & 1u) f |= H_ACT | H_ACT0;
    if ((a0w >> k) & 1u) f |= H_ORIG;
    if (!((p0w >> k) & 1u)) f |= H_NEW;
    return f;
}
__device__ __forceinline__ void lent_store(const DevBatch &B, const DevPack &K, const LEnt &e)
{
    double2 *m = reinterpret_cast<double2 *>(K.meta + e.s);
    __stcs(m + 0, make_double2(e.w, e.lik));
    __stcs(m + 1, make_double2(e.prior, e.ew));          // ew: helper.py:180
    const int t1 = tag_pack(e.side, e.lrn);
    if (e.rank != e.rank0 || t1 != e.tag0) *reinterpret_cast<int2 *>(tag_p(K, e.s)) = make_int2(e.rank, t1);
    if ((e.f & H_ACT0) && !(e.f & H_ACT)) bm_clear(K.act_nx, e.s);
}
__device__ __forceinline__ void lent_prior(LEnt &a, LEnt &b, int n)
{
    const unsigned m3 = H_PRES | H_EX | H_ACT;
    const bool ea = (a.f & m3) == m3, eb = n == 2 && (b.f & m3) == m3;
    const bool same = ea && eb && a.lay == b.lay;
    if (ea) a.prior = same ? 0.5 : 1.0; // helper.py:61: 1/len(group)
    if (eb) b.prior = same ? 0.5 : 1.0;
}
__device__ __forceinline__ void lent_reweight(const DevBatch &B, unsigned int *cnt, LEnt &a, LEnt &b, int n, double nodex, double thr)
{
    const unsigned m3 = H_PRES | H_EX | H_ACT;
    const bool ea = (a.f & m3) == m3, eb = n == 2 && (b.f & m3) == m3;
    if (!ea && !eb) return;
    const bool la = a.sx < nodex, lb = b.sx < nodex;
    int norm2 = 1;                                                    // len(set(x)) per side (helper.py:127,134)
    if (ea && eb && la == lb && a.sx != b.sx) norm2 = 2;
    const unsigned lf = n == 2 ? b.f : a.f;                           // stale `neighbour_num` (helper.py:131,138)
    if (!(lf & H_EX)) atomicOr(&cnt[CNT_REFERR], (unsigned)GTF_REF_KEY);
    const bool last_active = (lf & (H_EX | H_ACT)) == (H_EX | H_ACT);
    const int norm = last_active ? norm2 : 1;
    double denom = 0.0;                                               // dict order (helper.py:165-169)
    if (ea) denom += a.w * a.lik;
    if (eb) denom += b.w * b.lik;
    unsigned off = 0;
    if (ea) {
        double rw = (a.w * a.lik * a.prior) / denom;
        if (norm != 1) rw = rw / (double)norm;
        a.lrn = norm; a.side = la ? 1 : 2; a.w = rw; a.ew = rw;
        a.f |= H_RW;
        near_note(B, GTF_NEAR_REWEIGHT, a.s, rw, thr);
        if (rw < thr) { a.f &= ~H_ACT; off++; }
    }
    if (eb) {
        double rw = (b.w * b.lik * b.prior) / denom;
        if (norm != 1) rw = rw / (double)norm;
        b.lrn = norm; b.side = lb ? 1 : 2; b.w = rw; b.ew = rw;
        b.f |= H_RW;
        near_note(B, GTF_NEAR_REWEIGHT, b.s, rw, thr);
        if (rw < thr) { b.f &= ~H_ACT; off++; }
    }
    if (off) atomicAdd(&cnt[CNT_RWOFF], off);
}

#ifndef GTF_NODE2_THREADS
#define GTF_NODE2_THREADS 256
#endif
#ifndef GTF_NODE2_MINB
#define GTF_NODE2_MINB 4
#endif
__global__ void __launch_bounds__(GTF_NODE2_THREADS, GTF_NODE2_MINB) k_node2(DevBatch B, DevPack Kin, Prog P)
{
    if (Kin.counts[PK_STOP]) return;
    DevPack K = Kin;
    K.all_exist = Kin.counts[PK_MISSING] == 0; // every slot is an existing edge (counted when the bitmaps were packed)
    __shared__ unsigned int s_cnt[GTF_NCOUNTERS];
    __shared__ int s_n[HV_BINS + 1], s_base[HV_BINS + 1];
    __shared__ int s_list[HV_BINS + 1][GTF_NODE2_THREADS];
    const int tid = threadIdx.x;
    if (tid < GTF_NCOUNTERS) s_cnt[tid] = 0;
    if (tid <= HV_BINS) s_n[tid] = 0;
    __syncthreads();
    const int i = blockIdx.x * GTF_NODE2_THREADS + tid;
    unsigned n_act = 0, n_chg = 0;
    const bool force = K.counts[PK_FORCE] != 0;
    if (i < B.N && !force && K.node_static[i]) {
        // found static by an earlier iteration: activation only ever falls, so it stays static until the next forced pass --
        // no scan of its windows, only the restored merged_cov[1,1] (NaN: nothing to restore)
        const double cp = K.node_rest[i];
        if (cp == cp) B.m_p11_nx[i] = cp;
    } else if (i < B.N) {
        unsigned nf = B.node_ok[i];
        const int b0 = B.in_off[i], b1 = B.in_off[i + 1];
        int np = 0, e0 = -1, e1 = -1, deg = 0, chg = 0;
        unsigned a0_any = 0, f0 = 0, f1 = 0;          // f0 / f1: flag bits of the first two entries, cut out of the windows
        for (int c = b0; c < b1; c += 32) {
            const int nb = min(32, b1 - c);
            const unsigned mask = nb == 32 ? 0xffffffffu : ((1u << nb) - 1u);
            const unsigned ex = K.all_exist ? mask : (bm_win(K.exists, c) & mask);
            const unsigned pr = bm_win(K.pres, c) & mask;
            const unsigned p0 = bm_win(K.pres0, c);
            const unsigned anr = bm_win(K.act_nx, c), a0r = bm_win(K.act, c);
            const unsigned an = anr & ex, a0 = a0r & ex;
            deg += __popc(an);
            chg += __popc(an ^ a0);
            a0_any |= a0;
            if (pr) {
                if (np == 0) {
                    const int k0 = __ffs(pr) - 1;
                    e0 = c + k0; f0 = win_flags(ex, anr, a0r, p0, k0);
                    const unsigned r = pr & (pr - 1);
                    if (r) { const int k1 = __ffs(r) - 1; e1 = c + k1; f1 = win_flags(ex, anr, a0r, p0, k1); }
                } else if (np == 1) {
                    const int k1 = __ffs(pr) - 1;
                    e1 = c + k1; f1 = win_flags(ex, anr, a0r, p0, k1);
                }
                np += __popc(pr);
            }
        }
        // A node without an active in-edge at the start of the iteration receives no message and changes no flag: its
        // program would reproduce its last evaluation bit for bit (same entries, same priors, same cluster).  It is
        // skipped -- except that a cluster's merged_cov[1,1] is re-created by every evaluation while the node keeps
        // sending (quirk 2), so that value is restored.  After anything changed the packed state from outside, every
        // node is evaluated once (PK_FORCE).
        const bool is_static = a0_any == 0 && !force;
        if (force) K.node_static[i] = 0;              // (a forced pass evaluates every node and forgets the marks)
        if (is_static) {
            double cp = NAN;
            if (np > 2 && (nf & NF_OK)) {
                cp = K.mrec[i].cl_p11;
                if (cp == cp) B.m_p11_nx[i] = cp;
            }
            K.node_static[i] = 1;
            K.node_rest[i] = cp;
        } else if (np > 2 && (nf & NF_OK)) {
            const int bin = np <= 4 ? 0 : np <= 8 ? 1 : np <= 16 ? 2 : np <= 32 ? 3 : 4;
            GTF_BOUND(B, np <= b1 - b0);
            s_list[bin][atomicAdd(&s_n[bin], 1)] = i;
        } else {
            if (nf & NF_OK) {
                const int n = np;
                LEnt a, b;
                a.f = 0; b.f = 0; a.s = b.s = b0; a.lay = b.lay = -1; a.sx = b.sx = 0; a.w = b.w = a.lik = b.lik = 0;
                a.prior = b.prior = a.ew = b.ew = 0; a.side = b.side = a.rank = b.rank = 0; a.lrn = b.lrn = -1;
                a.rank0 = b.rank0 = a.tag0 = b.tag0 = 0;
                GTF_BOUND(B, n < 1 || (e0 >= b0 && e0 < b1));
                GTF_BOUND(B, n < 2 || (e1 > e0 && e1 < b1));
                if (n >= 1) lent_load(B, K, e0, a, f0);
                if (n == 2) lent_load(B, K, e1, b, f1);
                if (P.key == GTF_KEY_TSE) nf |= NF_DICT;          // every seeded node holds the dict (clustering.py:198)
                else if (B.has_uts[i]) nf |= NF_HASUTS | NF_DICT;
                const int nnew = ((a.f & H_NEW) != 0) + ((b.f & H_NEW) != 0);
                if (nnew) { // new entries enter the dict in ascending source order (extrapolate...py:419-447)
                    const int nxt = B.uts_next[i];
                    if (nnew == 2) {
                        const bool a_first = (*geo_p(K, e0)).src < (*geo_p(K, e1)).src;
                        a.rank = nxt + (a_first ? 0 : 1);
                        b.rank = nxt + (a_first ? 1 : 0);
                    } else if (a.f & H_NEW) a.rank = nxt; else b.rank = nxt;
                    B.uts_next[i] = nxt + nnew;
                    B.has_uts[i] = 1;
                    nf |= NF_DICT | NF_HASUTS;
                }
                if (n == 2 && b.rank < a.rank) { LEnt t = a; a = b; b = t; } // dict order
                const bool rdict = (nf & (NF_MULTI | NF_DICT)) == (NF_MULTI | NF_DICT);
                const bool ruts = (nf & (NF_MULTI | NF_HASUTS)) == (NF_MULTI | NF_HASUTS);
                if (n && P.pre_passes) {
                    const double nodex = B.x[i];
                    if (rdict) lent_prior(a, b, n);
                    if (ruts) lent_reweight(B, s_cnt, a, b, n, nodex, P.rw_thr);
                    if (rdict) lent_prior(a, b, n);
                    if (ruts) lent_reweight(B, s_cnt, a, b, n, nodex, P.rw_thr);
                }
                if (rdict) {
                    if (n == 0) atomicOr(&s_cnt[CNT_REFERR], (unsigned)GTF_REF_ZERO_DIV);
                    else {
                        const double mw = n == 2 ? 0.5 : 1.0; // helper.py:90
                        a.w = mw;
                        b.w = mw;
                        lent_prior(a, b, n);
                    }
                }
                if (n >= 1) {
                    if ((a.f & (H_EX | H_ACT0)) == (H_EX | H_ACT0) && !(a.f & H_ACT)) { deg--; chg += (a.f & H_ORIG) ? 1 : -1; }
                    lent_store(B, K, a);
                }
                if (n == 2) {
                    if ((b.f & (H_EX | H_ACT0)) == (H_EX | H_ACT0) && !(b.f & H_ACT)) { deg--; chg += (b.f & H_ORIG) ? 1 : -1; }
                    lent_store(B, K, b);
                }
                B.degree[i] = deg;
            }
            if (nf & NF_OK) {          // (edges of fragment sub-graphs left the list with their graph: not counted)
                n_act = deg;
                n_chg = chg;
            }
        }
    }
    n_act = __reduce_add_sync(0xffffffffu, n_act);
    n_chg = __reduce_add_sync(0xffffffffu, n_chg);
    if ((tid & 31) == 0) {
        if (n_act) atomicAdd(&s_cnt[CNT_ACTIVE], n_act);
        if (n_chg) atomicAdd(&s_cnt[CNT_CHANGED], n_chg);
    }
    __syncthreads();
    if (tid <= HV_BINS && s_n[tid]) s_base[tid] = atomicAdd(&K.counts[PK_HV0 + tid], s_n[tid]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k <= HV_BINS; k++)
        if (tid < s_n[k]) K.hv_list[(size_t)k * B.N + s_base[k] + tid] = s_list[k][tid];
    flush_counters(s_cnt, B.counters, tid);
}

// ------------------------------------------------------------------------------------------------ k_hv<G>
struct MergedOut {
    uint8_t *hm;      // has_merged flags to set
    MergedRec *rec;   // merged-state records to write (in place when the pass is committed, shadow otherwise)
    double *p11;      // accumulated merged_cov[1,1] of the next state
    double2 *ab;      // compact (a, b) of the committed state (k_send), NULL in uncommitted passes
};
__device__ __forceinline__ void merged_store(const MergedOut &MO, int i, const GtfState &m, double mprior)
{
    if (!MO.hm[i]) MO.hm[i] = 1;
    double2 *r = reinterpret_cast<double2 *>(MO.rec + i);
    r[0] = make_double2(m.a, m.b);
    r[1] = make_double2(m.c, m.p00);
    r[2] = make_double2(m.p01, m.p22);
    r[3] = make_double2(mprior, m.p11);
    if (MO.ab) MO.ab[i] = make_double2(m.a, m.b);
    MO.p11[i] = m.p11;
}
// the node was evaluated and formed no cluster: remember that (static nodes are skipped later, see k_node2)
__device__ __forceinline__ void merged_none(const MergedOut &MO, int i)
{
    if (MO.rec[i].cl_p11 == MO.rec[i].cl_p11) MO.rec[i].cl_p11 = NAN;
}

template <int G> __device__ __forceinline__ unsigned grp_min_u32(unsigned v)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <int G> __device__ __forceinline__ unsigned grp_add_u32(unsigned v)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int G> __device__ __forceinline__ unsigned grp_or_u32(unsigned v)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int G> __device__ __forceinline__ unsigned long long grp_min_u64(unsigned long long v)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
        unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t < v ? t : v;
    }
    return v;
}
// group arg-min over lanes with `have`: value and the first group lane holding it (-1: none)
template <int G> __device__ __forceinline__ int grp_argmin(double v, bool have, unsigned gmask, int gbase, double &vmin)
{
    const unsigned long long k = have ? dbl_key(v) : ~0ull;
    const unsigned long long mk = grp_min_u64<G>(k);
    const unsigned m = __ballot_sync(0xffffffffu, have && k == mk) & gmask;
    vmin = key_dbl(mk);
    return m ? __ffs(m) - 1 - gbase : -1;
}
template <int G> __device__ __forceinline__ void grp_shfl_info(const GtfInfo &in, int src, GtfInfo &out)
{
    const unsigned FULL = 0xffffffffu;
    out.s00 = __shfl_sync(FULL, in.s00, src, G); out.s01 = __shfl_sync(FULL, in.s01, src, G);
    out.s11 = __shfl_sync(FULL, in.s11, src, G); out.sq = __shfl_sync(FULL, in.sq, src, G);
    out.v0 = __shfl_sync(FULL, in.v0, src, G); out.v1 = __shfl_sync(FULL, in.v1, src, G);
    out.vc = __shfl_sync(FULL, in.vc, src, G); out.vt = __shfl_sync(FULL, in.vt, src, G);
}

#ifndef GTF_HV_WARPS
#define GTF_HV_WARPS 4
#endif
static_assert(GTF_HV_WARPS * 32 >= GTF_MAXD * (GTF_MAXD - 1) / 2, "one thread per entry of the pair table");
#ifndef GTF_HV_MINB
#define GTF_HV_MINB 5
#endif
// per-warp staging of the 32 dict entries a warp works on, indexed by (group base + DICT POSITION): lanes keep their entry
// in slot order (as loaded) and only this index carries the dict order, so nothing is ever permuted between lanes
struct HvWarp {
    double st[8][32];   // a b c tau p00 p01 p11 p22
    double pg[5][32];   // GtfPairGeo: I T Q A C
    double inf[8][32];  // information form: s00 s01 s11 sq v0 v1 vc vt
    double pr[32];      // prior after the re-weighting (merged_prior sums, clustering.py:234,266)
    double xs[32];      // w * likelihood of the active entries (denominator of the re-weighting, in dict order)
    double park[4][32]; // w, likelihood, prior, edge weight of MY entry while the clustering needs the registers
    int rk[32];         // dict stamps / source indices (to rank the entries)
};

template <int G>
__global__ void __launch_bounds__(GTF_HV_WARPS * 32, GTF_HV_MINB) k_hv(DevBatch B, DevPack Kin, Prog P, GtfGeom g, int bin,
                                                                         MergedOut MO)
{
    DevPack K = Kin;
    K.all_exist = Kin.counts[PK_MISSING] == 0; // every slot is an existing edge (counted when the bitmaps were packed)
    constexpr int MAXN = G == 4 ? 4 : G == 8 ? 8 : 15;           // largest dict that can cluster in this bin
    constexpr int R = (MAXN * (MAXN - 1) / 2 + G - 1) / G;       // pair rounds
    constexpr int NG = 32 / G;                                   // nodes per warp
    const unsigned FULL = 0xffffffffu;
    __shared__ unsigned int s_cnt[GTF_NCOUNTERS];
    __shared__ HvWarp s_warp[GTF_HV_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gl = lane % G, gbase = lane - gl;
    const unsigned gmask = G == 32 ? FULL : (((1u << G) - 1u) << gbase);
    if (tid < GTF_NCOUNTERS) s_cnt[tid] = 0;
    // 1/k and the pair -> row table in shared memory: per-lane indices would serialise in the constant cache
    __shared__ double s_recip[33];
    __shared__ uint8_t s_pair_i[GTF_MAXD * (GTF_MAXD - 1) / 2];
    if (tid < 33) s_recip[tid] = tid ? 1.0 / (double)tid : 0.0;
    if (tid < GTF_MAXD * (GTF_MAXD - 1) / 2) s_pair_i[tid] = c_pair_i[tid];
#define HV_RECIP(k) s_recip[k]
#define HV_PAIR_DECODE(p, i, j) do { (i) = s_pair_i[p]; (j) = (p) - (i) * ((i) - 1) / 2; } while (0)
    __syncthreads();
    const int count = K.counts[PK_HV0 + bin];
    const int32_t *list = K.hv_list + (size_t)bin * B.N;
    HvWarp &W = s_warp[warp];
    unsigned n_act = 0, n_chg = 0, n_off = 0, n_deact = 0, n_merged = 0, referr = 0;
    const int nwarps = gridDim.x * GTF_HV_WARPS;
    for (int base = (blockIdx.x * GTF_HV_WARPS + warp) * NG; base < count; base += nwarps * NG) {
        const int idx = base + lane / G;
        const bool gv = idx < count;
        const int i = gv ? list[idx] : 0;
        int b0 = 0, b1 = 0;
        unsigned nf = 0;
        if (gv) {
            b0 = B.in_off[i]; b1 = B.in_off[i + 1];
            nf = B.node_ok[i];
            if (P.key == GTF_KEY_TSE) nf |= NF_DICT;                 // every seeded node holds the dict (clustering.py:198)
            else if (B.has_uts[i]) nf |= NF_HASUTS | NF_DICT;
        }
        // ---- the node's bits: my entry = the gl-th present slot; totals over the slots that hold no entry
        int slot = -1, n = 0, deg_np = 0, chg_np = 0;
        unsigned fw = 0;                                          // flag bits of my entry, cut out of the windows
        for (int c = b0; c < b1; c += 32) {
            const int nb = min(32, b1 - c);
            const unsigned mask = nb == 32 ? FULL : ((1u << nb) - 1u);
            const unsigned ex = K.all_exist ? mask : (bm_win(K.exists, c) & mask);
            const unsigned pr = bm_win(K.pres, c) & mask;
            const unsigned p0 = bm_win(K.pres0, c);
            const unsigned anr = bm_win(K.act_nx, c), a0r = bm_win(K.act, c);
            const unsigned an = anr & ex, a0 = a0r & ex;
            deg_np += __popc(an & ~pr);
            chg_np += __popc((an ^ a0) & ~pr);
            const int cnt = __popc(pr);
            if (slot < 0 && gl < n + cnt) {
                const int k = (int)__fns(pr, 0, gl - n + 1);
                slot = c + k;
                fw = win_flags(ex, anr, a0r, p0, k);
            }
            n += cnt;
        }
        const bool valid = gv && gl < n;
        GTF_BOUND(B, !gv || (i >= 0 && i < B.N && n <= G && n >= 3));
        GTF_BOUND(B, !valid || (slot >= b0 && slot < b1));
        // ---- weight record, tag, geometry, flags of my entry
        double w = 0.0, lik = 0.0, prior = 0.0, sx = 0.0, ew = 0.0;
        int rank = 0x7fffffff, side = 0, lrn = -1, lay = -1000 - lane, src = 0, rank0 = 0, tag0 = 0;
        unsigned f = 0;
        if (valid) {
            const double2 m0 = __ldcs(reinterpret_cast<const double2 *>(K.meta + slot));
            const double2 m1 = __ldcs(reinterpret_cast<const double2 *>(K.meta + slot) + 1);
            const int2 tg = __ldcs(reinterpret_cast<const int2 *>(tag_p(K, slot)));
            w = m0.x; lik = m0.y; prior = m1.x; ew = m1.y;
            rank = tg.x; rank0 = tg.x; tag0 = tg.y;
            side = (int)(int8_t)(tg.y & 0xff);
            lrn = tg.y >> 16;
            const GeoRec gr = ld_geo(geo_p(K, slot));
            sx = gr.sx + 0.0; lay = gr.lay; src = gr.src;
            f = H_PRES | fw;
            if (prior != prior) { side = 0; lrn = 0; }        // entry just written by the extrapolation: no side / lr_layer_norm yet
        }
        const int nmax = (int)__reduce_max_sync(FULL, (unsigned)n);
        // ---- new entries enter the dict in ascending source order (extrapolate...py:419-447)
        const unsigned newm = __ballot_sync(FULL, (f & H_NEW) != 0);
        if (newm) {
            const int nxt = gv ? B.uts_next[i] : 0;
            __syncwarp();
            W.rk[lane] = src;
            __syncwarp();
            int before = 0;
            for (int t = 0; t < nmax; t++) before += ((newm >> (gbase + t)) & 1u) && W.rk[gbase + t] < src;
            if (f & H_NEW) rank = nxt + before;
            const int nnew = __popc(newm & gmask);
            if (nnew) {
                if (gl == 0) { B.uts_next[i] = nxt + nnew; B.has_uts[i] = 1; }
                nf |= NF_DICT | NF_HASUTS;
            }
        }
        // ---- dict position of my entry (ascending stamp)
        __syncwarp();
        W.rk[lane] = rank;
        __syncwarp();
        int pos = 0;
        for (int t = 0; t < nmax; t++) pos += (t < n) && W.rk[gbase + t] < rank;
        if (!valid) pos = gl;
        GTF_BOUND(B, pos >= 0 && pos < G);
        const int dpos = gbase + pos;
        const unsigned lastm = __ballot_sync(FULL, valid && pos == n - 1) & gmask;      // the lane holding the LAST dict key
        const int lastlane = lastm ? __ffs(lastm) - 1 : gbase;
        // ---- state record of my entry -> shared, with the per-entry parts of the pair chi2 and the information form
        double nodex = 0.0;
        const bool cl_node = G < 32 && gv && (nf & (NF_OK | NF_DICT)) == (NF_OK | NF_DICT) && n >= 3 && n <= GTF_MAXD; // clustering.py:207
        const bool cl_any = G < 32 && __any_sync(FULL, cl_node);
        {
            NodeXYZR X;
            X.x = X.y = X.z = X.r = 0.0;
            if (gv) X = K.xyzr[i];
            nodex = X.x;
            if (cl_any) {
                GtfState mine;
                mine.a = mine.b = mine.c = mine.tau = mine.p00 = mine.p01 = mine.p11 = mine.p22 = 0.0;
                double sz = 0.0, sr = 0.0;
                if (valid && cl_node) {
                    const double2 *stp = reinterpret_cast<const double2 *>(K.state + 8 * (size_t)slot);
                    const double2 v0 = __ldcs(stp + 0), v1 = __ldcs(stp + 1), v2 = __ldcs(stp + 2), v3 = __ldcs(stp + 3);
                    mine.a = v0.x; mine.b = v0.y; mine.c = v1.x; mine.tau = v1.y;
                    mine.p00 = v2.x; mine.p01 = v2.y; mine.p11 = v3.x; mine.p22 = v3.y;
                    if (src >= 0) { const NodeXYZR Sx = K.xyzr[src]; sz = Sx.z; sr = Sx.r; }
                }
                GtfPairGeo pg;
                gtf_pair_geo(sx, sz, sr, X.z, X.r, g, pg);
                GtfInfo mi;
                gtf_to_info(mine, mi);
                W.st[0][dpos] = mine.a; W.st[1][dpos] = mine.b; W.st[2][dpos] = mine.c; W.st[3][dpos] = mine.tau;
                W.st[4][dpos] = mine.p00; W.st[5][dpos] = mine.p01; W.st[6][dpos] = mine.p11; W.st[7][dpos] = mine.p22;
                W.pg[0][dpos] = pg.I; W.pg[1][dpos] = pg.T; W.pg[2][dpos] = pg.Q; W.pg[3][dpos] = pg.A; W.pg[4][dpos] = pg.C;
                W.inf[0][dpos] = mi.s00; W.inf[1][dpos] = mi.s01; W.inf[2][dpos] = mi.s11; W.inf[3][dpos] = mi.sq;
                W.inf[4][dpos] = mi.v0; W.inf[5][dpos] = mi.v1; W.inf[6][dpos] = mi.vc; W.inf[7][dpos] = mi.vt;
            }
        }
        const bool okd = (nf & (NF_OK | NF_MULTI | NF_DICT)) == (NF_OK | NF_MULTI | NF_DICT);
        const bool oku = (nf & (NF_OK | NF_MULTI | NF_HASUTS)) == (NF_OK | NF_MULTI | NF_HASUTS);
        const unsigned m3 = H_PRES | H_EX | H_ACT;
        const unsigned samelay = __match_any_sync(FULL, lay) & gmask;
        const unsigned samex = __match_any_sync(FULL, __double_as_longlong(sx)) & gmask;
        const bool isleft = sx < nodex;
        for (int pass = 0; pass < P.pre_passes; pass++) {                 // (2 in the iteration; 0 / 1 for cluster() on the seeds)
            // helper.py:30-63 compute_prior_probabilities
            {
                const bool el = (f & m3) == m3;
                const unsigned elm = __ballot_sync(FULL, el);
                if (okd && el) prior = HV_RECIP(__popc(samelay & elm));
            }
            // helper.py:99-200 side norm + reweight + prune
            if (__any_sync(FULL, oku)) {
                const bool el = oku && (f & m3) == m3;
                const bool left = el && isleft;
                const unsigned elm = __ballot_sync(FULL, el), leftm = __ballot_sync(FULL, left);
                const unsigned grpm = samex & (left ? leftm : (elm & ~leftm));
                const bool first = el && (__ffs(grpm) - 1 == lane);             // distinct x per side: len(set(coords))
                const int normL = __popc(__ballot_sync(FULL, first && left) & gmask);
                const int normR = __popc(__ballot_sync(FULL, first && !left) & gmask);
                const unsigned lf = __shfl_sync(FULL, f, lastlane);             // stale `neighbour_num`: LAST dict key
                const bool any_el = (elm & gmask) != 0;
                if (any_el && !(lf & H_EX) && gl == 0) referr |= GTF_REF_KEY;
                const bool last_active = (lf & (H_EX | H_ACT)) == (H_EX | H_ACT);
                __syncwarp();
                W.xs[dpos] = el ? w * lik : 0.0;                                // denominator in dict order (helper.py:165-169)
                __syncwarp();
                double denom = 0.0;
                for (int q = 0; q < nmax; q++)
                    if (q < n) denom += W.xs[gbase + q];
                if (el) {
                    const int norm = last_active ? (left ? normL : normR) : 1;
                    double rw = (w * lik * prior) / denom;
                    if (norm != 1) rw = rw / (double)norm;
                    lrn = norm; side = left ? 1 : 2;
                    w = rw; ew = rw;
                    f |= H_RW;
                    near_note(B, GTF_NEAR_REWEIGHT, slot, rw, P.rw_thr);
                    if (rw < P.rw_thr) { f &= ~H_ACT; n_off++; }
                }
            }
        }
        // ---- clustering.py:193-307
        bool clustered = false;
        GtfState merged;
        merged.a = merged.b = merged.c = merged.tau = merged.p00 = merged.p01 = merged.p11 = merged.p22 = 0.0;
        double mprior = 0.0;
        if (cl_any) {
            double thr = P.cl_kl;
            if (P.use_lut && gv) {
                const double ev = B.emp_var[i];
                const int lb = (ev == ev) ? (int)floor(ev / 0.05) : 27;
                thr = P.lut[max(0, min(27, lb))];
            }
            // my entry's weights leave the registers while the clustering runs
            W.park[0][lane] = w; W.park[1][lane] = lik; W.park[2][lane] = prior; W.park[3][lane] = ew;
            W.pr[dpos] = prior;
            __syncwarp();
            const int npairs = cl_node ? n * (n - 1) / 2 : 0;
            const int rmax = ((int)__reduce_max_sync(FULL, (unsigned)npairs) + G - 1) / G;
            double pv[R];
            double lbest = INFINITY;
            bool nz_any = false, nan_any = false;
#pragma unroll
            for (int r = 0; r < R; r++) {
                pv[r] = 0.0;
                if (r < rmax) {
                    const int p = r * G + gl;
                    if (p < npairs) {
                        int pi, pj;
                        HV_PAIR_DECODE(p, pi, pj);
                        const int li = gbase + pi, lj = gbase + pj;
                        GtfState si, sj;
                        si.a = W.st[0][li]; si.b = W.st[1][li]; si.p00 = W.st[4][li]; si.p01 = W.st[5][li]; si.p11 = W.st[6][li];
                        sj.a = W.st[0][lj]; sj.b = W.st[1][lj]; sj.p00 = W.st[4][lj]; sj.p01 = W.st[5][lj]; sj.p11 = W.st[6][lj];
                        si.c = si.tau = si.p22 = sj.c = sj.tau = sj.p22 = 0.0;   // (not part of the pairwise chi2)
                        GtfPairGeo gi, gj;
                        gi.I = W.pg[0][li]; gi.T = W.pg[1][li]; gi.Q = W.pg[2][li]; gi.A = W.pg[3][li]; gi.C = W.pg[4][li];
                        gj.I = W.pg[0][lj]; gj.T = W.pg[1][lj]; gj.Q = W.pg[2][lj]; gj.A = W.pg[3][lj]; gj.C = W.pg[4][lj];
                        const double v = gtf_pair_chi2_pre(si, sj, nodex, gi, gj, g);
                        pv[r] = v;
                        if (v != 0.0) {               // np.nonzero keeps NaN, drops +-0 (clustering.py:119)
                            nz_any = true;
                            if (v != v) nan_any = true; else lbest = fmin(lbest, v);
                        }
                    }
                }
            }
            nz_any = (__ballot_sync(FULL, nz_any) & gmask) != 0;
            nan_any = (__ballot_sync(FULL, nan_any) & gmask) != 0;
            if (cl_node && !nz_any && gl == 0) referr |= GTF_REF_EMPTY_MIN;    // np.min([]) -> ValueError
            const double best = key_dbl(grp_min_u64<G>(dbl_key(lbest)));
            bool go = cl_node && nz_any && !nan_any && best < P.cl_chi2;       // clustering.py:228 (nan < thr is False)
            if (cl_node && nz_any && !nan_any && gl == 0) near_note(B, GTF_NEAR_CLUSTER_CHI2, i, best, P.cl_chi2);
            // np.where(distances == smallest): all tied positions in row-major order (clustering.py:122-123)
            unsigned t1 = 1u << 30, t2 = 1u << 30, nm = 0, gone = 0;
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int p = r * G + gl;
                if (r < rmax && p < npairs && pv[r] == best) {
                    int pi, pj;
                    HV_PAIR_DECODE(p, pi, pj);
                    if (nm == 0) t1 = p; else if (nm == 1) t2 = p;
                    nm++;
                    gone |= (1u << pi) | (1u << pj);
                }
            }
            const unsigned pfirst = grp_min_u32<G>(t1);
            const unsigned p2 = grp_min_u32<G>(t1 == pfirst ? t2 : t1);
            nm = grp_add_u32<G>(nm);
            gone = grp_or_u32<G>(gone);
            int idx0 = 0, idx1 = 0;
            if (go) {
                GTF_BOUND(B, (int)pfirst < n * (n - 1) / 2);
                HV_PAIR_DECODE((int)pfirst, idx0, idx1);       // unique minimum: idx = [row, col]
                if (nm > 1) {                               // ties: idx = [rows..., cols...] -> idx[1] is the SECOND ROW
                    int jj;
                    HV_PAIR_DECODE((int)p2, idx1, jj);
                }
            }
            unsigned rem = go ? (((1u << n) - 1u) & ~gone) : 0u;   // (bits = dict positions)
            if (go && rem == 0) {                           // np.min([]) at :252
                if (gl == 0) referr |= GTF_REF_EMPTY_MIN;
                go = false;
            }
            if (__any_sync(FULL, go)) {
                GtfInfo M;
                {
                    const int l0 = gbase + idx0, l1 = gbase + idx1;  // clustering.py:231-233: Sigma^-1 = S_i + S_j
                    M.s00 = W.inf[0][l0]; M.s01 = W.inf[1][l0]; M.s11 = W.inf[2][l0]; M.sq = W.inf[3][l0];
                    M.v0 = W.inf[4][l0]; M.v1 = W.inf[5][l0]; M.vc = W.inf[6][l0]; M.vt = W.inf[7][l0];
                    M.s00 += W.inf[0][l1]; M.s01 += W.inf[1][l1]; M.s11 += W.inf[2][l1]; M.sq += W.inf[3][l1];
                    M.v0 += W.inf[4][l1]; M.v1 += W.inf[5][l1]; M.vc += W.inf[6][l1]; M.vt += W.inf[7][l1];
                    mprior = W.pr[l0] + W.pr[l1];                    // :234
                }
                gtf_from_info(M, merged);
                bool live = go;
                clustered = go;
                const bool mine_in = valid && cl_node;
                while (__any_sync(FULL, live)) {
                    const bool have = live && mine_in && ((rem >> pos) & 1u);
                    double kl = INFINITY;                            // clustering.py:107-112, both inverses at hand
                    if (have) {
                        const double tr = (W.st[4][dpos] - merged.p00) * (M.s00 - W.inf[0][dpos]) + (W.st[6][dpos] - merged.p11) * (M.s11 - W.inf[2][dpos]) +
                                          (W.st[7][dpos] - merged.p22) * (M.sq - W.inf[3][dpos]);
                        const double d0 = W.st[0][dpos] - merged.a, d1 = W.st[1][dpos] - merged.b, d2 = W.st[3][dpos] - merged.tau;
                        kl = tr + (d0 * d0 * (W.inf[0][dpos] + M.s00) + 2.0 * d0 * d1 * (W.inf[1][dpos] + M.s01) + d1 * d1 * (W.inf[2][dpos] + M.s11) +
                                   d2 * d2 * (W.inf[3][dpos] + M.sq));
                    }
                    const bool nan_kl = (__ballot_sync(FULL, have && kl != kl) & gmask) != 0;  // list.index(nan) -> ValueError
                    // arg-min; list.index: the FIRST occurrence in dict order among equal values
                    const unsigned long long kk = have ? dbl_key(kl) : ~0ull;
                    const unsigned long long mk = grp_min_u64<G>(kk);
                    const unsigned tie = __ballot_sync(FULL, have && kk == mk) & gmask;
                    int bk = __shfl_sync(FULL, pos, tie ? __ffs(tie) - 1 : lane);
                    if (__any_sync(FULL, (tie & (tie - 1)) != 0)) {     // (warp-uniform: the reduction shuffles across groups)
                        const int bmin = (int)grp_min_u32<G>(have && kk == mk ? (unsigned)pos : 99u);
                        if (tie & (tie - 1)) bk = bmin;
                    }
                    if (!tie) bk = -1;
                    const double bv = key_dbl(mk);
                    GTF_BOUND(B, bk < n);
                    const bool absorb = live && !nan_kl && bk >= 0 && bv < thr;                // clustering.py:261
                    if (live && !nan_kl && bk >= 0 && gl == 0) near_note(B, GTF_NEAR_CLUSTER_KL, i, bv, thr);
                    if (live && nan_kl) {
                        if (gl == 0) referr |= GTF_REF_NAN_INDEX;
                        clustered = false;
                        live = false;
                    } else if (absorb) {
                        const int lb = gbase + bk;                   // :263-265 merge_states(entry, merged)
                        M.s00 += W.inf[0][lb]; M.s01 += W.inf[1][lb]; M.s11 += W.inf[2][lb]; M.sq += W.inf[3][lb];
                        M.v0 += W.inf[4][lb]; M.v1 += W.inf[5][lb]; M.vc += W.inf[6][lb]; M.vt += W.inf[7][lb];
                        gtf_from_info(M, merged);
                        mprior = W.pr[lb] + mprior;                  // :266
                        rem &= ~(1u << bk);
                        if (rem == 0) live = false;                  // :283
                    } else
                        live = false;
                }
                // un-absorbed components: their in-edge is deactivated (clustering.py:297-321)
                if (clustered && valid && ((rem >> pos) & 1u) && (f & H_EX)) {
                    f &= ~H_ACT;
                    n_deact++;
                }
            }
            w = W.park[0][lane]; lik = W.park[1][lane]; prior = W.park[2][lane]; ew = W.park[3][lane];
        }
        // ---- degree (helper.py:67-73), mixture weights (helper.py:76-94), priors
        const unsigned actm = __ballot_sync(FULL, (f & (H_PRES | H_EX | H_ACT)) == (H_PRES | H_EX | H_ACT)) & gmask;
        const int deg = deg_np + __popc(actm);
        const unsigned chgm = __ballot_sync(FULL, (f & (H_PRES | H_EX)) == (H_PRES | H_EX) && (((f & H_ACT) != 0) != ((f & H_ORIG) != 0))) & gmask;
        if (gv && gl == 0) {
            if (nf & NF_OK) B.degree[i] = deg;
            n_act += deg;
            n_chg += chg_np + __popc(chgm);
        }
        if (okd && (f & H_PRES)) w = HV_RECIP(n);
        {
            const bool el = (f & m3) == m3;
            const unsigned elm = __ballot_sync(FULL, el);
            if (okd && el) prior = HV_RECIP(__popc(samelay & elm));
        }
        // ---- store
        if (valid) {
            double2 *m = reinterpret_cast<double2 *>(K.meta + slot);
            __stcs(m + 0, make_double2(w, lik));
            __stcs(m + 1, make_double2(prior, ew));              // ew: helper.py:180
            const int tw = tag_pack(side, lrn);
            if (rank != rank0 || tw != tag0) *reinterpret_cast<int2 *>(tag_p(K, slot)) = make_int2(rank, tw);
            if ((f & H_ACT0) && !(f & H_ACT)) bm_clear(K.act_nx, slot);
        }
        if (gv && gl == 0) {
            if (clustered) {
                merged_store(MO, i, merged, mprior);
                n_merged++;
            } else
                merged_none(MO, i);
        }
        __syncwarp();
    }
    n_act = __reduce_add_sync(FULL, n_act); n_chg = __reduce_add_sync(FULL, n_chg);
    n_off = __reduce_add_sync(FULL, n_off); n_deact = __reduce_add_sync(FULL, n_deact);
    n_merged = __reduce_add_sync(FULL, n_merged); referr = __reduce_or_sync(FULL, referr);
    if (lane == 0) {
        if (n_act) atomicAdd(&s_cnt[CNT_ACTIVE], n_act);
        if (n_chg) atomicAdd(&s_cnt[CNT_CHANGED], n_chg);
        if (n_off) atomicAdd(&s_cnt[CNT_RWOFF], n_off);
        if (n_deact) atomicAdd(&s_cnt[CNT_DEACT], n_deact);
        if (n_merged) atomicAdd(&s_cnt[CNT_MERGED], n_merged);
        if (referr) atomicOr(&s_cnt[CNT_REFERR], referr);
    }
    __syncthreads();
    flush_counters(s_cnt, B.counters, tid);
}

#undef HV_RECIP
#undef HV_PAIR_DECODE
// ------------------------------------------------------------------------------------------------ k_big
// dicts with more than 32 entries: one 32-thread CTA per node, the generic shared-memory node program of gtf_tile.cuh
// on a tile that holds just this node; lr_layer_norm lands in a shared array behind the tile.
#define GTF_BIG_SMEM (sizeof(TileSmem) + 16 + 2 * sizeof(double) * GTF_TILE_SLOTS + GTF_TILE_SLOTS)
__global__ void __launch_bounds__(32) k_big(DevBatch B, DevPack Kin, Prog P, GtfGeom g, MergedOut MO)
{
    DevPack K = Kin;
    K.all_exist = Kin.counts[PK_MISSING] == 0; // every slot is an existing edge (counted when the bitmaps were packed)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileSmem &sm = *reinterpret_cast<TileSmem *>(smem_raw);
    double *lrn_s = reinterpret_cast<double *>(smem_raw + ((sizeof(TileSmem) + 15) & ~(size_t)15));
    double *ew_s = lrn_s + GTF_TILE_SLOTS;
    uint8_t *fresh_s = reinterpret_cast<uint8_t *>(ew_s + GTF_TILE_SLOTS);
    __shared__ double mscr[8];
    __shared__ uint8_t hm_s[8];
    const int lane = threadIdx.x;
    const int count = K.counts[PK_BIG];
    const int32_t *list = K.hv_list + (size_t)HV_BINS * B.N;
    for (int idx = blockIdx.x; idx < count; idx += gridDim.x) {
        const int i = list[idx];
        const int gs0 = B.in_off[i], d = B.in_off[i + 1] - gs0;
        if (lane < GTF_NCOUNTERS) sm.cnt[lane] = 0;
        if (lane == 0) {
            sm.nbeg[0] = 0; sm.nbeg[1] = (uint16_t)d;
            unsigned nf = NF_DICT | B.node_ok[i];
            if (P.key == GTF_KEY_UTS && B.has_uts[i]) nf |= NF_HASUTS;
            sm.nflags[0] = (uint8_t)nf;
        }
        for (int ls = lane; ls < d; ls += 32) {
            const int s = gs0 + ls;
            const GeoRec gr = (*geo_p(K, s));
            unsigned f = 0, sd = 0;
            int rk = 0x7fffffff;
            fresh_s[ls] = 0;
            if (K.all_exist || bm_get(K.exists, s)) f |= F_EX;
            if (bm_get(K.act_nx, s)) f |= F_ACT;
            if (bm_get(K.act, s)) f |= F_ORIG;
            if (bm_get(K.pres, s)) {
                f |= F_PRES;
                const MetaRec m = K.meta[s];
                const TagRec t = (*tag_p(K, s));
                const bool fresh = m.prior != m.prior;
                if (fresh) fresh_s[ls] = 1;
                sd = SD_ORIGPRES | (fresh ? 0u : ((unsigned)t.side & 3u));
                const double2 *st = reinterpret_cast<const double2 *>(K.state + 8 * (size_t)s);
                const double2 v0 = st[0], v1 = st[1], v2 = st[2], v3 = st[3];
                sm.st[0][ls] = v0.x; sm.st[1][ls] = v0.y; sm.st[2][ls] = v1.x; sm.st[3][ls] = v1.y;
                sm.st[4][ls] = v2.x; sm.st[5][ls] = v2.y; sm.st[6][ls] = v3.x; sm.st[7][ls] = v3.y;
                sm.prior[ls] = m.prior; sm.w[ls] = m.w; sm.lik[ls] = m.lik;
                rk = t.rank;
                if (!bm_get(K.pres0, s)) f |= F_NEW;
                ew_s[ls] = m.ew;
            }
            sm.src[ls] = gr.src;
            sm.srcx[ls] = gr.sx;
            sm.layer[ls] = gr.lay;
            sm.rank[ls] = rk;
            sm.side[ls] = (uint8_t)sd;
            sm.flags[ls] = (uint8_t)f;
            lrn_s[ls] = -1.0;
        }
        __syncwarp();
        // the generic program writes mo[k][i]: point every mo[k] at a shared scratch (biased by -i), copy out afterwards
        double *mo[8];
        for (int k = 0; k < 8; k++) mo[k] = mscr + k - i;
        node_program_generic(sm, B, P, g, i, 0, gs0, 0, lane, P.key == GTF_KEY_UTS, hm_s - i, mo, lrn_s, ew_s);
        __syncwarp();
        if (lane == 0) {
            if (sm.nflags[0] & NF_CLUSTERED) {
                GtfState m;
                m.a = mscr[0]; m.b = mscr[1]; m.c = mscr[2]; m.p00 = mscr[3]; m.p01 = mscr[4]; m.p11 = mscr[5]; m.p22 = mscr[6];
                m.tau = 0.0;
                merged_store(MO, i, m, mscr[7]);
            } else
                merged_none(MO, i);
        }
        unsigned n_act = 0, n_chg = 0;
        for (int ls = lane; ls < d; ls += 32) {
            const int s = gs0 + ls;
            const unsigned f = sm.flags[ls];
            const bool a = f & F_ACT, a0 = f & F_ORIG;
            if (f & F_EX) { n_act += a; n_chg += a != a0; }
            if (!a && bm_get(K.act_nx, s)) bm_clear(K.act_nx, s);
            if (f & F_PRES) {
                MetaRec m;
                m.prior = sm.prior[ls]; m.w = sm.w[ls]; m.lik = sm.lik[ls]; m.ew = ew_s[ls];
                K.meta[s] = m;
                TagRec t = (*tag_p(K, s));
                if (fresh_s[ls]) { t.side = 0; t.lrn = 0; }
                t.rank = sm.rank[ls];
                if (f & F_RW) t.side = (int8_t)(sm.side[ls] & 3);
                if (lrn_s[ls] >= 0.0) t.lrn = (int16_t)lrn_s[ls];
                (*tag_p(K, s)) = t;
            }
        }
        n_act = __reduce_add_sync(0xffffffffu, n_act);
        n_chg = __reduce_add_sync(0xffffffffu, n_chg);
        __syncwarp();
        if (lane == 0) {
            if (n_act) atomicAdd(&B.counters[CNT_ACTIVE], (unsigned long long)n_act);
            if (n_chg) atomicAdd(&B.counters[CNT_CHANGED], (unsigned long long)n_chg);
        }
        flush_counters(sm.cnt, B.counters, lane);
        __syncwarp();
    }
}
