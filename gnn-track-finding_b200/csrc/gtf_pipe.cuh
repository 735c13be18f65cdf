// gtf_pipe.cuh -- "pipeline" form of the fused iteration: the same arithmetic and node programs as k_tile, split
// into small single-purpose kernels so that each runs at its own best occupancy (the fused tile kernel is pinned
// to 16 warps/SM by the 128 registers of the extrapolation and a 97 KB tile):
//   k_prefix      per-source multiple-scattering prefix                        (gtf_tile.cuh)
//   k_msg_list    thread per in-slot : which edges carry a message -> dense global list; active_nx := active
//   k_msg_exec    thread per message : extrapolate, chi2 gate, Kalman update   (extrapolate_merged_states.py:26-402)
//   k_node        thread per node    : scan its slot flags; <= 2 dict entries -> closed-form priors / reweight /
//                                      prune right here; >= 3 -> cooperative lists; merged state carried forward
//   k_heavy       warp per node      : cooperative nodes with <= 32 in-slots   (gtf_tile.cuh)
//   k_bignode     CTA per node       : cooperative nodes with more in-slots (generic shared-memory program)
// Included at the end of gtf_tile.cuh.
#pragma once

#define GTF_PIPE_THREADS 256

__global__ void __launch_bounds__(GTF_PIPE_THREADS) k_msg_list(DevBatch B)
{
    __shared__ int s_n, s_base;
    __shared__ int s_list[GTF_PIPE_THREADS];
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) s_n = 0;
    __syncthreads();
    const int s = blockIdx.x * GTF_PIPE_THREADS + tid;
    bool send = false;
    if (s < B.E) {
        const int src = B.in_src[s], dst = B.slot_dst[s];
        const uint8_t a = B.active[s];
        B.active_nx[s] = a;
        if (a == 1 && src >= 0 && (B.all_alive || (B.alive[src] && B.alive[dst])))
            send = (B.node_ok[dst] & (NF_OK | NF_MULTI)) == (NF_OK | NF_MULTI) && B.has_merged[src] != 0;
    }
    unsigned m = __ballot_sync(0xffffffffu, send);
    if (m) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_n, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (send) s_list[base + __popc(m & ((1u << lane) - 1u))] = s;
    }
    __syncthreads();
    if (s_n) {
        if (tid == 0) s_base = atomicAdd(B.msg_count, s_n); // one global atomic per block
        __syncthreads();
        if (tid < s_n) B.msg_list[s_base + tid] = s_list[tid];
    }
}

__global__ void __launch_bounds__(GTF_PIPE_THREADS, 2) k_msg_exec(DevBatch B, double chi2_cut, GtfGeom g)
{
    __shared__ unsigned int s_cnt[GTF_NCOUNTERS];
    const int tid = threadIdx.x;
    if (tid < GTF_NCOUNTERS) s_cnt[tid] = 0;
    __syncthreads();
    const int count = *B.msg_count;
    unsigned gated = 0, sent = 0;
    for (int q = blockIdx.x * GTF_PIPE_THREADS + tid; q < count; q += gridDim.x * GTF_PIPE_THREADS) {
        const int s = B.msg_list[q];
        const int u = B.in_src[s], v = B.slot_dst[s];
        GtfExtrapOut o;
        gtf_extrapolate(B.x[u], B.y[u], B.z[u], B.r[u], B.x[v], B.y[v], B.z[v], B.r[v], B.m_a[u], B.m_b[u], B.m_c[u],
                        B.m_p00[u], B.m_p01[u], B.slot_p11[s], B.m_p22[u], B.slot_vms[s], chi2_cut, g, o);
        B.uts_chi2[s] = o.chi2;
        sent++;
        if (o.pass) {
            int rs = B.rev_slot[s];
            double wv = NAN;
            if (rs >= 0 && B.tse_present[rs]) wv = B.tse_w[rs]; // extrapolate_merged_states.py:384
            else atomicOr(&s_cnt[CNT_REFERR], (unsigned)GTF_REF_NO_TSE);
            B.uts_a[s] = o.s.a; B.uts_b[s] = o.s.b; B.uts_c[s] = o.s.c; B.uts_tau[s] = o.s.tau;
            B.uts_p00[s] = o.s.p00; B.uts_p01[s] = o.s.p01; B.uts_p11[s] = o.s.p11; B.uts_p22[s] = o.s.p22;
            B.uts_lik[s] = o.lik;
            B.uts_w[s] = wv;
            B.uts_prior[s] = NAN; // a fresh dict entry has no prior / lr_layer_norm / side yet
            B.uts_lrn[s] = NAN;
            B.uts_side[s] = 0;
            if (!B.uts_present[s]) { B.uts_present[s] = 1; B.uts_rank[s] = GTF_NEWMARK; }
        } else {
            B.active_nx[s] = 0; // :393
            gated++;
        }
    }
    if (sent) atomicAdd(&s_cnt[CNT_SENT], sent);
    if (gated) atomicAdd(&s_cnt[CNT_GATED], gated);
    __syncthreads();
    if (tid < GTF_NCOUNTERS && s_cnt[tid]) {
        if (tid == CNT_REFERR) atomicOr(&B.counters[tid], (unsigned long long)s_cnt[tid]);
        else atomicAdd(&B.counters[tid], (unsigned long long)s_cnt[tid]);
    }
}

// one dict entry of a light node, read straight from global memory (slot s)
__device__ __forceinline__ void pipe_light_load(const DevBatch &B, int s, int node, LightEntry &e, int &rank)
{
    const int src = B.in_src[s];
    unsigned f = F_PRES;
    if (src >= 0 && (B.all_alive || (B.alive[src] && B.alive[node]))) f |= F_EX;
    if (B.active_nx[s] == 1) f |= F_ACT;
    rank = B.uts_rank[s];
    if (rank == GTF_NEWMARK) f |= F_NEW;
    e.ls = s;
    e.f = f;
    e.lay = src >= 0 ? B.layer[src] : -1;
    e.sx = (src >= 0 ? B.x[src] : 0.0) + 0.0;
    e.w = B.uts_w[s]; e.lik = B.uts_lik[s]; e.prior = B.uts_prior[s];
    e.side = 0;
}
__device__ __forceinline__ void pipe_light_store(const DevBatch &B, const LightEntry &e, int rank, bool was_active)
{
    const int s = e.ls;
    const bool act = (e.f & F_ACT) != 0;
    if (act != was_active) B.active_nx[s] = act ? 1 : 0;
    if (e.f & F_NEW) B.uts_rank[s] = rank;
    B.uts_w[s] = e.w;
    B.uts_prior[s] = e.prior;
    if (e.f & F_RW) B.uts_side[s] = (int8_t)e.side;
}

#define GTF_NODE_THREADS 128
__global__ void __launch_bounds__(GTF_NODE_THREADS) k_node(DevBatch B, Prog P)
{
    __shared__ unsigned int s_cnt[GTF_NCOUNTERS];
    __shared__ int s_nh, s_nb, s_hbase, s_bbase;
    __shared__ int s_heavy[GTF_NODE_THREADS], s_hslot[GTF_NODE_THREADS], s_big[GTF_NODE_THREADS];
    const int tid = threadIdx.x;
    if (tid < GTF_NCOUNTERS) s_cnt[tid] = 0;
    if (tid == 0) { s_nh = 0; s_nb = 0; }
    __syncthreads();
    const int i = blockIdx.x * GTF_NODE_THREADS + tid;
    unsigned n_act = 0, n_chg = 0;
    if (i < B.N) {
        unsigned nf = B.node_ok[i];
        const int b0 = B.in_off[i], b1 = B.in_off[i + 1];
        if (B.has_uts[i]) nf |= NF_HASUTS | NF_DICT;
        // one scan of the node's slots: dict entries, active in-degree, activation changes so far (the gate)
        int e0 = -1, e1 = -1, np = 0, deg = 0, chg = 0;
        for (int t = b0; t < b1; t++) {
            const bool a = B.active_nx[t] == 1, a0 = B.active[t] == 1;
            bool ex = true;
            if (!B.all_alive) { int src = B.in_src[t]; ex = src >= 0 && B.alive[src] && B.alive[i]; }
            deg += ex && a;
            chg += ex && (a != a0);
            if (B.uts_present[t]) {
                if (np == 0) e0 = t; else if (np == 1) e1 = t;
                np++;
            }
        }
        // merged state carried to the next buffers (cooperative kernels overwrite it when they form a cluster),
        // with the multiple-scattering term the reference accumulates on the node attribute (quirk 2)
        {
            uint8_t h = B.has_merged[i];
            B.has_merged_nx[i] = h;
            if (h) {
                B.m_a_nx[i] = B.m_a[i]; B.m_b_nx[i] = B.m_b[i]; B.m_c_nx[i] = B.m_c[i]; B.m_p00_nx[i] = B.m_p00[i];
                B.m_p01_nx[i] = B.m_p01[i]; B.m_p22_nx[i] = B.m_p22[i]; B.m_prior_nx[i] = B.m_prior[i];
                B.m_p11_nx[i] = B.node_p11tot[i];
            }
        }
        if (np > 2) {
            if (b1 - b0 <= 32) { int k = atomicAdd(&s_nh, 1); s_heavy[k] = i; s_hslot[k] = (b0 << 6) | (b1 - b0); }
            else s_big[atomicAdd(&s_nb, 1)] = i;
        } else {
            n_act = deg;
            n_chg = chg;
            if (nf & NF_OK) {
                const int n = np;
                LightEntry a, b;
                a.f = 0; b.f = 0; a.ls = b.ls = b0; a.lay = b.lay = -1; a.sx = b.sx = 0; a.w = b.w = a.lik = b.lik = 0;
                a.prior = b.prior = 0; a.side = b.side = 0;
                int ra = 0, rb = 0;
                if (n >= 1) pipe_light_load(B, e0, i, a, ra);
                if (n == 2) pipe_light_load(B, e1, i, b, rb);
                const unsigned fa0 = a.f, fb0 = b.f;
                int nnew = ((a.f & F_NEW) != 0) + ((b.f & F_NEW) != 0);
                if (nnew) { // new entries enter the dict in ascending source order (extrapolate...py:419-447)
                    int nxt = B.uts_next[i];
                    if (nnew == 2) {
                        bool a_first = B.in_src[e0] < B.in_src[e1];
                        ra = nxt + (a_first ? 0 : 1);
                        rb = nxt + (a_first ? 1 : 0);
                    } else if (a.f & F_NEW) ra = nxt; else rb = nxt;
                    B.uts_next[i] = nxt + nnew;
                    B.has_uts[i] = 1;
                    nf |= NF_DICT | NF_HASUTS;
                }
                bool swapped = false;
                if (n == 2 && rb < ra) { LightEntry t = a; a = b; b = t; int r = ra; ra = rb; rb = r; swapped = true; }
                const bool rdict = (nf & (NF_MULTI | NF_DICT)) == (NF_MULTI | NF_DICT);
                const bool ruts = (nf & (NF_MULTI | NF_HASUTS)) == (NF_MULTI | NF_HASUTS);
                const double nodex = B.x[i];
                if (n) {
                    if (rdict) light_prior(a, b, n);
                    if (ruts) light_reweight(s_cnt, a, b, n, nodex, P.rw_thr, B.edge_w, B.uts_lrn);
                    if (rdict) light_prior(a, b, n);
                    if (ruts) light_reweight(s_cnt, a, b, n, nodex, P.rw_thr, B.edge_w, B.uts_lrn);
                }
                if (rdict) {
                    if (n == 0) atomicOr(&s_cnt[CNT_REFERR], (unsigned)GTF_REF_ZERO_DIV);
                    else {
                        double mw = n == 2 ? 0.5 : 1.0;
                        a.w = mw;
                        b.w = mw;
                        light_prior(a, b, n);
                    }
                }
                const unsigned m2 = F_EX | F_ACT;
                const unsigned fa_before = swapped ? fb0 : fa0, fb_before = swapped ? fa0 : fb0;
                int lost = 0;
                if (n >= 1) {
                    bool was = (fa_before & m2) == m2, now = (a.f & m2) == m2;
                    lost += was && !now;
                    pipe_light_store(B, a, ra, (fa_before & F_ACT) != 0);
                    if ((fa_before & F_EX) && was != now) { // fold this entry's change into the node's change count
                        bool a0 = B.active[a.ls] == 1;
                        n_chg += ((now != a0) ? 1 : 0) - ((was != a0) ? 1 : 0);
                    }
                }
                if (n == 2) {
                    bool was = (fb_before & m2) == m2, now = (b.f & m2) == m2;
                    lost += was && !now;
                    pipe_light_store(B, b, rb, (fb_before & F_ACT) != 0);
                    if ((fb_before & F_EX) && was != now) {
                        bool a0 = B.active[b.ls] == 1;
                        n_chg += ((now != a0) ? 1 : 0) - ((was != a0) ? 1 : 0);
                    }
                }
                n_act -= lost;
                B.degree[i] = deg - lost;
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
    if (tid == 0) {
        if (s_nh) s_hbase = atomicAdd(B.heavy_count, s_nh);
        if (s_nb) s_bbase = atomicAdd(B.big_count, s_nb);
    }
    __syncthreads();
    if (tid < s_nh) { B.heavy_list[s_hbase + tid] = s_heavy[tid]; B.heavy_slot[s_hbase + tid] = s_hslot[tid]; }
    if (tid < s_nb) B.big_list[s_bbase + tid] = s_big[tid];
    if (tid < GTF_NCOUNTERS && s_cnt[tid]) {
        if (tid == CNT_REFERR) atomicOr(&B.counters[tid], (unsigned long long)s_cnt[tid]);
        else atomicAdd(&B.counters[tid], (unsigned long long)s_cnt[tid]);
    }
}

// cooperative nodes with more than 32 in-slots: one 32-thread CTA per node, generic shared-memory node program
__global__ void __launch_bounds__(32) k_bignode(DevBatch B, Prog P, GtfGeom g)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileSmem &sm = *reinterpret_cast<TileSmem *>(smem_raw);
    const int lane = threadIdx.x;
    const int count = *B.big_count;
    uint8_t *hm_out = B.has_merged_nx;
    double *const mo[8] = {B.m_a_nx, B.m_b_nx, B.m_c_nx, B.m_p00_nx, B.m_p01_nx, B.m_p11_nx, B.m_p22_nx, B.m_prior_nx};
    for (int idx = blockIdx.x; idx < count; idx += gridDim.x) {
        const int i = B.big_list[idx];
        const int gs0 = B.in_off[i], d = B.in_off[i + 1] - gs0;
        if (lane < GTF_NCOUNTERS) sm.cnt[lane] = 0;
        if (lane == 0) {
            sm.nbeg[0] = 0; sm.nbeg[1] = (uint16_t)d;
            unsigned nf = NF_DICT | B.node_ok[i];
            if (B.has_uts[i]) nf |= NF_HASUTS;
            sm.nflags[0] = (uint8_t)nf;
        }
        for (int ls = lane; ls < d; ls += 32) {
            const int s = gs0 + ls, src = B.in_src[s];
            unsigned f = 0, sd = 0;
            int rk = 0x7fffffff;
            if (src >= 0 && (B.all_alive || (B.alive[src] && B.alive[i]))) f |= F_EX;
            if (B.active_nx[s] == 1) f |= F_ACT;
            if (B.active[s] == 1) f |= F_ORIG;
            if (B.uts_present[s]) {
                f |= F_PRES;
                sd = SD_ORIGPRES | ((unsigned)B.uts_side[s] & 3u);
                sm.st[0][ls] = B.uts_a[s]; sm.st[1][ls] = B.uts_b[s]; sm.st[2][ls] = B.uts_c[s]; sm.st[3][ls] = B.uts_tau[s];
                sm.st[4][ls] = B.uts_p00[s]; sm.st[5][ls] = B.uts_p01[s]; sm.st[6][ls] = B.uts_p11[s]; sm.st[7][ls] = B.uts_p22[s];
                sm.prior[ls] = B.uts_prior[s]; sm.w[ls] = B.uts_w[s]; sm.lik[ls] = B.uts_lik[s];
                rk = B.uts_rank[s];
                if (rk == GTF_NEWMARK) f |= F_NEW;
            }
            sm.src[ls] = src;
            sm.srcx[ls] = src >= 0 ? B.x[src] : 0.0;
            sm.layer[ls] = src >= 0 ? B.layer[src] : -1;
            sm.rank[ls] = rk;
            sm.side[ls] = (uint8_t)sd;
            sm.flags[ls] = (uint8_t)f;
        }
        __syncwarp();
        node_program_generic(sm, B, P, g, i, 0, gs0, 0, lane, true, hm_out, mo, B.uts_lrn + gs0, B.edge_w + gs0);
        __syncwarp();
        unsigned n_act = 0, n_chg = 0;
        for (int ls = lane; ls < d; ls += 32) {
            const int s = gs0 + ls;
            unsigned f = sm.flags[ls];
            bool a = f & F_ACT, a0 = f & F_ORIG;
            if (f & F_EX) { n_act += a; n_chg += a != a0; }
            B.active_nx[s] = a ? 1 : 0;
            if (f & F_PRES) {
                B.uts_prior[s] = sm.prior[ls];
                B.uts_w[s] = sm.w[ls];
                if (f & F_RW) B.uts_side[s] = (int8_t)(sm.side[ls] & 3);
                if (f & F_NEW) B.uts_rank[s] = sm.rank[ls];
            } else if (sm.side[ls] & SD_ORIGPRES)
                B.uts_present[s] = 0;
        }
        n_act = __reduce_add_sync(0xffffffffu, n_act);
        n_chg = __reduce_add_sync(0xffffffffu, n_chg);
        __syncwarp();
        if (lane == 0) {
            if (n_act) atomicAdd(&B.counters[CNT_ACTIVE], (unsigned long long)n_act);
            if (n_chg) atomicAdd(&B.counters[CNT_CHANGED], (unsigned long long)n_chg);
        }
        if (lane < GTF_NCOUNTERS && sm.cnt[lane]) {
            if (lane == CNT_REFERR) atomicOr(&B.counters[lane], (unsigned long long)sm.cnt[lane]);
            else atomicAdd(&B.counters[lane], (unsigned long long)sm.cnt[lane]);
        }
        __syncwarp();
    }
}
