// gtf_b200.cu -- C-ABI (include/gtf.h) of the B200 message-passing path: batch object, stage launchers,
// seeding / component / extraction / tag-propagation kernels.  Built for sm_100a only; no CPU fallback.
#include <cub/device/device_radix_sort.cuh>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "gtf_tile.cuh"
#include "gtf_iter.cuh"
#ifndef GTF_FUSED_SX
#define GTF_FUSED_SX 0     // default of the run-time switch (environment GTF_FUSED_SX, read when a batch is created): send +
#endif                     // execute fused into the warp-specialised k_sx; 0: k_send -> global message list -> k_exec.
                           // Measured: k_sx saves 18 % of the iteration's DRAM traffic and is 7 % slower.
#ifndef GTF_EXEC_WAVES
#define GTF_EXEC_WAVES 1   // persistent grid = exactly the resident CTAs (measured: 0.1335 vs 0.138 ms with two waves)
#endif

// ------------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(GTF_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));           \
    } while (0)

#define TRY_(x)             \
    do {                    \
        int r__ = (x);      \
        if (r__) return r__; \
    } while (0)

struct Prog;
struct GtfGeom;
static int cluster_seeds_packed(gtf_batch *b, Prog P, const GtfGeom &gg, int pre_passes, gtf_stats *st, const gtf_geom *seed_first = nullptr);
static const bool g_tile_cluster = getenv("GTF_TILE_CLUSTER") != nullptr;   // debugging: cluster() on the seeds with the per-stage kernel

struct FieldInfo { const char *name; int elem; char ext; };
static const FieldInfo g_fields[GTF_NFIELDS] = {
#define EXT_N 'N'
#define EXT_N1 'n'
#define EXT_E 'E'
#define EXT_S 'S'
#define EXT_S1 's'
#define X(name, type, ext) {#name, (int)sizeof(type), EXT_##ext},
    GTF_FIELDS(X)
#undef X
};
static int64_t field_count(const gtf_batch *b, int f)
{
    switch (g_fields[f].ext) {
    case 'N': return b->N;
    case 'n': return (int64_t)b->N + 1;
    case 'E': return b->E;
    case 'S': return b->S;
    default: return (int64_t)b->S + 1;
    }
}

extern "C" int gtf_abi_version(void) { return GTF_ABI_VERSION; }
extern "C" const char *gtf_last_error(void) { return g_err.c_str(); }
extern "C" int gtf_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
extern "C" int gtf_field_count(void) { return GTF_NFIELDS; }
extern "C" const char *gtf_field_name(int f) { return (f >= 0 && f < GTF_NFIELDS) ? g_fields[f].name : nullptr; }
extern "C" int gtf_field_id(const char *name)
{
    for (int f = 0; f < GTF_NFIELDS; f++)
        if (!strcmp(name, g_fields[f].name)) return f;
    return -1;
}
extern "C" int64_t gtf_field_bytes(const gtf_batch *b, int f)
{
    if (!b || f < 0 || f >= GTF_NFIELDS) return -1;
    return field_count(b, f) * g_fields[f].elem;
}

// ------------------------------------------------------------------------------------------------ batch
template <typename T> static int dalloc(gtf_batch *b, T **p, int64_t count)
{
    size_t bytes = (size_t)(count > 0 ? count : 1) * sizeof(T) + 64;   // (+64: bulk copies read whole 16 B blocks around a range)
    CK(cudaMalloc((void **)p, bytes));
    CK(cudaMemsetAsync(*p, 0, bytes, b->stream));
    b->dev_bytes += bytes;
    return 0;
}
#define DA(ptr, count)                         \
    do {                                       \
        int r_ = dalloc(b, &(ptr), (count));   \
        if (r_) return r_;                     \
    } while (0)

static void sync_dev_view(gtf_batch *b)
{
    DevBatch &d = b->d;
    d.N = b->N; d.E = b->E; d.S = b->S; d.n_tiles = b->n_tiles;
    int f = 0;
#define X(name, type, ext) d.name = (type *)b->f[f++];
    GTF_FIELDS(X)
#undef X
    d.tile_begin = b->tile_begin;
}

static int batch_alloc(gtf_batch *b);
extern "C" int gtf_batch_create(int32_t N, int32_t E, int32_t S, int device, gtf_batch **out)
{
    if (!out || N < 0 || E < 0 || S < 0) return fail(GTF_E_ARG, "gtf_batch_create: bad sizes");
    if (gtf_device_count() <= device || device < 0) return fail(GTF_E_CUDA, "gtf_batch_create: no such CUDA device");
    CK(cudaSetDevice(device));
    gtf_batch *b = new gtf_batch();
    memset((void *)b, 0, sizeof(*b));
    b->N = N; b->E = E; b->S = S; b->device = device;
    b->capN = N; b->capE = E; b->capS = S;
    if (const char *fg = getenv("GTF_L2_FETCH")) {   // experiment switch: L2 -> DRAM fetch granularity (32 / 64 / 128 B)
        if (cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(fg)) != cudaSuccess) cudaGetLastError();
    }
    int rc = batch_alloc(b);
    if (rc) {                      // every pointer is zero-initialised: a partial batch is safe to destroy
        const std::string msg = g_err;
        gtf_batch_destroy(b);
        cudaGetLastError();
        g_err = msg;
        return rc;
    }
    *out = b;
    return 0;
}
static int batch_alloc(gtf_batch *b)
{
    const int N = b->N, E = b->E, S = b->S, device = b->device;
    CK(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&b->stream2, cudaStreamNonBlocking));
    for (int f = 0; f < GTF_NFIELDS; f++) {
        size_t bytes = (size_t)field_count(b, f) * g_fields[f].elem + 64;   // (+64: see dalloc)
        CK(cudaMalloc(&b->f[f], bytes));
        CK(cudaMemsetAsync(b->f[f], 0, bytes, b->stream));
        b->dev_bytes += bytes;
    }
    DevBatch &d = b->d;
    DA(d.sub_nalive, S);
    DA(d.node_ok, N);
    DA(b->n_dead, 1);
    DA(d.slot_p11, E); DA(d.slot_vms, E); DA(d.node_p11tot, N);
    DA(d.has_merged_nx, N); DA(d.m_p11_nx, N);
    DA(d.counters, GTF_NCOUNTERS_ALL);
    DA(d.near_log, GTF_NEAR_LOG);
    {
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        b->n_sm = prop.multiProcessorCount;
    }
    {
        // packed iteration layout (gtf_iter.cuh)
        DevPack &k = b->k;
        const int64_t words = ((int64_t)E + 31) / 32 + 2;
        DA(k.out_dst, E); DA(k.orec, E); DA(k.aux, E); DA(k.xyzr, N); DA(k.mrec, N); DA(k.mrec_nx, N); DA(k.srec, (int64_t)N + 1); DA(k.mab, N);
        DA(k.act, words); DA(k.act_nx, words); DA(k.pres, words); DA(k.exists, words); DA(k.pres0, words);
        DA(k.state, (int64_t)E * 8); DA(k.meta, E);
        DA(k.msg_desc, E); DA(k.msg_w, E);
        DA(k.msg_p11, E); DA(k.msg_vms, E);
        DA(k.hv_list, (int64_t)(HV_BINS + 1) * N);
        DA(k.c_edge, E); DA(k.c_rng, N); DA(k.node_static, N); DA(k.node_rest, N);
        CK(cudaMemset(k.node_static, 0, (size_t)(N ? N : 1)));
        DA(k.counts, PK_NCOUNTS);
        b->pack_static_stale = true;
        b->exists_stale = true;
        for (int q = 0; q < PG_N; q++) { b->pack_stale[q] = true; b->soa_stale[q] = false; }
        CK(cudaStreamCreateWithFlags(&b->stream3, cudaStreamNonBlocking));
        const char *ge = getenv("GTF_GRAPH");
        b->use_graph = !(ge && ge[0] == '0');
        const char *fe = getenv("GTF_FUSED_SX");
        b->fused_sx = fe ? fe[0] == '1' : GTF_FUSED_SX != 0;
        b->force_pending = true;
        b->force_dev = -1;
        CK(cudaEventCreateWithFlags(&b->ev_fork2, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&b->ev_join2, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&b->ev_join3, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&b->ev_tiles, cudaEventDisableTiming));
    }
    DA(b->accepted_total, N); DA(b->cand_root, N); DA(b->sub_has_inactive, S); DA(b->sub_first, S);
    DA(b->sort_keys, N); DA(b->sort_vals, N); DA(b->sort_keys2, N); DA(b->sort_vals2, N);
    DA(b->pv_xy, N); DA(b->pv_zr, N); DA(b->acc_now, N); DA(b->tags_a, N); DA(b->tags_b, N);
    DA(b->tile_begin, (int64_t)N + 2); DA(b->stile_begin, 4 * ((int64_t)N + 2)); // at most one tile per node (+ sentinel); int4 per k_send tile
    CK(cudaMallocHost((void **)&b->h_counters, sizeof(unsigned long long) * GTF_NCOUNTERS_ALL));
    CK(cudaMallocHost((void **)&b->h_loop_stats, sizeof(unsigned long long) * GTF_NCOUNTERS_ALL * GTF_LOOP_BURST));
    CK(cudaMallocHost((void **)&b->h_loop_done, sizeof(int)));
    CK(cudaMalloc((void **)&b->loop_stats, sizeof(unsigned long long) * GTF_NCOUNTERS_ALL * GTF_LOOP_BURST));
    CK(cudaMallocHost((void **)&b->h_tiles, sizeof(int32_t) * 5 * ((size_t)N + 2)));
    CK(cudaFuncSetAttribute(k_send, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SendSmem)));
    CK(cudaFuncSetAttribute(k_sx, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SxSmem)));
    CK(cudaFuncSetAttribute(k_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem)));
    CK(cudaFuncSetAttribute(k_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GTF_BIG_SMEM));
    sync_dev_view(b);
    CK(cudaStreamSynchronize(b->stream));
    return 0;
}

extern "C" int gtf_batch_destroy(gtf_batch *b)
{
    if (!b) return 0;
    cudaSetDevice(b->device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    for (int f = 0; f < GTF_NFIELDS; f++) cudaFree(b->f[f]);
    DevBatch &d = b->d;
    void *extra[] = {d.sub_nalive, d.node_ok, b->n_dead, d.slot_p11, d.slot_vms, d.node_p11tot, d.has_merged_nx, d.m_p11_nx,
                     d.counters, d.near_log, b->accepted_total, b->cand_root, b->sub_has_inactive, b->sub_first, b->sort_keys,
                     b->sort_vals, b->sort_keys2, b->sort_vals2, b->pv_xy, b->pv_zr, b->acc_now, b->tags_a, b->tags_b,
                     b->tile_begin, b->sort_tmp, b->cand_rows};
    for (void *p : extra) cudaFree(p);
    {
        DevPack &k = b->k;
        void *pk[] = {k.mab, k.srec, k.mrec, k.mrec_nx, k.out_dst, k.orec, k.aux, k.xyzr, k.act, k.act_nx, k.pres, k.exists, k.pres0, k.state, k.meta, k.msg_desc,
                      k.msg_w, k.msg_p11, k.msg_vms, k.hv_list, k.c_edge, k.c_rng, k.node_static, k.node_rest, k.counts, b->stile_begin};
        for (void *p : pk) cudaFree(p);
        for (int c = 0; c < 2; c++)
            for (int q = 0; q < 2; q++)
                for (int v = 0; v < 2; v++)
                    if (b->graphs[c][q][v].exec) cudaGraphExecDestroy(b->graphs[c][q][v].exec);
        if (b->stream3) cudaStreamDestroy(b->stream3);
        if (b->ev_fork2) cudaEventDestroy(b->ev_fork2);
        if (b->ev_join2) cudaEventDestroy(b->ev_join2);
        if (b->ev_join3) cudaEventDestroy(b->ev_join3);
        if (b->ev_tiles) cudaEventDestroy(b->ev_tiles);
        for (int q = 0; q < 6; q++) if (b->evk[q]) cudaEventDestroy(b->evk[q]);
    }
    if (b->h_counters) cudaFreeHost(b->h_counters);
    if (b->h_loop_stats) cudaFreeHost(b->h_loop_stats);
    if (b->h_loop_done) cudaFreeHost(b->h_loop_done);
    if (b->loop_stats) cudaFree(b->loop_stats);
    if (b->h_tiles) cudaFreeHost(b->h_tiles);
    if (b->stream) cudaStreamDestroy(b->stream);
    if (b->stream2) cudaStreamDestroy(b->stream2);
    delete b;
    return 0;
}

// ---- SoA fields <-> packed iteration layout: which side holds the newer copy, per group of fields
static int field_group(int f)
{
    switch (f) {
    case GTF_F_active: return PG_ACT;
    case GTF_F_uts_present: return PG_PRES;
    case GTF_F_uts_rank: case GTF_F_uts_a: case GTF_F_uts_b: case GTF_F_uts_c: case GTF_F_uts_tau: case GTF_F_uts_p00:
    case GTF_F_uts_p01: case GTF_F_uts_p11: case GTF_F_uts_p22: case GTF_F_uts_lik: case GTF_F_uts_prior: case GTF_F_uts_w:
    case GTF_F_uts_lrn: case GTF_F_uts_side: case GTF_F_edge_w: return PG_REC;
    case GTF_F_m_a: case GTF_F_m_b: case GTF_F_m_c: case GTF_F_m_p00: case GTF_F_m_p01: case GTF_F_m_p22: case GTF_F_m_prior:
        return PG_NODE;
    default: return -1;
    }
}
static bool field_is_pack_static(int f)
{
    return f == GTF_F_x || f == GTF_F_y || f == GTF_F_z || f == GTF_F_r || f == GTF_F_layer || f == GTF_F_in_src ||
           f == GTF_F_slot_dst || f == GTF_F_out_slot || f == GTF_F_rev_slot || f == GTF_F_in_off || f == GTF_F_out_off ||
           f == GTF_F_tse_w || f == GTF_F_tse_present; // (the carried seed weights live in the per-out-edge records)
}
// bring the SoA arrays of the groups in `mask` up to date (after packed iterations)
static int soa_sync(gtf_batch *b, unsigned mask)
{
    int d[PG_N];
    for (int q = 0; q < PG_N; q++) d[q] = ((mask >> q) & 1u) && b->soa_stale[q];
    if ((d[PG_ACT] || d[PG_PRES] || d[PG_REC]) && b->E) {
        k_unpack_slots<<<(b->E + 255) / 256, 256, 0, b->stream>>>(b->d, b->k, d[PG_ACT], d[PG_PRES], d[PG_REC]);
        CK(cudaGetLastError());
    }
    if (d[PG_NODE] && b->N) {
        k_unpack_nodes<<<(b->N + 255) / 256, 256, 0, b->stream>>>(b->d, b->k);
        CK(cudaGetLastError());
    }
    for (int q = 0; q < PG_N; q++) if ((mask >> q) & 1u) b->soa_stale[q] = false;
    return 0;
}
// a stage that works on the SoA arrays is about to run: SoA current, packed copy invalid afterwards
static int soa_for_stage(gtf_batch *b, bool writes)
{
    int r = soa_sync(b, (1u << PG_N) - 1u);
    if (r) return r;
    if (writes) {
        for (int q = 0; q < PG_N; q++) b->pack_stale[q] = true;
        b->pack_static_stale = true; // seed weights (tse_w) may change: per-out-edge records are rebuilt
        b->exists_stale = true;      // nodes may have been removed
    }
    return 0;
}

extern "C" int gtf_batch_upload(gtf_batch *b, int f, const void *host)
{
    if (!b || f < 0 || f >= GTF_NFIELDS || !host) return fail(GTF_E_ARG, "gtf_batch_upload: bad argument");
    CK(cudaSetDevice(b->device));
    const int grp = field_group(f);
    if (grp == PG_REC || grp == PG_NODE) { int r = soa_sync(b, 1u << grp); if (r) return r; } // the group's other fields must be current
    if (f == GTF_F_alive) { // the existing-edge bitmap is rebuilt from alive (together with the activation bitmap) at the next pack
        int r = soa_sync(b, 1u << PG_ACT); if (r) return r;
        b->pack_stale[PG_ACT] = true; b->exists_stale = true;
    }
    CK(cudaMemcpyAsync(b->f[f], host, (size_t)gtf_field_bytes(b, f), cudaMemcpyHostToDevice, b->stream));
    if (grp >= 0) { b->soa_stale[grp] = false; b->pack_stale[grp] = true; }
    if (field_is_pack_static(f)) b->pack_static_stale = true;
    if (f == GTF_F_alive || f == GTF_F_sub_state || f == GTF_F_sub) b->derived_dirty = true;
    if (f == GTF_F_has_merged || f == GTF_F_uts_next || f == GTF_F_has_uts || f == GTF_F_m_p11) b->force_pending = true;
    return 0;
}
extern "C" int gtf_batch_download_async(gtf_batch *b, int f, void *host)
{
    if (!b || f < 0 || f >= GTF_NFIELDS || !host) return fail(GTF_E_ARG, "gtf_batch_download_async: bad argument");
    CK(cudaSetDevice(b->device));
    const int grp = field_group(f);
    if (grp >= 0) { int r = soa_sync(b, 1u << grp); if (r) return r; }
    CK(cudaMemcpyAsync(host, b->f[f], (size_t)gtf_field_bytes(b, f), cudaMemcpyDeviceToHost, b->stream));
    return 0;
}
extern "C" int gtf_batch_download(gtf_batch *b, int f, void *host)
{
    int r = gtf_batch_download_async(b, f, host);
    if (r) return r;
    CK(cudaStreamSynchronize(b->stream));
    return 0;
}
extern "C" int gtf_batch_device_ptr(gtf_batch *b, int f, void **dptr)
{
    if (!b || f < 0 || f >= GTF_NFIELDS || !dptr) return fail(GTF_E_ARG, "gtf_batch_device_ptr: bad argument");
    const int grp = field_group(f);
    if (grp >= 0) { // the caller may write through the pointer: the SoA array becomes the authority for this group
        CK(cudaSetDevice(b->device));
        int r = soa_sync(b, 1u << grp);
        if (r) return r;
        b->pack_stale[grp] = true;
    }
    if (f == GTF_F_alive) { // (same invalidation as gtf_batch_upload: the caller may write through the pointer)
        CK(cudaSetDevice(b->device));
        int r = soa_sync(b, 1u << PG_ACT); if (r) return r;
        b->pack_stale[PG_ACT] = true; b->exists_stale = true;
    }
    if (field_is_pack_static(f)) b->pack_static_stale = true;
    if (f == GTF_F_alive || f == GTF_F_sub_state || f == GTF_F_sub) b->derived_dirty = true;
    if (f == GTF_F_has_merged || f == GTF_F_uts_next || f == GTF_F_has_uts || f == GTF_F_m_p11) b->force_pending = true;
    *dptr = b->f[f];
    return 0;
}
extern "C" int gtf_batch_near_threshold(gtf_batch *b, gtf_near_rec *out, int cap, int64_t *n)
{
    if (!b || !n) return fail(GTF_E_ARG, "gtf_batch_near_threshold: null argument");
    CK(cudaSetDevice(b->device));
    unsigned long long cnt = 0;
    CK(cudaMemcpyAsync(&cnt, b->d.counters + CNT_NEAR, sizeof(cnt), cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    *n = (int64_t)cnt;
    int k = (int)std::min<unsigned long long>(cnt, (unsigned long long)GTF_NEAR_LOG);
    if (cap < k) k = cap;
    if (out && k > 0) {
        CK(cudaMemcpyAsync(out, b->d.near_log, sizeof(gtf_near_rec) * (size_t)k, cudaMemcpyDeviceToHost, b->stream));
        CK(cudaStreamSynchronize(b->stream));
    }
    return 0;
}
extern "C" int gtf_batch_sync(gtf_batch *b)
{
    if (!b) return fail(GTF_E_ARG, "null batch");
    CK(cudaStreamSynchronize(b->stream));
    return 0;
}
extern "C" int gtf_batch_stream(gtf_batch *b, void **s)
{
    if (!b || !s) return fail(GTF_E_ARG, "null");
    *s = (void *)b->stream;
    return 0;
}
extern "C" int64_t gtf_batch_device_bytes(const gtf_batch *b) { return b ? b->dev_bytes : 0; }
extern "C" int64_t gtf_batch_iteration_launches(const gtf_batch *b) { return b ? b->launches : 0; }

// alive nodes per sub-graph
__global__ void k_sub_count(DevBatch B)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    // consecutive nodes mostly share a sub-graph: one atomic per distinct sub-graph in the warp
    const int sg = (i < B.N && B.alive[i]) ? B.sub[i] : -1 - (int)(threadIdx.x & 31);
    const unsigned same = __match_any_sync(0xffffffffu, sg);
    if (sg >= 0 && (int)(threadIdx.x & 31) == __ffs(same) - 1) atomicAdd(&B.sub_nalive[sg], __popc(same));
}
__global__ void k_node_ok(DevBatch B, unsigned long long *n_dead)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B.N) return;
    int sg = B.sub[i];
    unsigned f = 0;
    if (B.alive[i] && B.sub_state[sg] == GTF_SUB_INPLAY) f |= NF_OK;
    if (B.sub_nalive[sg] != 1) f |= NF_MULTI;
    B.node_ok[i] = (uint8_t)f;
    if (!B.alive[i]) atomicAdd(n_dead, 1ull);
}
// derived per-sub-graph / per-node flags; call after alive / sub_state changed
static int recount_subs(gtf_batch *b)
{
    CK(cudaMemsetAsync(b->d.sub_nalive, 0, sizeof(int32_t) * (b->S ? b->S : 1), b->stream));
    CK(cudaMemsetAsync(b->n_dead, 0, sizeof(unsigned long long), b->stream));
    if (b->N) {
        k_sub_count<<<(b->N + 255) / 256, 256, 0, b->stream>>>(b->d);
        k_node_ok<<<(b->N + 255) / 256, 256, 0, b->stream>>>(b->d, b->n_dead);
    }
    CK(cudaGetLastError());
    unsigned long long nd = 0;
    CK(cudaMemcpyAsync(&nd, b->n_dead, sizeof(unsigned long long), cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    b->d.all_alive = nd == 0;
    b->derived_dirty = false;
    return 0;
}

// node tiles of the per-stage kernel (whole nodes, <= GTF_TILE_NODES nodes, <= GTF_TILE_SLOTS in-slots) and source tiles of
// k_send (whole sources, <= GTF_SEND_SRCS sources, <= GTF_SEND_EDGES out-edges) from the two CSR offset arrays on the host
static int build_tiles(gtf_batch *b, const int32_t *in_off, const int32_t *out_off)
{
    const int N = b->N;
    if (N && (in_off[0] != 0 || in_off[N] != b->E)) return fail(GTF_E_STATE, "in_off does not span the slots");
    if (N && (out_off[0] != 0 || out_off[N] != b->E)) return fail(GTF_E_STATE, "out_off does not span the slots");
    if (b->ev_tiles_used) CK(cudaEventSynchronize(b->ev_tiles));   // the previous upload out of the staging buffer is done
    int32_t *t0 = b->h_tiles, *t1 = b->h_tiles + ((size_t)b->capN + 2);
    int nt = 0, i = 0;
    t0[0] = 0;
    while (i < N) {
        int start = i, slots = 0;
        while (i < N && (i - start) < GTF_TILE_NODES) {
            const int d = in_off[i + 1] - in_off[i];
            if (d < 0) return fail(GTF_E_STATE, "in_off not monotone");
            if (d > GTF_TILE_SLOTS) return fail(GTF_E_DEGREE, "node in-degree exceeds GTF_TILE_SLOTS");
            if (slots + d > GTF_TILE_SLOTS) break;
            slots += d;
            i++;
        }
        t0[++nt] = i;
    }
    int ns = 0, u = 0;
    while (u < N) {     // descriptor per tile: (first source, sources, first out-edge, out-edges)
        int start = u, edges = 0;
        while (u < N && (u - start) < GTF_SEND_SRCS) {
            const int dg = out_off[u + 1] - out_off[u];
            if (dg < 0) return fail(GTF_E_STATE, "out_off not monotone");
            if (dg > GTF_SEND_EDGES) return fail(GTF_E_DEGREE, "node out-degree exceeds GTF_SEND_EDGES");
            if (edges + dg > GTF_SEND_EDGES) break;
            edges += dg;
            u++;
        }
        t1[4 * ns + 0] = start; t1[4 * ns + 1] = u - start; t1[4 * ns + 2] = out_off[start]; t1[4 * ns + 3] = edges;
        ns++;
    }
    b->n_tiles = nt;
    b->n_stiles = ns;
    b->topo_gen++;
    CK(cudaMemcpyAsync(b->tile_begin, t0, sizeof(int32_t) * ((size_t)nt + 1), cudaMemcpyHostToDevice, b->stream));
    if (ns) CK(cudaMemcpyAsync(b->stile_begin, t1, sizeof(int32_t) * 4 * (size_t)ns, cudaMemcpyHostToDevice, b->stream));
    CK(cudaEventRecord(b->ev_tiles, b->stream));
    b->ev_tiles_used = true;
    return 0;
}

extern "C" int gtf_batch_finalize(gtf_batch *b)
{
    if (!b) return fail(GTF_E_ARG, "null batch");
    CK(cudaSetDevice(b->device));
    std::vector<int32_t> in_off((size_t)b->N + 1);
    CK(cudaMemcpyAsync(in_off.data(), b->f[GTF_F_in_off], sizeof(int32_t) * ((size_t)b->N + 1), cudaMemcpyDeviceToHost,
                       b->stream));
    CK(cudaStreamSynchronize(b->stream));
    std::vector<int32_t> out_off((size_t)b->N + 1);
    CK(cudaMemcpyAsync(out_off.data(), b->f[GTF_F_out_off], sizeof(int32_t) * ((size_t)b->N + 1), cudaMemcpyDeviceToHost,
                       b->stream));
    CK(cudaStreamSynchronize(b->stream));
    TRY_(build_tiles(b, in_off.data(), out_off.data()));
    sync_dev_view(b);
    int r = recount_subs(b);
    if (r) return r;
    CK(cudaStreamSynchronize(b->stream));
    b->finalized = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------ device-side ingest
// event_conversion.py:40-112 hands the stages a freshly built graph: every node alive, every sub-graph in play, no state
// dicts yet.  The host supplies only what defines the events (hits + the two CSR orders); everything else is derived or
// initialised here, on the device.
__global__ void k_derive_slot_dst(DevBatch B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B.N) return;
    for (int s = B.in_off[i]; s < B.in_off[i + 1]; s++) B.slot_dst[s] = i;
}
// rev_slot[s] for s = (u -> v): the slot of (v -> u), i.e. the in-slot of u whose source is v (-1: one-directional edge)
__global__ void k_derive_rev_slot(DevBatch B)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= B.E) return;
    const int u = B.in_src[s], v = B.slot_dst[s];
    int rs = -1;
    if (u >= 0)
        for (int t = B.in_off[u]; t < B.in_off[u + 1]; t++)
            if (B.in_src[t] == v) { rs = t; break; }
    B.rev_slot[s] = rs;
}
extern "C" int gtf_batch_load_events(gtf_batch *b, const gtf_events *ev)
{
    if (!b || !ev) return fail(GTF_E_ARG, "gtf_batch_load_events: null argument");
    const int N = ev->n_nodes, E = ev->n_slots, S = ev->n_subgraphs;
    if (N < 0 || E < 0 || S < 0 || N > b->capN || E > b->capE || S > b->capS)
        return fail(GTF_E_ARG, "gtf_batch_load_events: the events exceed the capacity the batch was created with");
    if (!ev->x || !ev->y || !ev->z || !ev->r || !ev->layer || !ev->volume || !ev->sub || !ev->sub_off || !ev->sub_event ||
        !ev->in_off || !ev->in_src || !ev->out_off || !ev->out_slot)
        return fail(GTF_E_ARG, "gtf_batch_load_events: missing array");
    CK(cudaSetDevice(b->device));
    b->N = N; b->E = E; b->S = S;
    b->finalized = false;
    cudaStream_t st = b->stream;
    // 1. the events: one asynchronous copy per array (pinned host memory makes them overlap the host work below)
    struct Up { int f; const void *p; } ups[] = {
        {GTF_F_x, ev->x}, {GTF_F_y, ev->y}, {GTF_F_z, ev->z}, {GTF_F_r, ev->r}, {GTF_F_layer, ev->layer},
        {GTF_F_volume, ev->volume}, {GTF_F_sub, ev->sub}, {GTF_F_sub_off, ev->sub_off}, {GTF_F_sub_event, ev->sub_event},
        {GTF_F_in_off, ev->in_off}, {GTF_F_in_src, ev->in_src}, {GTF_F_out_off, ev->out_off}, {GTF_F_out_slot, ev->out_slot}};
    bool given[GTF_NFIELDS] = {false};
    for (const Up &u : ups) {
        const size_t bytes = (size_t)gtf_field_bytes(b, u.f);
        if (bytes) CK(cudaMemcpyAsync(b->f[u.f], u.p, bytes, cudaMemcpyHostToDevice, st));
        given[u.f] = true;
    }
    // 2. every other array: the value a freshly converted event has (absent = NaN / -1 / 0; alive = 1)
    given[GTF_F_slot_dst] = given[GTF_F_rev_slot] = true;   // derived below
    for (int f = 0; f < GTF_NFIELDS; f++) {
        if (given[f]) continue;
        const size_t bytes = (size_t)gtf_field_bytes(b, f);
        if (!bytes) continue;
        int v = 0;
        if (g_fields[f].elem == 8 || f == GTF_F_uts_rank || f == GTF_F_label) v = 0xff;   // all-ones: NaN as f64, -1 as i32
        if (f == GTF_F_alive) v = 1;
        CK(cudaMemsetAsync(b->f[f], v, bytes, st));
    }
    if (N) {
        CK(cudaMemsetAsync(b->accepted_total, 0, (size_t)N, st));
        CK(cudaMemsetAsync(b->d.has_merged_nx, 0, (size_t)N, st));
    }
    // 3. tiles from the host copies of the offsets (while the copies fly), derived topology on the device
    TRY_(build_tiles(b, ev->in_off, ev->out_off));
    sync_dev_view(b);
    if (N) k_derive_slot_dst<<<(N + 255) / 256, 256, 0, st>>>(b->d);
    if (E) k_derive_rev_slot<<<(E + 255) / 256, 256, 0, st>>>(b->d);
    // 4. per-sub-graph counters / node flags (every node alive: no read-back needed)
    CK(cudaMemsetAsync(b->d.sub_nalive, 0, sizeof(int32_t) * (S ? S : 1), st));
    CK(cudaMemsetAsync(b->n_dead, 0, sizeof(unsigned long long), st));
    if (N) {
        k_sub_count<<<(N + 255) / 256, 256, 0, st>>>(b->d);
        k_node_ok<<<(N + 255) / 256, 256, 0, st>>>(b->d, b->n_dead);
    }
    CK(cudaGetLastError());
    b->d.all_alive = 1;
    b->derived_dirty = false;
    // 5. the packed copy is rebuilt from the fields at the next iteration
    b->pack_static_stale = true;
    b->exists_stale = true;
    for (int q = 0; q < PG_N; q++) { b->pack_stale[q] = true; b->soa_stale[q] = false; }
    b->force_pending = true;
    b->have_last_prog = false;
    b->finalized = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------ stats
static int counters_reset(gtf_batch *b)
{
    CK(cudaMemsetAsync(b->d.counters, 0, sizeof(unsigned long long) * GTF_NCOUNTERS_ALL, b->stream));
    return 0;
}
static int counters_read(gtf_batch *b, gtf_stats *st)
{
    CK(cudaMemcpyAsync(b->h_counters, b->d.counters, sizeof(unsigned long long) * GTF_NCOUNTERS_ALL, cudaMemcpyDeviceToHost,
                       b->stream));
    CK(cudaStreamSynchronize(b->stream));
    if (st) {
        st->nodes_merged = (int64_t)b->h_counters[CNT_MERGED];
        st->edges_deactivated = (int64_t)b->h_counters[CNT_DEACT];
        st->edges_sent = (int64_t)b->h_counters[CNT_SENT];
        st->edges_gated = (int64_t)b->h_counters[CNT_GATED];
        st->edges_reweight_off = (int64_t)b->h_counters[CNT_RWOFF];
        st->active_edges = (int64_t)b->h_counters[CNT_ACTIVE];
        st->active_changed = (int64_t)b->h_counters[CNT_CHANGED];
        st->ref_errors = (int64_t)b->h_counters[CNT_REFERR];
        st->near_threshold = (int64_t)b->h_counters[CNT_NEAR];
    }
    return 0;
}

static GtfGeom geom_of(const gtf_geom *g)
{
    GtfGeom o;
    o.sigma0xy = g->sigma0xy; o.sigma0rz = g->sigma0rz; o.sigma0rz2 = g->sigma0rz2; o.endcap = g->endcap_boundary;
    return o;
}
static GtfGeom geom_default()
{
    GtfGeom o;
    o.sigma0xy = 0.3; o.sigma0rz = 0.4; o.sigma0rz2 = 0.6; o.endcap = 550.0;
    return o;
}

static int launch_tile(gtf_batch *b, const Prog &P, const GtfGeom &g)
{
    if (!b->finalized) return fail(GTF_E_STATE, "batch not finalized (call gtf_batch_finalize after uploading the topology)");
    CK(cudaSetDevice(b->device));
    TRY_(soa_for_stage(b, true));
    if (b->n_tiles == 0) return 0;
    if (b->derived_dirty) {
        int r_ = recount_subs(b);
        if (r_) return r_;
    }
    k_tile<<<b->n_tiles, GTF_TILE_THREADS, sizeof(TileSmem), b->stream>>>(b->d, P, g);
    CK(cudaGetLastError());
    return 0;
}
static Prog make_prog(int key, int wb, std::initializer_list<int> ops)
{
    Prog P;
    memset(&P, 0, sizeof(P));
    int k = 0;
    for (int op : ops) P.ops[k++] = op;
    P.key = key;
    P.wb = wb;
    P.rw_thr = 0.1;
    return P;
}
static int launch_prefix(gtf_batch *b, const GtfGeom &g)
{
    TRY_(soa_for_stage(b, true));
    if (b->derived_dirty) {
        int r_ = recount_subs(b);
        if (r_) return r_;
    }
    if (b->N) k_prefix<<<(b->N + 127) / 128, 128, 0, b->stream>>>(b->d, g);
    CK(cudaGetLastError());
    return 0;
}
#define TRY(x)              \
    do {                    \
        int r_ = (x);       \
        if (r_) return r_;  \
    } while (0)

// ------------------------------------------------------------------------------------------------ seeding
// pack != 0: the entry also goes straight into the packed layout (what k_pack_tse would build from the fields afterwards:
// state / weight / tag / geometry records, activation = presence = every slot, existing-edge bits), for gtf_seed_cluster
__global__ void k_seed_slots(DevBatch B, DevPack K, GtfGeom g, int pack)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = s < B.E;
    int key = -1, i = 0;
    GtfState o;
    if (in) {
        i = B.slot_dst[s];
        int b0 = B.in_off[i], d = B.in_off[i + 1] - b0, k = s - b0;
        key = B.in_src[s];
        int other = B.in_src[b0 + d - 1 - k]; // quirk 5: tau of the mirrored neighbour
        double zA = B.z[i], rA = B.r[i];
        double dz = B.z[other] - zA, dr = B.r[other] - rA;
        double vt = gtf_var_tau(dz, dr, zA, B.z[other], g);
        gtf_seed_entry(B.x[i], B.y[i], zA, rA, B.x[key], B.y[key], B.z[key], B.r[key], dz / dr, vt * vt, g, o);
        B.tse_present[s] = 1;
        B.tse_a[s] = o.a; B.tse_b[s] = o.b; B.tse_c[s] = o.c; B.tse_tau[s] = o.tau;
        B.tse_p00[s] = o.p00; B.tse_p01[s] = o.p01; B.tse_p11[s] = o.p11; B.tse_p22[s] = o.p22;
    }
    if (!pack) return;
    const bool ex = in && key >= 0 && B.alive[key] && B.alive[i];
    const unsigned mex = __ballot_sync(0xffffffffu, ex), min_ = __ballot_sync(0xffffffffu, in);
    if ((threadIdx.x & 31) == 0 && min_) {
        K.act[s >> 5] = min_; K.pres[s >> 5] = min_; K.exists[s >> 5] = mex;       // initialize_edge_activation: all 1
        if (mex != min_) atomicAdd(&K.counts[PK_MISSING], __popc(min_ & ~mex));
    }
    if (!in) return;
    GeoRec gr;
    gr.sx = B.x[key]; gr.lay = B.layer[key]; gr.src = key;
    K.aux[s].g = gr;
    double2 *st = reinterpret_cast<double2 *>(K.state + 8 * (size_t)s);
    st[0] = make_double2(o.a, o.b);
    st[1] = make_double2(o.c, o.tau);
    st[2] = make_double2(o.p00, o.p01);
    st[3] = make_double2(o.p11, o.p22);
    MetaRec m;
    m.w = B.tse_w[s]; m.lik = 0.0; m.prior = B.tse_prior[s]; m.ew = B.edge_w[s];
    K.meta[s] = m;
    TagRec t;
    t.rank = s; t.side = 0; t.pad = 0; t.lrn = -1;
    K.aux[s].t = t;
}
// np.var of the xy edge gradients (helper.py:446), in set-iteration order = reversed slot order
__global__ void k_seed_nodes(DevBatch B)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B.N) return;
    int b0 = B.in_off[i], d = B.in_off[i + 1] - b0;
    double xA = B.x[i], yA = B.y[i], sum = 0.0;
    for (int q = d - 1; q >= 0; q--) { int nb = B.in_src[b0 + q]; sum += (B.y[nb] - yA) / (B.x[nb] - xA); }
    double mean = d ? sum / d : NAN, var = 0.0;
    for (int q = d - 1; q >= 0; q--) {
        int nb = B.in_src[b0 + q];
        double t = (B.y[nb] - yA) / (B.x[nb] - xA) - mean;
        var += t * t;
    }
    B.emp_var[i] = d ? var / d : NAN;
}

extern "C" int gtf_seed(gtf_batch *b, const gtf_geom *g)
{
    if (!b || !g) return fail(GTF_E_ARG, "gtf_seed: null argument");
    if (!b->finalized) return fail(GTF_E_STATE, "batch not finalized");
    CK(cudaSetDevice(b->device));
    TRY_(soa_for_stage(b, true));
    GtfGeom gg = geom_of(g);
    if (b->E) k_seed_slots<<<(b->E + 255) / 256, 256, 0, b->stream>>>(b->d, b->k, gg, 0);
    if (b->N) k_seed_nodes<<<(b->N + 255) / 256, 256, 0, b->stream>>>(b->d);
    CK(cudaGetLastError());
    return 0;
}
/* event_conversion.py:87-96 in one call: seed, initialize_edge_activation, compute_prior_probabilities, compute_mixture_weights,
 * node degrees -- the last three as ONE launch of the per-stage kernel */
extern "C" int gtf_seed_all(gtf_batch *b, const gtf_geom *g, gtf_stats *st)
{
    TRY(gtf_seed(b, g));
    CK(cudaMemsetAsync(b->f[GTF_F_active], 1, (size_t)(b->E ? b->E : 0), b->stream));
    TRY(counters_reset(b));
    Prog P = make_prog(GTF_KEY_TSE, WB_PRIOR | WB_W, {OP_PRIOR, OP_WEIGHTS, OP_DEGREE});
    TRY(launch_tile(b, P, geom_default()));
    return st ? counters_read(b, st) : 0;   // st == NULL: no read-back, the call stays asynchronous
}

extern "C" int gtf_initialize_edge_activation(gtf_batch *b)
{
    if (!b) return fail(GTF_E_ARG, "null batch");
    CK(cudaSetDevice(b->device));
    TRY_(soa_for_stage(b, true));
    CK(cudaMemsetAsync(b->f[GTF_F_active], 1, (size_t)(b->E ? b->E : 0), b->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------ stages
extern "C" int gtf_compute_prior_probabilities(gtf_batch *b, int key)
{
    if (!b) return fail(GTF_E_ARG, "null batch");
    Prog P = make_prog(key, WB_PRIOR, {OP_PRIOR});
    return launch_tile(b, P, geom_default());
}
extern "C" int gtf_compute_mixture_weights(gtf_batch *b, int key, gtf_stats *st)
{
    if (!b) return fail(GTF_E_ARG, "null batch");
    TRY(counters_reset(b));
    Prog P = make_prog(key, WB_W, {OP_WEIGHTS});
    TRY(launch_tile(b, P, geom_default()));
    return counters_read(b, st);
}
extern "C" int gtf_query_node_degree(gtf_batch *b)
{
    if (!b) return fail(GTF_E_ARG, "null batch");
    Prog P = make_prog(GTF_KEY_TSE, 0, {OP_DEGREE});
    return launch_tile(b, P, geom_default());
}
extern "C" int gtf_cluster(gtf_batch *b, int key, double chi2_thr, double kl_thr, const double *kl_lut,
                           const gtf_geom *g, gtf_stats *st)
{
    if (!b || !g) return fail(GTF_E_ARG, "gtf_cluster: null argument");
    TRY(counters_reset(b));
    Prog P = make_prog(key, WB_ACTIVE | WB_PRIOR | WB_W | WB_COUNT_ACTIVE, {OP_CLUSTER, OP_DEGREE, OP_WEIGHTS, OP_PRIOR});
    P.cl_chi2 = chi2_thr;
    P.cl_kl = kl_thr;
    if (kl_lut) { P.use_lut = 1; memcpy(P.lut, kl_lut, sizeof(double) * 28); }
    if (key == GTF_KEY_TSE && !g_tile_cluster) return cluster_seeds_packed(b, P, geom_of(g), 0, st);
    TRY(launch_tile(b, P, geom_of(g)));
    return st ? counters_read(b, st) : 0;   // st == NULL: no read-back, the call stays asynchronous
}
extern "C" int gtf_message_passing(gtf_batch *b, double chi2_cut, const gtf_geom *g, gtf_stats *st)
{
    if (!b || !g) return fail(GTF_E_ARG, "gtf_message_passing: null argument");
    if (!b->finalized) return fail(GTF_E_STATE, "batch not finalized");
    CK(cudaSetDevice(b->device));
    TRY(counters_reset(b));
    GtfGeom gg = geom_of(g);
    TRY(launch_prefix(b, gg));
    Prog P = make_prog(GTF_KEY_UTS, WB_ACTIVE | WB_PRESENT | WB_STATE | WB_PRIOR | WB_W | WB_UTSX | WB_COUNT_ACTIVE, {OP_E});
    P.chi2_cut = chi2_cut;
    TRY(launch_tile(b, P, gg));
    // the accumulated multiple-scattering term persists on the node attribute (quirk 2)
    CK(cudaMemcpyAsync(b->d.m_p11, b->d.node_p11tot, sizeof(double) * (size_t)b->N, cudaMemcpyDeviceToDevice, b->stream));
    return counters_read(b, st);
}
extern "C" int gtf_reweight(gtf_batch *b, int key, double threshold, gtf_stats *st)
{
    if (!b) return fail(GTF_E_ARG, "null batch");
    TRY(counters_reset(b));
    if (key == GTF_KEY_UTS) {
        Prog P = make_prog(key, WB_ACTIVE | WB_W | WB_UTSX | WB_EDGEW | WB_COUNT_ACTIVE, {OP_RW});
        P.rw_thr = threshold;
        TRY(launch_tile(b, P, geom_default()));
    }
    return counters_read(b, st);
}
extern "C" int gtf_extrapolate_stage(gtf_batch *b, double chi2_cut, const gtf_geom *g, gtf_stats *st)
{
    if (!b || !g) return fail(GTF_E_ARG, "gtf_extrapolate_stage: null argument");
    if (!b->finalized) return fail(GTF_E_STATE, "batch not finalized");
    CK(cudaSetDevice(b->device));
    TRY(counters_reset(b));
    GtfGeom gg = geom_of(g);
    TRY(launch_prefix(b, gg));
    Prog P = make_prog(GTF_KEY_UTS,
                       WB_ACTIVE | WB_PRESENT | WB_STATE | WB_PRIOR | WB_W | WB_UTSX | WB_EDGEW | WB_COUNT_ACTIVE,
                       {OP_E, OP_PRIOR, OP_RW, OP_PRIOR, OP_RW, OP_DEGREE});
    P.chi2_cut = chi2_cut;
    TRY(launch_tile(b, P, gg));
    CK(cudaMemcpyAsync(b->d.m_p11, b->d.node_p11tot, sizeof(double) * (size_t)b->N, cudaMemcpyDeviceToDevice, b->stream));
    return counters_read(b, st);
}
extern "C" int gtf_remove_state_metadata(gtf_batch *b, gtf_stats *st)
{
    if (!b) return fail(GTF_E_ARG, "null batch");
    TRY(counters_reset(b));
    Prog P1 = make_prog(GTF_KEY_TSE, WB_PRESENT | WB_PRIOR, {OP_POP, OP_PRIOR});
    TRY(launch_tile(b, P1, geom_default()));
    Prog P2 = make_prog(GTF_KEY_UTS, WB_ACTIVE | WB_PRESENT | WB_PRIOR | WB_W | WB_UTSX | WB_EDGEW | WB_COUNT_ACTIVE,
                        {OP_POP, OP_PRIOR, OP_RW});
    TRY(launch_tile(b, P2, geom_default()));
    return counters_read(b, st);
}

// ------------------------------------------------------------------------------------------------ fused iteration
static Prog fused_prog(const gtf_iter_params *p)
{
    Prog P = make_prog(GTF_KEY_UTS,
                       WB_ACTIVE | WB_PRESENT | WB_STATE | WB_PRIOR | WB_W | WB_UTSX | WB_EDGEW | WB_COUNT_ACTIVE,
                       {OP_E, OP_PRIOR, OP_RW, OP_PRIOR, OP_RW, OP_CLUSTER, OP_DEGREE, OP_WEIGHTS, OP_PRIOR});
    P.chi2_cut = p->chi2_cut;
    P.cl_chi2 = p->cluster_chi2;
    P.cl_kl = p->cluster_kl;
    P.rw_thr = p->reweight_threshold;
    P.pre_passes = 2;
    if (p->kl_lut) { P.use_lut = 1; memcpy(P.lut, p->kl_lut, sizeof(double) * 28); }
    return P;
}
// ---- packed pipeline (gtf_iter.cuh) --------------------------------------------------------------------------------
static int ensure_packed(gtf_batch *b)
{
    if (b->derived_dirty) TRY(recount_subs(b));
    const bool st = b->pack_static_stale;
    const bool any = st || b->pack_out_stale || b->exists_stale || b->pack_stale[0] || b->pack_stale[1] || b->pack_stale[2] || b->pack_stale[PG_NODE];
    if (!any) return 0;
    b->force_pending = true; // the packed state changes from outside: the next committed iteration evaluates every node
    DevPack &k = b->k;
    if (st || b->pack_out_stale) {
        if (b->E) k_pack_out<<<(b->E + 255) / 256, 256, 0, b->stream>>>(b->d, k);
    }
    if (b->N && (st || b->pack_stale[PG_NODE]))
        k_pack_nodes<<<(b->N + 256) / 256, 256, 0, b->stream>>>(b->d, k, st, b->pack_stale[PG_NODE]);   // N + 1 threads: srec sentinel
    if (!st && !b->exists_stale && !b->pack_stale[PG_REC]) {
        // only the flag bytes changed (a host that uploads them every iteration): bytes -> bits, nothing else
        if (b->E && (b->pack_stale[PG_ACT] || b->pack_stale[PG_PRES]))
            k_pack_bits<<<(b->E + 255) / 256, 256, 0, b->stream>>>(b->d, k, b->pack_stale[PG_ACT], b->pack_stale[PG_PRES]);
    } else {
        const bool act = b->pack_stale[PG_ACT] || b->exists_stale;
        if (act) CK(cudaMemsetAsync(k.counts + PK_MISSING, 0, sizeof(int), b->stream));
        if (b->E)
            k_pack_slots<<<(b->E + 255) / 256, 256, 0, b->stream>>>(b->d, k, st, act, b->pack_stale[PG_PRES], b->pack_stale[PG_REC]);
        b->exists_stale = false;
    }
    CK(cudaGetLastError());
    b->pack_static_stale = false;
    b->pack_out_stale = false;
    for (int q = 0; q < PG_N; q++) b->pack_stale[q] = false;
    return 0;
}
template <int G> static void launch_hv(gtf_batch *b, cudaStream_t s, const Prog &P, const GtfGeom &gg, int bin, const MergedOut &mo)
{
    k_hv<G><<<b->n_sm * GTF_HV_MINB, GTF_HV_WARPS * 32, 0, s>>>(b->d, b->k, P, gg, bin, mo);
}
// k_node2 + the cooperative bins (k_hv<4|8|16|32>, k_big) of one pass over the packed dict entries
static int issue_node_kernels(gtf_batch *b, const Prog &P, const GtfGeom &gg, bool commit, bool timed)
{
    DevPack &k = b->k;
    DevBatch &d = b->d;
    cudaStream_t s0 = b->stream;
    if (b->N) k_node2<<<(b->N + GTF_NODE2_THREADS - 1) / GTF_NODE2_THREADS, GTF_NODE2_THREADS, 0, s0>>>(d, k, P);
    if (timed) CK(cudaEventRecord(b->evk[3], s0));
    CK(cudaGetLastError());
    if (b->N) {
        MergedOut mo;
        mo.hm = commit ? d.has_merged : d.has_merged_nx;
        mo.rec = commit ? k.mrec : k.mrec_nx;
        mo.p11 = d.m_p11_nx; // k_begin / k_send wrote every node's accumulated value there; a new cluster replaces it
        mo.ab = commit ? k.mab : nullptr;
        // the bins are independent (disjoint nodes): run them side by side
        CK(cudaEventRecord(b->ev_fork2, s0));
        CK(cudaStreamWaitEvent(b->stream2, b->ev_fork2, 0));
        CK(cudaStreamWaitEvent(b->stream3, b->ev_fork2, 0));
        launch_hv<8>(b, s0, P, gg, 1, mo);
        launch_hv<16>(b, b->stream2, P, gg, 2, mo);
        launch_hv<4>(b, b->stream3, P, gg, 0, mo);
        launch_hv<32>(b, b->stream3, P, gg, 3, mo);
        k_big<<<b->n_sm * 2, 32, GTF_BIG_SMEM, b->stream2>>>(d, k, P, gg, mo);
        CK(cudaGetLastError());
        CK(cudaEventRecord(b->ev_join2, b->stream2));
        CK(cudaEventRecord(b->ev_join3, b->stream3));
        CK(cudaStreamWaitEvent(s0, b->ev_join2, 0));
        CK(cudaStreamWaitEvent(s0, b->ev_join3, 0));
    }
    return 0;
}
// the kernel launches of one iteration (k_begin .. k_hv / k_big), issued on the batch stream (and forked onto the side
// streams for the independent bins); also the body that is captured into a CUDA graph
static int issue_iteration(gtf_batch *b, const Prog &P, const GtfGeom &gg, int record_chi2, bool commit, bool timed, bool sparse = false)
{
    DevPack &k = b->k;
    DevBatch &d = b->d;
    const size_t words = ((size_t)b->E + 31) / 32 + 2;
    cudaStream_t s0 = b->stream;
    if (timed) CK(cudaEventRecord(b->evk[0], s0));
    {
        const int nthr = (int)std::max<size_t>(std::max<size_t>(words, (size_t)b->N), (size_t)GTF_NCOUNTERS_ALL);
        k_begin<<<(nthr + 255) / 256, 256, 0, s0>>>(d, k, (int)words, 1);   // (also resets the iteration's counters)
    }
    if (b->fused_sx) {
        // send + execute as one warp-specialised kernel (k_sx): the message list stays in shared memory
        if (b->n_stiles)
            k_sx<<<std::min(b->n_stiles, b->n_sm * GTF_SX_CTAS), 256, sizeof(SxSmem), s0>>>(
                d, k, reinterpret_cast<const int4 *>(b->stile_begin), b->n_stiles, P.chi2_cut, gg, record_chi2);
        if (timed) CK(cudaEventRecord(b->evk[1], s0));
        if (timed) CK(cudaEventRecord(b->evk[2], s0));
    } else {
        if (b->n_stiles)
            k_send<<<std::min(b->n_stiles, b->n_sm * GTF_SEND_MINB), GTF_SEND_THREADS, sizeof(SendSmem), s0>>>(
                d, k, reinterpret_cast<const int4 *>(b->stile_begin), b->n_stiles, gg);
        // inside a committed loop, once k_compact_out has listed the few out-edges still active (PK_SPARSE, decided on the
        // device), k_send returns at once and k_send_sparse sends; otherwise it is k_send_sparse that returns at once
        if (sparse && b->N) k_send_sparse<<<(b->N + 255) / 256, 256, 0, s0>>>(d, k, gg);
        if (timed) CK(cudaEventRecord(b->evk[1], s0));
        if (b->E) k_exec<<<b->n_sm * GTF_EXEC_MINB * GTF_EXEC_WAVES, GTF_EXEC_THREADS, 0, s0>>>(d, k, P.chi2_cut, gg, record_chi2);
        if (timed) CK(cudaEventRecord(b->evk[2], s0));
    }
    TRY_(issue_node_kernels(b, P, gg, commit, timed));
    if (timed) CK(cudaEventRecord(b->evk[4], s0));
    b->launches_per_iter = 1 + (b->n_stiles ? 1 : 0) + (!b->fused_sx && b->E ? 1 : 0) + (b->N ? 6 : 0) + (sparse && !b->fused_sx && b->N ? 1 : 0); // k_begin, k_sx | k_send, k_exec, k_node2 + k_hv x4 + k_big
    return 0;
}
// one iteration on the packed layout.  commit: the next state becomes the current one (merged states are written in
// place, the activation bitmap and the accumulated p11 swap); otherwise the merged states of this pass go to the shadow
// buffers and nothing the next pass reads is changed (profiling / benchmark entry point).
// The launch sequence is replayed from a CUDA graph (one per {committed, not} x {ping-pong parity}, re-captured when
// the parameters change): nine dependent launches cost more than the kernels themselves on a single event.
static int iterate_packed(gtf_batch *b, const gtf_iter_params *p, const GtfGeom &gg, gtf_stats *st, bool commit, bool sparse = false)
{
    TRY(ensure_packed(b));
    DevPack &k = b->k;
    DevBatch &d = b->d;
    const Prog P = fused_prog(p);
    cudaStream_t s0 = b->stream;
    {
        // nodes without an active in-edge are skipped (k_node2) unless every node has to be evaluated: after a (re)pack,
        // when the thresholds changed, and in uncommitted passes (their results go to shadow buffers)
        const bool changed = !b->have_last_prog || memcmp(&b->last_prog, &P, sizeof(Prog)) != 0 || memcmp(&b->last_geom, &gg, sizeof(GtfGeom)) != 0;
        const int fv = (b->force_pending || changed || !commit) ? 1 : 0;
        if (fv != b->force_dev) {
            CK(cudaMemsetAsync(k.counts + PK_FORCE, fv ? 0xff : 0, sizeof(int), s0));
            b->force_dev = fv;
        }
        if (commit) { b->last_prog = P; b->last_geom = gg; b->have_last_prog = true; b->force_pending = false; }
    }
    if (b->timing || !b->use_graph) {
        TRY(issue_iteration(b, P, gg, p->record_chi2, commit, b->timing, sparse));
        if (b->timing) {
            CK(cudaEventSynchronize(b->evk[4]));
            for (int q = 0; q < 4; q++) {
                float t = 0;
                CK(cudaEventElapsedTime(&t, b->evk[q], b->evk[q + 1]));
                b->t_k[q] += t;
            }
            b->t_count++;
        }
    } else {
        IterGraph &G = b->graphs[commit ? 1 : 0][b->parity][sparse ? 1 : 0];
        // the captured arguments hold sizes and pointers, not the tile tables' contents: a new batch of the same shape in the same
        // batch object replays the graph as it is; another shape updates the instantiated graph in place (cudaGraphExecUpdate:
        // tens of microseconds instead of an instantiation per graph and per loaded batch)
        const bool same = G.exec && memcmp(&G.P, &P, sizeof(Prog)) == 0 && memcmp(&G.g, &gg, sizeof(GtfGeom)) == 0 &&
                          G.record_chi2 == p->record_chi2 && G.n_stiles == b->n_stiles && G.stile == (const void *)b->stile_begin &&
                          G.N == b->N && G.E == b->E && G.S == b->S && G.n_tiles == b->n_tiles;
        if (!same) {
            cudaGraph_t graph = nullptr;
            CK(cudaStreamBeginCapture(s0, cudaStreamCaptureModeThreadLocal));
            int r = issue_iteration(b, P, gg, p->record_chi2, commit, false, sparse);
            cudaError_t e = cudaStreamEndCapture(s0, &graph);
            if (r) { if (graph) cudaGraphDestroy(graph); return r; }
            if (e != cudaSuccess) return fail(GTF_E_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
            if (G.exec) {
                cudaGraphExecUpdateResultInfo info;
                if (cudaGraphExecUpdate(G.exec, graph, &info) != cudaSuccess) {
                    cudaGetLastError();
                    cudaGraphExecDestroy(G.exec);
                    G.exec = nullptr;
                }
            }
            if (!G.exec) e = cudaGraphInstantiate(&G.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) return fail(GTF_E_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
            G.P = P; G.g = gg; G.record_chi2 = p->record_chi2; G.n_stiles = b->n_stiles; G.stile = (const void *)b->stile_begin;
            G.N = b->N; G.E = b->E; G.S = b->S; G.n_tiles = b->n_tiles;
        }
        CK(cudaGraphLaunch(G.exec, s0));
    }
    b->launches += b->launches_per_iter;
    for (int q = 0; q < 3; q++) b->soa_stale[q] = true; // (a dry pass also rewrites dict entries in place)
    if (commit) b->soa_stale[PG_NODE] = true;
    if (commit) {
        std::swap(k.act, k.act_nx);
        std::swap(b->f[GTF_F_m_p11], *(void **)&d.m_p11_nx);
        sync_dev_view(b);
        b->parity ^= 1;
    }
    if (st) return counters_read(b, st);
    return 0;
}

__global__ void k_count_flags(const uint8_t *f, int n, unsigned long long *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned m = __ballot_sync(0xffffffffu, i < n && f[i] != 0);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, (unsigned long long)__popc(m));
}
// cluster() on the SEED dict (clustering.py:149-376 with 'track_state_estimates') on the packed node kernels: the seed
// entries are packed like updated states (slot order = dict order), k_node2 / k_hv / k_big run without the re-weighting
// (pre_passes (prior) passes first: 0 = the stage as the reference defines it, 1 = preceded by the seed's own
// compute_prior_probabilities), then activation flags, weights and priors go back to the fields.
static int cluster_seeds_packed(gtf_batch *b, Prog P, const GtfGeom &gg, int pre_passes, gtf_stats *st, const gtf_geom *seed_first)
{
    if (!b->finalized) return fail(GTF_E_STATE, "batch not finalized");
    TRY_(soa_for_stage(b, true));                 // every field current; the packed records are about to hold the seed dict
    if (b->derived_dirty) TRY_(recount_subs(b));
    DevPack &k = b->k;
    DevBatch &d = b->d;
    cudaStream_t s0 = b->stream;
    P.key = GTF_KEY_TSE;
    P.pre_passes = pre_passes;
    {
        int n = 0;
        if (pre_passes) P.ops[n++] = OP_PRIOR;
        P.ops[n++] = OP_CLUSTER; P.ops[n++] = OP_DEGREE; P.ops[n++] = OP_WEIGHTS; P.ops[n++] = OP_PRIOR; P.ops[n] = OP_END;
    }
    const size_t words = ((size_t)b->E + 31) / 32 + 2;
    // does any node hold updated states?  (`has_uts` is current: soa_for_stage above)
    bool uts_empty = true;
    if (b->N) {
        CK(cudaMemsetAsync(b->n_dead, 0, sizeof(unsigned long long), s0));
        k_count_flags<<<(b->N + 255) / 256, 256, 0, s0>>>((const uint8_t *)b->f[GTF_F_has_uts], b->N, b->n_dead);
        unsigned long long nu = 0;
        CK(cudaMemcpyAsync(&nu, b->n_dead, sizeof(nu), cudaMemcpyDeviceToHost, s0));
        CK(cudaStreamSynchronize(s0));
        uts_empty = nu == 0;
    }
    CK(cudaMemsetAsync(d.counters, 0, sizeof(unsigned long long) * GTF_NCOUNTERS_ALL, s0));
    CK(cudaMemsetAsync(k.counts + PK_MISSING, 0, sizeof(int), s0));
    CK(cudaMemsetAsync(k.counts + PK_FORCE, 0xff, sizeof(int), s0));       // every node is evaluated
    b->force_dev = 1;
    if (b->N) k_pack_nodes<<<(b->N + 256) / 256, 256, 0, s0>>>(d, k, 1, 1);
    if (seed_first) {
        // gtf_seed_cluster: helper.py:238-452 + :24-25 with the entries written to the fields AND the packed records at once
        CK(cudaMemsetAsync(b->f[GTF_F_active], 1, (size_t)(b->E ? b->E : 0), s0));
        if (b->E) k_seed_slots<<<(b->E + 255) / 256, 256, 0, s0>>>(d, k, geom_of(seed_first), 1);
        if (b->N) k_seed_nodes<<<(b->N + 255) / 256, 256, 0, s0>>>(d);
    } else if (b->E)
        k_pack_tse<<<(b->E + 255) / 256, 256, 0, s0>>>(d, k);
    {
        const int nthr = (int)std::max<size_t>(words, (size_t)b->N);
        k_begin<<<(nthr + 255) / 256, 256, 0, s0>>>(d, k, (int)words, 0);
    }
    CK(cudaGetLastError());
    TRY_(issue_node_kernels(b, P, gg, true, false));
    // commit: the next activation bitmap and the carried merged_cov[1,1] become current
    std::swap(k.act, k.act_nx);
    std::swap(b->f[GTF_F_m_p11], *(void **)&d.m_p11_nx);
    sync_dev_view(b);
    b->parity ^= 1;
    if (b->E) k_unpack_tse<<<(b->E + 255) / 256, 256, 0, s0>>>(b->d, k);
    CK(cudaGetLastError());
    // what is where now: activation bits are current on both sides, the merged records are newer in the packed copy, the
    // geometry / node records were just rebuilt; the per-out-edge records carry the OLD seed weights.  The dict-entry
    // records hold the SEED dict: when the batch has no updated states at all (fresh events: the usual case) the presence
    // bitmap is simply cleared -- absent entries are never read --, otherwise the updated-state groups are re-packed from
    // the fields before the next iteration.
    b->soa_stale[PG_ACT] = false; b->pack_stale[PG_ACT] = false; b->exists_stale = false;
    b->soa_stale[PG_PRES] = b->soa_stale[PG_REC] = false;
    if (uts_empty) {
        CK(cudaMemsetAsync(k.pres, 0, sizeof(uint32_t) * words, s0));
        b->pack_stale[PG_PRES] = b->pack_stale[PG_REC] = false;
    } else
        b->pack_stale[PG_PRES] = b->pack_stale[PG_REC] = true;
    b->soa_stale[PG_NODE] = true; b->pack_stale[PG_NODE] = false;
    b->pack_static_stale = false;
    b->pack_out_stale = true;                     // the carried seed weights (per-out-edge records) changed
    b->force_pending = true;
    b->have_last_prog = false;
    return st ? counters_read(b, st) : 0;
}
extern "C" int gtf_seed_cluster(gtf_batch *b, const gtf_geom *g, double chi2_threshold, double kl_threshold, const double *kl_lut,
                                gtf_stats *st)
{
    if (!b || !g) return fail(GTF_E_ARG, "gtf_seed_cluster: null argument");
    if (!b->finalized) return fail(GTF_E_STATE, "batch not finalized");
    CK(cudaSetDevice(b->device));
    Prog P = make_prog(GTF_KEY_TSE, 0, {OP_END});
    P.cl_chi2 = chi2_threshold;
    P.cl_kl = kl_threshold;
    if (kl_lut) { P.use_lut = 1; memcpy(P.lut, kl_lut, sizeof(double) * 28); }
    return cluster_seeds_packed(b, P, geom_of(g), 1, st, g);
}

extern "C" int gtf_iterate_dry(gtf_batch *b, const gtf_iter_params *p, const gtf_geom *g, gtf_stats *st)
{
    if (!b || !p || !g) return fail(GTF_E_ARG, "gtf_iterate_dry: null argument");
    if (!b->finalized) return fail(GTF_E_STATE, "batch not finalized");
    CK(cudaSetDevice(b->device));
    return iterate_packed(b, p, geom_of(g), st, false);
}
// per-kernel timing of the fused iteration with CUDA events on the batch stream (bench.py roofline):
// enable=1 resets the accumulators; gtf_batch_timing returns the averages since then
extern "C" int gtf_batch_set_timing(gtf_batch *b, int enable)
{
    if (!b) return fail(GTF_E_ARG, "null batch");
    CK(cudaSetDevice(b->device));
    if (enable && !b->evk[0])
        for (int k = 0; k < 6; k++) CK(cudaEventCreate(&b->evk[k]));
    for (int k = 0; k < 5; k++) b->t_k[k] = 0.0;
    b->timing = enable != 0;
    b->t_count = 0;
    return 0;
}
extern "C" int gtf_batch_timing(gtf_batch *b, double *prefix_ms, double *tile_ms, double *heavy_ms, int *count)
{
    if (!b) return fail(GTF_E_ARG, "null batch");
    int n = b->t_count ? b->t_count : 1; // k_begin + k_send | k_exec + k_node2 | k_hv<*> + k_big
    if (prefix_ms) *prefix_ms = b->t_k[0] / n;
    if (tile_ms) *tile_ms = (b->t_k[1] + b->t_k[2]) / n;
    if (heavy_ms) *heavy_ms = b->t_k[3] / n;
    if (count) *count = b->t_count;
    return 0;
}
/* per-kernel averages of the packed pipeline: ms[0..3] = k_send, k_exec, k_node2, cooperative kernels */
extern "C" int gtf_batch_timing_kernels(gtf_batch *b, double *ms, int n_ms, int *count)
{
    if (!b || !ms) return fail(GTF_E_ARG, "null argument");
    int n = b->t_count ? b->t_count : 1;
    for (int q = 0; q < n_ms && q < 5; q++) ms[q] = b->t_k[q] / n;
    if (count) *count = b->t_count;
    return 0;
}
extern "C" int gtf_iterate(gtf_batch *b, const gtf_iter_params *p, const gtf_geom *g, int max_iter, int stop_when_converged,
                           gtf_stats *stats, int *n_done)
{
    if (!b || !p || !g) return fail(GTF_E_ARG, "gtf_iterate: null argument");
    if (!b->finalized) return fail(GTF_E_STATE, "batch not finalized");
    CK(cudaSetDevice(b->device));
    int it = 0;
    const bool need = stats != nullptr || stop_when_converged; // without either the call stays asynchronous (no counter read-back)
    if (!need || b->timing) {
        for (; it < max_iter; it++) {
            gtf_stats st;
            memset(&st, 0, sizeof(st));
            TRY(iterate_packed(b, p, geom_of(g), need ? &st : nullptr, true));
            if (stats) stats[it] = st;
            if (stop_when_converged && st.active_changed == 0) { it++; break; }
        }
    } else {
        // The loop runs on the device: up to GTF_LOOP_BURST iterations are queued back to back, each followed by k_iter_end,
        // which files the iteration's counters and raises the stop flag when no activation flag changed -- the iterations
        // queued behind it then do nothing.  One read-back per burst instead of one host round trip per iteration
        // (0.15 ms each: a third of a converged loop on a 128-event batch).
        DevPack &k = b->k;
        cudaStream_t s0 = b->stream;
        bool stopped = false;
        const bool may_compact = !b->fused_sx && max_iter >= 3 && b->N && b->E;   // (worth one more pass over the out-edges)
        CK(cudaMemsetAsync(k.counts + PK_SPARSE, 0, 2 * sizeof(int), s0));
        while (it < max_iter && !stopped) {
            const int burst = std::min(max_iter - it, (int)GTF_LOOP_BURST);
            CK(cudaMemsetAsync(k.counts + PK_STOP, 0, 2 * sizeof(int), s0));   // stop flag, iterations done
            for (int q = 0; q < burst; q++) {
                if (may_compact && it + q == 1) {        // after the loop's first iteration (the ping-pong pairs are swapped:
                    k_compact_out<<<(b->N + 255) / 256, 256, 0, s0>>>(b->d, k);   // k.act is the current bitmap)
                    b->launches++;
                }
                TRY(iterate_packed(b, p, geom_of(g), nullptr, true, may_compact && it + q >= 1));
                k_iter_end<<<1, 32, 0, s0>>>(b->d.counters, b->loop_stats, k.counts, stop_when_converged);
                b->launches++;
            }
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(b->h_loop_stats, b->loop_stats, sizeof(unsigned long long) * GTF_NCOUNTERS_ALL * burst,
                               cudaMemcpyDeviceToHost, s0));
            CK(cudaMemcpyAsync(b->h_loop_done, k.counts + PK_DONE, sizeof(int), cudaMemcpyDeviceToHost, s0));
            CK(cudaMemsetAsync(k.counts + PK_STOP, 0, sizeof(int), s0));
            if (it + burst >= max_iter) CK(cudaMemsetAsync(k.counts + PK_SPARSE, 0, sizeof(int), s0));
            CK(cudaStreamSynchronize(s0));
            const int done = *b->h_loop_done;
            if (done < 1 || done > burst) return fail(GTF_E_STATE, "gtf_iterate: loop bookkeeping out of range");
            if ((burst - done) & 1) {                    // the host swapped the ping-pong pairs once per QUEUED iteration
                std::swap(k.act, k.act_nx);
                std::swap(b->f[GTF_F_m_p11], *(void **)&b->d.m_p11_nx);
                sync_dev_view(b);
                b->parity ^= 1;
            }
            for (int q = 0; q < done; q++) {
                const unsigned long long *c = b->h_loop_stats + (size_t)q * GTF_NCOUNTERS_ALL;
                if (stats) {
                    gtf_stats &st = stats[it + q];
                    st.nodes_merged = (int64_t)c[CNT_MERGED]; st.edges_deactivated = (int64_t)c[CNT_DEACT];
                    st.edges_sent = (int64_t)c[CNT_SENT]; st.edges_gated = (int64_t)c[CNT_GATED];
                    st.edges_reweight_off = (int64_t)c[CNT_RWOFF]; st.active_edges = (int64_t)c[CNT_ACTIVE];
                    st.active_changed = (int64_t)c[CNT_CHANGED]; st.ref_errors = (int64_t)c[CNT_REFERR];
                    st.near_threshold = (int64_t)c[CNT_NEAR];
                }
            }
            it += done;
            if (done < burst) CK(cudaMemsetAsync(k.counts + PK_SPARSE, 0, sizeof(int), s0));
            stopped = done < burst || (stop_when_converged && b->h_loop_stats[(size_t)(done - 1) * GTF_NCOUNTERS_ALL + CNT_CHANGED] == 0);
        }
    }
    if (n_done) *n_done = it;
    return 0;
}

// ------------------------------------------------------------------------------------------------ components
__device__ __forceinline__ int uf_find(int32_t *p, int i)
{
    int r = i;
    while (true) {
        int q = p[r];
        if (q == r) break;
        int qq = p[q];
        if (qq != q) p[r] = qq; // path halving: benign race, qq is still an ancestor of r
        r = q;
    }
    return r;
}
__device__ __forceinline__ int uf_find_ro(const int32_t *p, int i)
{
    int r = i;
    while (true) {
        int q = p[r];
        if (q == r) break;
        r = q;
    }
    return r;
}
__global__ void k_cca_init(DevBatch B, uint8_t *has_inactive, int32_t *first)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B.S) { has_inactive[i] = 0; first[i] = 0x7fffffff; }
    if (i < B.N) B.label[i] = B.alive[i] ? i : -1;
}
__global__ void k_cca_first(DevBatch B, int32_t *first)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    // consecutive nodes mostly share a sub-graph: one atomic per distinct sub-graph in the warp (its lowest lane = lowest index)
    const int sg = (i < B.N && B.alive[i]) ? B.sub[i] : -1 - (int)(threadIdx.x & 31);
    const unsigned same = __match_any_sync(0xffffffffu, sg);
    if (sg >= 0 && (int)(threadIdx.x & 31) == __ffs(same) - 1) atomicMin(&first[sg], i);
}
__global__ void k_cca_edges(DevBatch B, uint8_t *has_inactive)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= B.E) return;
    int u = B.in_src[s], v = B.slot_dst[s];
    if (u < 0 || !B.alive[u] || !B.alive[v]) return;
    int sg = B.sub[v];
    if (B.sub_state[sg] != GTF_SUB_INPLAY) return;
    if (B.active[s] == 0) { has_inactive[sg] = 1; return; } // extract...py:335
    int32_t *p = B.label;
    int a = u, c = v;
    while (true) { // lock-free union: hook the larger root under the smaller one
        a = uf_find(p, a);
        c = uf_find(p, c);
        if (a == c) break;
        int hi = max(a, c), lo = min(a, c);
        int old = atomicCAS(&p[hi], hi, lo);
        if (old == hi) break;
        a = old; c = lo;
    }
}
__global__ void k_cca_final(DevBatch B, const uint8_t *has_inactive, const int32_t *first)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B.N || !B.alive[i]) return;
    int sg = B.sub[i];
    if (B.sub_state[sg] != GTF_SUB_INPLAY) return;
    // extract...py:343-344: with no inactive edge the WHOLE sub-graph is one candidate
    // read-only find: a compressing find here could overwrite another node's FINAL label with a stale ancestor
    B.label[i] = has_inactive[sg] ? uf_find_ro(B.label, i) : first[sg];
}
extern "C" int gtf_components(gtf_batch *b)
{
    if (!b) return fail(GTF_E_ARG, "null batch");
    if (!b->finalized) return fail(GTF_E_STATE, "batch not finalized");
    CK(cudaSetDevice(b->device));
    TRY_(soa_sync(b, 1u << PG_ACT));
    int n = b->N > b->S ? b->N : b->S;
    if (n == 0) return 0;
    k_cca_init<<<(n + 255) / 256, 256, 0, b->stream>>>(b->d, b->sub_has_inactive, b->sub_first);
    if (b->N) k_cca_first<<<(b->N + 255) / 256, 256, 0, b->stream>>>(b->d, b->sub_first);
    if (b->E) k_cca_edges<<<(b->E + 255) / 256, 256, 0, b->stream>>>(b->d, b->sub_has_inactive);
    // pointer-jump pass needs the unions finished: separate launch
    if (b->N) k_cca_final<<<(b->N + 255) / 256, 256, 0, b->stream>>>(b->d, b->sub_has_inactive, b->sub_first);
    CK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ extraction
#define GTF_MAX_CAND 64
__global__ void k_extract_keys(DevBatch B, int32_t *keys, int32_t *vals)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B.N) return;
    bool ok = B.alive[i] && B.sub_state[B.sub[i]] == GTF_SUB_INPLAY;
    keys[i] = ok ? B.label[i] : 0x7fffffff;
    vals[i] = i;
}
// one thread per component head (extract...py:409-456)
__global__ void k_extract_gate(DevBatch B, const int32_t *keys, const int32_t *vals, GtfGeom g, double pval_cut,
                               int numhits, double sep3d, double merge_dist, uint8_t *acc, int32_t *root_out,
                               double *pv_xy, double *pv_zr)
{
    int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= B.N) return;
    int key = keys[pos];
    if (key == 0x7fffffff || (pos > 0 && keys[pos - 1] == key)) return;
    int n = 1;
    while (pos + n < B.N && keys[pos + n] == key) n++;
    if (n < numhits) return; // :415
    if (n > GTF_MAX_CAND) {
        // A component this large passes the one-hit-per-layer test (:427-429, with at most two close pairs :58-151) only
        // on a detector with more than GTF_MAX_CAND - 2 distinct (volume, layer) ids.  Prove the rejection on its first
        // nodes (duplicates inside a subset are duplicates of the whole); if that fails, say so instead of skipping.
        const int K = min(n, 3 * GTF_MAX_CAND);
        int n2 = 0, bad = 0;
        for (int p = 0; p < K && !bad && n2 <= 2; p++) {
            const int mp = vals[pos + p], lp = B.volume[mp] * 1000 + B.layer[mp];
            int cnt = 0, first = 1;
            for (int q = 0; q < K; q++) {
                const int mq = vals[pos + q];
                if (B.volume[mq] * 1000 + B.layer[mq] == lp) { cnt++; if (q < p) first = 0; }
            }
            if (!first) continue;
            if (cnt == 2) n2++; else if (cnt != 1) bad = 1;
        }
        if (!bad && n2 <= 2) atomicOr(&B.counters[CNT_REFERR], (unsigned long long)GTF_STATUS_CAND_OVERFLOW);
        return;
    }
    int mem[GTF_MAX_CAND];
    int lay[GTF_MAX_CAND];
    double co[GTF_MAX_CAND][4];
    for (int k = 0; k < n; k++) {
        int m = vals[pos + k];
        mem[k] = m;
        lay[k] = B.volume[m] * 1000 + B.layer[m];
        co[k][0] = B.x[m]; co[k][1] = B.y[m]; co[k][2] = B.z[m]; co[k][3] = B.r[m];
    }
    // check_close_proximity_nodes (:58-151): at most two layers with exactly two nodes, all others one
    int n2 = 0, bad = 0;
    for (int p = 0; p < n; p++) {
        int cnt = 0, first = 1;
        for (int q = 0; q < n; q++)
            if (lay[q] == lay[p]) { cnt++; if (q < p) first = 0; }
        if (!first) continue;
        if (cnt == 2) n2++; else if (cnt != 1) bad = 1;
    }
    bool dropped[GTF_MAX_CAND];
    for (int k = 0; k < n; k++) dropped[k] = false;
    if (n2 > 0) {
        if (n2 > 2 || bad) return; // duplicates stay -> fails the one-hit-per-layer test (:429)
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++) {
                if (lay[q] != lay[p]) continue;
                double dx = co[p][0] - co[q][0], dy = co[p][1] - co[q][1], dz = co[p][2] - co[q][2];
                if (!(sqrt(dx * dx + dy * dy + dz * dz) <= merge_dist)) return; // :136-139
                double xm = (co[p][0] + co[q][0]) / 2, ym = (co[p][1] + co[q][1]) / 2, zm = (co[p][2] + co[q][2]) / 2;
                co[p][0] = xm; co[p][1] = ym; co[p][2] = zm; co[p][3] = sqrt(xm * xm + ym * ym);
                dropped[q] = true;
                break;
            }
    } else if (bad)
        return;
    int nk = 0;
    for (int k = 0; k < n; k++)
        if (!dropped[k]) {
            if (nk != k) { co[nk][0] = co[k][0]; co[nk][1] = co[k][1]; co[nk][2] = co[k][2]; co[nk][3] = co[k][3]; }
            nk++;
        }
    if (nk < numhits) return;
    for (int p = 1; p < nk; p++) { // stable sort by r, largest first (:434-436)
        double t0 = co[p][0], t1 = co[p][1], t2 = co[p][2], t3 = co[p][3];
        int q = p;
        while (q > 0 && co[q - 1][3] < t3) {
            co[q][0] = co[q - 1][0]; co[q][1] = co[q - 1][1]; co[q][2] = co[q - 1][2]; co[q][3] = co[q - 1][3];
            q--;
        }
        co[q][0] = t0; co[q][1] = t1; co[q][2] = t2; co[q][3] = t3;
    }
    double pxy, pzr;
    gtf_track_fit(co, nk, g.sigma0xy, g.sigma0rz, g.endcap, sep3d, pxy, pzr);
    pv_xy[key] = pxy;
    pv_zr[key] = pzr;
    if (pxy >= pval_cut && pzr >= pval_cut) // :442
        for (int k = 0; k < n; k++) { acc[mem[k]] = 1; root_out[mem[k]] = key; }
}
__global__ void k_extract_apply(DevBatch B, const uint8_t *acc, uint8_t *acc_total, unsigned long long *count)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B.N || !acc[i]) return;
    B.alive[i] = 0;
    acc_total[i] = 1;
    if (B.label[i] == i) atomicAdd(count, 1ull);
}
__global__ void k_sub_state(DevBatch B, int numhits)
{
    int sg = blockIdx.x * blockDim.x + threadIdx.x;
    if (sg >= B.S || B.sub_state[sg] != GTF_SUB_INPLAY) return;
    int left = B.sub_nalive[sg];
    if (left == 0) B.sub_state[sg] = GTF_SUB_EMPTY;            // extract...py:463-467
    else if (left < numhits) B.sub_state[sg] = GTF_SUB_FRAGMENT;
}
__global__ void k_fill_nan(double *a, double *c, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { a[i] = NAN; c[i] = NAN; }
}

extern "C" int gtf_extract(gtf_batch *b, const gtf_geom *g, double pval_cut, int numhits, double sep3d, double merge_dist,
                           int32_t *n_accepted, uint8_t *accepted, double *pval_xy, double *pval_zr)
{
    if (!b || !g) return fail(GTF_E_ARG, "gtf_extract: null argument");
    CK(cudaSetDevice(b->device));
    // extraction reads the activation flags and removes nodes: only the flag bytes are brought up to date; the dict-entry
    // and merged-state records stay packed, the existing-edge bitmap is rebuilt from `alive` before the next iteration
    TRY_(soa_sync(b, 1u << PG_ACT));
    b->exists_stale = true;
    TRY(gtf_components(b));
    if (n_accepted) *n_accepted = 0;
    if (b->N == 0) return 0;
    int nb = (b->N + 255) / 256;
    k_extract_keys<<<nb, 256, 0, b->stream>>>(b->d, b->sort_keys, b->sort_vals);
    size_t need = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, need, b->sort_keys, b->sort_keys2, b->sort_vals, b->sort_vals2, b->N, 0, 32,
                                       b->stream));
    if (need > b->sort_tmp_bytes) {
        if (b->sort_tmp) cudaFree(b->sort_tmp);
        CK(cudaMalloc(&b->sort_tmp, need));
        b->sort_tmp_bytes = need;
    }
    CK(cub::DeviceRadixSort::SortPairs(b->sort_tmp, need, b->sort_keys, b->sort_keys2, b->sort_vals, b->sort_vals2, b->N, 0,
                                       32, b->stream));
    CK(cudaMemsetAsync(b->acc_now, 0, (size_t)b->N, b->stream));
    k_fill_nan<<<nb, 256, 0, b->stream>>>(b->pv_xy, b->pv_zr, b->N);
    TRY(counters_reset(b));
    k_extract_gate<<<(b->N + 63) / 64, 64, 0, b->stream>>>(b->d, b->sort_keys2, b->sort_vals2, geom_of(g), pval_cut, numhits,
                                                           sep3d, merge_dist, b->acc_now, b->cand_root, b->pv_xy, b->pv_zr);
    k_extract_apply<<<nb, 256, 0, b->stream>>>(b->d, b->acc_now, b->accepted_total, b->d.counters + CNT_MERGED);
    TRY(recount_subs(b));
    if (b->S) k_sub_state<<<(b->S + 255) / 256, 256, 0, b->stream>>>(b->d, numhits);
    // sub-graphs that just became fragments / empty leave the list (extract...py:463-467): their nodes are no longer in play
    // for any later stage, so the per-node flags derived from sub_state are rebuilt
    CK(cudaMemsetAsync(b->n_dead, 0, sizeof(unsigned long long), b->stream));
    k_node_ok<<<nb, 256, 0, b->stream>>>(b->d, b->n_dead);
    CK(cudaGetLastError());
    if (accepted) CK(cudaMemcpyAsync(accepted, b->acc_now, (size_t)b->N, cudaMemcpyDeviceToHost, b->stream));
    if (pval_xy) CK(cudaMemcpyAsync(pval_xy, b->pv_xy, sizeof(double) * (size_t)b->N, cudaMemcpyDeviceToHost, b->stream));
    if (pval_zr) CK(cudaMemcpyAsync(pval_zr, b->pv_zr, sizeof(double) * (size_t)b->N, cudaMemcpyDeviceToHost, b->stream));
    gtf_stats st;
    TRY(counters_read(b, &st));
    if (n_accepted) *n_accepted = (int32_t)st.nodes_merged;
    if (st.ref_errors & GTF_STATUS_CAND_OVERFLOW)
        return fail(GTF_E_DEGREE, "gtf_extract: a component with more than GTF_MAX_CAND (64) nodes could be a one-hit-per-layer candidate");
    return 0;
}

// ------------------------------------------------------------------------------------------------ tag propagation
__global__ void k_tag_sweep(DevBatch B, const int32_t *tin, int32_t *tout, unsigned long long *cnt)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B.N) return;
    int32_t t = tin[i];
    bool work = false;
    if (B.alive[i] && B.sub_state[B.sub[i]] == GTF_SUB_INPLAY) {
        double ri = B.r[i];
        for (int o = B.out_off[i]; o < B.out_off[i + 1]; o++) {
            int v = B.slot_dst[B.out_slot[o]];
            if (B.alive[v] && !(B.r[v] > ri)) { // keep successors with radius <= own (tag_propagation.py:99-110)
                work = true;
                t = max(t, tin[v]);             // :139-149 (max, although the script calls it "smallest")
            }
        }
    }
    tout[i] = t;
    if (work) {
        atomicAdd(&cnt[0], 1ull);
        if (t != tin[i]) atomicAdd(&cnt[1], 1ull);
    }
}
extern "C" int gtf_tag_propagate(gtf_batch *b, double threshold, int32_t *tags, int max_sweeps, int *n_sweeps)
{
    if (!b || !tags) return fail(GTF_E_ARG, "gtf_tag_propagate: null argument");
    if (!b->finalized) return fail(GTF_E_STATE, "batch not finalized");
    CK(cudaSetDevice(b->device));
    int sweeps = 0;
    if (b->N) {
        CK(cudaMemcpyAsync(b->tags_a, tags, sizeof(int32_t) * (size_t)b->N, cudaMemcpyHostToDevice, b->stream));
        int32_t *cur = b->tags_a, *nxt = b->tags_b;
        double frac = 1.0;
        while (frac > threshold && sweeps < max_sweeps) {
            TRY(counters_reset(b));
            k_tag_sweep<<<(b->N + 255) / 256, 256, 0, b->stream>>>(b->d, cur, nxt, b->d.counters);
            CK(cudaGetLastError());
            TRY(counters_read(b, nullptr));
            unsigned long long nwork = b->h_counters[0], flipped = b->h_counters[1];
            if (nwork == 0) break;
            std::swap(cur, nxt);
            frac = (double)flipped / (double)nwork;
            sweeps++;
        }
        CK(cudaMemcpyAsync(tags, cur, sizeof(int32_t) * (size_t)b->N, cudaMemcpyDeviceToHost, b->stream));
        CK(cudaStreamSynchronize(b->stream));
    }
    if (n_sweeps) *n_sweeps = sweeps;
    return 0;
}

// ------------------------------------------------------------------------------------------------ candidate table
// candidate table: keys = candidate id (root node index) of accepted nodes, INT_MAX otherwise; a stable radix sort by key
// leaves the accepted nodes first, grouped by candidate, ascending node index inside a candidate -- and candidates in
// event order, because events own contiguous node ranges
__global__ void k_cand_keys(DevBatch B, const uint8_t *acc_total, const int32_t *root, int32_t *keys, int32_t *vals,
                            unsigned long long *count)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool a = i < B.N && acc_total[i];
    if (i < B.N) { keys[i] = a ? root[i] : 0x7fffffff; vals[i] = i; }
    const unsigned m = __ballot_sync(0xffffffffu, a);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, (unsigned long long)__popc(m));
}
__global__ void k_cand_rows(DevBatch B, const int32_t *keys, const int32_t *vals, int32_t *rows, long long n)
{
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int i = vals[k];
    rows[3 * k + 0] = B.sub_event[B.sub[i]];
    rows[3 * k + 1] = keys[k];
    rows[3 * k + 2] = i;
}
// count, sort and materialise the rows on the device; *n_rows rows at b->cand_rows (fill == false: count only)
static int candidates_build(gtf_batch *b, bool fill, int64_t cap_rows, int64_t *n_rows)
{
    *n_rows = 0;
    if (b->N == 0) return 0;
    TRY(counters_reset(b));
    k_cand_keys<<<(b->N + 255) / 256, 256, 0, b->stream>>>(b->d, b->accepted_total, b->cand_root, b->sort_keys, b->sort_vals,
                                                           b->d.counters);
    CK(cudaGetLastError());
    if (fill) {
        size_t need = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, need, b->sort_keys, b->sort_keys2, b->sort_vals, b->sort_vals2, b->N, 0, 32,
                                           b->stream));
        if (need > b->sort_tmp_bytes) {
            if (b->sort_tmp) cudaFree(b->sort_tmp);
            CK(cudaMalloc(&b->sort_tmp, need));
            b->sort_tmp_bytes = need;
        }
        CK(cub::DeviceRadixSort::SortPairs(b->sort_tmp, need, b->sort_keys, b->sort_keys2, b->sort_vals, b->sort_vals2, b->N, 0,
                                           32, b->stream));
    }
    TRY(counters_read(b, nullptr));
    *n_rows = (int64_t)b->h_counters[0];
    if (fill && *n_rows > 0) {
        const int64_t nw = *n_rows < cap_rows ? *n_rows : cap_rows;
        if (!b->cand_rows) CK(cudaMalloc((void **)&b->cand_rows, sizeof(int32_t) * 3 * (size_t)(b->capN ? b->capN : 1)));
        k_cand_rows<<<(unsigned)((nw + 255) / 256), 256, 0, b->stream>>>(b->d, b->sort_keys2, b->sort_vals2, b->cand_rows, (long long)nw);
        CK(cudaGetLastError());
    }
    return 0;
}
extern "C" int gtf_candidates(gtf_batch *b, int32_t *table_host, int64_t cap_rows, int64_t *n_rows)
{
    if (!b || !n_rows) return fail(GTF_E_ARG, "gtf_candidates: null argument");
    CK(cudaSetDevice(b->device));
    const bool fill = cap_rows > 0 && table_host;
    TRY(candidates_build(b, fill, cap_rows, n_rows));
    if (fill && *n_rows > 0) {
        const int64_t nw = *n_rows < cap_rows ? *n_rows : cap_rows;
        CK(cudaMemcpyAsync(table_host, b->cand_rows, sizeof(int32_t) * 3 * (size_t)nw, cudaMemcpyDeviceToHost, b->stream));
        CK(cudaStreamSynchronize(b->stream));
    }
    return 0;
}
extern "C" int gtf_candidates_device(gtf_batch *b, int32_t **rows_dev, int64_t *n_rows)
{
    if (!b || !n_rows || !rows_dev) return fail(GTF_E_ARG, "gtf_candidates_device: null argument");
    CK(cudaSetDevice(b->device));
    TRY(candidates_build(b, true, (int64_t)b->N, n_rows));
    CK(cudaStreamSynchronize(b->stream));
    *rows_dev = b->cand_rows;
    return 0;
}

// ------------------------------------------------------------------------------------------------ pairwise KL diagnostic
// KLDistance between every pair of components of every group (node) with general 3x3 covariances: the inner function of
// the reference's KL-LUT training-data generator (learn_KL_*_model: compute_KL_distance.py:11-21,
// clustering_updated_states_test.py:175-233).  One thread per pair; output order: group by group, (i, j < i) row-major.
__global__ void k_kl_pairs(const double *mean, const double *cov, const int32_t *off, const long long *poff, int n_groups,
                           double *out, long long n_pairs)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    int lo = 0, hi = n_groups;             // group of pair p: last group whose first pair is <= p
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (poff[mid] <= p) lo = mid; else hi = mid;
    }
    const int q = (int)(p - poff[lo]);
    int i = (int)((1.0 + sqrt(1.0 + 8.0 * (double)q)) * 0.5);
    while (i * (i - 1) / 2 > q) i--;
    while ((i + 1) * i / 2 <= q) i++;
    const int j = q - i * (i - 1) / 2;
    const int a = off[lo] + i, b = off[lo] + j;
    out[p] = gtf_kl_general(mean + 3 * (size_t)a, cov + 9 * (size_t)a, mean + 3 * (size_t)b, cov + 9 * (size_t)b);
}
extern "C" int gtf_kl_pairs(int device, const double *mean, const double *cov, const int32_t *off, int32_t n_groups,
                            double *out, int64_t cap, int64_t *n_pairs)
{
    if (!mean || !cov || !off || !n_pairs || n_groups < 0) return fail(GTF_E_ARG, "gtf_kl_pairs: bad argument");
    if (gtf_device_count() <= device || device < 0) return fail(GTF_E_CUDA, "gtf_kl_pairs: no such CUDA device");
    CK(cudaSetDevice(device));
    std::vector<long long> poff((size_t)n_groups + 1, 0);
    for (int gidx = 0; gidx < n_groups; gidx++) {
        const long long n = off[gidx + 1] - off[gidx];
        if (n < 0) return fail(GTF_E_ARG, "gtf_kl_pairs: offsets not monotone");
        poff[gidx + 1] = poff[gidx] + n * (n - 1) / 2;
    }
    const long long np = poff[n_groups];
    *n_pairs = np;
    if (!out || np == 0) return 0;
    if (cap < np) return fail(GTF_E_ARG, "gtf_kl_pairs: output buffer too small");
    const size_t M = (size_t)off[n_groups];
    double *d_mean = nullptr, *d_cov = nullptr, *d_out = nullptr;
    int32_t *d_off = nullptr;
    long long *d_poff = nullptr;
    cudaError_t e = cudaSuccess;
    auto done = [&](int rc) {
        cudaFree(d_mean); cudaFree(d_cov); cudaFree(d_out); cudaFree(d_off); cudaFree(d_poff);
        return rc;
    };
#define CKF(call) do { e = (call); if (e != cudaSuccess) return done(fail(GTF_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e))); } while (0)
    CKF(cudaMalloc((void **)&d_mean, sizeof(double) * 3 * (M ? M : 1)));
    CKF(cudaMalloc((void **)&d_cov, sizeof(double) * 9 * (M ? M : 1)));
    CKF(cudaMalloc((void **)&d_out, sizeof(double) * (size_t)np));
    CKF(cudaMalloc((void **)&d_off, sizeof(int32_t) * ((size_t)n_groups + 1)));
    CKF(cudaMalloc((void **)&d_poff, sizeof(long long) * ((size_t)n_groups + 1)));
    CKF(cudaMemcpy(d_mean, mean, sizeof(double) * 3 * M, cudaMemcpyHostToDevice));
    CKF(cudaMemcpy(d_cov, cov, sizeof(double) * 9 * M, cudaMemcpyHostToDevice));
    CKF(cudaMemcpy(d_off, off, sizeof(int32_t) * ((size_t)n_groups + 1), cudaMemcpyHostToDevice));
    CKF(cudaMemcpy(d_poff, poff.data(), sizeof(long long) * ((size_t)n_groups + 1), cudaMemcpyHostToDevice));
    k_kl_pairs<<<(unsigned)((np + 255) / 256), 256>>>(d_mean, d_cov, d_off, d_poff, n_groups, d_out, np);
    CKF(cudaGetLastError());
    CKF(cudaMemcpy(out, d_out, sizeof(double) * (size_t)np, cudaMemcpyDeviceToHost));
#undef CKF
    return done(0);
}

// ------------------------------------------------------------------------------------------------ stand-alone helpers
// The pure helper functions of the reference (clustering/clustering.py:11-124, extrapolate_merged_states.py:26) for
// callers that use them outside the stages: tiny inputs, one launch each, same device arithmetic as the kernels.
struct SmallBuf {       // device scratch for a handful of doubles, freed on scope exit
    double *d = nullptr;
    ~SmallBuf() { if (d) cudaFree(d); }
};
__global__ void k_pairwise_chi2(const double *sv, const double *cov, const double *node, const double *nbr, int n, GtfGeom g, double *out)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n * n) return;
    const int i = p / n, j = p % n;
    double v = 0.0;
    if (j < i) {        // lower triangle, zeros elsewhere (clustering.py:80-86)
        GtfState si, sj;
        si.a = sv[3 * i]; si.b = sv[3 * i + 1]; si.p00 = cov[9 * i]; si.p01 = cov[9 * i + 1]; si.p11 = cov[9 * i + 4];
        sj.a = sv[3 * j]; sj.b = sv[3 * j + 1]; sj.p00 = cov[9 * j]; sj.p01 = cov[9 * j + 1]; sj.p11 = cov[9 * j + 4];
        si.c = si.tau = si.p22 = sj.c = sj.tau = sj.p22 = 0.0;
        v = gtf_pair_chi2(si, sj, node[0], node[2], node[3], nbr[4 * i], nbr[4 * i + 2], nbr[4 * i + 3], nbr[4 * j], nbr[4 * j + 2],
                          nbr[4 * j + 3], g);
    }
    out[p] = v;
}
extern "C" int gtf_pairwise_chi2(int device, int32_t n, const double *edge_svs, const double *edge_covs, const double *node_coords,
                                 const double *neighbour_coords, double sigma0rz, double sigma0rz2, double endcap_boundary, double *out)
{
    if (n < 0 || !edge_svs || !edge_covs || !node_coords || !neighbour_coords || !out) return fail(GTF_E_ARG, "gtf_pairwise_chi2: bad argument");
    if (gtf_device_count() <= device || device < 0) return fail(GTF_E_CUDA, "gtf_pairwise_chi2: no such CUDA device");
    if (n == 0) return 0;
    CK(cudaSetDevice(device));
    SmallBuf B;
    const size_t nin = (size_t)n * 16 + 4, nout = (size_t)n * n;
    CK(cudaMalloc((void **)&B.d, sizeof(double) * (nin + nout)));
    double *sv = B.d, *cov = sv + 3 * n, *node = cov + 9 * n, *nbr = node + 4, *o = B.d + nin;
    CK(cudaMemcpy(sv, edge_svs, sizeof(double) * 3 * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(cov, edge_covs, sizeof(double) * 9 * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(node, node_coords, sizeof(double) * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(nbr, neighbour_coords, sizeof(double) * 4 * n, cudaMemcpyHostToDevice));
    GtfGeom g;
    g.sigma0xy = 0.0; g.sigma0rz = sigma0rz; g.sigma0rz2 = sigma0rz2; g.endcap = endcap_boundary;
    k_pairwise_chi2<<<(unsigned)((nout + 127) / 128), 128>>>(sv, cov, node, nbr, n, g, o);
    CK(cudaGetLastError());
    CK(cudaMemcpy(out, o, sizeof(double) * nout, cudaMemcpyDeviceToHost));
    return 0;
}
__global__ void k_merge_states(const double *in, double *out) { gtf_merge_general(in, in + 3, in + 12, in + 15, out, out + 3); }
extern "C" int gtf_merge_states(int device, const double *mean1, const double *cov1, const double *mean2, const double *cov2,
                                double *merged_mean, double *merged_cov)
{
    if (!mean1 || !cov1 || !mean2 || !cov2 || !merged_mean || !merged_cov) return fail(GTF_E_ARG, "gtf_merge_states: null argument");
    if (gtf_device_count() <= device || device < 0) return fail(GTF_E_CUDA, "gtf_merge_states: no such CUDA device");
    CK(cudaSetDevice(device));
    SmallBuf B;
    CK(cudaMalloc((void **)&B.d, sizeof(double) * 36));
    double h[24], o[12];
    memcpy(h, mean1, 24); memcpy(h + 3, cov1, 72); memcpy(h + 12, mean2, 24); memcpy(h + 15, cov2, 72);
    CK(cudaMemcpy(B.d, h, sizeof(h), cudaMemcpyHostToDevice));
    k_merge_states<<<1, 1>>>(B.d, B.d + 24);
    CK(cudaGetLastError());
    CK(cudaMemcpy(o, B.d + 24, sizeof(o), cudaMemcpyDeviceToHost));
    memcpy(merged_mean, o, 24); memcpy(merged_cov, o + 3, 72);
    return 0;
}
// parabolic-model seeding of the KL look-up-table training pipeline (learn_KL_parabolic_model/.../utils.py:221-299): thread per
// (node, neighbour) pair
__global__ void k_seed_parabolic(const double *node_xy, const double *nbr_xy, int64_t n, double s0, double sA, double sB, double *sv, double *cov)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s[3], c[9];
    gtf_seed_parabolic(node_xy[2 * i], node_xy[2 * i + 1], nbr_xy[2 * i], nbr_xy[2 * i + 1], s0, sA, sB, s, c);
    for (int k = 0; k < 3; k++) sv[3 * i + k] = s[k];
    for (int k = 0; k < 9; k++) cov[9 * i + k] = c[k];
}
extern "C" int gtf_seed_parabolic_pairs(int device, int64_t n, const double *node_xy, const double *nbr_xy, double sigma0, double sigmaA,
                                        double sigmaB, double *state, double *cov)
{
    if (n < 0 || !node_xy || !nbr_xy || !state || !cov) return fail(GTF_E_ARG, "gtf_seed_parabolic_pairs: bad argument");
    if (gtf_device_count() <= device || device < 0) return fail(GTF_E_CUDA, "gtf_seed_parabolic_pairs: no such CUDA device");
    if (n == 0) return 0;
    CK(cudaSetDevice(device));
    SmallBuf B;
    CK(cudaMalloc((void **)&B.d, sizeof(double) * (size_t)n * 16));
    double *dn = B.d, *db = dn + 2 * n, *ds = db + 2 * n, *dc = ds + 3 * n;
    CK(cudaMemcpy(dn, node_xy, sizeof(double) * 2 * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, nbr_xy, sizeof(double) * 2 * n, cudaMemcpyHostToDevice));
    k_seed_parabolic<<<(unsigned)((n + 127) / 128), 128>>>(dn, db, n, sigma0, sigmaA, sigmaB, ds, dc);
    CK(cudaGetLastError());
    CK(cudaMemcpy(state, ds, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cov, dc, sizeof(double) * 9 * n, cudaMemcpyDeviceToHost));
    return 0;
}
// in: node xyzr[4], neighbour xyzr[4], state[3], block covariance (p00 p01 p11 p22), chi2 cut; out: gtf_edge_result as doubles
__global__ void k_extrapolate_one(const double *in, GtfGeom g, double *out)
{
    const double a = in[8], b = in[9];
    const double dr = in[7] - in[3], dz = in[6] - in[2];
    const double var_ms = gtf_var_ms(a, b, in[4], dr, dz, in[2], g.endcap);       // extrapolate...py:114-124
    const double p11 = in[14] + var_ms;                                           // :127-128 (the caller's matrix is updated too)
    GtfExtrapOut o;
    gtf_extrapolate(in[0], in[1], in[2], in[3], in[4], in[5], in[6], in[7], a, b, in[10], in[11], in[12], p11, in[15], var_ms, in[16], g, o);
    out[0] = o.pass; out[1] = o.chi2; out[2] = var_ms;
    if (o.pass) {
        out[3] = o.lik; out[4] = o.s.a; out[5] = o.s.b; out[6] = o.s.c; out[7] = o.s.tau;
        out[8] = o.s.p00; out[9] = o.s.p01; out[10] = o.s.p11; out[11] = o.s.p22;
    }
}
extern "C" int gtf_extrapolate_validate(int device, const double *node_xyzr, const double *neighbour_xyzr, const double *state,
                                        double *state_cov, double chi2_cut, const gtf_geom *g, gtf_edge_result *out)
{
    if (!node_xyzr || !neighbour_xyzr || !state || !state_cov || !g || !out) return fail(GTF_E_ARG, "gtf_extrapolate_validate: null argument");
    if (gtf_device_count() <= device || device < 0) return fail(GTF_E_CUDA, "gtf_extrapolate_validate: no such CUDA device");
    if (state_cov[2] != 0.0 || state_cov[5] != 0.0 || state_cov[6] != 0.0 || state_cov[7] != 0.0)
        return fail(GTF_E_ARG, "gtf_extrapolate_validate: the covariance must have the block form every stored state has "
                               "(row / column 2 zero off the diagonal: helper.py:423-425, extrapolate_merged_states.py:363-365)");
    CK(cudaSetDevice(device));
    SmallBuf B;
    CK(cudaMalloc((void **)&B.d, sizeof(double) * 32));
    double h[17], o[12] = {0};
    memcpy(h, node_xyzr, 32); memcpy(h + 4, neighbour_xyzr, 32); memcpy(h + 8, state, 24);
    h[11] = state_cov[0]; h[12] = state_cov[1]; h[14] = state_cov[4]; h[15] = state_cov[8]; h[13] = 0.0; h[16] = chi2_cut;
    CK(cudaMemcpy(B.d, h, sizeof(h), cudaMemcpyHostToDevice));
    k_extrapolate_one<<<1, 1>>>(B.d, geom_of(g), B.d + 20);
    CK(cudaGetLastError());
    CK(cudaMemcpy(o, B.d + 20, sizeof(o), cudaMemcpyDeviceToHost));
    memset(out, 0, sizeof(*out));
    out->pass = (int32_t)o[0]; out->chi2 = o[1]; out->var_ms = o[2];
    state_cov[4] += o[2];             // `merged_cov[1, 1] += var_ms` mutates the caller's matrix (quirk 2)
    if (out->pass) {
        out->likelihood = o[3];
        out->state[0] = o[4]; out->state[1] = o[5]; out->state[2] = o[6]; out->tau = o[7];
        out->cov[0] = o[8]; out->cov[1] = o[9]; out->cov[2] = o[10]; out->cov[3] = o[11];
    }
    return 0;
}
