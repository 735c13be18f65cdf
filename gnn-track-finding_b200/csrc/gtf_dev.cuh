// gtf_dev.cuh -- device-side view of an event batch, op codes of the per-node program, batch object.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/gtf.h"
#include "gtf_math.cuh"

enum {
#define X(name, type, ext) GTF_F_##name,
    GTF_FIELDS(X)
#undef X
        GTF_NFIELDS
};

// every array of gtf_fields.h as a typed device pointer + derived / scratch arrays
struct DevBatch {
    int N, E, S, n_tiles;
#define X(name, type, ext) type *name;
    GTF_FIELDS(X)
#undef X
    // derived topology
    const int32_t *tile_begin;   // [n_tiles + 1] node ranges, <= TILE_SLOTS slots and <= TILE_NODES nodes each
    int32_t *sub_nalive;         // [S] alive nodes per sub-graph ("len(subGraph.nodes()) == 1: continue")
    uint8_t *node_ok;            // [N] bit0: alive and sub-graph in play, bit1: sub-graph has != 1 nodes (derived)
    int all_alive;               // no node removed yet: every edge exists, the alive[] gathers can be skipped
    // per-source multiple-scattering prefix (extrapolate_merged_states.py:127-128, quirk 2)
    double *slot_p11;            // [E] merged_cov[1,1] as seen by this edge
    double *slot_vms;            // [E] var_ms of this edge
    double *node_p11tot;         // [N] merged_cov[1,1] after all successors
    // shadow buffers of the iteration: accumulated merged_cov[1,1] of the next state (ping-pong with m_p11, quirk 2);
    // has_merged flags of an uncommitted pass
    uint8_t *has_merged_nx;
    double *m_p11_nx;
    unsigned long long *counters; // [GTF_NCOUNTERS]
    gtf_near_rec *near_log;       // [GTF_NEAR_LOG] decisions within 1e-9 relative of their threshold (counters[CNT_NEAR] counts them)
};
#define GTF_NEAR_LOG 256
#define GTF_LOOP_BURST 32      // iterations of gtf_iterate queued between two host read-backs

// ---- packed iteration layout (gtf_iter.cuh): per-slot records + bitmaps, built from / written back to the SoA
// fields by k_pack_* / k_unpack_slots.  The SoA arrays stay the exchange format of the C-ABI.
struct __align__(32) MetaRec {   // the part of an updated_track_states entry the re-weighting touches: one 32 B sector,
    double w, lik, prior;        // always written whole (a partial-sector write costs a DRAM read-modify-write)
    double ew;                   // G[src][dst]['mixture_weight'] (helper.py:180)
};
struct __align__(8) TagRec {     // rarely changing part of the entry: written only when a value changes
    int32_t rank;                // dict insertion stamp
    int8_t side, pad;            // helper.py:129-139 'side' (0 none, 1 left, 2 right)
    int16_t lrn;                 // lr_layer_norm: -1 untouched since packing (SoA value stands), 0 NaN, k > 0 the integer norm
};
struct __align__(16) GeoRec {    // static per-slot view of the source hit
    double sx;                   // x of the source hit
    int32_t lay, src;            // its layer id; its node index (-1: ghost slot)
};
struct __align__(32) AuxRec { GeoRec g; TagRec t; int32_t pad[2]; };
struct __align__(32) NodeXYZR { double x, y, z, r; };
// per out-edge (successor order), static between per-stage calls: what a message needs besides the source's merged state
struct __align__(32) OutRec {
    double sin_t, xk, rdz;       // sin(theta) of the segment, x of the destination hit, |dr| / |dz|  (Highland term, extrapolate...py:114-124)
    double w;                    // mixture weight the message carries = the source's seed entry for this neighbour (:384); GTF_NO_TSE bits: none
};
#define GTF_NO_TSE_BITS 0x7ff8dead00000001ll
struct __align__(16) SrcRec {     // static per-source record of k_send: one 16 B read instead of two arrays
    double z;                     // z of the hit (end-cap side of the Highland term)
    int32_t off, pad;             // out_off[u]
};
struct __align__(32) MergedRec {  // merged_state / merged_cov / merged_prior of a node
    double a, b, c, p00, p01, p22, prior;
    double cl_p11;                // merged_cov[1,1] of the cluster formed at the node's last evaluation, NaN: none.  (The live
};                                // value accumulates multiple scattering in the m_p11 ping-pong pair: quirk 2.)

static_assert(sizeof(MetaRec) == 32 && sizeof(AuxRec) == 32 && sizeof(NodeXYZR) == 32 && sizeof(OutRec) == 32 && sizeof(MergedRec) == 64,
              "packed records are whole 32 B sectors");

enum { HV_BINS = 4, PK_MSG = 0, PK_HV0 = 1, PK_BIG = 5, PK_MISSING = 6, PK_FORCE = 7,
       PK_STOP = 8,   // the committed loop of gtf_iterate has converged on the device: queued iterations do nothing
       PK_DONE = 9,   // iterations of that loop that ran
       PK_SPARSE = 10, // the loop's out-edges have been compacted to the active ones: k_send_sparse sends, k_send does nothing
       PK_CLIST = 11, // entries of the compacted list handed out so far
       PK_NCOUNTS = 12 };

struct DevPack {
    // static, derived from the topology and the hit coordinates
    int32_t *out_dst;            // [E] out-CSR order: destination node
    OutRec *orec;                // [E] out-CSR order
    AuxRec *aux;                 // [E] geometry + tag
    NodeXYZR *xyzr;              // [N]
    SrcRec *srec;                // [N + 1]
    MergedRec *mrec, *mrec_nx;   // [N] merged state of every node (64 B: two whole sectors); _nx: shadow for uncommitted passes
    double2 *mab;                // [N] compact copy of the COMMITTED merged (a, b): what k_send needs of a source, contiguous
                                 // per tile (a 16 B gather out of the 64 B records costs a 64 B DRAM fetch per source)
    int all_exist;               // every slot is an existing edge (no ghost slot, no removed node): set on the device from counts[PK_MISSING]
    // mutable slot state
    uint32_t *act, *act_nx, *pres, *exists; // bitmaps over slots, 2 zero words of padding
    uint32_t *pres0;             // presence bitmap as it was when the iteration started (an entry not in it is new)
    double *state;               // [E][8] a b c tau p00 p01 p11 p22
    MetaRec *meta;               // [E]
    // message list of one iteration, source-major (a source's messages are contiguous, in successor order)
    int4 *msg_desc;              // [E] (slot, source, destination, -); slot bit 31: the source has no seed entry for this neighbour
    double *msg_w;               // [E] mixture weight carried by the message (extrapolate...py:384)
    double *msg_p11, *msg_vms;   // [E] merged_cov[1,1] as the edge sees it (quirk 2), its multiple-scattering term
    // the ACTIVE out-edges of every source, compacted inside a committed loop once few are left (activation only ever
    // falls inside a loop): k_compact_out -> k_send_sparse
    int4 *c_edge;                // [E] (slot, destination, out-edge index, source), a source's edges contiguous and in successor order
    int2 *c_rng;                 // [N] (first entry, entries) of the source in c_edge
    // nodes found without an active in-edge (k_node2): skipped without a scan until the next forced pass
    uint8_t *node_static;        // [N]
    double *node_rest;           // [N] their merged_cov[1,1] to restore every iteration (quirk 2), NaN: none
    int32_t *hv_list;            // [(HV_BINS + 1) * N] cooperative nodes binned by dict size: <=4, <=8, <=16, <=32, more
    int *counts;                 // [PK_NCOUNTS] messages, 4 bins, big, missing slots, 'evaluate every node' flag, loop stop / done
};

enum {
    CNT_MERGED = 0, CNT_DEACT, CNT_SENT, CNT_GATED, CNT_RWOFF, CNT_ACTIVE, CNT_CHANGED, CNT_REFERR, GTF_NCOUNTERS,
    CNT_NEAR = GTF_NCOUNTERS, GTF_NCOUNTERS_ALL     // (CNT_NEAR is bumped in global memory directly: a rare event)
};
// in-kernel bounds checks of the debug build (compute-sanitizer is not available on the pool): a violated condition sets a
// status bit that every stats read-back returns; the release build compiles them away
#ifdef GTF_DEBUG_BOUNDS
#define GTF_BOUND(B, cond) do { if (!(cond)) atomicOr(&(B).counters[CNT_REFERR], (unsigned long long)GTF_STATUS_BOUNDS); } while (0)
#else
#define GTF_BOUND(B, cond) ((void)0)
#endif
// a decision `value (<|<=) threshold` was just taken: note it when it is a boundary-flip candidate
#ifdef GTF_NEAR_NOINLINE      // (measured: a real call costs the hot kernels more than the inlined rare path, 0.138 vs 0.134 ms in k_exec)
#define GTF_NEAR_ATTR __noinline__
#else
#define GTF_NEAR_ATTR __forceinline__
#endif
__device__ GTF_NEAR_ATTR void near_log(unsigned long long *counters, gtf_near_rec *log, int kind, int index, double value, double threshold)
{
    const unsigned long long k = atomicAdd(&counters[CNT_NEAR], 1ull);
    if (k < GTF_NEAR_LOG) {
        gtf_near_rec r;
        r.kind = kind; r.index = index; r.value = value; r.threshold = threshold;
        log[k] = r;
    }
}
__device__ __forceinline__ void near_note(const DevBatch &B, int kind, int index, double value, double threshold)
{
    if (fabs(value - threshold) <= GTF_NEAR_RTOL * fabs(threshold)) near_log(B.counters, B.near_log, kind, index, value, threshold);
}

// per-node program executed by the tile kernel
enum { PG_ACT = 0, PG_PRES = 1, PG_REC = 2, PG_NODE = 3, PG_N = 4 };
enum { OP_END = 0, OP_E, OP_PRIOR, OP_RW, OP_CLUSTER, OP_DEGREE, OP_WEIGHTS, OP_POP };

// write-back masks
enum {
    WB_ACTIVE = 1, WB_PRESENT = 2, WB_STATE = 4, WB_PRIOR = 8, WB_W = 16, WB_UTSX = 32 /* lik, lrn, side, rank */,
    WB_EDGEW = 64, WB_COUNT_ACTIVE = 256
};

struct Prog {
    int ops[12];
    int key;          // working dict: GTF_KEY_TSE / GTF_KEY_UTS
    int wb;           // WB_* mask
    int use_lut;
    int pre_passes;   // (prior, reweight) passes before the clustering in the packed node kernels: 2 = the fused iteration
    double chi2_cut, cl_chi2, cl_kl, rw_thr;
    double lut[28];
};

#ifndef GTF_TILE_SLOTS
#define GTF_TILE_SLOTS 768
#endif
#ifndef GTF_TILE_NODES
#define GTF_TILE_NODES 255
#endif
#ifndef GTF_TILE_THREADS
#define GTF_TILE_THREADS 256
#endif
#define GTF_MAXD 15
#ifndef GTF_TILE_MINB
#define GTF_TILE_MINB 2
#endif

// a captured iteration (issue_iteration in gtf_b200.cu) and the parameters it was captured with
struct IterGraph {
    cudaGraphExec_t exec;
    Prog P;
    GtfGeom g;
    int record_chi2, n_stiles;
    const void *stile;
    int N, E, S, n_tiles;      // batch shape the kernel arguments (sizes, grids) were captured for
};

struct gtf_batch {
    int N, E, S, device;
    int capN, capE, capS;      // allocated capacity (gtf_batch_create); gtf_batch_load_events may load smaller batches
    int topo_gen;              // incremented whenever N / E / S or a tile table changes
    int32_t *h_tiles;          // pinned staging for the two tile tables
    cudaEvent_t ev_tiles;      // the upload out of h_tiles
    bool ev_tiles_used;
    cudaStream_t stream, stream2;
    void *f[GTF_NFIELDS];
    DevBatch d;
    bool finalized, derived_dirty;
    unsigned long long *n_dead;
    int n_tiles;
    int32_t *tile_begin;
    int n_stiles;
    int32_t *stile_begin;      // k_send tiles as int4 (first source, sources, first out-edge, out-edges): whole sources,
                               // <= GTF_SEND_SRCS sources and <= GTF_SEND_EDGES out-edges
    unsigned long long *h_counters; // pinned
    unsigned long long *loop_stats, *h_loop_stats; // [GTF_LOOP_BURST][GTF_NCOUNTERS_ALL] per-iteration counters of a queued loop (device, pinned)
    int *h_loop_done;          // pinned
    int64_t dev_bytes;
    // extraction scratch
    uint8_t *accepted_total;   // [N] nodes accepted by any gtf_extract so far
    int32_t *cand_root;        // [N] root (candidate id) of accepted nodes
    uint8_t *sub_has_inactive; // [S]
    int32_t *sub_first;        // [S]
    void *sort_tmp;
    size_t sort_tmp_bytes;
    int32_t *sort_keys, *sort_vals, *sort_keys2, *sort_vals2;
    double *pv_xy, *pv_zr;     // [N]
    uint8_t *acc_now;          // [N]
    int32_t *tags_a, *tags_b;  // [N]
    int32_t *cand_rows;        // [3 N] candidate table staging (allocated on first use)
    int n_sm;
    // packed iteration layout
    DevPack k;
    bool exists_stale;         // alive flags changed since the existing-edge bitmap was built
    bool pack_static_stale;    // topology / coordinates / seed weights changed since the static part was built
    bool pack_out_stale;       // only the per-out-edge records (carried seed weights) are out of date
    bool pack_stale[4], soa_stale[4]; // per group (PG_ACT, PG_PRES, PG_REC, PG_NODE): which side holds the newer state
    cudaStream_t stream3;
    cudaEvent_t ev_fork2, ev_join2, ev_join3;
    cudaEvent_t evk[6];
    bool force_pending;        // the packed state changed from outside: the next committed iteration evaluates every node
    int force_dev;             // value of counts[PK_FORCE] on the device (-1 unknown)
    bool have_last_prog;       // thresholds / geometry of the last committed iteration (a change re-evaluates every node)
    Prog last_prog;
    GtfGeom last_geom;
    bool use_graph;            // replay the iteration from a CUDA graph (GTF_GRAPH=0: plain launches)
    bool fused_sx;             // k_send + k_exec as the one warp-specialised kernel k_sx (GTF_FUSED_SX=1)
    int parity;                // which half of the ping-pong pairs (act / act_nx, m_p11 / m_p11_nx) is current
    IterGraph graphs[2][2][2]; // [committed][parity][sparse-send variant]
    double t_k[5];             // send, exec, node, heavy, (spare)
    // optional per-kernel timing of the iteration (CUDA events on the batch stream)
    bool timing;
    int t_count;
    long long launches;        // kernels launched by packed iterations so far (graph replays count their kernel nodes)
    int launches_per_iter;
};
