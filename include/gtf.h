/* gtf.h -- C-ABI of libgtf_b200.so: the B200 (sm_100a) message-passing hot path of
 * nishalad95/GNN-track-finding behind plain C entry points.
 *
 * The reference has no FFI; its de-facto boundary is "Python function taking a list of nx.DiGraph (or a
 * directory of {i}_subgraph.gpickle) and mutating it" (SURVEY.md §8b).  Each entry point below names the
 * reference function (path relative to /root/reference/src) whose effect on the graph attributes it
 * reproduces on the flat layout of gtf_fields.h.  INTEGRATION.md shows the ctypes stub a maintainer of
 * the reference would add.
 *
 * Conventions: every function returns 0 on success or a negative GTF_E_* code (message via
 * gtf_last_error()).  `host` pointers are plain host memory (pinned or not); the library copies.  A batch
 * owns its device buffers and one CUDA stream; calls on one batch are not re-entrant.  There is NO CPU
 * fallback: without a CUDA device every compute call fails with GTF_E_CUDA.
 */
#ifndef GTF_H
#define GTF_H
#include <stdint.h>
#include "gtf_fields.h"

#ifdef __cplusplus
extern "C" {
#endif

#define GTF_ABI_VERSION 3

#define GTF_E_CUDA (-1)    /* CUDA runtime error / no device */
#define GTF_E_ARG (-2)     /* bad argument */
#define GTF_E_STATE (-3)   /* batch not finalized / topology missing */
#define GTF_E_DEGREE (-4)  /* a node has more in-slots than one tile holds (GTF_TILE_SLOTS) */
#define GTF_E_NOMEM (-5)

/* bits of gtf_stats.ref_errors: points where the reference itself would raise */
#define GTF_REF_EMPTY_MIN 1 /* np.min([])                clustering/clustering.py:116,120  ValueError        */
#define GTF_REF_NAN_INDEX 2 /* list.index(nan)           clustering/clustering.py:117      ValueError        */
#define GTF_REF_ZERO_DIV 4  /* 1/len({})                 utilities/helper.py:90            ZeroDivisionError */
#define GTF_REF_KEY 8       /* G[u][v] on a removed edge utilities/helper.py:131,138       KeyError          */
#define GTF_REF_NO_TSE 16   /* missing seed entry        extrapolate/extrapolate_merged_states.py:384 KeyError */
/* not a reference error: gtf_extract met a component of more than 64 nodes that could still be one-hit-per-layer (a
 * detector with > 62 distinct (volume, layer) ids); the call fails with GTF_E_DEGREE instead of skipping it silently */
#define GTF_STATUS_CAND_OVERFLOW 32
/* debug builds only (nvcc -DGTF_DEBUG_BOUNDS, tools/debug_bounds.sh): an index computed inside a kernel left its array */
#define GTF_STATUS_BOUNDS 64

typedef struct gtf_batch gtf_batch;

typedef struct {
    double sigma0xy, sigma0rz, sigma0rz2, endcap_boundary; /* run_gnn_trackml_mod.sh:11-21 */
} gtf_geom;

typedef struct {
    int64_t nodes_merged;       /* nodes that got a (new) merged state        clustering.py:291-294 */
    int64_t edges_deactivated;  /* un-absorbed components switched off        clustering.py:311-321 */
                                /* (gtf_iterate counts the nodes it evaluated: from the second committed iteration on, nodes
                                   without an active in-edge are skipped -- their evaluation would change nothing) */
    int64_t edges_sent;         /* extrapolate_validate calls                 extrapolate...py:433  */
    int64_t edges_gated;        /* chi2 > cut                                 extrapolate...py:393  */
    int64_t edges_reweight_off; /* weight < threshold                         helper.py:186-187     */
    int64_t active_edges;       /* existing edges with activated == 1 after the call */
    int64_t active_changed;     /* edges whose flag changed in the call (convergence test) */
    int64_t ref_errors;         /* OR of GTF_REF_* */
    int64_t near_threshold;     /* decisions taken within GTF_NEAR_RTOL (1e-9 relative) of their threshold in this call: the
                                   "boundary flip candidates" of SURVEY.md 8d; the records are read with
                                   gtf_batch_near_threshold */
} gtf_stats;

/* one decision whose value lies within 1e-9 relative of its threshold (an fp64 implementation with a different operation
 * order may take it the other way) */
#define GTF_NEAR_RTOL 1e-9
#define GTF_NEAR_GATE 0         /* chi2 <= chi2CutFactor     extrapolate_merged_states.py:298   index = slot */
#define GTF_NEAR_REWEIGHT 1     /* reweight < 0.1            utilities/helper.py:186            index = slot */
#define GTF_NEAR_CLUSTER_CHI2 2 /* smallest_dist < chi2_thr  clustering/clustering.py:228       index = node */
#define GTF_NEAR_CLUSTER_KL 3   /* smallest_dist < KL_thr    clustering/clustering.py:261       index = node */
typedef struct {
    int32_t kind, index;
    double value, threshold;
} gtf_near_rec;

typedef struct {
    double chi2_cut;            /* extrapolation gate            run_gnn_trackml_mod.sh:28  (2.0)   */
    double cluster_chi2, cluster_kl; /* cluster on updated states run_gnn_trackml_mod.sh:112 (1000, 100) */
    double reweight_threshold;  /* helper.py:145 (0.1) */
    const double *kl_lut;       /* host pointer to 28 kl_max values or NULL (scalar threshold) */
    int32_t record_chi2;        /* != 0: also store every message's gate chi2 in `uts_chi2` -- a diagnostic the reference
                                   only appends to a CSV (extrapolate_merged_states.py:150-292), not part of the graph
                                   state; off by default in the fused iteration (one scattered 8 B write per message) */
} gtf_iter_params;

int gtf_abi_version(void);
const char *gtf_last_error(void);
int gtf_device_count(void);

/* ---- batch life cycle and data movement ------------------------------------------------------- */
int gtf_batch_create(int32_t n_nodes, int32_t n_slots, int32_t n_subgraphs, int device, gtf_batch **out);
int gtf_batch_destroy(gtf_batch *b);
int gtf_field_count(void);
const char *gtf_field_name(int field_id);
int gtf_field_id(const char *name);
int64_t gtf_field_bytes(const gtf_batch *b, int field_id);
/* host -> device / device -> host copies of one array of gtf_fields.h (async on the batch stream;
 * download synchronises before returning) */
int gtf_batch_upload(gtf_batch *b, int field_id, const void *host);
int gtf_batch_download(gtf_batch *b, int field_id, void *host);
/* same without the synchronisation (host must be pinned; call gtf_batch_sync before reading it) */
int gtf_batch_download_async(gtf_batch *b, int field_id, void *host);
int gtf_batch_device_ptr(gtf_batch *b, int field_id, void **dptr);
/* after the topology arrays (in_off, in_src, slot_dst, out_off, out_slot, rev_slot, sub, sub_off,
 * sub_state, alive) are uploaded: builds the node tiles and per-sub-graph counters */
int gtf_batch_finalize(gtf_batch *b);
/* ---- device-side ingest: a whole batch of freshly converted events in ONE call --------------------------------------
 * What trackml_mod/event_conversion.py:40-112 + utilities/helper.py:465-520 (construct_graph) hand to the stages: hits and
 * the directed graph, nothing else.  The host supplies only the arrays that DEFINE the events (44 B per hit + 8 B per
 * directed edge; pinned memory makes the copies asynchronous); slot_dst / rev_slot are derived on the device, every
 * node is alive, every sub-graph in play, all state arrays start absent (NaN / -1 / 0), the tiles are built, and the batch
 * is finalized.  n_nodes / n_slots / n_subgraphs may be smaller than the capacity given to gtf_batch_create, so one batch
 * object (and its device memory) serves a stream of batches.  Follow with gtf_seed_all, gtf_cluster, gtf_iterate, ... */
typedef struct {
    int32_t n_nodes, n_slots, n_subgraphs;
    const double *x, *y, *z, *r;       /* [N] hit coordinates                         GNN_Measurement.py:1-9            */
    const int32_t *layer, *volume;     /* [N] in_volume_layer_id, volume_id           utilities/helper.py:497-508       */
    const int32_t *sub;                /* [N] sub-graph of the node                   event_conversion.py:76-84         */
    const int32_t *sub_off;            /* [S+1] node range of each sub-graph                                             */
    const int32_t *sub_event;          /* [S] event id of each sub-graph (candidate table)                              */
    const int32_t *in_off, *in_src;    /* [N+1], [E] in-CSR by destination, dict (insertion) order  helper.py:280,375    */
    const int32_t *out_off, *out_slot; /* [N+1], [E] out-CSR by source, successor order   extrapolate...py:430          */
} gtf_events;
int gtf_batch_load_events(gtf_batch *b, const gtf_events *ev);
/* the boundary-flip candidates of the most recent stage call / iteration on this batch: up to `cap` records into `out`
 * (the library keeps the first 256), *n = how many the call counted */
int gtf_batch_near_threshold(gtf_batch *b, gtf_near_rec *out, int cap, int64_t *n);
int gtf_batch_sync(gtf_batch *b);
int gtf_batch_stream(gtf_batch *b, void **cuda_stream);
int64_t gtf_batch_device_bytes(const gtf_batch *b);
/* kernels launched by gtf_iterate / gtf_iterate_dry on this batch so far (a CUDA-graph replay counts its kernel nodes) */
int64_t gtf_batch_iteration_launches(const gtf_batch *b);

/* ---- per-stage entry points (same effect as the reference function named) ---------------------- */
/* utilities/helper.py:238-452 compute_track_state_estimates (slot order supplies the neighbour order) */
int gtf_seed(gtf_batch *b, const gtf_geom *g);
/* event_conversion.py:87-96 in one call: gtf_seed + initialize_edge_activation + compute_prior_probabilities(seeds) +
 * compute_mixture_weights(seeds) + node degrees (the last three as one kernel launch) */
int gtf_seed_all(gtf_batch *b, const gtf_geom *g, gtf_stats *st);
/* gtf_seed_all followed by gtf_cluster('track_state_estimates', ...) -- event_conversion.py:87-96 + iteration 1 of
 * run_gnn_trackml_mod.sh:89 -- as ONE pass of the packed node kernels over the freshly seeded dicts */
int gtf_seed_cluster(gtf_batch *b, const gtf_geom *g, double chi2_threshold, double kl_threshold, const double *kl_lut,
                     gtf_stats *st);
/* utilities/helper.py:24-25 initialize_edge_activation */
int gtf_initialize_edge_activation(gtf_batch *b);
/* utilities/helper.py:30-63 compute_prior_probabilities(GraphList, key) */
int gtf_compute_prior_probabilities(gtf_batch *b, int key);
/* utilities/helper.py:76-94 compute_mixture_weights(GraphList, key) */
int gtf_compute_mixture_weights(gtf_batch *b, int key, gtf_stats *st);
/* utilities/helper.py:67-73 query_node_degree_in_edges for every node -> `degree` */
int gtf_query_node_degree(gtf_batch *b);
/* clustering/clustering.py:149-376 cluster(): per-node chi2/KL clustering, simultaneous deactivation,
 * degree, mixture weights, priors.  kl_lut: host pointer to 28 doubles (LUT mode) or NULL. */
int gtf_cluster(gtf_batch *b, int key, double chi2_threshold, double kl_threshold, const double *kl_lut,
                const gtf_geom *g, gtf_stats *st);
/* extrapolate/extrapolate_merged_states.py:406-447 message_passing() */
int gtf_message_passing(gtf_batch *b, double chi2_cut, const gtf_geom *g, gtf_stats *st);
/* utilities/helper.py:143-200 reweight(subGraphs, 'updated_track_states') */
int gtf_reweight(gtf_batch *b, int key, double threshold, gtf_stats *st);
/* extrapolate/extrapolate_merged_states.py:552-567 main(): message_passing, (prior, reweight) x2, degree */
int gtf_extrapolate_stage(gtf_batch *b, double chi2_cut, const gtf_geom *g, gtf_stats *st);
/* update/remove_state_metadata.py:31-53 */
int gtf_remove_state_metadata(gtf_batch *b, gtf_stats *st);

/* ---- fused iteration: [message_passing, prior, reweight, prior, reweight, cluster(updated states), degree, weights, priors] --- */
/* Runs on a packed copy of the mutable state (activation / presence bitmaps, 64 B state + 32 B weight records per dict
 * entry, 64 B merged-state records per node) that the library builds from the fields on entry and writes back lazily
 * when a field is downloaded or a per-stage entry point is called.  Per iteration: k_send (message list + per-source
 * multiple-scattering prefix), k_exec (extrapolate, chi2 gate, Kalman update), k_node2 (nodes holding <= 2 components:
 * priors, reweight x2, prune), k_hv<4|8|16|32> / k_big (>= 3 components: the same + pairwise chi2 + greedy KL merge).
 * The launch sequence is replayed from a CUDA graph (environment GTF_GRAPH=0: plain launches).
 * Runs `max_iter` iterations or stops early when an iteration leaves the active-edge bitmap unchanged (SURVEY.md §8d
 * "converged").  stats[i] receives iteration i's counters (may be NULL); *n_done the number of iterations run.
 * With stats == NULL and stop_when_converged == 0 the call does not synchronise (no counter read-back). */
int gtf_iterate(gtf_batch *b, const gtf_iter_params *p, const gtf_geom *g, int max_iter, int stop_when_converged,
                gtf_stats *stats, int *n_done);
/* the same iteration, ONE pass, NOT committed: reads the current state, rewrites the dict entries in place (with the
 * same values on every call) and sends activation flags / merged states / accumulated p11 to shadow buffers, so the
 * next call does identical work (benchmark / profiling entry point; the stats are those of a committed pass).
 * Side effects that DO persist: the dict entries written by the pass (presence bits, state / weight records, insertion
 * stamps `uts_rank`, `uts_next`, `has_uts`) -- exactly what a committed pass would write, and what the next pass
 * overwrites with the same values; the activation flags, merged states and accumulated merged_cov[1,1] do not change. */
int gtf_iterate_dry(gtf_batch *b, const gtf_iter_params *p, const gtf_geom *g, gtf_stats *st);

/* per-kernel timing of the iteration (CUDA events recorded on the batch stream): enable != 0 resets the
 * accumulators; gtf_batch_timing returns averages over the calls since: prefix_ms = k_begin + k_send,
 * tile_ms = k_exec + k_node2, heavy_ms = k_hv<*> + k_big */
int gtf_batch_set_timing(gtf_batch *b, int enable);
int gtf_batch_timing(gtf_batch *b, double *prefix_ms, double *tile_ms, double *heavy_ms, int *count);
/* per-kernel averages of the packed pipeline: ms[0..3] = k_send, k_exec, k_node2, cooperative kernels (k_hv<*>, k_big) */
int gtf_batch_timing_kernels(gtf_batch *b, double *ms, int n_ms, int *count);

/* ---- candidate extraction ------------------------------------------------------------------- */
/* extract/extract_track_candidates.py:332-346 CCA: weakly connected components over active edges ->
 * `label` (smallest node index of the component; -1 for removed nodes) */
int gtf_components(gtf_batch *b);
/* extract/extract_track_candidates.py:402-467: components -> one-hit-per-layer / close-pair merge /
 * KF fit p-value gate -> accepted nodes removed, sub-graph states updated.
 * accepted (u8[N]), pval_xy / pval_zr (f64[N], at the component's root index) are optional host outputs.
 * A candidate holds at most 64 nodes (one hit per (volume, layer) id plus two close pairs: enough for any detector with
 * <= 62 such ids, TrackML has 48); larger components are rejected after their layer duplicates are proven, else the call
 * fails with GTF_E_DEGREE. */
int gtf_extract(gtf_batch *b, const gtf_geom *g, double pval_cut, int numhits, double sep3d, double merge_dist,
                int32_t *n_accepted, uint8_t *accepted, double *pval_xy, double *pval_zr);
/* tag_propagation/tag_propagation.py:64-164: Jacobi max-label propagation from lower-radius successors until
 * the fraction of flipped tags <= threshold.  tags: host i32[N], in = initial tag, out = final tag. */
int gtf_tag_propagate(gtf_batch *b, double threshold, int32_t *tags, int max_sweeps, int *n_sweeps);
/* rows (event_id, candidate_id = root node index, node index) of every node accepted so far, sorted by (event, candidate,
 * node) on the device (radix sort by candidate id), table copied to host.  Returns the row count in *n_rows (may exceed
 * cap; only cap rows written); table_host == NULL: count only. */
int gtf_candidates(gtf_batch *b, int32_t *table_host, int64_t cap_rows, int64_t *n_rows);
/* the same table left on the device (valid until the next gtf_candidates* call on this batch): *rows_dev points at
 * n_rows x 3 int32 -- what a multi-GPU driver hands to NCCL for the final gather */
int gtf_candidates_device(gtf_batch *b, int32_t **rows_dev, int64_t *n_rows);

/* ---- diagnostic: pairwise KL between the components of every group (general 3x3 covariances) ------------------- */
/* The inner function of the reference's KL-threshold LUT training-data generator (learn_KL_linear_model /
 * learn_KL_parabolic_model: compute_KL_distance.py:11-21, clustering_updated_states_test.py:175-233; KLDistance with the
 * element-wise trace of clustering/clustering.py:90-94).  mean: host f64[M][3], cov: host f64[M][9], off: host i32[G+1]
 * (components of group g are off[g]..off[g+1]).  Writes KL(i, j) for j < i, group by group, to out (host f64[*n_pairs];
 * out == NULL: count only).  Needs no batch. */
int gtf_kl_pairs(int device, const double *mean, const double *cov, const int32_t *off, int32_t n_groups, double *out,
                 int64_t cap, int64_t *n_pairs);

/* ---- the reference's stand-alone helper functions (tiny inputs, one launch each; same device arithmetic as the kernels) --- */
/* clustering/clustering.py:80-86 calc_pairwise_distances_chi2 (= :11-78 mahalanobis_distance for every pair j < i):
 * edge_svs f64[n][3], edge_covs f64[n][3][3] (only the symmetric [0:2, 0:2] block enters), node_coords (x, y, z, r),
 * neighbour_coords f64[n][4]; out f64[n][n], lower triangle filled, zeros elsewhere.  Needs no batch. */
int gtf_pairwise_chi2(int device, int32_t n, const double *edge_svs, const double *edge_covs, const double *node_coords,
                      const double *neighbour_coords, double sigma0rz, double sigma0rz2, double endcap_boundary, double *out);
/* clustering/clustering.py:97-105 merge_states for general 3x3 covariances (row-major f64[9]) */
int gtf_merge_states(int device, const double *mean1, const double *cov1, const double *mean2, const double *cov2,
                     double *merged_mean, double *merged_cov);
/* learn_KL_parabolic_model/src/generate_training_data/utils.py:221-299 compute_track_state_estimates (with :197-218
 * rotate_track) -- the PARABOLIC-model seeding behind the KL look-up-table training data, not the seeding of the reconstruction
 * (gtf_seed): for each of the n (node, neighbour) pairs (xy f64[n][2] each) the `edge_state_vector` f64[n][3] and the full
 * `edge_covariance` f64[n][3][3] = H^-1 diag(sigma0^2, sigmaA^2, sigmaB^2) H^-T (the reference: 4.0, 0.1, 0.1).  Needs no batch. */
int gtf_seed_parabolic_pairs(int device, int64_t n, const double *node_xy, const double *nbr_xy, double sigma0, double sigmaA,
                             double sigmaB, double *state, double *cov);
/* extrapolate/extrapolate_merged_states.py:26-402 extrapolate_validate for ONE edge node -> neighbour: state (a, b, c) and its
 * covariance (row-major f64[9], block form) at `node`; like the reference it adds the multiple-scattering variance to
 * state_cov[1][1] IN PLACE (:127-128) before extrapolating.  pass == 0: chi2 > cut, the caller deactivates the edge (:393). */
typedef struct {
    int32_t pass, pad;
    double chi2, var_ms, likelihood;
    double state[3];   /* updated (a, b, c) = edge_state_vector                       */
    double tau;        /* joint_vector = (state[0], state[1], tau)                    */
    double cov[4];     /* p00 p01 p11 p22 of edge_covariance = joint_vector_covariance */
} gtf_edge_result;
int gtf_extrapolate_validate(int device, const double *node_xyzr, const double *neighbour_xyzr, const double *state,
                             double *state_cov, double chi2_cut, const gtf_geom *g, gtf_edge_result *out);

#ifdef __cplusplus
}
#endif
#endif
