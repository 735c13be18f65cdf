/* gtf_oracle.h -- CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A literal, single-threaded (optionally events-parallel) C restatement of the reference's
 * Python hot path over the flat layout of include/gtf_fields.h.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it; the product path
 * (gnn-track-finding_b200/, the CUDA C-ABI) never does.
 *
 * Pinning: every function here is checked against outputs of the UNMODIFIED reference run in the
 * build container (tests/golden/make_golden.py -> tests/golden/*.npz) and, for KL + parabolic
 * seeding, against the reference's one shipped known-answer file.  The Kalman
 * predict/update arithmetic follows filterpy 1.4.5's published algorithm; filterpy itself is absent
 * and unpinned upstream, so that boundary is "parity unpinned" by reference tests.
 */
#ifndef GTF_ORACLE_H
#define GTF_ORACLE_H
#include <stdint.h>
#include "../include/gtf_fields.h"

typedef struct {
    int32_t N, E, S, pad_;
#define X(name, type, ext) type *name;
    GTF_FIELDS(X)
#undef X
} gtfo_arrays;

/* error bits returned (OR-ed) by the stage functions: the reference would raise at these points */
#define GTFO_ERR_EMPTY_MIN 1   /* np.min of an empty array (clustering.py:116,120)      -> ValueError        */
#define GTFO_ERR_NAN_INDEX 2   /* list.index(nan)             (clustering.py:117)       -> ValueError        */
#define GTFO_ERR_ZERO_DIV 4    /* 1/len({})                   (helper.py:90)            -> ZeroDivisionError */
#define GTFO_ERR_KEY 8         /* G[u][v] on a removed edge   (helper.py:131,138)       -> KeyError          */
#define GTFO_ERR_NO_TSE 16     /* missing track_state_estimates entry (extrapolate...py:384) -> KeyError    */

typedef struct {
    double sigma0xy, sigma0rz, sigma0rz2, endcap_boundary;
} gtfo_geom;

typedef struct {
    int64_t nodes_with_state, nodes_merged, edges_deactivated; /* cluster */
    int64_t edges_sent, edges_gated;                           /* message passing */
    int64_t edges_reweight_off;                                /* reweight */
} gtfo_stats;

int gtfo_seed(gtfo_arrays *A, const gtfo_geom *g);
int gtfo_initialize_edge_activation(gtfo_arrays *A);
int gtfo_compute_prior_probabilities(gtfo_arrays *A, int key);
int gtfo_compute_mixture_weights(gtfo_arrays *A, int key);
int gtfo_query_node_degree(gtfo_arrays *A);
/* kl_lut: NULL for the scalar threshold, else 28 kl_max values indexed by floor(emp_var/0.05) */
int gtfo_cluster(gtfo_arrays *A, int key, double chi2_thr, double kl_thr, const double *kl_lut,
                 const gtfo_geom *g, gtfo_stats *st);
int gtfo_message_passing(gtfo_arrays *A, double chi2_cut, const gtfo_geom *g, gtfo_stats *st);
int gtfo_reweight(gtfo_arrays *A, int key, double threshold, gtfo_stats *st);
int gtfo_remove_state_metadata(gtfo_arrays *A, gtfo_stats *st);
/* weakly connected components per in-play sub-graph; label = smallest node index of the component.
 * Reproduces CCA's quirk: a sub-graph with no inactive edge is ONE component (extract...py:343-344). */
int gtfo_cca(gtfo_arrays *A);
/* candidate acceptance (extract...py:402-467): returns number accepted; accepted[i]=1 marks nodes,
 * pvals (2 per label-root node index) optional.  Removes accepted nodes, updates sub_state. */
int gtfo_extract(gtfo_arrays *A, const gtfo_geom *g, double pval_cut, int numhits, double sep3d,
                 double merge_dist, uint8_t *accepted, double *pval_xy, double *pval_zr);
/* tag_propagation.py:64-164 on in-play sub-graphs; tags out = final tag per node (node index), returns sweeps */
int gtfo_tag_propagation(gtfo_arrays *A, double threshold, int32_t *tags, int max_sweeps);

/* pure helpers exposed for unit tests (clustering.py:11-124) */
double gtfo_kl_distance(const double m1[3], const double c1[9], const double m2[3], const double c2[9]);
void gtfo_merge_states(const double m1[3], const double c1[9], const double m2[3], const double c2[9],
                       double mm[3], double mc[9]);
double gtfo_mahalanobis(const double m1[3], const double c1[9], const double m2[3], const double c2[9],
                        const double node[4], const double nb1[4], const double nb2[4], double sigma0rz,
                        double sigma0rz2, double endcap);
double gtfo_chi2_sf(double x, double k);
/* learn_KL_parabolic_model/src/generate_training_data/utils.py:221-299 for one (node, neighbour) pair */
void gtfo_seed_parabolic(const double node_xy[2], const double nbr_xy[2], double sigma0, double sigmaA, double sigmaB,
                         double sv[3], double cov[9]);
#endif
