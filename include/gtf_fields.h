/* gtf_fields.h -- the flat event-batch layout shared by the C-ABI (include/gtf.h), the CUDA
 * kernels, the Python host side and the CPU oracle.  One X-macro row per array:
 *     X(name, c_type, extent)      extent in {N, N1 (=N+1), E, S, S1 (=S+1)}
 *
 * What each array replaces in the reference (all paths under /root/reference/src):
 *   nodes      GNN_Measurement.py:1-9 + node attributes set in utilities/helper.py:497-508.
 *              Nodes are stored in graph-iteration order, sub-graphs concatenated; a node removed by
 *              extraction (extract/extract_track_candidates.py:460-462) keeps its row, alive=0.
 *   in-CSR     one slot per entry (neighbour -> node) of the node's `track_state_estimates` dict, slots
 *              in dict insertion order (helper.py:375-441).  The same slot carries the directed edge
 *              attribute G[neighbour][node] (helper.py:24-25,180) and the `updated_track_states` entry
 *              keyed by that neighbour (extrapolate/extrapolate_merged_states.py:441-447).
 *   out-CSR    slot ids of a node's out-edges in G.successors(node) order (extrapolate...py:430).
 *   tse_*      track_state_estimates[neighbour]  (helper.py:432-441, prior :63, mixture_weight :94)
 *   uts_*      updated_track_states[neighbour]   (extrapolate...py:375-385, helper.py:129-139,177)
 *              uts_rank = insertion stamp (dict order = ascending rank), uts_next = next stamp.
 *   m_*        merged_state / merged_cov / merged_prior node attributes (clustering/clustering.py:291-293)
 * Covariances: (p00 p01 p11 p22); row/col 2 of every stored matrix is zero off the diagonal
 * (helper.py:423-425, extrapolate...py:363-365) and the matrix is symmetric up to rounding.
 */
#ifndef GTF_FIELDS_H
#define GTF_FIELDS_H

#define GTF_FIELDS(X)                 \
    X(x, double, N)                   \
    X(y, double, N)                   \
    X(z, double, N)                   \
    X(r, double, N)                   \
    X(layer, int32_t, N)              \
    X(volume, int32_t, N)             \
    X(sub, int32_t, N)                \
    X(alive, uint8_t, N)              \
    X(sub_off, int32_t, S1)           \
    X(sub_state, uint8_t, S)          \
    X(sub_event, int32_t, S)          \
    X(in_off, int32_t, N1)            \
    X(in_src, int32_t, E)             \
    X(slot_dst, int32_t, E)           \
    X(out_off, int32_t, N1)           \
    X(out_slot, int32_t, E)           \
    X(rev_slot, int32_t, E)           \
    X(active, uint8_t, E)             \
    X(edge_w, double, E)              \
    X(tse_present, uint8_t, E)        \
    X(tse_a, double, E)               \
    X(tse_b, double, E)               \
    X(tse_c, double, E)               \
    X(tse_tau, double, E)             \
    X(tse_p00, double, E)             \
    X(tse_p01, double, E)             \
    X(tse_p11, double, E)             \
    X(tse_p22, double, E)             \
    X(tse_prior, double, E)           \
    X(tse_w, double, E)               \
    X(has_uts, uint8_t, N)            \
    X(uts_next, int32_t, N)           \
    X(uts_present, uint8_t, E)        \
    X(uts_rank, int32_t, E)           \
    X(uts_a, double, E)               \
    X(uts_b, double, E)               \
    X(uts_c, double, E)               \
    X(uts_tau, double, E)             \
    X(uts_p00, double, E)             \
    X(uts_p01, double, E)             \
    X(uts_p11, double, E)             \
    X(uts_p22, double, E)             \
    X(uts_lik, double, E)             \
    X(uts_prior, double, E)           \
    X(uts_w, double, E)               \
    X(uts_lrn, double, E)             \
    X(uts_side, int8_t, E)            \
    X(uts_chi2, double, E)            \
    X(has_merged, uint8_t, N)         \
    X(m_a, double, N)                 \
    X(m_b, double, N)                 \
    X(m_c, double, N)                 \
    X(m_p00, double, N)               \
    X(m_p01, double, N)               \
    X(m_p11, double, N)               \
    X(m_p22, double, N)               \
    X(m_prior, double, N)             \
    X(degree, int32_t, N)             \
    X(label, int32_t, N)              \
    X(emp_var, double, N)

/* sub_state values */
#define GTF_SUB_INPLAY 0   /* still in the list handed to the next stage ("remaining") */
#define GTF_SUB_FRAGMENT 1 /* 1..numhits-1 nodes left after extraction (extract...py:464-465) */
#define GTF_SUB_EMPTY 2    /* every node extracted */

/* which state dict a stage works on (clustering.py:149 `track_state_key`) */
#define GTF_KEY_TSE 0 /* 'track_state_estimates' */
#define GTF_KEY_UTS 1 /* 'updated_track_states'  */

#endif
