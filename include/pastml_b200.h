/*
 * pastml_b200 -- C ABI of the B200-native ancestral-character-reconstruction engine.
 *
 * This is the drop-in boundary for the numerical hot path of evolbioinfo/pastml (pure Python + numpy):
 *   pastml/ml.py          get_bottom_up_loglikelihood, calculate_top_down_likelihood,
 *                         calculate_marginal_likelihoods, choose_ancestral_states_{joint,map,mppa}
 *   pastml/parsimony.py   uppass, acctran, downpass, deltran, get_num_parsimonious_steps
 *   pastml/models/*.py    Model.get_Pij_t
 * The reference has no FFI of its own; these entry points are what a ctypes binding inside pastml would bind
 * (INTEGRATION.md shows that binding).  Plain pointers and sizes only, no torch / CUDA types.
 *
 * Conventions
 *   - Return value 0 = success; non-zero = error, text in pml_last_error() (per calling thread).
 *   - Host pointers are borrowed for the duration of the call.  Device memory belongs to the handle.
 *   - One handle = one forest x C characters ("columns") that share the number of states K and the model
 *     family.  A handle owns one CUDA stream; calls on one handle must not overlap, different handles may be
 *     used from different threads.
 *   - Nodes are numbered in level order over the forest: all roots, then all depth-1 nodes, ...; siblings are
 *     contiguous and ordered by parent, so `parent[]` is non-decreasing.  For one tree this is ete3's default
 *     traverse() order, i.e. the row order of the reference's marginal-probability table (ml.py:498-502).
 *   - State sets are bit masks in a uint64 (bit k = state k); K <= 62.
 *   - "Parameter set": the reference evaluates its objective for one parameter vector at a time
 *     (ml.py:202-208); here B parameter sets x C columns are evaluated by one launch sequence.  params is
 *     [B][C][pml_param_stride()] doubles:
 *        [0] sf  [1] tau  [2] tau_factor               (Model.transform_t, models/__init__.py:269-270, 39-42)
 *        F81 family (F81/JC/EFT): [3] mu = 1/(1 - pi.pi)  [4..4+K) pi                    (F81Model.py:18-46)
 *        HKY:                     [3] kappa               [4..8)  pi (A,C,G,T)             (HKYModel.py:44-82)
 *        EIG (CUSTOM_RATES/JTT):  [3] unused  [4..4+K) pi  then d[K], A[K*K], Ainv[K*K] row-major
 *                                                                                       (generator.py:16-65)
 */
#ifndef PASTML_B200_H
#define PASTML_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pml_problem pml_problem;

/* model families (what get_Pij_t computes) */
#define PML_MODEL_F81 0
#define PML_MODEL_HKY 1
#define PML_MODEL_EIG 2

/* which allowed-state masks a likelihood pass uses (the reference keeps them in <col>_ALLOWED_STATES) */
#define PML_MASK_ANNOTATION 0   /* initialize_allowed_states, ml.py:293-318 */
#define PML_MASK_JOINT 1        /* after choose_ancestral_states_joint  */
#define PML_MASK_MAP 2          /* after choose_ancestral_states_map    */
#define PML_MASK_MPPA 3         /* after choose_ancestral_states_mppa   */
#define PML_MASK_CUSTOM 4       /* uploaded with pml_set_custom_masks (e.g. parsimony states, ml.py:724) */

/* parsimony methods (bit flags) */
#define PML_PARS_ACCTRAN 1
#define PML_PARS_DOWNPASS 2
#define PML_PARS_DELTRAN 4

#define PML_REUSE_BOTTOM_UP 2
#define PML_RESULTS_ON_DEVICE 4

/* error codes */
#define PML_OK 0
#define PML_ERR_INVALID 1
#define PML_ERR_CUDA 2
#define PML_ERR_ZERO_LIKELIHOOD 3   /* PastMLLikelihoodError of ml.py:139-145: see pml_zero_likelihood_info */
#define PML_ERR_NO_DEVICE 4

const char *pml_last_error(void);
int pml_version(void);
int pml_device_count(void);
int pml_param_stride(int32_t model_kind, int32_t n_states);

/* Creates a problem: topology (level-ordered), C columns with K states, annotation masks.
 *   annot_mask[C][N]  annotated state set of each node, 0 = none (=> all states allowed, ml.py:311-313)
 *   has_state[C][N]   node carries the annotation feature at all (ml.py:329-331), drives zero-branch clusters
 *   max_param_sets    capacity B of parameter sets per launch (>= 1)
 *   device            CUDA device ordinal */
int pml_problem_create(int32_t n_nodes, const int32_t *parent, const double *dist,
                       int32_t n_columns, int32_t n_states, int32_t model_kind,
                       const uint64_t *annot_mask, const uint8_t *has_state,
                       int32_t max_param_sets, int32_t device, pml_problem **out);
void pml_problem_destroy(pml_problem *p);

/* get_bottom_up_loglikelihood(tree, character, model, is_marginal, alter) summed over the forest's trees
 * (ml.py:82-121, 206-207) for n_sets x C (parameter set, column) pairs in one launch sequence.
 *   mask_kind  PML_MASK_*;   is_marginal 1 = sum-product, 0 = max-product (no pointers kept);
 *   alter      1 = apply the zero-branch cluster alteration where tau == 0 (ml.py:100-102, 352-387)
 *   out_loglh  [n_sets][C]
 * On PML_ERR_ZERO_LIKELIHOOD the other lanes' values are still written (the failing lane gets -inf). */
int pml_loglik(pml_problem *p, const double *params, int32_t n_sets, int32_t mask_kind, int32_t is_marginal,
               int32_t alter, double *out_loglh);

/* JOINT reconstruction (ml.py:687-695): max-product bottom-up pass on the annotation masks with argmax
 * pointers, unalter_zero_node_joint_states, top-down traceback.  params = one set [C][stride].
 *   out_loglh[C]        log_likelihood_restricted_JOINT            (may be NULL)
 *   out_state[N][C]     JOINT_STATE index of every node (int32)    (may be NULL)
 * Leaves the JOINT masks and states on the device for later passes. */
int pml_joint(pml_problem *p, const double *params, double *out_loglh, int32_t *out_state);

/* Marginal reconstruction (ml.py:697-714, 735-742): bottom-up, top-down, marginal likelihoods,
 * probabilities, MAP and MPPA state selection, fused in one level-synchronous sweep.
 *   force_joint        bit 0: MPPA keeps room for the JOINT state (needs a previous pml_joint call);
 *                      bit 1 (PML_REUSE_BOTTOM_UP): the bottom-up state of the immediately preceding pml_loglik call
 *                      (same parameters, one set, annotation masks, marginal, alter) is reused instead of recomputed;
 *                      bit 2 (PML_RESULTS_ON_DEVICE): probabilities / node log-likelihoods are computed into the
 *                      handle's device buffers even when the host pointers are NULL
 *   out_loglh[C]       marginal log-likelihood                                     (may be NULL)
 *   out_probs[N][C][K] marginal probabilities, ml.py:486-502                        (may be NULL)
 *   out_node_log10lh[N][C]  log10 of the node's total marginal likelihood = log10(sum LIKELIHOOD) -
 *                      LIKELIHOOD_SF of the reference before MAP masking            (may be NULL)
 *   out_map_mask[N][C], out_mppa_mask[N][C]   selected state sets                   (may be NULL)
 *   out_k_hist[C][K+1] histogram of the number of states MPPA kept per node         (may be NULL)
 * The MPPA outputs are final only after pml_mppa_finalize. */
int pml_marginal(pml_problem *p, const double *params, int32_t force_joint,
                 double *out_loglh, double *out_probs, double *out_node_log10lh,
                 uint64_t *out_map_mask, uint64_t *out_mppa_mask, int64_t *out_k_hist);

/* The reference selects MPPA states after the MAP-restricted likelihood pass, whose zero-branch alteration
 * may have replaced the '.initial' masks MPPA multiplies by (ml.py:541-542, 384-385).  Call this after
 * pml_loglik(PML_MASK_MAP, alter=1) to replay that for the few affected nodes.  Outputs as in pml_marginal. */
int pml_mppa_finalize(pml_problem *p, int32_t force_joint, uint64_t *out_mppa_mask, int64_t *out_k_hist);

/* Parsimony (parsimony.py:245-331): uppass, then the requested methods.  out_mask[m][N][C] for m in
 * ACCTRAN, DOWNPASS, DELTRAN order restricted to the requested ones; out_steps[m][C]. */
int pml_parsimony(pml_problem *p, int32_t methods, uint64_t *out_mask, int64_t *out_steps);

/* Uploads [C][N] masks into the PML_MASK_CUSTOM slot. */
int pml_set_custom_masks(pml_problem *p, const uint64_t *mask);
/* Downloads a mask slot as [N][C]. */
int pml_get_masks(pml_problem *p, int32_t mask_kind, uint64_t *out_mask);
/* Nodes that currently carry the reference's '<col>_ALLOWED_STATES.initial' feature: out_has[N][C] (0/1),
 * out_initial[N][C]. */
int pml_get_initial(pml_problem *p, uint8_t *out_has, uint64_t *out_initial);

/* The bottom-up state left by the last single-set sum-product pass (pml_loglik with n_sets = 1, or pml_marginal): the
 * mantissa vectors out_vec[NI][K][C] and exponents out_ex[NI][C] of the internal nodes (slot order = internal nodes by
 * increasing node index; a tip's vector is its allowed-state mask), the allowed-state masks in force during that
 * pass out_allowed[N][C] (after the zero-branch alteration) and out_altered[N][C] = 1 where the alteration changed
 * the node's mask.  Input of marginal_counts (ml.py:753-862). */
int pml_get_bottom_up(pml_problem *p, double *out_vec, int32_t *out_ex, uint64_t *out_allowed, uint8_t *out_altered);

/* After PML_ERR_ZERO_LIKELIHOOD: first failing (column, parameter set, parent node, child node). */
int pml_zero_likelihood_info(pml_problem *p, int32_t *column, int32_t *param_set, int32_t *parent, int32_t *child);

/* After PML_ERR_ZERO_LIKELIHOOD from pml_loglik: (parent, child) of every lane of that call, -1 where the lane
 * was fine.  out_pairs[n_lanes][2], lane = param_set * C + column. */
int pml_zero_likelihood_all(pml_problem *p, int32_t *out_pairs, int32_t n_lanes);

/* ---- Subtree sharding of one big tree: a tree is cut at a depth level; every GPU owns the
 * clades below the cut as a forest, the part above the cut ("top tree") is replicated and sees the clade roots as
 * tips whose bottom-up vectors are supplied from outside.  One exchange per sweep: the clade-root vectors (all-gather
 * between the GPUs, done by the caller), and the top-down vectors handed back to the clade roots.
 *   pml_get_root_partials  after pml_loglik on a clade forest: bottom-up vectors of its roots,
 *                          out_vec[T][K][J] mantissas, out_ex[T][J] power-of-two exponents (J = n_sets * C)
 *   pml_set_tip_partials   on the top tree: attach such vectors to the n listed tips (same set on every call)
 *   pml_get_tip_td         after pml_marginal on the top tree: top-down vectors computed for those tips,
 *                          out_vec[n][K][C], out_ex[n][C]
 *   pml_set_root_td        on a clade forest before pml_marginal: top-down vectors arriving at its roots
 *                          (vec[T][K][C], ex[T][C]) instead of the all-ones vector of a true root (ml.py:274-277) */
int pml_get_root_partials(pml_problem *p, int32_t n_sets, double *out_vec, int32_t *out_ex);
int pml_set_tip_partials(pml_problem *p, int32_t n, const int32_t *nodes, const double *vec, const int32_t *ex,
                         int32_t n_sets);
int pml_get_tip_td(pml_problem *p, double *out_vec, int32_t *out_ex);
int pml_set_root_td(pml_problem *p, const double *vec, const int32_t *ex);

/* The optimiser's objective over an explicit list of lanes (replaces the n_sets x all-columns grid of pml_loglik when
 * only some columns are still being fitted, ml.py:202-208): lane i evaluates column lane_columns[i] with the parameter
 * block params[i][stride]; annotation masks, sum-product, alteration on.  n_lanes must be a multiple of 4 (pad with
 * copies of a lane) and at most max_param_sets * C.  Available when pml_supports_lane_lists() is non-zero (F81 family,
 * no external tables).  Zero-likelihood lanes are reported like pml_loglik (pml_zero_likelihood_all: one pair per lane). */
int pml_supports_lane_lists(pml_problem *p);
int pml_loglik_lanes(pml_problem *p, const double *params, const int32_t *lane_columns, int32_t n_lanes,
                     double *out_loglh);

/* Device buffers of destroyed problems are kept in a per-process cache and reused by the next problem of the same
 * shape (cudaMalloc/cudaFree of GBs cost more than the sweeps).  This returns the cached blocks to the driver. */
int pml_trim_memory(void);

/* Timing of the last call on this handle: device time between the first and last kernel of the call, measured
 * with CUDA events on the handle's stream (ms), and the number of kernels launched. */
int pml_last_call_stats(pml_problem *p, double *device_ms, int64_t *kernel_launches);

/* Expected numbers of state transitions over the forest from n_repetitions ancestral scenarios sampled from the
 * marginal posterior -- pastml.ml.marginal_counts (pastml/ml.py:753-862): the alteration of zero-branch clusters and
 * the bottom-up sweep run first, then the roots' states are drawn from their marginal probabilities and every branch
 * draws the child's state given the parent's, level by level (Philox4x32-10 keyed by `seed`, counted by repetition,
 * node and column: same seed, same result).  out_counts: [n_columns][K][K], row = from, column = to, per repetition;
 * the diagonal holds the lineages that stay (ml.py:856-858). */
int pml_marginal_counts(pml_problem *p, const double *params, int32_t n_repetitions, uint64_t seed, double *out_counts);

/* Which kernels the last bottom-up and the last top-down launch sequence of this handle consisted of, as text:
 * "bottom_up:bu_coop_kernel=29/6011,bu_tpi_kernel=24/993988;top_down:td_coop_kernel=54/1999999"
 * (kernel = launches per sequence / items -- internal nodes or nodes -- those launches cover).
 * Measurement only (bench.py names the kernels of its roofline from this instead of guessing); no reference
 * counterpart. */
int pml_last_kernels(pml_problem *p, char *buf, int32_t buf_len);

/* ---- The same split with the exchange kept on the device (pastml_b200/sharded.py, bench.py --gpus N).  A pml_shard ties
 * one rank's clade forest and its copy of the top tree (same device).  boundary_top[n_clades]: the clade roots as tip ids
 * of the top tree, in boundary order; clade_owner[n_clades]: the rank that owns each clade (the clade forest's roots are
 * this rank's clades in boundary order).  Per sweep:
 *   pml_shard_bottom_up   clades bottom-up (masks of the given kind), clade-root runs packed and all-gathered over NCCL
 *                         on the clade problem's stream -- no host synchronisation;
 *   pml_shard_finish      gathered runs -> external tips of the top tree, top tree bottom-up, log-likelihood of the whole
 *                         tree (the same on every rank; ml.py:82-121); out_loglh null = nothing fetched, no host sync;
 *   pml_shard_top_down    fused top-down / marginal / MAP / MPPA sweep (ml.py:240-290, 431-465, 505-595): top tree on
 *                         every rank, the rank's clade roots receive their top-down vectors on the device, clades;
 *                         force_joint as in pml_marginal (after pml_shard_joint_finish);
 *   pml_shard_results     outputs of either problem (which: 0 = top tree, 1 = clade forest), shapes as pml_marginal.
 * NCCL is resolved from the running process (libnccl.so.2): pml_shard_unique_id on rank 0, the 128 bytes broadcast by the
 * caller, pml_shard_connect on every rank.  pml_shard_exchange_local replaces the all-gather when the ranks are
 * simulated in one process on one device (tests). */
typedef struct pml_shard pml_shard;
int pml_shard_create(pml_problem *clades, pml_problem *top, int32_t rank, int32_t world, int32_t n_clades,
                     const int32_t *boundary_top, const int32_t *clade_owner, pml_shard **out);
void pml_shard_destroy(pml_shard *sh);
int pml_shard_unique_id(uint8_t *out_id128);
int pml_shard_connect(pml_shard *sh, const uint8_t *id128);
int pml_shard_bottom_up(pml_shard *sh, const double *params, int32_t n_sets, int32_t mask_kind);
int pml_shard_exchange_local(pml_shard **shards, int32_t world);
int pml_shard_finish(pml_shard *sh, const double *params, int32_t n_sets, double *out_loglh);
int pml_shard_top_down(pml_shard *sh, int32_t want_probs, int32_t want_node_lh, int32_t force_joint);
/* JOINT on the sharded tree (ml.py:598-622, 687-695): max-product sweep of the clades + all-gather, then the top tree's
 * max-product sweep, the joint log-likelihood of the whole tree (out_loglh [C], the same on every rank), the top tree's
 * traceback, and the clades' traceback started from the states the top tree chose for their roots.  Afterwards
 * pml_shard_bottom_up(mask_kind = PML_MASK_JOINT) gives the JOINT-restricted likelihood and pml_shard_top_down may
 * force the joint state into the MPPA sets (ml.py:545). */
int pml_shard_joint_bottom_up(pml_shard *sh, const double *params);
int pml_shard_joint_finish(pml_shard *sh, const double *params, double *out_loglh);
int pml_shard_joint_states(pml_shard *sh, int32_t which, int32_t *out_state);
int pml_shard_results(pml_shard *sh, int32_t which, double *out_probs, double *out_node_log10lh, uint64_t *out_map_mask,
                      uint64_t *out_mppa_mask, int64_t *out_k_hist);
/* device-timed sharded MPPA step, n_iter times (CUDA events on the clade stream around each step); phase_ms: clades
 * bottom-up + pack, exchange, top tree bottom-up, top tree top-down, clades top-down */
int pml_shard_bench_step(pml_shard *sh, const double *params, int32_t n_iter, int32_t flush_l2, double *avg_ms,
                         double *phase_ms /* [5] of the last step, or null */);

/* Device-resident variant used by bench.py: runs a pass n_iter times back to back on resident inputs and returns
 * the average / minimum device time per pass (CUDA events on the handle's stream).  which: 0 = bottom-up
 * objective (annotation masks, marginal, alter) for n_sets parameter sets; 1 = the marginal sweep (bottom-up +
 * fused top-down/marginal/MAP/MPPA, force_joint off); 2 = JOINT (max-product bottom-up + traceback); 3 = the
 * fused top-down sweep alone; 4 = like 1 with force_joint on (needs pml_joint).  flush_l2 != 0 writes a 256 MiB
 * buffer between iterations, outside the timed intervals. */
int pml_bench_pass(pml_problem *p, const double *params, int32_t n_sets, int32_t which, int32_t n_iter,
                   int32_t flush_l2, double *avg_ms, double *min_ms, int64_t *kernel_launches);

/* ------------------------------------------------------------------------------------------------------------
 * Array-native newick ingest (host only; no GPU needed).  Replaces pastml.tree.read_forest / read_tree and the ete3
 * traversal for large trees (pastml/tree.py:176-222): every ';'-terminated tree of the text becomes part of ONE
 * forest numbered in level order (all roots, then all depth-1 nodes, ...; children left to right) -- the node index
 * space of pml_problem_create.  Labels of internal nodes are names (ete3 format 1), a missing branch length is 1.0
 * (0 for a root), negative lengths become 0 (tree.py:194-198), [comments] are skipped.
 * pml_newick_name_nodes applies pastml.tree.name_tree (tree.py:76-105) to every tree (use_tree_suffix: "_<tree>"
 * suffixes for forests of several trees).  pml_newick_fill copies parent[N], dist[N] and, when both are given, the
 * names as a concatenated byte buffer with name_offsets[N + 1]. */
typedef struct pml_newick pml_newick;
const char *pml_newick_last_error(void);
int pml_newick_parse(const char *text, int64_t len, pml_newick **out);
int64_t pml_newick_num_nodes(const pml_newick *nw);
int64_t pml_newick_num_trees(const pml_newick *nw);
int64_t pml_newick_names_bytes(const pml_newick *nw);
int pml_newick_name_nodes(pml_newick *nw, int use_tree_suffix);

/* A table of floats as text, the way the reference writes marginal probabilities (acr.py:45-73,
 * DataFrame.to_csv(sep='\t', index_label='node')): `header` + '\n', then per row its label (labels: one UTF-8 blob,
 * label r = [label_offsets[r], label_offsets[r + 1])), a tab and the Python repr() of every value ([n_rows][n_cols];
 * empty cell for NaN).  Host only.  Returns 0, or 1 with pml_newick_last_error(). */
int pml_write_float_table(const char *path, const char *header, const char *labels, const int64_t *label_offsets,
                          const double *values, int64_t n_rows, int32_t n_cols);
int pml_newick_fill(const pml_newick *nw, int32_t *parent, double *dist, int64_t *name_offsets, char *names);
void pml_newick_free(pml_newick *nw);

#ifdef __cplusplus
}
#endif
#endif
