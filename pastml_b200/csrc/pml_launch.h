// Table of host-side launchers for one (model family, padded state count) instantiation of the templated kernels.
// Every instantiation lives in its own translation unit (pml_inst.cu compiled with -DPML_FAMILY=.. -DPML_KP=..)
// so that the build parallelises; the engine picks the table once per problem.
#pragma once
#include <cstdlib>
#include "pml_kernels.cuh"
#include "pml_counts.cuh"

namespace pml {

struct KernelSet {
    void (*bu)(cudaStream_t, const BuArgs &, bool maxp, int slot_begin, int n_slots);
    void (*root_loglik)(cudaStream_t, const TreeDev &, const ZcDev &, const LaneDev &, const uint64_t *mask,
                        int alter, const double *bu, const int *bu_ex, const int *roots, int n_roots,
                        int is_marginal, double *out);
    void (*td)(cudaStream_t, const TdArgs &, int node_begin, int n_nodes);
    // marginal / MAP / MPPA of every node from stored top-down vectors (after a vector-only top-down sweep)
    void (*td_select)(cudaStream_t, const TdArgs &, const double *td_all, const int *td_ex_all, int n_nodes);
    void (*mppa_fix)(cudaStream_t, int n_rows, const int *row_node, int C, int K, uint64_t kmask, const ZcDev &,
                     const uint64_t *base_mask, const uint64_t *prev_initial, const uint8_t *prev_hasinit,
                     const double *zc_lh, uint8_t *zc_k, const int *joint_state, int force_joint,
                     uint64_t *mppa_mask, unsigned long long *k_hist);
    // scenario sampling for marginal_counts (pml_counts.cuh): the roots' states, then one level of branches
    void (*counts)(cudaStream_t, const CountsArgs &, bool roots);
};

// Launch of a level kernel with programmatic stream serialization (see pdl_wait in pml_coop.cuh); PASTML_B200_NO_PDL=1
// launches plainly (A/B measurements)
template <class... KArgs, class... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    static const bool off = getenv("PASTML_B200_NO_PDL") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = off ? 0 : 1;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

inline unsigned grid_for(long long items) {
    long long g = (items + kBlock - 1) / kBlock;
    if (g < 1) g = 1;
    if (g > 0x7fffffffLL) g = 0x7fffffffLL;
    return (unsigned) g;
}

// family: 0 = F81, 1 = HKY, 2 = EIG (PML_MODEL_*); returns nullptr when the combination is not built
const KernelSet *kernel_set(int family, int kp);

// Warp-cooperative F81 sweeps (pml_coop.cuh): G threads per (node, lane) item, KT states per thread.
struct CoopSet {
    void (*bu)(cudaStream_t, const BuArgs &, int slot_begin, int n_slots);
    void (*td)(cudaStream_t, const TdArgs &, int node_begin, int n_nodes);
    int G, KT, JT;   // JT > 0: built for exactly JT lanes and K == G * KT
    // TMA-staged variant of the top-down sweep (JT > 0 only; problems without external tables), else nullptr
    void (*td_tma)(cudaStream_t, const TdArgs &, int node_begin, int n_nodes);
};
const CoopSet *coop_set(int g, int kt, int jt);   // nullptr when not built

// Thread-per-item F81 bottom-up sweep with TMA-staged operands (pml_tpi.cuh)
int tpi_tile_lanes(int K, int J);   // lanes per tile for a pass with J lanes, 0 = not applicable
void tpi_bu_launch(int jt, cudaStream_t, const BuArgs &, int slot_begin, int n_slots);

// Bottom-up sweep of the eigen-decomposition models on the FP64 tensor cores (pml_mma.cu): lanes share one decomposition
bool mma_bu_applicable(int K, int J);
void mma_bu_launch(cudaStream_t, const BuArgs &, int slot_begin, int n_slots);
// max-product (JOINT) sweep with P formed on the tensor cores, one warp per (node, lane); lanes share one decomposition
bool mma_maxp_applicable(int K);
void mma_maxp_launch(cudaStream_t, const BuArgs &, int slot_begin, int n_slots);
// top-down vectors of one level on the tensor cores (needs BuArgs::v_all of the preceding bottom-up sweep): writes
// td_all [N][K][C] / td_ex_all [N][C]
void mma_td_launch(cudaStream_t, const TdArgs &, const double *v_all, double *td_all, int *td_ex_all, int node_begin,
                   int n_nodes);

// Derived topology arrays of a level-ordered forest, built on the device from parent[] / dist[] (pml_topo.cu).
// Stage 1 counts children, checks the level order and scans; stage 2 (once the number of internal nodes is known and
// the slot-indexed arrays are allocated) fills every array the sweeps read.
struct TopoOut {
    int *n_children, *first_child, *slot, *pslot, *zc_index;   // [N]
    int *int_nodes, *fc_slot, *nc_slot;                        // [NI]
    int4 *bu_desc;                                             // [2 NI]
    int4 *td_desc;                                             // [N]
};
constexpr int kTopoNotBeforeChild = 1, kTopoNotSorted = 2, kTopoZeroBranch = 4;
size_t topo_scan_bytes(int n_nodes);
// flags_ni (device, 2 ints): [0] = OR of the kTopo* bits, [1] = number of internal nodes; scan: [N] packed prefix sums
void topo_stage1(cudaStream_t, int n_nodes, const int *parent, const double *dist, int *n_children,
                 unsigned long long *scan, void *scan_tmp, size_t scan_tmp_bytes, int *flags_ni);
void topo_stage2(cudaStream_t, int n_nodes, int n_roots, int n_internal, const int *parent, const double *dist,
                 const unsigned long long *scan, const TopoOut &out, const int *level_ptr, int n_level_ptr,
                 int *int_level_ptr);

}  // namespace pml
