// One (model family, KP) instantiation of the level kernels.  Compile with -DPML_FAMILY={0,1,2} -DPML_KP=<n>.
#include <algorithm>
#include "pml_launch.h"

#ifndef PML_FAMILY
#error "PML_FAMILY must be defined"
#endif
#ifndef PML_KP
#error "PML_KP must be defined"
#endif

#define PML_CAT2(a, b, c) pml_ks_##a##_##b##c
#define PML_CAT(a, b) PML_CAT2(a, b, )

namespace pml {

#if PML_FAMILY == 0
template <int KP> using TheLane = F81Lane<KP>;
#elif PML_FAMILY == 1
template <int KP> using TheLane = HKYLane<KP>;
#else
template <int KP> using TheLane = EigLane<KP>;
#endif

namespace {
constexpr int KP = PML_KP;

constexpr int kFamily = TheLane<KP>::family;
constexpr size_t kMaxDynSmem = 200 * 1024;

template <class Kern>
void allow_big_smem(Kern kern, size_t bytes) {
    // kernels need an opt-in above 48 KB of dynamic shared memory (many parameter sets x columns)
    if (bytes > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMaxDynSmem);
}

void bu_launch(cudaStream_t st, const BuArgs &a, bool maxp, int s0, int ns) {
    const long long items = (long long) ns * a.ln.J;
    const size_t sh = sizeof(double) * staged_doubles(kFamily, a.ln.K, a.ln.J);
    if (maxp) {
        allow_big_smem(bu_level_kernel<KP, TheLane, true>, sh);
        launch_pdl(bu_level_kernel<KP, TheLane, true>, dim3(grid_for(items)), dim3(kBlock), sh, st, a, s0, ns);
    } else {
        allow_big_smem(bu_level_kernel<KP, TheLane, false>, sh);
        launch_pdl(bu_level_kernel<KP, TheLane, false>, dim3(grid_for(items)), dim3(kBlock), sh, st, a, s0, ns);
    }
}

void root_launch(cudaStream_t st, const TreeDev &t, const ZcDev &zc, const LaneDev &ln, const uint64_t *mask,
                 int alter, const double *bu, const int *bu_ex, const int *roots, int n_roots, int is_marginal,
                 double *out) {
    root_loglik_kernel<<<grid_for(ln.J), kBlock, 0, st>>>(t, zc, ln, mask, alter, bu, bu_ex, roots, n_roots,
                                                          is_marginal, out);
}

size_t td_smem(const TdArgs &a) {
    return sizeof(double) * staged_doubles(kFamily, a.ln.K, a.ln.C) + sizeof(unsigned) * a.ln.C * (a.ln.K + 1);
}

void td_launch(cudaStream_t st, const TdArgs &a, int s0, int ns) {
    const size_t sh = td_smem(a);
    allow_big_smem(td_level_kernel<KP, TheLane>, sh);
    launch_pdl(td_level_kernel<KP, TheLane>, dim3(grid_for((long long) ns * a.ln.C)), dim3(kBlock), sh, st, a, s0, ns);
}

void td_select_launch(cudaStream_t st, const TdArgs &a, const double *td_all, const int *td_ex_all, int n_nodes) {
    const size_t sh = sizeof(double) * staged_fields(TheLane<KP>::family, a.ln.K) * a.ln.C + sizeof(unsigned) * a.ln.C * (a.ln.K + 1);
    allow_big_smem(td_select_kernel<KP, TheLane>, sh);
    long long g = ((long long) n_nodes * a.ln.C + kBlock - 1) / kBlock;
    if (g > 148LL * 16) g = 148LL * 16;   // grid-stride: the per-block histogram is flushed once
    launch_pdl(td_select_kernel<KP, TheLane>, dim3((unsigned) (g < 1 ? 1 : g)), dim3(kBlock), sh, st, a, td_all, td_ex_all, n_nodes);
}

void mppa_fix_launch(cudaStream_t st, int n_rows, const int *row_node, int C, int K, uint64_t kmask, const ZcDev &zc,
                     const uint64_t *base_mask, const uint64_t *prev_initial, const uint8_t *prev_hasinit,
                     const double *zc_lh, uint8_t *zc_k, const int *joint_state, int force_joint, uint64_t *mppa_mask,
                     unsigned long long *k_hist) {
    mppa_fix_kernel<KP><<<grid_for((long long) n_rows * C), kBlock, 0, st>>>(
        n_rows, row_node, C, K, kmask, zc, base_mask, prev_initial, prev_hasinit, zc_lh, zc_k, joint_state,
        force_joint, mppa_mask, k_hist);
}
void counts_launch(cudaStream_t st, const CountsArgs &a, bool roots) {
    const size_t sh = counts_smem_bytes(kFamily, a.ln.K, a.ln.C);
    const dim3 grid((unsigned) std::max(1, std::min(a.n_slots, 148 * 8)), (unsigned) a.ln.C);
    if (roots) {
        allow_big_smem(counts_root_kernel<KP, TheLane>, sh);
        counts_root_kernel<KP, TheLane><<<grid, 128, sh, st>>>(a);
    } else {
        allow_big_smem(counts_level_kernel<KP, TheLane>, sh);
        counts_level_kernel<KP, TheLane><<<grid, 128, sh, st>>>(a);
    }
}
}  // namespace

#define PML_KS_NAME3(f, k) pml_ks_##f##_##k
#define PML_KS_NAME2(f, k) PML_KS_NAME3(f, k)
#define PML_KS_NAME PML_KS_NAME2(PML_FAMILY, PML_KP)

const KernelSet *PML_KS_NAME() {
    static const KernelSet ks = {bu_launch, root_launch, td_launch, td_select_launch, mppa_fix_launch, counts_launch};
    return &ks;
}

}  // namespace pml
