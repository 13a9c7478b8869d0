// One (G, KT) instantiation of the warp-cooperative F81 sweeps.  Compile with -DPML_G=<threads per item>
// -DPML_KT=<states per thread>; optionally -DPML_JT=<lanes> for a variant with compile-time lane count and
// K == G * KT (address arithmetic folded into immediates).
#include "pml_coop.cuh"
#include "pml_launch.h"

#if !defined(PML_G) || !defined(PML_KT)
#error "PML_G and PML_KT must be defined"
#endif

namespace pml {
namespace {
#ifndef PML_JT
#define PML_JT 0
#endif
constexpr int G = PML_G, KT = PML_KT, JT = PML_JT;
constexpr bool KX = JT != 0;
constexpr size_t kMaxDynSmem = 200 * 1024;
#ifndef PML_COOP_GRID
#define PML_COOP_GRID 8
#endif

unsigned coop_grid(int n, int lanes, int blocks_per_sm = PML_COOP_GRID) {
    const int per_block = (kCoopBlock / 32) * (32 / (G * lanes));
    long long g = ((long long) n + per_block - 1) / per_block;
    const long long cap = 148LL * blocks_per_sm;   // one resident wave; the kernels loop (and prefetch) beyond it
    if (g > cap) g = cap;
    return (unsigned) (g < 1 ? 1 : g);
}

// shared memory the TMA variants need per SM for their resident blocks, as a carve-out percentage (set once)
template <class Kernel>
void prefer_smem(Kernel kernel, size_t per_block, bool &done) {
    if (done) return;
    done = true;
    const size_t per_sm = (per_block + 1024) * coop_min_blocks(KT, true);   // + 1 KB reserved per block
    int pct = (int) ((per_sm * 100 + 228 * 1024 - 1) / (228 * 1024));
    if (pct > 100) pct = 100;
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    if (per_block > 48 * 1024)
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMaxDynSmem);
}

void bu_launch(cudaStream_t st, const BuArgs &a, int s0, int ns) {
    if (JT > 0 && a.ln.J > JT) {
        // more lanes than the compile-time tile: one grid row per tile of JT lanes
        const int tiles = a.ln.J / JT;
        const size_t sh = sizeof(double) * (4 + a.ln.K) * JT;
        unsigned gx = coop_grid(ns, JT);
        const unsigned per_row = (148u * PML_COOP_GRID + tiles - 1) / tiles;
        if (gx > per_row) gx = per_row;
        launch_pdl(bu_coop_kernel<KT, G, JT, KX, true>, dim3(gx, tiles), dim3(kCoopBlock), sh, st, a, s0, ns);
        return;
    }
    const size_t sh = sizeof(double) * (4 + a.ln.K) * a.ln.J;
    if (sh > 48 * 1024)
        cudaFuncSetAttribute(bu_coop_kernel<KT, G, JT, KX, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMaxDynSmem);
    launch_pdl(bu_coop_kernel<KT, G, JT, KX, false>, dim3(coop_grid(ns, a.ln.J)), dim3(kCoopBlock), sh, st, a, s0, ns);
}

void td_launch(cudaStream_t st, const TdArgs &a, int n0, int nn) {
    const size_t sh = sizeof(double) * (4 + a.ln.K) * a.ln.C + sizeof(unsigned) * a.ln.C * (a.ln.K + 1);
    if (sh > 48 * 1024)
        cudaFuncSetAttribute(td_coop_kernel<KT, G, JT, KX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMaxDynSmem);
    launch_pdl(td_coop_kernel<KT, G, JT, KX>, dim3(coop_grid(nn, a.ln.C)), dim3(kCoopBlock), sh, st, a, n0, nn);
}

#if PML_JT > 0
constexpr int kTmaBlocks = coop_min_blocks(KT, true);   // persistent grid: one resident wave
// TMA-staged top-down sweep (compile-time lanes, K == G * KT, no external tables): see pml_coop.cuh
void td_tma_launch(cudaStream_t st, const TdArgs &a, int n0, int nn) {
    const size_t sh = td_tma_smem(G * KT, JT, G);
    static bool done = false;
    prefer_smem(td_coop_kernel<KT, G, JT, KX, true>, sh, done);
    launch_pdl(td_coop_kernel<KT, G, JT, KX, true>, dim3(coop_grid(nn, JT, kTmaBlocks)), dim3(kCoopBlock), sh, st, a, n0, nn);
}
#endif
}  // namespace

#define PML_CS_NAME3(g, k, j) pml_coop_##g##_##k##_##j
#define PML_CS_NAME2(g, k, j) PML_CS_NAME3(g, k, j)
#define PML_CS_NAME PML_CS_NAME2(PML_G, PML_KT, PML_JT)

const CoopSet *PML_CS_NAME() {
#if PML_JT > 0
    static const CoopSet cs = {bu_launch, td_launch, G, KT, JT, td_tma_launch};
#else
    static const CoopSet cs = {bu_launch, td_launch, G, KT, JT, nullptr};
#endif
    return &cs;
}

}  // namespace pml
