// One KT (states per thread, two threads per item) instantiation set of the single-launch top-down sweep
// (td_flow_kernel, pml_flow.cuh).  Compile with -DPML_KT=<states per thread>; -DPML_KX=1 adds the variants that know
// K == 2 KT at compile time.
#include "pml_flow.cuh"
#include "pml_launch.h"
#include "pml_flow_host.h"

#if !defined(PML_KT)
#error "PML_KT must be defined"
#endif
#ifndef PML_KX
#define PML_KX 0
#endif

namespace pml {
namespace {
constexpr int KT = PML_KT, G = 2;

template <int JT, bool KX>
void td_launch_t(cudaStream_t st, const FlowTdArgs &a, int tiles) {
    const int K = a.ln.K;
    int block = PML_FLOW_BLOCK;
    while (block > 32 && flow_td_smem(K, G, JT, block) > kFlowSmemCap / 2) block /= 2;
    const size_t sh = flow_td_smem(K, G, JT, block);
    static int resident[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    static size_t configured[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 7;
    if (sh != configured[dev]) {
        resident[dev] = resident_blocks(td_flow_kernel<KT, G, JT, KX>, block, sh);
        configured[dev] = sh;
    }
    const unsigned gx = flow_grid_x(resident[dev], tiles, a.sched.n_chunks, block);
    td_flow_kernel<KT, G, JT, KX><<<dim3(gx, (unsigned) tiles), block, sh, st>>>(a);
}

template <bool KX>
void td_launch_j(int jt, cudaStream_t st, const FlowTdArgs &a, int tiles) {
    switch (jt) {
        case 2: td_launch_t<2, KX>(st, a, tiles); break;
        case 4: td_launch_t<4, KX>(st, a, tiles); break;
        case 8: td_launch_t<8, KX>(st, a, tiles); break;
        default: td_launch_t<16, KX>(st, a, tiles); break;
    }
}
}  // namespace

#define PML_FTD_NAME2(k) pml_flow_td_##k
#define PML_FTD_NAME(k) PML_FTD_NAME2(k)

void PML_FTD_NAME(PML_KT)(int jt, cudaStream_t st, const FlowTdArgs &a, int tiles) {
#if PML_KX
    if (a.ln.K == G * KT) { td_launch_j<true>(jt, st, a, tiles); return; }
#endif
    td_launch_j<false>(jt, st, a, tiles);
}

}  // namespace pml
