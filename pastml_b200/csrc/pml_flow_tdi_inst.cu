// One KP (register-array length) instantiation set of the thread-per-item single-launch top-down sweep
// (td_tpi_flow_kernel, pml_flow_tdi.cuh).  Compile with -DPML_KP=<padded state count>; -DPML_KX=1 adds the variants that
// know K == KP at compile time.
#include "pml_flow_tdi.cuh"
#include "pml_launch.h"
#include "pml_flow_host.h"

#if !defined(PML_KP)
#error "PML_KP must be defined"
#endif
#ifndef PML_KX
#define PML_KX 0
#endif

namespace pml {
namespace {
constexpr int KP = PML_KP;

template <int JT, bool KX>
void tdi_launch_t(cudaStream_t st, const FlowTdArgs &a, int tiles) {
    const int K = a.ln.K;
    int block = PML_FLOW_BLOCK;
    while (block > 32 && flow_tdi_smem(K, JT, block) > kFlowSmemCap / 2) block /= 2;
    const size_t sh = flow_tdi_smem(K, JT, block);
    static int resident[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    static size_t configured[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 7;
    if (sh != configured[dev]) {
        resident[dev] = resident_blocks(td_tpi_flow_kernel<KP, JT, KX>, block, sh);
        configured[dev] = sh;
    }
    const unsigned gx = flow_grid_x(resident[dev], tiles, a.sched.n_chunks, block);
    td_tpi_flow_kernel<KP, JT, KX><<<dim3(gx, (unsigned) tiles), block, sh, st>>>(a);
}

template <bool KX>
void tdi_launch_j(int jt, cudaStream_t st, const FlowTdArgs &a, int tiles) {
    switch (jt) {
        case 2: tdi_launch_t<2, KX>(st, a, tiles); break;
        case 4: tdi_launch_t<4, KX>(st, a, tiles); break;
        case 8: tdi_launch_t<8, KX>(st, a, tiles); break;
        default: tdi_launch_t<16, KX>(st, a, tiles); break;
    }
}
}  // namespace

#define PML_FTDI_NAME2(k) pml_flow_tdi_##k
#define PML_FTDI_NAME(k) PML_FTDI_NAME2(k)

void PML_FTDI_NAME(PML_KP)(int jt, cudaStream_t st, const FlowTdArgs &a, int tiles) {
#if PML_KX
    if (a.ln.K == KP) { tdi_launch_j<true>(jt, st, a, tiles); return; }
#endif
    tdi_launch_j<false>(jt, st, a, tiles);
}

}  // namespace pml
