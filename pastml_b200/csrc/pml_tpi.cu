// Host-side launcher of the thread-per-item, TMA-staged F81 bottom-up sweep (pml_tpi.cuh).
#include <cstdlib>
#include "pml_tpi.cuh"
#include "pml_launch.h"

namespace pml {
namespace {

constexpr size_t kSmemPerSm = 227 * 1024;

template <int JT>
void bu_launch_t(cudaStream_t st, const BuArgs &a, int s0, int ns) {
    int block = PML_TPI_BLOCK;
    while (block > 32 && tpi_bu_smem(a.ln.K, JT, block) > 56 * 1024) block /= 2;
    const size_t sh = tpi_bu_smem(a.ln.K, JT, block);
    int resident = (int) (kSmemPerSm / (sh + 1024));
    const int by_threads = PML_TPI_BU_MINB * PML_TPI_BLOCK / block;
    if (resident > by_threads) resident = by_threads;
    if (resident < 1) resident = 1;
    static size_t configured = 0;
    if (sh > configured) {   // the attribute only ever has to grow
        cudaFuncSetAttribute(bu_tpi_kernel<JT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sh);
        cudaFuncSetAttribute(bu_tpi_kernel<JT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        configured = sh;
    }
    const int tiles = a.ln.J / JT;
    const int per_block = (block / 32) * (32 / JT);
    long long gx = ((long long) ns + per_block - 1) / per_block;
    long long cap = (148LL * resident + tiles - 1) / tiles;   // persistent grid: one resident wave over all tiles
    if (cap < 1) cap = 1;
    if (gx > cap) gx = cap;
    launch_pdl(bu_tpi_kernel<JT>, dim3((unsigned) gx, (unsigned) tiles), dim3(block), sh, st, a, s0, ns);
}

}  // namespace

// lanes per tile the TMA-staged sweep would use for a pass with J lanes, 0 when it does not apply
int tpi_tile_lanes(int K, int J) {
    if (K < 2 || K > 62) return 0;
    for (int jt : {32, 16, 8, 4})
        if (J == jt) return jt;   // one tile: the plain layout
    // more lanes (parameter sets x columns): the widest tile that divides them, tile-major layout -- a wider tile
    // shares a node's descriptor, mask words and bulk copies between more lanes
    static const int max_tile = getenv("PASTML_B200_TPI_MAX_TILE") ? atoi(getenv("PASTML_B200_TPI_MAX_TILE")) : 32;
    for (int jt : {32, 16, 8, 4})
        if (jt <= max_tile && J % jt == 0) return jt;
    return 0;
}

void tpi_bu_launch(int jt, cudaStream_t st, const BuArgs &a, int s0, int ns) {
    switch (jt) {
        case 4: bu_launch_t<4>(st, a, s0, ns); break;
        case 8: bu_launch_t<8>(st, a, s0, ns); break;
        case 16: bu_launch_t<16>(st, a, s0, ns); break;
        default: bu_launch_t<32>(st, a, s0, ns); break;
    }
}

}  // namespace pml
