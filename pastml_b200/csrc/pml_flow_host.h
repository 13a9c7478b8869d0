// Launch geometry shared by the translation units of the single-launch sweeps.
#pragma once
#include <cuda_runtime.h>

namespace pml {

template <class Kernel>
inline int resident_blocks(Kernel kernel, int block, size_t smem) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, block, smem) != cudaSuccess || n < 1) {
        cudaGetLastError();
        n = 1;
    }
    return n;
}

// persistent grid: what is resident at once over all the tiles, but no more blocks than there are chunks of work
inline unsigned flow_grid_x(int resident, int tiles, int n_chunks, int block) {
    long long gx = (148LL * resident + tiles - 1) / tiles;
    const long long need = ((long long) n_chunks + (block / 32) - 1) / (block / 32);
    if (gx > need) gx = need;
    return (unsigned) (gx < 1 ? 1 : gx);
}

constexpr size_t kFlowSmemCap = 225 * 1024;

}  // namespace pml
