// Topology derivation on the device: from parent[] / dist[] of a level-ordered forest (roots first, parents
// non-decreasing -- the order pastml's own traversal of ete3 trees yields, pastml/tree.py:read_forest) to the arrays
// the sweeps read: children counts, first child, slot of every internal node, per-slot bottom-up descriptors and
// per-node top-down descriptors.  On the host this pass cost ~25 ms for a 2M-node tree (a third of the end-to-end
// call); here it is three small kernels and one prefix sum behind the 24 MB upload of parent[] and dist[].
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "pml_launch.h"

namespace pml {
namespace {

constexpr int kTopoBlock = 256;

inline unsigned topo_grid(long long n) { return (unsigned) ((n + kTopoBlock - 1) / kTopoBlock); }

__global__ void topo_count_kernel(int N, const int *__restrict__ parent, const double *__restrict__ dist,
                                  int *__restrict__ n_children, int *__restrict__ flags) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= N) return;
    const int q = parent[v];
    int bad = 0;
    if (q >= v) bad |= kTopoNotBeforeChild;
    if (v > 0 && q < parent[v - 1]) bad |= kTopoNotSorted;
    if (q >= 0 && q < v) {
        atomicAdd(n_children + q, 1);
        if (dist[v] == 0.) bad |= kTopoZeroBranch;
    }
    if (bad) atomicOr(flags, bad);
}

// (children << 32) | is_internal: one 64-bit prefix sum gives both the first-child offsets and the slots
struct PackCount {
    __host__ __device__ unsigned long long operator()(int n) const {
        return ((unsigned long long) (unsigned) n << 32) | (n > 0 ? 1ULL : 0ULL);
    }
};

__global__ void topo_total_kernel(int N, const int *__restrict__ n_children, const unsigned long long *__restrict__ scan,
                                  int *__restrict__ flags_ni) {
    flags_ni[1] = (int) (scan[N - 1] & 0xffffffffULL) + (n_children[N - 1] > 0 ? 1 : 0);
}

__global__ void topo_fill_kernel(int N, int n_roots, const int *__restrict__ parent, const double *__restrict__ dist,
                                 const unsigned long long *__restrict__ scan, TopoOut o) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= N) return;
    const unsigned long long sc = scan[v];
    const int n = o.n_children[v];
    const int s = n ? (int) (sc & 0xffffffffULL) : -1;
    const int f = n_roots + (int) (sc >> 32);
    const int q = parent[v];
    const int pq = q < 0 ? -1 : (int) (scan[q] & 0xffffffffULL);   // a parent is internal by definition
    o.slot[v] = s;
    o.first_child[v] = f;
    o.pslot[v] = pq;
    o.zc_index[v] = -1;
    const double d = dist[v];
    o.td_desc[v] = make_int4(pq, s, __double2loint(d), __double2hiint(d));
    if (s >= 0) {
        o.int_nodes[s] = v;
        o.fc_slot[s] = f;
        o.nc_slot[s] = n;
        const int s0 = o.n_children[f] ? (int) (scan[f] & 0xffffffffULL) : -1;
        const int s1 = n > 1 ? (o.n_children[f + 1] ? (int) (scan[f + 1] & 0xffffffffULL) : -1) : -1;
        const double d0 = dist[f], d1 = n > 1 ? dist[f + 1] : 0.;
        o.bu_desc[2 * (size_t) s] = make_int4(f, n, s0, s1);
        o.bu_desc[2 * (size_t) s + 1] = make_int4(__double2loint(d0), __double2hiint(d0), __double2loint(d1), __double2hiint(d1));
    }
}

// internal nodes in front of each level boundary
__global__ void topo_levels_kernel(int N, int n, const int *__restrict__ level_ptr, const unsigned long long *__restrict__ scan,
                                   int n_internal, int *__restrict__ int_level_ptr) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n) return;
    const int at = level_ptr[l];
    int_level_ptr[l] = at >= N ? n_internal : (int) (scan[at] & 0xffffffffULL);
}

}  // namespace

size_t topo_scan_bytes(int n_nodes) {
    size_t bytes = 0;
    cub::TransformInputIterator<unsigned long long, PackCount, const int *> in(nullptr, PackCount());
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, (unsigned long long *) nullptr, n_nodes);
    return bytes;
}

void topo_stage1(cudaStream_t stream, int N, const int *parent, const double *dist, int *n_children,
                 unsigned long long *scan, void *scan_tmp, size_t scan_tmp_bytes, int *flags_ni) {
    cudaMemsetAsync(n_children, 0, sizeof(int) * (size_t) N, stream);
    cudaMemsetAsync(flags_ni, 0, 2 * sizeof(int), stream);
    topo_count_kernel<<<topo_grid(N), kTopoBlock, 0, stream>>>(N, parent, dist, n_children, flags_ni);
    cub::TransformInputIterator<unsigned long long, PackCount, const int *> in(n_children, PackCount());
    cub::DeviceScan::ExclusiveSum(scan_tmp, scan_tmp_bytes, in, scan, N, stream);
    topo_total_kernel<<<1, 1, 0, stream>>>(N, n_children, scan, flags_ni);
}

void topo_stage2(cudaStream_t stream, int N, int n_roots, int n_internal, const int *parent, const double *dist,
                 const unsigned long long *scan, const TopoOut &out, const int *level_ptr, int n_level_ptr,
                 int *int_level_ptr) {
    topo_fill_kernel<<<topo_grid(N), kTopoBlock, 0, stream>>>(N, n_roots, parent, dist, scan, out);
    topo_levels_kernel<<<topo_grid(n_level_ptr), kTopoBlock, 0, stream>>>(N, n_level_ptr, level_ptr, scan, n_internal,
                                                                          int_level_ptr);
}

}  // namespace pml
