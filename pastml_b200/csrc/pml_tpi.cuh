// Thread-per-item sweeps with TMA-staged operands for the F81 family (F81 / JC / EFT).
//
// One thread owns one (node, lane) item and walks its K states sequentially, so nothing is replicated over a group
// of threads and no reduction crosses threads -- about a third of the warp instructions of the cooperative kernels
// (pml_coop.cuh).  What made thread-per-item slow before (pml_kernels.cuh: three K-vectors in registers, 180+
// registers, dependent loads with few bytes in flight) is gone because the operands never sit in registers: on
// the [slot][k][lane] layout the vector of a node for JT lanes is ONE contiguous run of K * JT * 8 bytes, which one
// thread per node fetches with a bulk asynchronous copy (cp.async.bulk, completion on a per-warp mbarrier) into the
// warp's staging slot; the items then stream over shared memory twice (weighted sum, then the product).  A warp keeps
// 2 * 32 * (8 K + 4) bytes in flight (10.5 KB for K = 20) against 2.5 KB for the register kernels, which is what
// the HBM latency-bandwidth product asks for at 16 warps per SM.
//
// Lanes: JT lanes of a tile are adjacent threads, 32 / JT nodes share a warp.  More lanes than JT (parameter sets x
// columns of the batched finite-difference objective): the pass uses the tile-major layout [slot][tile][k][JT]
// (identical to the plain layout when there is one tile), blockIdx.y selects the tile.
#pragma once
#include "pml_coop.cuh"

namespace pml {

#ifndef PML_TPI_BLOCK
#define PML_TPI_BLOCK 128
#endif
#ifndef PML_TPI_BU_MINB
#define PML_TPI_BU_MINB 4
#endif

__host__ __device__ constexpr int tpi_bu_sub_bytes(int K, int JT) { return 2 * (K * JT * 8 + JT * 4); }
inline size_t tpi_bu_smem(int K, int JT, int block) {
    return (size_t) round16((4 + K) * JT * 8) + (size_t) (block / 32) * (32 / JT) * tpi_bu_sub_bytes(K, JT) + 8 * (block / 32);
}

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// power-of-two rescale of one value with the factors exponent_rescale<> would use for a vector whose maximum has
// exponent `e` (subnormal maxima are lifted by 2^200 first)
struct Rescale {
    double lift, s1, s2;
    int removed;
    __device__ __forceinline__ double apply(double v) const { return ((v * lift) * s1) * s2; }
};
__device__ __forceinline__ bool needs_rescale(double vmax) {
    if (!(vmax > 0.) || isinf(vmax)) return false;
    const int e = exponent_of(vmax);
    return e < -kKeep || e > kKeep;
}
__device__ __forceinline__ Rescale make_rescale(double vmax) {
    Rescale rs{1., 1., 1., 0};
    int e = exponent_of(vmax);
    int extra = 0;
    if (e == -1023) {
        rs.lift = 0x1p+200;
        e = exponent_of(vmax * 0x1p+200);
        extra = -200;
    }
    const int e1 = e / 2, e2 = e - e1;
    rs.s1 = pow2(-e1);
    rs.s2 = pow2(-e2);
    rs.removed = e + extra;
    return rs;
}

// calc_node_bu_likelihood (ml.py:124-148) for the internal nodes [slot_begin, slot_begin + n_slots) of one level,
// marginal (sum-product) mode.
template <int JT>
__global__ void __launch_bounds__(PML_TPI_BLOCK, PML_TPI_BU_MINB) bu_tpi_kernel(BuArgs a, int slot_begin, int n_slots) {
    extern __shared__ double s_par[];
    constexpr int NPW = 32 / JT;
    const int K = a.ln.K, C = a.ln.C;
    const int T = gridDim.y, tile = blockIdx.y, j0 = tile * JT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / JT, jl = lane - sub * JT;
    const int j = j0 + jl, c = lane_column(a.ln, j);
    const int VB = K * JT * 8, EB = JT * 4, sub_bytes = 2 * (VB + EB);
    unsigned char *st_base = reinterpret_cast<unsigned char *>(s_par) + round16((4 + K) * JT * 8);
    unsigned char *mine = st_base + (size_t) (warp * NPW + sub) * sub_bytes;
    const uint32_t bar = smem_addr(st_base + (size_t) (blockDim.x >> 5) * NPW * sub_bytes + 8 * warp);
    const uint32_t stage = smem_addr(mine);
    double *in0 = reinterpret_cast<double *>(mine) + jl;   // child 0, state k at in0[k * JT]
    double *in1 = in0 + K * JT;                            // child 1
    const int *sex = reinterpret_cast<const int *>(mine + 2 * VB);   // exponents: [child][JT]
    if (lane == 0) {
        mbar_init(bar, NPW);
        mbar_fence_init();
    }
    {
        // lane parameters of this tile, transposed to [field][lane]
        const int total = (4 + K) * JT;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int f = i / JT, jj = i - f * JT;
            s_par[i] = a.ln.params[(size_t) (j0 + jj) * a.ln.stride + f];
        }
        __syncthreads();
    }
    const double sf = s_par[jl], tau = s_par[JT + jl], tauf = s_par[2 * JT + jl], mu = s_par[3 * JT + jl];
    const double *spi = s_par + 4 * JT + jl;   // pi_k at spi[k * JT]
    const bool use_alt = a.alter && tau == 0.;
    const bool mu_inf = isinf(mu);
    const int warp_global = blockIdx.x * (blockDim.x >> 5) + warp;
    const int n_warps = gridDim.x * (blockDim.x >> 5);
    const size_t vec_stride = (size_t) K * JT;   // doubles between (slot, tile) vectors

    const int stride = n_warps * NPW;
    int base = warp_global * NPW;
    int4 cu_top = make_int4(0, 0, 0, 0), cu_len = make_int4(0, 0, 0, 0), nx_top = cu_top, nx_len = cu_len;
    int cu_node = 0, nx_node = 0;
    uint64_t cu_mo = 0, cu_m0 = 0, cu_m1 = 0;
    auto load_desc = [&](int b, int4 &top, int4 &len, int &node) {
        const int sl = b + sub;
        if (sl < n_slots) {
            const size_t slot = (size_t) slot_begin + sl;
            top = a.t.bu_desc[2 * slot];
            len = a.t.bu_desc[2 * slot + 1];
            node = a.t.int_nodes[slot];
        } else {
            top = make_int4(0, 0, -1, -1);
        }
    };
    auto load_masks = [&](int b, const int4 &top, int node, uint64_t &mo, uint64_t &m0, uint64_t &m1) {
        if (b + sub >= n_slots) return;
        mo = a.mask[(size_t) node * C + c];
        m0 = (top.y == 2 && top.z < 0) ? a.mask[(size_t) top.x * C + c] : 0ULL;
        m1 = (top.y == 2 && top.w < 0) ? a.mask[(size_t) (top.x + 1) * C + c] : 0ULL;
    };
    // one thread per node arms the barrier and issues the copies of the node's two child vectors + exponent rows
    auto issue = [&](const int4 &top) {
        if (jl != 0) return;
        const bool bin = top.y == 2;
        const int cs0 = top.z, cs1 = top.w;
        mbar_arrive_expect(bar, bin ? ((cs0 >= 0 ? VB + EB : 0) + (cs1 >= 0 ? VB + EB : 0)) : 0);
        if (!bin) return;
        if (cs0 >= 0) {
            bulk_g2s(stage, a.bu + ((size_t) cs0 * T + tile) * vec_stride, VB, bar);
            bulk_g2s(stage + 2 * VB, a.bu_ex + ((size_t) cs0 * T + tile) * JT, EB, bar);
        }
        if (cs1 >= 0) {
            bulk_g2s(stage + VB, a.bu + ((size_t) cs1 * T + tile) * vec_stride, VB, bar);
            bulk_g2s(stage + 2 * VB + EB, a.bu_ex + ((size_t) cs1 * T + tile) * JT, EB, bar);
        }
    };
    uint32_t phase = 0;
    pdl_launch_dependents();
    if (base < n_slots) {
        load_desc(base, cu_top, cu_len, cu_node);
        load_masks(base, cu_top, cu_node, cu_mo, cu_m0, cu_m1);
        if (base + stride < n_slots) load_desc(base + stride, nx_top, nx_len, nx_node);
        pdl_wait();   // everything above (topology, annotation masks) is independent of the previous level
        issue(cu_top);
    }
    for (; base < n_slots; base += stride) {
        const bool on = base + sub < n_slots;
        const int slot = slot_begin + base + sub;
        const int4 d_top = cu_top, d_len = cu_len;
        const int node = cu_node;
        const int fc = d_top.x, nc = d_top.y;
        const bool bin = nc == 2;
        const uint64_t own = eff_from_raw(cu_mo, a.t, a.zc, C, node, c, use_alt, a.ln.kmask);
        // the exponentials of the two branches do not need the vectors: evaluate them while the copies land
        double e0 = 0., e1 = 0.;
        uint64_t m0 = 0ULL, m1 = 0ULL;
        if (on && bin) {
            if (!mu_inf) {
                e0 = exp_nonpos(-mu * ((__hiloint2double(d_len.y, d_len.x) + tau) * tauf * sf));
                e1 = exp_nonpos(-mu * ((__hiloint2double(d_len.w, d_len.z) + tau) * tauf * sf));
            }
            if (d_top.z < 0) m0 = eff_from_raw(cu_m0, a.t, a.zc, C, fc, c, use_alt, a.ln.kmask);
            if (d_top.w < 0) m1 = eff_from_raw(cu_m1, a.t, a.zc, C, fc + 1, c, use_alt, a.ln.kmask);
        }
        // requests for the batches behind this one go out before the arithmetic: mask words of the next batch (its
        // descriptor is here already), descriptor of the one after
        const bool has_next = base + stride < n_slots;
        uint64_t n_mo = 0, n_m0 = 0, n_m1 = 0;
        int4 n2_top = make_int4(0, 0, -1, -1), n2_len = make_int4(0, 0, 0, 0);
        int n2_node = 0;
        if (has_next) {
            load_masks(base + stride, nx_top, nx_node, n_mo, n_m0, n_m1);
            if (base + 2 * stride < n_slots) load_desc(base + 2 * stride, n2_top, n2_len, n2_node);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        double *dst = a.bu + ((size_t) slot * T + tile) * vec_stride + jl;
        int ex = 0;
        int bad_child = -1;
        if (on && bin) {
            // tip children: the vector is the allowed-state mask, written into the (unused) staging column so that the
            // two walks below are branch-free
            if (d_top.z < 0) {
                uint64_t m = m0;
                for (int k = 0; k < K; ++k, m >>= 1) in0[k * JT] = __hiloint2double((m & 1ULL) ? 0x3ff00000 : 0, 0);
            } else {
                ex += sex[jl];
            }
            if (d_top.w < 0) {
                uint64_t m = m1;
                for (int k = 0; k < K; ++k, m >>= 1) in1[k * JT] = __hiloint2double((m & 1ULL) ? 0x3ff00000 : 0, 0);
            } else {
                ex += sex[JT + jl];
            }
            // first walk: frequency-weighted sums of the two child vectors (two chains each for the FMA latency)
            double s0a = 0., s0b = 0., s1a = 0., s1b = 0.;
            int k = 0;
#pragma unroll 2
            for (; k + 1 < K; k += 2) {
                const double pa = spi[k * JT], pb = spi[(k + 1) * JT];
                s0a = fma(pa, in0[k * JT], s0a);
                s0b = fma(pb, in0[(k + 1) * JT], s0b);
                s1a = fma(pa, in1[k * JT], s1a);
                s1b = fma(pb, in1[(k + 1) * JT], s1b);
            }
            if (k < K) {
                const double pa = spi[k * JT];
                s0a = fma(pa, in0[k * JT], s0a);
                s1a = fma(pa, in1[k * JT], s1a);
            }
            const double b0 = (1. - e0) * (s0a + s0b), b1 = (1. - e1) * (s1a + s1b);
            // second walk: acc_k = allowed_k * (e0 x0_k + b0) * (e1 x1_k + b1), written straight to memory
            double vmax = 0.;
            uint64_t ow = own;
#pragma unroll 4
            for (k = 0; k < K; ++k, ow >>= 1) {
                double v = fma(e0, in0[k * JT], b0);   // >= 0, the reference's np.maximum(., 0) is a no-op here
                v = (ow & 1ULL) ? v : 0.;              // 10 ** log10(allowed), ml.py:127
                v *= fma(e1, in1[k * JT], b1);
                vmax = fmax(vmax, v);
                dst[k * JT] = v;
            }
            if (vmax == 0.) {
                // all entries zero (PastMLLikelihoodError of ml.py:139-145): which child emptied the vector
                double after0 = 0.;
                for (k = 0; k < K; ++k) after0 = fmax(after0, bit_as_double(own, k) * fma(e0, in0[k * JT], b0));
                bad_child = (after0 == 0.) ? fc : fc + 1;
            } else if (needs_rescale(vmax)) {
                // rare: the maximum left the exponent window, rewrite the vector scaled by an exact power of two
                const Rescale rs = make_rescale(vmax);
                for (k = 0; k < K; ++k) {
                    double v = bit_as_double(own, k);
                    v *= fma(e0, in0[k * JT], b0);
                    v *= fma(e1, in1[k * JT], b1);
                    dst[k * JT] = rs.apply(v);
                }
                ex += rs.removed;
            }
        } else if (on) {
            // any other number of children: the node's staging column holds the running product, the children are read
            // from memory one after the other (no copies were issued for this node)
            double *col = in0;
            for (int k = 0; k < K; ++k) col[k * JT] = bit_as_double(own, k);
            for (int q = 0; q < nc; ++q) {
                const int ch = fc + q;
                const int cs = a.t.slot[ch];
                const double *src = (cs >= 0) ? a.bu + ((size_t) cs * T + tile) * vec_stride + jl : nullptr;
                uint64_t m = 0ULL;
                if (cs < 0) m = eff_mask(a.mask, a.t, a.zc, C, ch, c, use_alt, a.ln.kmask);
                const double e = mu_inf ? 0. : exp_nonpos(-mu * ((a.t.dist[ch] + tau) * tauf * sf));
                double s = 0.;
                for (int k = 0; k < K; ++k) s = fma(spi[k * JT], src ? src[k * JT] : bit_as_double(m, k), s);
                const double b = (1. - e) * s;
                double mx = 0.;
                for (int k = 0; k < K; ++k) {
                    const double v = col[k * JT] * fma(e, src ? src[k * JT] : bit_as_double(m, k), b);
                    col[k * JT] = v;
                    mx = fmax(mx, v);
                }
                if (cs >= 0) ex += a.bu_ex[((size_t) cs * T + tile) * JT + jl];
                if (mx == 0. && bad_child < 0) bad_child = ch;
                if (needs_rescale(mx)) {
                    const Rescale rs = make_rescale(mx);
                    for (int k = 0; k < K; ++k) col[k * JT] = rs.apply(col[k * JT]);
                    ex += rs.removed;
                }
            }
            for (int k = 0; k < K; ++k) dst[k * JT] = col[k * JT];
        }
        if (on) {
            if (bad_child >= 0) {
                if (atomicCAS(a.err + 2 * j, -1, node) == -1) a.err[2 * j + 1] = bad_child;
            }
            a.bu_ex[((size_t) slot * T + tile) * JT + jl] = ex;
        }
        // the staging slot is free again: fetch the next batch (generic-proxy writes to the slot, tip masks or the
        // running product, are ordered before the copies of the async proxy)
        fence_proxy_async();
        __syncwarp();
        if (has_next) {
            issue(nx_top);
            cu_top = nx_top;
            cu_len = nx_len;
            cu_node = nx_node;
            cu_mo = n_mo;
            cu_m0 = n_m0;
            cu_m1 = n_m1;
            nx_top = n2_top;
            nx_len = n2_len;
            nx_node = n2_node;
        }
    }
}

}  // namespace pml
