// Warp-cooperative sweeps for the F81 family (F81 / JC / EFT), the headline path.
//
// A (node, lane) item is shared by G threads of one warp, each holding KT states (k = r + G i, r = rank in the
// group): the register footprint drops from ~3 K doubles to ~3 KT, so 5-6 blocks of 128 threads stay resident per SM
// instead of 2-3, and on the [slot][k][lane] layout the G x J threads of a node read and write one contiguous run
// (256 B per request for K = 20 states x 8 lanes).  The only cross-thread quantities of the F81 recursion are
// sum_k pi_k L_k and max_k L_k, two butterfly reductions over the group.  Used when G * J <= 32 (otherwise the
// thread-per-item kernels of pml_kernels.cuh already coalesce over the lanes).
#pragma once
#include "pml_kernels.cuh"

namespace pml {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kCoopBlock = 128;

// Reductions over the G threads that share one (node, lane) item.  Lane layout inside a warp:
// lane = sub * (G * J) + r * J + j.  When J is a power of two (PJ) the partner of a butterfly step is
// lane ^ (step * J): SHFL.BFLY does the address arithmetic for free; otherwise the partner lane is computed.
template <int G, bool PJ>
struct GroupOps {
    static constexpr int kSteps = (G >= 32) ? 5 : (G >= 16) ? 4 : (G >= 8) ? 3 : (G >= 4) ? 2 : (G >= 2) ? 1 : 0;
    int lane, r, J;
    __device__ __forceinline__ GroupOps(int lane_, int r_, int J_) : lane(lane_), r(r_), J(J_) {}
    template <class T>
    __device__ __forceinline__ T xchg(T v, int q) const {
        if (PJ) return __shfl_xor_sync(kFullMask, v, (1 << q) * J);
        return __shfl_sync(kFullMask, v, lane + ((r ^ (1 << q)) - r) * J);
    }
    __device__ __forceinline__ double sum(double v) const {
#pragma unroll
        for (int q = 0; q < kSteps; ++q) v += xchg(v, q);
        return v;
    }
    __device__ __forceinline__ double max(double v) const {
#pragma unroll
        for (int q = 0; q < kSteps; ++q) v = fmax(v, xchg(v, q));
        return v;
    }
    __device__ __forceinline__ unsigned long long bit_or(unsigned long long v) const {
#pragma unroll
        for (int q = 0; q < kSteps; ++q) v |= xchg(v, q);
        return v;
    }
    // arg max with the reference's tie rule (first index): larger value wins, equal values -> smaller index
    __device__ __forceinline__ void argmax(double &v, int &k) const {
#pragma unroll
        for (int q = 0; q < kSteps; ++q) {
            const double ov = xchg(v, q);
            const int ok = xchg(k, q);
            if (ov > v || (ov == v && ok < k)) { v = ov; k = ok; }
        }
    }
    // value held by the group member of rank `src`
    __device__ __forceinline__ double from_rank(double v, int src) const {
        return __shfl_sync(kFullMask, v, lane + (src - r) * J);
    }
};

// exp(x) for x <= 0 (every use here is exp(-mu t), t >= 0): Cody-Waite reduction x = n ln2 + r, |r| <= ln2 / 2,
// degree-13 Taylor polynomial (truncation 4e-18), exact scaling by 2^n through the exponent field.  About a third
// of the instructions of the library exp (no special-case paths: results below 2^-1021 flush to zero).
__device__ __forceinline__ double exp_nonpos(double x) {
    if (!(x > -708.)) return (x != x) ? x : 0.;
    const double t = fma(x, 1.4426950408889634074, 6755399441055744.0);   // round(x / ln2) in the low mantissa bits
    const int n = __double2loint(t);
    const double nf = t - 6755399441055744.0;
    double r = fma(nf, -6.93147180369123816490e-01, x);
    r = fma(nf, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;            // 1/13!
    p = fma(p, r, 2.08767569878681e-09);          // 1/12!
    p = fma(p, r, 2.505210838544172e-08);         // 1/11!
    p = fma(p, r, 2.755731922398589e-07);         // 1/10!
    p = fma(p, r, 2.7557319223985893e-06);        // 1/9!
    p = fma(p, r, 2.48015873015873e-05);          // 1/8!
    p = fma(p, r, 1.984126984126984e-04);         // 1/7!
    p = fma(p, r, 1.388888888888889e-03);         // 1/6!
    p = fma(p, r, 8.333333333333333e-03);         // 1/5!
    p = fma(p, r, 4.1666666666666664e-02);        // 1/4!
    p = fma(p, r, 1.6666666666666666e-01);        // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.);
    p = fma(p, r, 1.);
    return __hiloint2double(__double2hiint(p) + n * 1048576, __double2loint(p));
}

// 1 / x for a normal positive x: hardware approximation (2^-23) + two Newton steps, ~1 ulp.  The fused top-down
// step divides K values per (node, lane) item; __drcp_rn's exact rounding and special-case handling cost three
// times as many instructions and buy nothing at the 1e-10 tolerance.  Tiny or huge x take the exact path.
__device__ __forceinline__ double fast_rcp(double x) {
    if (!(x > 0x1p-900 && x < 0x1p+900)) return __drcp_rn(x);
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double t = fma(-x, r, 1.);
    r = fma(r, t, r);
    t = fma(-x, r, 1.);
    return fma(r, t, r);
}

// ------------------------------------------------------------------------------------------ TMA staging
//
// The TMA variants of the sweeps (template flag TMA) take the K-vectors of the NEXT node of a warp's queue with bulk
// asynchronous copies (cp.async.bulk, global -> shared, completion counted on a per-warp mbarrier) while the warp
// computes the current node out of registers: on the [slot][k][lane] layout the whole vector of a node for all the
// lanes is ONE contiguous run (K * J * 8 bytes, 1 280 B for K = 20 x 8 lanes), so a step needs 2-3 copies issued by
// one thread instead of 2-3 x KT dependent loads in every thread, and no vector load is left on the dependent path.
// One staging slot per warp: it is drained into registers at the top of the step and refilled right away.
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__host__ __device__ constexpr int round16(int x) { return (x + 15) & ~15; }


// dynamic shared memory of the TMA variants (host side of the layout the kernels carve up)
__host__ __device__ constexpr int td_tma_sub_bytes(int K, int C) { return 3 * (K * C * 8 + C * 4); }
inline size_t td_tma_smem(int K, int C, int G) {
    return (size_t) round16((4 + K) * C * 8 + C * (K + 1) * 4) + (size_t) (kCoopBlock / 32) * (32 / (G * C)) * td_tma_sub_bytes(K, C) +
           8 * (kCoopBlock / 32);
}

struct CoopLane {   // where a thread sits: which node of the warp's batch, which lane (column), which rank
    int sub, r, j, c, npw;
    bool lane_on;
};

#ifndef PML_COOP_MINB
#define PML_COOP_MINB 8
#endif

// resident blocks per SM the kernels are compiled for: 64 registers per thread up to 5 states per thread (measured
// best for K = 20: more warps in flight beat the few spills), fewer blocks when a thread holds more states
constexpr int coop_min_blocks(int kt) { return kt <= 5 ? PML_COOP_MINB : kt <= 8 ? 5 : 3; }
// the TMA variants keep the next batch's descriptors and mask words in registers and do not need as many warps to
// cover the memory latency (the copies run ahead of the arithmetic): fewer resident blocks, more registers each
#ifndef PML_TMA_MINB
#define PML_TMA_MINB 5
#endif
#ifndef PML_TMA_MINB_WIDE
#define PML_TMA_MINB_WIDE 4
#endif
constexpr int coop_min_blocks(int kt, bool tma) { return tma ? (kt <= 5 ? PML_TMA_MINB : kt <= 8 ? 4 : PML_TMA_MINB_WIDE) : coop_min_blocks(kt); }

template <int G>
__device__ __forceinline__ CoopLane coop_lane(int J, int C) {
    const int lane = threadIdx.x & 31;
    const int gj = G * J;
    CoopLane cl;
    cl.npw = 32 / gj;
    cl.sub = lane / gj;
    const int rem = lane - cl.sub * gj;
    cl.r = rem / J;
    cl.j = rem - cl.r * J;
    cl.c = cl.j % C;
    cl.lane_on = cl.sub < cl.npw;
    return cl;
}

template <int KT, int G>
__device__ __forceinline__ int exponent_rescale(double (&v)[KT], double vmax, bool always) {
    // same policy as rescale<> of the thread-per-item kernels; vmax is the group-wide maximum
    if (!(vmax > 0.) || isinf(vmax)) return 0;
    int e = exponent_of(vmax);
    if (!always && e >= -kKeep && e <= kKeep) return 0;
    int extra = 0;
    if (e == -1023) {
#pragma unroll
        for (int i = 0; i < KT; ++i) v[i] *= 0x1p+200;
        e = exponent_of(vmax * 0x1p+200);
        extra = -200;
    }
    if (e != 0) {
        const int e1 = e / 2, e2 = e - e1;
        const double s1 = pow2(-e1), s2 = pow2(-e2);
#pragma unroll
        for (int i = 0; i < KT; ++i) v[i] = (v[i] * s1) * s2;
    }
    return e + extra;
}

// loads this thread's KT states of a node's bottom-up vector (or of its tip mask); returns the exponent
// state index r + G i is a real state (always true when K == G * KT, which the KX instantiations know)
#define PML_VALID(i) (KX || (r + G * (i) < K))

// K and J are passed explicitly so that the instantiations with compile-time K / J fold the address arithmetic
template <int KT, int G, bool KX>
__device__ __forceinline__ int coop_load(const TreeDev &t, const ZcDev &zc, const ExtDev &ext, const LaneDev &ln,
                                         const uint64_t *mask, const double *bu, const int *bu_ex, int node, int cs,
                                         int r, int j, int c, bool use_alt, double (&in)[KT], int K, int J) {
    if (cs >= 0) {
        const double *src = bu + ((size_t) (cs * K + r) * J + j);
#pragma unroll
        for (int i = 0; i < KT; ++i) in[i] = PML_VALID(i) ? src[(G * i) * J] : 0.;
        return bu_ex[(size_t) cs * J + j];
    }
    const uint64_t raw = mask[(size_t) node * ln.C + c];
    if (raw & kExtFlag) {   // a clade root cut off by subtree sharding: its vector comes from the ExtDev tables
        const int e = ext.index[node];
        const double *src = ext.bu + ((size_t) (e * K + r) * J + j);
#pragma unroll
        for (int i = 0; i < KT; ++i) in[i] = PML_VALID(i) ? src[(G * i) * J] : 0.;
        return ext.bu_ex[(size_t) e * J + j];
    }
    uint64_t m = raw;
    if (use_alt && (m & kMemberFlag)) m = zc.alt[(size_t) t.zc_index[node] * ln.C + c];
    m = (m & ln.kmask) >> r;
#pragma unroll
    for (int i = 0; i < KT; ++i) in[i] = bit_as_double(m, G * i);
    return 0;
}

// ------------------------------------------------------------------------------------------ bottom-up

// effective allowed-state mask from the raw annotation word (eff_mask with the word already loaded)
__device__ __forceinline__ uint64_t eff_from_raw(uint64_t raw, const TreeDev &t, const ZcDev &zc, int C, int node, int c,
                                                 bool use_alt, uint64_t kmask) {
    if (use_alt && (raw & kMemberFlag)) raw = zc.alt[(size_t) t.zc_index[node] * C + c];
    return raw & kmask;
}

// JT > 0: the number of lanes J is the compile-time constant JT; KX: K == G * KT exactly.
// TILED: the pass has more lanes than one warp batch can take (many parameter sets x columns): blockIdx.y selects a
// tile of JT consecutive lanes, memory keeps its full leading dimension `ld` = total lanes.
// (Problems without external tables take the TMA-staged thread-per-item sweep of pml_tpi.cuh instead; this kernel
// serves the subtree-sharded problems and the lane counts the staged sweep does not tile.)
template <int KT, int G, int JT, bool KX, bool TILED>
__global__ void __launch_bounds__(kCoopBlock, coop_min_blocks(KT)) bu_coop_kernel(BuArgs a, int slot_begin, int n_slots) {
    extern __shared__ double s_par[];
    const int J = JT ? JT : a.ln.J, K = KX ? G * KT : a.ln.K, C = a.ln.C;
    const int ld = TILED ? a.ln.J : J;                 // leading dimension of bu / bu_ex / ext tables
    const int j0 = TILED ? (int) blockIdx.y * J : 0;   // first lane of this block's tile
    {
        // stage the tile's lane parameters, transposed to [field][lane]
        const int total = (4 + K) * J;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int f = i / J, jj = i - f * J;
            s_par[i] = a.ln.params[(size_t) (j0 + jj) * a.ln.stride + f];
        }
        __syncthreads();
    }
    const CoopLane cl = coop_lane<G>(J, C);
    const int r = cl.r;
    const int j = j0 + cl.j;                 // lane index in memory
    const int c = TILED ? j % C : cl.c;
    const GroupOps<G, (JT > 0 && (JT & (JT - 1)) == 0)> grp((int) (threadIdx.x & 31), r, J);
    const int jl = cl.j;                     // lane index inside the tile (shared-memory column)
    const double sf = s_par[jl], tau = s_par[J + jl], tauf = s_par[2 * J + jl], mu = s_par[3 * J + jl];
    double pi[KT];
#pragma unroll
    for (int i = 0; i < KT; ++i) pi[i] = PML_VALID(i) ? s_par[(4 + r + G * i) * J + jl] : 0.;
    const bool use_alt = a.alter && tau == 0.;
    const bool mu_inf = isinf(mu);
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;

    // the loop is software-pipelined by one step: the packed descriptor (and node id) of the next batch is requested
    // before the current batch is processed, so that only the vector loads remain on the dependent path
    const int stride = n_warps * cl.npw;
    int base = warp_global * cl.npw;
    int4 nx_top = make_int4(0, 0, 0, 0), nx_len = make_int4(0, 0, 0, 0);
    int nx_node = 0;
    auto load_desc = [&](int b, int4 &top, int4 &len, int &node) {
        const int sl = b + cl.sub;
        const size_t slot = (size_t) slot_begin + ((cl.lane_on && sl < n_slots) ? sl : b);
        top = a.t.bu_desc[2 * slot];
        len = a.t.bu_desc[2 * slot + 1];
        node = a.t.int_nodes[slot];
    };
    pdl_launch_dependents();
    if (base < n_slots) load_desc(base, nx_top, nx_len, nx_node);
    pdl_wait();   // everything above is independent of the previous level
    for (; base < n_slots; base += stride) {
        const int sl = base + cl.sub;
        const bool on = cl.lane_on && sl < n_slots;
        const int slot = slot_begin + (on ? sl : base);
        const int4 d_top = nx_top, d_len = nx_len;
        const int node = nx_node;
        if (base + stride < n_slots) load_desc(base + stride, nx_top, nx_len, nx_node);
        const int fc = d_top.x, nc = d_top.y;
        double acc[KT];
        {
            const uint64_t m = eff_mask(a.mask, a.t, a.zc, C, node, c, use_alt, a.ln.kmask) >> r;
#pragma unroll
            for (int i = 0; i < KT; ++i) acc[i] = bit_as_double(m, G * i);   // 10 ** log10(allowed), ml.py:127
        }
        int ex = 0;
        int bad_child = -1;
        if (__all_sync(kFullMask, nc == 2)) {
            // binary nodes: both children's vectors are requested before any arithmetic; rank 0 and rank 1 of the
            // group evaluate one exp each and hand it to the others
            double in0[KT], in1[KT];
            const int cs0 = d_top.z, cs1 = d_top.w;
            const double dmine = (r & 1) ? __hiloint2double(d_len.w, d_len.z) : __hiloint2double(d_len.y, d_len.x);
            ex += coop_load<KT, G, KX>(a.t, a.zc, a.ext, a.ln, a.mask, a.bu, a.bu_ex, fc, cs0, r, j, c, use_alt, in0, K, ld);
            ex += coop_load<KT, G, KX>(a.t, a.zc, a.ext, a.ln, a.mask, a.bu, a.bu_ex, fc + 1, cs1, r, j, c, use_alt, in1, K, ld);
            const double emine = mu_inf ? 0. : exp_nonpos(-mu * ((dmine + tau) * tauf * sf));
            const double e0 = grp.from_rank(emine, 0), e1 = grp.from_rank(emine, 1);
            double s0 = 0., s1 = 0.;
#pragma unroll
            for (int i = 0; i < KT; ++i) { s0 = fma(pi[i], in0[i], s0); s1 = fma(pi[i], in1[i], s1); }
            s0 = grp.sum(s0);
            s1 = grp.sum(s1);
            const double b0 = (1. - e0) * s0, b1 = (1. - e1) * s1;
            double m0 = 0., m1 = 0.;
#pragma unroll
            for (int i = 0; i < KT; ++i) {
                acc[i] *= fma(e0, in0[i], b0);   // >= 0, the reference's np.maximum(., 0) is a no-op here
                m0 = fmax(m0, acc[i]);
                acc[i] *= fma(e1, in1[i], b1);
                m1 = fmax(m1, acc[i]);
            }
            m1 = grp.max(m1);
            if (__any_sync(kFullMask, m1 == 0.)) {   // error path only: which child emptied the vector
                m0 = grp.max(m0);
                if (m1 == 0.) bad_child = (m0 == 0.) ? fc : fc + 1;
            }
            ex += exponent_rescale<KT, G>(acc, m1, false);
        } else {
            const int ncmax = __reduce_max_sync(kFullMask, nc);
            for (int q = 0; q < ncmax; ++q) {
                const bool valid = q < nc;
                const int ch = fc + (valid ? q : 0);
                double in[KT];
                const int exc = coop_load<KT, G, KX>(a.t, a.zc, a.ext, a.ln, a.mask, a.bu, a.bu_ex, ch, a.t.slot[ch], r, j, c,
                                                 use_alt, in, K, ld);
                const double e = mu_inf ? 0. : exp_nonpos(-mu * ((a.t.dist[ch] + tau) * tauf * sf));
                double s = 0.;
#pragma unroll
                for (int i = 0; i < KT; ++i) s = fma(pi[i], in[i], s);
                s = grp.sum(s);
                const double b = (1. - e) * s;
                double mx = 0.;
#pragma unroll
                for (int i = 0; i < KT; ++i) {
                    if (valid) acc[i] *= fma(e, in[i], b);
                    mx = fmax(mx, acc[i]);
                }
                mx = grp.max(mx);
                if (valid) {
                    ex += exc;
                    if (mx == 0. && bad_child < 0) bad_child = ch;
                    ex += exponent_rescale<KT, G>(acc, mx, false);
                }
            }
        }
        if (on) {
            if (bad_child >= 0 && r == 0) {   // all entries zero: PastMLLikelihoodError of ml.py:139-145
                if (atomicCAS(a.err + 2 * j, -1, node) == -1) a.err[2 * j + 1] = bad_child;
            }
            double *dst = a.bu + ((size_t) (slot * K + r) * ld + j);
#pragma unroll
            for (int i = 0; i < KT; ++i)
                if (PML_VALID(i)) dst[(G * i) * ld] = acc[i];
            if (r == 0) a.bu_ex[(size_t) slot * ld + j] = ex;
        }
    }
}

// ------------------------------------------------------------------------------------------ top-down sweep

// node-centric fused top-down / marginal / MAP / MPPA step (see td_level_kernel) with the K-vector split over G threads.
// TMA: the node's own bottom-up vector and its parent's bottom-up and top-down vectors are staged one step ahead by
// bulk copies (three contiguous K * C * 8-byte runs), the mask word and the joint state are prefetched in registers.
template <int KT, int G, int JT, bool KX, bool TMA = false>
__global__ void __launch_bounds__(kCoopBlock, coop_min_blocks(KT, TMA)) td_coop_kernel(TdArgs a, int node_begin, int n_nodes) {
    extern __shared__ double s_par[];
    static_assert(!TMA || (JT > 0 && KX && (32 % (G * (JT ? JT : 1))) == 0), "TMA staging needs compile-time lanes");
    const int C = JT ? JT : a.ln.C, K = KX ? G * KT : a.ln.K;
    const int nf = 4 + K;
    unsigned int *sh_hist = reinterpret_cast<unsigned int *>(s_par + (size_t) nf * C);   // [C][K+1]
    for (int i = threadIdx.x; i < C * (K + 1); i += blockDim.x) sh_hist[i] = 0u;
    const CoopLane cl = coop_lane<G>(C, C);
    constexpr int VB = G * KT * (JT ? JT : 1) * 8, EB = (JT ? JT : 1) * 4;   // bytes of one staged vector / exponent row
    uint32_t bar = 0, stage = 0, phase = 0;
    const double *st_vec = nullptr;
    if constexpr (TMA) {
        unsigned char *st_base = reinterpret_cast<unsigned char *>(s_par) + round16(nf * C * 8 + C * (K + 1) * 4);
        constexpr int sub_bytes = td_tma_sub_bytes(G * KT, JT);
        constexpr int npw = 32 / (G * JT);
        const int warp = threadIdx.x >> 5;
        unsigned char *mine = st_base + (warp * npw + cl.sub) * sub_bytes;
        bar = smem_addr(st_base + (kCoopBlock / 32) * npw * sub_bytes + 8 * warp);
        stage = smem_addr(mine);
        st_vec = reinterpret_cast<const double *>(mine);
        if ((threadIdx.x & 31) == 0) {
            mbar_init(bar, npw);
            mbar_fence_init();
        }
    }
    stage_params(s_par, a.ln, nf);
    const int r = cl.r, c = cl.j;
    const GroupOps<G, (JT > 0 && (JT & (JT - 1)) == 0)> grp((int) (threadIdx.x & 31), r, C);
    const double sf = s_par[c], tau = s_par[C + c], tauf = s_par[2 * C + c], mu = s_par[3 * C + c];
    double pi[KT];
#pragma unroll
    for (int i = 0; i < KT; ++i) pi[i] = PML_VALID(i) ? s_par[(4 + r + G * i) * C + c] : 0.;
    const bool use_alt = tau == 0.;   // ml.py:702-703
    const bool mu_inf = isinf(mu);
    const bool fj = a.force_joint != 0;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const int n_root_slots = a.t.n_root_slots;

    const int stride = n_warps * cl.npw;
    int base = warp_global * cl.npw;
    int4 nx_desc = make_int4(-1, -1, 0, 0);
    int4 cu_desc = make_int4(-1, -1, 0, 0);   // TMA: the batch whose copies are in flight
    uint64_t cu_raw = 0;                      // TMA: its mask word
    int cu_js = 0;                            //      and joint state
    auto node_of = [&](int b) {
        const int nd = b + cl.sub;
        return node_begin + ((cl.lane_on && nd < n_nodes) ? nd : b);
    };
    // the copies of one batch, issued by one thread per node: own vector (internal nodes), parent's bottom-up vector,
    // parent's top-down vector (parents that are not roots), each with its exponent row
    auto issue = [&](const int4 &d) {
        if (cl.r != 0 || cl.j != 0) return;
        const int ps = d.x, cs = d.y;
        const bool ptd = ps >= n_root_slots;
        mbar_arrive_expect(bar, (cs >= 0 ? VB + EB : 0) + (ps >= 0 ? VB + EB : 0) + (ptd ? VB + EB : 0));
        if (cs >= 0) {
            bulk_g2s(stage, a.bu + (size_t) cs * K * C, VB, bar);
            bulk_g2s(stage + 3 * VB, a.bu_ex + (size_t) cs * C, EB, bar);
        }
        if (ps >= 0) {
            bulk_g2s(stage + VB, a.bu + (size_t) ps * K * C, VB, bar);
            bulk_g2s(stage + 3 * VB + EB, a.bu_ex + (size_t) ps * C, EB, bar);
        }
        if (ptd) {
            bulk_g2s(stage + 2 * VB, a.td + (size_t) ps * K * C, VB, bar);
            bulk_g2s(stage + 3 * VB + 2 * EB, a.td_ex + (size_t) ps * C, EB, bar);
        }
    };
    if constexpr (TMA) pdl_launch_dependents();
    if (base < n_nodes) {
        if constexpr (TMA) {
            const int n0 = node_of(base);
            cu_desc = a.t.td_desc[n0];
            cu_raw = a.mask[(size_t) n0 * C + c];
            if (fj) cu_js = a.joint_state[(size_t) n0 * C + c];
            if (base + stride < n_nodes) nx_desc = a.t.td_desc[node_of(base + stride)];
            pdl_wait();   // everything above (topology, masks, joint states) is independent of the previous level
            issue(cu_desc);
        } else {
            nx_desc = a.t.td_desc[node_of(base)];
        }
    }
    if constexpr (!TMA) {
        pdl_launch_dependents();
        pdl_wait();
    }
    for (; base < n_nodes; base += stride) {
        const int nd = base + cl.sub;
        const bool on = cl.lane_on && nd < n_nodes;
        const int node = node_begin + (on ? nd : base);
        const int4 desc = TMA ? cu_desc : nx_desc;   // {parent slot, own slot, dist}
        const uint64_t raw_mask = cu_raw;
        const int js_pre = cu_js;
        const int ps = desc.x;
        const int cs = desc.y;
        const size_t nc = (size_t) node * C + c;
        const bool has_parent = ps >= 0;

        double Lc[KT], X[KT];
        int exc = 0, pex = 0, ex_total;
        uint64_t base_mask, allowed;
        if constexpr (TMA) {
            mbar_wait(bar, phase);
            phase ^= 1u;
            base_mask = raw_mask;
            allowed = eff_from_raw(raw_mask, a.t, a.zc, C, node, c, use_alt, a.ln.kmask);
            const int *st_ex = reinterpret_cast<const int *>(st_vec + 3 * (VB / 8));
            if (cs >= 0) {
#pragma unroll
                for (int i = 0; i < KT; ++i) Lc[i] = st_vec[(r + G * i) * C + c];
                exc = st_ex[c];
            } else {   // a tip: its vector is its allowed-state mask
                const uint64_t m = allowed >> r;
#pragma unroll
                for (int i = 0; i < KT; ++i) Lc[i] = bit_as_double(m, G * i);
            }
            ex_total = exc;
            // parent part of the node's top-down vector: TD_p * L_p (all ones above a root, ml.py:274-277)
            if (has_parent) {
#pragma unroll
                for (int i = 0; i < KT; ++i) X[i] = st_vec[VB / 8 + (r + G * i) * C + c];
                pex = st_ex[C + c];
                if (ps >= n_root_slots) {
#pragma unroll
                    for (int i = 0; i < KT; ++i) X[i] *= st_vec[2 * (VB / 8) + (r + G * i) * C + c];
                    pex += st_ex[2 * C + c];
                } else if (a.ext.root_td) {   // child of a clade root: the root's top-down vector came from above the cut
                    const int rn = a.t.int_nodes[ps];
                    const double *src2 = a.ext.root_td + ((size_t) (rn * K + r) * C + c);
#pragma unroll
                    for (int i = 0; i < KT; ++i) X[i] *= src2[(G * i) * C];
                    pex += a.ext.root_td_ex[(size_t) rn * C + c];
                }
            } else if (a.ext.root_td) {   // a clade root (subtree sharding): its top-down vector arrives from above the cut
                const double *src = a.ext.root_td + ((size_t) (node * K + r) * C + c);
#pragma unroll
                for (int i = 0; i < KT; ++i) X[i] = src[(G * i) * C];
                ex_total += a.ext.root_td_ex[(size_t) node * C + c];
            } else {
#pragma unroll
                for (int i = 0; i < KT; ++i) X[i] = 1.;
            }
            __syncwarp();
            if (base + stride < n_nodes) {
                issue(nx_desc);
                cu_desc = nx_desc;
                const int n1 = node_of(base + stride);
                cu_raw = a.mask[(size_t) n1 * C + c];
                if (fj) cu_js = a.joint_state[(size_t) n1 * C + c];
                if (base + 2 * stride < n_nodes) nx_desc = a.t.td_desc[node_of(base + 2 * stride)];
            }
        } else {
            if (base + stride < n_nodes) nx_desc = a.t.td_desc[node_of(base + stride)];
            exc = coop_load<KT, G, KX>(a.t, a.zc, a.ext, a.ln, a.mask, a.bu, a.bu_ex, node, cs, r, c, c, use_alt, Lc, K, C);
            base_mask = a.mask[nc];
            allowed = eff_mask(a.mask, a.t, a.zc, C, node, c, use_alt, a.ln.kmask);
            ex_total = exc;
            // parent part of the node's top-down vector: TD_p * L_p (all ones above a root, ml.py:274-277)
            if (has_parent) {
                const size_t off = (size_t) (ps * K + r) * C + c;
                const double *src = a.bu + off;
#pragma unroll
                for (int i = 0; i < KT; ++i) X[i] = PML_VALID(i) ? src[(G * i) * C] : 0.;
                pex = a.bu_ex[(size_t) ps * C + c];
                if (ps >= n_root_slots) {
                    const double *src2 = a.td + off;
#pragma unroll
                    for (int i = 0; i < KT; ++i) X[i] *= PML_VALID(i) ? src2[(G * i) * C] : 0.;
                    pex += a.td_ex[(size_t) ps * C + c];
                } else if (a.ext.root_td) {   // child of a clade root: the root's top-down vector came from above the cut
                    const int rn = a.t.int_nodes[ps];
                    const double *src2 = a.ext.root_td + ((size_t) (rn * K + r) * C + c);
#pragma unroll
                    for (int i = 0; i < KT; ++i) X[i] *= PML_VALID(i) ? src2[(G * i) * C] : 0.;
                    pex += a.ext.root_td_ex[(size_t) rn * C + c];
                }
            } else if (a.ext.root_td) {   // a clade root: its top-down vector arrives from above the cut
                const double *src = a.ext.root_td + ((size_t) (node * K + r) * C + c);
#pragma unroll
                for (int i = 0; i < KT; ++i) X[i] = PML_VALID(i) ? src[(G * i) * C] : 0.;
                ex_total += a.ext.root_td_ex[(size_t) node * C + c];
            } else {
#pragma unroll
                for (int i = 0; i < KT; ++i) X[i] = 1.;
            }
        }
        const uint64_t flag = base_mask & kMemberFlag;
        // the branch above the node: rank 0 evaluates the exponential for its group
        double e = 0.;
        if (r == 0 && !mu_inf) e = exp_nonpos(-mu * ((__hiloint2double(desc.w, desc.z) + tau) * tauf * sf));
        e = grp.from_rank(e, 0);
        const double ome = 1. - e;
        // node_contribution = P L_node (ml.py:280); q = TD_p L_p / contribution, contribution <= 0 -> 1 (ml.py:281-283)
        double s = 0.;
#pragma unroll
        for (int i = 0; i < KT; ++i) s = fma(pi[i], Lc[i], s);
        s = grp.sum(s);
        double q[KT];
        {
            // the reciprocals: the range check of fast_rcp is hoisted out of the loop (one vote per warp instead of a
            // branch per element); tiny / huge contributions take the exact path for the whole batch
            // contribution_k = e L_k + (1 - e) s >= (1 - e) s: when that floor is a normal number away from the ends of
            // the exponent range (practically always: only zero-length branches and vanishing vectors fail it), every
            // contribution is, and the reciprocals take the short path without a range check per element
            const double os = ome * s;
            if (__all_sync(kFullMask, os > 0x1p-900 && s < 0x1p+800)) {
#pragma unroll
                for (int i = 0; i < KT; ++i) {
                    const double cv = fma(e, Lc[i], os);
                    double rr;
                    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rr) : "d"(cv));
                    double tt = fma(-cv, rr, 1.);
                    rr = fma(rr, tt, rr);
                    tt = fma(-cv, rr, 1.);
                    rr = fma(rr, tt, rr);
                    q[i] = X[i] * rr;
                }
            } else {
#pragma unroll
                for (int i = 0; i < KT; ++i) {
                    const double cv = fma(e, Lc[i], os);   // >= 0: e in [0, 1] and the inputs are non-negative
                    q[i] = (cv > 0.) ? X[i] * fast_rcp(cv) : X[i];
                }
            }
        }
        double s2 = 0.;
#pragma unroll
        for (int i = 0; i < KT; ++i) s2 = fma(pi[i], q[i], s2);
        s2 = grp.sum(s2);
        // keeping the vector in range (ml.py:284-286): the frequency-weighted sum is already reduced over the group
        // and is within a factor 1/min(pi) of the maximum, so its exponent decides the power-of-two rescale
        int qex = pex - exc;
        double scale = 1.;
        if (s2 > 0. && !isinf(s2)) {
            const int e2 = exponent_of(s2);
            if (e2 < -kKeep || e2 > kKeep) {
                const int ee = (e2 == -1023) ? -1022 : e2;
                scale = pow2(-ee / 2) * pow2(-(ee - ee / 2));
                if (!(scale > 0.) || isinf(scale)) { scale = (ee > 0) ? 0x1p-1000 : 0x1p+1000; qex += (ee > 0) ? 1000 : -1000; }
                else qex += ee;
            }
        }
        double tdv[KT];
#pragma unroll
        for (int i = 0; i < KT; ++i)
            tdv[i] = PML_VALID(i) ? fma(e, q[i], ome * s2) * scale : 0.;   // ml.py:288-289 (already >= 0)
        if (has_parent) {
            if (cs >= 0) {
                if (on) {
                    double *dst = a.td + ((size_t) (cs * K + r) * C + c);
#pragma unroll
                    for (int i = 0; i < KT; ++i)
                        if (PML_VALID(i)) dst[(G * i) * C] = tdv[i];
                    if (r == 0) a.td_ex[(size_t) cs * C + c] = qex;
                }
            }
            else if (!TMA && (base_mask & kExtFlag)) {   // hand the vector to the clade below the cut
                if (on) {
                    const int e_row = a.ext.index[node];
                    double *dst = a.ext.td + ((size_t) (e_row * K + r) * C + c);
#pragma unroll
                    for (int i = 0; i < KT; ++i)
                        if (PML_VALID(i)) dst[(G * i) * C] = tdv[i];
                    if (r == 0) a.ext.td_ex[(size_t) e_row * C + c] = qex;
                }
            }
            ex_total = qex + exc;
        }
        // marginal likelihood vector (ml.py:454-465): L * TD * pi * allowed
        double m[KT];
        double sum = 0.;
        const uint64_t allowed_r = allowed >> r;
#pragma unroll
        for (int i = 0; i < KT; ++i) {
            const double tdi = has_parent ? tdv[i] : X[i];   // roots: all ones, or the vector from above the cut
            m[i] = ((allowed_r >> (G * i)) & 1ULL) ? Lc[i] * tdi * pi[i] : 0.;
            sum += m[i];
        }
        sum = grp.sum(sum);
        double inv = __drcp_rn(sum);
        if (on) {
            if (a.probs) {
                double *dst = a.probs + nc * K + r;
#pragma unroll
                for (int i = 0; i < KT; ++i)
                    if (PML_VALID(i)) dst[G * i] = m[i] * inv;
            }
            // the logarithm is taken by log10_finish_kernel over the whole array afterwards: here it would be
            // evaluated by one rank per group with the rest of the warp idle
            if (a.node_log10lh && r == 0) {
                a.node_log10lh[nc] = sum;
                a.node_ex[nc] = ex_total;
            }
        }
        // members of zero-branch clusters: remember the vector, apply '.initial' (ml.py:593-594)
        uint64_t init = a.ln.kmask;
        size_t zr = 0;
        if (flag) {
            zr = (size_t) a.t.zc_index[node] * C + c;
            if (on) {
                double *z = a.zc_lh + zr * K + r;
#pragma unroll
                for (int i = 0; i < KT; ++i)
                    if (PML_VALID(i)) z[G * i] = m[i];
            }
            if (a.zc.hasinit[zr]) init = a.zc.initial[zr];
        }
        if (__any_sync(kFullMask, flag != 0ULL)) {   // rare: some node of the warp's batch sits in a zero-branch cluster
            double sum2 = 0.;
            const uint64_t init_r = init >> r;
#pragma unroll
            for (int i = 0; i < KT; ++i) {
                m[i] = ((init_r >> (G * i)) & 1ULL) ? m[i] : 0.;
                sum2 += m[i];
            }
            sum2 = grp.sum(sum2);
            if (flag) { sum = sum2; inv = 1. / sum2; }
        }
        // MAP: first arg max (ml.py:595)
        double bv = -1.;
        int best = 0;
#pragma unroll
        for (int i = 0; i < KT; ++i)
            if (PML_VALID(i) && m[i] > bv) { bv = m[i]; best = r + G * i; }
        grp.argmax(bv, best);
        const int ji = fj ? (TMA ? js_pre : a.joint_state[nc]) : best;
        // MPPA (ml.py:539-572), see mppa_select in pml_kernels.cuh for the derivation
        int best_k = 1;
        uint64_t sel = 1ULL << best;
        bool done = true;
        if (!(sum > 0.)) { sel = a.ln.kmask; best_k = K; }
        else if (!(bv * inv >= 0.76 && ji == best)) done = false;
        if (!__all_sync(kFullMask, done)) {
            double pj = 0.;
#pragma unroll
            for (int i = 0; i < KT; ++i) pj += (r + G * i == ji) ? m[i] * inv : 0.;
            pj = grp.sum(pj);
            unsigned used = 0u, best_used = 0u;
            double run = 0., best_score = CUDART_INF;
            bool joint_seen = false;
            int bk = K;
            for (int k = 1; k <= K; ++k) {
                if (__all_sync(kFullMask, done)) break;
                double lv = -1.;
                int lk = 1 << 20;
#pragma unroll
                for (int i = 0; i < KT; ++i)
                    if (PML_VALID(i) && !((used >> i) & 1u) && m[i] > lv) { lv = m[i]; lk = r + G * i; }
                grp.argmax(lv, lk);
                if (!done) {
                    if (lk < K && lk >= r && (lk - r) % G == 0) used |= 1u << ((lk - r) / G);
                    const double prev = run;
                    run += lv * inv;
                    joint_seen = joint_seen || (lk == ji);
                    const double tk = (!fj || joint_seen) ? run : pj + prev;
                    const double score = (1. - 2. * tk) / (double) k;
                    if (score < best_score) { best_score = score; bk = k; best_used = used; }
                    else if (score > best_score + 1e-9) done = true;
                    if (k == K) done = true;
                }
            }
            unsigned long long mine = 0ULL;
#pragma unroll
            for (int i = 0; i < KT; ++i)
                if ((best_used >> i) & 1u) mine |= 1ULL << (r + G * i);
            mine = grp.bit_or(mine);
            if (sum > 0. && !(bv * inv >= 0.76 && ji == best)) { sel = mine; best_k = bk; }
        }
        if (on && r == 0) {
            // (the marker of an external tip travels with the state sets: restricted passes on these masks must still
            // take the tip's vector from the tables of the clade below the cut)
            a.map_mask[nc] = (1ULL << best) | flag | (base_mask & kExtFlag);
            a.mppa_mask[nc] = sel | flag | (base_mask & kExtFlag);
            if (flag) a.zc_k[zr] = (uint8_t) best_k;
            atomicAdd(sh_hist + c * (K + 1) + best_k, 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * (K + 1); i += blockDim.x)
        if (sh_hist[i]) atomicAdd(a.k_hist + i, (unsigned long long) sh_hist[i]);
}

}  // namespace pml
