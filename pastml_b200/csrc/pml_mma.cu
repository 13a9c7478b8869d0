// Bottom-up sweep of the eigen-decomposition models (JTT, CUSTOM_RATES) on the FP64 tensor cores.
//
// P(t) L = A (exp(d t') o (A^-1 L)) (generator.py:54-65) costs 2 K^2 fused multiply-adds per (branch, lane); when
// all the lanes share one decomposition (JTT with its fixed frequencies: the columns differ only in the scaling
// factor, which enters through exp(d t')), the two products of a node are small GEMMs against the SAME matrices:
// [K x K] . [K x 8 lanes].  One warp takes one (node, tile of 8 lanes): the child's vector, [k][lane] in memory, is
// exactly the col-major B operand of mma.sync.m8n8k4 (one coalesced 256-byte load per step of 4 states), the
// matrices sit in registers as A fragments (2 x ceil(K/8) x ceil(K/4) doubles per thread), the accumulators are
// the 8 x 8 D tiles.  Between the two products the D tiles are scaled by exp(d_m t'_n) and turned into B fragments
// through a 1.5 KB shared-memory slot per warp.
//
// Used for K <= 24, lanes a multiple of 8, sum-product mode, no external (subtree-sharding) tips; everything else
// takes the thread-per-item kernels of pml_kernels.cuh.
#include <algorithm>
#include "pml_coop.cuh"
#include "pml_launch.h"

namespace pml {
namespace {

constexpr int kMmaKP = 24;            // largest padded state count: 3 row tiles of 8, 6 steps of 4
constexpr int kMmaBlock = 128;
#ifndef PML_MMA_MINB
#define PML_MMA_MINB 8   // latency-bound: resident warps beat registers; with the A fragments in shared memory 64 registers
                         // hold the rest (JTT step, 1M tips x 8 columns: 4 blocks 10.6 ms, 6 blocks 10.0 ms; fragments in
                         // shared memory: 6 blocks 8.9 ms, 8 blocks 8.74 ms, 10 blocks 9.3 ms)
#endif
constexpr int kMmaMinBlocks = PML_MMA_MINB;
#ifndef PML_MMA_SMEM_FRAG
#define PML_MMA_SMEM_FRAG 1
#endif

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// exp_nonpos (pml_coop.cuh) on N arguments at once: the Horner steps run across the arguments, so that each of the 14
// coefficients is materialised once per batch instead of once per exponential
template <int N>
__device__ __forceinline__ void exp_nonpos_n(const double (&x)[N], double (&y)[N]) {
    double r[N], p[N];
    int n[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double t = fma(x[i], 1.4426950408889634074, 6755399441055744.0);
        n[i] = __double2loint(t);
        const double nf = t - 6755399441055744.0;
        r[i] = fma(nf, -6.93147180369123816490e-01, x[i]);
        r[i] = fma(nf, -1.90821492927058770002e-10, r[i]);
        p[i] = 1.6059043836821613e-10;
    }
    const double c[13] = {2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07, 2.7557319223985893e-06,
                          2.48015873015873e-05, 1.984126984126984e-04, 1.388888888888889e-03, 8.333333333333333e-03,
                          4.1666666666666664e-02, 1.6666666666666666e-01, 0.5, 1., 1.};
#pragma unroll
    for (int k = 0; k < 13; ++k) {
#pragma unroll
        for (int i = 0; i < N; ++i) p[i] = fma(p[i], r[i], c[k]);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double v = __hiloint2double(__double2hiint(p[i]) + n[i] * 1048576, __double2loint(p[i]));
        y[i] = (x[i] > -708.) ? v : ((x[i] != x[i]) ? x[i] : 0.);
    }
}

// group-wide maximum over the 8 row groups of a D tile column (lanes that differ in bits 2..4)
__device__ __forceinline__ double colmax(double v) {
    v = fmax(v, __shfl_xor_sync(kFullMask, v, 4));
    v = fmax(v, __shfl_xor_sync(kFullMask, v, 8));
    return fmax(v, __shfl_xor_sync(kFullMask, v, 16));
}

// MT row tiles of 8 states, KS steps of 4 states: the smallest pair that covers K (all-zero tiles are not issued)
template <int MT, int KS>
__global__ void __launch_bounds__(kMmaBlock, kMmaMinBlocks) bu_mma_kernel(BuArgs a, int slot_begin, int n_slots) {
    __shared__ __align__(16) double s_w[kMmaBlock / 32][kMmaKP * 8];
    const int K = a.ln.K, J = a.ln.J, C = a.ln.C;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t4 = lane & 3;   // D tile: row g, columns 2 t4, 2 t4 + 1; A: (g, t4); B: (t4, g)
    const int tiles = J >> 3;
    double *sw = s_w[warp];

    // the shared decomposition, from lane 0's parameter block: A fragments of A^-1 and of A, this thread's eigenvalues
    const double *dv = a.ln.params + 4 + K, *Am = dv + K, *Ai = Am + (size_t) K * K;
    double dm[MT];
#if PML_MMA_SMEM_FRAG
    // the A fragments of both matrices are the same for every warp of the block: kept in shared memory ([fragment][lane],
    // conflict-free), which frees 4 MT KS registers per thread -- at 6 blocks per SM they would be spilled
    __shared__ double s_fi[MT * KS][32], s_fa[MT * KS][32];
#define F_INV(mt, s) s_fi[(mt) * KS + (s)][lane]
#define F_A(mt, s) s_fa[(mt) * KS + (s)][lane]
#else
    double f_inv[MT][KS], f_a[MT][KS];
#define F_INV(mt, s) f_inv[mt][s]
#define F_A(mt, s) f_a[mt][s]
#endif
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int row = 8 * mt + g;
        dm[mt] = (row < K) ? dv[row] : 0.;
#if PML_MMA_SMEM_FRAG
        if (warp == 0)
#endif
#pragma unroll
        for (int s = 0; s < KS; ++s) {
            const int col = 4 * s + t4;
            const bool in = row < K && col < K;
            F_INV(mt, s) = in ? Ai[row * K + col] : 0.;
            F_A(mt, s) = in ? Am[row * K + col] : 0.;
        }
    }
#if PML_MMA_SMEM_FRAG
    __syncthreads();
#endif
    pdl_launch_dependents();   // the fragments above are loaded while the previous level drains
    // the lane tile comes from blockIdx.y, so that everything that depends on it only is loop-invariant
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const int n_items = n_slots;
    const int tile = blockIdx.y;
    const int j0 = tile * 8;
    const int jb = j0 + g;              // the lane this thread feeds as B operand
    const int jd = j0 + 2 * t4;         // the two lanes of this thread's D elements: jd, jd + 1
    const int cb = jb % C, cd0 = jd % C, cd1 = (jd + 1) % C;
    const double *p0 = a.ln.params + (size_t) jd * a.ln.stride, *p1 = p0 + a.ln.stride;
    const double sf0 = p0[0], tau0 = p0[1], tf0 = p0[2], sf1 = p1[0], tau1 = p1[1], tf1 = p1[2];
    const bool alt_b = a.alter && a.ln.params[(size_t) jb * a.ln.stride + 1] == 0.;
    (void) tiles;
    // the packed descriptor (and node id) of the next item is requested before the current one is processed
    int4 nx_top = make_int4(0, 0, 0, 0), nx_len = nx_top;
    int nx_node = 0;
    if (warp_global < n_items) {
        const size_t s0 = (size_t) slot_begin + (size_t) warp_global;
        nx_top = a.t.bu_desc[2 * s0];
        nx_len = a.t.bu_desc[2 * s0 + 1];
        nx_node = a.t.int_nodes[s0];
    }
    pdl_wait();   // from here on the previous level's vectors are read
    for (int w = warp_global; w < n_items; w += n_warps) {
        const int slot = slot_begin + w;
        const int4 top = nx_top, len = nx_len;
        const int node = nx_node;
        if (w + n_warps < n_items) {
            const size_t s1 = (size_t) slot_begin + (size_t) (w + n_warps);
            nx_top = a.t.bu_desc[2 * s1];
            nx_len = a.t.bu_desc[2 * s1 + 1];
            nx_node = a.t.int_nodes[s1];
        }
        const int fc = top.x, nc = top.y;
        double acc[MT][2];
        {
            const uint64_t m0 = eff_mask(a.mask, a.t, a.zc, C, node, cd0, a.alter && tau0 == 0., a.ln.kmask);
            const uint64_t m1 = eff_mask(a.mask, a.t, a.zc, C, node, cd1, a.alter && tau1 == 0., a.ln.kmask);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {   // 10 ** log10(allowed), ml.py:127
                acc[mt][0] = (8 * mt + g < K) ? bit_as_double(m0, 8 * mt + g) : 0.;
                acc[mt][1] = (8 * mt + g < K) ? bit_as_double(m1, 8 * mt + g) : 0.;
            }
        }
        int ex0 = 0, ex1 = 0, bad0 = -1, bad1 = -1;
        for (int q = 0; q < nc; ++q) {
            const int ch = fc + q;
            const int cs = (q == 0) ? top.z : (q == 1) ? top.w : a.t.slot[ch];
            const double dist = (q == 0) ? __hiloint2double(len.y, len.x) : (q == 1) ? __hiloint2double(len.w, len.z) : a.t.dist[ch];
            // the child's vector as B fragments: states 4 s + t4 of lane jb
            double lf[KS];
            if (cs >= 0) {
                const double *src = a.bu + (size_t) cs * K * J + jb;
#pragma unroll
                for (int s = 0; s < KS; ++s) lf[s] = (4 * s + t4 < K) ? src[(size_t) (4 * s + t4) * J] : 0.;
            } else {
                const uint64_t m = eff_mask(a.mask, a.t, a.zc, C, ch, cb, alt_b, a.ln.kmask);
#pragma unroll
                for (int s = 0; s < KS; ++s) lf[s] = bit_as_double(m, 4 * s + t4);
            }
            // W = A^-1 L
            double wv[MT][2];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                wv[mt][0] = wv[mt][1] = 0.;
#pragma unroll
                for (int s = 0; s < KS; ++s) dmma(wv[mt][0], wv[mt][1], F_INV(mt, s), lf[s]);
            }
            // W o exp(d t'), handed over as B fragments through the warp's shared-memory slot
            const double t0 = (dist + tau0) * tf0 * sf0, t1 = (dist + tau1) * tf1 * sf1;
            {
                double xs[2 * MT], es[2 * MT];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {   // rows >= K carry d = 0 and W = 0
                    xs[2 * mt] = dm[mt] * t0;
                    xs[2 * mt + 1] = dm[mt] * t1;
                }
                exp_nonpos_n<2 * MT>(xs, es);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
                    *reinterpret_cast<double2 *>(sw + (8 * mt + g) * 8 + 2 * t4) = make_double2(wv[mt][0] * es[2 * mt], wv[mt][1] * es[2 * mt + 1]);
            }
            __syncwarp();
            double bf[KS];
#pragma unroll
            for (int s = 0; s < KS; ++s) bf[s] = sw[(4 * s + t4) * 8 + g];
            __syncwarp();
            // V = A (W o exp(d t')); acc *= max(V, 0)
            double m0 = 0., m1 = 0.;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                double v0 = 0., v1 = 0.;
#pragma unroll
                for (int s = 0; s < KS; ++s) dmma(v0, v1, F_A(mt, s), bf[s]);
                if (a.v_all && 8 * mt + g < K)   // kept for the top-down sweep: its node_contribution (ml.py:280)
                    *reinterpret_cast<double2 *>(a.v_all + ((size_t) ch * K + 8 * mt + g) * J + jd) = make_double2(v0, v1);
                acc[mt][0] *= fmax(v0, 0.);
                acc[mt][1] *= fmax(v1, 0.);
                m0 = fmax(m0, acc[mt][0]);
                m1 = fmax(m1, acc[mt][1]);
            }
            m0 = colmax(m0);
            m1 = colmax(m1);
            if (cs >= 0) {
                ex0 += a.bu_ex[(size_t) cs * J + jd];
                ex1 += a.bu_ex[(size_t) cs * J + jd + 1];
            }
            if (m0 == 0. && bad0 < 0) bad0 = ch;
            if (m1 == 0. && bad1 < 0) bad1 = ch;
            // keep the vectors in range (same policy as rescale<> of the thread-per-item kernels), lane by lane
            if (m0 > 0. && !isinf(m0)) {
                int e = exponent_of(m0);
                if (e < -kKeep || e > kKeep) {
                    int extra = 0;
                    double lift = 1.;
                    if (e == -1023) { lift = 0x1p+200; e = exponent_of(m0 * 0x1p+200); extra = -200; }
                    const double s1 = pow2(-(e / 2)), s2 = pow2(-(e - e / 2));
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) acc[mt][0] = ((acc[mt][0] * lift) * s1) * s2;
                    ex0 += e + extra;
                }
            }
            if (m1 > 0. && !isinf(m1)) {
                int e = exponent_of(m1);
                if (e < -kKeep || e > kKeep) {
                    int extra = 0;
                    double lift = 1.;
                    if (e == -1023) { lift = 0x1p+200; e = exponent_of(m1 * 0x1p+200); extra = -200; }
                    const double s1 = pow2(-(e / 2)), s2 = pow2(-(e - e / 2));
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) acc[mt][1] = ((acc[mt][1] * lift) * s1) * s2;
                    ex1 += e + extra;
                }
            }
        }
        // the D tiles are the node's vector: rows = states, two adjacent lanes per thread
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const int row = 8 * mt + g;
            if (row < K) *reinterpret_cast<double2 *>(a.bu + ((size_t) slot * K + row) * J + jd) = make_double2(acc[mt][0], acc[mt][1]);
        }
        if (g == 0) {
            a.bu_ex[(size_t) slot * J + jd] = ex0;
            a.bu_ex[(size_t) slot * J + jd + 1] = ex1;
            // all entries zero: PastMLLikelihoodError of ml.py:139-145
            if (bad0 >= 0 && atomicCAS(a.err + 2 * jd, -1, node) == -1) a.err[2 * jd + 1] = bad0;
            if (bad1 >= 0 && atomicCAS(a.err + 2 * (jd + 1), -1, node) == -1) a.err[2 * (jd + 1) + 1] = bad1;
        }
    }
}

// Top-down vectors of one level (calc_node_td_likelihood, ml.py:273-290) on the tensor cores, one warp per (node, tile
// of 8 lanes).  The node's contribution P L_node was formed -- and kept, BuArgs::v_all -- by the bottom-up sweep, so the
// step is ONE application of P: q = TD_p L_p / contribution in the D-tile layout (rows = states, two adjacent lanes
// per thread, read straight from memory as double2), turned into B fragments through the warp's shared-memory slot,
// W = A^-1 q, W o exp(d t'), TD = A (.).  Vectors only: the marginal / MAP / MPPA half of the step has no dependency
// between nodes and runs once for the whole forest afterwards (td_select_kernel).  Top-down vectors are stored for
// every node that has a parent, tips included, indexed by node.
template <int MT, int KS>
__global__ void __launch_bounds__(kMmaBlock, kMmaMinBlocks) td_mma_kernel(TdArgs a, const double *v_all, double *td_all,
                                                                        int *td_ex_all, int node_begin, int n_nodes) {
    __shared__ __align__(16) double s_w[kMmaBlock / 32][kMmaKP * 8];
    const int K = a.ln.K, C = a.ln.C;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int tiles = C >> 3;
    double *sw = s_w[warp];
    const double *dv = a.ln.params + 4 + K, *Am = dv + K, *Ai = Am + (size_t) K * K;
    double dm[MT];
#if PML_MMA_SMEM_FRAG
    // the A fragments of both matrices are the same for every warp of the block: kept in shared memory ([fragment][lane],
    // conflict-free), which frees 4 MT KS registers per thread -- at 6 blocks per SM they would be spilled
    __shared__ double s_fi[MT * KS][32], s_fa[MT * KS][32];
#define F_INV(mt, s) s_fi[(mt) * KS + (s)][lane]
#define F_A(mt, s) s_fa[(mt) * KS + (s)][lane]
#else
    double f_inv[MT][KS], f_a[MT][KS];
#define F_INV(mt, s) f_inv[mt][s]
#define F_A(mt, s) f_a[mt][s]
#endif
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int row = 8 * mt + g;
        dm[mt] = (row < K) ? dv[row] : 0.;
#if PML_MMA_SMEM_FRAG
        if (warp == 0)
#endif
#pragma unroll
        for (int s = 0; s < KS; ++s) {
            const int col = 4 * s + t4;
            const bool in = row < K && col < K;
            F_INV(mt, s) = in ? Ai[row * K + col] : 0.;
            F_A(mt, s) = in ? Am[row * K + col] : 0.;
        }
    }
#if PML_MMA_SMEM_FRAG
    __syncthreads();
#endif
    pdl_launch_dependents();
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const int n_items = n_nodes;
    const int n_root_slots = a.t.n_root_slots;
    // the lane tile comes from blockIdx.y: this thread's two lanes and their parameters are loop-invariant
    const int tile = blockIdx.y;
    const int jd = tile * 8 + 2 * t4;   // the two lanes of this thread's D elements
    const double *p0 = a.ln.params + (size_t) jd * a.ln.stride, *p1 = p0 + a.ln.stride;
    const double sf0 = p0[0], tau0 = p0[1], tf0 = p0[2], sf1 = p1[0], tau1 = p1[1], tf1 = p1[2];
    (void) tiles;
    int4 nx_desc = make_int4(-1, -1, 0, 0);
    int nx_par = -1;
    if (warp_global < n_items) {
        const int n0 = node_begin + warp_global;
        nx_desc = a.t.td_desc[n0];
        nx_par = a.t.parent[n0];
    }
    pdl_wait();   // from here on the previous level's vectors are read
    for (int w = warp_global; w < n_items; w += n_warps) {
        const int node = node_begin + w;
        const int4 desc = nx_desc;
        const int par = nx_par;
        if (w + n_warps < n_items) {
            const int n1 = node_begin + w + n_warps;
            nx_desc = a.t.td_desc[n1];
            nx_par = a.t.parent[n1];
        }
        const int ps = desc.x, cs = desc.y;
        if (ps < 0) continue;   // a root: all ones, never stored (ml.py:274-277)
        const double dist = __hiloint2double(desc.w, desc.z);
        const double t0 = (dist + tau0) * tf0 * sf0, t1 = (dist + tau1) * tf1 * sf1;
        // q = TD_p L_p / contribution, contribution <= 0 -> 1 (ml.py:281-283), in the D layout
        double q[MT][2];
        int pex0 = a.bu_ex[(size_t) ps * C + jd], pex1 = a.bu_ex[(size_t) ps * C + jd + 1];
        const bool ptd = ps >= n_root_slots;
        if (ptd) {
            pex0 += td_ex_all[(size_t) par * C + jd];
            pex1 += td_ex_all[(size_t) par * C + jd + 1];
        }
        double m0 = 0., m1 = 0.;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const int row = 8 * mt + g;
            q[mt][0] = q[mt][1] = 0.;
            if (row < K) {
                double2 x = *reinterpret_cast<const double2 *>(a.bu + ((size_t) ps * K + row) * C + jd);
                if (ptd) {
                    const double2 t = *reinterpret_cast<const double2 *>(td_all + ((size_t) par * K + row) * C + jd);
                    x.x *= t.x;
                    x.y *= t.y;
                }
                const double2 cv = *reinterpret_cast<const double2 *>(v_all + ((size_t) node * K + row) * C + jd);
                q[mt][0] = (cv.x > 0.) ? x.x * __drcp_rn(cv.x) : x.x;
                q[mt][1] = (cv.y > 0.) ? x.y * __drcp_rn(cv.y) : x.y;
                m0 = fmax(m0, q[mt][0]);
                m1 = fmax(m1, q[mt][1]);
            }
        }
        int ex0 = pex0, ex1 = pex1;
        if (cs >= 0) {
            ex0 -= a.bu_ex[(size_t) cs * C + jd];
            ex1 -= a.bu_ex[(size_t) cs * C + jd + 1];
        }
        // keep q in range (ml.py:284-286), lane by lane, before it goes through the products
        m0 = colmax(m0);
        m1 = colmax(m1);
        auto rescale_lane = [&](double mx, int which, int &ex, bool always, double (&v)[MT][2]) {
            if (!(mx > 0.) || isinf(mx)) return;
            int e = exponent_of(mx);
            if (!always && e >= -kKeep && e <= kKeep) return;
            int extra = 0;
            double lift = 1.;
            if (e == -1023) { lift = 0x1p+200; e = exponent_of(mx * 0x1p+200); extra = -200; }
            const double s1 = pow2(-(e / 2)), s2 = pow2(-(e - e / 2));
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) v[mt][which] = ((v[mt][which] * lift) * s1) * s2;
            ex += e + extra;
        };
        rescale_lane(m0, 0, ex0, false, q);
        rescale_lane(m1, 1, ex1, false, q);
        // q as B fragments: through the warp's slot (D layout in, [state][lane] out)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
            *reinterpret_cast<double2 *>(sw + (8 * mt + g) * 8 + 2 * t4) = make_double2(q[mt][0], q[mt][1]);
        __syncwarp();
        double bf[KS];
#pragma unroll
        for (int s = 0; s < KS; ++s) bf[s] = sw[(4 * s + t4) * 8 + g];
        __syncwarp();
        // W = A^-1 q;  W o exp(d t')
        double wv[MT][2];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            wv[mt][0] = wv[mt][1] = 0.;
#pragma unroll
            for (int s = 0; s < KS; ++s) dmma(wv[mt][0], wv[mt][1], F_INV(mt, s), bf[s]);
        }
        {
            double xs[2 * MT], es[2 * MT];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                xs[2 * mt] = dm[mt] * t0;
                xs[2 * mt + 1] = dm[mt] * t1;
            }
            exp_nonpos_n<2 * MT>(xs, es);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
                *reinterpret_cast<double2 *>(sw + (8 * mt + g) * 8 + 2 * t4) = make_double2(wv[mt][0] * es[2 * mt], wv[mt][1] * es[2 * mt + 1]);
        }
        __syncwarp();
#pragma unroll
        for (int s = 0; s < KS; ++s) bf[s] = sw[(4 * s + t4) * 8 + g];
        __syncwarp();
        // TD = max(A (W o exp(d t')), 0) (ml.py:288-289), normalised so that the stored maximum sits in [1, 2)
        double td[MT][2];
        m0 = m1 = 0.;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            double v0 = 0., v1 = 0.;
#pragma unroll
            for (int s = 0; s < KS; ++s) dmma(v0, v1, F_A(mt, s), bf[s]);
            td[mt][0] = fmax(v0, 0.);
            td[mt][1] = fmax(v1, 0.);
            m0 = fmax(m0, td[mt][0]);
            m1 = fmax(m1, td[mt][1]);
        }
        m0 = colmax(m0);
        m1 = colmax(m1);
        rescale_lane(m0, 0, ex0, true, td);
        rescale_lane(m1, 1, ex1, true, td);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const int row = 8 * mt + g;
            if (row < K) *reinterpret_cast<double2 *>(td_all + ((size_t) node * K + row) * C + jd) = make_double2(td[mt][0], td[mt][1]);
        }
        if (g == 0) {
            td_ex_all[(size_t) node * C + jd] = ex0;
            td_ex_all[(size_t) node * C + jd + 1] = ex1;
        }
    }
}

// Max-product bottom-up sweep (JOINT: calc_node_bu_likelihood with is_marginal = False, ml.py:124-148) for lanes that
// share one decomposition.  The max-product needs the entries of P(t') themselves, K^3 multiply-adds per (branch,
// lane) -- on the CUDA cores that is 125 ms for the 1M-tip JTT workload.  Here one warp takes one (node, lane) item and
// forms P = A (exp(d t') o A^-1) on the tensor cores, an MT x NT grid of 8 x 8 D tiles from constant A fragments and
// B fragments that are the constant A^-1 fragments scaled by this branch's exp(d_k t') (one exponential per state, computed
// by lane k of the warp and fetched with shuffles).  A thread then holds P[8 mt + g][8 nt + 2 t4 (+1)]: it multiplies
// by the child's vector, keeps its best (value, child state) per row, and the four threads of a row group finish the
// arg-max with two shuffle steps (ties to the lower state index, as np.argmax).  Column tiles are processed one after
// the other so that only MT D tiles are live.
#ifndef PML_MAXP_MINB
#define PML_MAXP_MINB 5   // (JOINT on the same workload: 3 blocks 26.1 ms, 5 blocks 22.8 ms, 6 blocks 24.2 ms)
#endif
template <int MT, int KS>
__global__ void __launch_bounds__(kMmaBlock, PML_MAXP_MINB) bu_maxp_mma_kernel(BuArgs a, int slot_begin, int n_slots) {
    const int K = a.ln.K, J = a.ln.J, C = a.ln.C;
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const double *dv = a.ln.params + 4 + K, *Am = dv + K, *Ai = Am + (size_t) K * K;
    // A fragments of A: (row 8 mt + g, col 4 s + t4); B fragments of A^-1: (row 4 s + t4, col 8 nt + g)
    // both fragment sets are the same for every warp of the block: shared memory, [fragment][lane]
    __shared__ double s_fa[MT * KS][32], s_fb[KS * MT][32];
    if (threadIdx.x < 32) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                const int r = 8 * mt + g, c4 = 4 * s + t4;
                s_fa[mt * KS + s][lane] = (r < K && c4 < K) ? Am[r * K + c4] : 0.;
                s_fb[s * MT + mt][lane] = (c4 < K && r < K) ? Ai[c4 * K + r] : 0.;
            }
    }
    __syncthreads();
    const double d_lane = (lane < K) ? dv[lane] : 0.;
    pdl_launch_dependents();
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    // the lane comes from blockIdx.y: its column and parameters are loop-invariant
    const int j = blockIdx.y;
    const int c = j % C;
    const double *pj = a.ln.params + (size_t) j * a.ln.stride;
    const double sf = pj[0], tau = pj[1], tf = pj[2];
    const bool use_alt = a.alter && tau == 0.;
    pdl_wait();
    for (int w = warp_global; w < n_slots; w += n_warps) {
        const int slot = slot_begin + w;
        const int4 top = a.t.bu_desc[2 * (size_t) slot], len = a.t.bu_desc[2 * (size_t) slot + 1];
        const int node = a.t.int_nodes[slot];
        const int fc = top.x, nc = top.y;
        double acc[MT];
        {
            const uint64_t m = eff_mask(a.mask, a.t, a.zc, C, node, c, use_alt, a.ln.kmask);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) acc[mt] = (8 * mt + g < K) ? bit_as_double(m, 8 * mt + g) : 0.;
        }
        int ex = 0, bad = -1;
        for (int q = 0; q < nc; ++q) {
            const int ch = fc + q;
            const int cs = (q == 0) ? top.z : (q == 1) ? top.w : a.t.slot[ch];
            const double dist = (q == 0) ? __hiloint2double(len.y, len.x) : (q == 1) ? __hiloint2double(len.w, len.z) : a.t.dist[ch];
            const double tt = (dist + tau) * tf * sf;
            // exp(d_k t'): lane k computes state k's
            double e_lane;
            {
                double xs[1] = {d_lane * tt}, es[1];
                exp_nonpos_n<1>(xs, es);
                e_lane = es[0];
            }
            double ek[KS];
#pragma unroll
            for (int s = 0; s < KS; ++s) ek[s] = __shfl_sync(kFullMask, e_lane, (4 * s + t4) & 31);
            // the child's vector at this thread's columns
            const uint64_t cm = (cs >= 0) ? 0 : eff_mask(a.mask, a.t, a.zc, C, ch, c, use_alt, a.ln.kmask);
            const double *src = a.bu + (size_t) (cs >= 0 ? cs : 0) * K * J + j;
            double best[MT];
            int arg[MT];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) { best[mt] = -CUDART_INF; arg[mt] = 0; }
#pragma unroll
            for (int nt = 0; nt < MT; ++nt) {
                const int i0 = 8 * nt + 2 * t4, i1 = i0 + 1;
                double l0 = 0., l1 = 0.;
                if (cs >= 0) {
                    if (i0 < K) l0 = src[(size_t) i0 * J];
                    if (i1 < K) l1 = src[(size_t) i1 * J];
                } else {
                    l0 = bit_as_double(cm, i0);
                    l1 = bit_as_double(cm, i1);
                }
                double bfr[KS];   // B fragments of exp(d t') o A^-1 for this column tile
#pragma unroll
                for (int s = 0; s < KS; ++s) bfr[s] = s_fb[s * MT + nt][lane] * ek[s];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    double p0 = 0., p1 = 0.;
#pragma unroll
                    for (int s = 0; s < KS; ++s) dmma(p0, p1, s_fa[mt * KS + s][lane], bfr[s]);
                    const double x0 = __dmul_rn(p0, l0), x1 = __dmul_rn(p1, l1);
                    if (i0 < K && x0 > best[mt]) { best[mt] = x0; arg[mt] = i0; }
                    if (i1 < K && x1 > best[mt]) { best[mt] = x1; arg[mt] = i1; }
                }
            }
            // finish the arg-max over the four threads of a row group (columns interleave: lower index wins ties)
            double vmax = 0.;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                for (int sh = 1; sh <= 2; sh <<= 1) {
                    const double ob = __shfl_xor_sync(kFullMask, best[mt], sh);
                    const int oa = __shfl_xor_sync(kFullMask, arg[mt], sh);
                    if (ob > best[mt] || (ob == best[mt] && oa < arg[mt])) { best[mt] = ob; arg[mt] = oa; }
                }
                const int row = 8 * mt + g;
                if (a.ptr && t4 == 0 && row < K) a.ptr[((size_t) ch * C + c) * K + row] = (uint8_t) arg[mt];
                acc[mt] *= (row < K) ? fmax(best[mt], 0.) : 0.;
                vmax = fmax(vmax, acc[mt]);
            }
            vmax = colmax(vmax);
            if (cs >= 0) ex += a.bu_ex[(size_t) cs * J + j];
            if (vmax == 0. && bad < 0) bad = ch;
            const bool last = q == nc - 1;
            if (vmax > 0. && !isinf(vmax)) {
                int e = exponent_of(vmax);
                if (last || e < -kKeep || e > kKeep) {
                    int extra = 0;
                    double lift = 1.;
                    if (e == -1023) { lift = 0x1p+200; e = exponent_of(vmax * 0x1p+200); extra = -200; }
                    const double s1 = pow2(-(e / 2)), s2 = pow2(-(e - e / 2));
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) acc[mt] = ((acc[mt] * lift) * s1) * s2;
                    ex += e + extra;
                }
            }
        }
        if (t4 == 0) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
                if (8 * mt + g < K) a.bu[((size_t) slot * K + 8 * mt + g) * J + j] = acc[mt];
        }
        if (lane == 0) {
            a.bu_ex[(size_t) slot * J + j] = ex;
            if (bad >= 0 && atomicCAS(a.err + 2 * j, -1, node) == -1) a.err[2 * j + 1] = bad;
        }
    }
}

}  // namespace

bool mma_maxp_applicable(int K) { return K > 8 && K <= kMmaKP; }

void mma_maxp_launch(cudaStream_t st, const BuArgs &a, int s0, int ns) {
    long long g = ((long long) ns + kMmaBlock / 32 - 1) / (kMmaBlock / 32);
    const long long cap = std::max(1LL, 148LL * PML_MAXP_MINB / a.ln.J);   // persistent grid, blockIdx.y = lane
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    const int K = a.ln.K;
    const dim3 grid((unsigned) g, (unsigned) a.ln.J);
    if (K <= 16) launch_pdl(bu_maxp_mma_kernel<2, 4>, grid, dim3(kMmaBlock), 0, st, a, s0, ns);
    else if (K <= 20) launch_pdl(bu_maxp_mma_kernel<3, 5>, grid, dim3(kMmaBlock), 0, st, a, s0, ns);
    else launch_pdl(bu_maxp_mma_kernel<3, 6>, grid, dim3(kMmaBlock), 0, st, a, s0, ns);
}

namespace {
}  // namespace

bool mma_bu_applicable(int K, int J) { return K >= 2 && K <= kMmaKP && J >= 8 && J % 8 == 0; }

void mma_bu_launch(cudaStream_t st, const BuArgs &a, int s0, int ns) {
    const int tiles = a.ln.J / 8;
    long long g = ((long long) ns + kMmaBlock / 32 - 1) / (kMmaBlock / 32);
    const long long cap = std::max(1LL, 148LL * kMmaMinBlocks / tiles);   // persistent grid, blockIdx.y = lane tile
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    const int K = a.ln.K;
    const dim3 grid((unsigned) g, (unsigned) tiles);
    if (K <= 8) launch_pdl(bu_mma_kernel<1, 2>, grid, dim3(kMmaBlock), 0, st, a, s0, ns);
    else if (K <= 16) launch_pdl(bu_mma_kernel<2, 4>, grid, dim3(kMmaBlock), 0, st, a, s0, ns);
    else if (K <= 20) launch_pdl(bu_mma_kernel<3, 5>, grid, dim3(kMmaBlock), 0, st, a, s0, ns);
    else launch_pdl(bu_mma_kernel<3, 6>, grid, dim3(kMmaBlock), 0, st, a, s0, ns);
}

}  // namespace pml

namespace pml {
void mma_td_launch(cudaStream_t st, const TdArgs &a, const double *v_all, double *td_all, int *td_ex_all, int n0, int nn) {
    const int tiles = a.ln.C / 8;
    long long g = ((long long) nn + kMmaBlock / 32 - 1) / (kMmaBlock / 32);
    const long long cap = std::max(1LL, 148LL * kMmaMinBlocks / tiles);   // persistent grid, blockIdx.y = lane tile
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    const int K = a.ln.K;
    const dim3 grid((unsigned) g, (unsigned) tiles);
    if (K <= 8) launch_pdl(td_mma_kernel<1, 2>, grid, dim3(kMmaBlock), 0, st, a, v_all, td_all, td_ex_all, n0, nn);
    else if (K <= 16) launch_pdl(td_mma_kernel<2, 4>, grid, dim3(kMmaBlock), 0, st, a, v_all, td_all, td_ex_all, n0, nn);
    else if (K <= 20) launch_pdl(td_mma_kernel<3, 5>, grid, dim3(kMmaBlock), 0, st, a, v_all, td_all, td_ex_all, n0, nn);
    else launch_pdl(td_mma_kernel<3, 6>, grid, dim3(kMmaBlock), 0, st, a, v_all, td_all, td_ex_all, n0, nn);
}
}  // namespace pml
