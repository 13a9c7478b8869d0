// Sampling of ancestral scenarios from the marginal posterior on the device: the K x K expected numbers of state
// transitions of pastml.ml.marginal_counts (/root/reference/pastml/ml.py:753-862).
//
// The reference draws, per node, how many of the n_repetitions scenarios are in each state: the root's counts from
// its marginal probabilities, then for every branch and every parent state j with count c_j a multinomial of c_j
// draws from  P(child = i | parent = j) ~ L_child[i] P_ij(t) pi_i allowed_child[i]  (ml.py:820-838).  Here every
// repetition keeps its own state along the tree (one byte per internal node and repetition, two levels alive at a
// time): a child's state is drawn from row (parent's state) of the same table, which gives the same distribution of
// the per-node counts and of everything derived from them.  Random numbers: Philox4x32-10 keyed by the seed and
// counted by (repetition, node, column), so a result does not depend on the launch geometry.
//
// One block works on one (parent, column) at a time: conditional tables of a child [K][K] in shared memory (thread j
// computes row j with the model's own matrix-vector product: P e_j is column j of P), repetitions spread over the
// threads, transitions counted in a shared [K][K] histogram, then the count-level bookkeeping of the reference
// (altered zero-branch clusters redistribute counts instead of adding transitions, ml.py:806-852) and one flush of the
// block's K x K partial result at the end.
#pragma once

#include "pml_kernels.cuh"

namespace pml {

struct CountsArgs {
    TreeDev t;
    ZcDev zc;
    ExtDev ext;
    LaneDev ln;             // one parameter set: lanes = columns
    const uint64_t *mask;   // annotation masks [N][C]
    const double *bu;       // bottom-up vectors of the preceding sum-product sweep
    const int *bu_ex;
    int n_rep;
    unsigned long long seed;
    const uint8_t *st_cur;  // [level width][C][n_rep] states of this level's internal nodes (row = slot - slot_begin)
    uint8_t *st_next;       // the same for the next level (row = slot - next_slot_begin)
    int slot_begin, n_slots, next_slot_begin;
    double *result;         // [C][K][K], accumulated
};

__device__ __forceinline__ double philox_uniform(unsigned long long seed, unsigned c0, unsigned c1, unsigned c2) {
    unsigned k0 = (unsigned) seed, k1 = (unsigned) (seed >> 32);
    unsigned x0 = c0, x1 = c1, x2 = c2, x3 = 0x243F6A88u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
        const unsigned y0 = hi1 ^ x1 ^ k0, y2 = hi0 ^ x3 ^ k1;
        x0 = y0; x1 = lo1; x2 = y2; x3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    const unsigned long long bits = (((unsigned long long) x0 << 32) | x1) >> 11;
    return (double) bits * 0x1p-53;   // [0, 1)
}

// first i with u < cdf[i] (cdf non-decreasing, cdf[K-1] == 1)
__device__ __forceinline__ int draw_from(const double *cdf, int K, double u) {
    int i = 0;
    while (i < K - 1 && u >= cdf[i]) ++i;
    return i;
}

__host__ __device__ inline size_t counts_smem_bytes(int family, int K, int C) {
    // staged parameters | cdf [K][K] | result [K][K] | psci, same, sc, norm [4][K] doubles | trans [K][K] | pc [K] ints
    return sizeof(double) * (staged_doubles(family, K, C) + 2 * (size_t) K * K + 4 * (size_t) K) +
           sizeof(int) * ((size_t) K * K + K + 4);
}

// states of the (internal) roots: drawn from bu * pi * allowed  (the root's top-down vector is all ones, ml.py:789-797)
template <int KP, template <int> class Lane>
__global__ void __launch_bounds__(128) counts_root_kernel(CountsArgs a) {
    extern __shared__ double s_par[];
    const int C = a.ln.C, K = a.ln.K, c = blockIdx.y;
    const int nf = staged_fields(Lane<KP>::family, K);
    const bool with_eig = eig_staged(Lane<KP>::family, K, C);
    double *cdf = s_par + staged_doubles(Lane<KP>::family, K, C);
    stage_params(s_par, a.ln, nf, with_eig);
    Lane<KP> lane;
    lane.init(LaneSmem{s_par + c, C}, a.ln.params + (size_t) c * a.ln.stride, K,
              with_eig ? s_par + (size_t) nf * C + (size_t) c * eig_lane_stride(K) : nullptr);
    const bool use_alt = lane.tau == 0.;
    for (int s = a.slot_begin + blockIdx.x; s < a.slot_begin + a.n_slots; s += gridDim.x) {
        const int node = a.t.int_nodes[s];
        __syncthreads();
        if (threadIdx.x == 0) {
            double L[KP];
            uint64_t m;
            load_child<KP>(a.t, a.zc, a.ext, a.ln, a.mask, a.bu, a.bu_ex, node, c, c, use_alt, L, m);
            m = eff_mask(a.mask, a.t, a.zc, C, node, c, use_alt, a.ln.kmask);
            double tot = 0.;
            for (int k = 0; k < K; ++k) {
                tot += L[k] * lane.pi(k) * bit_as_double(m, k);
                cdf[k] = tot;
            }
            for (int k = 0; k < K; ++k) cdf[k] = (k == K - 1) ? 1. : cdf[k] / tot;
        }
        __syncthreads();
        uint8_t *dst = a.st_next + ((size_t) (s - a.next_slot_begin) * C + c) * a.n_rep;
        for (int r = threadIdx.x; r < a.n_rep; r += blockDim.x)
            dst[r] = (uint8_t) draw_from(cdf, K, philox_uniform(a.seed, (unsigned) r, (unsigned) node, (unsigned) c));
    }
}

template <int KP, template <int> class Lane>
__global__ void __launch_bounds__(128) counts_level_kernel(CountsArgs a) {
    extern __shared__ double s_par[];
    const int C = a.ln.C, K = a.ln.K, c = blockIdx.y, KK = K * K;
    const int nf = staged_fields(Lane<KP>::family, K);
    const bool with_eig = eig_staged(Lane<KP>::family, K, C);
    double *cdf = s_par + staged_doubles(Lane<KP>::family, K, C);
    double *res = cdf + KK, *psci = res + KK, *same = psci + K, *sc = same + K, *norm = sc + K;
    int *trans = reinterpret_cast<int *>(norm + K), *pc = trans + KK, *flags = pc + K;
    stage_params(s_par, a.ln, nf, with_eig);
    Lane<KP> lane;
    lane.init(LaneSmem{s_par + c, C}, a.ln.params + (size_t) c * a.ln.stride, K,
              with_eig ? s_par + (size_t) nf * C + (size_t) c * eig_lane_stride(K) : nullptr);
    const bool use_alt = lane.tau == 0.;
    const double n_rep = (double) a.n_rep;
    for (int e = threadIdx.x; e < KK; e += blockDim.x) res[e] = 0.;

    // is (node, c) altered by the zero-branch alteration, and its '.initial' states
    auto altered = [&](int node, uint64_t &initial) {
        initial = 0;
        if (!use_alt || !(a.mask[(size_t) node * C + c] & kMemberFlag)) return false;
        const size_t row = (size_t) a.t.zc_index[node] * C + c;
        if (!a.zc.altflag[row]) return false;
        initial = a.zc.initial[row] & a.ln.kmask;
        return true;
    };
    // counts restricted to the node's own states, rescaled to n_rep (ml.py:806-812, 846-852); in: v[K], out: v[K]
    auto redistribute = [&](double *v, uint64_t initial) {   // thread 0 only
        double tot = 0.;
        for (int k = 0; k < K; ++k) { v[k] *= bit_as_double(initial, k); tot += v[k]; }
        if (tot > 0.) { for (int k = 0; k < K; ++k) v[k] = n_rep * v[k] / tot; }
        else {
            const double n_init = (double) __popcll(initial);
            for (int k = 0; k < K; ++k) v[k] = n_rep * bit_as_double(initial, k) / n_init;
        }
    };

    for (int s = a.slot_begin + blockIdx.x; s < a.slot_begin + a.n_slots; s += gridDim.x) {
        const int parent = a.t.int_nodes[s];
        const uint8_t *st_p = a.st_cur + ((size_t) (s - a.slot_begin) * C + c) * a.n_rep;
        __syncthreads();
        for (int k = threadIdx.x; k < K; k += blockDim.x) { pc[k] = 0; same[k] = 0.; }
        __syncthreads();
        for (int r = threadIdx.x; r < a.n_rep; r += blockDim.x) atomicAdd(pc + st_p[r], 1);
        __syncthreads();
        uint64_t p_initial;
        const bool p_alt = altered(parent, p_initial);
        if (threadIdx.x == 0) {
            for (int k = 0; k < K; ++k) psci[k] = (double) pc[k];
            if (p_alt) redistribute(psci, p_initial);
        }
        const int first = a.t.first_child[parent], n_ch = a.t.n_children[parent];
        for (int node = first; node < first + n_ch; ++node) {
            __syncthreads();
            // ---- conditional tables: row j = cumulative P(child = . | parent = j)
            if (threadIdx.x < K) {
                const int j = threadIdx.x;
                double L[KP], ej[KP], col[KP];
                uint64_t m;
                load_child<KP>(a.t, a.zc, a.ext, a.ln, a.mask, a.bu, a.bu_ex, node, c, c, use_alt, L, m);
                m = eff_mask(a.mask, a.t, a.zc, C, node, c, use_alt, a.ln.kmask);
#pragma unroll
                for (int k = 0; k < KP; ++k) ej[k] = (k == j) ? 1. : 0.;
                const typename Lane<KP>::Branch br = lane.branch(a.t.dist[node], K);
                lane.matvec(br, ej, col, K);   // col[i] = P_ij(t): transpose(get_Pij_t)[j][i], ml.py:821
                double tot = 0.;
#pragma unroll
                for (int k = 0; k < KP; ++k)
                    if (k < K) {
                        tot += L[k] * col[k] * lane.pi(k) * bit_as_double(m, k);
                        cdf[j * K + k] = tot;
                    }
                for (int k = 0; k < K; ++k) cdf[j * K + k] = (k == K - 1) ? 1. : cdf[j * K + k] / tot;
            }
            for (int e = threadIdx.x; e < KK; e += blockDim.x) trans[e] = 0;
            __syncthreads();
            // ---- one draw per repetition
            const int cs = a.t.slot[node];
            uint8_t *st_c = cs >= 0 ? a.st_next + ((size_t) (cs - a.next_slot_begin) * C + c) * a.n_rep : nullptr;
            for (int r = threadIdx.x; r < a.n_rep; r += blockDim.x) {
                const int j = st_p[r];
                const int i = draw_from(cdf + j * K, K, philox_uniform(a.seed, (unsigned) r, (unsigned) node, (unsigned) c));
                atomicAdd(trans + j * K + i, 1);
                if (st_c) st_c[r] = (uint8_t) i;
            }
            __syncthreads();
            // ---- bookkeeping on the counts (ml.py:826-858)
            uint64_t n_initial;
            const bool n_alt = altered(node, n_initial);
            if (!p_alt && !n_alt) {
                for (int e = threadIdx.x; e < KK; e += blockDim.x) res[e] += (double) trans[e];
                for (int k = threadIdx.x; k < K; k += blockDim.x) same[k] += (double) trans[k * K + k];
            } else {
                for (int i = threadIdx.x; i < K; i += blockDim.x) {
                    int n = 0;
                    for (int j = 0; j < K; ++j) n += trans[j * K + i];
                    sc[i] = (double) n;
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    if (n_alt) redistribute(sc, n_initial);
                    double tot = 0.;
                    for (int k = 0; k < K; ++k) tot += sc[k];
                    for (int k = 0; k < K; ++k) norm[k] = sc[k] / tot;
                }
                __syncthreads();
                for (int e = threadIdx.x; e < KK; e += blockDim.x) {
                    const int i = e / K, k = e - i * K;
                    if (psci[i] > 0.) res[e] += norm[k] * psci[i];
                }
                for (int i = threadIdx.x; i < K; i += blockDim.x)
                    if (psci[i] > 0.) same[i] += norm[i] * psci[i];
            }
        }
        __syncthreads();
        for (int k = threadIdx.x; k < K; k += blockDim.x) res[k * K + k] -= fmin(psci[k], same[k]);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < KK; e += blockDim.x)
        if (res[e] != 0.) atomicAdd(a.result + (size_t) c * KK + e, res[e]);
    (void) flags;
}

}  // namespace pml
