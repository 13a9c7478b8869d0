// Device code of the pastml_b200 engine (sm_100a).  All arithmetic is fp64 on the CUDA cores: the hot path is
// HBM-bound (O(K) flops per node-state for the F81 family), so nothing here is shaped into a GEMM.
//
// Work decomposition: one thread per (internal node of a depth level, lane) where a lane is one
// (parameter set, column) pair; the K-vector of the node lives in registers.  Partial likelihoods are stored
// k-major / lane-minor, [slot][k][lane], so that the lanes of a warp read and write contiguous runs.
//
// Scaling: the reference keeps log10 likelihoods and shifts them by real-valued factors (rescale_log,
// /root/reference/pastml/ml.py:151-171).  Here a node vector is a mantissa vector whose maximum is kept in [1, 2)
// plus an exact power-of-two exponent (int32): true value = mant * 2^ex.  The likelihood is invariant to the
// choice of shifts, so results agree with the reference to rounding while the per-child log10/pow calls vanish.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>

namespace pml {

constexpr uint64_t kMemberFlag = 1ULL << 63;  // (node, column) belongs to a zero-branch cluster (see alter_kernel)
constexpr uint64_t kExtFlag = 1ULL << 62;     // tip whose bottom-up vector is supplied from outside (a clade cut off by
                                              // subtree sharding): see ExtDev
#ifndef PML_BLOCK
#define PML_BLOCK 128
#endif
// (instantiations with at most 4 states ask for 6 blocks per SM: HKY on 1M tips 2.09 -> 1.93 ms per step)
#ifndef PML_BU_MINB
#define PML_BU_MINB 3
#endif
#ifndef PML_TD_MINB
#define PML_TD_MINB 3
#endif
#ifndef PML_TDSEL_MINB
#define PML_TDSEL_MINB 3   // td_select_kernel: issue-bound, more resident warps pay (measured, profiles/README.md r2_w)
#endif
constexpr int kBlock = PML_BLOCK;

struct TreeDev {
    int n_nodes;
    int n_internal;
    const int *parent;       // [N]
    const int *first_child;  // [N]
    const int *n_children;   // [N]
    const int *slot;         // [N]  rank among internal nodes or -1
    const int *int_nodes;    // [NI] slot -> node
    const double *dist;      // [N]
    const int *zc_index;     // [N]  row in the zero-cluster table or -1
    const int *fc_slot;      // [NI] first child of the slot's node (saves a dependent load in the sweeps)
    const int *nc_slot;      // [NI] number of children of the slot's node
    const int *pslot;        // [N]  slot of the parent, -1 for roots
    int n_root_slots;        // internal roots occupy slots [0, n_root_slots)
    const int4 *bu_desc;     // [NI][2] packed {first child, #children, slot of child 0, slot of child 1} +
                             //         {dist of child 0, dist of child 1}: everything the bottom-up step needs to
                             //         issue its vector loads after ONE dependent load
    const int4 *td_desc;     // [N]     packed {parent slot, own slot, dist (2 words)} for the top-down step
};

struct ZcDev {              // zero-branch cluster table, rows x C
    const uint64_t *alt;     // effective mask of the member for the current pass
    const uint8_t *altflag;  // member was altered by the current pass' alteration
    const uint64_t *initial; // '<col>_ALLOWED_STATES.initial'
    const uint8_t *hasinit;
    const uint8_t *touched;  // '.initial' was rewritten since the last marginal sweep
};

// exact n / d for 0 <= n < 2^31 with one 32x32->64 multiply (m = floor(2^(31+s) / d) + 1, s = ceil(log2 d))
struct FastDiv {
    unsigned m, s, d;
    __host__ static FastDiv make(unsigned d) {
        unsigned s = 0;
        while ((1ull << s) < d) ++s;
        return FastDiv{(unsigned) (((1ull << (31 + s)) / d) + 1ull), s, d};
    }
    __device__ __forceinline__ unsigned div(unsigned n) const {
        return (unsigned) (((unsigned long long) n * m) >> (31 + s));
    }
};

struct LaneDev {
    int C;                   // columns
    int J;                   // lanes = sets * C
    int K;                   // states
    int stride;              // doubles per lane in params
    const double *params;    // [J][stride]
    uint64_t kmask;          // (1 << K) - 1
    FastDiv divJ, divC;
    int jt;                  // 0: bottom-up vectors laid out [slot][k][J]; else tile-major [slot][J / jt][k][jt]
                             // (passes of the TMA-staged bottom-up sweep with more lanes than one tile)
    const int *lane_col;     // null: lane j works on column j % C; else an explicit column per lane (objective passes
                             // over the columns that are still being fitted, pml_loglik_lanes)
};

__device__ __forceinline__ int lane_column(const LaneDev &ln, int j) { return ln.lane_col ? ln.lane_col[j] : j % ln.C; }

// index of state k, lane j of a slot's bottom-up vector in the layout of this pass
__device__ __forceinline__ size_t bu_at(const LaneDev &ln, int slot, int k, int j) {
    if (ln.jt > 0 && ln.jt < ln.J) {
        const int t = j / ln.jt, jl = j - t * ln.jt;
        return (((size_t) slot * (ln.J / ln.jt) + t) * ln.K + k) * ln.jt + jl;
    }
    return ((size_t) slot * ln.K + k) * ln.J + j;
}

// Launches cover at most this many (node, lane) items, so that item indices fit 31 bits
constexpr long long kMaxItemsPerLaunch = 1LL << 30;

// ------------------------------------------------------------------------------------------ helpers

// Programmatic dependent launch: the level kernels of a sweep are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so the blocks of level N + 1 become resident while level N drains and
// run their prologue (parameter staging, barrier set-up, first descriptors -- nothing level N produces) before
// pdl_wait(), which returns once level N has completed and its stores are visible.  Without the launch attribute both
// instructions are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ int exponent_of(double x) {  // floor(log2 x) for positive normal x
    return ((__double2hiint(x) >> 20) & 0x7ff) - 1023;
}

__device__ __forceinline__ double pow2(int e) {  // exact 2^e for -1022 <= e <= 1023
    return __hiloint2double((e + 1023) << 20, 0);
}

__device__ __forceinline__ uint64_t eff_mask(const uint64_t *base, const TreeDev &t, const ZcDev &zc, int C,
                                             int node, int c, bool use_alt, uint64_t kmask) {
    uint64_t m = base[(size_t) node * C + c];
    if (use_alt && (m & kMemberFlag)) m = zc.alt[(size_t) t.zc_index[node] * C + c];
    return m & kmask;
}

// 1.0 where bit k of m is set, else 0.0, without an int -> double conversion
__device__ __forceinline__ double bit_as_double(uint64_t m, int k) {
    const unsigned w = (k < 32) ? (unsigned) m : (unsigned) (m >> 32);
    return __hiloint2double(((w >> (k & 31)) & 1u) ? 0x3ff00000 : 0, 0);
}

// Keeps a non-negative vector in range: when the exponent of its maximum leaves [-kKeep, kKeep] (or `always`)
// the vector is multiplied by an exact power of two that brings the maximum into [1, 2).  Returns the exponent
// removed.  vmax = 0 (or NaN) is left alone: the caller reports the all-zero vector.
constexpr int kKeep = 96;
template <int KP>
__device__ __forceinline__ int rescale(double (&v)[KP], double vmax, bool always) {
    if (!(vmax > 0.) || isinf(vmax)) return 0;
    int e = exponent_of(vmax);
    if (!always && e >= -kKeep && e <= kKeep) return 0;
    int extra = 0;
    if (e == -1023) {  // subnormal maximum: lift it first
#pragma unroll
        for (int k = 0; k < KP; ++k) v[k] *= 0x1p+200;
        e = exponent_of(vmax * 0x1p+200);
        extra = -200;
    }
    if (e != 0) {
        // 2^-e alone may be out of range: apply it in two exact halves
        const int e1 = e / 2, e2 = e - e1;
        const double s1 = pow2(-e1), s2 = pow2(-e2);
#pragma unroll
        for (int k = 0; k < KP; ++k) v[k] = (v[k] * s1) * s2;
    }
    return e + extra;
}

template <int KP>
__device__ __forceinline__ double vec_max(const double (&v)[KP]) {
    // two chains to shorten the dependency path
    double m0 = 0., m1 = 0.;
#pragma unroll
    for (int k = 0; k + 1 < KP; k += 2) { m0 = fmax(m0, v[k]); m1 = fmax(m1, v[k + 1]); }
    if (KP & 1) m0 = fmax(m0, v[KP - 1]);
    return fmax(m0, m1);
}

// ------------------------------------------------------------------------------------------ models
// Per-lane parameters are staged once per block in shared memory, transposed to [field][lane] so that the lanes of
// a warp read consecutive words.  A lane model exposes: branch(dist) -> Branch, matvec(out = max(P(t) in, 0)),
// maxvec(out_i = max_j P_ij in_j with the first argmax), pi(k).  Entries k >= K of every register vector are zero
// and stay zero, so arithmetic runs unguarded over the padded length KP and only memory accesses are guarded.

// Subtree sharding: when one big tree is cut into clades owned by different GPUs, the part of
// the tree above the cut sees every clade root as a tip whose bottom-up vector comes from another problem (possibly
// another GPU, after an all-gather), and every clade forest sees top-down vectors arriving at its roots.
struct ExtDev {
    const int *index;       // [N] tip node -> row of the tables below (valid where the mask carries kExtFlag)
    const double *bu;       // [rows][K][J] bottom-up mantissas of the external tips
    const int *bu_ex;       // [rows][J]
    double *td;             // [rows][K][C] top-down vectors computed for the external tips (output of the sweep)
    int *td_ex;             // [rows][C]
    const double *root_td;  // [T][K][C] top-down vectors arriving at the roots, or null (roots start from all ones)
    const int *root_td_ex;  // [T][C]
};

struct LaneSmem {
    const double *s;   // staged fields of this lane: field f at s[f * J]
    int J;
    __device__ __forceinline__ double f(int i) const { return s[i * J]; }
};

// number of leading parameter fields staged in shared memory ([field][lane])
__host__ __device__ inline int staged_fields(int family, int K) { return family == 1 ? 8 : 4 + K; }

// The eigen-decomposition models also stage the eigenvalues and both matrices (2 K + 2 K^2 doubles per lane, 6.6 KB for
// K = 20) when they fit: every matrix element is read once per matrix-vector product, by all the nodes a warp holds
// for that lane.  Lane-major ([lane][element], so that the unrolled products keep immediate offsets), lane stride
// = 1 mod 16 doubles: the 8 or 16 lanes of a warp fall into distinct banks.
constexpr size_t kEigStageBytes = 160 * 1024;
__host__ __device__ inline int eig_elements(int K) { return 2 * K + 2 * K * K; }
__host__ __device__ inline int eig_lane_stride(int K) {
    const int n = eig_elements(K);
    return n + (17 - n % 16) % 16;
}
__host__ __device__ inline bool eig_staged(int family, int K, int J) {
    return family == 2 && ((size_t) staged_fields(family, K) + eig_lane_stride(K)) * J * sizeof(double) <= kEigStageBytes;
}
// doubles of the staging area of a block
__host__ __device__ inline size_t staged_doubles(int family, int K, int J) {
    return (size_t) staged_fields(family, K) * J + (eig_staged(family, K, J) ? (size_t) eig_lane_stride(K) * J : 0);
}

__device__ __forceinline__ void stage_params(double *smem, const LaneDev &ln, int nf, bool with_eig = false) {
    const int total = nf * ln.J;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int f = i / ln.J, j = i - f * ln.J;
        smem[i] = ln.params[(size_t) j * ln.stride + f];
    }
    if (with_eig) {   // [lane][d, A, A^-1] behind the fields
        const int ne = eig_elements(ln.K), ls = eig_lane_stride(ln.K);
        double *dst = smem + total;
        for (int i = threadIdx.x; i < ne * ln.J; i += blockDim.x) {
            const int j = i / ne, e = i - j * ne;
            dst[(size_t) j * ls + e] = ln.params[(size_t) j * ln.stride + 4 + ln.K + e];
        }
    }
    __syncthreads();
}

template <int KP>
struct F81Lane {  // F81 / JC / EFT: P_ij = (1 - e) pi_j + delta_ij e, e = exp(-mu t')  (F81Model.py:28-46)
    static constexpr int family = 0;
    LaneSmem sm;
    double sf, tau, tauf, mu;
    struct Branch { double e, ome; };

    __device__ __forceinline__ void init(const LaneSmem &l, const double *, int, const double *) {
        sm = l;
        sf = l.f(0); tau = l.f(1); tauf = l.f(2); mu = l.f(3);
    }
    __device__ __forceinline__ double pi(int k) const { return sm.f(4 + k); }
    __device__ __forceinline__ Branch branch(double dist, int) const {
        const double t = (dist + tau) * tauf * sf;                 // Model.transform_t
        const double e = isinf(mu) ? 0. : exp(-mu * t);            // F81Model.py:44-45
        return Branch{e, 1. - e};
    }
    __device__ __forceinline__ double weighted_sum(const double (&in)[KP], int K) const {
        double s0 = 0., s1 = 0.;
#pragma unroll
        for (int k = 0; k + 1 < KP; k += 2) {
            if (k < K) s0 = fma(pi(k), in[k], s0);
            if (k + 1 < K) s1 = fma(pi(k + 1), in[k + 1], s1);
        }
        if ((KP & 1) && KP - 1 < K) s0 = fma(pi(KP - 1), in[KP - 1], s0);
        return s0 + s1;
    }
    __device__ __forceinline__ void matvec(const Branch &b, const double (&in)[KP], double (&out)[KP], int K) const {
        const double base = b.ome * weighted_sum(in, K);
#pragma unroll
        for (int k = 0; k < KP; ++k) out[k] = (k < K) ? fmax(fma(b.e, in[k], base), 0.) : 0.;
    }
    // acc *= max(P in, 0) without a temporary vector
    __device__ __forceinline__ void mul_matvec(const Branch &b, const double (&in)[KP], double (&acc)[KP], int K) const {
        const double base = b.ome * weighted_sum(in, K);
#pragma unroll
        for (int k = 0; k < KP; ++k) acc[k] *= fmax(fma(b.e, in[k], base), 0.);
    }
    // max_j P_ij in_j in O(K): off-diagonal candidates are w_j = ((1-e) pi_j) in_j for every row, the diagonal one
    // is ((1-e) pi_i + e) in_i; products are formed exactly like numpy forms P * L (no contraction).
    __device__ __forceinline__ void maxvec(const Branch &b, const double (&in)[KP], double (&out)[KP],
                                           uint8_t (&arg)[KP], int K) const {
        double t1 = -CUDART_INF, t2 = -CUDART_INF;
        int i1 = 0, i2 = 0;
#pragma unroll
        for (int k = 0; k < KP; ++k)
            if (k < K) {
                const double w = __dmul_rn(__dmul_rn(b.ome, pi(k)), in[k]);
                if (w > t1) { t2 = t1; i2 = i1; t1 = w; i1 = k; }
                else if (w > t2) { t2 = w; i2 = k; }
            }
#pragma unroll
        for (int k = 0; k < KP; ++k)
            if (k < K) {
                const double dv = __dmul_rn(__dadd_rn(__dmul_rn(b.ome, pi(k)), b.e), in[k]);
                const double ov = (i1 == k) ? t2 : t1;
                const int oi = (i1 == k) ? i2 : i1;
                double best; int bi;
                if (K == 1 || dv > ov) { best = dv; bi = k; }
                else if (dv < ov) { best = ov; bi = oi; }
                else { best = dv; bi = k < oi ? k : oi; }
                out[k] = fmax(best, 0.);
                arg[k] = (uint8_t) bi;
            } else { out[k] = 0.; arg[k] = 0; }
    }
};

template <int KP>
struct HKYLane {  // HKY85 closed form, states A, C, G, T  (HKYModel.py:44-82); KP == 4
    static constexpr int family = 1;
    LaneSmem sm;
    double sf, tau, tauf, kappa;
    double p4[4];
    struct Branch { double P[4][4]; };

    __device__ __forceinline__ void init(const LaneSmem &l, const double *, int, const double *) {
        sm = l;
        sf = l.f(0); tau = l.f(1); tauf = l.f(2); kappa = l.f(3);
#pragma unroll
        for (int k = 0; k < 4; ++k) p4[k] = l.f(4 + k);
    }
    __device__ __forceinline__ double pi(int k) const { return sm.f(4 + k); }
    __device__ __forceinline__ Branch branch(double dist, int) const {
        const double t = (dist + tau) * tauf * sf;
        const double pa = p4[0], pc = p4[1], pg = p4[2], pt = p4[3];
        const double pag = pa + pg, pct = pc + pt;
        const double beta = .5 / (pag * pct + kappa * (pa * pg + pc * pt));
        const double e1 = exp(-beta * t);
        const double ect = exp(-beta * t * (1. + pct * (kappa - 1.))) / pct;
        const double eag = exp(-beta * t * (1. + pag * (kappa - 1.))) / pag;
        const double ct_sum = (pct + pag * e1) / pct;
        const double ag_sum = (pag + pct * e1) / pag;
        Branch b;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) b.P[i][j] = (1. - e1) * p4[j];
        b.P[3][3] = pt * ct_sum + pc * ect;
        b.P[3][1] = pc * ct_sum - pc * ect;
        b.P[1][3] = pt * ct_sum - pt * ect;
        b.P[1][1] = pc * ct_sum + pt * ect;
        b.P[0][0] = pa * ag_sum + pg * eag;
        b.P[0][2] = pg * ag_sum - pg * eag;
        b.P[2][0] = pa * ag_sum - pa * eag;
        b.P[2][2] = pg * ag_sum + pa * eag;
        return b;
    }
    __device__ __forceinline__ void matvec(const Branch &b, const double (&in)[KP], double (&out)[KP], int) const {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double s = b.P[i][0] * in[0];
#pragma unroll
            for (int j = 1; j < 4; ++j) s = fma(b.P[i][j], in[j], s);
            out[i] = fmax(s, 0.);
        }
    }
    __device__ __forceinline__ void mul_matvec(const Branch &b, const double (&in)[KP], double (&acc)[KP], int K) const {
        double v[KP];
        matvec(b, in, v, K);
#pragma unroll
        for (int k = 0; k < KP; ++k) acc[k] *= v[k];
    }
    __device__ __forceinline__ void maxvec(const Branch &b, const double (&in)[KP], double (&out)[KP],
                                           uint8_t (&arg)[KP], int) const {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double best = __dmul_rn(b.P[i][0], in[0]);
            int bi = 0;
#pragma unroll
            for (int j = 1; j < 4; ++j) {
                const double x = __dmul_rn(b.P[i][j], in[j]);
                if (x > best) { best = x; bi = j; }
            }
            out[i] = fmax(best, 0.);
            arg[i] = (uint8_t) bi;
        }
    }
};

template <int KP>
struct EigLane {  // CUSTOM_RATES / JTT: P = A diag(exp(d t')) A^-1  (generator.py:54-65)
    static constexpr int family = 2;
    LaneSmem sm;
    double sf, tau, tauf;
    const double *d, *A, *Ainv;  // this lane's eigen-decomposition: staged in shared memory, else in the parameter block
    struct Branch { double ed[KP]; };

    __device__ __forceinline__ void init(const LaneSmem &l, const double *gp, int K, const double *staged) {
        sm = l;
        sf = l.f(0); tau = l.f(1); tauf = l.f(2);
        d = staged ? staged : gp + 4 + K;
        A = d + K;
        Ainv = A + K * K;
    }
    __device__ __forceinline__ double pi(int k) const { return sm.f(4 + k); }
    __device__ __forceinline__ Branch branch(double dist, int K) const {
        const double t = (dist + tau) * tauf * sf;
        Branch b;
#pragma unroll
        for (int k = 0; k < KP; ++k) b.ed[k] = (k < K) ? exp(d[k] * t) : 0.;
        return b;
    }
    // out = A (ed o (A^-1 in)): 2 K^2 fused multiply-adds instead of forming P (2 K^3)
    __device__ __forceinline__ void matvec(const Branch &b, const double (&in)[KP], double (&out)[KP], int K) const {
        double w[KP];
        if constexpr (KP > 32) {
            // large state spaces (codon-sized rate matrices): rolled loops -- the unrolled 2 KP^2 multiply-adds of the
            // 48- and 62-state instantiations take minutes to compile; the vectors live in local memory here
#pragma unroll 1
            for (int m = 0; m < K; ++m) {
                double s = 0.;
#pragma unroll 1
                for (int j = 0; j < K; ++j) s = fma(Ainv[m * K + j], in[j], s);
                w[m] = s * b.ed[m];
            }
#pragma unroll 1
            for (int i = 0; i < KP; ++i) {
                double s = 0.;
                if (i < K) {
#pragma unroll 1
                    for (int m = 0; m < K; ++m) s = fma(A[i * K + m], w[m], s);
                }
                out[i] = fmax(s, 0.);
            }
            return;
        }
#pragma unroll
        for (int m = 0; m < KP; ++m) {
            double s = 0.;
            if (m < K) {
#pragma unroll
                for (int j = 0; j < KP; ++j)
                    if (j < K) s = fma(Ainv[m * K + j], in[j], s);
            }
            w[m] = s * b.ed[m];
        }
#pragma unroll
        for (int i = 0; i < KP; ++i) {
            double s = 0.;
            if (i < K) {
#pragma unroll
                for (int m = 0; m < KP; ++m)
                    if (m < K) s = fma(A[i * K + m], w[m], s);
            }
            out[i] = fmax(s, 0.);
        }
    }
    __device__ __forceinline__ void mul_matvec(const Branch &b, const double (&in)[KP], double (&acc)[KP], int K) const {
        double v[KP];
        matvec(b, in, v, K);
#pragma unroll
        for (int k = 0; k < KP; ++k) acc[k] *= v[k];
    }
    __device__ __forceinline__ void maxvec(const Branch &b, const double (&in)[KP], double (&out)[KP],
                                           uint8_t (&arg)[KP], int K) const {
#pragma unroll 1
        for (int i = 0; i < K; ++i) {
            double row[KP];
#pragma unroll
            for (int m = 0; m < KP; ++m) row[m] = (m < K) ? A[i * K + m] * b.ed[m] : 0.;
            double best = -CUDART_INF;
            int bi = 0;
#pragma unroll 1
            for (int j = 0; j < K; ++j) {
                double s = 0.;
#pragma unroll
                for (int m = 0; m < KP; ++m)
                    if (m < K) s = fma(row[m], Ainv[m * K + j], s);
                double inj = 0.;
#pragma unroll
                for (int q = 0; q < KP; ++q) inj = (q == j) ? in[q] : inj;
                const double x = __dmul_rn(s, inj);
                if (x > best) { best = x; bi = j; }
            }
#pragma unroll
            for (int q = 0; q < KP; ++q)
                if (q == i) { out[q] = fmax(best, 0.); arg[q] = (uint8_t) bi; }
        }
#pragma unroll
        for (int q = 0; q < KP; ++q)
            if (q >= K) { out[q] = 0.; arg[q] = 0; }
    }
};

// ------------------------------------------------------------------------------------------ zero clusters

// get_zero_clusters_with_states + alter_zero_node_allowed_states (ml.py:321-387) for one (cluster, column) per
// thread.  Writes the effective masks of all members into the cluster table; when the members share no state the
// union is assigned to all of them and (if record_col[c]) the old masks become their '.initial'.
static __global__ void alter_kernel(int n_clusters, const int *cl_col, const int *cl_ptr, const int *cl_mem,
                             const int *zc_index, const uint64_t *src, int C, uint64_t kmask,
                             const uint8_t *record_col, uint64_t *alt, uint8_t *altflag, uint64_t *initial,
                             uint8_t *hasinit, uint8_t *touched) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_clusters) return;
    const int c = cl_col[q];
    uint64_t common = ~0ULL, uni = 0ULL;
    for (int i = cl_ptr[q]; i < cl_ptr[q + 1]; ++i) {
        const uint64_t v = src[(size_t) cl_mem[i] * C + c] & kmask;
        common &= v;
        uni |= v;
    }
    const bool altered = (common == 0ULL) && record_col[c];   // the reference alters only while tau == 0 (ml.py:101)
    for (int i = cl_ptr[q]; i < cl_ptr[q + 1]; ++i) {
        const int node = cl_mem[i];
        const size_t r = (size_t) zc_index[node] * C + c;
        const uint64_t v = src[(size_t) node * C + c] & kmask;
        alt[r] = altered ? uni : v;
        altflag[r] = altered ? 1 : 0;
        if (altered) {
            initial[r] = v;
            hasinit[r] = 1;
            touched[r] = 1;
        }
    }
}

// ------------------------------------------------------------------------------------------ bottom-up

struct BuArgs {
    TreeDev t;
    ZcDev zc;
    ExtDev ext;
    LaneDev ln;
    const uint64_t *mask;   // [N][C]
    int alter;              // apply the cluster table where the lane's tau == 0
    double *bu;             // [NI][K][J]
    int *bu_ex;             // [NI][J]
    uint8_t *ptr;           // joint mode: [N][C][K] argmax pointers stored on the child (may be null)
    int *err;               // [J][2] first failing (parent, child), initialised to -1
    double *v_all;          // tensor-core sweep, one parameter set: [N][K][C] the products P(t) L_child it forms anyway,
                            // kept for the top-down sweep (they are its node_contribution, ml.py:280); may be null
};

// loads the K-vector of a child: a stored mantissa vector for internal nodes, the allowed-state mask for tips
// (tips: BOTTOM_UP_LIKELIHOOD = allowed states, SF = 0)
template <int KP>
__device__ __forceinline__ int load_child(const TreeDev &t, const ZcDev &zc, const ExtDev &ext, const LaneDev &ln,
                                          const uint64_t *mask, const double *bu, const int *bu_ex, int ch, int j, int c,
                                          bool use_alt, double (&in)[KP], uint64_t &tip_mask) {
    const int cs = t.slot[ch];
    if (cs >= 0) {
        const double *src = bu + ((size_t) cs * ln.K) * ln.J + j;
#pragma unroll
        for (int k = 0; k < KP; ++k) in[k] = (k < ln.K) ? src[(size_t) k * ln.J] : 0.;
        tip_mask = 0ULL;
        return bu_ex[(size_t) cs * ln.J + j];
    }
    if (mask[(size_t) ch * ln.C + c] & kExtFlag) {   // a clade root cut off by subtree sharding
        const int e = ext.index[ch];
        const double *src = ext.bu + ((size_t) e * ln.K) * ln.J + j;
#pragma unroll
        for (int k = 0; k < KP; ++k) in[k] = (k < ln.K) ? src[(size_t) k * ln.J] : 0.;
        tip_mask = ln.kmask;
        return ext.bu_ex[(size_t) e * ln.J + j];
    }
    tip_mask = eff_mask(mask, t, zc, ln.C, ch, c, use_alt, ln.kmask);
#pragma unroll
    for (int k = 0; k < KP; ++k) in[k] = bit_as_double(tip_mask, k);
    return 0;
}

// calc_node_bu_likelihood (ml.py:124-148) for ONE (internal node, lane): the children's vectors in, the node's vector
// and exponent out.  s_par: the block's staged parameters.
template <int KP, template <int> class Lane, bool MAXP>
__device__ __forceinline__ void bu_item(const BuArgs &a, const double *s_par, int nf, bool with_eig, int slot, int j) {
    const int J = a.ln.J, K = a.ln.K, C = a.ln.C;
    const int c = j - (int) a.ln.divC.div((unsigned) j) * C;
    Lane<KP> lane;
    lane.init(LaneSmem{s_par + j, J}, a.ln.params + (size_t) j * a.ln.stride, K,
              with_eig ? s_par + (size_t) nf * J + (size_t) j * eig_lane_stride(K) : nullptr);
    const bool use_alt = a.alter && lane.tau == 0.;
    const int fc = a.t.fc_slot[slot], nc = a.t.nc_slot[slot];
    const int node = a.t.int_nodes[slot];

    double acc[KP];
    {
        const uint64_t m = eff_mask(a.mask, a.t, a.zc, C, node, c, use_alt, a.ln.kmask);
#pragma unroll
        for (int k = 0; k < KP; ++k) acc[k] = bit_as_double(m, k);   // 10 ** log10(allowed), ml.py:127
    }
    int ex = 0;
    if (!MAXP && nc == 2) {
        // binary node: both children's vectors are requested before any arithmetic (twice the bytes in flight)
        double in0[KP], in1[KP];
        uint64_t tm;
        const double d0 = a.t.dist[fc], d1 = a.t.dist[fc + 1];
        ex += load_child<KP>(a.t, a.zc, a.ext, a.ln, a.mask, a.bu, a.bu_ex, fc, j, c, use_alt, in0, tm);
        ex += load_child<KP>(a.t, a.zc, a.ext, a.ln, a.mask, a.bu, a.bu_ex, fc + 1, j, c, use_alt, in1, tm);
        const typename Lane<KP>::Branch b0 = lane.branch(d0, K), b1 = lane.branch(d1, K);
        lane.mul_matvec(b0, in0, acc, K);
        const double vmax0 = vec_max<KP>(acc);
        lane.mul_matvec(b1, in1, acc, K);
        const double vmax = vec_max<KP>(acc);
        if (vmax == 0.) {   // all entries zero: PastMLLikelihoodError of ml.py:139-145
            if (atomicCAS(a.err + 2 * j, -1, node) == -1) a.err[2 * j + 1] = (vmax0 == 0.) ? fc : fc + 1;
        }
        ex += rescale<KP>(acc, vmax, true);
    } else {
        for (int ch = fc; ch < fc + nc; ++ch) {
            const typename Lane<KP>::Branch br = lane.branch(a.t.dist[ch], K);
            double in[KP];
            uint64_t tm;
            ex += load_child<KP>(a.t, a.zc, a.ext, a.ln, a.mask, a.bu, a.bu_ex, ch, j, c, use_alt, in, tm);
            if (MAXP) {
                double v[KP];
                uint8_t arg[KP];
                lane.maxvec(br, in, v, arg, K);
                if (a.ptr) {
                    uint8_t *dst = a.ptr + ((size_t) ch * C + c) * K;
#pragma unroll
                    for (int k = 0; k < KP; ++k)
                        if (k < K) dst[k] = arg[k];
                }
#pragma unroll
                for (int k = 0; k < KP; ++k) acc[k] *= v[k];
            } else {
                lane.mul_matvec(br, in, acc, K);
            }
            const double vmax = vec_max<KP>(acc);
            if (vmax == 0.) {   // all entries zero: PastMLLikelihoodError of ml.py:139-145
                if (atomicCAS(a.err + 2 * j, -1, node) == -1) a.err[2 * j + 1] = ch;
            }
            ex += rescale<KP>(acc, vmax, ch == fc + nc - 1);   // always normalised when the node is complete
        }
    }
    double *dst = a.bu + ((size_t) slot * K) * J + j;
#pragma unroll
    for (int k = 0; k < KP; ++k)
        if (k < K) dst[(size_t) k * J] = acc[k];
    a.bu_ex[(size_t) slot * J + j] = ex;
}

// ... for the internal nodes [slot_begin, slot_begin + n_slots) of one level.
template <int KP, template <int> class Lane, bool MAXP>
__global__ void __launch_bounds__(kBlock, (KP <= 4 ? 6 : KP <= 20 ? PML_BU_MINB : 1)) bu_level_kernel(BuArgs a, int slot_begin, int n_slots) {
    extern __shared__ double s_par[];
    const int J = a.ln.J, K = a.ln.K;
    const int nf = staged_fields(Lane<KP>::family, K);
    const bool with_eig = eig_staged(Lane<KP>::family, K, J);
    pdl_launch_dependents();
    stage_params(s_par, a.ln, nf, with_eig);
    pdl_wait();   // the staging above is independent of the previous level
    const unsigned n_items = (unsigned) n_slots * (unsigned) J;
    for (unsigned g = blockIdx.x * blockDim.x + threadIdx.x; g < n_items; g += gridDim.x * blockDim.x) {
        const unsigned sl = a.ln.divJ.div(g);
        bu_item<KP, Lane, MAXP>(a, s_par, nf, with_eig, slot_begin + (int) sl, (int) (g - sl * (unsigned) J));
    }
}

// ln(sum_i L_root,i pi_i) - SF ln 10 summed over the roots (ml.py:112-121); a root may be a tip.
static __global__ void root_loglik_kernel(TreeDev t, ZcDev zc, LaneDev ln, const uint64_t *mask, int alter,
                                          const double *bu, const int *bu_ex, const int *roots, int n_roots,
                                          int is_marginal, double *out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ln.J) return;
    const int K = ln.K, J = ln.J, c = lane_column(ln, j);
    const double *p = ln.params + (size_t) j * ln.stride;
    const bool use_alt = alter && p[1] == 0.;
    double total = 0.;
    for (int r = 0; r < n_roots; ++r) {
        const int node = roots[r];
        const int s = t.slot[node];
        double acc = is_marginal ? 0. : -CUDART_INF;
        int ex = 0;
        if (s >= 0) ex = bu_ex[(size_t) s * J + j];
        const uint64_t m = (s >= 0) ? 0ULL : eff_mask(mask, t, zc, ln.C, node, c, use_alt, ln.kmask);
        for (int k = 0; k < K; ++k) {
            const double l = (s >= 0) ? bu[bu_at(ln, s, k, j)] : bit_as_double(m, k);
            const double x = l * p[4 + k];
            acc = is_marginal ? acc + x : fmax(acc, x);
        }
        total += log(acc) + (double) ex * 0.69314718055994530942;
    }
    out[j] = total;
}

// ------------------------------------------------------------------------------------------ joint traceback

struct JointArgs {
    TreeDev t;
    ZcDev zc;
    LaneDev ln;
    const uint64_t *base_mask;  // annotation masks (carry the member flags)
    const double *bu;           // joint bottom-up mantissas
    const uint8_t *ptr;         // [N][C][K]
    int *state;                 // [N][C]
    uint64_t *out_mask;         // [N][C]
    const int *root_state;      // [n_roots][C] states the roots take instead of their own arg-max (clade forests of a
                                // tree cut by subtree: the state comes from the part of the tree above the cut), or null
};

// choose_ancestral_states_joint (ml.py:598-622) + unalter_zero_node_joint_states (ml.py:408-428) for the nodes
// [node_begin, node_begin + n_nodes) of one level.
static __global__ void joint_trace_kernel(JointArgs a, int node_begin, int n_nodes) {
    pdl_launch_dependents();
    pdl_wait();
    const int C = a.ln.C, K = a.ln.K;
    const long long n_items = (long long) n_nodes * C;
    for (long long g = (long long) blockIdx.x * blockDim.x + threadIdx.x; g < n_items;
         g += (long long) gridDim.x * blockDim.x) {
        const int node = node_begin + (int) (g / C);
        const int c = (int) (g % C);
        const int par = a.t.parent[node];
        const uint64_t base = a.base_mask[(size_t) node * C + c];
        int s;
        if (par < 0 && a.root_state) {
            // the pointer of this node was followed above the cut; what is left is the node's own un-altering
            s = a.root_state[(size_t) node * C + c];
            if (base & kMemberFlag) {
                const size_t r = (size_t) a.t.zc_index[node] * C + c;
                if (a.zc.altflag[r]) {
                    const uint64_t init = a.zc.initial[r];
                    if (!((init >> s) & 1ULL) && init) s = __ffsll((long long) init) - 1;
                }
            }
        } else if (par < 0) {
            const double *p = a.ln.params + (size_t) c * a.ln.stride;
            const int sl = a.t.slot[node];
            const bool use_alt = p[1] == 0.;
            const uint64_t m = eff_mask(a.base_mask, a.t, a.zc, C, node, c, use_alt, a.ln.kmask);
            double best = -CUDART_INF;
            s = 0;
            for (int k = 0; k < K; ++k) {
                const double l = (sl >= 0) ? a.bu[((size_t) sl * K + k) * C + c] : bit_as_double(m, k);
                const double x = __dmul_rn(l, p[4 + k]);
                if (x > best) { best = x; s = k; }
            }
        } else {
            s = a.ptr[((size_t) node * C + c) * K + a.state[(size_t) par * C + c]];
            if (base & kMemberFlag) {
                const size_t r = (size_t) a.t.zc_index[node] * C + c;
                if (a.zc.altflag[r]) {
                    const uint64_t init = a.zc.initial[r];
                    if (!((init >> s) & 1ULL) && init) s = __ffsll((long long) init) - 1;
                }
            }
        }
        a.state[(size_t) node * C + c] = s;
        a.out_mask[(size_t) node * C + c] = (1ULL << s) | (base & (kMemberFlag | kExtFlag));   // markers stay
    }
}

// ------------------------------------------------------------------------------------------ top-down sweep

struct TdArgs {
    TreeDev t;
    ZcDev zc;
    ExtDev ext;
    LaneDev ln;                 // J == C (one parameter set)
    const uint64_t *mask;       // annotation masks
    const double *bu;
    const int *bu_ex;
    double *td;                 // [NI][K][C]
    int *td_ex;                 // [NI][C]
    const int *joint_state;     // [N][C] or null
    int force_joint;
    double *probs;              // [N][C][K] or null
    double *node_log10lh;       // [N][C] or null
    int *node_ex;               // [N][C] exponent of the raw sums the cooperative kernel leaves in node_log10lh
    uint64_t *map_mask;         // [N][C]
    uint64_t *mppa_mask;        // [N][C]
    unsigned long long *k_hist; // [C][K+1]
    double *zc_lh;              // [rows][C][K]  marginal mantissas of cluster members (for pml_mppa_finalize)
    uint8_t *zc_k;              // [rows][C]     best_k chosen here
    int raw_node_lh = 0;        // the thread-per-item kernels leave (sum, exponent) like the cooperative ones: the sweep
                                // ends with log10_finish_kernel over all nodes (mixed sweeps: fused thin levels)
};

// MPPA state selection (choose_ancestral_states_mppa, ml.py:539-572) on a likelihood vector already multiplied by
// the '.initial' mask where the node has one.
//
// The reference sorts the probabilities, and for k = 1..n scores the array against (0, .., 0, 1/k, .., 1/k) with a
// squared distance, keeping the first minimum; the distance equals sum(p^2) + (1 - 2 T_k) / k with T_k the sum of
// the k entries that receive 1/k (the joint state first when it is forced, then the others by decreasing
// probability).  (1 - 2 T_k) / k is unimodal in k, so states are extracted one at a time by decreasing likelihood
// (ties to the lower index, which is the reference's stable order, ml.py:562-565) until the score has clearly
// passed its minimum.  The selected set is the first best_k extracted states.
template <int KP>
__device__ __forceinline__ uint64_t mppa_select(const double (&lh)[KP], double inv_sum, int K, int joint_index,
                                                bool force_joint, int &best_k) {
    double pj = 0.;
    if (force_joint) {
#pragma unroll
        for (int k = 0; k < KP; ++k) pj = (k == joint_index) ? lh[k] * inv_sum : pj;
    }
    uint64_t used = 0ULL, best_sel = 0ULL;
    double run = 0.;             // sum of the extracted probabilities
    double best_score = CUDART_INF;
    bool joint_seen = false;
    best_k = K;
    for (int k = 1; k <= K; ++k) {
        // next state by decreasing likelihood, first index on ties
        double bv = -1.;
        int bi = 0;
#pragma unroll
        for (int i = 0; i < KP; ++i) {
            const bool free_i = i < K && !((used >> i) & 1ULL);
            if (free_i && lh[i] > bv) { bv = lh[i]; bi = i; }
        }
        const double prev = run;
        run += bv * inv_sum;
        used |= 1ULL << bi;
        joint_seen = joint_seen || (bi == joint_index);
        const double tk = (!force_joint || joint_seen) ? run : pj + prev;
        const double score = (1. - 2. * tk) / (double) k;
        if (score < best_score) { best_score = score; best_k = k; best_sel = used; }
        else if (score > best_score + 1e-9) break;   // past the minimum of a unimodal sequence
    }
    return best_sel;
}

// calc_node_marginal_likelihood + probabilities + MAP + MPPA for one node (ml.py:454-465, 498-500, 577-595, 539-572).
// m[] comes in as L * TD (unscaled marginal without frequencies) and is consumed.
template <int KP, class LaneT>
__device__ __forceinline__ void emit_node(const TdArgs &a, const LaneT &lane, int node, int c, double (&m)[KP],
                                          uint64_t allowed, int ex, unsigned int *sh_hist) {
    const int K = a.ln.K, C = a.ln.C;
    double sum = 0.;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        m[k] = (k < K && ((allowed >> k) & 1ULL)) ? m[k] * lane.pi(k) : 0.;
        sum += m[k];
    }
    const size_t nc = (size_t) node * C + c;
    double inv = 1. / sum;
    if (a.probs) {
        double *dst = a.probs + nc * K;
#pragma unroll
        for (int k = 0; k < KP; ++k)
            if (k < K) dst[k] = m[k] * inv;
    }
    if (a.node_log10lh) {
        if (a.raw_node_lh) {
            a.node_log10lh[nc] = sum;
            a.node_ex[nc] = ex;
        } else {
            a.node_log10lh[nc] = log10(sum) + (double) ex * 0.30102999566398119521;
        }
    }
    const uint64_t base = a.mask[nc];
    const uint64_t flag = base & kMemberFlag;
    if (flag) {
        const size_t r = (size_t) a.t.zc_index[node] * C + c;
        double *z = a.zc_lh + r * K;
#pragma unroll
        for (int k = 0; k < KP; ++k)
            if (k < K) z[k] = m[k];
        if (a.zc.hasinit[r]) {
            const uint64_t init = a.zc.initial[r];
            sum = 0.;
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                m[k] = ((init >> k) & 1ULL) ? m[k] : 0.;   // lh *= initial, ml.py:593-594
                sum += m[k];
            }
            inv = 1. / sum;
        }
    }
    int best = 0;   // first arg max (ml.py:595)
    double bv = m[0];
#pragma unroll
    for (int k = 1; k < KP; ++k)
        if (m[k] > bv) { bv = m[k]; best = k; }
    a.map_mask[nc] = (1ULL << best) | flag | (base & kExtFlag);   // external tips keep their marker
    int best_k;
    uint64_t sel;
    const int ji = a.force_joint ? a.joint_state[nc] : best;
    if (!(sum > 0.)) { sel = a.ln.kmask; best_k = K; }   // NaN probabilities: no k beats +inf, all states kept
    else if (bv * inv >= 0.76 && ji == best) { sel = 1ULL << best; best_k = 1; }   // k = 1 is optimal when p1 > 3/4
    else sel = mppa_select<KP>(m, inv, K, ji, a.force_joint != 0, best_k);
    a.mppa_mask[nc] = sel | flag | (base & kExtFlag);
    if (flag) a.zc_k[(size_t) a.t.zc_index[node] * C + c] = (uint8_t) best_k;
    atomicAdd(sh_hist + c * (K + 1) + best_k, 1u);
}

// calc_node_td_likelihood (ml.py:273-290) fused with the node's marginal step, node-centric: one thread per
// (node of a depth level, column) reads the parent's bottom-up and top-down vectors and the node's own bottom-up
// vector (or tip mask), forms the node's top-down vector, stores it if the node is internal, and emits the node's
// marginal probabilities, MAP and MPPA states.  Roots take the all-ones top-down vector (ml.py:274-277).
// Shared memory: staged parameters, then the block's k histogram.
// One (node, column) of the fused top-down step; sh_hist: the block's k histogram [C][K + 1]
template <int KP, template <int> class Lane>
__device__ __forceinline__ void td_item(const TdArgs &a, const double *s_par, int nf, bool with_eig, unsigned int *sh_hist,
                                        int node, int c) {
    const int C = a.ln.C, K = a.ln.K;
    Lane<KP> lane;
    lane.init(LaneSmem{s_par + c, C}, a.ln.params + (size_t) c * a.ln.stride, K,
              with_eig ? s_par + (size_t) nf * C + (size_t) c * eig_lane_stride(K) : nullptr);
    const bool use_alt = lane.tau == 0.;   // ml.py:702-703
    const int ps = a.t.pslot[node];
    const int cs = a.t.slot[node];

    double Lc[KP], X[KP];
    uint64_t mc;
    const int exc = load_child<KP>(a.t, a.zc, a.ext, a.ln, a.mask, a.bu, a.bu_ex, node, c, c, use_alt, Lc, mc);
    if (cs >= 0) mc = eff_mask(a.mask, a.t, a.zc, C, node, c, use_alt, a.ln.kmask);
    int ex_total;
    if (ps < 0) {
        ex_total = exc;
        if (a.ext.root_td) {   // a clade root: its top-down vector arrives from the part of the tree above the cut
            const double *src = a.ext.root_td + ((size_t) node * K) * C + c;
#pragma unroll
            for (int k = 0; k < KP; ++k) X[k] = (k < K) ? Lc[k] * src[(size_t) k * C] : 0.;
            ex_total += a.ext.root_td_ex[(size_t) node * C + c];
        } else {
#pragma unroll
            for (int k = 0; k < KP; ++k) X[k] = Lc[k];   // TD(root) = 1
        }
    } else {
        // parent part of the node's top-down vector: TD_p * L_p, divided below by the node's own contribution
        {
            const double *src = a.bu + ((size_t) ps * K) * C + c;
#pragma unroll
            for (int k = 0; k < KP; ++k) X[k] = (k < K) ? src[(size_t) k * C] : 0.;
        }
        int pex = a.bu_ex[(size_t) ps * C + c];
        if (ps >= a.t.n_root_slots) {
            const double *src = a.td + ((size_t) ps * K) * C + c;
#pragma unroll
            for (int k = 0; k < KP; ++k) X[k] *= (k < K) ? src[(size_t) k * C] : 0.;
            pex += a.td_ex[(size_t) ps * C + c];
        } else if (a.ext.root_td) {   // child of a clade root: the root's top-down vector came from above the cut
            const int rn = a.t.int_nodes[ps];
            const double *src = a.ext.root_td + ((size_t) rn * K) * C + c;
#pragma unroll
            for (int k = 0; k < KP; ++k) X[k] *= (k < K) ? src[(size_t) k * C] : 0.;
            pex += a.ext.root_td_ex[(size_t) rn * C + c];
        }
        const typename Lane<KP>::Branch br = lane.branch(a.t.dist[node], K);
        double q[KP];
        lane.matvec(br, Lc, q, K);                                   // node_contribution, ml.py:280
#pragma unroll
        for (int k = 0; k < KP; ++k) q[k] = (q[k] > 0.) ? X[k] * __drcp_rn(q[k]) : X[k];   // ml.py:281-283
        int qex = pex - exc + rescale<KP>(q, vec_max<KP>(q), false);                    // ml.py:284-286
        lane.matvec(br, q, X, K);                                    // ml.py:288-289
        if (cs >= 0) {
            qex += rescale<KP>(X, vec_max<KP>(X), true);
            double *dst = a.td + ((size_t) cs * K) * C + c;
#pragma unroll
            for (int k = 0; k < KP; ++k)
                if (k < K) dst[(size_t) k * C] = X[k];
            a.td_ex[(size_t) cs * C + c] = qex;
        } else if (a.mask[(size_t) node * C + c] & kExtFlag) {   // hand the vector to the clade below the cut
            const int e = a.ext.index[node];
            double *dst = a.ext.td + ((size_t) e * K) * C + c;
#pragma unroll
            for (int k = 0; k < KP; ++k)
                if (k < K) dst[(size_t) k * C] = X[k];
            a.ext.td_ex[(size_t) e * C + c] = qex;
        }
#pragma unroll
        for (int k = 0; k < KP; ++k) X[k] *= Lc[k];
        ex_total = qex + exc;
    }
    emit_node<KP>(a, lane, node, c, X, mc, ex_total, sh_hist);
}

template <int KP, template <int> class Lane>
__global__ void __launch_bounds__(kBlock, (KP <= 4 ? 6 : KP <= 20 ? PML_TD_MINB : 1)) td_level_kernel(TdArgs a, int node_begin, int n_nodes) {
    extern __shared__ double s_par[];
    const int C = a.ln.C, K = a.ln.K;
    const int nf = staged_fields(Lane<KP>::family, K);
    const bool with_eig = eig_staged(Lane<KP>::family, K, C);
    unsigned int *sh_hist = reinterpret_cast<unsigned int *>(s_par + staged_doubles(Lane<KP>::family, K, C));   // [C][K+1]
    for (int i = threadIdx.x; i < C * (K + 1); i += blockDim.x) sh_hist[i] = 0u;
    pdl_launch_dependents();
    stage_params(s_par, a.ln, nf, with_eig);
    pdl_wait();   // the staging above is independent of the previous level
    const unsigned n_items = (unsigned) n_nodes * (unsigned) C;
    for (unsigned g = blockIdx.x * blockDim.x + threadIdx.x; g < n_items; g += gridDim.x * blockDim.x) {
        const unsigned nd = a.ln.divC.div(g);
        td_item<KP, Lane>(a, s_par, nf, with_eig, sh_hist, node_begin + (int) nd, (int) (g - nd * (unsigned) C));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * (K + 1); i += blockDim.x)
        if (sh_hist[i]) atomicAdd(a.k_hist + i, (unsigned long long) sh_hist[i]);
}

// Second half of the top-down step on its own: marginal likelihoods, probabilities, MAP and MPPA states of every node
// from the bottom-up vectors and top-down vectors that a vector-only sweep (the tensor-core one, pml_mma.cu) left in
// td_all [N][K][C] / td_ex_all [N][C] (indexed by node, tips included).  No dependency between nodes: one launch for
// the whole forest.
template <int KP, template <int> class Lane>
__global__ void __launch_bounds__(kBlock, (KP <= 20 ? PML_TDSEL_MINB : 1)) td_select_kernel(TdArgs a, const double *td_all,
                                                                                      const int *td_ex_all, int n_nodes) {
    extern __shared__ double s_par[];
    const int C = a.ln.C, K = a.ln.K;
    const int nf = staged_fields(Lane<KP>::family, K);
    unsigned int *sh_hist = reinterpret_cast<unsigned int *>(s_par + (size_t) nf * C);   // [C][K+1]
    for (int i = threadIdx.x; i < C * (K + 1); i += blockDim.x) sh_hist[i] = 0u;
    pdl_launch_dependents();
    stage_params(s_par, a.ln, nf, false);
    pdl_wait();
    const unsigned n_items = (unsigned) n_nodes * (unsigned) C;   // (< 2^31: the vectors alone would not fit otherwise)
    for (unsigned g = blockIdx.x * blockDim.x + threadIdx.x; g < n_items; g += gridDim.x * blockDim.x) {
        const unsigned nd = a.ln.divC.div(g);
        const int node = (int) nd, c = (int) (g - nd * (unsigned) C);
        // the node's top-down vector does not depend on anything looked up: requested first, so that it travels while
        // the slot -> bottom-up vector chain resolves (a root's entry is never written: read, and ignored below)
        double T[KP];
        {
            const double *src = td_all + ((size_t) node * K) * C + c;
#pragma unroll
            for (int k = 0; k < KP; ++k) T[k] = (k < K) ? src[(size_t) k * C] : 0.;
        }
        const int t_ex = td_ex_all[(size_t) node * C + c];
        const int par = a.t.parent[node];
        Lane<KP> lane;
        lane.init(LaneSmem{s_par + c, C}, a.ln.params + (size_t) c * a.ln.stride, K, nullptr);
        const bool use_alt = lane.tau == 0.;   // ml.py:702-703
        const int cs = a.t.slot[node];
        const uint64_t mc = eff_mask(a.mask, a.t, a.zc, C, node, c, use_alt, a.ln.kmask);
        double X[KP];
        int ex_total = 0;
        if (cs >= 0) {
            const double *src = a.bu + ((size_t) cs * K) * C + c;
#pragma unroll
            for (int k = 0; k < KP; ++k) X[k] = (k < K) ? src[(size_t) k * C] : 0.;
            ex_total = a.bu_ex[(size_t) cs * C + c];
        } else {
#pragma unroll
            for (int k = 0; k < KP; ++k) X[k] = bit_as_double(mc, k);
        }
        if (par >= 0) {   // (a root's top-down vector is all ones, ml.py:274-277)
#pragma unroll
            for (int k = 0; k < KP; ++k) X[k] *= T[k];
            ex_total += t_ex;
        }
        emit_node<KP>(a, lane, node, c, X, mc, ex_total, sh_hist);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * (K + 1); i += blockDim.x)
        if (sh_hist[i]) atomicAdd(a.k_hist + i, (unsigned long long) sh_hist[i]);
}

// pml_mppa_finalize: redo the MPPA choice of cluster members whose '.initial' was replaced by the MAP-restricted
// pass' alteration (ml.py:384-385, 541-542).  One thread per (table row, column).
template <int KP>
__global__ void mppa_fix_kernel(int n_rows, const int *row_node, int C, int K, uint64_t kmask, ZcDev zc,
                                const uint64_t *base_mask, const uint64_t *prev_initial, const uint8_t *prev_hasinit,
                                const double *zc_lh, uint8_t *zc_k, const int *joint_state, int force_joint,
                                uint64_t *mppa_mask, unsigned long long *k_hist) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_rows * C) return;
    const int row = g / C, c = g % C;
    const size_t r = (size_t) row * C + c;
    const int node = row_node[row];
    const size_t nc = (size_t) node * C + c;
    if (!(base_mask[nc] & kMemberFlag) || !zc.touched[r]) return;   // '.initial' not rewritten since MAP used it
    double m[KP];
    const uint64_t i0 = prev_hasinit[r] ? prev_initial[r] : kmask;    // applied in place by MAP (ml.py:593-594)
    const uint64_t i1 = zc.initial[r];                                // applied again by MPPA (ml.py:541-542)
    double msum = 0.;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        m[k] = (k < K && ((i0 >> k) & (i1 >> k) & 1ULL)) ? zc_lh[r * K + k] : 0.;
        msum += m[k];
    }
    int best_k;
    uint64_t sel;
    if (msum > 0.) sel = mppa_select<KP>(m, 1. / msum, K, force_joint ? joint_state[nc] : 0, force_joint != 0, best_k);
    else { sel = kmask; best_k = K; }
    const int old_k = zc_k[r];
    mppa_mask[nc] = sel | kMemberFlag;
    zc_k[r] = (uint8_t) best_k;
    if (old_k != best_k) {
        atomicAdd(k_hist + c * (K + 1) + old_k, (unsigned long long) -1LL);
        atomicAdd(k_hist + c * (K + 1) + best_k, 1ULL);
    }
}

// ------------------------------------------------------------------------------------------ parsimony

// get_most_common_states (parsimony.py:76-89) over `first` plus the masks m[0..n): states with the maximal count.
__device__ __forceinline__ uint64_t most_common(uint64_t first, bool has_first, const uint64_t *m, size_t stride,
                                                int n, int skip, uint64_t kmask) {
    const int total = n + (has_first ? 1 : 0) - (skip >= 0 ? 1 : 0);
    if (total <= 6) {
        // bit-sliced 3-bit counters: c2 c1 c0 hold, per state, how many sets contain it
        uint64_t c0 = 0, c1 = 0, c2 = 0;
        auto add = [&](uint64_t x) {
            uint64_t carry = c0 & x;
            c0 ^= x;
            const uint64_t carry2 = c1 & carry;
            c1 ^= carry;
            c2 ^= carry2;
        };
        if (has_first) add(first & kmask);
        for (int i = 0; i < n; ++i)
            if (i != skip) add(m[(size_t) i * stride] & kmask);
        uint64_t cand = kmask;
        if (cand & c2) cand &= c2;
        if (cand & c1) cand &= c1;
        if (cand & c0) cand &= c0;
        return (c0 | c1 | c2) ? cand : 0ULL;
    }
    int best = 0;
    uint64_t out = 0ULL;
    for (int s = 0; s < 64 && ((kmask >> s) & 1ULL); ++s) {
        int cnt = has_first ? (int) ((first >> s) & 1ULL) : 0;
        for (int i = 0; i < n; ++i)
            if (i != skip) cnt += (int) ((m[(size_t) i * stride] >> s) & 1ULL);
        if (cnt > best) { best = cnt; out = 1ULL << s; }
        else if (cnt == best && cnt > 0) out |= 1ULL << s;
    }
    return out;
}

// uppass (parsimony.py:92-122) for the internal nodes of one level; down[] was initialised to the annotation sets.
static __global__ void pars_uppass_kernel(TreeDev t, int C, uint64_t kmask, uint64_t *down, int slot_begin, int n_slots) {
    pdl_launch_dependents();
    pdl_wait();
    const long long n_items = (long long) n_slots * C;
    for (long long g = (long long) blockIdx.x * blockDim.x + threadIdx.x; g < n_items;
         g += (long long) gridDim.x * blockDim.x) {
        const int node = t.int_nodes[slot_begin + (int) (g / C)], c = (int) (g % C);
        const uint64_t mc = most_common(0, false, down + (size_t) t.first_child[node] * C + c, C, t.n_children[node],
                                        -1, kmask);
        const uint64_t own = down[(size_t) node * C + c];
        const uint64_t inter = own & mc;
        if (inter) down[(size_t) node * C + c] = inter;
    }
}

// acctran (parsimony.py:125-158): the children of the internal nodes of one level.
static __global__ void pars_acctran_kernel(TreeDev t, int C, const uint64_t *down, uint64_t *out, int slot_begin,
                                    int n_slots) {
    pdl_launch_dependents();
    pdl_wait();
    const long long n_items = (long long) n_slots * C;
    for (long long g = (long long) blockIdx.x * blockDim.x + threadIdx.x; g < n_items;
         g += (long long) gridDim.x * blockDim.x) {
        const int node = t.int_nodes[slot_begin + (int) (g / C)], c = (int) (g % C);
        const uint64_t own = out[(size_t) node * C + c];
        const int fc = t.first_child[node];
        for (int ch = fc; ch < fc + t.n_children[node]; ++ch) {
            const uint64_t d = down[(size_t) ch * C + c];
            const uint64_t inter = own & d;
            out[(size_t) ch * C + c] = inter ? inter : d;
        }
    }
}

// downpass (parsimony.py:161-210): an internal node's own state from up(node) and its children's down sets, then
// up() of its children; tip children are finished on the spot.  res[] starts as the pre-uppass sets.
static __global__ void pars_downpass_kernel(TreeDev t, int C, uint64_t kmask, const uint64_t *down, uint64_t *up,
                                     uint64_t *res, int slot_begin, int n_slots) {
    pdl_launch_dependents();
    pdl_wait();
    const long long n_items = (long long) n_slots * C;
    for (long long g = (long long) blockIdx.x * blockDim.x + threadIdx.x; g < n_items;
         g += (long long) gridDim.x * blockDim.x) {
        const int node = t.int_nodes[slot_begin + (int) (g / C)], c = (int) (g % C);
        const size_t nc = (size_t) node * C + c;
        const uint64_t upn = (t.parent[node] < 0) ? kmask : up[nc];
        const int fc = t.first_child[node], n = t.n_children[node];
        const uint64_t *kids = down + (size_t) fc * C + c;
        const uint64_t du = most_common(upn, true, kids, C, n, -1, kmask);
        const uint64_t pre = res[nc];
        res[nc] = (du & pre) ? (du & pre) : pre;
        for (int i = 0; i < n; ++i) {
            const int ch = fc + i;
            const uint64_t upc = most_common(upn, true, kids, C, n, i, kmask);
            const size_t cc = (size_t) ch * C + c;
            if (t.slot[ch] >= 0) up[cc] = upc;
            else {
                const uint64_t prec = res[cc];
                res[cc] = (upc & prec) ? (upc & prec) : prec;
            }
        }
    }
}

// deltran (parsimony.py:213-242): children of the internal nodes of one level, on top of the DOWNPASS result.
static __global__ void pars_deltran_kernel(TreeDev t, int C, uint64_t *res, int slot_begin, int n_slots) {
    pdl_launch_dependents();
    pdl_wait();
    const long long n_items = (long long) n_slots * C;
    for (long long g = (long long) blockIdx.x * blockDim.x + threadIdx.x; g < n_items;
         g += (long long) gridDim.x * blockDim.x) {
        const int node = t.int_nodes[slot_begin + (int) (g / C)], c = (int) (g % C);
        const uint64_t own = res[(size_t) node * C + c];
        const int fc = t.first_child[node];
        for (int ch = fc; ch < fc + t.n_children[node]; ++ch) {
            const uint64_t inter = res[(size_t) ch * C + c] & own;
            if (inter) res[(size_t) ch * C + c] = inter;
        }
    }
}

// get_num_parsimonious_steps (parsimony.py:360-380): min-plus DP, one thread per (internal node, column);
// num[slot][k][C] int32 (INF for states outside the node's set).
static __global__ void pars_steps_kernel(TreeDev t, int C, int K, const uint64_t *states, int *num, int slot_begin,
                                  int n_slots) {
    pdl_launch_dependents();
    pdl_wait();
    const int INF = 1 << 29;
    const long long n_items = (long long) n_slots * C;
    for (long long g = (long long) blockIdx.x * blockDim.x + threadIdx.x; g < n_items;
         g += (long long) gridDim.x * blockDim.x) {
        const int slot = slot_begin + (int) (g / C);
        const int node = t.int_nodes[slot], c = (int) (g % C);
        const uint64_t own = states[(size_t) node * C + c];
        const int fc = t.first_child[node], n = t.n_children[node];
        int *dst = num + ((size_t) slot * K) * C + c;
        for (int s = 0; s < K; ++s) dst[(size_t) s * C] = ((own >> s) & 1ULL) ? 0 : INF;
        for (int ch = fc; ch < fc + n; ++ch) {
            const int cs = t.slot[ch];
            const uint64_t cm = states[(size_t) ch * C + c];
            const int *src = (cs >= 0) ? num + ((size_t) cs * K) * C + c : nullptr;
            int mn = INF;
            for (int s = 0; s < K; ++s) {
                const int v = src ? src[(size_t) s * C] : (((cm >> s) & 1ULL) ? 0 : INF);
                mn = min(mn, v);
            }
            for (int s = 0; s < K; ++s)
                if ((own >> s) & 1ULL) {
                    const int v = src ? src[(size_t) s * C] : (((cm >> s) & 1ULL) ? 0 : INF);
                    dst[(size_t) s * C] += min(v, mn + 1);
                }
        }
    }
}

static __global__ void pars_steps_root_kernel(TreeDev t, int C, int K, const uint64_t *states, const int *num,
                                       const int *roots, int n_roots, long long *out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    long long total = 0;
    for (int r = 0; r < n_roots; ++r) {
        const int node = roots[r], s0 = t.slot[node];
        if (s0 < 0) continue;   // a single-node tree needs no change
        int mn = 1 << 29;
        for (int s = 0; s < K; ++s) mn = min(mn, num[((size_t) s0 * K + s) * C + c]);
        total += mn;
    }
    out[c] = total;
}

// ------------------------------------------------------------------------------------------ utilities

// [C][N] host layout -> [N][C] device layout, adding the allowed-states rule "no annotation => all states"
// (initialize_allowed_states, ml.py:311-313) and the cluster-member flag.
static __global__ void init_masks_kernel(int N, int C, uint64_t kmask, const uint64_t *annot_cn, const uint8_t *member_cn,
                                  uint64_t *base_nc, int apply_default) {
    const long long g = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long) N * C) return;
    const int node = (int) (g / C), c = (int) (g % C);
    uint64_t m = annot_cn[(size_t) c * N + node] & kmask;
    if (apply_default && !m) m = kmask;
    if (member_cn && member_cn[(size_t) c * N + node]) m |= kMemberFlag;
    base_nc[g] = m;
}

// node_log10lh holds sum of the node's marginal mantissas, node_ex its power-of-two exponent: take the logarithm
// (all lanes busy, instead of one rank per group inside the sweep)
// (HBM-bound: 20 bytes per entry; two entries per thread through 16- and 8-byte accesses, persistent grid)
static __global__ void __launch_bounds__(256) log10_finish_kernel(long long n, double *v, const int *ex) {
    pdl_launch_dependents();
    pdl_wait();
    const long long pairs = n >> 1;
    double2 *v2 = reinterpret_cast<double2 *>(v);
    const int2 *e2 = reinterpret_cast<const int2 *>(ex);
    for (long long g = (long long) blockIdx.x * blockDim.x + threadIdx.x; g < pairs; g += (long long) gridDim.x * blockDim.x) {
        double2 x = v2[g];
        const int2 e = e2[g];
        x.x = log10(x.x) + (double) e.x * 0.30102999566398119521;
        x.y = log10(x.y) + (double) e.y * 0.30102999566398119521;
        v2[g] = x;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) v[n - 1] = log10(v[n - 1]) + (double) ex[n - 1] * 0.30102999566398119521;
}

// marks the listed tips as external (their vectors live in the ExtDev tables) in every column
static __global__ void mark_ext_kernel(int n, const int *nodes, int C, uint64_t *mask, int *ext_index) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n * C) return;
    const int i = g / C, c = g % C;
    mask[(size_t) nodes[i] * C + c] |= kExtFlag;
    if (c == 0) ext_index[nodes[i]] = i;
}

// gathers the bottom-up vectors of the roots: [T][K][J] mantissas and [T][J] exponents (tip roots: their masks)
static __global__ void root_partials_kernel(TreeDev t, ZcDev zc, LaneDev ln, const uint64_t *mask, int alter,
                                            const double *bu, const int *bu_ex, int n_roots, double *out,
                                            int *out_ex) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_roots * ln.J) return;
    const int node = g / ln.J, j = g % ln.J, c = j % ln.C;
    const int s = t.slot[node];
    const double *p = ln.params + (size_t) j * ln.stride;
    const uint64_t m = (s >= 0) ? 0ULL : eff_mask(mask, t, zc, ln.C, node, c, alter && p[1] == 0., ln.kmask);
    for (int k = 0; k < ln.K; ++k)
        out[((size_t) node * ln.K + k) * ln.J + j] = (s >= 0) ? bu[bu_at(ln, s, k, j)] : bit_as_double(m, k);
    out_ex[(size_t) node * ln.J + j] = (s >= 0) ? bu_ex[(size_t) s * ln.J + j] : 0;
}

static __global__ void strip_flag_kernel(long long n, uint64_t kmask, const uint64_t *in, uint64_t *out) {
    const long long g = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n) out[g] = in[g] & kmask;
}

static __global__ void merge_flag_kernel(long long n, const uint64_t *base, const uint64_t *src, uint64_t *dst) {
    const long long g = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n) dst[g] = src[g] | (base[g] & kMemberFlag);
}

static __global__ void fill_int_kernel(long long n, int value, int *out) {
    const long long g = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n) out[g] = value;
}

static __global__ void flush_kernel(long long n, float *buf, float v) {
    for (long long g = (long long) blockIdx.x * blockDim.x + threadIdx.x; g < n;
         g += (long long) gridDim.x * blockDim.x)
        buf[g] = v;
}

}  // namespace pml
