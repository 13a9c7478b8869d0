// Single-launch ("dataflow") sweeps for the F81 family (F81 / JC / EFT): the headline path.
//
// A level-synchronous sweep is ~55 dependent launches, and on a 1M-tip tree 30 of them are narrower than one wave of
// the GPU: each pays a launch + drain (~3-5 us) for microseconds of work, and every fat level pays a ramp and a tail.
// Here a sweep is ONE persistent kernel.  The host lays the nodes out in a dependency-respecting order (the
// "schedule", pml_flow.cu: level order, cut into batches that never hold a node together with its child); warps take
// chunks of consecutive batches from a ticket counter and, before the bulk copies of a batch are issued, check one
// ready flag per operand vector that was produced by this same sweep (bottom-up: the children, top-down: the parent).
// A flag is the sweep's epoch number, written with release semantics after the warp has stored the vector; the
// consumer reads it with acquire semantics and only then lets the async proxy (cp.async.bulk) fetch the vector.
//
// Progress: tickets are handed out in schedule order and a warp works through its batches in that order, blocking
// only when it holds no unpublished work; the lowest unfinished batch always belongs to a warp that is running and
// all its operands are finished, so the sweep cannot deadlock whatever the number of resident blocks.  In the fat part
// of a tree the operands were produced long before (a full level earlier) and nothing ever waits; in the thin parts
// (the first and last dozen levels) the wait is the true dependency latency instead of a kernel boundary.
//
// Vectors of the flow path carry their power-of-two exponent as one more row: a (slot, tile) run is
// [K + 1][JT] doubles, row K = the exponents (int32 in the low word), so one bulk copy brings both (the level
// kernels keep separate exponent arrays and their own buffers; the engine never mixes the two in one pass).
#pragma once
#include "pml_tpi.cuh"

namespace pml {

// ------------------------------------------------------------------------------------------ control block
//
// ctl (unsigned words): [0] error (a spin ran out), [1] epoch of the running sweep, [16 + 16 t] chunk ticket of
// tile t (64-byte spacing).  flow_begin_kernel bumps the epoch and clears the tickets; flags are never cleared.
constexpr int kFlowCtlTile0 = 16, kFlowCtlStride = 16;
__host__ __device__ constexpr int flow_ctl_words(int tiles) { return kFlowCtlTile0 + kFlowCtlStride * tiles; }

static __global__ void flow_begin_kernel(unsigned *ctl, int tiles) {
    const int t = threadIdx.x + blockIdx.x * blockDim.x;
    if (t < tiles) ctl[kFlowCtlTile0 + kFlowCtlStride * t] = 0u;
    if (t == 0) ctl[1] += 1u;
}

struct FlowSeg {   // a run of batches handed out in chunks of `chunk` batches
    int first_chunk, first_batch, n_batches, chunk;
};
constexpr int kFlowMaxSeg = 192;

struct FlowSched {
    const int4 *desc;      // per scheduled node (padding entries have slot -1), layout depends on the sweep
    const int4 *seg;       // [n_seg] FlowSeg
    int n_seg, n_batches, n_chunks;
};

// Flags are polled with RELAXED loads (they go to L2: no L1 invalidation per poll, which an acquire load costs).  What
// the poll orders is the bulk copy that follows it: the copy is issued only after the loaded value has been compared
// (a control dependency), is preceded by a proxy fence, and reads L2 -- where the producer's release put the vector
// before the flag.  The few generic loads of produced vectors (nodes that are not binary) use ld.global.cg.
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// orders this thread's generic-proxy view (the acquire above) before the async-proxy reads it issues next
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// blocks until the flag carries this sweep's epoch; a spin that runs out (seconds) is a scheduling bug: trap rather
// than hang the device
__device__ __forceinline__ void flag_wait(const unsigned *flag, unsigned epoch, unsigned *ctl) {
    unsigned spins = 0;
    while (ld_relaxed_u32(flag) != epoch) {
        __nanosleep(32);
        if (++spins > (1u << 25)) {
            ctl[0] = 1u;
            __threadfence_system();
            __trap();
        }
    }
}

__device__ __forceinline__ void st_relaxed_u32(unsigned *p, unsigned v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// The stream of batches of one warp: chunks of consecutive batches taken from the tile's ticket counter.  The next
// ticket is requested (lane 0's atomic) when a chunk is opened and only consumed when the chunk is used up.
struct FlowTickets {
    const int4 *seg;       // shared-memory copy of the segment table
    int n_seg, n_chunks;
    unsigned *counter;
    int cur, end;          // batches left in the open chunk: [cur, end)
    unsigned pending;      // lane 0: the requested ticket
    int s;                 // segment cursor (tickets only grow)
    bool last;             // the batch just returned closes its chunk
    __device__ __forceinline__ void init(const int4 *seg_, int n_seg_, int n_chunks_, unsigned *counter_) {
        seg = seg_; n_seg = n_seg_; n_chunks = n_chunks_; counter = counter_;
        cur = end = 0; s = 0; last = true;
        request();
    }
    __device__ __forceinline__ void request() {
        if ((threadIdx.x & 31) == 0) pending = atomicAdd(counter, 1u);
    }
    // next batch index, or -1 when the schedule is exhausted
    __device__ __forceinline__ int next() {
        if (cur >= end) {
            if (end < 0) return -1;
            const unsigned c = __shfl_sync(0xffffffffu, pending, 0);
            if (c >= (unsigned) n_chunks) { end = -1; last = true; return -1; }
            while (s + 1 < n_seg && (int) c >= seg[s + 1].x) ++s;
            const int4 g = seg[s];
            cur = g.y + ((int) c - g.x) * g.w;
            end = min(cur + g.w, g.y + g.z);
            request();
        }
        last = cur + 1 >= end;
        return cur++;
    }
};

// Descriptors of the coming batches travel through a small per-warp ring in shared memory, fetched by bulk copies two
// batches ahead: prefetching them into registers made the compiler spill, and a spill store waits for the very load it
// spills.  Slot = sequence number of the batch mod kFlowRing; one mbarrier per slot, armed by lane 0.
constexpr int kFlowRing = 4;
struct FlowDescRing {
    unsigned char *base;   // kFlowRing slots of `bytes` bytes
    uint32_t bar0;         // kFlowRing mbarriers, 8 bytes apart
    int bytes;             // descriptor bytes of one batch
    const unsigned char *src;
    __device__ __forceinline__ void init(unsigned char *base_, uint32_t bar0_, int bytes_, const void *src_) {
        base = base_; bar0 = bar0_; bytes = bytes_; src = static_cast<const unsigned char *>(src_);
        if ((threadIdx.x & 31) == 0)
            for (int i = 0; i < kFlowRing; ++i) mbar_init(bar0 + 8 * i, 1);
    }
    __device__ __forceinline__ void fetch(int seq, int batch) {   // whole warp calls, lane 0 acts
        if (batch < 0 || (threadIdx.x & 31) != 0) return;
        const int sl = seq & (kFlowRing - 1);
        mbar_arrive_expect(bar0 + 8 * sl, bytes);
        bulk_g2s(smem_addr(base + sl * bytes), src + (size_t) batch * bytes, bytes, bar0 + 8 * sl);
    }
    __device__ __forceinline__ const int4 *get(int seq) const {   // blocks until the slot's copy has landed
        const int sl = seq & (kFlowRing - 1);
        mbar_wait(bar0 + 8 * sl, (unsigned) (seq / kFlowRing) & 1u);
        return reinterpret_cast<const int4 *>(base + sl * bytes);
    }
};

// Publication.  A finished vector is announced by storing the sweep's epoch into its flag AFTER a gpu-scope fence -- and
// that fence costs a round trip through the memory system (~1 us under load).  In the fat part of the tree nobody waits
// for a vector until a whole level later, so a warp collects the slots it has finished (in shared memory, per issuer
// thread) and announces them with ONE fence when the chunk of batches it took from the ticket counter is done, at a
// point of the next step where the stores are old.  In the thin parts (the next batch is not ready, or no next batch)
// it announces at once: the consumers may be waiting, and nothing may stay unannounced while this warp waits.
constexpr int kFlowPend = 8;   // batches per chunk at most (chunk_for in pml_flow.cu)
struct FlowPending {
    int *list;     // this issuer thread's kFlowPend entries
    int n;
    __device__ __forceinline__ void init(int *list_) { list = list_; n = 0; }
    __device__ __forceinline__ void add(int slot) {
        if (n == kFlowPend) return;   // cannot happen: flushed at every chunk end
        list[n++] = slot;
    }
    // every lane of the warp calls this at the same point; the warp barrier that ordered all lanes' stores before it
    // is the caller's
    __device__ __forceinline__ void flush(unsigned *flags, size_t stride, unsigned epoch) {
        if (n > 0) {
            fence_acq_rel_gpu();
            for (int i = 0; i < n; ++i) st_relaxed_u32(flags + (size_t) list[i] * stride, epoch);
            n = 0;
        }
    }
};

// ------------------------------------------------------------------------------------------ bottom-up sweep

struct FlowBuArgs {
    TreeDev t;
    ZcDev zc;
    LaneDev ln;             // J = all lanes of the pass (T tiles of JT); params [j_real][stride]
    int j_real;             // lanes that carry parameters; the padding lanes behind them repeat the last one
    const uint64_t *mask;   // [N][C]
    int alter;
    double *vec;            // [NI][T][K + 1][JT] bottom-up vectors with their exponent row
    int *err;               // [J][2]
    FlowSched sched;        // desc: 3 x int4 per scheduled node: {first child, #children, slot child 0, slot child 1},
                            //       {dist child 0, dist child 1}, {node, own slot, -, -}
    unsigned *ctl;
    unsigned *flags;        // [NI][T]
};

#ifndef PML_FLOW_BLOCK
#define PML_FLOW_BLOCK 128
#endif
#ifndef PML_FLOW_BU_MINB
#define PML_FLOW_BU_MINB 4
#endif

__host__ __device__ constexpr int flow_bu_stage_bytes(int K, int JT) { return 2 * (K + 1) * JT * 8; }   // per node
// per warp: staging of the batch's child runs, descriptor ring, pending list, mbarriers (stage + ring)
__host__ __device__ constexpr int flow_bu_warp_bytes(int K, int JT) {
    return ((32 / JT) * flow_bu_stage_bytes(K, JT) + kFlowRing * (32 / JT) * 48 + (32 / JT) * kFlowPend * 4 + 8 * (1 + kFlowRing) + 127) & ~127;
}
__host__ __device__ inline size_t flow_bu_smem(int K, int JT, int block) {
    return (size_t) round16((4 + K) * JT * 8) + (size_t) kFlowMaxSeg * 16 + (size_t) (block / 32) * flow_bu_warp_bytes(K, JT);
}

// a tip's vector is its allowed-state mask: written into the (unused) staging column so that the walks are branch-free
template <int KC, int JT>
__device__ __forceinline__ void flow_expand_mask(double *col, uint64_t m, int K) {
    if (KC) {
#pragma unroll
        for (int k = 0; k < (KC ? KC : 1); ++k) col[k * JT] = __hiloint2double(((m >> k) & 1ULL) ? 0x3ff00000 : 0, 0);
    } else {
        for (int k = 0; k < K; ++k, m >>= 1) col[k * JT] = __hiloint2double((m & 1ULL) ? 0x3ff00000 : 0, 0);
    }
}

// calc_node_bu_likelihood (ml.py:124-148), marginal (sum-product) mode, for every internal node of the forest in one
// launch.  One thread per (node, lane) item streaming the staged child columns from shared memory (see pml_tpi.cuh
// for the arithmetic); KC > 0: the number of states is the compile-time constant KC (walks fully unrolled, offsets
// folded into the instructions).
//
// Steady state of a warp (fat part of the tree: the next batch's operands are always finished):
//   top      the descriptor copy of the batch after next goes out; relaxed loads of the NEXT batch's operand flags
//            (looked at after the arithmetic) and of its mask words
//   wait     this batch's copies (issued at the end of the previous step)
//   walk 1   frequency-weighted sums
//   publish  when the previous batch closed a chunk: its chunk's flags, one fence (the stores are a walk old)
//   walk 2   products, stored straight to memory
//   reissue  the stage is free: the next batch's copies go out (they fly during the next step's preamble)
// When the next batch is not ready (thin parts of the tree) the warp publishes at once and blocks on the flags.
template <int KC, int JT>
__global__ void __launch_bounds__(PML_FLOW_BLOCK, PML_FLOW_BU_MINB) bu_flow_kernel(FlowBuArgs a) {
    extern __shared__ double s_par[];
    constexpr int NPW = 32 / JT;
    const int K = KC ? KC : a.ln.K, C = a.ln.C;
    const int KS = K + 1;
    const int T = gridDim.y, tile = blockIdx.y, j0 = tile * JT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / JT, jl = lane - sub * JT;
    const int j = j0 + jl, c = lane_column(a.ln, min(j, a.j_real - 1));
    const int VB = KS * JT * 8;   // bytes of one (slot, tile) run
    unsigned char *seg_base = reinterpret_cast<unsigned char *>(s_par) + round16((4 + K) * JT * 8);
    int4 *s_seg = reinterpret_cast<int4 *>(seg_base);
    unsigned char *wbase = seg_base + kFlowMaxSeg * 16 + (size_t) warp * flow_bu_warp_bytes(K, JT);
    const int node_bytes = 2 * VB;
    unsigned char *mine = wbase + (size_t) sub * node_bytes;
    unsigned char *ring_base = wbase + (size_t) NPW * node_bytes;
    int *pend_base = reinterpret_cast<int *>(ring_base + kFlowRing * NPW * 48);
    const uint32_t bar = smem_addr(reinterpret_cast<unsigned char *>(pend_base) + NPW * kFlowPend * 4);
    const uint32_t stage = smem_addr(mine);
    FlowDescRing ring;
    ring.init(ring_base, bar + 8, NPW * 48, a.sched.desc);
    if (lane == 0) {
        mbar_init(bar, NPW);
        mbar_fence_init();
    }
    {
        const int total = (4 + K) * JT;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int f = i / JT, jj = i - f * JT;
            s_par[i] = a.ln.params[(size_t) min(j0 + jj, a.j_real - 1) * a.ln.stride + f];
        }
        for (int i = threadIdx.x; i < a.sched.n_seg; i += blockDim.x) s_seg[i] = a.sched.seg[i];
        __syncthreads();
    }
    const unsigned epoch = ld_volatile_u32(a.ctl + 1);
    const double sf = s_par[jl], tau = s_par[JT + jl], tauf = s_par[2 * JT + jl], mu = s_par[3 * JT + jl];
    const double *spi = s_par + 4 * JT + jl;   // pi_k at spi[k * JT]
    const bool use_alt = a.alter && tau == 0.;
    const bool mu_inf = isinf(mu);
    const size_t run = (size_t) KS * JT;       // doubles of one (slot, tile) run
    unsigned *flags = a.flags + tile;          // flag of slot s at flags[s * T]
    double *in0 = reinterpret_cast<double *>(mine) + jl;   // child 0, state k at in0[k * JT]; exponent at row K
    double *in1 = in0 + run;                               // child 1
    FlowPending pending;
    pending.init(pend_base + sub * kFlowPend);

    struct Desc { int4 top, len; int node, slot; };
    auto read_desc = [&](int seq, int b, Desc &d) {
        if (b < 0) { d.top = make_int4(0, 0, -1, -1); d.slot = -1; d.node = 0; d.len = make_int4(0, 0, 0, 0); return; }
        const int4 *p = ring.get(seq) + 3 * sub;
        d.top = p[0];
        d.len = p[1];
        const int4 w = p[2];
        d.node = w.x;
        d.slot = w.y;
    };
    // blocks until the operands of this thread's node are finished (issuer threads; non-binary nodes: every child)
    auto wait_ready = [&](const int4 &top, int slot) {
        if (jl != 0 || slot < 0) return;
        if (top.y == 2) {
            if (top.z >= 0) flag_wait(flags + (size_t) top.z * T, epoch, a.ctl);
            if (top.w >= 0) flag_wait(flags + (size_t) top.w * T, epoch, a.ctl);
            return;
        }
        for (int q = 0; q < top.y; ++q) {
            const int cs = a.t.slot[top.x + q];
            if (cs >= 0) flag_wait(flags + (size_t) cs * T, epoch, a.ctl);
        }
    };
    // one thread per node arms the barrier and issues the copies of the node's two child runs
    auto issue = [&](const int4 &top, int slot) {
        if (jl != 0) return;
        const bool bin = slot >= 0 && top.y == 2;
        const int cs0 = top.z, cs1 = top.w;
        mbar_arrive_expect(bar, bin ? ((cs0 >= 0 ? VB : 0) + (cs1 >= 0 ? VB : 0)) : 0);
        if (!bin) return;
        fence_proxy_async_global();
        if (cs0 >= 0) bulk_g2s(stage, a.vec + ((size_t) cs0 * T + tile) * run, VB, bar);
        if (cs1 >= 0) bulk_g2s(stage + VB, a.vec + ((size_t) cs1 * T + tile) * run, VB, bar);
    };

    FlowTickets tk;
    tk.init(s_seg, a.sched.n_seg, a.sched.n_chunks, a.ctl + kFlowCtlTile0 + kFlowCtlStride * tile);
    __syncwarp();
    int b_cu = tk.next();
    bool last_cu = tk.last;
    int b_nx = tk.next();
    bool last_nx = tk.last;
    ring.fetch(0, b_cu);
    ring.fetch(1, b_nx);
    Desc cu;
    uint64_t cu_mo = 0, cu_m0 = 0, cu_m1 = 0;
    uint32_t phase = 0;
    bool flush_due = false;   // the previous batch closed its chunk in steady state: announce after walk 1
    int seq = 0;
    if (b_cu >= 0) {
        read_desc(0, b_cu, cu);
        wait_ready(cu.top, cu.slot);
        issue(cu.top, cu.slot);
        if (cu.slot >= 0) {
            cu_mo = a.mask[(size_t) cu.node * C + c];
            if (cu.top.y == 2 && cu.top.z < 0) cu_m0 = a.mask[(size_t) cu.top.x * C + c];
            if (cu.top.y == 2 && cu.top.w < 0) cu_m1 = a.mask[(size_t) (cu.top.x + 1) * C + c];
        }
    }
    while (b_cu >= 0) {
        const int b_n2 = tk.next();
        const bool last_n2 = tk.last;
        ring.fetch(seq + 2, b_n2);
        const bool on = cu.slot >= 0;
        const int slot = cu.slot, node = cu.node;
        const int4 d_top = cu.top, d_len = cu.len;
        const int fc = d_top.x, nc = d_top.y;
        const bool bin = nc == 2;
        // the next batch: its descriptor is in the ring by now; flags and mask words are requested here and looked at
        // after the arithmetic
        int4 n_top = make_int4(0, 0, -1, -1);
        int n_slot = -1;
        unsigned nf0 = epoch, nf1 = epoch;
        uint64_t n_mo = 0, n_m0 = 0, n_m1 = 0;
        if (b_nx >= 0) {
            const int4 *p = ring.get(seq + 1) + 3 * sub;
            n_top = p[0];
            const int4 w = p[2];
            n_slot = w.y;
            if (n_slot >= 0) {
                if (jl == 0 && n_top.y == 2) {
                    if (n_top.z >= 0) nf0 = ld_relaxed_u32(flags + (size_t) n_top.z * T);
                    if (n_top.w >= 0) nf1 = ld_relaxed_u32(flags + (size_t) n_top.w * T);
                }
                n_mo = a.mask[(size_t) w.x * C + c];
                if (n_top.y == 2 && n_top.z < 0) n_m0 = a.mask[(size_t) n_top.x * C + c];
                if (n_top.y == 2 && n_top.w < 0) n_m1 = a.mask[(size_t) (n_top.x + 1) * C + c];
            }
        }
        const uint64_t own = on ? eff_from_raw(cu_mo, a.t, a.zc, C, node, c, use_alt, a.ln.kmask) : 0ULL;
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
        mbar_wait(bar, phase);
        phase ^= 1u;
        double *dst = a.vec + ((size_t) slot * T + tile) * run + jl;
        int ex = 0;
        int bad_child = -1;
        double b0 = 0., b1 = 0.;
        if (on && bin) {
            if (d_top.z < 0) {
                flow_expand_mask<KC, JT>(in0, m0, K);
            } else {
                ex += __double2loint(in0[K * JT]);
            }
            if (d_top.w < 0) {
                flow_expand_mask<KC, JT>(in1, m1, K);
            } else {
                ex += __double2loint(in1[K * JT]);
            }
            // walk 1: frequency-weighted sums of the two child vectors (two chains each for the FMA latency)
            double s0a = 0., s0b = 0., s1a = 0., s1b = 0.;
            if (KC) {
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    const double pk = spi[k * JT];
                    if (k & 1) { s0b = fma(pk, in0[k * JT], s0b); s1b = fma(pk, in1[k * JT], s1b); }
                    else { s0a = fma(pk, in0[k * JT], s0a); s1a = fma(pk, in1[k * JT], s1a); }
                }
            } else {
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
            }
            b0 = (1. - e0) * (s0a + s0b);
            b1 = (1. - e1) * (s1a + s1b);
        }
        // the previous chunk's vectors were stored at least a walk ago: announcing them now finds the stores drained
        // (the warp barrier at the end of the previous step ordered every lane's stores before this fence)
        if (flush_due) pending.flush(flags, T, epoch);
        flush_due = false;
        if (on && bin) {
            // walk 2: acc_k = allowed_k * (e0 x0_k + b0) * (e1 x1_k + b1), written straight to memory
            double vmax = 0.;
            if (KC) {
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    double v = fma(e0, in0[k * JT], b0);   // >= 0, the reference's np.maximum(., 0) is a no-op here
                    v = ((own >> k) & 1ULL) ? v : 0.;      // 10 ** log10(allowed), ml.py:127
                    v *= fma(e1, in1[k * JT], b1);
                    vmax = fmax(vmax, v);
                    dst[k * JT] = v;
                }
            } else {
                uint64_t ow = own;
#pragma unroll 4
                for (int k = 0; k < K; ++k, ow >>= 1) {
                    double v = fma(e0, in0[k * JT], b0);
                    v = (ow & 1ULL) ? v : 0.;
                    v *= fma(e1, in1[k * JT], b1);
                    vmax = fmax(vmax, v);
                    dst[k * JT] = v;
                }
            }
            if (vmax == 0.) {
                // all entries zero (PastMLLikelihoodError of ml.py:139-145): which child emptied the vector
                double after0 = 0.;
                for (int k = 0; k < K; ++k) after0 = fmax(after0, bit_as_double(own, k) * fma(e0, in0[k * JT], b0));
                bad_child = (after0 == 0.) ? fc : fc + 1;
            } else if (needs_rescale(vmax)) {
                // rare: the maximum left the exponent window, rewrite the vector scaled by an exact power of two
                const Rescale rs = make_rescale(vmax);
                for (int k = 0; k < K; ++k) {
                    double v = bit_as_double(own, k);
                    v *= fma(e0, in0[k * JT], b0);
                    v *= fma(e1, in1[k * JT], b1);
                    dst[k * JT] = rs.apply(v);
                }
                ex += rs.removed;
            }
        } else if (on) {
            // any other number of children: the staging column holds the running product, the children are read from
            // memory one after the other (their flags were checked by this node's issuer thread before the barrier)
            double *col = in0;
            for (int k = 0; k < K; ++k) col[k * JT] = bit_as_double(own, k);
            for (int q = 0; q < nc; ++q) {
                const int ch = fc + q;
                const int cs = a.t.slot[ch];
                const double *src = (cs >= 0) ? a.vec + ((size_t) cs * T + tile) * run + jl : nullptr;
                uint64_t m = 0ULL;
                if (cs < 0) m = eff_mask(a.mask, a.t, a.zc, C, ch, c, use_alt, a.ln.kmask);
                const double e = mu_inf ? 0. : exp_nonpos(-mu * ((a.t.dist[ch] + tau) * tauf * sf));
                double sw = 0.;
                for (int k = 0; k < K; ++k) sw = fma(spi[k * JT], src ? __ldcg(src + k * JT) : bit_as_double(m, k), sw);
                const double b = (1. - e) * sw;
                double mx = 0.;
                for (int k = 0; k < K; ++k) {
                    const double v = col[k * JT] * fma(e, src ? __ldcg(src + k * JT) : bit_as_double(m, k), b);
                    col[k * JT] = v;
                    mx = fmax(mx, v);
                }
                if (cs >= 0) ex += __double2loint(__ldcg(src + K * JT));
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
            dst[K * JT] = __hiloint2double(0, ex);
            if (jl == 0) pending.add(slot);
        }
        // the staging slot is free again (generic-proxy writes to it, tip masks or the running product, are ordered
        // before the copies of the async proxy); every lane's stores are ordered before the warp barrier, so that the
        // issuer's fence -- now or in a later step -- covers them
        fence_proxy_async();
        __syncwarp();
        const bool nx_ready = b_nx >= 0 && __all_sync(0xffffffffu, nf0 == epoch && nf1 == epoch && (n_slot < 0 || n_top.y == 2));
        if (nx_ready) {
            issue(n_top, n_slot);
            flush_due = last_cu;
        } else {
            // thin part of the tree (or the end): nothing may stay unannounced while this warp waits for others
            pending.flush(flags, T, epoch);
            if (b_nx >= 0) {
                wait_ready(n_top, n_slot);
                issue(n_top, n_slot);
            }
        }
        ++seq;
        b_cu = b_nx;
        last_cu = last_nx;
        b_nx = b_n2;
        last_nx = last_n2;
        if (b_cu >= 0) read_desc(seq, b_cu, cu);
        cu_mo = n_mo; cu_m0 = n_m0; cu_m1 = n_m1;
    }
}

// ------------------------------------------------------------------------------------------ top-down sweep

struct FlowTdArgs {
    TreeDev t;
    ZcDev zc;
    LaneDev ln;                 // one parameter set: J = C real lanes, padded to T tiles of JT
    const uint64_t *mask;       // annotation masks [N][C]
    const double *bu;           // [NI][T][K + 1][JT] (flow layout, from the bottom-up sweep)
    double *td;                 // [NI][T][K + 1][JT]
    const int *joint_state;     // [N][C] or null
    int force_joint;
    double *probs;              // [N][C][K] or null
    double *node_log10lh;       // [N][C] or null: raw sums, finished by log10_finish_kernel
    int *node_ex;               // [N][C]
    uint64_t *map_mask;         // [N][C]
    uint64_t *mppa_mask;        // [N][C]
    unsigned long long *k_hist; // [C][K+1]
    double *zc_lh;              // [rows][C][K]
    uint8_t *zc_k;              // [rows][C]
    FlowSched sched;            // desc: 2 x int4 per scheduled node: {parent slot, own slot, dist (2 words)},
                                //       {node, parent region of the batch (-1: a root), 1 = issues the parent's copies, -}
    unsigned *ctl;
    unsigned *flags;            // [NI][T] top-down vector of the slot is written
};

#ifndef PML_FLOW_TD_MINB
#define PML_FLOW_TD_MINB 4
#endif

// nodes per warp batch and parent regions per batch (siblings share their parent's two runs)
__host__ __device__ constexpr int flow_td_npw(int G, int JT) { return 32 / (G * JT); }
__host__ __device__ constexpr int flow_td_npar(int G, int JT) { return flow_td_npw(G, JT) > 1 ? flow_td_npw(G, JT) / 2 : 1; }
__host__ __device__ constexpr int flow_td_stage_bytes(int K, int G, int JT) {
    return (flow_td_npw(G, JT) + 2 * flow_td_npar(G, JT)) * (K + 1) * JT * 8;
}
// per warp: staging of the batch's runs, descriptor ring, pending list, mbarriers (stage + ring)
__host__ __device__ constexpr int flow_td_warp_bytes(int K, int G, int JT) {
    return (flow_td_stage_bytes(K, G, JT) + kFlowRing * flow_td_npw(G, JT) * 32 + flow_td_npw(G, JT) * kFlowPend * 4 + 8 * (1 + kFlowRing) + 127) & ~127;
}
__host__ __device__ inline size_t flow_td_smem(int K, int G, int JT, int block) {
    return (size_t) round16((4 + K) * JT * 8 + JT * (K + 1) * 4) + (size_t) kFlowMaxSeg * 16 +
           (size_t) (block / 32) * flow_td_warp_bytes(K, G, JT);
}

// calc_node_td_likelihood (ml.py:273-290) + calc_node_marginal_likelihood (ml.py:454-465) +
// convert_likelihoods_to_probabilities (ml.py:486-502) + choose_ancestral_states_map / _mppa (ml.py:505-595), fused and
// node-centric as td_coop_kernel (pml_coop.cuh), for every node of the forest in one launch.  G threads share a
// (node, lane) item, KT states each; the node's own bottom-up run and its parent's bottom-up and top-down runs are
// staged by bulk copies, the parent's once for all the siblings of a batch.  The stage is drained into registers at
// the top of a step and refilled at once when the next batch's parents are finished (always, in the fat part of the
// tree), so the copies fly under the whole step.
template <int KT, int G, int JT, bool KX>
__global__ void __launch_bounds__(PML_FLOW_BLOCK, PML_FLOW_TD_MINB) td_flow_kernel(FlowTdArgs a) {
    extern __shared__ double s_par[];
    static_assert(32 % (G * JT) == 0, "a warp holds whole (node, lane) groups");
    constexpr int NPW = flow_td_npw(G, JT), NPAR = flow_td_npar(G, JT);
    const int K = KX ? G * KT : a.ln.K, C = a.ln.C;
    const int KS = K + 1;
    const int nf = 4 + K;
    const int T = gridDim.y, tile = blockIdx.y, j0 = tile * JT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / (G * JT), rem = lane - sub * (G * JT), r = rem / JT, jl = rem - r * JT;
    const int cg = min(j0 + jl, C - 1);          // column this thread works on (padding lanes repeat the last one)
    const bool lane_real = j0 + jl < C;
    const int VB = KS * JT * 8;
    const size_t run = (size_t) KS * JT;
    unsigned int *sh_hist = reinterpret_cast<unsigned int *>(s_par + (size_t) nf * JT);   // [JT][K+1]
    unsigned char *seg_base = reinterpret_cast<unsigned char *>(s_par) + round16(nf * JT * 8 + JT * (K + 1) * 4);
    int4 *s_seg = reinterpret_cast<int4 *>(seg_base);
    unsigned char *my_stage = seg_base + kFlowMaxSeg * 16 + (size_t) warp * flow_td_warp_bytes(K, G, JT);
    unsigned char *ring_base = my_stage + flow_td_stage_bytes(K, G, JT);
    int *pend_base = reinterpret_cast<int *>(ring_base + kFlowRing * NPW * 32);
    const uint32_t bar = smem_addr(reinterpret_cast<unsigned char *>(pend_base) + NPW * kFlowPend * 4);
    FlowDescRing ring;
    ring.init(ring_base, bar + 8, NPW * 32, a.sched.desc);
    if (lane == 0) {
        mbar_init(bar, NPW);
        mbar_fence_init();
    }
    for (int i = threadIdx.x; i < JT * (K + 1); i += blockDim.x) sh_hist[i] = 0u;
    {
        const int total = nf * JT;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int f = i / JT, jj = i - f * JT;
            s_par[i] = a.ln.params[(size_t) min(j0 + jj, C - 1) * a.ln.stride + f];
        }
        for (int i = threadIdx.x; i < a.sched.n_seg; i += blockDim.x) s_seg[i] = a.sched.seg[i];
        __syncthreads();
    }
    const unsigned epoch = ld_volatile_u32(a.ctl + 1);
    const GroupOps<G, true> grp(lane, r, JT);
    const double sf = s_par[jl], tau = s_par[JT + jl], tauf = s_par[2 * JT + jl], mu = s_par[3 * JT + jl];
    double pi[KT];
#pragma unroll
    for (int i = 0; i < KT; ++i) pi[i] = PML_VALID(i) ? s_par[(4 + r + G * i) * JT + jl] : 0.;
    const bool use_alt = tau == 0.;   // ml.py:702-703
    const bool mu_inf = isinf(mu);
    const bool fj = a.force_joint != 0;
    const int n_root_slots = a.t.n_root_slots;
    unsigned *flags = a.flags + tile;   // flag of slot s at flags[s * T]
    const bool issuer = r == 0 && jl == 0;
    FlowPending pending;
    pending.init(pend_base + sub * kFlowPend);

    // descriptor of this thread's node in a batch: {parent slot, own slot, dist}, {node, parent region, issues parent}
    struct Desc { int4 d; int node, pidx, pissue; };
    auto read_desc = [&](int seq, int b, Desc &d) {
        if (b < 0) { d.d = make_int4(-1, -1, 0, 0); d.node = -1; d.pidx = -1; d.pissue = 0; return; }
        const int4 *p = ring.get(seq) + 2 * sub;
        d.d = p[0];
        const int4 w = p[1];
        d.node = w.x;
        d.pidx = w.y;
        d.pissue = w.z;
    };
    // does this thread have a flag to look at for the batch: the first sibling's issuer, when the parent is not a root
    auto has_flag = [&](const Desc &d) { return issuer && d.node >= 0 && d.pissue && d.d.x >= n_root_slots; };
    // stage layout: [own run of node 0..NPW-1][parent region 0: bottom-up run, top-down run][region 1] ...
    auto issue = [&](const Desc &d) {
        if (!issuer) return;
        const int ps = d.d.x, cs = d.d.y;
        const bool have = d.node >= 0;
        const bool pb = have && d.pissue && ps >= 0, ptd = pb && ps >= n_root_slots;
        mbar_arrive_expect(bar, (have && cs >= 0 ? VB : 0) + (pb ? VB : 0) + (ptd ? VB : 0));
        if (!have) return;
        fence_proxy_async_global();
        const uint32_t stage = smem_addr(my_stage);
        if (cs >= 0) bulk_g2s(stage + sub * VB, a.bu + ((size_t) cs * T + tile) * run, VB, bar);
        if (pb) bulk_g2s(stage + (NPW + 2 * d.pidx) * VB, a.bu + ((size_t) ps * T + tile) * run, VB, bar);
        if (ptd) bulk_g2s(stage + (NPW + 2 * d.pidx + 1) * VB, a.td + ((size_t) ps * T + tile) * run, VB, bar);
    };

    FlowTickets tk;
    tk.init(s_seg, a.sched.n_seg, a.sched.n_chunks, a.ctl + kFlowCtlTile0 + kFlowCtlStride * tile);
    __syncwarp();
    int b_cu = tk.next();
    bool last_cu = tk.last;
    int b_nx = tk.next();
    bool last_nx = tk.last;
    ring.fetch(0, b_cu);
    ring.fetch(1, b_nx);
    Desc cu;
    uint64_t cu_raw = 0;
    int cu_js = 0;
    uint32_t phase = 0;
    bool flush_due = false;   // the previous batch closed its chunk in steady state: announce before this step's stores
    int seq = 0;
    read_desc(0, b_cu, cu);
    if (b_cu >= 0) {
        if (has_flag(cu)) flag_wait(flags + (size_t) cu.d.x * T, epoch, a.ctl);
        issue(cu);
        if (cu.node >= 0) {
            cu_raw = a.mask[(size_t) cu.node * C + cg];
            if (fj) cu_js = a.joint_state[(size_t) cu.node * C + cg];
        }
    }
    while (b_cu >= 0) {
        const int b_n2 = tk.next();
        const bool last_n2 = tk.last;
        ring.fetch(seq + 2, b_n2);
        const bool have = cu.node >= 0;
        const bool on = have && lane_real;
        const int node = have ? cu.node : 0;
        const int4 desc = cu.d;
        const int ps = desc.x, cs = desc.y;
        const bool has_parent = ps >= 0;
        const size_t nc = (size_t) node * C + cg;
        const uint64_t raw_mask = cu_raw;
        const int js_pre = cu_js;
        // the next batch: descriptor from the ring, the parent's flag as seen now (looked at once the stage has been
        // drained), mask word and joint state
        Desc nx;
        read_desc(seq + 1, b_nx, nx);
        const unsigned nx_flag = has_flag(nx) ? ld_relaxed_u32(flags + (size_t) nx.d.x * T) : epoch;
        uint64_t n_raw = 0;
        int n_js = 0;
        if (nx.node >= 0) {
            n_raw = a.mask[(size_t) nx.node * C + cg];
            if (fj) n_js = a.joint_state[(size_t) nx.node * C + cg];
        }

        const double *st_vec = reinterpret_cast<const double *>(my_stage);
        const double *own_run = st_vec + (size_t) sub * run;
        const double *par_run = st_vec + (size_t) (NPW + 2 * max(cu.pidx, 0)) * run;
        double Lc[KT], X[KT];
        int exc = 0, pex = 0, ex_total;
        mbar_wait(bar, phase);
        phase ^= 1u;
        const uint64_t base_mask = raw_mask;
        const uint64_t allowed = eff_from_raw(raw_mask, a.t, a.zc, C, node, cg, use_alt, a.ln.kmask);
        if (cs >= 0) {
#pragma unroll
            for (int i = 0; i < KT; ++i) Lc[i] = PML_VALID(i) ? own_run[(r + G * i) * JT + jl] : 0.;
            exc = __double2loint(own_run[K * JT + jl]);
        } else {   // a tip: its vector is its allowed-state mask
            const uint64_t m = allowed >> r;
#pragma unroll
            for (int i = 0; i < KT; ++i) Lc[i] = bit_as_double(m, G * i);
        }
        ex_total = exc;
        // parent part of the node's top-down vector: TD_p * L_p (all ones above a root, ml.py:274-277)
        if (has_parent) {
#pragma unroll
            for (int i = 0; i < KT; ++i) X[i] = PML_VALID(i) ? par_run[(r + G * i) * JT + jl] : 0.;
            pex = __double2loint(par_run[K * JT + jl]);
            if (ps >= n_root_slots) {
#pragma unroll
                for (int i = 0; i < KT; ++i) X[i] *= PML_VALID(i) ? par_run[run + (r + G * i) * JT + jl] : 0.;
                pex += __double2loint(par_run[run + K * JT + jl]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < KT; ++i) X[i] = 1.;
        }
        // the stage is drained into registers: refill it right away when the next batch's parents are finished;
        // otherwise right after this batch's own vectors have been announced
        __syncwarp();
        const bool nx_issued = b_nx >= 0 && __all_sync(kFullMask, nx_flag == epoch);
        if (nx_issued) issue(nx);
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
            // contribution_k = e L_k + (1 - e) s >= (1 - e) s: when that floor is a normal number away from the ends of
            // the exponent range (practically always), every contribution is, and the reciprocals take the short path
            // -- hardware approximation + two Newton steps, ~1 ulp -- without a range check per element
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
        // the previous chunk's top-down vectors were stored most of a step ago: announcing them now finds the stores
        // drained (the warp barrier after the drain ordered every lane's stores before this fence)
        if (flush_due) pending.flush(flags, T, epoch);
        flush_due = false;
        const bool publishes = issuer && have && has_parent && cs >= 0;
        if (has_parent) {
            if (cs >= 0 && have) {
                double *dst = a.td + ((size_t) cs * T + tile) * run + jl;
#pragma unroll
                for (int i = 0; i < KT; ++i)
                    if (PML_VALID(i)) dst[(r + G * i) * JT] = tdv[i];
                if (r == 0) dst[K * JT] = __hiloint2double(0, qex);
            }
            ex_total = qex + exc;
        }
        if (publishes) pending.add(cs);
        if (nx_issued) {
            flush_due = last_cu;   // steady state: announced with its chunk, in a later step
        } else {
            // the next batch is not ready (thin part of the tree) or there is none: announce at once -- the children
            // may be waiting, and nothing may stay unannounced while this warp waits for others
            __syncwarp();
            pending.flush(flags, T, epoch);
            if (b_nx >= 0) {
                if (has_flag(nx)) flag_wait(flags + (size_t) nx.d.x * T, epoch, a.ctl);
                issue(nx);
            }
        }
        // marginal likelihood vector (ml.py:454-465): L * TD * pi * allowed
        double m[KT];
        double sum = 0.;
        const uint64_t allowed_r = allowed >> r;
#pragma unroll
        for (int i = 0; i < KT; ++i) {
            const double tdi = has_parent ? tdv[i] : X[i];   // roots: all ones
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
            if (a.node_log10lh && r == 0) {
                a.node_log10lh[nc] = sum;
                a.node_ex[nc] = ex_total;
            }
        }
        // members of zero-branch clusters: remember the vector, apply '.initial' (ml.py:593-594)
        uint64_t init = a.ln.kmask;
        size_t zr = 0;
        if (flag && have) {
            zr = (size_t) a.t.zc_index[node] * C + cg;
            if (on) {
                double *z = a.zc_lh + zr * K + r;
#pragma unroll
                for (int i = 0; i < KT; ++i)
                    if (PML_VALID(i)) z[G * i] = m[i];
            }
            if (a.zc.hasinit[zr]) init = a.zc.initial[zr];
        }
        if (__any_sync(kFullMask, flag != 0ULL && have)) {   // rare: some node of the batch sits in a zero-branch cluster
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
        const int ji = fj ? js_pre : best;
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
            double runsum = 0., best_score = CUDART_INF;
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
                    const double prev = runsum;
                    runsum += lv * inv;
                    joint_seen = joint_seen || (lk == ji);
                    const double tk_ = (!fj || joint_seen) ? runsum : pj + prev;
                    const double score = (1. - 2. * tk_) / (double) k;
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
            a.map_mask[nc] = (1ULL << best) | flag;
            a.mppa_mask[nc] = sel | flag;
            if (flag) a.zc_k[zr] = (uint8_t) best_k;
            atomicAdd(sh_hist + jl * (K + 1) + best_k, 1u);
        }
        ++seq;
        b_cu = b_nx;
        last_cu = last_nx;
        b_nx = b_n2;
        last_nx = last_n2;
        cu = nx;
        cu_raw = n_raw;
        cu_js = n_js;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < JT * (K + 1); i += blockDim.x) {
        const int jj = i / (K + 1), k = i - jj * (K + 1);
        if (sh_hist[i] && j0 + jj < C) atomicAdd(a.k_hist + (size_t) (j0 + jj) * (K + 1) + k, (unsigned long long) sh_hist[i]);
    }
}

// ln(sum_i L_root,i pi_i) + ex ln 2 summed over the roots (ml.py:112-121) on the flow layout; a root may be a tip.
static __global__ void root_loglik_flow_kernel(TreeDev t, ZcDev zc, LaneDev ln, const uint64_t *mask, int alter,
                                               const double *vec, int JT, const int *roots, int n_roots, double *out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ln.J) return;
    const int K = ln.K, c = lane_column(ln, j);
    const int T = (ln.J + JT - 1) / JT, tile = j / JT, jl = j - tile * JT;
    const double *p = ln.params + (size_t) j * ln.stride;
    const bool use_alt = alter && p[1] == 0.;
    double total = 0.;
    for (int r = 0; r < n_roots; ++r) {
        const int node = roots[r];
        const int s = t.slot[node];
        const double *v = vec + (((size_t) max(s, 0) * T + tile) * (K + 1)) * JT + jl;
        const int ex = (s >= 0) ? __double2loint(v[(size_t) K * JT]) : 0;
        const uint64_t m = (s >= 0) ? 0ULL : eff_mask(mask, t, zc, ln.C, node, c, use_alt, ln.kmask);
        double acc = 0.;
        for (int k = 0; k < K; ++k) acc += ((s >= 0) ? v[(size_t) k * JT] : bit_as_double(m, k)) * p[4 + k];
        total += log(acc) + (double) ex * 0.69314718055994530942;
    }
    out[j] = total;
}

}  // namespace pml
