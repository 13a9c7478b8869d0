// Single-launch top-down sweep, thread-per-item: one thread owns a (node, lane) item and walks its K states
// sequentially, streaming the staged runs from shared memory.
//
// The cooperative version (td_flow_kernel, pml_flow.cuh: two threads per item, everything in registers) spends ~690
// warp instructions per node, most of them not arithmetic: every scalar of the step (branch exponential, exponents,
// masks, flags, the MPPA loop control, address arithmetic) is replicated over the threads that share an item, and
// every reduction is a shuffle.  Here nothing is replicated and nothing crosses threads.  What made thread-per-item
// slow before (three K-vectors in registers: 236 registers, 8 warps per SM) is avoided by keeping ONE K-vector in
// registers and the rest in the staging slot:
//   pass A   s = sum_k pi_k L_k                                  reads the node's own column
//   pass B   q_k = TDp_k BUp_k / (e L_k + (1 - e) s), sum_k pi_k q_k;  q stays in registers, the own column is
//            overwritten in place by w_k = allowed_k L_k pi_k (the column belongs to this thread alone)
//   pass C   td_k = e q_k + (1 - e) s2 -> memory;  m_k = w_k td_k replaces q_k in registers
//   pass D   probabilities, MAP, MPPA from the registers
// The stage is free after pass C: the next batch's copies fly under pass D and the next step's preamble.
// Layout, schedule, flags, tickets, descriptor ring and chunked publication are those of pml_flow.cuh.
#pragma once
#include "pml_flow.cuh"

namespace pml {

#ifndef PML_FLOW_TDI_MINB
#define PML_FLOW_TDI_MINB 4
#endif

__host__ __device__ constexpr int flow_tdi_npw(int JT) { return 32 / JT; }
__host__ __device__ constexpr int flow_tdi_npar(int JT) { return flow_tdi_npw(JT) > 1 ? flow_tdi_npw(JT) / 2 : 1; }
__host__ __device__ constexpr int flow_tdi_stage_bytes(int K, int JT) {
    return (flow_tdi_npw(JT) + 2 * flow_tdi_npar(JT)) * (K + 1) * JT * 8;
}
__host__ __device__ constexpr int flow_tdi_warp_bytes(int K, int JT) {
    return (flow_tdi_stage_bytes(K, JT) + kFlowRing * flow_tdi_npw(JT) * 32 + flow_tdi_npw(JT) * kFlowPend * 4 + 8 * (1 + kFlowRing) + 127) & ~127;
}
__host__ __device__ inline size_t flow_tdi_smem(int K, int JT, int block) {
    return (size_t) round16((4 + K) * JT * 8 + JT * (K + 1) * 4) + (size_t) kFlowMaxSeg * 16 +
           (size_t) (block / 32) * flow_tdi_warp_bytes(K, JT);
}

// KP: register-array length (K <= KP); KX: K == KP exactly.
template <int KP, int JT, bool KX>
__global__ void __launch_bounds__(PML_FLOW_BLOCK, PML_FLOW_TDI_MINB) td_tpi_flow_kernel(FlowTdArgs a) {
    extern __shared__ double s_par[];
    constexpr int NPW = flow_tdi_npw(JT), NPAR = flow_tdi_npar(JT);
    const int K = KX ? KP : a.ln.K, C = a.ln.C;
    const int KS = K + 1;
    const int nf = 4 + K;
    const int T = gridDim.y, tile = blockIdx.y, j0 = tile * JT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / JT, jl = lane - sub * JT;
    const int cg = min(j0 + jl, C - 1);          // column this thread works on (padding lanes repeat the last one)
    const bool lane_real = j0 + jl < C;
    const int VB = KS * JT * 8;
    const size_t run = (size_t) KS * JT;
    unsigned int *sh_hist = reinterpret_cast<unsigned int *>(s_par + (size_t) nf * JT);   // [JT][K+1]
    unsigned char *seg_base = reinterpret_cast<unsigned char *>(s_par) + round16(nf * JT * 8 + JT * (K + 1) * 4);
    int4 *s_seg = reinterpret_cast<int4 *>(seg_base);
    unsigned char *my_stage = seg_base + kFlowMaxSeg * 16 + (size_t) warp * flow_tdi_warp_bytes(K, JT);
    unsigned char *ring_base = my_stage + flow_tdi_stage_bytes(K, JT);
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
    const double sf = s_par[jl], tau = s_par[JT + jl], tauf = s_par[2 * JT + jl], mu = s_par[3 * JT + jl];
    const double *spi = s_par + 4 * JT + jl;   // pi_k at spi[k * JT]
    const bool use_alt = tau == 0.;   // ml.py:702-703
    const bool mu_inf = isinf(mu);
    const bool fj = a.force_joint != 0;
    const int n_root_slots = a.t.n_root_slots;
    unsigned *flags = a.flags + tile;   // flag of slot s at flags[s * T]
    const bool issuer = jl == 0;
    FlowPending pending;
    pending.init(pend_base + sub * kFlowPend);
    double *own = reinterpret_cast<double *>(my_stage) + (size_t) sub * run + jl;   // this item's column: state k at own[k * JT]

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
    bool flush_due = false;
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
        const int ps = cu.d.x, cs = cu.d.y;
        const bool has_parent = ps >= 0, ptd = ps >= n_root_slots;
        const size_t nc = (size_t) node * C + cg;
        const uint64_t base_mask = cu_raw;
        const int js_pre = cu_js;
        // the next batch: descriptor from the ring, the parent's flag as seen now (looked at after pass C), mask word
        // and joint state
        Desc nx;
        read_desc(seq + 1, b_nx, nx);
        const unsigned nx_flag = has_flag(nx) ? ld_relaxed_u32(flags + (size_t) nx.d.x * T) : epoch;
        uint64_t n_raw = 0;
        int n_js = 0;
        if (nx.node >= 0) {
            n_raw = a.mask[(size_t) nx.node * C + cg];
            if (fj) n_js = a.joint_state[(size_t) nx.node * C + cg];
        }
        const uint64_t allowed = eff_from_raw(base_mask, a.t, a.zc, C, node, cg, use_alt, a.ln.kmask);
        const uint64_t flag = base_mask & kMemberFlag;
        double e = 0.;
        if (!mu_inf) e = exp_nonpos(-mu * ((__hiloint2double(cu.d.w, cu.d.z) + tau) * tauf * sf));
        const double ome = 1. - e;
        const double *pbu = reinterpret_cast<const double *>(my_stage) + (size_t) (NPW + 2 * max(cu.pidx, 0)) * run + jl;
        const double *ptdv = pbu + run;

        mbar_wait(bar, phase);
        phase ^= 1u;
        int exc = 0, pex = 0;
        if (cs >= 0) {
            exc = __double2loint(own[K * JT]);
        } else {   // a tip: its vector is its allowed-state mask, written into the (unused) own column
            uint64_t m = allowed;
#pragma unroll
            for (int k = 0; k < KP; ++k, m >>= 1)
                if (KX || k < K) own[k * JT] = __hiloint2double((m & 1ULL) ? 0x3ff00000 : 0, 0);
        }
        if (has_parent) {
            pex = __double2loint(pbu[K * JT]);
            if (ptd) pex += __double2loint(ptdv[K * JT]);
        }
        // pass A: node_contribution needs s = sum_k pi_k L_k (ml.py:280)
        double s;
        {
            double sa = 0., sb = 0.;
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                if (KX || k < K) {
                    if (k & 1) sb = fma(spi[k * JT], own[k * JT], sb);
                    else sa = fma(spi[k * JT], own[k * JT], sa);
                }
            }
            s = sa + sb;
        }
        // pass B: q = TD_p L_p / contribution, contribution <= 0 -> 1 (ml.py:281-283); contribution_k >= (1 - e) s, so
        // when that floor is a normal number away from the ends of the exponent range every reciprocal takes the short
        // path (hardware approximation + two Newton steps, ~1 ulp)
        double q[KP];
        double s2;
        {
            const double os = ome * s;
            const bool fast = __all_sync(kFullMask, os > 0x1p-900 && s < 0x1p+800);
            double s2a = 0., s2b = 0.;
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                if (KX || k < K) {
                    const double Lk = own[k * JT];
                    double Xk = 1.;   // all ones above a root (ml.py:274-277)
                    if (has_parent) {
                        Xk = pbu[k * JT];
                        if (ptd) Xk *= ptdv[k * JT];
                    }
                    const double cv = fma(e, Lk, os);
                    double qk;
                    if (fast) {
                        double rr;
                        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rr) : "d"(cv));
                        double tt = fma(-cv, rr, 1.);
                        rr = fma(rr, tt, rr);
                        tt = fma(-cv, rr, 1.);
                        rr = fma(rr, tt, rr);
                        qk = Xk * rr;
                    } else {
                        qk = (cv > 0.) ? Xk * fast_rcp(cv) : Xk;
                    }
                    const double pk = spi[k * JT];
                    if (k & 1) s2b = fma(pk, qk, s2b);
                    else s2a = fma(pk, qk, s2a);
                    q[k] = qk;
                    own[k * JT] = ((allowed >> k) & 1ULL) ? Lk * pk : 0.;   // w_k: L * pi * allowed (ml.py:454-458)
                } else {
                    q[k] = 0.;
                }
            }
            s2 = s2a + s2b;
        }
        // keeping the vector in range (ml.py:284-286): the frequency-weighted sum is within a factor 1/min(pi) of the
        // maximum, so its exponent decides the power-of-two rescale
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
        // the previous chunk's top-down vectors were stored most of a step ago: announcing them now finds the stores
        // drained (the warp barrier of the previous step ordered every lane's stores before this fence)
        if (flush_due) pending.flush(flags, T, epoch);
        flush_due = false;
        // pass C: TD = P (TD_p L_p / contribution) (ml.py:288-289), stored for internal nodes; marginal m = w * TD
        const bool stores_td = have && has_parent && cs >= 0;
        double *dst = a.td + ((size_t) max(cs, 0) * T + tile) * run + jl;
        double sum = 0.;
        {
            const double base = ome * s2;
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                if (KX || k < K) {
                    const double td = fma(e, q[k], base) * scale;   // already >= 0
                    if (stores_td) dst[k * JT] = td;
                    const double mk = has_parent ? own[k * JT] * td : own[k * JT];   // roots: all ones
                    q[k] = mk;
                    sum += mk;
                }
            }
        }
        int ex_total = exc;
        if (has_parent) {
            if (stores_td) dst[K * JT] = __hiloint2double(0, qex);
            ex_total = qex + exc;
        }
        if (stores_td && issuer) pending.add(cs);
        // the stage is free: refill it when the next batch's parents are finished (always, in the fat part of the
        // tree); otherwise announce this batch at once and wait
        fence_proxy_async();
        __syncwarp();
        const bool nx_ready = b_nx >= 0 && __all_sync(kFullMask, nx_flag == epoch);
        if (nx_ready) {
            issue(nx);
            flush_due = last_cu;
        } else {
            pending.flush(flags, T, epoch);
            if (b_nx >= 0) {
                if (has_flag(nx)) flag_wait(flags + (size_t) nx.d.x * T, epoch, a.ctl);
                issue(nx);
            }
        }
        // pass D: probabilities (ml.py:498-500), node likelihood, MAP (ml.py:577-595), MPPA (ml.py:539-572)
        double inv = __drcp_rn(sum);
        if (on) {
            if (a.probs) {
                double *pd = a.probs + nc * K;
#pragma unroll
                for (int k = 0; k < KP; ++k)
                    if (KX || k < K) pd[k] = q[k] * inv;
            }
            if (a.node_log10lh) {
                a.node_log10lh[nc] = sum;
                a.node_ex[nc] = ex_total;
            }
        }
        size_t zr = 0;
        if (flag && have) {   // members of zero-branch clusters: remember the vector, apply '.initial' (ml.py:593-594)
            zr = (size_t) a.t.zc_index[node] * C + cg;
            if (on) {
                double *z = a.zc_lh + zr * K;
#pragma unroll
                for (int k = 0; k < KP; ++k)
                    if (KX || k < K) z[k] = q[k];
            }
            if (a.zc.hasinit[zr]) {
                const uint64_t init = a.zc.initial[zr];
                sum = 0.;
#pragma unroll
                for (int k = 0; k < KP; ++k) {
                    q[k] = ((init >> k) & 1ULL) ? q[k] : 0.;
                    sum += q[k];
                }
                inv = 1. / sum;
            }
        }
        int best = 0;   // first arg max (ml.py:595)
        double bv = q[0];
#pragma unroll
        for (int k = 1; k < KP; ++k)
            if ((KX || k < K) && q[k] > bv) { bv = q[k]; best = k; }
        const int ji = fj ? js_pre : best;
        int best_k = 1;
        uint64_t sel = 1ULL << best;
        if (!(sum > 0.)) { sel = a.ln.kmask; best_k = K; }   // NaN probabilities: no k beats +inf, all states kept
        else if (!(bv * inv >= 0.76 && ji == best)) sel = mppa_select<KP>(q, inv, K, ji, fj, best_k);
        if (on) {
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

}  // namespace pml
