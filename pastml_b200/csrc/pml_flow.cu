// Host side of the single-launch sweeps (pml_flow.cuh): schedules, launch geometry, bottom-up instantiations.
#include <algorithm>
#include <cstdlib>
#include <string>
#include <vector>

#include "pml_flow.cuh"
#include "pml_launch.h"
#include "pml_flow_host.h"

namespace pml {

// ------------------------------------------------------------------------------------------ schedules

namespace {

// chunk size for a level of nb batches: the fat levels hand out several batches per ticket (fewer atomics on one
// address), the thin ones one batch per ticket so that a level spreads over as many warps as it has batches
int chunk_for(long long nb) {
    const long long warps = 148LL * 16;
    long long c = nb / (2 * warps);
    if (c < 1) c = 1;
    if (c > 8) c = 8;
    return (int) c;
}

struct SegBuilder {
    std::vector<FlowSeg> seg;
    int n_batches = 0, n_chunks = 0;
    void add_level(int nb) {
        if (nb <= 0) return;
        const int ch = chunk_for(nb);
        if (ch == 1 && !seg.empty() && seg.back().chunk == 1) {   // thin levels share a segment
            seg.back().n_batches += nb;
        } else {
            seg.push_back(FlowSeg{n_chunks, n_batches, nb, ch});
        }
        n_batches += nb;
        n_chunks += (nb + ch - 1) / ch;
    }
    void add_segment(int nb, int ch) {   // a run of batches that is not one level (interleaved runs)
        if (nb <= 0) return;
        seg.push_back(FlowSeg{n_chunks, n_batches, nb, ch});
        n_batches += nb;
        n_chunks += (nb + ch - 1) / ch;
    }
    void finish(FlowHostSched &out) {
        if ((int) seg.size() > kFlowMaxSeg) {   // cannot happen for real trees; one batch per ticket everywhere
            seg.assign(1, FlowSeg{0, 0, n_batches, 1});
            n_chunks = n_batches;
        }
        out.seg.resize(seg.size() * 4);
        for (size_t i = 0; i < seg.size(); ++i) {
            out.seg[4 * i] = seg[i].first_chunk;
            out.seg[4 * i + 1] = seg[i].first_batch;
            out.seg[4 * i + 2] = seg[i].n_batches;
            out.seg[4 * i + 3] = seg[i].chunk;
        }
        out.n_seg = (int) seg.size();
        out.n_batches = n_batches;
        out.n_chunks = n_chunks;
    }
};

}  // namespace

// Schedules.  The base order is level order (bottom-up: deepest level first; top-down: roots first), cut into batches
// that never cross a level, so a batch never holds a node together with one of its operands' producers.  Level
// order has a weakness: the deepest levels of a big tree are thin (a few nodes on long root-to-tip paths), and a run
// of thin levels is a chain of dependent steps during which almost every warp would sit on a batch it cannot start.
// The thin run at the deep end is therefore INTERLEAVED with filler batches that depend on nothing in the sweep:
//   bottom-up  the thin tail comes first; filler = cherries (nodes whose children are all tips) of the fat levels,
//              which can run at any time before their parents;
//   top-down   the thin tail comes last; filler = batches made of tips only (nobody waits for a tip), held back from
//              the fat levels and run after their parents.
// One batch in (1 + kFlowFill) is then chain-limited, the others keep the machine busy.  The thin run at the ROOT end
// (last dozen levels bottom-up, first dozen top-down) has no filler by construction: everything else depends on it or
// it depends on everything else.
namespace {
constexpr int kFlowFill = 7;         // filler batches per chain-limited batch
// a level with fewer batches than two per resident warp is thin (PASTML_B200_FLOW_THIN overrides: tests, tuning)
int flow_thin() { return getenv("PASTML_B200_FLOW_THIN") ? atoi(getenv("PASTML_B200_FLOW_THIN")) : 2 * 148 * 16; }
bool flow_interleave() {   // PASTML_B200_FLOW_SCHED=level: plain level order (A/B measurements)
    static const bool off = getenv("PASTML_B200_FLOW_SCHED") && std::string(getenv("PASTML_B200_FLOW_SCHED")) == "level";
    return !off;
}
}  // namespace

// bottom-up: internal-node slots, npw per batch
void flow_schedule_bu(const std::vector<int> &level_ptr, const int *parent, int npw, FlowHostSched &out) {
    out.order.clear();
    out.aux.clear();
    SegBuilder sb;
    const int n_levels = (int) level_ptr.size() - 1;
    const int N = level_ptr[n_levels];
    // slots and cherries: slot = rank among internal nodes; a cherry has no internal child
    std::vector<int> nch(N, 0), slot(N, -1);
    for (int v = 0; v < N; ++v)
        if (parent[v] >= 0) nch[parent[v]]++;
    std::vector<char> has_internal_child(N, 0);
    int NI = 0;
    for (int v = 0; v < N; ++v) {
        if (nch[v]) {
            slot[v] = NI++;
            if (parent[v] >= 0) has_internal_child[parent[v]] = 1;
        }
    }
    std::vector<int> int_ptr(n_levels + 1, 0);   // slot range of every level
    for (int l = 0, s = 0; l < n_levels; ++l) {
        for (int v = level_ptr[l]; v < level_ptr[l + 1]; ++v) s += nch[v] ? 1 : 0;
        int_ptr[l + 1] = s;
    }
    std::vector<char> cherry(NI, 0);
    for (int v = 0; v < N; ++v)
        if (nch[v] && !has_internal_child[v]) cherry[slot[v]] = 1;
    auto batches_of = [&](int l) { return (int_ptr[l + 1] - int_ptr[l] + npw - 1) / npw; };
    // the thin run at the deep end
    int tail_first = n_levels;   // levels [tail_first, n_levels) are the tail
    while (tail_first > 0 && batches_of(tail_first - 1) < flow_thin()) --tail_first;
    const bool all_thin = tail_first == 0;
    long long tail_batches = 0;
    for (int l = tail_first; l < n_levels; ++l) tail_batches += batches_of(l);
    std::vector<char> used(NI, 0);
    auto emit_level = [&](int l, bool skip_used) {
        int count = 0;
        for (int s = int_ptr[l]; s < int_ptr[l + 1]; ++s) {
            if (skip_used && used[s]) continue;
            out.order.push_back(s);
            ++count;
        }
        const int nb = (count + npw - 1) / npw;
        for (int i = count; i < nb * npw; ++i) out.order.push_back(-1);
        return nb;
    };
    if (all_thin || !flow_interleave() || tail_batches == 0) {
        for (int l = n_levels - 1; l >= 0; --l) sb.add_level(emit_level(l, false));
        sb.finish(out);
        return;
    }
    // filler: cherries of the fat levels, deepest fat level first
    std::vector<int> filler;
    const long long want = tail_batches * kFlowFill * npw;
    for (int l = tail_first - 1; l >= 0 && (long long) filler.size() < want; --l) {
        if (batches_of(l) < flow_thin()) continue;
        for (int s = int_ptr[l]; s < int_ptr[l + 1] && (long long) filler.size() < want; ++s)
            if (cherry[s]) { filler.push_back(s); used[s] = 1; }
    }
    const int fill = (int) std::min<long long>(kFlowFill, (long long) filler.size() / npw / std::max<long long>(tail_batches, 1));
    size_t f = 0;
    int nb1 = 0;
    for (int l = n_levels - 1; l >= tail_first; --l) {
        const int s0 = int_ptr[l], ns = int_ptr[l + 1] - s0;
        for (int b = 0; b * npw < ns; ++b) {
            for (int i = 0; i < npw; ++i) out.order.push_back(b * npw + i < ns ? s0 + b * npw + i : -1);
            ++nb1;
            for (int q = 0; q < fill; ++q) {
                for (int i = 0; i < npw; ++i) out.order.push_back(filler[f++]);
                ++nb1;
            }
        }
    }
    // filler that was reserved but not placed goes back to its level
    for (; f < filler.size(); ++f) used[filler[f]] = 0;
    sb.add_segment(nb1, 1);
    for (int l = tail_first - 1; l >= 0; --l) sb.add_level(emit_level(l, true));
    sb.finish(out);
}

// top-down: nodes; a batch holds at most npw nodes with at most npar distinct parents (siblings share the staged parent
// runs).  aux = (parent region, issues-the-parent-copies) per scheduled node
void flow_schedule_td(const std::vector<int> &level_ptr, const int *parent, int npw, int npar, FlowHostSched &out) {
    out.order.clear();
    out.aux.clear();
    SegBuilder sb;
    const int n_levels = (int) level_ptr.size() - 1;
    const int N = level_ptr[n_levels];
    std::vector<char> internal(N, 0);
    for (int v = 0; v < N; ++v)
        if (parent[v] >= 0) internal[parent[v]] = 1;
    // batches of every level, greedily: [first node, count] + "tips only"
    struct Batch { int first, count; char tips_only; };
    std::vector<Batch> batches;
    std::vector<int> level_batch_ptr(n_levels + 1, 0);
    for (int lvl = 0; lvl < n_levels; ++lvl) {
        const int n1 = level_ptr[lvl + 1];
        int v = level_ptr[lvl];
        while (v < n1) {
            Batch b{v, 0, 1};
            int regions = 0, last_parent = -2;
            while (v < n1 && b.count < npw) {
                const int q = parent[v];
                if (q >= 0 && q != last_parent) {
                    if (regions == npar) break;
                    ++regions;
                    last_parent = q;
                }
                if (internal[v]) b.tips_only = 0;
                ++b.count;
                ++v;
            }
            batches.push_back(b);
        }
        level_batch_ptr[lvl + 1] = (int) batches.size();
    }
    auto emit = [&](const Batch &b) {
        int regions = 0, last_parent = -2;
        for (int i = 0; i < b.count; ++i) {
            const int v = b.first + i, q = parent[v];
            int region = -1, pissue = 0;
            if (q >= 0) {
                if (q != last_parent) { ++regions; last_parent = q; pissue = 1; }
                region = regions - 1;
            }
            out.order.push_back(v);
            out.aux.push_back(region);
            out.aux.push_back(pissue);
        }
        for (int i = b.count; i < npw; ++i) {
            out.order.push_back(-1);
            out.aux.push_back(-1);
            out.aux.push_back(0);
        }
    };
    auto batches_of = [&](int l) { return level_batch_ptr[l + 1] - level_batch_ptr[l]; };
    int tail_first = n_levels;
    while (tail_first > 0 && batches_of(tail_first - 1) < flow_thin()) --tail_first;
    long long tail_batches = 0;
    for (int l = tail_first; l < n_levels; ++l) tail_batches += batches_of(l);
    if (tail_first == 0 || !flow_interleave() || tail_batches == 0) {
        for (int l = 0; l < n_levels; ++l) {
            for (int b = level_batch_ptr[l]; b < level_batch_ptr[l + 1]; ++b) emit(batches[b]);
            sb.add_level(batches_of(l));
        }
        sb.finish(out);
        return;
    }
    // filler: tips-only batches of the fat levels, deepest fat level first, held back for the tail
    std::vector<char> held(batches.size(), 0);
    std::vector<int> filler;
    const long long want = tail_batches * kFlowFill;
    for (int l = tail_first - 1; l >= 0 && (long long) filler.size() < want; --l) {
        if (batches_of(l) < flow_thin()) continue;
        for (int b = level_batch_ptr[l + 1] - 1; b >= level_batch_ptr[l] && (long long) filler.size() < want; --b)
            if (batches[b].tips_only) { filler.push_back(b); held[b] = 1; }
    }
    const int fill = (int) std::min<long long>(kFlowFill, (long long) filler.size() / std::max<long long>(tail_batches, 1));
    while ((long long) filler.size() > fill * tail_batches) { held[filler.back()] = 0; filler.pop_back(); }
    for (int l = 0; l < tail_first; ++l) {
        int nb = 0;
        for (int b = level_batch_ptr[l]; b < level_batch_ptr[l + 1]; ++b)
            if (!held[b]) { emit(batches[b]); ++nb; }
        sb.add_level(nb);
    }
    size_t f = 0;
    int nb3 = 0;
    for (int l = tail_first; l < n_levels; ++l) {
        for (int b = level_batch_ptr[l]; b < level_batch_ptr[l + 1]; ++b) {
            emit(batches[b]);
            ++nb3;
            for (int q = 0; q < fill; ++q) { emit(batches[filler[f++]]); ++nb3; }
        }
    }
    sb.add_segment(nb3, 1);
    sb.finish(out);
}

// descriptors in schedule order, gathered on the device from the per-slot / per-node tables
static __global__ void flow_gather_bu_kernel(long long n, const int *order, const int4 *bu_desc, const int *int_nodes,
                                             int4 *out) {
    const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = order[i];
    if (s < 0) {
        out[3 * i] = make_int4(0, 0, -1, -1);
        out[3 * i + 1] = make_int4(0, 0, 0, 0);
        out[3 * i + 2] = make_int4(0, -1, 0, 0);
        return;
    }
    out[3 * i] = bu_desc[2 * (size_t) s];
    out[3 * i + 1] = bu_desc[2 * (size_t) s + 1];
    out[3 * i + 2] = make_int4(int_nodes[s], s, 0, 0);
}

static __global__ void flow_gather_td_kernel(long long n, const int *order, const int *aux, const int4 *td_desc, int4 *out) {
    const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int v = order[i];
    if (v < 0) {
        out[2 * i] = make_int4(-1, -1, 0, 0);
        out[2 * i + 1] = make_int4(-1, -1, 0, 0);
        return;
    }
    out[2 * i] = td_desc[v];
    out[2 * i + 1] = make_int4(v, aux[2 * i], aux[2 * i + 1], 0);
}

void flow_gather_bu(cudaStream_t st, long long n, const int *order, const int4 *bu_desc, const int *int_nodes, int4 *out) {
    if (n > 0) flow_gather_bu_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, st>>>(n, order, bu_desc, int_nodes, out);
}
void flow_gather_td(cudaStream_t st, long long n, const int *order, const int *aux, const int4 *td_desc, int4 *out) {
    if (n > 0) flow_gather_td_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, st>>>(n, order, aux, td_desc, out);
}

int flow_ctl_word_count(int tiles) { return flow_ctl_words(tiles); }

void flow_begin(cudaStream_t st, unsigned *ctl, int tiles) {
    flow_begin_kernel<<<(tiles + 127) / 128, 128, 0, st>>>(ctl, tiles);
}

void flow_root_loglik(cudaStream_t st, const TreeDev &t, const ZcDev &zc, const LaneDev &ln, const uint64_t *mask, int alter,
                      const double *vec, int jt, const int *roots, int n_roots, double *out) {
    root_loglik_flow_kernel<<<grid_for(ln.J), kBlock, 0, st>>>(t, zc, ln, mask, alter, vec, jt, roots, n_roots, out);
}

// [slot][T][K + 1][JT] -> the level kernels' [slot][K][C] + [slot][C] (callers that read the bottom-up state back)
static __global__ void flow_unpack_kernel(long long n, int K, int C, int JT, const double *vec, double *out_vec, int *out_ex) {
    const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;   // over slot x (K + 1) x C
    if (i >= n) return;
    const int c = (int) (i % C);
    const int k = (int) ((i / C) % (K + 1));
    const long long slot = i / ((long long) C * (K + 1));
    const int T = (C + JT - 1) / JT, tile = c / JT, jl = c - tile * JT;
    const double v = vec[((slot * T + tile) * (K + 1) + k) * JT + jl];
    if (k < K) out_vec[(slot * K + k) * C + c] = v;
    else out_ex[slot * C + c] = __double2loint(v);
}
void flow_unpack(cudaStream_t st, long long n_slots, int K, int C, int jt, const double *vec, double *out_vec, int *out_ex) {
    const long long n = n_slots * (K + 1) * C;
    if (n > 0) flow_unpack_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, st>>>(n, K, C, jt, vec, out_vec, out_ex);
}

// ------------------------------------------------------------------------------------------ launch geometry

// lanes per tile for a pass with J lanes: the candidate with the least padded work, small tiles slightly penalised
// (short runs per bulk copy), ties to the wider tile
int flow_tile_lanes(int J, int max_jt) {
    int best = 0;
    double best_cost = 1e300;
    for (int jt : {32, 16, 8, 4, 2}) {
        if (jt > max_jt) continue;
        const int tiles = (J + jt - 1) / jt;
        const double cost = (double) tiles * jt * (1. + 2. / jt);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = jt; }
    }
    return best;
}

namespace {

template <int KC, int JT>
void bu_launch_s(cudaStream_t st, const FlowBuArgs &a, int tiles) {
    const int K = a.ln.K;
    int block = PML_FLOW_BLOCK;
    // keep at least two blocks per SM: the stages of one block may take up to half of the shared memory
    while (block > 32 && flow_bu_smem(K, JT, block) > kFlowSmemCap / 2) block /= 2;
    const size_t sh = flow_bu_smem(K, JT, block);
    static int resident[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    static size_t configured[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 7;
    if (sh != configured[dev]) {
        resident[dev] = resident_blocks(bu_flow_kernel<KC, JT>, block, sh);
        configured[dev] = sh;
    }
    const unsigned gx = flow_grid_x(resident[dev], tiles, a.sched.n_chunks, block);
    bu_flow_kernel<KC, JT><<<dim3(gx, (unsigned) tiles), block, sh, st>>>(a);
}

template <int KC>
void bu_launch_k(int jt, cudaStream_t st, const FlowBuArgs &a, int tiles) {
    switch (jt) {
        case 2: bu_launch_s<KC, 2>(st, a, tiles); break;
        case 4: bu_launch_s<KC, 4>(st, a, tiles); break;
        case 8: bu_launch_s<KC, 8>(st, a, tiles); break;
        case 16: bu_launch_s<KC, 16>(st, a, tiles); break;
        default: bu_launch_s<KC, 32>(st, a, tiles); break;
    }
}

}  // namespace

void flow_bu_launch(int jt, cudaStream_t st, const FlowBuArgs &a, int tiles) {
    if (a.sched.n_batches <= 0) return;
    static const bool generic = getenv("PASTML_B200_FLOW_GENERIC_K") != nullptr;   // A/B: no compile-time state counts
    if (!generic && jt >= 4 && jt <= 16) {   // the common alphabets get unrolled walks
        switch (a.ln.K) {
            case 20: bu_launch_k<20>(jt, st, a, tiles); return;
            case 10: bu_launch_k<10>(jt, st, a, tiles); return;
            case 4: bu_launch_k<4>(jt, st, a, tiles); return;
            default: break;
        }
    }
    bu_launch_k<0>(jt, st, a, tiles);
}

// ------------------------------------------------------------------------------------------ top-down dispatch
// keep in sync with FLOW_TD_KTS in pastml_b200/build.py
#ifdef PML_ONLY_F81_20   /* development builds (tools/exp_variants.py) */
#define PML_FTD_ALL(X) X(10)
#else
#define PML_FTD_ALL(X) X(1) X(2) X(3) X(4) X(5) X(6) X(8) X(10) X(12) X(16)
#endif
#define PML_FTD_DECL(k) void pml_flow_td_##k(int jt, cudaStream_t, const FlowTdArgs &, int tiles);
PML_FTD_ALL(PML_FTD_DECL)

namespace {
int flow_td_kt(int K) {
#ifdef PML_ONLY_F81_20
    return K == 20 ? 10 : 0;
#else
    const int kts[] = {1, 2, 3, 4, 5, 6, 8, 10, 12, 16};
    for (int kt : kts)
        if (2 * kt >= K) return kt;
    return 0;
#endif
}
}  // namespace

// thread-per-item variant (pml_flow_tdi.cuh): register-array lengths; keep in sync with FLOW_TDI_KPS in build.py
#ifdef PML_ONLY_F81_20
#define PML_FTDI_ALL(X) X(20)
#else
#define PML_FTDI_ALL(X) X(2) X(4) X(8) X(10) X(12) X(16) X(20) X(24) X(32)
#endif
#define PML_FTDI_DECL(k) void pml_flow_tdi_##k(int jt, cudaStream_t, const FlowTdArgs &, int tiles);
PML_FTDI_ALL(PML_FTDI_DECL)

namespace {
int flow_tdi_kp(int K) {
#ifdef PML_ONLY_F81_20
    return K == 20 ? 20 : 0;
#else
    const int kps[] = {2, 4, 8, 10, 12, 16, 20, 24, 32};
    for (int kp : kps)
        if (kp >= K) return kp;
    return 0;
#endif
}
// PASTML_B200_FLOW_TD=g2: the cooperative sweep (two threads per item) instead of thread-per-item (A/B measurements)
bool flow_td_thread_per_item() {
    static const bool g2 = getenv("PASTML_B200_FLOW_TD") && std::string(getenv("PASTML_B200_FLOW_TD")) == "g2";
    return !g2;
}
}  // namespace

bool flow_td_available(int K, int jt) {
    if (!(jt == 2 || jt == 4 || jt == 8 || jt == 16) || K < 2) return false;
    return flow_td_thread_per_item() ? flow_tdi_kp(K) > 0 : flow_td_kt(K) > 0;
}

void flow_td_geometry(int jt, int *npw, int *npar) {
    *npw = flow_td_thread_per_item() ? 32 / jt : 32 / (2 * jt);
    *npar = *npw > 1 ? *npw / 2 : 1;
}

void flow_td_launch(int K, int jt, cudaStream_t st, const FlowTdArgs &a, int tiles) {
    if (a.sched.n_batches <= 0) return;
#define PML_FTD_CASE(k) case k: pml_flow_td_##k(jt, st, a, tiles); return;
#define PML_FTDI_CASE(k) case k: pml_flow_tdi_##k(jt, st, a, tiles); return;
    if (flow_td_thread_per_item()) {
        switch (flow_tdi_kp(K)) {
            PML_FTDI_ALL(PML_FTDI_CASE)
            default: return;
        }
    }
    switch (flow_td_kt(K)) {
        PML_FTD_ALL(PML_FTD_CASE)
        default: return;
    }
}

}  // namespace pml
