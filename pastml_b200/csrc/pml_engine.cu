// Host side of the pastml_b200 engine: problem handles, level-synchronous launch sequences, the C ABI declared
// in include/pastml_b200.h.  Built by nvcc for sm_100a only (see pastml_b200/build.py).
#include <dlfcn.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/pastml_b200.h"
#include "pml_launch.h"

using namespace pml;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(PML_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));               \
    } while (0)

// Device memory is recycled through a small per-process cache: creating and destroying problems of the same
// shape (the end-to-end path, one problem per call) would otherwise spend longer in cudaMalloc / cudaFree than in
// the kernels.  pml_trim_memory() returns the cached blocks to the driver.
class DevicePool {
    std::mutex mu_;
    std::multimap<std::pair<int, size_t>, void *> free_;   // (device, bytes) -> block
    size_t cached_ = 0;
    static constexpr size_t kMaxCached = 48ull << 30;
public:
    // small blocks are rounded up to a power of two so that problems of slightly different shapes still hit the cache
    static size_t rounded(size_t bytes) {
        if (bytes >= (1u << 20)) return bytes;
        size_t r = 512;
        while (r < bytes) r <<= 1;
        return r;
    }
    cudaError_t get(size_t bytes, void **out) {
        bytes = rounded(bytes);
        int dev = 0;
        cudaGetDevice(&dev);
        {
            std::lock_guard<std::mutex> lock(mu_);
            auto it = free_.find({dev, bytes});
            if (it != free_.end()) {
                *out = it->second;
                free_.erase(it);
                cached_ -= bytes;
                return cudaSuccess;
            }
        }
        cudaError_t rc = cudaMalloc(out, bytes);
        if (rc != cudaSuccess) {   // give the cache back and retry once
            trim();
            cudaGetLastError();
            rc = cudaMalloc(out, bytes);
        }
        return rc;
    }
    void put(void *ptr, size_t bytes) {
        int dev = 0;
        cudaGetDevice(&dev);
        bytes = rounded(bytes);
        std::lock_guard<std::mutex> lock(mu_);
        if (cached_ + bytes > kMaxCached) { cudaFree(ptr); return; }
        free_.insert({{dev, bytes}, ptr});
        cached_ += bytes;
    }
    void trim() {
        std::lock_guard<std::mutex> lock(mu_);
        int cur = 0;
        cudaGetDevice(&cur);
        for (auto &kv : free_) { cudaSetDevice(kv.first.first); cudaFree(kv.second); }
        cudaSetDevice(cur);
        free_.clear();
        cached_ = 0;
    }
};
DevicePool g_pool;

template <typename T>
struct DevBuf {
    T *ptr = nullptr;
    size_t count = 0;
    cudaError_t alloc(size_t n) {
        release();
        count = n;
        if (n == 0) return cudaSuccess;
        return g_pool.get(n * sizeof(T), (void **) &ptr);
    }
    void release() {
        if (ptr) g_pool.put(ptr, count * sizeof(T));
        ptr = nullptr;
        count = 0;
    }
    ~DevBuf() { release(); }
};

int kp_for(int K) {
    const int sizes[] = {2, 4, 8, 12, 16, 20, 24, 32, 48, 62};
    for (int s : sizes)
        if (K <= s) return s;
    return -1;
}

}  // namespace

struct pml_problem {
    int device = 0;
    int N = 0, NI = 0, C = 0, K = 0, KP = 0, kind = 0, B = 1, stride = 0, n_levels = 0, n_roots = 0;
    uint64_t kmask = 0;
    std::vector<int> level_ptr, int_level_ptr;
    cudaStream_t stream = nullptr, copy_stream = nullptr;   // copy_stream: bulk uploads / result downloads beside the sweeps
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_copy = nullptr;
    // topology
    DevBuf<int> parent, first_child, n_children, slot, int_nodes, zc_index, roots, fc_slot, nc_slot, pslot;
    DevBuf<int4> bu_desc, td_desc;
    DevBuf<int> d_level_ptr, d_int_level_ptr;   // level boundaries on the device
    int n_root_slots = 0;
    DevBuf<double> dist;
    // zero clusters
    int n_clusters = 0, n_rows = 0;
    DevBuf<int> cl_col, cl_ptr, cl_mem, row_node;
    DevBuf<uint64_t> zc_alt, zc_initial, zc_prev_initial;
    DevBuf<uint8_t> zc_altflag, zc_hasinit, zc_prev_hasinit, zc_k, zc_touched, record_col;
    DevBuf<double> zc_lh;
    // masks [N][C]
    DevBuf<uint64_t> mask[5];
    bool mask_valid[5] = {false, false, false, false, false};
    // state
    DevBuf<double> params, bu, td, probs, node_lh, out_lh;
    DevBuf<int> bu_ex, td_ex, err, joint_state, steps_num, node_ex;
    DevBuf<uint8_t> ptr;
    DevBuf<unsigned long long> k_hist;
    DevBuf<uint64_t> scratch_a, scratch_b, scratch_c;
    DevBuf<long long> steps_out;
    DevBuf<float> flush;
    bool joint_done = false, marginal_done = false;
    // host staging
    std::vector<double> h_out, h_params;   // h_params: the parameter block that is on the device (upload_params)
    std::vector<int> h_err;
    int zero_info[4] = {-1, -1, -1, -1};
    double last_ms = 0.;
    long long last_launches = 0, launches = 0;
    const KernelSet *ks = nullptr;
    // which kernels a launch sequence consists of (name -> launches), reported by pml_last_kernels
    typedef std::map<std::string, std::pair<int, long long>> KernelLog;   // launches, items (slots or nodes)
    struct GraphEntry { cudaGraphExec_t exec; long long launches; KernelLog log; };
    KernelLog *cur_log = nullptr;        // where the enqueue functions count their launches
    KernelLog last_log[2];               // [0] the last bottom-up sequence, [1] the last top-down sequence
    void note(const char *kernel, long long items) {
        if (!cur_log) return;
        auto &e = (*cur_log)[kernel];
        e.first++;
        e.second += items;
    }
    std::map<std::array<long long, 6>, GraphEntry> graphs;

    TreeDev tree() const {
        return TreeDev{N, NI, parent.ptr, first_child.ptr, n_children.ptr, slot.ptr, int_nodes.ptr, dist.ptr,
                       zc_index.ptr, fc_slot.ptr, nc_slot.ptr, pslot.ptr, n_root_slots, bu_desc.ptr, td_desc.ptr};
    }
    // subtree sharding: external tip vectors / root top-down vectors
    DevBuf<int> lane_col;      // explicit column per lane (pml_loglik_lanes), B * C entries, allocated on first use
    bool eig_shared = false;   // EIG: every lane of the uploaded parameters carries the same d, A, A^-1
    // tensor-core top-down sweep (shared decomposition): the products P L_child kept by the bottom-up sweep and the
    // top-down vectors of every node, all indexed by node
    DevBuf<double> v_all, td_all;
    DevBuf<int> td_ex_all;
    bool v_valid = false;      // v_all belongs to the bottom-up state that is in bu / bu_ex
    int n_ext = 0, ext_sets = 0;
    bool have_root_td = false;
    DevBuf<int> ext_index, ext_bu_ex, ext_td_ex, root_td_ex, ext_nodes;
    DevBuf<double> ext_bu, ext_td, root_td;
    ExtDev ext() const {
        return ExtDev{ext_index.ptr, ext_bu.ptr, ext_bu_ex.ptr, ext_td.ptr, ext_td_ex.ptr,
                      have_root_td ? root_td.ptr : nullptr, root_td_ex.ptr};
    }
    ZcDev zc() const { return ZcDev{zc_alt.ptr, zc_altflag.ptr, zc_initial.ptr, zc_hasinit.ptr, zc_touched.ptr}; }
    LaneDev lanes(int n_sets) const {
        return LaneDev{C, n_sets * C, K, stride, params.ptr, kmask, FastDiv::make((unsigned) (n_sets * C)),
                       FastDiv::make((unsigned) C), 0, nullptr};
    }
};

namespace {

const CoopSet *pick_coop(const pml_problem *p, int J, bool wide = false, bool allow_tiled = true);
inline bool tpi_enabled() {   // PASTML_B200_NO_TPI=1: the cooperative kernels instead (A/B measurements)
    static const bool off = getenv("PASTML_B200_NO_TPI") != nullptr;
    return !off;
}
// levels narrower than this (about half a wave of the staged kernel) take the low-latency cooperative bottom-up kernel;
// PASTML_B200_THIN=<nodes> overrides (measured on the 1M-tip workload: 0 -> 0.740 ms, 5000 -> 0.716 ms per sweep)
inline int thin_level_threshold() {
    static const int n = getenv("PASTML_B200_THIN") ? atoi(getenv("PASTML_B200_THIN")) : 5000;
    return n;
}
inline bool tma_enabled() {   // PASTML_B200_NO_TMA=1: the register-path cooperative kernels (A/B measurements)
    static const bool off = getenv("PASTML_B200_NO_TMA") != nullptr;
    return !off;
}

// the top-down sweep can run on the tensor cores: eigen-decomposition model with one decomposition for all the columns
// (JTT), columns a multiple of 8, no external tables (PASTML_B200_NO_MMA / PASTML_B200_NO_MMA_TD: A/B measurements)
bool mma_td_wanted(const pml_problem *p) {
    static const bool off = getenv("PASTML_B200_NO_MMA") != nullptr || getenv("PASTML_B200_NO_MMA_TD") != nullptr;
    return !off && p->kind == PML_MODEL_EIG && p->eig_shared && !p->ext_index.ptr && !p->have_root_td &&
           mma_bu_applicable(p->K, p->C);
}

void run_root_loglik(pml_problem *p, const LaneDev &ln, const uint64_t *mask, int alter, int is_marginal) {
    p->ks->root_loglik(p->stream, p->tree(), p->zc(), ln, mask, alter, p->bu.ptr, p->bu_ex.ptr, p->roots.ptr,
                       p->n_roots, is_marginal, p->out_lh.ptr);
    p->launches++;
}

// zero-branch cluster alteration of `src` masks; record[c] tells whether '.initial' is written for column c
void run_alter(pml_problem *p, const uint64_t *src) {
    if (!p->n_clusters) return;
    alter_kernel<<<grid_for(p->n_clusters), kBlock, 0, p->stream>>>(
        p->n_clusters, p->cl_col.ptr, p->cl_ptr.ptr, p->cl_mem.ptr, p->zc_index.ptr, src, p->C, p->kmask,
        p->record_col.ptr, p->zc_alt.ptr, p->zc_altflag.ptr, p->zc_initial.ptr, p->zc_hasinit.ptr, p->zc_touched.ptr);
    p->launches++;
}

// uploads params [n_sets][C][stride] and the per-column "tau == 0 somewhere" flags
int upload_params(pml_problem *p, const double *params, int n_sets) {
    const size_t n = (size_t) n_sets * p->C * p->stride;
    // the same block as last time (restricted-likelihood passes, repeated sweeps): nothing to copy, nothing to wait for
    if (p->h_params.size() == n && std::memcmp(p->h_params.data(), params, n * sizeof(double)) == 0) return PML_OK;
    p->h_params.assign(params, params + n);
    CUDA_TRY(cudaMemcpyAsync(p->params.ptr, params, n * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    std::vector<uint8_t> rec(p->C, 0);
    for (int b = 0; b < n_sets; ++b)
        for (int c = 0; c < p->C; ++c)
            if (params[((size_t) b * p->C + c) * p->stride + 1] == 0.) rec[c] = 1;
    CUDA_TRY(cudaMemcpyAsync(p->record_col.ptr, rec.data(), p->C, cudaMemcpyHostToDevice, p->stream));
    if (p->kind == PML_MODEL_EIG) {
        // do all the lanes share one eigen-decomposition (JTT with its fixed frequencies)?  Then the products of a
        // node are GEMMs against the same matrices and the bottom-up sweep can run on the tensor cores
        const size_t off = 4 + (size_t) p->K, len = (size_t) p->stride - off;
        bool same = true;
        for (size_t j = 1; same && j < (size_t) n_sets * p->C; ++j)
            same = std::memcmp(params + off, params + j * p->stride + off, len * sizeof(double)) == 0;
        p->eig_shared = same;
    }
    CUDA_TRY(cudaStreamSynchronize(p->stream));   // rec is a stack-local host buffer
    return PML_OK;
}

// the bottom-up sweep on resident params/masks; leaves out_lh on the device
// lanes per tile of the TMA-staged bottom-up sweep for a sum-product pass with J lanes, 0 = other kernels
// (the tile-major layout only for passes with several parameter sets: the top-down sweep, which always runs on one
// set, reads the plain layout)
int bu_tile_lanes(const pml_problem *p, int J) {
    if (p->kind != PML_MODEL_F81 || p->ext_index.ptr || !tpi_enabled()) return 0;
    const int jt = tpi_tile_lanes(p->K, J);
    return (jt == J || J > p->C) ? jt : 0;
}

// n_lanes > 0: the pass runs over an explicit list of n_lanes lanes (p->lane_col) instead of n_sets x all columns
void enqueue_bottom_up(pml_problem *p, int n_sets, const uint64_t *mask, int is_marginal, int alter, bool keep_ptr,
                       int n_lanes = 0) {
    LaneDev ln = p->lanes(n_sets);
    if (n_lanes > 0) {
        ln.J = n_lanes;
        ln.divJ = FastDiv::make((unsigned) n_lanes);
        ln.lane_col = p->lane_col.ptr;
    }
    const int tpi_jt = n_lanes > 0 ? tpi_tile_lanes(p->K, n_lanes) : (is_marginal ? bu_tile_lanes(p, ln.J) : 0);
    ln.jt = tpi_jt;
    if (alter) run_alter(p, mask);
    fill_int_kernel<<<grid_for(2LL * ln.J), kBlock, 0, p->stream>>>(2LL * ln.J, -1, p->err.ptr);
    p->launches++;
    BuArgs a{p->tree(), p->zc(), p->ext(), ln, mask, alter, p->bu.ptr, p->bu_ex.ptr, keep_ptr ? p->ptr.ptr : nullptr,
             p->err.ptr, nullptr};
    const bool maxp = !is_marginal;
    const CoopSet *coop = maxp ? nullptr : pick_coop(p, ln.J);
    static const bool no_mma = getenv("PASTML_B200_NO_MMA") != nullptr;
    const bool mma = !maxp && p->kind == PML_MODEL_EIG && p->eig_shared && !p->ext_index.ptr && !no_mma &&
                     mma_bu_applicable(p->K, ln.J);
    const bool mma_maxp = maxp && p->kind == PML_MODEL_EIG && p->eig_shared && !p->ext_index.ptr && !no_mma &&
                          n_lanes == 0 && mma_maxp_applicable(p->K);
    // the state pass of the marginal pipeline (one set, annotation masks, alter) also keeps P L_child for the top-down sweep
    if (mma && n_sets == 1 && n_lanes == 0 && alter && mask == p->mask[PML_MASK_ANNOTATION].ptr) a.v_all = p->v_all.ptr;
    // thread-per-item sweep with TMA-staged child vectors (F81 family, sum-product mode, no external tips)
    for (int lvl = p->n_levels - 1; lvl >= 0; --lvl) {
        const int s0 = p->int_level_ptr[lvl], ns = p->int_level_ptr[lvl + 1] - s0;
        if (ns <= 0) continue;
        if (mma_maxp) {
            mma_maxp_launch(p->stream, a, s0, ns);
            p->launches++;
            p->note("bu_maxp_mma_kernel", ns);
            continue;
        }
        if (mma) {
            mma_bu_launch(p->stream, a, s0, ns);
            p->launches++;
            p->note("bu_mma_kernel", ns);
            continue;
        }
        if (tpi_jt) {
            // thin levels are latency-bound: the cooperative kernel (a node's states split over a thread group, plain
            // loads) finishes a single wave sooner than the staged thread-per-item kernel
            // (not for lane lists: the cooperative kernels take lane j's column to be j % C)
            const bool thin = coop && !ln.lane_col && tpi_jt == ln.J && ns < thin_level_threshold();
            if (thin) coop->bu(p->stream, a, s0, ns);
            else tpi_bu_launch(tpi_jt, p->stream, a, s0, ns);
            p->launches++;
            p->note(thin ? "bu_coop_kernel" : "bu_tpi_kernel", ns);
            continue;
        }
        if (coop) {
            coop->bu(p->stream, a, s0, ns);
            p->launches++;
            p->note("bu_coop_kernel", ns);
            continue;
        }
        const int chunk = (int) std::max<long long>(1, kMaxItemsPerLaunch / ln.J);
        for (int off = 0; off < ns; off += chunk) {
            p->ks->bu(p->stream, a, maxp, s0 + off, std::min(chunk, ns - off));
            p->launches++;
            p->note("bu_level_kernel", std::min(chunk, ns - off));
        }
    }
    run_root_loglik(p, ln, mask, alter, is_marginal);
}

// hook(node_end): called behind the launches that complete the results of nodes [0, node_end)
using LevelHook = std::function<void(int)>;
void enqueue_top_down(pml_problem *p, int force_joint, bool want_probs, bool want_node_lh, const LevelHook *hook = nullptr);

// Launch sequences are captured once into CUDA graphs (one per distinct combination of arguments) and replayed:
// a sweep is ~55 dependent launches, half of them over levels too narrow to hide the launch latency.
// PASTML_B200_NO_GRAPH=1 launches the kernels directly (ncu captures, debugging).
template <class F>
void run_graphed(pml_problem *p, const std::array<long long, 6> &key, F enqueue) {
    static const bool disabled = getenv("PASTML_B200_NO_GRAPH") != nullptr;
    const int kind = key[0] == 1 ? 1 : 0;
    auto logged = [&](pml_problem::KernelLog &log) {
        log.clear();
        p->cur_log = &log;
        enqueue();
        p->cur_log = nullptr;
    };
    if (disabled) { logged(p->last_log[kind]); return; }
    auto it = p->graphs.find(key);
    if (it == p->graphs.end()) {
        // first use: launched directly (a one-shot call -- create, one pass, destroy -- would pay capture and
        // instantiation for a single replay); the sequence is captured when it comes back
        it = p->graphs.emplace(key, pml_problem::GraphEntry{nullptr, 0, {}}).first;
        logged(it->second.log);
        p->last_log[kind] = it->second.log;
        return;
    }
    if (!it->second.exec) {
        if (it->second.launches < 0) {   // capture failed before: direct launches for this handle
            logged(it->second.log);
            p->last_log[kind] = it->second.log;
            return;
        }
        const long long l0 = p->launches;
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        bool ok = cudaStreamBeginCapture(p->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            logged(it->second.log);
            ok = cudaStreamEndCapture(p->stream, &graph) == cudaSuccess && graph != nullptr;
        }
        if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
        const long long n = p->launches - l0;
        p->launches = l0;
        if (!ok) {
            cudaGetLastError();
            it->second.launches = -1;
            logged(it->second.log);
            p->last_log[kind] = it->second.log;
            return;
        }
        it->second.exec = exec;
        it->second.launches = n;
    }
    cudaGraphLaunch(it->second.exec, p->stream);
    p->launches += it->second.launches;
    p->last_log[kind] = it->second.log;
}

void run_bottom_up(pml_problem *p, int n_sets, const uint64_t *mask, int is_marginal, int alter, bool keep_ptr,
                   int n_lanes = 0) {
    // everything that selects kernels at capture time belongs to the key: eig_shared decides between the tensor-core
    // sweep (one decomposition for all lanes) and the per-lane kernels, and it changes when the frequencies do
    run_graphed(p, {n_lanes > 0 ? 2 : 0, n_lanes > 0 ? n_lanes : n_sets, (long long) (uintptr_t) mask, is_marginal, alter,
                    (keep_ptr ? 1 : 0) + 2 * (p->eig_shared ? 1 : 0) + 4 * (p->v_all.ptr ? 1 : 0) + 8 * (long long) p->n_ext},
                [&]() { enqueue_bottom_up(p, n_sets, mask, is_marginal, alter, keep_ptr, n_lanes); });
    // does v_all (the products P L_child kept by the tensor-core sweep) still go with the vectors in bu?
    const bool state_pass = is_marginal && n_sets == 1 && n_lanes == 0 && alter && mask == p->mask[PML_MASK_ANNOTATION].ptr;
    p->v_valid = state_pass && p->v_all.ptr && mma_td_wanted(p);
}

void run_top_down(pml_problem *p, int force_joint, bool want_probs, bool want_node_lh) {
    const bool mma = mma_td_wanted(p) && p->v_valid;
    run_graphed(p, {1, force_joint, want_probs ? 1 : 0, want_node_lh ? 1 : 0, p->n_ext, (p->have_root_td ? 1 : 0) + (mma ? 2 : 0)},
                [&]() { enqueue_top_down(p, force_joint, want_probs, want_node_lh); });
}

int fetch_loglik(pml_problem *p, int n_sets, double *out, int n_lanes = 0) {
    const int J = n_lanes > 0 ? n_lanes : n_sets * p->C;
    p->h_out.resize(J);
    p->h_err.resize(2 * (size_t) J);
    CUDA_TRY(cudaMemcpyAsync(p->h_out.data(), p->out_lh.ptr, J * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaMemcpyAsync(p->h_err.data(), p->err.ptr, 2 * (size_t) J * sizeof(int), cudaMemcpyDeviceToHost,
                             p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    int rc = PML_OK;
    for (int j = 0; j < J; ++j) {
        if (p->h_err[2 * j] >= 0) {
            if (rc == PML_OK) {
                p->zero_info[0] = j % p->C;
                p->zero_info[1] = j / p->C;
                p->zero_info[2] = p->h_err[2 * j];
                p->zero_info[3] = p->h_err[2 * j + 1];
            }
            rc = PML_ERR_ZERO_LIKELIHOOD;
            p->h_out[j] = -INFINITY;
        }
        if (out) out[j] = p->h_out[j];
    }
    if (rc != PML_OK) g_last_error = "zero likelihood: a parent and a child connected by a zero-length branch have "
                                     "non-intersecting states";
    return rc;
}

struct CallTimer {
    pml_problem *p;
    long long l0;
    explicit CallTimer(pml_problem *pp) : p(pp), l0(pp->launches) { cudaEventRecord(p->ev0, p->stream); }
    void stop() {
        cudaEventRecord(p->ev1, p->stream);
        cudaEventSynchronize(p->ev1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, p->ev0, p->ev1);
        p->last_ms = ms;
        p->last_launches = p->launches - l0;
    }
};

int ensure_marginal_buffers(pml_problem *p, bool want_probs, bool want_node_lh) {
    const size_t nik = (size_t) p->NI * p->K * p->C;
    if (!p->td.ptr && nik) CUDA_TRY(p->td.alloc(nik));
    if (!p->td_ex.ptr && p->NI) CUDA_TRY(p->td_ex.alloc((size_t) p->NI * p->C));
    if (want_probs && !p->probs.ptr) CUDA_TRY(p->probs.alloc((size_t) p->N * p->C * p->K));
    if (want_node_lh && !p->node_lh.ptr) CUDA_TRY(p->node_lh.alloc((size_t) p->N * p->C));
    if (want_node_lh && !p->node_ex.ptr) CUDA_TRY(p->node_ex.alloc((size_t) p->N * p->C));
    if (!p->k_hist.ptr) CUDA_TRY(p->k_hist.alloc((size_t) p->C * (p->K + 1)));
    if (!p->mask[PML_MASK_MAP].ptr) CUDA_TRY(p->mask[PML_MASK_MAP].alloc((size_t) p->N * p->C));
    if (!p->mask[PML_MASK_MPPA].ptr) CUDA_TRY(p->mask[PML_MASK_MPPA].alloc((size_t) p->N * p->C));
    return PML_OK;
}

// the fused top-down sweep on resident bottom-up state
void enqueue_top_down(pml_problem *p, int force_joint, bool want_probs, bool want_node_lh, const LevelHook *hook) {
    const LaneDev ln = p->lanes(1);
    cudaMemsetAsync(p->k_hist.ptr, 0, sizeof(unsigned long long) * p->C * (p->K + 1), p->stream);
    TdArgs a{p->tree(), p->zc(), p->ext(), ln, p->mask[PML_MASK_ANNOTATION].ptr, p->bu.ptr, p->bu_ex.ptr, p->td.ptr,
             p->td_ex.ptr, p->joint_state.ptr, force_joint, want_probs ? p->probs.ptr : nullptr,
             want_node_lh ? p->node_lh.ptr : nullptr, p->node_ex.ptr, p->mask[PML_MASK_MAP].ptr, p->mask[PML_MASK_MPPA].ptr,
             p->k_hist.ptr, p->zc_lh.ptr, p->zc_k.ptr};
    if (mma_td_wanted(p) && p->v_valid) {
        // vectors level by level on the tensor cores, then marginals / MAP / MPPA of the whole forest in one launch
        for (int lvl = 1; lvl < p->n_levels; ++lvl) {
            const int n0 = p->level_ptr[lvl], nn = p->level_ptr[lvl + 1] - n0;
            mma_td_launch(p->stream, a, p->v_all.ptr, p->td_all.ptr, p->td_ex_all.ptr, n0, nn);
            p->launches++;
            p->note("td_mma_kernel", nn);
        }
        p->ks->td_select(p->stream, a, p->td_all.ptr, p->td_ex_all.ptr, p->N);
        p->launches++;
        p->note("td_select_kernel", p->N);
        if (hook) (*hook)(p->N);
        return;
    }
    const bool tma_ok = !p->ext_index.ptr && tma_enabled();   // (top-down vectors arriving at the roots are fine)
    const CoopSet *coop = pick_coop(p, ln.C, tma_ok && !getenv("PASTML_B200_COOP_G"), false);
    const bool td_tma = coop && coop->td_tma && !p->ext_index.ptr && tma_enabled();
    a.raw_node_lh = coop ? 1 : 0;   // one finishing pass takes the logarithms when the cooperative kernels run the sweep
    for (int lvl = 0; lvl < p->n_levels; ++lvl) {
        const int n0 = p->level_ptr[lvl], nn = p->level_ptr[lvl + 1] - n0;
        if (coop) {
            (td_tma ? coop->td_tma : coop->td)(p->stream, a, n0, nn);
            p->launches++;
            p->note("td_coop_kernel", nn);
            if (hook) (*hook)(n0 + nn);
            continue;
        }
        const int chunk = (int) std::max<long long>(1, kMaxItemsPerLaunch / ln.C);
        for (int off = 0; off < nn; off += chunk) {
            p->ks->td(p->stream, a, n0 + off, std::min(chunk, nn - off));
            p->launches++;
            p->note("td_level_kernel", std::min(chunk, nn - off));
        }
        if (hook) (*hook)(n0 + nn);
    }
    if (coop && want_node_lh) {
        const long long n = (long long) p->N * p->C;
        launch_pdl(log10_finish_kernel, dim3((unsigned) std::min<long long>(148 * 8, (n / 2 + 255) / 256 + 1)), dim3(256), 0, p->stream, n,
                   p->node_lh.ptr, (const int *) p->node_ex.ptr);
        p->launches++;
    }
}

int ensure_mma_td_buffers(pml_problem *p) {
    if (!mma_td_wanted(p) || p->v_all.ptr) return PML_OK;
    const size_t n = (size_t) p->N * p->K * p->C;
    CUDA_TRY(p->v_all.alloc(n));
    CUDA_TRY(p->td_all.alloc(n));
    CUDA_TRY(p->td_ex_all.alloc((size_t) p->N * p->C));
    return PML_OK;
}

int ensure_joint_buffers(pml_problem *p);

// choose_ancestral_states_joint level by level on the pointers the max-product sweep left; root_state: see JointArgs
void enqueue_joint_trace(pml_problem *p, const int *root_state) {
    JointArgs a{p->tree(), p->zc(), p->lanes(1), p->mask[PML_MASK_ANNOTATION].ptr, p->bu.ptr, p->ptr.ptr,
                p->joint_state.ptr, p->mask[PML_MASK_JOINT].ptr, root_state};
    for (int lvl = 0; lvl < p->n_levels; ++lvl) {
        const int n0 = p->level_ptr[lvl], nn = p->level_ptr[lvl + 1] - n0;
        launch_pdl(joint_trace_kernel, dim3(grid_for((long long) nn * p->C)), dim3(kBlock), 0, p->stream, a, n0, nn);
        p->launches++;
    }
}

int ensure_joint_buffers(pml_problem *p) {
    const size_t nc = (size_t) p->N * p->C;
    if (!p->ptr.ptr) CUDA_TRY(p->ptr.alloc(nc * p->K));
    if (!p->joint_state.ptr) CUDA_TRY(p->joint_state.alloc(nc));
    if (!p->mask[PML_MASK_JOINT].ptr) CUDA_TRY(p->mask[PML_MASK_JOINT].alloc(nc));
    return PML_OK;
}

int copy_masks_out(pml_problem *p, const uint64_t *dev, uint64_t *host) {
    const long long n = (long long) p->N * p->C;
    if (!p->scratch_a.ptr) CUDA_TRY(p->scratch_a.alloc(n));
    strip_flag_kernel<<<grid_for(n), kBlock, 0, p->stream>>>(n, p->kmask, dev, p->scratch_a.ptr);
    p->launches++;
    CUDA_TRY(cudaMemcpyAsync(host, p->scratch_a.ptr, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    return PML_OK;
}

int copy_khist_out(pml_problem *p, int64_t *out) {
    CUDA_TRY(cudaMemcpyAsync(out, p->k_hist.ptr, sizeof(int64_t) * p->C * (p->K + 1), cudaMemcpyDeviceToHost,
                             p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    return PML_OK;
}

}  // namespace

namespace pml {
#define PML_DECL(f, k) const KernelSet *pml_ks_##f##_##k();
#define PML_CASE(f, k) if (family == f && kp == k) return pml_ks_##f##_##k();
#ifdef PML_ONLY_F81_20   /* development builds (tools/exp_variants.py) */
#define PML_ALL(X) X(0, 20)
#else
#define PML_ALL(X) X(0, 2) X(0, 4) X(0, 8) X(0, 12) X(0, 16) X(0, 20) X(0, 24) X(0, 32) X(0, 48) X(0, 62) \
                   X(1, 4) X(2, 2) X(2, 4) X(2, 8) X(2, 12) X(2, 16) X(2, 20) X(2, 24) X(2, 32) X(2, 48) X(2, 62)
#endif
PML_ALL(PML_DECL)
const KernelSet *kernel_set(int family, int kp) {
    PML_ALL(PML_CASE)
    return nullptr;
}
}  // namespace pml

namespace pml {
#define PML_CDECL(g, k, j) const CoopSet *pml_coop_##g##_##k##_##j();
#define PML_CCASE(g, k, j) if (gg == g && kt == k && jt == j) return pml_coop_##g##_##k##_##j();
#ifdef PML_ONLY_F81_20
#define PML_COOP_ALL(X) X(4, 5, 0) X(4, 5, 8) X(2, 10, 8)
#else
#define PML_COOP_ALL(X) X(2, 2, 0) X(2, 3, 0) X(2, 4, 0) X(2, 5, 0) X(2, 6, 0) X(2, 8, 0) X(2, 10, 0) X(2, 12, 0) X(2, 16, 0) \
                        X(4, 2, 0) X(4, 3, 0) X(4, 4, 0) X(4, 5, 0) X(4, 6, 0) X(4, 8, 0) X(4, 10, 0) X(4, 12, 0) X(4, 16, 0) \
                        X(8, 2, 0) X(8, 3, 0) X(8, 4, 0) X(8, 5, 0) X(8, 6, 0) X(8, 8, 0) X(16, 2, 0) X(16, 3, 0) X(16, 4, 0) \
                        X(32, 2, 0) X(4, 5, 8) X(2, 5, 16) X(2, 5, 8) X(4, 5, 4) X(2, 2, 16) X(2, 2, 8) X(2, 10, 8) \
                        X(4, 3, 8) X(4, 4, 8) X(4, 2, 8) X(2, 3, 8) X(2, 4, 8) X(4, 6, 8) X(4, 8, 8) X(2, 6, 8)
#endif
PML_COOP_ALL(PML_CDECL)
const CoopSet *coop_set(int gg, int kt, int jt) {
    PML_COOP_ALL(PML_CCASE)
    return nullptr;
}
}  // namespace pml

namespace {
// Picks the cooperative configuration for K states and J lanes: G * J <= 32 threads per node, KT = the smallest
// built size >= ceil(K / G); least padding first, then the larger group.  nullptr -> thread-per-item kernels.
const CoopSet *pick_coop(const pml_problem *p, int J, bool wide, bool allow_tiled) {
    static const bool disabled = getenv("PASTML_B200_NO_COOP") != nullptr;
    if (disabled || p->kind != PML_MODEL_F81 || p->K < 3) return nullptr;
    if (J > 16) {
        // many lanes (parameter sets x columns): tiles of 8 lanes through the compile-time variants, when they fit
        // (bottom-up sweep only: the top-down kernels take all the columns of a node in one warp batch)
        if (!allow_tiled || J % 8) return nullptr;
        for (int g = 4; g >= 2; g >>= 1)
            if (p->K % g == 0) {
                const CoopSet *cs = coop_set(g, p->K / g, 8);
                if (cs) return cs;
            }
        return nullptr;
    }
    if (wide && p->K % 2 == 0 && p->K / 2 <= 10) {
        // the TMA-staged top-down sweep runs best with two threads per item (measured: K = 20, 2.50 against 2.59 ms)
        const CoopSet *cs = coop_set(2, p->K / 2, J);
        if (cs && cs->td_tma) return cs;
    }
    const int kts[] = {2, 3, 4, 5, 6, 8, 10, 12, 16};
    const CoopSet *best = nullptr;
    double best_score = 1e30;
    static const int forced_g = getenv("PASTML_B200_COOP_G") ? atoi(getenv("PASTML_B200_COOP_G")) : 0;   // tuning aid
    for (int g = 2; g <= 32 && g * J <= 32; g <<= 1) {
        if (forced_g && g != forced_g) continue;
        const int need = (p->K + g - 1) / g;
        for (int kt : kts) {
            if (kt < need) continue;
            const CoopSet *cs = coop_set(g, kt, 0);
            if (g * kt == p->K) {   // a variant with compile-time lane count and state count, when built
                const CoopSet *exact = coop_set(g, kt, J);
                if (exact) cs = exact;
            }
            if (!cs) break;
            const double score = (double) (g * kt - p->K) / p->K + (kt < 3 ? 0.5 : 0.) + (kt > 8 ? 0.25 : 0.) - 1e-3 * g;
            if (score < best_score) { best_score = score; best = cs; }
            break;
        }
    }
    return best;
}
}  // namespace

// ====================================================================================== C ABI

extern "C" {

const char *pml_last_error(void) { return g_last_error.c_str(); }

int pml_version(void) { return 100; }

int pml_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int pml_param_stride(int32_t model_kind, int32_t K) {
    if (K < 1) return -1;
    switch (model_kind) {
        case PML_MODEL_F81: return 4 + K;
        case PML_MODEL_HKY: return 8;
        case PML_MODEL_EIG: return 4 + 2 * K + 2 * K * K;
        default: return -1;
    }
}

int pml_problem_create(int32_t n_nodes, const int32_t *parent, const double *dist, int32_t n_columns,
                       int32_t n_states, int32_t model_kind, const uint64_t *annot_mask, const uint8_t *has_state,
                       int32_t max_param_sets, int32_t device, pml_problem **out) {
    if (!out) return fail(PML_ERR_INVALID, "out is null");
    *out = nullptr;
    if (n_nodes < 1 || !parent || !dist || !annot_mask || !has_state)
        return fail(PML_ERR_INVALID, "empty forest or null input arrays");
    if (n_columns < 1 || max_param_sets < 1) return fail(PML_ERR_INVALID, "need at least one column and one set");
    if (n_states < 1 || n_states > 62) return fail(PML_ERR_INVALID, "the number of states must be in [1, 62]");
    if (model_kind == PML_MODEL_HKY && n_states != 4) return fail(PML_ERR_INVALID, "HKY needs exactly 4 states");
    if (model_kind < 0 || model_kind > PML_MODEL_EIG) return fail(PML_ERR_INVALID, "unknown model family");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return fail(PML_ERR_NO_DEVICE, "no CUDA device available: pastml_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(PML_ERR_INVALID, "bad device ordinal");
    CUDA_TRY(cudaSetDevice(device));

    const int N = n_nodes, C = n_columns, K = n_states;
    const size_t nc = (size_t) N * C;
    pml_problem *p = new pml_problem();
    p->device = device;
#define PML_TRY_FREE(expr)                                                                               \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            std::string m = std::string(#expr) + ": " + cudaGetErrorString(_e);                          \
            pml_problem_destroy(p);                                                                      \
            return fail(PML_ERR_CUDA, m);                                                                \
        }                                                                                                \
    } while (0)
#define PML_FAIL_FREE(msg)                                                                               \
    do {                                                                                                 \
        pml_problem_destroy(p);                                                                          \
        return fail(PML_ERR_INVALID, msg);                                                               \
    } while (0)
    PML_TRY_FREE(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    PML_TRY_FREE(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
    PML_TRY_FREE(cudaEventCreateWithFlags(&p->ev_copy, cudaEventDisableTiming));
    PML_TRY_FREE(cudaEventCreate(&p->ev0));
    PML_TRY_FREE(cudaEventCreate(&p->ev1));

    // ---- the annotation masks ([C][N], by far the largest input) start moving first, on the copy stream; the
    // topology is derived on the device meanwhile
    DevBuf<uint64_t> tmp_mask;
    PML_TRY_FREE(tmp_mask.alloc(nc));
    PML_TRY_FREE(p->parent.alloc(N));
    PML_TRY_FREE(p->dist.alloc(N));
    PML_TRY_FREE(cudaMemcpyAsync(p->parent.ptr, parent, sizeof(int) * (size_t) N, cudaMemcpyHostToDevice, p->stream));
    PML_TRY_FREE(cudaMemcpyAsync(p->dist.ptr, dist, sizeof(double) * (size_t) N, cudaMemcpyHostToDevice, p->stream));
    PML_TRY_FREE(cudaEventRecord(p->ev_copy, p->stream));
    PML_TRY_FREE(cudaStreamWaitEvent(p->copy_stream, p->ev_copy, 0));   // (keeps the two small uploads in front)
    PML_TRY_FREE(cudaMemcpyAsync(tmp_mask.ptr, annot_mask, nc * sizeof(uint64_t), cudaMemcpyHostToDevice, p->copy_stream));

    // ---- topology (level order: parents non-decreasing and before their children): children counts, level-order
    // checks and the slot / first-child prefix sums on the device (pml_topo.cu)
    DevBuf<unsigned long long> scan;
    DevBuf<uint8_t> scan_tmp;
    DevBuf<int> flags_ni;
    DevBuf<int> &d_level_ptr = p->d_level_ptr, &d_int_level_ptr = p->d_int_level_ptr;
    const size_t scan_bytes = topo_scan_bytes(N);
    PML_TRY_FREE(scan.alloc(N));
    PML_TRY_FREE(scan_tmp.alloc(std::max<size_t>(scan_bytes, 1)));
    PML_TRY_FREE(flags_ni.alloc(2));
    PML_TRY_FREE(p->n_children.alloc(N));
    topo_stage1(p->stream, N, p->parent.ptr, p->dist.ptr, p->n_children.ptr, scan.ptr, scan_tmp.ptr, scan_bytes, flags_ni.ptr);
    int h_flags_ni[2] = {0, 0};
    PML_TRY_FREE(cudaMemcpyAsync(h_flags_ni, flags_ni.ptr, sizeof(h_flags_ni), cudaMemcpyDeviceToHost, p->stream));
    // (the level boundaries need the host copy of parent[] only: binary searches, done while the device works)
    const int n_roots = (int) (std::lower_bound(parent, parent + N, 0) - parent);
    std::vector<int> level_ptr{0, n_roots};
    bool disconnected = false;
    while (n_roots >= 1 && level_ptr.back() < N && (int) level_ptr.size() <= N + 1) {
        // the next level ends where parents stop falling into the current one: parent[] is sorted
        const int hi = level_ptr.back();
        const int nxt = (int) (std::lower_bound(parent + hi, parent + N, hi) - parent);
        if (nxt == hi) { disconnected = true; break; }
        level_ptr.push_back(nxt);
    }
    PML_TRY_FREE(cudaStreamSynchronize(p->stream));
    if (h_flags_ni[0] & kTopoNotBeforeChild) PML_FAIL_FREE("nodes are not in level order (a parent must precede its children)");
    if (h_flags_ni[0] & kTopoNotSorted) PML_FAIL_FREE("nodes are not in level order (parents must be non-decreasing)");
    if (n_roots < 1) PML_FAIL_FREE("roots must come first in level order");
    if (disconnected) PML_FAIL_FREE("disconnected nodes");
    const bool any_zero = (h_flags_ni[0] & kTopoZeroBranch) != 0;
    const int NI = h_flags_ni[1];
    const int n_levels = (int) level_ptr.size() - 1;
    std::vector<int> roots(n_roots);
    for (int r = 0; r < n_roots; ++r) roots[r] = r;
    std::vector<int> int_level_ptr(n_levels + 1, 0);
    PML_TRY_FREE(p->first_child.alloc(N));
    PML_TRY_FREE(p->slot.alloc(N));
    PML_TRY_FREE(p->pslot.alloc(N));
    PML_TRY_FREE(p->zc_index.alloc(N));
    PML_TRY_FREE(p->td_desc.alloc(N));
    PML_TRY_FREE(p->int_nodes.alloc(NI));
    PML_TRY_FREE(p->fc_slot.alloc(NI));
    PML_TRY_FREE(p->nc_slot.alloc(NI));
    PML_TRY_FREE(p->bu_desc.alloc(2 * (size_t) NI));
    PML_TRY_FREE(d_level_ptr.alloc(n_levels + 1));
    PML_TRY_FREE(d_int_level_ptr.alloc(n_levels + 1));
    PML_TRY_FREE(cudaMemcpyAsync(d_level_ptr.ptr, level_ptr.data(), sizeof(int) * (n_levels + 1), cudaMemcpyHostToDevice, p->stream));
    {
        TopoOut o{p->n_children.ptr, p->first_child.ptr, p->slot.ptr, p->pslot.ptr, p->zc_index.ptr,
                  p->int_nodes.ptr, p->fc_slot.ptr, p->nc_slot.ptr, p->bu_desc.ptr, p->td_desc.ptr};
        topo_stage2(p->stream, N, n_roots, NI, p->parent.ptr, p->dist.ptr, scan.ptr, o, d_level_ptr.ptr, n_levels + 1,
                    d_int_level_ptr.ptr);
    }
    PML_TRY_FREE(cudaMemcpyAsync(int_level_ptr.data(), d_int_level_ptr.ptr, sizeof(int) * (n_levels + 1), cudaMemcpyDeviceToHost, p->stream));
    PML_TRY_FREE(cudaStreamSynchronize(p->stream));

    // ---- zero-branch clusters per column (get_zero_clusters_with_states, ml.py:321-349)
    std::vector<int> cl_col, cl_ptr{0}, cl_mem, row_node;
    std::vector<uint8_t> member;
    std::vector<int> zc_index;
    if (any_zero) {
        zc_index.assign(N, -1);
        std::vector<int> top(N);
        for (int v = 0; v < N; ++v) {
            const int q = parent[v];
            top[v] = (q >= 0 && dist[v] == 0.) ? top[q] : v;
        }
        member.assign((size_t) C * N, 0);
        std::vector<int> cnt(N), start(N + 1), fill(N), order;
        for (int c = 0; c < C; ++c) {
            const uint8_t *hs = has_state + (size_t) c * N;
            std::fill(cnt.begin(), cnt.end(), 0);
            for (int v = 0; v < N; ++v)
                if (hs[v]) cnt[top[v]]++;
            start[0] = 0;
            for (int v = 0; v < N; ++v) start[v + 1] = start[v] + cnt[v];
            order.assign(start[N], 0);
            std::fill(fill.begin(), fill.end(), 0);
            for (int v = 0; v < N; ++v)
                if (hs[v]) order[start[top[v]] + fill[top[v]]++] = v;
            for (int t = 0; t < N; ++t) {
                if (cnt[t] < 2) continue;
                cl_col.push_back(c);
                for (int i = start[t]; i < start[t + 1]; ++i) {
                    const int v = order[i];
                    cl_mem.push_back(v);
                    member[(size_t) c * N + v] = 1;
                    if (zc_index[v] < 0) { zc_index[v] = (int) row_node.size(); row_node.push_back(v); }
                }
                cl_ptr.push_back((int) cl_mem.size());
            }
        }
    }

    p->N = N; p->NI = NI; p->C = C; p->K = K; p->kind = model_kind; p->B = max_param_sets;
    p->KP = (model_kind == PML_MODEL_HKY) ? 4 : kp_for(K);
    p->ks = kernel_set(model_kind, p->KP);
    if (!p->ks) PML_FAIL_FREE("no kernels built for this model family / number of states");
    p->stride = pml_param_stride(model_kind, K);
    p->kmask = (K >= 64) ? ~0ULL : ((1ULL << K) - 1ULL);
    p->n_levels = n_levels;
    p->n_roots = (int) roots.size();
    p->level_ptr = level_ptr;
    p->int_level_ptr = int_level_ptr;
    p->n_clusters = (int) cl_col.size();
    p->n_rows = (int) row_node.size();

#define PML_UPLOAD(buf, vec)                                                                             \
    do {                                                                                                 \
        PML_TRY_FREE((buf).alloc((vec).size()));                                                         \
        if (!(vec).empty())                                                                              \
            PML_TRY_FREE(cudaMemcpy((buf).ptr, (vec).data(), (vec).size() * sizeof((vec)[0]), cudaMemcpyHostToDevice)); \
    } while (0)
    if (any_zero)
        PML_TRY_FREE(cudaMemcpy(p->zc_index.ptr, zc_index.data(), sizeof(int) * (size_t) N, cudaMemcpyHostToDevice));
    p->n_root_slots = int_level_ptr[1];
    PML_UPLOAD(p->roots, roots);
    PML_UPLOAD(p->cl_col, cl_col);
    PML_UPLOAD(p->cl_ptr, cl_ptr);
    PML_UPLOAD(p->cl_mem, cl_mem);
    PML_UPLOAD(p->row_node, row_node);
    const size_t rc = (size_t) std::max(p->n_rows, 1) * C;
    PML_TRY_FREE(p->zc_alt.alloc(rc));
    PML_TRY_FREE(p->zc_initial.alloc(rc));
    PML_TRY_FREE(p->zc_prev_initial.alloc(rc));
    PML_TRY_FREE(p->zc_altflag.alloc(rc));
    PML_TRY_FREE(p->zc_hasinit.alloc(rc));
    PML_TRY_FREE(p->zc_prev_hasinit.alloc(rc));
    PML_TRY_FREE(p->zc_k.alloc(rc));
    PML_TRY_FREE(p->zc_touched.alloc(rc));
    PML_TRY_FREE(cudaMemsetAsync(p->zc_touched.ptr, 0, rc, p->stream));
    PML_TRY_FREE(p->zc_lh.alloc(rc * K));
    PML_TRY_FREE(p->record_col.alloc(C));
    PML_TRY_FREE(cudaMemsetAsync(p->zc_alt.ptr, 0, rc * sizeof(uint64_t), p->stream));
    PML_TRY_FREE(cudaMemsetAsync(p->zc_initial.ptr, 0, rc * sizeof(uint64_t), p->stream));
    PML_TRY_FREE(cudaMemsetAsync(p->zc_altflag.ptr, 0, rc, p->stream));
    PML_TRY_FREE(cudaMemsetAsync(p->zc_hasinit.ptr, 0, rc, p->stream));
    PML_TRY_FREE(cudaMemsetAsync(p->zc_k.ptr, 0, rc, p->stream));
    PML_TRY_FREE(cudaMemsetAsync(p->zc_lh.ptr, 0, rc * K * sizeof(double), p->stream));

    PML_TRY_FREE(p->mask[PML_MASK_ANNOTATION].alloc(nc));   // the other mask sets are allocated when first produced
    {
        // [C][N] -> [N][C] on the device, behind the upload on the copy stream
        DevBuf<uint8_t> tmp_mem;
        if (!member.empty()) {
            PML_TRY_FREE(tmp_mem.alloc(nc));
            PML_TRY_FREE(cudaMemcpyAsync(tmp_mem.ptr, member.data(), nc, cudaMemcpyHostToDevice, p->copy_stream));
        }
        init_masks_kernel<<<grid_for((long long) nc), kBlock, 0, p->copy_stream>>>(N, C, p->kmask, tmp_mask.ptr, tmp_mem.ptr,
                                                                                   p->mask[PML_MASK_ANNOTATION].ptr, 1);
        PML_TRY_FREE(cudaStreamSynchronize(p->copy_stream));
        p->mask_valid[PML_MASK_ANNOTATION] = true;
    }
    const size_t J = (size_t) p->B * C;
    PML_TRY_FREE(p->params.alloc(J * p->stride));
    PML_TRY_FREE(p->bu.alloc(std::max<size_t>((size_t) NI * K * J, 1)));
    PML_TRY_FREE(p->bu_ex.alloc(std::max<size_t>((size_t) NI * J, 1)));
    PML_TRY_FREE(p->err.alloc(2 * J));
    PML_TRY_FREE(p->out_lh.alloc(J));
#undef PML_UPLOAD
#undef PML_TRY_FREE
#undef PML_FAIL_FREE
    *out = p;
    return PML_OK;
}

void pml_problem_destroy(pml_problem *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    for (auto &kv : p->graphs)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    p->graphs.clear();
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    if (p->ev_copy) cudaEventDestroy(p->ev_copy);
    if (p->copy_stream) { cudaStreamSynchronize(p->copy_stream); cudaStreamDestroy(p->copy_stream); }
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
}

int pml_loglik(pml_problem *p, const double *params, int32_t n_sets, int32_t mask_kind, int32_t is_marginal,
               int32_t alter, double *out_loglh) {
    if (!p || !params) return fail(PML_ERR_INVALID, "null argument");
    if (n_sets < 1 || n_sets > p->B) return fail(PML_ERR_INVALID, "n_sets exceeds the capacity given at creation");
    if (mask_kind < 0 || mask_kind > PML_MASK_CUSTOM || !p->mask_valid[mask_kind])
        return fail(PML_ERR_INVALID, "the requested mask set has not been produced yet");
    CUDA_TRY(cudaSetDevice(p->device));
    int rc = upload_params(p, params, n_sets);
    if (rc) return rc;
    CallTimer timer(p);
    run_bottom_up(p, n_sets, p->mask[mask_kind].ptr, is_marginal, alter, false);
    timer.stop();
    CUDA_TRY(cudaGetLastError());
    return fetch_loglik(p, n_sets, out_loglh);
}

int pml_supports_lane_lists(pml_problem *p) {
    return p && p->kind == PML_MODEL_F81 && !p->ext_index.ptr && tpi_enabled() && p->K >= 2;
}

int pml_loglik_lanes(pml_problem *p, const double *params, const int32_t *lane_columns, int32_t n_lanes,
                     double *out_loglh) {
    if (!p || !params || !lane_columns) return fail(PML_ERR_INVALID, "null argument");
    if (!pml_supports_lane_lists(p)) return fail(PML_ERR_INVALID, "lane lists need the TMA-staged bottom-up sweep (F81 family)");
    if (n_lanes < 1 || n_lanes > p->B * p->C) return fail(PML_ERR_INVALID, "n_lanes exceeds the capacity given at creation");
    if (!tpi_tile_lanes(p->K, n_lanes)) return fail(PML_ERR_INVALID, "the number of lanes must be a multiple of 4");
    CUDA_TRY(cudaSetDevice(p->device));
    if (!p->lane_col.ptr) CUDA_TRY(p->lane_col.alloc((size_t) p->B * p->C));
    std::vector<uint8_t> rec(p->C, 0);
    for (int j = 0; j < n_lanes; ++j) {
        const int c = lane_columns[j];
        if (c < 0 || c >= p->C) return fail(PML_ERR_INVALID, "lane column out of range");
        if (params[(size_t) j * p->stride + 1] == 0.) rec[c] = 1;
    }
    p->h_params.clear();   // the lane-list layout replaces whatever upload_params left on the device
    CUDA_TRY(cudaMemcpyAsync(p->params.ptr, params, (size_t) n_lanes * p->stride * sizeof(double), cudaMemcpyHostToDevice,
                             p->stream));
    CUDA_TRY(cudaMemcpyAsync(p->lane_col.ptr, lane_columns, (size_t) n_lanes * sizeof(int), cudaMemcpyHostToDevice, p->stream));
    CUDA_TRY(cudaMemcpyAsync(p->record_col.ptr, rec.data(), p->C, cudaMemcpyHostToDevice, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));   // rec is a stack-local host buffer
    CallTimer timer(p);
    run_bottom_up(p, 1, p->mask[PML_MASK_ANNOTATION].ptr, 1, 1, false, n_lanes);
    timer.stop();
    CUDA_TRY(cudaGetLastError());
    const int rc = fetch_loglik(p, 1, out_loglh, n_lanes);
    if (rc == PML_ERR_ZERO_LIKELIHOOD) {   // (column, lane) of the first failing lane
        const int j = p->zero_info[1] * p->C + p->zero_info[0];
        p->zero_info[0] = lane_columns[j];
        p->zero_info[1] = j;
    }
    return rc;
}

int pml_joint(pml_problem *p, const double *params, double *out_loglh, int32_t *out_state) {
    if (!p || !params) return fail(PML_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(p->device));
    int rc = ensure_joint_buffers(p);
    if (rc) return rc;
    rc = upload_params(p, params, 1);
    if (rc) return rc;
    CallTimer timer(p);
    const uint64_t *base = p->mask[PML_MASK_ANNOTATION].ptr;
    run_bottom_up(p, 1, base, 0, 1, true);
    enqueue_joint_trace(p, nullptr);
    timer.stop();
    CUDA_TRY(cudaGetLastError());
    p->mask_valid[PML_MASK_JOINT] = true;
    p->joint_done = true;
    rc = fetch_loglik(p, 1, out_loglh);
    if (out_state) {
        CUDA_TRY(cudaMemcpyAsync(out_state, p->joint_state.ptr, sizeof(int) * (size_t) p->N * p->C,
                                 cudaMemcpyDeviceToHost, p->stream));
        CUDA_TRY(cudaStreamSynchronize(p->stream));
    }
    return rc;
}

int pml_marginal(pml_problem *p, const double *params, int32_t force_joint, double *out_loglh, double *out_probs,
                 double *out_node_log10lh, uint64_t *out_map_mask, uint64_t *out_mppa_mask, int64_t *out_k_hist) {
    if (!p || !params) return fail(PML_ERR_INVALID, "null argument");
    // bit 1 of force_joint (PML_REUSE_BOTTOM_UP): the bottom-up state left by the previous pml_loglik call (one
    // parameter set, annotation masks, marginal, alter) is still current -- skip recomputing it
    const bool reuse_bu = (force_joint & 2) != 0;
    // bit 2 (PML_RESULTS_ON_DEVICE): probabilities and node log-likelihoods are computed and kept in the handle's
    // device buffers even though no host pointer is given (device-resident timing of the full sweep)
    const bool on_device = (force_joint & 4) != 0;
    force_joint &= 1;
    if (force_joint && !p->joint_done)
        return fail(PML_ERR_INVALID, "force_joint needs the JOINT states: call pml_joint first");
    CUDA_TRY(cudaSetDevice(p->device));
    const bool do_probs = out_probs != nullptr || on_device, do_node_lh = out_node_log10lh != nullptr || on_device;
    int rc = ensure_marginal_buffers(p, do_probs, do_node_lh);
    if (rc) return rc;
    rc = upload_params(p, params, 1);
    if (rc) return rc;
    rc = ensure_mma_td_buffers(p);
    if (rc) return rc;
    CallTimer timer(p);
    // ml.py:700-706: allowed states are re-initialised and altered, then BU(alter=False), TD, marginals
    // (a reused bottom-up state must come with the products the tensor-core top-down sweep reads)
    if (!reuse_bu || (mma_td_wanted(p) && !p->v_valid)) run_bottom_up(p, 1, p->mask[PML_MASK_ANNOTATION].ptr, 1, 1, false);
    // Large results leave while the sweep is still running: a level's probabilities and state sets are final when
    // its launch retires, so they are handed to the copy stream level by level (in pieces of >= 8 MB) instead of
    // after the last level.  (Not with zero-branch clusters or external tips: their masks carry marker bits that are
    // stripped on the way out, and clusters are revisited after the sweep.)
    const size_t nc = (size_t) p->N * p->C;
    static const bool no_stream_out = getenv("PASTML_B200_NO_STREAM_OUT") != nullptr;
    const bool stream_out = (out_probs || out_map_mask || out_mppa_mask) && p->n_rows == 0 && p->n_ext == 0 &&
                            !p->ext_index.ptr && nc * p->K * sizeof(double) >= (32u << 20) && !no_stream_out;
    cudaError_t copy_rc = cudaSuccess;
    if (stream_out) {
        size_t done = 0;
        const size_t row = (size_t) p->C, piece = std::max<size_t>(1, (8u << 20) / (row * p->K * sizeof(double)));
        const LevelHook hook = [&](int node_end) {
            if ((size_t) node_end - done < piece && node_end < p->N) return;
            const size_t n = ((size_t) node_end - done) * row, at = done * row;
            auto keep = [&](cudaError_t e) { if (copy_rc == cudaSuccess) copy_rc = e; };
            keep(cudaEventRecord(p->ev_copy, p->stream));
            keep(cudaStreamWaitEvent(p->copy_stream, p->ev_copy, 0));
            if (out_probs)
                keep(cudaMemcpyAsync(out_probs + at * p->K, p->probs.ptr + at * p->K, n * p->K * sizeof(double),
                                     cudaMemcpyDeviceToHost, p->copy_stream));
            if (out_map_mask)
                keep(cudaMemcpyAsync(out_map_mask + at, p->mask[PML_MASK_MAP].ptr + at, n * sizeof(uint64_t),
                                     cudaMemcpyDeviceToHost, p->copy_stream));
            if (out_mppa_mask)
                keep(cudaMemcpyAsync(out_mppa_mask + at, p->mask[PML_MASK_MPPA].ptr + at, n * sizeof(uint64_t),
                                     cudaMemcpyDeviceToHost, p->copy_stream));
            done = (size_t) node_end;
        };
        p->last_log[1].clear();
        p->cur_log = &p->last_log[1];
        enqueue_top_down(p, force_joint, do_probs, do_node_lh, &hook);
        p->cur_log = nullptr;
    } else {
        run_top_down(p, force_joint, do_probs, do_node_lh);
    }
    // remember the '.initial' state MAP saw: pml_mppa_finalize needs it
    if (p->n_rows) {
        cudaMemcpyAsync(p->zc_prev_initial.ptr, p->zc_initial.ptr, sizeof(uint64_t) * (size_t) p->n_rows * p->C,
                        cudaMemcpyDeviceToDevice, p->stream);
        cudaMemcpyAsync(p->zc_prev_hasinit.ptr, p->zc_hasinit.ptr, (size_t) p->n_rows * p->C,
                        cudaMemcpyDeviceToDevice, p->stream);
        // from here on, alterations that rewrite '.initial' are what pml_mppa_finalize has to replay
        cudaMemsetAsync(p->zc_touched.ptr, 0, (size_t) p->n_rows * p->C, p->stream);
    }
    timer.stop();
    CUDA_TRY(cudaGetLastError());
    p->mask_valid[PML_MASK_MAP] = p->mask_valid[PML_MASK_MPPA] = true;
    p->marginal_done = true;
    rc = fetch_loglik(p, 1, out_loglh);
    if (out_probs && !stream_out)
        CUDA_TRY(cudaMemcpyAsync(out_probs, p->probs.ptr, nc * p->K * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    if (out_node_log10lh)
        CUDA_TRY(cudaMemcpyAsync(out_node_log10lh, p->node_lh.ptr, nc * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    if (stream_out) {   // (before anything below may return: the copy stream writes into the caller's buffers)
        const cudaError_t e = cudaStreamSynchronize(p->copy_stream);
        if (copy_rc == cudaSuccess) copy_rc = e;
    }
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    CUDA_TRY(copy_rc);
    if (!stream_out) {
        if (out_map_mask) { int r2 = copy_masks_out(p, p->mask[PML_MASK_MAP].ptr, out_map_mask); if (r2) return r2; }
        if (out_mppa_mask) { int r2 = copy_masks_out(p, p->mask[PML_MASK_MPPA].ptr, out_mppa_mask); if (r2) return r2; }
    }
    if (out_k_hist) { int r2 = copy_khist_out(p, out_k_hist); if (r2) return r2; }
    return rc;
}

int pml_marginal_counts(pml_problem *p, const double *params, int32_t n_repetitions, uint64_t seed, double *out_counts) {
    if (!p || !params || !out_counts) return fail(PML_ERR_INVALID, "null argument");
    if (n_repetitions < 1) return fail(PML_ERR_INVALID, "need at least one repetition");
    if (p->ext_index.ptr) return fail(PML_ERR_INVALID, "scenario sampling is not available on a problem with external tips");
    CUDA_TRY(cudaSetDevice(p->device));
    int rc = upload_params(p, params, 1);
    if (rc) return rc;
    // ml.py:786-791: allowed states re-initialised and altered, bottom-up likelihoods (the top-down vector of a root
    // is all ones, and the conditional draws below need no other top-down vector)
    run_bottom_up(p, 1, p->mask[PML_MASK_ANNOTATION].ptr, 1, 1, false);
    rc = fetch_loglik(p, 1, nullptr);
    if (rc) return rc;
    const int K = p->K, C = p->C;
    size_t width = 1;
    for (int l = 0; l < p->n_levels; ++l) width = std::max<size_t>(width, (size_t) (p->int_level_ptr[l + 1] - p->int_level_ptr[l]));
    DevBuf<uint8_t> st[2];
    DevBuf<double> result;
    CUDA_TRY(st[0].alloc(width * C * (size_t) n_repetitions));
    CUDA_TRY(st[1].alloc(width * C * (size_t) n_repetitions));
    CUDA_TRY(result.alloc((size_t) C * K * K));
    CUDA_TRY(cudaMemsetAsync(result.ptr, 0, sizeof(double) * (size_t) C * K * K, p->stream));
    const LaneDev ln = p->lanes(1);
    CountsArgs a{p->tree(), p->zc(), p->ext(), ln, p->mask[PML_MASK_ANNOTATION].ptr, p->bu.ptr, p->bu_ex.ptr,
                 n_repetitions, (unsigned long long) seed, nullptr, st[0].ptr, 0, p->int_level_ptr[1], 0, result.ptr};
    if (a.n_slots > 0) {
        p->ks->counts(p->stream, a, true);
        p->launches++;
    }
    int cur = 0;
    for (int l = 0; l < p->n_levels; ++l) {
        const int s0 = p->int_level_ptr[l], ns = p->int_level_ptr[l + 1] - s0;
        if (ns <= 0) continue;
        a.st_cur = st[cur].ptr;
        a.st_next = st[cur ^ 1].ptr;
        a.slot_begin = s0;
        a.n_slots = ns;
        a.next_slot_begin = p->int_level_ptr[l + 1];
        p->ks->counts(p->stream, a, false);
        p->launches++;
        cur ^= 1;
    }
    CUDA_TRY(cudaMemcpyAsync(out_counts, result.ptr, sizeof(double) * (size_t) C * K * K, cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    CUDA_TRY(cudaGetLastError());
    for (size_t i = 0; i < (size_t) C * K * K; ++i) out_counts[i] /= (double) n_repetitions;
    return PML_OK;
}

int pml_mppa_finalize(pml_problem *p, int32_t force_joint, uint64_t *out_mppa_mask, int64_t *out_k_hist) {
    if (!p) return fail(PML_ERR_INVALID, "null argument");
    if (!p->marginal_done) return fail(PML_ERR_INVALID, "call pml_marginal first");
    CUDA_TRY(cudaSetDevice(p->device));
    CallTimer timer(p);
    if (p->n_rows) {
        p->ks->mppa_fix(p->stream, p->n_rows, p->row_node.ptr, p->C, p->K, p->kmask, p->zc(),
                        p->mask[PML_MASK_ANNOTATION].ptr, p->zc_prev_initial.ptr, p->zc_prev_hasinit.ptr,
                        p->zc_lh.ptr, p->zc_k.ptr, p->joint_state.ptr, force_joint, p->mask[PML_MASK_MPPA].ptr,
                        p->k_hist.ptr);
        p->launches++;
    }
    timer.stop();
    CUDA_TRY(cudaGetLastError());
    if (out_mppa_mask) { int r2 = copy_masks_out(p, p->mask[PML_MASK_MPPA].ptr, out_mppa_mask); if (r2) return r2; }
    if (out_k_hist) { int r2 = copy_khist_out(p, out_k_hist); if (r2) return r2; }
    return PML_OK;
}

int pml_parsimony(pml_problem *p, int32_t methods, uint64_t *out_mask, int64_t *out_steps) {
    if (!p) return fail(PML_ERR_INVALID, "null argument");
    if (!(methods & (PML_PARS_ACCTRAN | PML_PARS_DOWNPASS | PML_PARS_DELTRAN)))
        return fail(PML_ERR_INVALID, "no parsimony method requested");
    CUDA_TRY(cudaSetDevice(p->device));
    const int C = p->C, K = p->K;
    const long long nc = (long long) p->N * C;
    if (!p->scratch_a.ptr) CUDA_TRY(p->scratch_a.alloc(nc));
    if (!p->scratch_b.ptr) CUDA_TRY(p->scratch_b.alloc(nc));
    if (!p->scratch_c.ptr) CUDA_TRY(p->scratch_c.alloc(nc));
    if (!p->steps_num.ptr) CUDA_TRY(p->steps_num.alloc(std::max<size_t>((size_t) p->NI * K * C, 1)));
    if (!p->steps_out.ptr) CUDA_TRY(p->steps_out.alloc(C));
    uint64_t *down = p->scratch_a.ptr, *res = p->scratch_b.ptr, *up = p->scratch_c.ptr;
    const TreeDev t = p->tree();
    CallTimer timer(p);
    // initialise_parsimonious_states (parsimony.py:51-73): annotation sets, all states where there is none
    strip_flag_kernel<<<grid_for(nc), kBlock, 0, p->stream>>>(nc, p->kmask, p->mask[PML_MASK_ANNOTATION].ptr, down);
    p->launches++;
    for (int lvl = p->n_levels - 1; lvl >= 0; --lvl) {
        const int s0 = p->int_level_ptr[lvl], ns = p->int_level_ptr[lvl + 1] - s0;
        if (ns <= 0) continue;
        launch_pdl(pars_uppass_kernel, dim3(grid_for((long long) ns * C)), dim3(kBlock), 0, p->stream, t, C, p->kmask, down, s0, ns);
        p->launches++;
    }
    int slot_out = 0;
    auto steps_of = [&](const uint64_t *states, int64_t *dst) -> int {
        for (int lvl = p->n_levels - 1; lvl >= 0; --lvl) {
            const int s0 = p->int_level_ptr[lvl], ns = p->int_level_ptr[lvl + 1] - s0;
            if (ns <= 0) continue;
            launch_pdl(pars_steps_kernel, dim3(grid_for((long long) ns * C)), dim3(kBlock), 0, p->stream, t, C, K, states, p->steps_num.ptr, s0, ns);
            p->launches++;
        }
        pars_steps_root_kernel<<<grid_for(C), kBlock, 0, p->stream>>>(t, C, K, states, p->steps_num.ptr, p->roots.ptr,
                                                                    p->n_roots, p->steps_out.ptr);
        p->launches++;
        if (dst) {
            CUDA_TRY(cudaMemcpyAsync(dst, p->steps_out.ptr, sizeof(long long) * C, cudaMemcpyDeviceToHost, p->stream));
            CUDA_TRY(cudaStreamSynchronize(p->stream));
        }
        return PML_OK;
    };
    auto emit = [&](const uint64_t *states) -> int {
        int rc = steps_of(states, out_steps ? out_steps + (size_t) slot_out * C : nullptr);
        if (rc) return rc;
        if (out_mask) {
            CUDA_TRY(cudaMemcpyAsync(out_mask + (size_t) slot_out * nc, states, nc * sizeof(uint64_t),
                                     cudaMemcpyDeviceToHost, p->stream));
            CUDA_TRY(cudaStreamSynchronize(p->stream));
        }
        ++slot_out;
        return PML_OK;
    };
    if (methods & PML_PARS_ACCTRAN) {
        // roots take their uppass sets (parsimony.py:152-153)
        CUDA_TRY(cudaMemcpyAsync(res, down, sizeof(uint64_t) * (size_t) p->n_roots * C, cudaMemcpyDeviceToDevice,
                                 p->stream));
        for (int lvl = 0; lvl < p->n_levels; ++lvl) {
            const int s0 = p->int_level_ptr[lvl], ns = p->int_level_ptr[lvl + 1] - s0;
            if (ns <= 0) continue;
            launch_pdl(pars_acctran_kernel, dim3(grid_for((long long) ns * C)), dim3(kBlock), 0, p->stream, t, C, (const uint64_t *) down, res, s0, ns);
            p->launches++;
        }
        int rc = emit(res);
        if (rc) return rc;
    }
    if (methods & (PML_PARS_DOWNPASS | PML_PARS_DELTRAN)) {
        // the PARSIMONY feature starts as the pre-uppass sets (parsimony.py:72-73)
        strip_flag_kernel<<<grid_for(nc), kBlock, 0, p->stream>>>(nc, p->kmask, p->mask[PML_MASK_ANNOTATION].ptr, res);
        p->launches++;
        for (int lvl = 0; lvl < p->n_levels; ++lvl) {
            const int s0 = p->int_level_ptr[lvl], ns = p->int_level_ptr[lvl + 1] - s0;
            if (ns <= 0) continue;
            launch_pdl(pars_downpass_kernel, dim3(grid_for((long long) ns * C)), dim3(kBlock), 0, p->stream, t, C, p->kmask, (const uint64_t *) down, up, res, s0, ns);
            p->launches++;
        }
        if (methods & PML_PARS_DOWNPASS) {
            int rc = emit(res);
            if (rc) return rc;
        }
        if (methods & PML_PARS_DELTRAN) {
            for (int lvl = 0; lvl < p->n_levels; ++lvl) {
                const int s0 = p->int_level_ptr[lvl], ns = p->int_level_ptr[lvl + 1] - s0;
                if (ns <= 0) continue;
                launch_pdl(pars_deltran_kernel, dim3(grid_for((long long) ns * C)), dim3(kBlock), 0, p->stream, t, C, res, s0, ns);
                p->launches++;
            }
            int rc = emit(res);
            if (rc) return rc;
        }
    }
    timer.stop();
    CUDA_TRY(cudaGetLastError());
    return PML_OK;
}

int pml_set_custom_masks(pml_problem *p, const uint64_t *mask) {
    if (!p || !mask) return fail(PML_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(p->device));
    const size_t nc = (size_t) p->N * p->C;
    if (!p->scratch_a.ptr) CUDA_TRY(p->scratch_a.alloc(nc));
    if (!p->scratch_b.ptr) CUDA_TRY(p->scratch_b.alloc(nc));
    if (!p->mask[PML_MASK_CUSTOM].ptr) CUDA_TRY(p->mask[PML_MASK_CUSTOM].alloc(nc));
    CUDA_TRY(cudaMemcpyAsync(p->scratch_a.ptr, mask, nc * sizeof(uint64_t), cudaMemcpyHostToDevice, p->stream));
    // member flags are taken from the annotation masks: [C][N] -> [N][C], then OR the flags in
    init_masks_kernel<<<grid_for((long long) nc), kBlock, 0, p->stream>>>(p->N, p->C, p->kmask, p->scratch_a.ptr, nullptr,
                                                                          p->scratch_b.ptr, 0);
    merge_flag_kernel<<<grid_for((long long) nc), kBlock, 0, p->stream>>>((long long) nc, p->mask[PML_MASK_ANNOTATION].ptr,
                                                                          p->scratch_b.ptr, p->mask[PML_MASK_CUSTOM].ptr);
    p->launches += 2;
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    CUDA_TRY(cudaGetLastError());
    p->mask_valid[PML_MASK_CUSTOM] = true;
    return PML_OK;
}

int pml_get_masks(pml_problem *p, int32_t mask_kind, uint64_t *out_mask) {
    if (!p || !out_mask) return fail(PML_ERR_INVALID, "null argument");
    if (mask_kind < 0 || mask_kind > PML_MASK_CUSTOM || !p->mask_valid[mask_kind])
        return fail(PML_ERR_INVALID, "the requested mask set has not been produced yet");
    CUDA_TRY(cudaSetDevice(p->device));
    return copy_masks_out(p, p->mask[mask_kind].ptr, out_mask);
}

int pml_get_initial(pml_problem *p, uint8_t *out_has, uint64_t *out_initial) {
    if (!p || !out_has || !out_initial) return fail(PML_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(p->device));
    const size_t nc = (size_t) p->N * p->C;
    std::memset(out_has, 0, nc);
    std::memset(out_initial, 0, nc * sizeof(uint64_t));
    if (!p->n_rows) return PML_OK;
    const size_t rc = (size_t) p->n_rows * p->C;
    std::vector<uint64_t> init(rc);
    std::vector<uint8_t> has(rc);
    std::vector<int> rows(p->n_rows);
    CUDA_TRY(cudaMemcpy(init.data(), p->zc_initial.ptr, rc * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(has.data(), p->zc_hasinit.ptr, rc, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(rows.data(), p->row_node.ptr, sizeof(int) * p->n_rows, cudaMemcpyDeviceToHost));
    for (int r = 0; r < p->n_rows; ++r)
        for (int c = 0; c < p->C; ++c)
            if (has[(size_t) r * p->C + c]) {
                out_has[(size_t) rows[r] * p->C + c] = 1;
                out_initial[(size_t) rows[r] * p->C + c] = init[(size_t) r * p->C + c];
            }
    return PML_OK;
}

int pml_get_bottom_up(pml_problem *p, double *out_vec, int32_t *out_ex, uint64_t *out_allowed, uint8_t *out_altered) {
    if (!p || !out_vec || !out_ex || !out_allowed || !out_altered) return fail(PML_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(p->device));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    const size_t nc = (size_t) p->N * p->C;
    if (p->NI) {
        CUDA_TRY(cudaMemcpy(out_vec, p->bu.ptr, sizeof(double) * (size_t) p->NI * p->K * p->C, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(out_ex, p->bu_ex.ptr, sizeof(int) * (size_t) p->NI * p->C, cudaMemcpyDeviceToHost));
    }
    CUDA_TRY(cudaMemcpy(out_allowed, p->mask[PML_MASK_ANNOTATION].ptr, nc * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < nc; ++i) out_allowed[i] &= p->kmask;
    std::memset(out_altered, 0, nc);
    if (!p->n_rows) return PML_OK;
    const size_t rc = (size_t) p->n_rows * p->C;
    std::vector<uint64_t> alt(rc);
    std::vector<uint8_t> flag(rc);
    std::vector<int> rows(p->n_rows);
    CUDA_TRY(cudaMemcpy(alt.data(), p->zc_alt.ptr, rc * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(flag.data(), p->zc_altflag.ptr, rc, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(rows.data(), p->row_node.ptr, sizeof(int) * p->n_rows, cudaMemcpyDeviceToHost));
    for (int r = 0; r < p->n_rows; ++r)
        for (int c = 0; c < p->C; ++c)
            if (flag[(size_t) r * p->C + c]) {
                out_altered[(size_t) rows[r] * p->C + c] = 1;
                out_allowed[(size_t) rows[r] * p->C + c] = alt[(size_t) r * p->C + c] & p->kmask;
            }
    return PML_OK;
}

int pml_zero_likelihood_info(pml_problem *p, int32_t *column, int32_t *param_set, int32_t *parent, int32_t *child) {
    if (!p) return fail(PML_ERR_INVALID, "null argument");
    if (column) *column = p->zero_info[0];
    if (param_set) *param_set = p->zero_info[1];
    if (parent) *parent = p->zero_info[2];
    if (child) *child = p->zero_info[3];
    return PML_OK;
}

int pml_zero_likelihood_all(pml_problem *p, int32_t *out_pairs, int32_t n_lanes) {
    if (!p || !out_pairs) return fail(PML_ERR_INVALID, "null argument");
    for (int j = 0; j < n_lanes; ++j) {
        const bool have = (size_t) (2 * j + 1) < p->h_err.size();
        out_pairs[2 * j] = have ? p->h_err[2 * j] : -1;
        out_pairs[2 * j + 1] = have ? p->h_err[2 * j + 1] : -1;
    }
    return PML_OK;
}

int pml_get_root_partials(pml_problem *p, int32_t n_sets, double *out_vec, int32_t *out_ex) {
    if (!p || !out_vec || !out_ex) return fail(PML_ERR_INVALID, "null argument");
    if (n_sets < 1 || n_sets > p->B) return fail(PML_ERR_INVALID, "bad n_sets");
    CUDA_TRY(cudaSetDevice(p->device));
    LaneDev ln = p->lanes(n_sets);
    ln.jt = bu_tile_lanes(p, ln.J);   // the layout the last sum-product pass with this many lanes left behind
    const size_t nv = (size_t) p->n_roots * p->K * ln.J, ne = (size_t) p->n_roots * ln.J;
    DevBuf<double> vec;
    DevBuf<int> ex;
    CUDA_TRY(vec.alloc(nv));
    CUDA_TRY(ex.alloc(ne));
    root_partials_kernel<<<grid_for((long long) ne), kBlock, 0, p->stream>>>(
        p->tree(), p->zc(), ln, p->mask[PML_MASK_ANNOTATION].ptr, 1, p->bu.ptr, p->bu_ex.ptr, p->n_roots, vec.ptr, ex.ptr);
    p->launches++;
    CUDA_TRY(cudaMemcpyAsync(out_vec, vec.ptr, nv * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaMemcpyAsync(out_ex, ex.ptr, ne * sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    CUDA_TRY(cudaGetLastError());
    return PML_OK;
}

int pml_set_tip_partials(pml_problem *p, int32_t n, const int32_t *nodes, const double *vec, const int32_t *ex,
                         int32_t n_sets) {
    if (!p || !nodes || !vec || !ex || n < 1) return fail(PML_ERR_INVALID, "bad argument");
    if (n_sets < 1 || n_sets > p->B) return fail(PML_ERR_INVALID, "bad n_sets");
    CUDA_TRY(cudaSetDevice(p->device));
    const int J = n_sets * p->C;
    if (p->n_ext != n) {
        // first call (or a different cut): validate and mark the tips
        std::vector<int> slot_host(p->N);
        CUDA_TRY(cudaMemcpy(slot_host.data(), p->slot.ptr, sizeof(int) * p->N, cudaMemcpyDeviceToHost));
        for (int i = 0; i < n; ++i)
            if (nodes[i] < 0 || nodes[i] >= p->N || slot_host[nodes[i]] >= 0)
                return fail(PML_ERR_INVALID, "external partials can only be attached to tips");
        if (p->n_ext) return fail(PML_ERR_INVALID, "the set of external tips cannot change on a handle");
        CUDA_TRY(p->ext_index.alloc(p->N));
        CUDA_TRY(cudaMemsetAsync(p->ext_index.ptr, 0xff, sizeof(int) * p->N, p->stream));
        CUDA_TRY(p->ext_nodes.alloc(n));
        CUDA_TRY(cudaMemcpyAsync(p->ext_nodes.ptr, nodes, sizeof(int) * n, cudaMemcpyHostToDevice, p->stream));
        mark_ext_kernel<<<grid_for((long long) n * p->C), kBlock, 0, p->stream>>>(
            n, p->ext_nodes.ptr, p->C, p->mask[PML_MASK_ANNOTATION].ptr, p->ext_index.ptr);
        p->launches++;
        CUDA_TRY(p->ext_td.alloc((size_t) n * p->K * p->C));
        CUDA_TRY(p->ext_td_ex.alloc((size_t) n * p->C));
        p->n_ext = n;
    }
    if (!p->ext_bu.ptr) {
        // sized for the handle's capacity of parameter sets once, so that the tables never move (the captured
        // launch graphs keep pointing at them) whatever n_sets the following calls use
        CUDA_TRY(p->ext_bu.alloc((size_t) n * p->K * p->B * p->C));
        CUDA_TRY(p->ext_bu_ex.alloc((size_t) n * p->B * p->C));
    }
    p->ext_sets = n_sets;
    CUDA_TRY(cudaMemcpyAsync(p->ext_bu.ptr, vec, sizeof(double) * (size_t) n * p->K * J, cudaMemcpyHostToDevice, p->stream));
    CUDA_TRY(cudaMemcpyAsync(p->ext_bu_ex.ptr, ex, sizeof(int) * (size_t) n * J, cudaMemcpyHostToDevice, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    CUDA_TRY(cudaGetLastError());
    return PML_OK;
}

int pml_get_tip_td(pml_problem *p, double *out_vec, int32_t *out_ex) {
    if (!p || !out_vec || !out_ex) return fail(PML_ERR_INVALID, "null argument");
    if (!p->n_ext || !p->marginal_done) return fail(PML_ERR_INVALID, "needs external tips and a marginal sweep");
    CUDA_TRY(cudaSetDevice(p->device));
    CUDA_TRY(cudaMemcpyAsync(out_vec, p->ext_td.ptr, sizeof(double) * (size_t) p->n_ext * p->K * p->C,
                             cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaMemcpyAsync(out_ex, p->ext_td_ex.ptr, sizeof(int) * (size_t) p->n_ext * p->C, cudaMemcpyDeviceToHost,
                             p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    return PML_OK;
}

int pml_set_root_td(pml_problem *p, const double *vec, const int32_t *ex) {
    if (!p || !vec || !ex) return fail(PML_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(p->device));
    const size_t nv = (size_t) p->n_roots * p->K * p->C, ne = (size_t) p->n_roots * p->C;
    if (!p->root_td.ptr) {
        CUDA_TRY(p->root_td.alloc(nv));
        CUDA_TRY(p->root_td_ex.alloc(ne));
    }
    CUDA_TRY(cudaMemcpyAsync(p->root_td.ptr, vec, nv * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    CUDA_TRY(cudaMemcpyAsync(p->root_td_ex.ptr, ex, ne * sizeof(int), cudaMemcpyHostToDevice, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    p->have_root_td = true;
    return PML_OK;
}

}  // extern "C" (the sharding block below opens its own)

// ---------------------------------------------------------------------------------- device-resident subtree sharding
//
// One rank's share of a tree cut by subtree: its clade forest and the replicated top tree on one
// device, with the exchange kept on the device.  Per sweep: clades bottom-up -> pack the clade-root runs -> ALL-GATHER
// (NCCL on the clade problem's stream; the library is taken from the process at run time, so a single-GPU user never
// needs it) -> unpack into the top tree's external-tip tables -> top tree (bottom-up, log-likelihood on every rank,
// top-down) -> the rank's own clade roots receive their top-down vectors -> clades top-down.  Nothing goes through
// the host inside a step; the two problems' streams are ordered with events.

namespace {

// the handful of NCCL entry points, resolved with dlopen (signatures of nccl.h; ncclFloat64 = 8)
struct NcclApi {
    struct UniqueId { char internal[128]; };
    int (*GetUniqueId)(UniqueId *) = nullptr;
    int (*CommInitRank)(void **, int, UniqueId, int) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
    std::string why;
};
NcclApi &nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, []() {
        void *h = nullptr;
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) { api.why = "libnccl.so.2 cannot be loaded (import torch first, or put NCCL on the library path)"; return; }
        api.GetUniqueId = (int (*)(NcclApi::UniqueId *)) dlsym(h, "ncclGetUniqueId");
        api.CommInitRank = (int (*)(void **, int, NcclApi::UniqueId, int)) dlsym(h, "ncclCommInitRank");
        api.AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t)) dlsym(h, "ncclAllGather");
        api.CommDestroy = (int (*)(void *)) dlsym(h, "ncclCommDestroy");
        api.GetErrorString = (const char *(*)(int)) dlsym(h, "ncclGetErrorString");
        api.ok = api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy && api.GetErrorString;
        if (!api.ok) api.why = "NCCL symbols missing";
    });
    return api;
}

// clade-root runs of this rank, [root][K + 1][J]: K mantissa rows and the exponents (as doubles, exact)
__global__ void shard_pack_kernel(TreeDev t, LaneDev ln, const double *bu, const int *bu_ex, int n_roots, double *out) {
    const long long g = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const int K = ln.K, J = ln.J;
    if (g >= (long long) n_roots * (K + 1) * J) return;
    const int j = (int) (g % J), k = (int) ((g / J) % (K + 1)), root = (int) (g / ((long long) J * (K + 1)));
    const int s = t.slot[root];   // clade roots are internal nodes
    out[g] = (k < K) ? bu[bu_at(ln, s, k, j)] : (double) bu_ex[(size_t) s * J + j];
}

// gathered runs -> the top tree's external-tip tables (rows in boundary order)
__global__ void shard_unpack_kernel(const double *gathered, const int *row_src, int n_rows, int K, int J, double *ext_bu,
                                    int *ext_bu_ex) {
    const long long g = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long) n_rows * (K + 1) * J) return;
    const int j = (int) (g % J), k = (int) ((g / J) % (K + 1)), row = (int) (g / ((long long) J * (K + 1)));
    const double v = gathered[((size_t) row_src[row] * (K + 1) + k) * J + j];
    if (k < K) ext_bu[((size_t) row * K + k) * J + j] = v;
    else ext_bu_ex[(size_t) row * J + j] = (int) v;
}

// top-down vectors the top tree computed for this rank's clade roots -> the clade forest's roots
__global__ void shard_root_td_kernel(const double *ext_td, const int *ext_td_ex, const int *my_rows, int n_mine, int K, int C,
                                     double *root_td, int *root_td_ex) {
    const long long g = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long) n_mine * (K + 1) * C) return;
    const int c = (int) (g % C), k = (int) ((g / C) % (K + 1)), i = (int) (g / ((long long) C * (K + 1)));
    const int row = my_rows[i];
    if (k < K) root_td[((size_t) i * K + k) * C + c] = ext_td[((size_t) row * K + k) * C + c];
    else root_td_ex[(size_t) i * C + c] = ext_td_ex[(size_t) row * C + c];
}

// joint states the top tree chose for this rank's clade roots -> preset root states of the clade forest
__global__ void shard_root_state_kernel(const int *top_state, const int *my_top_nodes, int n_mine, int C, int *root_state) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_mine * C) return;
    const int i = g / C, c = g - i * C;
    root_state[g] = top_state[(size_t) my_top_nodes[i] * C + c];
}

}  // namespace

struct pml_shard {
    pml_problem *clades = nullptr, *top = nullptr;
    int rank = 0, world = 1, n_clades = 0, n_mine = 0, n_max = 0;
    DevBuf<int> row_src, my_rows, my_top_nodes, root_state;   // my_top_nodes: this rank's clade roots as nodes of the top tree
    DevBuf<double> xsend, xbuf;
    size_t slot_doubles = 0;   // capacity of one rank's part of the exchange: n_max * (K + 1) * B * C
    cudaEvent_t ev_gather = nullptr, ev_top = nullptr;
    void *comm = nullptr;
    int last_sets = 0, last_mask_kind = 0;
    cudaEvent_t phase_event = nullptr;   // benchmark instrumentation: recorded between the pack and the collective
};

extern "C" {

int pml_shard_create(pml_problem *clades, pml_problem *top, int32_t rank, int32_t world, int32_t n_clades,
                     const int32_t *boundary_top, const int32_t *clade_owner, pml_shard **out) {
    if (!out) return fail(PML_ERR_INVALID, "out is null");
    *out = nullptr;
    if (!clades || !top || !boundary_top || !clade_owner || world < 1 || rank < 0 || rank >= world || n_clades < world)
        return fail(PML_ERR_INVALID, "bad shard description");
    if (clades->device != top->device || clades->K != top->K || clades->C != top->C || clades->B != top->B ||
        clades->kind != top->kind)
        return fail(PML_ERR_INVALID, "the clade forest and the top tree must share device, states, columns and capacity");
    CUDA_TRY(cudaSetDevice(top->device));
    std::vector<int> count(world, 0), local(n_clades), mine;
    for (int i = 0; i < n_clades; ++i) {
        const int r = clade_owner[i];
        if (r < 0 || r >= world) return fail(PML_ERR_INVALID, "clade owner out of range");
        local[i] = count[r]++;
        if (r == rank) mine.push_back(i);
    }
    const int n_max = *std::max_element(count.begin(), count.end());
    if ((int) mine.size() != clades->n_roots)
        return fail(PML_ERR_INVALID, "the clade forest must have one root per clade of this rank, in boundary order");
    std::vector<int> row_src(n_clades);
    for (int i = 0; i < n_clades; ++i) row_src[i] = clade_owner[i] * n_max + local[i];
    // the top tree's external tips: allocate and mark (zero vectors for now)
    {
        const size_t J = (size_t) top->B * top->C;
        std::vector<double> zv((size_t) n_clades * top->K * J, 0.);
        std::vector<int32_t> ze((size_t) n_clades * J, 0);
        const int rc = pml_set_tip_partials(top, n_clades, boundary_top, zv.data(), ze.data(), top->B);
        if (rc) return rc;
    }
    pml_shard *sh = new pml_shard();
    sh->clades = clades; sh->top = top; sh->rank = rank; sh->world = world;
    sh->n_clades = n_clades; sh->n_mine = (int) mine.size(); sh->n_max = n_max;
    sh->slot_doubles = (size_t) n_max * (top->K + 1) * top->B * top->C;
#define PML_SH_TRY(expr)                                                                                 \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            std::string m = std::string(#expr) + ": " + cudaGetErrorString(_e);                          \
            pml_shard_destroy(sh);                                                                       \
            return fail(PML_ERR_CUDA, m);                                                                \
        }                                                                                                \
    } while (0)
    PML_SH_TRY(sh->row_src.alloc(n_clades));
    PML_SH_TRY(sh->my_rows.alloc(std::max<size_t>(mine.size(), 1)));
    PML_SH_TRY(sh->xsend.alloc(sh->slot_doubles));
    PML_SH_TRY(sh->xbuf.alloc(sh->slot_doubles * world));
    PML_SH_TRY(cudaMemset(sh->xsend.ptr, 0, sh->slot_doubles * sizeof(double)));
    PML_SH_TRY(cudaMemcpy(sh->row_src.ptr, row_src.data(), sizeof(int) * n_clades, cudaMemcpyHostToDevice));
    PML_SH_TRY(cudaMemcpy(sh->my_rows.ptr, mine.data(), sizeof(int) * mine.size(), cudaMemcpyHostToDevice));
    {
        std::vector<int> top_nodes(mine.size());
        for (size_t i = 0; i < mine.size(); ++i) top_nodes[i] = boundary_top[mine[i]];
        PML_SH_TRY(sh->my_top_nodes.alloc(std::max<size_t>(mine.size(), 1)));
        PML_SH_TRY(sh->root_state.alloc(std::max<size_t>(mine.size(), 1) * top->C));
        PML_SH_TRY(cudaMemcpy(sh->my_top_nodes.ptr, top_nodes.data(), sizeof(int) * mine.size(), cudaMemcpyHostToDevice));
    }
    PML_SH_TRY(cudaEventCreateWithFlags(&sh->ev_gather, cudaEventDisableTiming));
    PML_SH_TRY(cudaEventCreateWithFlags(&sh->ev_top, cudaEventDisableTiming));
    // the clade forest's roots take their top-down vectors from the top tree from now on
    const size_t nv = (size_t) clades->n_roots * clades->K * clades->C, ne = (size_t) clades->n_roots * clades->C;
    if (!clades->root_td.ptr) {
        PML_SH_TRY(clades->root_td.alloc(nv));
        PML_SH_TRY(clades->root_td_ex.alloc(ne));
        PML_SH_TRY(cudaMemset(clades->root_td.ptr, 0, nv * sizeof(double)));
        PML_SH_TRY(cudaMemset(clades->root_td_ex.ptr, 0, ne * sizeof(int)));
    }
    clades->have_root_td = true;
#undef PML_SH_TRY
    *out = sh;
    return PML_OK;
}

void pml_shard_destroy(pml_shard *sh) {
    if (!sh) return;
    cudaSetDevice(sh->top->device);
    cudaStreamSynchronize(sh->clades->stream);
    cudaStreamSynchronize(sh->top->stream);
    if (sh->comm && nccl().ok) nccl().CommDestroy(sh->comm);
    if (sh->ev_gather) cudaEventDestroy(sh->ev_gather);
    if (sh->ev_top) cudaEventDestroy(sh->ev_top);
    delete sh;
}

int pml_shard_unique_id(uint8_t *out_id128) {
    if (!out_id128) return fail(PML_ERR_INVALID, "null argument");
    NcclApi &api = nccl();
    if (!api.ok) return fail(PML_ERR_INVALID, api.why);
    NcclApi::UniqueId id;
    const int rc = api.GetUniqueId(&id);
    if (rc) return fail(PML_ERR_CUDA, std::string("ncclGetUniqueId: ") + api.GetErrorString(rc));
    std::memcpy(out_id128, id.internal, 128);
    return PML_OK;
}

int pml_shard_connect(pml_shard *sh, const uint8_t *id128) {
    if (!sh || !id128) return fail(PML_ERR_INVALID, "null argument");
    NcclApi &api = nccl();
    if (!api.ok) return fail(PML_ERR_INVALID, api.why);
    CUDA_TRY(cudaSetDevice(sh->top->device));
    NcclApi::UniqueId id;
    std::memcpy(id.internal, id128, 128);
    const int rc = api.CommInitRank(&sh->comm, sh->world, id, sh->rank);
    if (rc) return fail(PML_ERR_CUDA, std::string("ncclCommInitRank: ") + api.GetErrorString(rc));
    return PML_OK;
}

// phase A: this rank's clades bottom-up, the clade-root runs packed and (when connected) all-gathered; no host sync
int pml_shard_bottom_up(pml_shard *sh, const double *params, int32_t n_sets, int32_t mask_kind) {
    if (!sh || !params) return fail(PML_ERR_INVALID, "null argument");
    pml_problem *p = sh->clades;
    if (n_sets < 1 || n_sets > p->B) return fail(PML_ERR_INVALID, "n_sets exceeds the capacity given at creation");
    if (mask_kind < 0 || mask_kind > PML_MASK_CUSTOM || !p->mask_valid[mask_kind])
        return fail(PML_ERR_INVALID, "the requested mask set has not been produced yet");
    CUDA_TRY(cudaSetDevice(p->device));
    int rc = upload_params(p, params, n_sets);
    if (rc) return rc;
    run_bottom_up(p, n_sets, p->mask[mask_kind].ptr, 1, 1, false);
    LaneDev ln = p->lanes(n_sets);
    ln.jt = bu_tile_lanes(p, ln.J);
    const long long n = (long long) p->n_roots * (p->K + 1) * ln.J;
    shard_pack_kernel<<<grid_for(n), kBlock, 0, p->stream>>>(p->tree(), ln, p->bu.ptr, p->bu_ex.ptr, p->n_roots, sh->xsend.ptr);
    p->launches++;
    sh->last_sets = n_sets;
    sh->last_mask_kind = mask_kind;
    const size_t count = (size_t) sh->n_max * (p->K + 1) * ln.J;   // doubles per rank for this pass
    if (sh->phase_event) cudaEventRecord(sh->phase_event, p->stream);
    if (sh->comm) {
        const int nrc = nccl().AllGather(sh->xsend.ptr, sh->xbuf.ptr, count, 8 /* ncclFloat64 */, sh->comm, p->stream);
        if (nrc) return fail(PML_ERR_CUDA, std::string("ncclAllGather: ") + nccl().GetErrorString(nrc));
    }
    CUDA_TRY(cudaEventRecord(sh->ev_gather, p->stream));
    CUDA_TRY(cudaGetLastError());
    return PML_OK;
}

// single-process stand-in for the all-gather (tests; ranks simulated on one device): every shard's packed runs are
// copied into every shard's gather buffer.  Call between pml_shard_bottom_up and pml_shard_finish.
int pml_shard_exchange_local(pml_shard **shards, int32_t world) {
    if (!shards || world < 1) return fail(PML_ERR_INVALID, "bad argument");
    for (int r = 0; r < world; ++r)
        if (!shards[r] || shards[r]->world != world || shards[r]->comm) return fail(PML_ERR_INVALID, "shards do not form a local group");
    const pml_problem *p0 = shards[0]->clades;
    const size_t count = (size_t) shards[0]->n_max * (p0->K + 1) * shards[0]->last_sets * p0->C;
    for (int r = 0; r < world; ++r) CUDA_TRY(cudaStreamSynchronize(shards[r]->clades->stream));
    for (int dst = 0; dst < world; ++dst)
        for (int src = 0; src < world; ++src)
            CUDA_TRY(cudaMemcpy(shards[dst]->xbuf.ptr + (size_t) src * count, shards[src]->xsend.ptr, count * sizeof(double),
                                cudaMemcpyDeviceToDevice));
    for (int r = 0; r < world; ++r) CUDA_TRY(cudaEventRecord(shards[r]->ev_gather, shards[r]->clades->stream));
    return PML_OK;
}

// phase B: the gathered runs become the top tree's external tips; top tree bottom-up; log-likelihoods [n_sets][C] of the
// whole tree (identical on every rank).  out_loglh null: nothing is fetched and the host is not synchronised.
int pml_shard_finish(pml_shard *sh, const double *params, int32_t n_sets, double *out_loglh) {
    if (!sh || !params) return fail(PML_ERR_INVALID, "null argument");
    pml_problem *p = sh->top;
    if (n_sets != sh->last_sets) return fail(PML_ERR_INVALID, "n_sets differs from the bottom-up phase");
    CUDA_TRY(cudaSetDevice(p->device));
    int rc = upload_params(p, params, n_sets);
    if (rc) return rc;
    CUDA_TRY(cudaStreamWaitEvent(p->stream, sh->ev_gather, 0));
    const int J = n_sets * p->C;
    const long long n = (long long) sh->n_clades * (p->K + 1) * J;
    shard_unpack_kernel<<<grid_for(n), kBlock, 0, p->stream>>>(sh->xbuf.ptr, sh->row_src.ptr, sh->n_clades, p->K, J,
                                                             p->ext_bu.ptr, p->ext_bu_ex.ptr);
    p->launches++;
    p->ext_sets = n_sets;
    // the top tree's own nodes take the same kind of masks (the clade roots, its external tips, were restricted by
    // their own problem already)
    if (!p->mask_valid[sh->last_mask_kind]) return fail(PML_ERR_INVALID, "the top tree has not produced this mask set yet");
    run_bottom_up(p, n_sets, p->mask[sh->last_mask_kind].ptr, 1, 1, false);
    CUDA_TRY(cudaGetLastError());
    if (!out_loglh) return PML_OK;
    return fetch_loglik(p, n_sets, out_loglh);
}

// the fused top-down sweep of the sharded tree after pml_shard_bottom_up + pml_shard_finish with one parameter set:
// top tree on every rank, then this rank's clades from the top-down vectors of their roots.  Outputs stay in the two
// problems' device buffers (fetch them with pml_shard_results); the host is synchronised once, at the end.
int pml_shard_top_down(pml_shard *sh, int32_t want_probs, int32_t want_node_lh, int32_t force_joint) {
    if (!sh) return fail(PML_ERR_INVALID, "null argument");
    if (sh->last_sets != 1) return fail(PML_ERR_INVALID, "the top-down sweep takes one parameter set");
    pml_problem *t = sh->top, *c = sh->clades;
    force_joint = force_joint ? 1 : 0;
    if (force_joint && !(t->joint_done && c->joint_done))
        return fail(PML_ERR_INVALID, "force_joint needs the JOINT states: call pml_shard_joint_finish first");
    CUDA_TRY(cudaSetDevice(t->device));
    int rc = ensure_marginal_buffers(t, want_probs != 0, want_node_lh != 0);
    if (rc) return rc;
    rc = ensure_marginal_buffers(c, want_probs != 0, want_node_lh != 0);
    if (rc) return rc;
    run_top_down(t, force_joint, want_probs != 0, want_node_lh != 0);
    const long long n = (long long) sh->n_mine * (t->K + 1) * t->C;
    shard_root_td_kernel<<<grid_for(n), kBlock, 0, t->stream>>>(t->ext_td.ptr, t->ext_td_ex.ptr, sh->my_rows.ptr, sh->n_mine, t->K,
                                                              t->C, c->root_td.ptr, c->root_td_ex.ptr);
    t->launches++;
    CUDA_TRY(cudaEventRecord(sh->ev_top, t->stream));
    CUDA_TRY(cudaStreamWaitEvent(c->stream, sh->ev_top, 0));
    run_top_down(c, force_joint, want_probs != 0, want_node_lh != 0);
    CUDA_TRY(cudaGetLastError());
    for (pml_problem *p : {t, c}) {
        p->mask_valid[PML_MASK_MAP] = p->mask_valid[PML_MASK_MPPA] = true;
        p->marginal_done = true;
        if (p->n_rows) {
            cudaMemcpyAsync(p->zc_prev_initial.ptr, p->zc_initial.ptr, sizeof(uint64_t) * (size_t) p->n_rows * p->C,
                            cudaMemcpyDeviceToDevice, p->stream);
            cudaMemcpyAsync(p->zc_prev_hasinit.ptr, p->zc_hasinit.ptr, (size_t) p->n_rows * p->C, cudaMemcpyDeviceToDevice, p->stream);
            cudaMemsetAsync(p->zc_touched.ptr, 0, (size_t) p->n_rows * p->C, p->stream);
        }
    }
    CUDA_TRY(cudaStreamSynchronize(t->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return PML_OK;
}

// JOINT on the sharded tree (ml.py:598-622, 687-695), two phases around the exchange like the sum-product pass.
// Phase A: max-product sweep of this rank's clades (pointers kept), clade-root vectors packed and all-gathered.
int pml_shard_joint_bottom_up(pml_shard *sh, const double *params) {
    if (!sh || !params) return fail(PML_ERR_INVALID, "null argument");
    pml_problem *p = sh->clades;
    CUDA_TRY(cudaSetDevice(p->device));
    int rc = ensure_joint_buffers(p);
    if (rc) return rc;
    rc = upload_params(p, params, 1);
    if (rc) return rc;
    run_bottom_up(p, 1, p->mask[PML_MASK_ANNOTATION].ptr, 0, 1, true);
    const LaneDev ln = p->lanes(1);   // (the max-product sweep writes the plain layout)
    const long long n = (long long) p->n_roots * (p->K + 1) * ln.J;
    shard_pack_kernel<<<grid_for(n), kBlock, 0, p->stream>>>(p->tree(), ln, p->bu.ptr, p->bu_ex.ptr, p->n_roots, sh->xsend.ptr);
    p->launches++;
    sh->last_sets = 1;
    sh->last_mask_kind = PML_MASK_ANNOTATION;
    const size_t count = (size_t) sh->n_max * (p->K + 1) * ln.J;
    if (sh->comm) {
        const int nrc = nccl().AllGather(sh->xsend.ptr, sh->xbuf.ptr, count, 8 /* ncclFloat64 */, sh->comm, p->stream);
        if (nrc) return fail(PML_ERR_CUDA, std::string("ncclAllGather: ") + nccl().GetErrorString(nrc));
    }
    CUDA_TRY(cudaEventRecord(sh->ev_gather, p->stream));
    CUDA_TRY(cudaGetLastError());
    return PML_OK;
}

// Phase B: the top tree's max-product sweep on the gathered vectors, its joint log-likelihood (the whole tree's, the same
// on every rank) and traceback; the states it chose for this rank's clade roots start the clades' traceback.
int pml_shard_joint_finish(pml_shard *sh, const double *params, double *out_loglh) {
    if (!sh || !params) return fail(PML_ERR_INVALID, "null argument");
    pml_problem *t = sh->top, *c = sh->clades;
    if (sh->last_sets != 1) return fail(PML_ERR_INVALID, "call pml_shard_joint_bottom_up first");
    CUDA_TRY(cudaSetDevice(t->device));
    int rc = ensure_joint_buffers(t);
    if (rc) return rc;
    rc = upload_params(t, params, 1);
    if (rc) return rc;
    CUDA_TRY(cudaStreamWaitEvent(t->stream, sh->ev_gather, 0));
    const long long n = (long long) sh->n_clades * (t->K + 1) * t->C;
    shard_unpack_kernel<<<grid_for(n), kBlock, 0, t->stream>>>(sh->xbuf.ptr, sh->row_src.ptr, sh->n_clades, t->K, t->C,
                                                             t->ext_bu.ptr, t->ext_bu_ex.ptr);
    t->launches++;
    t->ext_sets = 1;
    run_bottom_up(t, 1, t->mask[PML_MASK_ANNOTATION].ptr, 0, 1, true);
    enqueue_joint_trace(t, nullptr);
    if (sh->n_mine) {
        shard_root_state_kernel<<<grid_for((long long) sh->n_mine * t->C), kBlock, 0, t->stream>>>(
            t->joint_state.ptr, sh->my_top_nodes.ptr, sh->n_mine, t->C, sh->root_state.ptr);
        t->launches++;
    }
    CUDA_TRY(cudaEventRecord(sh->ev_top, t->stream));
    CUDA_TRY(cudaStreamWaitEvent(c->stream, sh->ev_top, 0));
    enqueue_joint_trace(c, sh->root_state.ptr);
    CUDA_TRY(cudaGetLastError());
    for (pml_problem *p : {t, c}) {
        p->mask_valid[PML_MASK_JOINT] = true;
        p->joint_done = true;
    }
    rc = fetch_loglik(t, 1, out_loglh);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return rc;
}

// joint states [N][C] of one of the two problems after pml_shard_joint_finish (which: 0 = top tree, 1 = clade forest)
int pml_shard_joint_states(pml_shard *sh, int32_t which, int32_t *out_state) {
    if (!sh || which < 0 || which > 1 || !out_state) return fail(PML_ERR_INVALID, "bad argument");
    pml_problem *p = which ? sh->clades : sh->top;
    if (!p->joint_done) return fail(PML_ERR_INVALID, "call pml_shard_joint_finish first");
    CUDA_TRY(cudaSetDevice(p->device));
    CUDA_TRY(cudaMemcpyAsync(out_state, p->joint_state.ptr, sizeof(int) * (size_t) p->N * p->C, cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    return PML_OK;
}

// results of the last pml_shard_top_down for one of the two problems (which: 0 = top tree, 1 = clade forest); any
// pointer may be null.  Shapes as in pml_marginal.
int pml_shard_results(pml_shard *sh, int32_t which, double *out_probs, double *out_node_log10lh, uint64_t *out_map_mask,
                      uint64_t *out_mppa_mask, int64_t *out_k_hist) {
    if (!sh || which < 0 || which > 1) return fail(PML_ERR_INVALID, "bad argument");
    pml_problem *p = which ? sh->clades : sh->top;
    if (!p->marginal_done) return fail(PML_ERR_INVALID, "call pml_shard_top_down first");
    CUDA_TRY(cudaSetDevice(p->device));
    const size_t nc = (size_t) p->N * p->C;
    if (out_probs) {
        if (!p->probs.ptr) return fail(PML_ERR_INVALID, "probabilities were not requested");
        CUDA_TRY(cudaMemcpyAsync(out_probs, p->probs.ptr, nc * p->K * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    }
    if (out_node_log10lh) {
        if (!p->node_lh.ptr) return fail(PML_ERR_INVALID, "node likelihoods were not requested");
        CUDA_TRY(cudaMemcpyAsync(out_node_log10lh, p->node_lh.ptr, nc * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    if (out_map_mask) { int r2 = copy_masks_out(p, p->mask[PML_MASK_MAP].ptr, out_map_mask); if (r2) return r2; }
    if (out_mppa_mask) { int r2 = copy_masks_out(p, p->mask[PML_MASK_MPPA].ptr, out_mppa_mask); if (r2) return r2; }
    if (out_k_hist) { int r2 = copy_khist_out(p, out_k_hist); if (r2) return r2; }
    return PML_OK;
}

// device-timed benchmark of the sharded MPPA step (bottom-up, exchange, top tree, top-down): n_iter steps back to back,
// CUDA events on the clade stream around the whole run; the communicator must be connected unless world == 1
int pml_shard_bench_step(pml_shard *sh, const double *params, int32_t n_iter, int32_t flush_l2, double *avg_ms,
                         double *phase_ms) {
    if (!sh || !params || n_iter < 1) return fail(PML_ERR_INVALID, "bad argument");
    // (development: PASTML_B200_SHARD_BENCH_LOCAL=1 times one rank's share alone, on the gather buffer a previous
    // pml_shard_exchange_local left behind)
    const bool local_only = !sh->comm && sh->world > 1 && getenv("PASTML_B200_SHARD_BENCH_LOCAL") != nullptr;
    if (!sh->comm && sh->world > 1 && !local_only) return fail(PML_ERR_INVALID, "connect the shard first");
    pml_problem *t = sh->top, *c = sh->clades;
    CUDA_TRY(cudaSetDevice(t->device));
    int rc = ensure_marginal_buffers(t, true, true);
    if (rc) return rc;
    rc = ensure_marginal_buffers(c, true, true);
    if (rc) return rc;
    const long long flush_n = 64LL << 20;
    if (flush_l2 && !c->flush.ptr) CUDA_TRY(c->flush.alloc(flush_n));
    double total = 0.;
    // phase boundaries of the last step (phase_ms[5]: clades bottom-up + pack, exchange, top tree bottom-up, top tree
    // top-down, clades top-down)
    cudaEvent_t pe[4] = {nullptr, nullptr, nullptr, nullptr};
    if (phase_ms) for (auto &e : pe) CUDA_TRY(cudaEventCreate(&e));
    for (int it = 0; it < n_iter; ++it) {
        if (flush_l2) flush_kernel<<<148 * 8, 256, 0, c->stream>>>(flush_n, c->flush.ptr, (float) it);
        cudaEventRecord(c->ev0, c->stream);
        sh->phase_event = phase_ms ? pe[0] : nullptr;
        rc = pml_shard_bottom_up(sh, params, 1, PML_MASK_ANNOTATION);
        if (rc) return rc;
        if (phase_ms) cudaEventRecord(pe[1], c->stream);
        if (!sh->comm && !local_only) {   // world == 1: the gather is this rank's own runs
            CUDA_TRY(cudaMemcpyAsync(sh->xbuf.ptr, sh->xsend.ptr, (size_t) sh->n_max * (c->K + 1) * c->C * sizeof(double),
                                     cudaMemcpyDeviceToDevice, c->stream));
            CUDA_TRY(cudaEventRecord(sh->ev_gather, c->stream));
        }
        rc = pml_shard_finish(sh, params, 1, nullptr);
        if (rc) return rc;
        if (phase_ms) cudaEventRecord(pe[2], t->stream);
        run_top_down(t, 0, true, true);
        const long long n = (long long) sh->n_mine * (t->K + 1) * t->C;
        shard_root_td_kernel<<<grid_for(n), kBlock, 0, t->stream>>>(t->ext_td.ptr, t->ext_td_ex.ptr, sh->my_rows.ptr, sh->n_mine,
                                                                  t->K, t->C, c->root_td.ptr, c->root_td_ex.ptr);
        CUDA_TRY(cudaEventRecord(sh->ev_top, t->stream));
        if (phase_ms) cudaEventRecord(pe[3], t->stream);
        CUDA_TRY(cudaStreamWaitEvent(c->stream, sh->ev_top, 0));
        run_top_down(c, 0, true, true);
        cudaEventRecord(c->ev1, c->stream);
        CUDA_TRY(cudaEventSynchronize(c->ev1));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->ev0, c->ev1);
        total += ms;
        if (phase_ms && it == n_iter - 1) {
            float x = 0.f;
            cudaEventElapsedTime(&x, c->ev0, pe[0]); phase_ms[0] = x;
            cudaEventElapsedTime(&x, pe[0], pe[1]); phase_ms[1] = x;
            cudaEventElapsedTime(&x, pe[1], pe[2]); phase_ms[2] = x;
            cudaEventElapsedTime(&x, pe[2], pe[3]); phase_ms[3] = x;
            cudaEventElapsedTime(&x, pe[3], c->ev1); phase_ms[4] = x;
        }
    }
    sh->phase_event = nullptr;
    if (phase_ms) for (auto &e : pe) cudaEventDestroy(e);
    CUDA_TRY(cudaStreamSynchronize(t->stream));
    CUDA_TRY(cudaGetLastError());
    for (pml_problem *p : {t, c}) { p->mask_valid[PML_MASK_MAP] = p->mask_valid[PML_MASK_MPPA] = true; p->marginal_done = true; }
    if (avg_ms) *avg_ms = total / n_iter;
    return PML_OK;
}

}  // extern "C"

extern "C" {

int pml_trim_memory(void) {
    g_pool.trim();
    return PML_OK;
}

int pml_last_call_stats(pml_problem *p, double *device_ms, int64_t *kernel_launches) {
    if (!p) return fail(PML_ERR_INVALID, "null argument");
    if (device_ms) *device_ms = p->last_ms;
    if (kernel_launches) *kernel_launches = p->last_launches;
    return PML_OK;
}

int pml_last_kernels(pml_problem *p, char *buf, int32_t buf_len) {
    if (!p || !buf || buf_len < 1) return fail(PML_ERR_INVALID, "null argument");
    std::string out;
    const char *label[2] = {"bottom_up", "top_down"};
    for (int k = 0; k < 2; ++k) {
        if (k) out += ";";
        out += label[k];
        out += ":";
        bool first = true;
        for (const auto &kv : p->last_log[k]) {
            if (!first) out += ",";
            first = false;
            out += kv.first + "=" + std::to_string(kv.second.first) + "/" + std::to_string(kv.second.second);
        }
    }
    std::snprintf(buf, (size_t) buf_len, "%s", out.c_str());
    return PML_OK;
}

int pml_bench_pass(pml_problem *p, const double *params, int32_t n_sets, int32_t which, int32_t n_iter,
                   int32_t flush_l2, double *avg_ms, double *min_ms, int64_t *kernel_launches) {
    if (!p || !params || n_iter < 1) return fail(PML_ERR_INVALID, "bad argument");
    if (n_sets < 1 || n_sets > p->B) return fail(PML_ERR_INVALID, "n_sets exceeds the capacity given at creation");
    if (which != 0 && n_sets != 1) return fail(PML_ERR_INVALID, "this pass takes one parameter set");
    if (which < 0 || which > 4) return fail(PML_ERR_INVALID, "unknown pass");
    CUDA_TRY(cudaSetDevice(p->device));
    const int force_joint = (which == 4) ? 1 : 0;
    if (which == 1 || which == 3 || which == 4) {
        int rc = ensure_marginal_buffers(p, true, true);
        if (rc) return rc;
        if (force_joint && !p->joint_done) return fail(PML_ERR_INVALID, "call pml_joint first");
    }
    if (which == 2) { int r2 = ensure_joint_buffers(p); if (r2) return r2; }
    int rc = upload_params(p, params, n_sets);
    if (rc) return rc;
    if (which == 1 || which == 3 || which == 4) { rc = ensure_mma_td_buffers(p); if (rc) return rc; }
    const long long flush_n = 64LL << 20;   // 256 MiB of floats > 126 MB L2
    if (flush_l2 && !p->flush.ptr) CUDA_TRY(p->flush.alloc(flush_n));
    double total = 0., best = 1e300;
    const uint64_t *base = p->mask[PML_MASK_ANNOTATION].ptr;
    if (which == 3) {   // the top-down sweep alone needs a resident bottom-up state
        run_bottom_up(p, 1, base, 1, 1, false);
    }
    const long long l0 = p->launches;
    for (int it = 0; it < n_iter; ++it) {
        if (flush_l2) flush_kernel<<<148 * 8, 256, 0, p->stream>>>(flush_n, p->flush.ptr, (float) it);
        cudaEventRecord(p->ev0, p->stream);
        if (which == 0) run_bottom_up(p, n_sets, base, 1, 1, false);
        else if (which == 1 || which == 4) {
            run_bottom_up(p, 1, base, 1, 1, false);
            run_top_down(p, force_joint, true, true);
        } else if (which == 3) {
            run_top_down(p, 0, true, true);
        } else {
            run_bottom_up(p, 1, base, 0, 1, true);
            enqueue_joint_trace(p, nullptr);
        }
        cudaEventRecord(p->ev1, p->stream);
        CUDA_TRY(cudaEventSynchronize(p->ev1));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, p->ev0, p->ev1);
        total += ms;
        best = std::min(best, (double) ms);
    }
    CUDA_TRY(cudaGetLastError());
    if (which == 1 || which == 3 || which == 4) p->mask_valid[PML_MASK_MAP] = p->mask_valid[PML_MASK_MPPA] = p->marginal_done = true;
    if (which == 2) { p->mask_valid[PML_MASK_JOINT] = true; p->joint_done = true; }
    if (avg_ms) *avg_ms = total / n_iter;
    if (min_ms) *min_ms = best;
    if (kernel_launches) *kernel_launches = (p->launches - l0) / n_iter;
    return PML_OK;
}

}  // extern "C"
