// gbp_ba.cu -- handle, host-side graph compiler and the C ABI of libgbp_b200.so.
// See include/gbp_b200.h for the contract and the reference methods each entry point replaces.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/gbp_b200.h"
#include "gbp_kernels.cuh"

using namespace gbp;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return fail(GBP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define CHECK_H(h)                                             \
    if (!(h)) return fail(GBP_ERR_INVALID, "null gbp_handle"); \
    CU(cudaSetDevice((h)->device))

// One cudaMalloc per graph: every device array is carved out of a single arena (cudaMalloc / cudaFree
// cost ~0.5 ms each and cudaFree synchronises the device; a graph has ~30 arrays).
struct Arena {
    char* base = nullptr;
    size_t size = 0, used = 0;
    static size_t round_up(size_t b) { return (b + 255) & ~size_t(255); }
    void* take(size_t bytes) {
        void* p = base ? base + used : nullptr;
        used += round_up(bytes ? bytes : 1);
        return p;
    }
};

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    bool owned = false;
    cudaError_t alloc(size_t count) {   // stand-alone allocation (temporaries)
        n = count;
        owned = true;
        return cudaMalloc(reinterpret_cast<void**>(&p), std::max<size_t>(count, 1) * sizeof(T));
    }
    void carve(Arena& a, size_t count) {   // with a.base == nullptr this only measures
        n = count;
        p = reinterpret_cast<T*>(a.take(count * sizeof(T)));
    }
    void release() {
        if (p && owned) cudaFree(p);
        p = nullptr;
    }
    size_t bytes() const { return n * sizeof(T); }
};

}  // namespace

struct gbp_ba_graph {
    int device = 0;
    cudaStream_t stream = nullptr;
    gbp_config cfg{};
    Intrinsics K{};
    int C = 0, L = 0;
    long long F = 0;
    int T = 32;        // edges per tile
    int n_tiles = 0;
    long long n_slots = 0;
    bool robust = false;
    bool priors_set = false;
    bool streaming = false;   // HBM-bound build of the sweep kernel: factored keyframe messages + early issue + L2 prefetch
    int cam_w = CAM_M;   // doubles per stored factor->keyframe message: 27, or 18 with the factored layout (streaming)
    int pf_dist = 0;     // L2 prefetch distance in tiles (streaming build only)
    long long launches = 0;
    int K_chunks = 1;         // landmark chunks of the keyframe-side sums (gbp_config)
    int belief_lanes = 0;     // lanes per landmark in belief_kernel: 0 = by graph size (GBP_TUNE_BELIEF_LANES)

    // host copies (factor order)
    std::vector<int> h_slot_of_factor, h_file_of_factor, h_adj;

    // device state
    DevBuf<Tile> tiles;
    DevBuf<int> lmk_idx, iters, flags, slot_of_factor, lmk_ptr, lmk_slots, cam_tile_ptr, cam_tiles, cam_chunk_ptr, tile_pos;
    DevBuf<double> z, linpoint, msg_cam, msg_lmk, sigma2a;
    DevBuf<double> cam_belief, lmk_belief, cam_prior, lmk_prior, cam_partial, cam_mu0, lmk_mu0, cam_mu, lmk_mu, cam_chol;
    DevBuf<double> tile_partial, tile_metric, metric_out, edge_max, tile_max, cam_max;

    std::map<int, cudaGraphExec_t> graphs;  // key: stages
    Arena arena;
    size_t upload_off = 0, upload_bytes = 0;   // the static tables of the graph: ONE contiguous region = one host->device copy
    size_t zero_off = 0, zero_bytes = 0;       // everything gbp_ba_reset clears: ONE contiguous region = one memset
    void* stage = nullptr;                     // page-locked staging block of that copy (from the shell cache)
    size_t stage_bytes = 0;
    cudaEvent_t snap_event = nullptr;
    // fused [iteration + metrics + copies to pinned host buffers] graphs, keyed by stages; valid for snap_ptrs
    std::map<int, cudaGraphExec_t> snap_graphs;
    void* snap_ptr = nullptr;
    size_t snap_bytes = 0;   // metric_out .. end of lmk_belief (start of the arena)

    // multi-GPU (gbp_ba_attach_comm): this graph is one rank's share of a landmark-partitioned graph
    void* comm = nullptr;            // gbp_comm_s* (not owned)
    int rank = 0, nranks = 1;
    double* gather = nullptr;        // [nranks x K_chunks][C][27] chunk sums of every rank, rank order = chunk order (own cudaMalloc)
    cudaStream_t side = nullptr;     // high-priority stream of the keyframe branch (chunk sums -> all-gather -> keyframe beliefs)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;

    ~gbp_ba_graph();
};

// ----------------------------------------------------------------------------------------
// Shell cache.  What a destroyed graph leaves behind for the next one: its device arena, its page-locked staging block
// and -- when the next graph has the SAME shape (sizes, tiling, parameters, intrinsics => the same arena layout and the
// same kernel arguments) -- its instantiated CUDA graphs.  A client that solves problem after problem
// (create_ba_graph -> iterate -> destroy, i.e. ba.py once per file) then pays neither cudaMalloc / cudaFree (~0.5 ms each,
// cudaFree synchronises the device; single create calls of 5-55 ms were measured in round 1) nor cudaGraphInstantiate.
// Bounded: at most g_cache_max_shells shells, arenas above g_cache_max_arena bytes are never kept.
// ----------------------------------------------------------------------------------------
namespace {

struct ShapeKey {           // compared bytewise: always built by make_key (zero-filled first)
    int device, C, L, T, n_tiles, cam_w, pf_dist, streaming, robust, resident;
    long long F, n_slots;
    double cfg_d[4];
    int cfg_i[10];
    long long cfg_l[2];
    double K[4];
};

struct Shell {
    int device = 0;
    char* arena = nullptr;
    size_t arena_size = 0;
    void* stage = nullptr;
    size_t stage_bytes = 0;
    bool has_key = false;
    ShapeKey key;
    std::map<int, cudaGraphExec_t> graphs, snap_graphs;
    void* snap_ptr = nullptr;
    void drop_graphs() {
        for (auto& kv : graphs) cudaGraphExecDestroy(kv.second);
        for (auto& kv : snap_graphs) cudaGraphExecDestroy(kv.second);
        graphs.clear();
        snap_graphs.clear();
        snap_ptr = nullptr;
        has_key = false;
    }
    void free_all() {
        drop_graphs();
        if (arena) cudaFree(arena);
        if (stage) cudaFreeHost(stage);
        arena = nullptr;
        stage = nullptr;
    }
};

std::mutex g_cache_mu;
std::vector<Shell> g_shells;                 // oldest first
int g_cache_max_shells = 4;
size_t g_cache_max_arena = size_t(1) << 30;
long long g_cache_stats[4] = {0, 0, 0, 0};   // creates, arena reuses, graph reuses, shells evicted

ShapeKey make_key(const gbp_ba_graph* g);

// ONE library stream per device (plus one high-priority stream for the keyframe branch of multi-GPU iterations), created on first
// use and kept for the life of the process: handles created with stream = NULL share it.  Measured on fr1desk (scripts/
// diag_second_graph*.py): as soon as CUDA graphs have been launched on a SECOND stream of the process, every graph replay -- on
// both streams, for the rest of the process -- runs ~5 % slower (8.68 -> 9.1 us per iteration); a second graph on the SAME stream
// costs nothing.  So a stream per handle (the obvious design) is the slow one.
cudaStream_t library_stream(int device, bool high_priority) {
    static std::mutex mu;
    static std::map<int, cudaStream_t> streams[2];
    std::lock_guard<std::mutex> lk(mu);
    auto& m = streams[high_priority ? 1 : 0];
    auto it = m.find(device);
    if (it != m.end()) return it->second;
    cudaStream_t s = nullptr;
    int lo = 0, hi = 0;
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    if (high_priority && cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess) return nullptr;
    if (cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, high_priority ? hi : 0) != cudaSuccess) return nullptr;
    m[device] = s;
    return s;
}

// Best shell for a new graph: the one of the same shape (its graphs stay valid), else the smallest arena that fits.
bool cache_take(const ShapeKey& key, int device, size_t arena_need, Shell* out) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    g_cache_stats[0]++;
    int best = -1;
    for (int i = 0; i < (int)g_shells.size(); ++i) {
        const Shell& sh = g_shells[i];
        if (sh.device != device || sh.arena_size < arena_need) continue;
        if (sh.has_key && memcmp(&sh.key, &key, sizeof(key)) == 0) { best = i; break; }
        if (best < 0 || sh.arena_size < g_shells[best].arena_size) best = i;
    }
    if (best < 0) return false;
    *out = std::move(g_shells[best]);
    g_shells.erase(g_shells.begin() + best);
    if (!(out->has_key && memcmp(&out->key, &key, sizeof(key)) == 0)) out->drop_graphs();
    g_cache_stats[1]++;
    if (out->has_key) g_cache_stats[2]++;
    return true;
}

void cache_put(Shell&& sh) {
    if (!sh.arena || sh.arena_size > g_cache_max_arena || g_cache_max_shells <= 0) { sh.free_all(); return; }
    std::lock_guard<std::mutex> lk(g_cache_mu);
    g_shells.push_back(std::move(sh));
    while ((int)g_shells.size() > g_cache_max_shells) {
        g_shells.front().free_all();
        g_shells.erase(g_shells.begin());
        g_cache_stats[3]++;
    }
}

}  // namespace

static void comm_teardown(gbp_ba_graph* g);   // gbp_dist.cu.inc

gbp_ba_graph::~gbp_ba_graph() {
    if (snap_event) cudaEventDestroy(snap_event);
    const bool had_comm = comm != nullptr;
    if (had_comm) comm_teardown(this);   // drops the CUDA graphs (they hold collective nodes of this communicator) first
    // everything else goes back to the shell cache (the stream was synchronised by gbp_ba_destroy; a graph that dies
    // on an error path of gbp_ba_create has nothing in flight that reads the arena after its failed call returned)
    Shell sh;
    sh.device = device;
    sh.arena = arena.base; sh.arena_size = arena.size;
    sh.stage = stage; sh.stage_bytes = stage_bytes;
    sh.graphs = std::move(graphs); sh.snap_graphs = std::move(snap_graphs); sh.snap_ptr = snap_ptr;
    sh.key = make_key(this);
    sh.has_key = arena.base != nullptr && !had_comm;
    if (arena.base || stage) {
        if (stream) cudaStreamSynchronize(stream);
        cache_put(std::move(sh));
    } else {
        sh.free_all();
    }
}

namespace {

ShapeKey make_key(const gbp_ba_graph* g) {
    ShapeKey k;
    memset(&k, 0, sizeof(k));
    k.device = g->device; k.C = g->C; k.L = g->L; k.T = g->T; k.n_tiles = g->n_tiles; k.cam_w = g->cam_w; k.pf_dist = g->pf_dist;
    k.streaming = g->streaming ? 1 : 0; k.robust = g->robust ? 1 : 0; k.resident = 0;
    k.F = g->F; k.n_slots = g->n_slots;
    k.cfg_d[0] = g->cfg.gauss_noise_std; k.cfg_d[1] = g->cfg.eta_damping; k.cfg_d[2] = g->cfg.beta; k.cfg_d[3] = g->cfg.Nstds;
    k.cfg_i[0] = g->cfg.num_undamped_iters; k.cfg_i[1] = g->cfg.min_linear_iters; k.cfg_i[2] = g->cfg.loss;
    k.cfg_i[3] = g->cfg.tile_edges; k.cfg_i[4] = g->cfg.lmk_block; k.cfg_i[5] = g->cfg.kernel_variant;
    k.cfg_i[6] = g->cfg.lmk_chunks; k.cfg_i[7] = g->cfg.lmk_chunk_first; k.cfg_i[8] = g->cfg.lmk_chunks_total; k.cfg_i[9] = g->K_chunks;
    k.cfg_l[0] = g->cfg.lmk_first; k.cfg_l[1] = g->cfg.lmk_total;
    k.K[0] = g->K.fx; k.K[1] = g->K.fy; k.K[2] = g->K.cx; k.K[3] = g->K.cy;
    return k;
}

SweepParams sweep_params(gbp_ba_graph* g, int stages) {
    SweepParams p{};
    p.tiles = g->tiles.p; p.lmk_idx = g->lmk_idx.p; p.z = g->z.p; p.linpoint = g->linpoint.p;
    p.msg_cam = g->msg_cam.p; p.msg_lmk = g->msg_lmk.p; p.iters = g->iters.p; p.flags = g->flags.p;
    p.sigma2a = g->sigma2a.p; p.cam_belief = g->cam_belief.p; p.cam_chol = g->cam_chol.p; p.lmk_belief = g->lmk_belief.p;
    p.tile_partial = g->tile_partial.p; p.tile_pos = g->tile_pos.p; p.K = g->K;
    p.var0 = g->cfg.gauss_noise_std * g->cfg.gauss_noise_std;
    p.eta_damping = g->cfg.eta_damping; p.beta = g->cfg.beta; p.nstds = g->cfg.Nstds;
    p.num_undamped = g->cfg.num_undamped_iters; p.min_linear = g->cfg.min_linear_iters;
    p.loss = g->cfg.loss; p.stages = stages; p.n_tiles = g->n_tiles; p.pf_dist = g->pf_dist;
    return p;
}

template <int T, bool ROBUST, bool STREAM>
void launch_sweep_k(gbp_ba_graph* g, const SweepParams& p, size_t smem) {
    // the complete synchronous iteration gets the build with compile-time stages (a non-robust graph ignores the robustify stage)
    const int st = p.stages & ST_FULL;
    if ((st | (ROBUST ? 0 : ST_ROBUSTIFY)) == ST_FULL) sweep_kernel<T, ROBUST, STREAM, ST_FULL><<<g->n_tiles, T, smem, g->stream>>>(p);
    else sweep_kernel<T, ROBUST, STREAM, 0><<<g->n_tiles, T, smem, g->stream>>>(p);
}

template <int T>
int launch_sweep_t(gbp_ba_graph* g, int stages) {
    const SweepParams p = sweep_params(g, stages);
    if (g->streaming) {
        // HBM-bound graphs (more than 8192 tiles, or kernel_variant 2): factored keyframe messages, early issue, far-ahead
        // L2 prefetch (10 M-factor graph, same box: 1.257 ms per launch for the descriptor-first full-row build, 0.995 ms for this one)
        constexpr int TP = T <= 64 ? T : 64;      // T <= 64 checked at creation
        constexpr size_t smem = sweep_smem_bytes<TP, true>();
        static_assert(smem <= 48 * 1024, "factored sweep tile must fit the default dynamic shared memory limit");
        if (g->robust) launch_sweep_k<TP, true, true>(g, p, smem);
        else launch_sweep_k<TP, false, true>(g, p, smem);
    } else {
        constexpr size_t smem = sweep_smem_bytes<T, false>();
        static_assert(smem <= 48 * 1024, "sweep tile must fit the default dynamic shared memory limit");
        if (g->robust) launch_sweep_k<T, true, false>(g, p, smem);
        else launch_sweep_k<T, false, false>(g, p, smem);
    }
    g->launches++;
    CU(cudaGetLastError());
    return GBP_OK;
}

int launch_sweep(gbp_ba_graph* g, int stages) {
    if (g->n_tiles == 0) return GBP_OK;
    switch (g->T) {
        case 32: return launch_sweep_t<32>(g, stages);
        case 64: return launch_sweep_t<64>(g, stages);
        default: return launch_sweep_t<128>(g, stages);
    }
}

// landmark beliefs + keyframe partial sums (+ keyframe beliefs when finalise); parts: bit0 keyframes, bit1 landmarks
int launch_belief(gbp_ba_graph* g, int finalise, int parts = 3, cudaStream_t stream = nullptr) {
    if (!stream) stream = g->stream;
    BeliefParams p{};
    p.msg_lmk = g->msg_lmk.p; p.lmk_prior = g->lmk_prior.p; p.lmk_belief = g->lmk_belief.p;
    p.lmk_ptr = g->lmk_ptr.p; p.lmk_slots = g->lmk_slots.p; p.tile_partial = g->tile_partial.p;
    p.cam_tile_ptr = g->cam_tile_ptr.p; p.cam_tiles = g->cam_tiles.p; p.cam_prior = g->cam_prior.p;
    p.cam_belief = g->cam_belief.p; p.cam_chol = g->cam_chol.p; p.cam_partial = g->cam_partial.p; p.cam_mu = g->cam_mu.p; p.lmk_mu = g->lmk_mu.p;
    p.cam_chunk_ptr = g->cam_chunk_ptr.p; p.K = g->K_chunks;
    p.L = g->L; p.C = g->C; p.finalise = finalise; p.parts = parts;
    // small graphs are latency-bound: a whole warp per landmark gathers a degree-46 landmark in 2 dependent rounds; from ~50 k
    // landmarks on one thread per landmark wins (125 k landmarks / 1.25 M factors, the per-rank share at 8 GPUs: 45 vs 57 us)
    const int lanes = g->belief_lanes ? g->belief_lanes : (g->L >= 49152 ? 1 : (g->L > 8192 ? 8 : 32));
    const int per_cta = 128 / lanes;
    const int blocks = ((parts & 2) ? (g->L + per_cta - 1) / per_cta : 0) + ((parts & 1) ? (g->C + 3) / 4 : 0);
    if (blocks == 0) return GBP_OK;
    switch (lanes) {
        case 1: belief_kernel<1><<<blocks, 128, 0, stream>>>(p); break;
        case 8: belief_kernel<8><<<blocks, 128, 0, stream>>>(p); break;
        default: belief_kernel<32><<<blocks, 128, 0, stream>>>(p); break;
    }
    g->launches++;
    CU(cudaGetLastError());
    return GBP_OK;
}

template <int T>
int launch_metric_t(gbp_ba_graph* g) {
    MetricParams p{};
    p.tiles = g->tiles.p; p.lmk_idx = g->lmk_idx.p; p.z = g->z.p; p.iters = g->iters.p; p.sigma2a = g->sigma2a.p;
    p.cam_belief = g->cam_belief.p; p.lmk_belief = g->lmk_belief.p; p.tile_metric = g->tile_metric.p; p.K = g->K;
    p.var0 = g->cfg.gauss_noise_std * g->cfg.gauss_noise_std; p.robust = g->robust ? 1 : 0;
    metric_kernel<T><<<g->n_tiles, T, 0, g->stream>>>(p);
    CU(cudaGetLastError());
    return GBP_OK;
}

template <int T>
int launch_lammax_t(gbp_ba_graph* g) {
    edge_lammax_kernel<T><<<g->n_tiles, T, 0, g->stream>>>(g->tiles.p, g->linpoint.p, g->sigma2a.p, g->robust ? 1 : 0,
                                                           g->cfg.gauss_noise_std * g->cfg.gauss_noise_std, g->K,
                                                           g->edge_max.p, g->tile_max.p);
    CU(cudaGetLastError());
    return GBP_OK;
}

template <int T>
int launch_init_t(gbp_ba_graph* g) {
    init_edges_kernel<T><<<g->n_tiles, T, 0, g->stream>>>(g->tiles.p, g->lmk_idx.p, g->cam_belief.p, g->lmk_belief.p,
                                                          g->cfg.gauss_noise_std * g->cfg.gauss_noise_std, g->linpoint.p,
                                                          g->iters.p, g->flags.p, g->sigma2a.p);
    CU(cudaGetLastError());
    return GBP_OK;
}

#define DISPATCH_T(g, fn)                         \
    ((g)->T == 32 ? fn<32>(g) : (g)->T == 64 ? fn<64>(g) : fn<128>(g))

struct FieldInfo {
    int indexed;   // 0 = keyframes, 1 = landmarks, 2 = factors
    int words;     // row width in 4-byte words
    bool writable;
};

bool field_info(int field, FieldInfo* fi) {
    switch (field) {
        case GBP_F_CAM_BELIEF: *fi = {0, CAM_B * 2, true}; return true;
        case GBP_F_LMK_BELIEF: *fi = {1, LMK_B * 2, true}; return true;
        case GBP_F_CAM_PRIOR: *fi = {0, CAM_M * 2, true}; return true;
        case GBP_F_LMK_PRIOR: *fi = {1, LMK_M * 2, true}; return true;
        case GBP_F_MSG_CAM: *fi = {2, CAM_M * 2, true}; return true;
        case GBP_F_MSG_LMK: *fi = {2, LMK_M * 2, true}; return true;
        case GBP_F_LINPOINT: *fi = {2, 18, true}; return true;
        case GBP_F_ITERS_SINCE_RELIN: *fi = {2, 1, true}; return true;
        case GBP_F_FLAGS: *fi = {2, 1, true}; return true;
        case GBP_F_ADAPTIVE_VAR: *fi = {2, 2, true}; return true;
        case GBP_F_MEASUREMENT: *fi = {2, 4, false}; return true;
        case GBP_F_JACOBIAN_B: *fi = {2, 40, false}; return true;
        case GBP_F_ADJ: *fi = {2, 2, false}; return true;
        case GBP_F_FILE_INDEX: *fi = {2, 1, false}; return true;
        case GBP_F_CAM_PARTIAL: *fi = {0, CAM_M * 2, false}; return true;
        case GBP_F_CAM_MU: *fi = {0, 12, false}; return true;
        case GBP_F_LMK_MU: *fi = {1, 6, false}; return true;
        default: return false;
    }
}

void* field_dev_ptr(gbp_ba_graph* g, int field) {
    switch (field) {
        case GBP_F_CAM_BELIEF: return g->cam_belief.p;
        case GBP_F_LMK_BELIEF: return g->lmk_belief.p;
        case GBP_F_CAM_PRIOR: return g->cam_prior.p;
        case GBP_F_LMK_PRIOR: return g->lmk_prior.p;
        case GBP_F_MSG_CAM: return g->msg_cam.p;
        case GBP_F_MSG_LMK: return g->msg_lmk.p;
        case GBP_F_LINPOINT: return g->linpoint.p;
        case GBP_F_ITERS_SINCE_RELIN: return g->iters.p;
        case GBP_F_FLAGS: return g->flags.p;
        case GBP_F_ADAPTIVE_VAR: return g->sigma2a.p;
        case GBP_F_MEASUREMENT: return g->z.p;
        case GBP_F_CAM_PARTIAL: return g->cam_partial.p;
        case GBP_F_CAM_MU: return g->cam_mu.p;
        case GBP_F_LMK_MU: return g->lmk_mu.p;
        default: return nullptr;
    }
}

int comm_all_reduce(gbp_ba_graph* g, double* dev, size_t n, bool max_op);   // gbp_dist.cu.inc: in place over the ranks, on the handle's stream
int enqueue_beliefs(gbp_ba_graph* g);   // gbp_dist.cu.inc: belief update after a sweep, with the keyframe exchange when a communicator is attached

// one synchronous iteration on the handle's stream(s): sweep, then the belief update
int enqueue_iteration(gbp_ba_graph* g, int stages) {
    int rc = launch_sweep(g, stages);
    if (rc == GBP_OK && (stages & ST_BELIEFS)) rc = enqueue_beliefs(g);
    return rc;
}

// CUDA graph of `reps` consecutive iterations [sweep, beliefs] x reps (fewer, longer launches: the gap between two
// graph launches is larger than the gap between two nodes of one graph, which matters at 10 us per iteration)
int get_graph(gbp_ba_graph* g, int stages, cudaGraphExec_t* out, int reps = 1) {
    const int key = stages | (reps << 8);
    auto it = g->graphs.find(key);
    if (it != g->graphs.end()) { *out = it->second; return GBP_OK; }
    cudaGraph_t graph = nullptr;
    CU(cudaStreamBeginCapture(g->stream, cudaStreamCaptureModeThreadLocal));
    const long long before = g->launches;
    int rc = GBP_OK;
    for (int r = 0; r < reps && rc == GBP_OK; ++r) rc = enqueue_iteration(g, stages);
    g->launches = before;  // capture does not execute
    cudaError_t e = cudaStreamEndCapture(g->stream, &graph);
    if (rc != GBP_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return fail(GBP_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(GBP_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    g->graphs[key] = exec;
    *out = exec;
    return GBP_OK;
}

// kernels of this library per synchronous iteration: sweep + beliefs, or with a communicator sweep, landmark beliefs, keyframe
// chunk sums, keyframe beliefs (the collective's own kernel is not counted)
int launches_per_iteration(const gbp_ba_graph* g) { return (g->n_tiles > 0 ? 1 : 0) + (g->comm ? 3 : 1); }

int iteration_stages(int robustify, int local_relin) {
    // synchronous_iteration (gbp/gbp.py:86-92) for a graph with nonlinear factors
    int st = ST_MESSAGES | ST_BELIEFS;
    if (robustify) st |= ST_ROBUSTIFY;
    if (local_relin) st |= ST_RELIN | ST_LOCAL_DAMPING;
    return st;
}


// ----------------------------------------------------------------------------------------
// Host graph compiler: measurement list -> engine storage order.  Pure host code (no CUDA call), so it is also
// exported on its own (gbp_plan_*) and tested on CPU.  Replaces the construction loops of create_ba_graph
// (gbp/gbp_ba.py:128-143), which are O(C F) in the reference.
//   factor order  = stable sort of the measurements by camera (the order the reference creates its factors in)
//   storage order = runs of equal (landmark block, camera), factor order inside a run; every run is cut into tiles of
//                   <= T edges, a tile owns T consecutive slots, padding only at the end of a run's last tile
// ----------------------------------------------------------------------------------------
struct GraphPlan {
    int T = 32;
    long long lblock = 1;
    std::vector<Tile> tiles;
    std::vector<int> slot_of_factor, file_of_factor, adj;      // factor order; adj = (camera, landmark) pairs
    std::vector<int> lmk_idx;                                  // [slots] landmark of the edge in that slot (0 in padding)
    std::vector<double> z;                                     // [slots][2] (only when measurements were given)
    std::vector<int> lmk_ptr, lmk_slots, cam_tile_ptr, cam_tiles;
    std::vector<int> tile_pos;                                 // [tiles] position of a tile in cam_tiles (the order its partial sums are stored in)
    int n_chunks = 1;
    std::vector<int> tile_chunk;                               // [tiles] landmark chunk of every tile
    std::vector<int> cam_chunk_ptr;                            // [C][chunks + 1] positions in cam_tiles where a keyframe's chunks start
    long long n_slots() const { return (long long)tiles.size() * T; }
};

// tile_edges / lmk_block: 0 = automatic (same rules for every caller)
int choose_tiling(int tile_edges, long long lmk_block, int L, int64_t F, int* T_out, long long* lblock_out) {
    int T = tile_edges;
    if (T == 0) T = F >= 64LL * 148 * 6 ? 64 : 32;   // measured: 64-edge tiles beat 128 on large graphs; 32 spreads small ones
    if (T != 32 && T != 64 && T != 128) return fail(GBP_ERR_INVALID, "tile_edges must be 0, 32, 64 or 128");
    long long lblock = lmk_block;
    if (lblock <= 0) lblock = ((long long)L * LMK_B * 8 <= (24LL << 20)) ? std::max(L, 1) : 262144;
    *T_out = T;
    *lblock_out = lblock;
    return GBP_OK;
}

// chunk_bounds: [chunks + 1] landmark boundaries of the chunks (0 ... L); a landmark block never straddles a chunk
int plan_graph(int T, long long lblock, const std::vector<long long>& chunk_bounds, int C, int L, int64_t F, const int32_t* cam_id,
               const int32_t* lmk_id, const double* z, GraphPlan* plan) {
    // Counting sorts only, one pass per table (13 k edges: ~0.2 ms; 10 M edges: ~0.5 s).  Positions fit int32 (checked below).
    if (F >= (1LL << 31) / 2) return fail(GBP_ERR_INVALID, "graph too large for int32 slots");
    plan->T = T;
    plan->lblock = lblock;
    const int K = std::max<int>(1, (int)chunk_bounds.size() - 1);
    plan->n_chunks = K;
    const int Cn = std::max(C, 1);
    // factor order = stable sort of the measurement list by camera (gbp/gbp_ba.py:128-130); validation in the counting pass
    std::vector<int> cam_start((size_t)C + 1, 0);
    for (int64_t i = 0; i < F; ++i) {
        const int c = cam_id[i], l = lmk_id[i];
        if ((unsigned)c >= (unsigned)C) return fail(GBP_ERR_INVALID, "measurement %lld: camera id %d out of range", (long long)i, c);
        if ((unsigned)l >= (unsigned)L) return fail(GBP_ERR_INVALID, "measurement %lld: landmark id %d out of range", (long long)i, l);
        cam_start[(size_t)c + 1]++;
    }
    for (int c = 0; c < C; ++c) cam_start[c + 1] += cam_start[c];
    // landmark blocks: every chunk is cut into blocks of lblock landmarks; blocks are numbered chunk by chunk
    std::vector<int> blk_of_lmk((size_t)L), chunk_of_blk;
    for (int k = 0; k < K; ++k) {
        const long long l0 = chunk_bounds.size() > 1 ? chunk_bounds[k] : 0, l1 = chunk_bounds.size() > 1 ? chunk_bounds[k + 1] : L;
        for (long long b0 = l0; b0 < l1; b0 += lblock) {
            const int b = (int)chunk_of_blk.size();
            std::fill(blk_of_lmk.begin() + b0, blk_of_lmk.begin() + std::min(l1, b0 + lblock), b);
            chunk_of_blk.push_back(k);
        }
    }
    if (chunk_of_blk.empty()) chunk_of_blk.push_back(0);
    const long long nb = (long long)chunk_of_blk.size();
    const size_t nkeys = (size_t)nb * (size_t)Cn;
    if (nkeys >= (size_t)1 << 31) return fail(GBP_ERR_INVALID, "too many (landmark block, keyframe) runs");

    // one scatter pass: factor f of measurement i; its (keyframe, landmark), its run key = block * C + keyframe; run sizes
    plan->file_of_factor.resize((size_t)F);
    plan->adj.resize((size_t)F * 2);
    std::vector<int> key((size_t)F);
    std::vector<int> run_count(nkeys + 1, 0);
    {
        int* fof = plan->file_of_factor.data();
        int* adj = plan->adj.data();
        std::vector<int> pos(cam_start.begin(), cam_start.end() - 1);
        for (int64_t i = 0; i < F; ++i) {
            const int c = cam_id[i], l = lmk_id[i];
            const int f = pos[c]++;
            fof[f] = (int)i;
            adj[2 * (size_t)f] = c;
            adj[2 * (size_t)f + 1] = l;
            const int k = blk_of_lmk[(size_t)l] * Cn + c;
            key[(size_t)f] = k;
            run_count[(size_t)k]++;
        }
    }
    // storage order: (landmark block, camera) runs, factor order inside a run; tiles per run
    std::vector<Tile>& tiles = plan->tiles;
    tiles.clear();
    plan->tile_chunk.clear();
    std::vector<int> run_slot(nkeys, 0);
    {
        long long tcount = 0;
        for (size_t k = 0; k < nkeys; ++k) {
            if (tcount * T >= (1LL << 31) - 2 * T) return fail(GBP_ERR_INVALID, "graph too large for int32 slots");
            run_slot[k] = (int)(tcount * T);
            int left = run_count[k];
            const int cam = (int)(k % (size_t)Cn), chunk = chunk_of_blk[k / (size_t)Cn];
            while (left > 0) {
                Tile t;
                t.cam = cam;
                t.count = std::min(left, T);
                tiles.push_back(t);
                plan->tile_chunk.push_back(chunk);
                left -= t.count;
                ++tcount;
            }
        }
    }
    const long long n_slots = plan->n_slots();
    if (n_slots >= (1LL << 31)) return fail(GBP_ERR_INVALID, "graph too large for int32 slots");
    // one pass in factor order: the slot of every factor (runs are contiguous in slots except for the padding of their last
    // tile, which lies at the END of the run, so consecutive positions are correct), the slot-ordered inputs, landmark degrees
    plan->slot_of_factor.resize((size_t)F);
    plan->lmk_idx.assign((size_t)n_slots, 0);
    plan->z.assign(z ? (size_t)n_slots * 2 : 0, 0.0);
    plan->lmk_ptr.assign((size_t)L + 1, 0);
    {
        int* sof = plan->slot_of_factor.data();
        int* lidx = plan->lmk_idx.data();
        double* zs = plan->z.data();
        int* lptr = plan->lmk_ptr.data();
        const int* adj = plan->adj.data();
        const int* fof = plan->file_of_factor.data();
        for (int64_t f = 0; f < F; ++f) {
            const int s = run_slot[(size_t)key[(size_t)f]]++;
            const int l = adj[2 * (size_t)f + 1];
            sof[f] = s;
            lidx[s] = l;
            lptr[(size_t)l + 1]++;
            if (z) {
                const size_t i = (size_t)fof[f];
                zs[2 * (size_t)s] = z[2 * i];
                zs[2 * (size_t)s + 1] = z[2 * i + 1];
            }
        }
    }
    // CSR by landmark over slots (factor order inside a landmark = adj_factors order)
    plan->lmk_slots.resize((size_t)F);
    for (int l = 0; l < L; ++l) plan->lmk_ptr[l + 1] += plan->lmk_ptr[l];
    {
        std::vector<int> pos(plan->lmk_ptr.begin(), plan->lmk_ptr.end() - 1);
        int* ls = plan->lmk_slots.data();
        const int* adj = plan->adj.data();
        const int* sof = plan->slot_of_factor.data();
        for (int64_t f = 0; f < F; ++f) ls[pos[(size_t)adj[2 * (size_t)f + 1]]++] = sof[f];
    }
    // CSR by camera over tiles
    plan->cam_tile_ptr.assign((size_t)C + 1, 0);
    plan->cam_tiles.assign(tiles.size(), 0);
    for (const Tile& t : tiles) plan->cam_tile_ptr[(size_t)t.cam + 1]++;
    for (int c = 0; c < C; ++c) plan->cam_tile_ptr[c + 1] += plan->cam_tile_ptr[c];
    {
        std::vector<int> pos(plan->cam_tile_ptr.begin(), plan->cam_tile_ptr.end() - 1);
        for (size_t t = 0; t < tiles.size(); ++t) plan->cam_tiles[(size_t)pos[tiles[t].cam]++] = (int)t;
    }
    plan->tile_pos.assign(tiles.size(), 0);
    for (size_t q = 0; q < tiles.size(); ++q) plan->tile_pos[(size_t)plan->cam_tiles[q]] = (int)q;
    // a keyframe's tiles are listed in tile order = chunk-major: where its chunks start
    plan->cam_chunk_ptr.assign((size_t)C * (K + 1), 0);
    for (int c = 0; c < C; ++c) {
        int* ptr = &plan->cam_chunk_ptr[(size_t)c * (K + 1)];
        int q = plan->cam_tile_ptr[c];
        for (int k = 0; k < K; ++k) {
            ptr[k] = q;
            while (q < plan->cam_tile_ptr[c + 1] && plan->tile_chunk[(size_t)plan->cam_tiles[(size_t)q]] == k) ++q;
        }
        ptr[K] = q;
        if (q != plan->cam_tile_ptr[c + 1]) return fail(GBP_ERR_INVALID, "internal: tiles of keyframe %d are not in chunk order", c);
    }
    return GBP_OK;
}

// Automatic number of landmark chunks: the largest power of two <= 8 that leaves a chunk at least 125000 landmarks.  Every chunk
// boundary ends the (chunk, keyframe) runs of edges and with them a partly filled tile per keyframe, so small chunks cost padding
// (125 k landmarks / 1000 keyframes / 10 observations cut into 8 chunks: 22 % padding slots, sweep 6 % slower); 8 chunks of the
// 1 M-landmark graph cost 1.2 %.
int auto_chunks(long long L) {
    int K = 1;
    while (K < 8 && L / (2 * K) >= 125000) K *= 2;
    return K;
}

// The chunking of a graph: local landmark boundaries of its chunks from the configuration (see gbp_config), 0 = automatic.
int chunk_bounds_of(const gbp_config* cfg, int L, std::vector<long long>* cb) {
    int K = cfg ? cfg->lmk_chunks : 0;
    long long first = 0, total = 0, lmk_first = 0, lmk_total = L;
    if (K <= 0) {
        K = auto_chunks(L);
        total = K;
    } else {
        first = cfg->lmk_chunk_first; total = cfg->lmk_chunks_total; lmk_first = cfg->lmk_first; lmk_total = cfg->lmk_total;
        if (K > 4096 || first < 0 || total < K || first + K > total || lmk_first < 0 || lmk_total < lmk_first + L)
            return fail(GBP_ERR_INVALID, "inconsistent landmark chunking (%d chunks from %lld of %lld)", K, first, total);
    }
    cb->assign((size_t)K + 1, 0);
    for (int k = 0; k <= K; ++k) (*cb)[k] = lmk_total * (first + k) / total - lmk_first;
    if ((*cb)[0] != 0 || (*cb)[K] != L) return fail(GBP_ERR_INVALID, "landmark chunks do not cover this graph's %d landmarks", L);
    return GBP_OK;
}

}  // namespace

extern "C" {

const char* gbp_last_error(void) { return g_err.c_str(); }
int gbp_abi_version(void) { return GBP_B200_ABI_VERSION; }

int gbp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

static int ba_create_impl(const gbp_config* cfg, int32_t C, int32_t L, int64_t F, const int32_t* cam_id, const int32_t* lmk_id,
                  const double* z, const double* cam_mu0, const double* lmk_mu0, const double K[4], int device,
                  void* stream, gbp_handle* out) {
    if (!cfg || !out || !K) return fail(GBP_ERR_INVALID, "null argument");
    if (C < 0 || L < 0 || F < 0) return fail(GBP_ERR_INVALID, "negative size");
    if (F > 0 && (!cam_id || !lmk_id || !z)) return fail(GBP_ERR_INVALID, "null measurement arrays");
    if ((C > 0 && !cam_mu0) || (L > 0 && !lmk_mu0)) return fail(GBP_ERR_INVALID, "null initial means");
    if (cfg->loss < 0 || cfg->loss > 2) return fail(GBP_ERR_INVALID, "unknown loss %d", cfg->loss);
    if (!(cfg->gauss_noise_std > 0)) return fail(GBP_ERR_INVALID, "gauss_noise_std must be > 0");
    *out = nullptr;
    if (gbp_device_count() <= 0)
        return fail(GBP_ERR_NO_DEVICE, "no CUDA device visible: gbp_b200 has no CPU fallback for the BA sweep");
    if (device < 0 || device >= gbp_device_count()) return fail(GBP_ERR_INVALID, "device %d out of range", device);
    CU(cudaSetDevice(device));
    std::unique_ptr<gbp_ba_graph> owner(new gbp_ba_graph());   // freed on every early return / exception below
    gbp_ba_graph* g = owner.get();
    g->device = device;
    g->cfg = *cfg;
    g->K = Intrinsics{K[0], K[1], K[2], K[3]};
    g->C = C; g->L = L; g->F = F;
    g->robust = cfg->loss != GBP_LOSS_NONE;
    if (stream) {
        g->stream = reinterpret_cast<cudaStream_t>(stream);
    } else {
        g->stream = library_stream(device, /*high_priority=*/false);
        if (!g->stream) return fail(GBP_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(cudaGetLastError()));
    }

    // ---------------- host graph compiler ----------------
    int T = 32;
    long long lblock = 1;
    {
        int rc = choose_tiling(cfg->tile_edges, cfg->lmk_block, L, F, &T, &lblock);
        if (rc != GBP_OK) return rc;
    }
    if (cfg->kernel_variant < 0 || cfg->kernel_variant > 2) return fail(GBP_ERR_INVALID, "kernel_variant must be 0 (automatic), 1 or 2");
    if (cfg->kernel_variant == 2 && T == 128) return fail(GBP_ERR_INVALID, "kernel_variant 2 (streaming build) needs tile_edges 32 or 64");
    g->T = T;
    GraphPlan plan;
    {
        std::vector<long long> cb;
        int rc = chunk_bounds_of(cfg, L, &cb);
        if (rc == GBP_OK) rc = plan_graph(T, lblock, cb, C, L, F, cam_id, lmk_id, z, &plan);
        if (rc != GBP_OK) return rc;
    }
    g->K_chunks = plan.n_chunks;
    const std::vector<Tile>& tiles = plan.tiles;
    g->n_tiles = (int)tiles.size();
    g->n_slots = plan.n_slots();
    g->h_file_of_factor = std::move(plan.file_of_factor);
    g->h_adj = std::move(plan.adj);
    g->h_slot_of_factor = std::move(plan.slot_of_factor);
    // Which build of the sweep kernel: graphs of more than 8192 tiles stream ~7 GB per sweep and get the streaming build
    // (factored keyframe messages: 144 B less per edge; no descriptor wait in the prologue; every CTA prefetches the streams
    // of the tile ~38 k edges ahead into L2 -- A/B on the 10 M-factor graph: flat optimum between 400 and 750 tiles of 64
    // edges, 1800 and more thrash L2; the distance shrinks with the graph so that it stays below one wave of tiles).
    g->streaming = cfg->kernel_variant == 2 || (cfg->kernel_variant == 0 && g->n_tiles > 8192 && T <= 64);
    if (g->streaming) {
        g->cam_w = CAM_MF;
        g->pf_dist = g->n_tiles > 8192 ? (int)std::min<long long>(38400 / T, std::max(1, g->n_tiles / 32)) : 0;
    }

    // ---------------- device allocation + upload ----------------
    auto bail = [&](cudaError_t e, const char* what) { return fail(GBP_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e)); };
    cudaError_t e;
    const size_t S = (size_t)g->n_slots;
    Shell shell;
    bool have_shell = false;   // an arena (and with the same shape: instantiated graphs) left by an earlier graph
    size_t zero_off = 0, zero_end = 0;
    for (int pass = 0; pass < 2; ++pass) {   // pass 0 measures, pass 1 carves
        Arena& A = g->arena;
        A.used = 0;
#define ALLOC(buf, n) g->buf.carve(A, n)
        // snapshot region: metrics | keyframe means | landmark means, contiguous -> ONE small device->host copy
        ALLOC(metric_out, 4); ALLOC(cam_mu, (size_t)C * 6); ALLOC(lmk_mu, (size_t)L * 3);
        g->snap_bytes = A.used;
        ALLOC(cam_belief, (size_t)C * CAM_B); ALLOC(lmk_belief, (size_t)L * LMK_B);
        // upload region: the static tables of the graph, contiguous -> ONE host->device copy from a page-locked block
        g->upload_off = A.used;
        ALLOC(tiles, tiles.size()); ALLOC(lmk_idx, S); ALLOC(z, S * 2); ALLOC(slot_of_factor, (size_t)F);
        ALLOC(lmk_ptr, (size_t)L + 1); ALLOC(lmk_slots, (size_t)F); ALLOC(cam_tile_ptr, (size_t)C + 1); ALLOC(cam_tiles, tiles.size());
        ALLOC(cam_mu0, (size_t)C * 6); ALLOC(lmk_mu0, (size_t)L * 3); ALLOC(cam_chunk_ptr, plan.cam_chunk_ptr.size());
        ALLOC(tile_pos, plan.tile_pos.size());
        g->upload_bytes = A.used - g->upload_off;
        // zero region: everything gbp_ba_reset clears, contiguous -> ONE memset
        zero_off = A.used;
        ALLOC(msg_cam, S * (size_t)g->cam_w); ALLOC(msg_lmk, S * LMK_M);
        ALLOC(cam_prior, (size_t)C * CAM_M); ALLOC(lmk_prior, (size_t)L * LMK_M); ALLOC(cam_partial, (size_t)g->K_chunks * C * CAM_M);
        ALLOC(tile_partial, tiles.size() * CAM_M); ALLOC(edge_max, S); ALLOC(tile_max, tiles.size()); ALLOC(cam_max, (size_t)C);
        ALLOC(cam_chol, (size_t)C * CHOL6);
        zero_end = A.used;
        ALLOC(iters, S); ALLOC(flags, S); ALLOC(linpoint, S * 9); ALLOC(sigma2a, S);
        ALLOC(tile_metric, tiles.size() * 3);
#undef ALLOC
        if (pass == 0) {
            A.size = A.used;
            const ShapeKey key = make_key(g);
            have_shell = cache_take(key, device, A.size, &shell);
            if (have_shell) {
                A.base = shell.arena; A.size = shell.arena_size;
                g->stage = shell.stage; g->stage_bytes = shell.stage_bytes;
                g->graphs = std::move(shell.graphs); g->snap_graphs = std::move(shell.snap_graphs); g->snap_ptr = shell.snap_ptr;
                shell.arena = nullptr; shell.stage = nullptr;
            } else if ((e = cudaMalloc(reinterpret_cast<void**>(&A.base), A.size)) != cudaSuccess) {
                A.base = nullptr;
                return bail(e, "cudaMalloc arena");
            }
        }
    }
    g->zero_off = zero_off;
    g->zero_bytes = zero_end - zero_off;
    {
        struct Up { const void* src; size_t bytes; void* dst; };
        const Up ups[] = {
            {tiles.data(), tiles.size() * sizeof(Tile), g->tiles.p}, {plan.lmk_idx.data(), plan.lmk_idx.size() * 4, g->lmk_idx.p},
            {plan.z.data(), plan.z.size() * 8, g->z.p}, {g->h_slot_of_factor.data(), g->h_slot_of_factor.size() * 4, g->slot_of_factor.p},
            {plan.lmk_ptr.data(), plan.lmk_ptr.size() * 4, g->lmk_ptr.p}, {plan.lmk_slots.data(), plan.lmk_slots.size() * 4, g->lmk_slots.p},
            {plan.cam_tile_ptr.data(), plan.cam_tile_ptr.size() * 4, g->cam_tile_ptr.p}, {plan.cam_tiles.data(), plan.cam_tiles.size() * 4, g->cam_tiles.p},
            {cam_mu0, (size_t)C * 48, g->cam_mu0.p}, {lmk_mu0, (size_t)L * 24, g->lmk_mu0.p},
            {plan.cam_chunk_ptr.data(), plan.cam_chunk_ptr.size() * 4, g->cam_chunk_ptr.p},
            {plan.tile_pos.data(), plan.tile_pos.size() * 4, g->tile_pos.p}};
        constexpr size_t STAGE_MAX = size_t(32) << 20;   // larger graphs upload table by table (a page-locked block that size costs more than it saves)
        if (g->upload_bytes <= STAGE_MAX) {
            if (g->stage_bytes < g->upload_bytes) {
                if (g->stage) cudaFreeHost(g->stage);
                g->stage = nullptr;
                g->stage_bytes = 0;
                const size_t want = std::max(g->upload_bytes, size_t(1) << 20);
                if ((e = cudaHostAlloc(&g->stage, want, cudaHostAllocDefault)) != cudaSuccess) { g->stage = nullptr; return bail(e, "cudaHostAlloc staging block"); }
                g->stage_bytes = want;
            }
            char* up0 = g->arena.base + g->upload_off;
            for (const Up& u : ups)
                if (u.bytes) memcpy(static_cast<char*>(g->stage) + (static_cast<char*>(u.dst) - up0), u.src, u.bytes);
            if ((e = cudaMemcpyAsync(up0, g->stage, g->upload_bytes, cudaMemcpyHostToDevice, g->stream)) != cudaSuccess) return bail(e, "upload of the graph tables");
        } else {
            for (const Up& u : ups)
                if (u.bytes && (e = cudaMemcpyAsync(u.dst, u.src, u.bytes, cudaMemcpyHostToDevice, g->stream)) != cudaSuccess) return bail(e, "upload of a graph table");
        }
    }
    {
        int rc = gbp_ba_reset(g);
        if (rc != GBP_OK) { return rc; }
    }
    *out = owner.release();
    return GBP_OK;
}


// No C++ exception may cross the C ABI: the graph compiler allocates host vectors proportional to the graph.
int gbp_ba_create(const gbp_config* cfg, int32_t C, int32_t L, int64_t F, const int32_t* cam_id, const int32_t* lmk_id,
                  const double* z, const double* cam_mu0, const double* lmk_mu0, const double K[4], int device,
                  void* stream, gbp_handle* out) {
    try {
        return ba_create_impl(cfg, C, L, F, cam_id, lmk_id, z, cam_mu0, lmk_mu0, K, device, stream, out);
    } catch (const std::bad_alloc&) {
        return fail(GBP_ERR_INVALID, "out of host memory while compiling the graph (C=%d L=%d F=%lld)", C, L, (long long)F);
    } catch (const std::exception& e) {
        return fail(GBP_ERR_INVALID, "graph compiler: %s", e.what());
    }
}

int gbp_cache_configure(int32_t max_shells, int64_t max_arena_bytes) {
    if (max_shells < 0 || max_arena_bytes < 0) return fail(GBP_ERR_INVALID, "negative cache limit");
    std::vector<Shell> drop;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        g_cache_max_shells = max_shells;
        g_cache_max_arena = (size_t)max_arena_bytes;
        for (size_t i = 0; i < g_shells.size();) {
            if (g_shells[i].arena_size > g_cache_max_arena) {
                drop.push_back(std::move(g_shells[i]));
                g_shells.erase(g_shells.begin() + i);
            } else {
                ++i;
            }
        }
        while ((int)g_shells.size() > g_cache_max_shells) {
            drop.push_back(std::move(g_shells.front()));
            g_shells.erase(g_shells.begin());
        }
    }
    for (Shell& sh : drop) { cudaSetDevice(sh.device); sh.free_all(); }
    return GBP_OK;
}

int gbp_cache_stats(int64_t out[6]) {
    if (!out) return fail(GBP_ERR_INVALID, "null out");
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (int i = 0; i < 4; ++i) out[i] = g_cache_stats[i];
    out[4] = (int64_t)g_shells.size();
    size_t bytes = 0;
    for (const Shell& sh : g_shells) bytes += sh.arena_size;
    out[5] = (int64_t)bytes;
    return GBP_OK;
}

int gbp_ba_reset(gbp_handle h) {
    CHECK_H(h);
    gbp_ba_graph* g = h;
    if (g->zero_bytes) CU(cudaMemsetAsync(g->arena.base + g->zero_off, 0, g->zero_bytes, g->stream));   // messages, priors, partial sums, prior scans
    // beliefs: eta = Lambda = 0, mu = initial means; edges linearised at those means
    if (g->C > 0) init_belief_kernel<<<(g->C + 127) / 128, 128, 0, g->stream>>>(g->cam_mu0.p, g->C, 6, CAM_B, g->cam_belief.p, g->cam_mu.p);
    if (g->L > 0) init_belief_kernel<<<(g->L + 127) / 128, 128, 0, g->stream>>>(g->lmk_mu0.p, g->L, 3, LMK_B, g->lmk_belief.p, g->lmk_mu.p);
    CU(cudaGetLastError());
    if (g->n_tiles > 0) {
        int rc = DISPATCH_T(g, launch_init_t);
        if (rc != GBP_OK) return rc;
    }
    g->priors_set = false;
    return GBP_OK;
}

int gbp_ba_destroy(gbp_handle h) {
    if (!h) return GBP_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    delete h;
    return GBP_OK;
}

int gbp_ba_sizes(gbp_handle h, int64_t out[6]) {
    if (!h || !out) return fail(GBP_ERR_INVALID, "null argument");
    out[0] = h->C; out[1] = h->L; out[2] = h->F; out[3] = h->n_tiles; out[4] = h->T; out[5] = h->n_slots;
    return GBP_OK;
}

int gbp_ba_layout(gbp_handle h, int64_t out[4]) {
    if (!h || !out) return fail(GBP_ERR_INVALID, "null argument");
    out[0] = h->cam_w;                 // doubles per stored factor->keyframe message: 27 (full) or 18 (factored)
    out[1] = h->pf_dist;               // L2 prefetch distance in tiles (0 = off)
    out[2] = h->streaming ? 2 : 1;     // sweep kernel build in use (gbp_config.kernel_variant after the automatic choice)
    out[3] = h->K_chunks;              // landmark chunks of the keyframe-side sums
    return GBP_OK;
}

int gbp_ba_prior_scan(gbp_handle h, double* cam_max) {
    CHECK_H(h);
    if (h->n_tiles > 0) {
        int rc = DISPATCH_T(h, launch_lammax_t);
        if (rc != GBP_OK) return rc;
    }
    if (h->C > 0) {
        cam_max_kernel<<<(h->C + 127) / 128, 128, 0, h->stream>>>(h->tile_max.p, h->cam_tile_ptr.p, h->cam_tiles.p, h->C, h->cam_max.p);
        CU(cudaGetLastError());
        if (h->comm) {   // the maximum over the factors of a keyframe on EVERY rank (gbp/gbp_ba.py:25-30 over the whole graph)
            int rc = comm_all_reduce(h, h->cam_max.p, (size_t)h->C, /*max*/ true);
            if (rc != GBP_OK) return rc;
        }
        if (cam_max) CU(cudaMemcpyAsync(cam_max, h->cam_max.p, (size_t)h->C * 8, cudaMemcpyDeviceToHost, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    return GBP_OK;
}

int gbp_ba_generate_priors(gbp_handle h, double weaker_factor, const double* cam_max) {
    CHECK_H(h);
    if (!(weaker_factor > 0)) return fail(GBP_ERR_INVALID, "weaker_factor must be > 0");
    if (cam_max) {
        // caller already ran gbp_ba_prior_scan on every rank and reduced the maxima
        if (h->C > 0) CU(cudaMemcpyAsync(h->cam_max.p, cam_max, (size_t)h->C * 8, cudaMemcpyHostToDevice, h->stream));
    } else {
        int rc = gbp_ba_prior_scan(h, nullptr);
        if (rc != GBP_OK) return rc;
    }
    if (h->C > 0) cam_prior_kernel<<<(h->C + 127) / 128, 128, 0, h->stream>>>(h->cam_max.p, weaker_factor, h->C, h->cam_belief.p, h->cam_prior.p);
    if (h->L > 0) lmk_prior_kernel<<<(h->L + 127) / 128, 128, 0, h->stream>>>(h->edge_max.p, h->lmk_ptr.p, h->lmk_slots.p, weaker_factor, h->L, h->lmk_belief.p, h->lmk_prior.p);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->stream));
    h->priors_set = true;
    return GBP_OK;
}

int gbp_ba_set_priors(gbp_handle h, const double* cam_lam, const double* lmk_lam) {
    CHECK_H(h);
    if ((h->C > 0 && !cam_lam) || (h->L > 0 && !lmk_lam)) return fail(GBP_ERR_INVALID, "null prior precision");
    if (h->C > 0) {
        CU(cudaMemcpy2DAsync(h->cam_prior.p + 6, CAM_M * 8, cam_lam, 21 * 8, 21 * 8, h->C, cudaMemcpyHostToDevice, h->stream));
        prior_eta_kernel<6><<<(h->C + 127) / 128, 128, 0, h->stream>>>(h->C, h->cam_belief.p, CAM_B, 27, h->cam_prior.p);
    }
    if (h->L > 0) {
        CU(cudaMemcpy2DAsync(h->lmk_prior.p + 3, LMK_M * 8, lmk_lam, 6 * 8, 6 * 8, h->L, cudaMemcpyHostToDevice, h->stream));
        prior_eta_kernel<3><<<(h->L + 127) / 128, 128, 0, h->stream>>>(h->L, h->lmk_belief.p, LMK_B, 9, h->lmk_prior.p);
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->stream));
    h->priors_set = true;
    return GBP_OK;
}

int gbp_ba_scale_priors(gbp_handle h, double factor) {
    CHECK_H(h);
    const long long nc = (long long)h->C * CAM_M, nl = (long long)h->L * LMK_M;
    if (nc > 0) scale_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, h->stream>>>(h->cam_prior.p, nc, factor);
    if (nl > 0) scale_kernel<<<(unsigned)((nl + 255) / 256), 256, 0, h->stream>>>(h->lmk_prior.p, nl, factor);
    CU(cudaGetLastError());
    return GBP_OK;
}

int gbp_ba_sweep_local(gbp_handle h, int stages) {
    CHECK_H(h);
    if (stages & (ST_ROBUSTIFY | ST_RELIN | ST_MESSAGES | ST_BELIEFS)) {
        // the sweep kernel also produces the per-tile keyframe sums when beliefs are requested
        int rc = launch_sweep(h, stages);
        if (rc != GBP_OK) return rc;
    }
    if (stages & ST_BELIEFS) return launch_belief(h, 0, (stages & GBP_STAGE_DEFER_LANDMARKS) ? 1 : 3);
    return GBP_OK;
}

int gbp_ba_landmark_update(gbp_handle h) {
    CHECK_H(h);
    return launch_belief(h, 0, 2);
}

int gbp_ba_cam_update(gbp_handle h, const double* partials_dev, int nranks) {
    CHECK_H(h);
    if (h->C == 0) return GBP_OK;
    const double* src = partials_dev ? partials_dev : h->cam_partial.p;
    if (!partials_dev) nranks = h->K_chunks;
    if (nranks < 1) return fail(GBP_ERR_INVALID, "the number of partial sums must be >= 1");
    cam_update_kernel<<<(h->C + 3) / 4, 128, 0, h->stream>>>(src, nranks, h->C, h->cam_prior.p, h->cam_belief.p, h->cam_mu.p, h->cam_chol.p);
    h->launches++;
    CU(cudaGetLastError());
    return GBP_OK;
}

int gbp_ba_iterate(gbp_handle h, int n_iters, int robustify, int local_relin) {
    CHECK_H(h);
    if (n_iters < 0) return fail(GBP_ERR_INVALID, "n_iters < 0");
    if (!h->priors_set) return fail(GBP_ERR_STATE, "priors not set: call gbp_ba_generate_priors / gbp_ba_set_priors first");
    const int st = iteration_stages(robustify, local_relin);
    const int per_iter = launches_per_iteration(h);
    int left = n_iters;
    if (h->n_tiles <= 8192) {
        // small graphs only: there the gaps between graph launches are a visible share of a ~9 us iteration, so long runs replay
        // graphs of 32 and 8 iterations (fr1desk: 10.3 us per iteration with one iteration per launch, 9.1 with 8, 8.x with 32)
        for (int reps : {32, 8}) {
            if (left < reps) continue;
            cudaGraphExec_t exec;
            int rc = get_graph(h, st, &exec, reps);
            if (rc != GBP_OK) return rc;
            for (; left >= reps; left -= reps) CU(cudaGraphLaunch(exec, h->stream));
        }
    }
    if (left > 0) {
        cudaGraphExec_t exec;
        int rc = get_graph(h, st, &exec);
        if (rc != GBP_OK) return rc;
        for (; left > 0; --left) CU(cudaGraphLaunch(exec, h->stream));
    }
    h->launches += (long long)per_iter * n_iters;
    return GBP_OK;
}

int gbp_ba_update_beliefs(gbp_handle h) {
    CHECK_H(h);
    int rc = launch_sweep(h, ST_BELIEFS);  // only the per-tile sums of the stored messages
    if (rc != GBP_OK) return rc;
    return enqueue_beliefs(h);
}

namespace {
int enqueue_metrics(gbp_ba_graph* h) {
    if (h->n_tiles > 0) {
        int rc = DISPATCH_T(h, launch_metric_t);
        if (rc != GBP_OK) return rc;
        reduce_rows_kernel<3><<<1, 256, 0, h->stream>>>(h->tile_metric.p, h->n_tiles, h->metric_out.p);
        h->launches += 2;
        CU(cudaGetLastError());
    } else {
        CU(cudaMemsetAsync(h->metric_out.p, 0, 3 * sizeof(double), h->stream));
    }
    if (h->comm) return comm_all_reduce(h, h->metric_out.p, 3, /*max*/ false);
    return GBP_OK;
}
}  // namespace

int gbp_ba_metrics(gbp_handle h, double out[3]) {
    CHECK_H(h);
    if (!out) return fail(GBP_ERR_INVALID, "null out");
    out[0] = out[1] = out[2] = 0.0;
    if (h->n_tiles == 0 && !h->comm) return GBP_OK;
    int rc = enqueue_metrics(h);   // with a communicator: summed over the ranks (a rank without edges still joins the sum)
    if (rc != GBP_OK) return rc;
    CU(cudaMemcpyAsync(out, h->metric_out.p, 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return GBP_OK;
}


int gbp_ba_snapshot_layout(gbp_handle h, uint64_t out[4]) {
    if (!h || !out) return fail(GBP_ERR_INVALID, "null argument");
    out[0] = h->snap_bytes;
    out[1] = (uint64_t)((char*)h->metric_out.p - h->arena.base);
    out[2] = (uint64_t)((char*)h->cam_mu.p - h->arena.base);
    out[3] = (uint64_t)((char*)h->lmk_mu.p - h->arena.base);
    return GBP_OK;
}

int gbp_ba_snapshot_async(gbp_handle h, void* region) {
    CHECK_H(h);
    if (!region) return fail(GBP_ERR_INVALID, "null region");
    if (!h->snap_event) CU(cudaEventCreateWithFlags(&h->snap_event, cudaEventDisableTiming));
    int rc = enqueue_metrics(h);
    if (rc != GBP_OK) return rc;
    CU(cudaMemcpyAsync(region, h->arena.base, h->snap_bytes, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaEventRecord(h->snap_event, h->stream));
    return GBP_OK;
}

int gbp_ba_iterate_snapshot(gbp_handle h, int robustify, int local_relin, void* region) {
    CHECK_H(h);
    if (!region) return fail(GBP_ERR_INVALID, "null region");
    if (!h->priors_set) return fail(GBP_ERR_STATE, "priors not set: call gbp_ba_generate_priors / gbp_ba_set_priors first");
    if (!h->snap_event) CU(cudaEventCreateWithFlags(&h->snap_event, cudaEventDisableTiming));
    const int st = iteration_stages(robustify, local_relin);
    if (h->snap_ptr != region) {
        for (auto& kv : h->snap_graphs) cudaGraphExecDestroy(kv.second);
        h->snap_graphs.clear();
        h->snap_ptr = region;
    }
    auto it = h->snap_graphs.find(st);
    cudaGraphExec_t exec = nullptr;
    if (it != h->snap_graphs.end()) {
        exec = it->second;
    } else {
        cudaGraph_t graph = nullptr;
        CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        const long long before = h->launches;
        int rc = enqueue_iteration(h, st);
        if (rc == GBP_OK) rc = enqueue_metrics(h);
        cudaError_t e = cudaSuccess;
        if (rc == GBP_OK) e = cudaMemcpyAsync(region, h->arena.base, h->snap_bytes, cudaMemcpyDeviceToHost, h->stream);
        h->launches = before;
        cudaError_t e2 = cudaStreamEndCapture(h->stream, &graph);
        if (rc != GBP_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (e != cudaSuccess || e2 != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            return fail(GBP_ERR_CUDA, "capture of the iteration+snapshot graph failed: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
        }
        e = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) return fail(GBP_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
        h->snap_graphs[st] = exec;
    }
    CU(cudaGraphLaunch(exec, h->stream));
    h->launches += launches_per_iteration(h) + (h->n_tiles > 0 ? 2 : 0);
    CU(cudaEventRecord(h->snap_event, h->stream));
    return GBP_OK;
}

int gbp_ba_snapshot_wait(gbp_handle h) {
    CHECK_H(h);
    if (h->snap_event) CU(cudaEventSynchronize(h->snap_event));
    return GBP_OK;
}

void* gbp_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (gbp_device_count() <= 0) return nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void gbp_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int gbp_ba_read(gbp_handle h, int field, void* host_dst, size_t bytes) {
    CHECK_H(h);
    FieldInfo fi;
    if (!field_info(field, &fi)) return fail(GBP_ERR_INVALID, "unknown field %d", field);
    const long long rows = field == GBP_F_CAM_PARTIAL ? (long long)h->K_chunks * h->C : fi.indexed == 0 ? h->C : fi.indexed == 1 ? h->L : h->F;
    const size_t need = (size_t)rows * fi.words * 4;
    if (bytes != need) return fail(GBP_ERR_INVALID, "field %d: expected %zu bytes, got %zu", field, need, bytes);
    if (need == 0) return GBP_OK;
    if (!host_dst) return fail(GBP_ERR_INVALID, "null destination");
    if (field == GBP_F_ADJ) { memcpy(host_dst, h->h_adj.data(), need); return GBP_OK; }
    if (field == GBP_F_FILE_INDEX) { memcpy(host_dst, h->h_file_of_factor.data(), need); return GBP_OK; }
    if (fi.indexed != 2) {
        CU(cudaMemcpyAsync(host_dst, field_dev_ptr(h, field), need, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        return GBP_OK;
    }
    DevBuf<uint32_t> tmp;
    CU(tmp.alloc((size_t)rows * fi.words));
    const long long n = rows * fi.words;
    if (field == GBP_F_MSG_CAM && h->cam_w == CAM_MF) {
        export_msg_cam_factored_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, h->stream>>>(reinterpret_cast<double*>(tmp.p), h->msg_cam.p,
                                                                                         h->slot_of_factor.p, rows);
    } else if (field == GBP_F_JACOBIAN_B) {
        export_jb_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, h->stream>>>(h->slot_of_factor.p, rows, h->linpoint.p, h->z.p, h->K,
                                                                            reinterpret_cast<double*>(tmp.p));
    } else {
        gather_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(tmp.p, reinterpret_cast<const uint32_t*>(field_dev_ptr(h, field)),
                                                                           h->slot_of_factor.p, rows, fi.words);
    }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(host_dst, tmp.p, need, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    tmp.release();
    if (e != cudaSuccess) return fail(GBP_ERR_CUDA, "read field %d: %s", field, cudaGetErrorString(e));
    return GBP_OK;
}

int gbp_ba_write(gbp_handle h, int field, const void* host_src, size_t bytes) {
    CHECK_H(h);
    FieldInfo fi;
    if (!field_info(field, &fi)) return fail(GBP_ERR_INVALID, "unknown field %d", field);
    if (!fi.writable) return fail(GBP_ERR_INVALID, "field %d is read-only", field);
    const long long rows = fi.indexed == 0 ? h->C : fi.indexed == 1 ? h->L : h->F;
    const size_t need = (size_t)rows * fi.words * 4;
    if (bytes != need) return fail(GBP_ERR_INVALID, "field %d: expected %zu bytes, got %zu", field, need, bytes);
    if (need == 0) return GBP_OK;
    if (!host_src) return fail(GBP_ERR_INVALID, "null source");
    if (fi.indexed != 2) {
        CU(cudaMemcpyAsync(field_dev_ptr(h, field), host_src, need, cudaMemcpyHostToDevice, h->stream));
        if (field == GBP_F_CAM_BELIEF) refresh_mu_kernel<6><<<(h->C + 127) / 128, 128, 0, h->stream>>>(h->cam_belief.p, h->C, CAM_B, h->cam_mu.p, h->cam_chol.p);
        if (field == GBP_F_LMK_BELIEF) refresh_mu_kernel<3><<<(h->L + 127) / 128, 128, 0, h->stream>>>(h->lmk_belief.p, h->L, LMK_B, h->lmk_mu.p, nullptr);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(h->stream));
        if (field == GBP_F_CAM_PRIOR || field == GBP_F_LMK_PRIOR) h->priors_set = true;
        return GBP_OK;
    }
    DevBuf<uint32_t> tmp;
    CU(tmp.alloc((size_t)rows * fi.words));
    const long long n = rows * fi.words;
    cudaError_t e = cudaMemcpyAsync(tmp.p, host_src, need, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
        if (field == GBP_F_MSG_CAM && h->cam_w == CAM_MF)
            import_msg_cam_factored_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, h->stream>>>(h->msg_cam.p, reinterpret_cast<const double*>(tmp.p),
                                                                                             h->slot_of_factor.p, rows);
        else
            scatter_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(reinterpret_cast<uint32_t*>(field_dev_ptr(h, field)), tmp.p,
                                                                                h->slot_of_factor.p, rows, fi.words);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    tmp.release();
    if (e != cudaSuccess) return fail(GBP_ERR_CUDA, "write field %d: %s", field, cudaGetErrorString(e));
    return GBP_OK;
}

int gbp_ba_fill_iters(gbp_handle h, int32_t value) {
    CHECK_H(h);
    if (value < 0) return fail(GBP_ERR_INVALID, "iters_since_relin must be >= 0");
    if (h->n_slots > 0) {
        fill_iters_kernel<<<(unsigned)((h->n_slots + 255) / 256), 256, 0, h->stream>>>(h->iters.p, h->n_slots, value);
        h->launches++;
        CU(cudaGetLastError());
    }
    return GBP_OK;
}

int gbp_ba_device_ptr(gbp_handle h, int field, void** dev_ptr, size_t* bytes) {
    if (!h || !dev_ptr) return fail(GBP_ERR_INVALID, "null argument");
    FieldInfo fi;
    if (!field_info(field, &fi) || fi.indexed == 2) return fail(GBP_ERR_INVALID, "field %d has no stable device layout", field);
    *dev_ptr = field_dev_ptr(h, field);
    if (bytes) *bytes = (size_t)(fi.indexed == 0 ? h->C : h->L) * fi.words * 4 * (field == GBP_F_CAM_PARTIAL ? (size_t)h->K_chunks : 1);
    return GBP_OK;
}

int gbp_ba_set_params(gbp_handle h, double eta_damping, double beta, int32_t num_undamped_iters, int32_t min_linear_iters) {
    CHECK_H(h);
    CU(cudaStreamSynchronize(h->stream));
    h->cfg.eta_damping = eta_damping; h->cfg.beta = beta;
    h->cfg.num_undamped_iters = num_undamped_iters; h->cfg.min_linear_iters = min_linear_iters;
    for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second);  // parameters are baked into captured launches
    h->graphs.clear();
    for (auto& kv : h->snap_graphs) cudaGraphExecDestroy(kv.second);
    h->snap_graphs.clear();
    return GBP_OK;
}

int gbp_ba_synchronize(gbp_handle h) {
    CHECK_H(h);
    CU(cudaStreamSynchronize(h->stream));
    return GBP_OK;
}

int gbp_ba_tune(gbp_handle h, int knob, int64_t value) {
    CHECK_H(h);
    switch (knob) {
        case GBP_TUNE_BELIEF_LANES:
            if (value != 0 && value != 1 && value != 8 && value != 32) return fail(GBP_ERR_INVALID, "lanes per landmark: 0 (automatic), 1, 8 or 32");
            CU(cudaStreamSynchronize(h->stream));
            h->belief_lanes = (int)value;
            for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second);
            h->graphs.clear();
            for (auto& kv : h->snap_graphs) cudaGraphExecDestroy(kv.second);
            h->snap_graphs.clear();
            return GBP_OK;
        case GBP_TUNE_PREFETCH_TILES:
            if (value < 0) return fail(GBP_ERR_INVALID, "prefetch distance must be >= 0");
            CU(cudaStreamSynchronize(h->stream));
            h->pf_dist = h->streaming ? (int)value : 0;
            for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second);      // the distance is baked into captured launches
            h->graphs.clear();
            for (auto& kv : h->snap_graphs) cudaGraphExecDestroy(kv.second);
            h->snap_graphs.clear();
            return GBP_OK;
        default:
            return fail(GBP_ERR_INVALID, "unknown tuning knob %d", knob);
    }
}

int gbp_ba_time_iterations(gbp_handle h, int n_iters, int robustify, int local_relin, int per_kernel, float* ms_total,
                           float* ms_msg_kernel) {
    CHECK_H(h);
    if (n_iters <= 0) return fail(GBP_ERR_INVALID, "n_iters must be > 0");
    if (!h->priors_set) return fail(GBP_ERR_STATE, "priors not set");
    const int st = iteration_stages(robustify, local_relin);
    if (ms_total) *ms_total = 0.f;
    if (ms_msg_kernel) *ms_msg_kernel = 0.f;
    if (!per_kernel) {
        struct Pair {
            cudaEvent_t e0 = nullptr, e1 = nullptr;
            ~Pair() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
        } ev;
        CU(cudaEventCreate(&ev.e0)); CU(cudaEventCreate(&ev.e1));
        cudaGraphExec_t exec;
        int rc = get_graph(h, st, &exec);
        if (rc != GBP_OK) return rc;
        CU(cudaStreamSynchronize(h->stream));
        CU(cudaEventRecord(ev.e0, h->stream));
        for (int i = 0; i < n_iters; ++i) CU(cudaGraphLaunch(exec, h->stream));
        CU(cudaEventRecord(ev.e1, h->stream));
        CU(cudaEventSynchronize(ev.e1));
        h->launches += (long long)launches_per_iteration(h) * n_iters;
        if (ms_total) CU(cudaEventElapsedTime(ms_total, ev.e0, ev.e1));
        return GBP_OK;
    }
    struct Events {      // destroyed on every return path
        std::vector<cudaEvent_t> v;
        ~Events() { for (cudaEvent_t e : v) if (e) cudaEventDestroy(e); }
    } evs;
    try {
        evs.v.assign((size_t)2 * n_iters + 2, nullptr);
    } catch (const std::bad_alloc&) {
        return fail(GBP_ERR_INVALID, "out of host memory");
    }
    std::vector<cudaEvent_t>& ev = evs.v;
    for (auto& e : ev) CU(cudaEventCreate(&e));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaEventRecord(ev[2 * n_iters], h->stream));
    for (int i = 0; i < n_iters; ++i) {
        CU(cudaEventRecord(ev[2 * i], h->stream));
        int rc = launch_sweep(h, st);
        if (rc != GBP_OK) return rc;
        CU(cudaEventRecord(ev[2 * i + 1], h->stream));
        rc = enqueue_beliefs(h);
        if (rc != GBP_OK) return rc;
    }
    CU(cudaEventRecord(ev[2 * n_iters + 1], h->stream));
    CU(cudaEventSynchronize(ev[2 * n_iters + 1]));
    float tot = 0.f, acc = 0.f;
    CU(cudaEventElapsedTime(&tot, ev[2 * n_iters], ev[2 * n_iters + 1]));
    for (int i = 0; i < n_iters; ++i) {
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, ev[2 * i], ev[2 * i + 1]));
        acc += t;
    }
    if (ms_total) *ms_total = tot;
    if (ms_msg_kernel) *ms_msg_kernel = acc;
    return GBP_OK;
}

int64_t gbp_ba_launch_count(gbp_handle h) { return h ? h->launches : 0; }

int gbp_reprojection_eval(const double* x, int64_t n, const double K[4], int device, double* out_h, double* out_J) {
    if (!x || !K || !out_h || !out_J || n < 0) return fail(GBP_ERR_INVALID, "bad argument");
    if (gbp_device_count() <= 0) return fail(GBP_ERR_NO_DEVICE, "no CUDA device visible: gbp_b200 has no CPU fallback");
    if (n == 0) return GBP_OK;
    CU(cudaSetDevice(device));
    DevBuf<double> dx, dh, dj;
    cudaError_t e = dx.alloc((size_t)n * 9);
    if (e == cudaSuccess) e = dh.alloc((size_t)n * 2);
    if (e == cudaSuccess) e = dj.alloc((size_t)n * 18);
    if (e == cudaSuccess) e = cudaMemcpy(dx.p, x, (size_t)n * 72, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        reprojection_eval_kernel<<<(unsigned)((n + 127) / 128), 128>>>(dx.p, n, Intrinsics{K[0], K[1], K[2], K[3]}, dh.p, dj.p);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out_h, dh.p, (size_t)n * 16, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(out_J, dj.p, (size_t)n * 144, cudaMemcpyDeviceToHost);
    dx.release(); dh.release(); dj.release();
    if (e != cudaSuccess) return fail(GBP_ERR_CUDA, "reprojection_eval: %s", cudaGetErrorString(e));
    return GBP_OK;
}


// ---- the host graph compiler on its own (no GPU needed): what gbp_ba_create lays out, for inspection and tests
struct gbp_plan_s {
    GraphPlan plan;
    int C = 0, L = 0;
    int64_t F = 0;
};

int gbp_plan_create(int32_t tile_edges, int32_t lmk_block, const int64_t* chunking, int32_t C, int32_t L, int64_t F, const int32_t* cam_id,
                    const int32_t* lmk_id, gbp_plan* out) {
    if (!out) return fail(GBP_ERR_INVALID, "null argument");
    *out = nullptr;
    if (C < 0 || L < 0 || F < 0) return fail(GBP_ERR_INVALID, "negative size");
    if (F > 0 && (!cam_id || !lmk_id)) return fail(GBP_ERR_INVALID, "null measurement arrays");
    try {
        std::unique_ptr<gbp_plan_s> p(new gbp_plan_s());
        p->C = C; p->L = L; p->F = F;
        int T = 32;
        long long lblock = 1;
        int rc = choose_tiling(tile_edges, lmk_block, L, F, &T, &lblock);
        std::vector<long long> cb;
        gbp_config cc{};
        if (chunking) {
            cc.lmk_chunks = (int32_t)chunking[0]; cc.lmk_chunk_first = (int32_t)chunking[1]; cc.lmk_chunks_total = (int32_t)chunking[2];
            cc.lmk_first = chunking[3]; cc.lmk_total = chunking[4];
        }
        if (rc == GBP_OK) rc = chunk_bounds_of(&cc, L, &cb);
        if (rc == GBP_OK) rc = plan_graph(T, lblock, cb, C, L, F, cam_id, lmk_id, nullptr, &p->plan);
        if (rc != GBP_OK) return rc;
        *out = p.release();
        return GBP_OK;
    } catch (const std::bad_alloc&) {
        return fail(GBP_ERR_INVALID, "out of host memory while compiling the graph");
    }
}

int gbp_plan_sizes(gbp_plan p, int64_t out[6]) {
    if (!p || !out) return fail(GBP_ERR_INVALID, "null argument");
    out[0] = p->C; out[1] = p->L; out[2] = p->F; out[3] = (int64_t)p->plan.tiles.size(); out[4] = p->plan.T; out[5] = p->plan.n_slots();
    return GBP_OK;
}

int gbp_plan_copy(gbp_plan p, int32_t* tiles, int32_t* slot_of_factor, int32_t* file_of_factor, int32_t* adj, int32_t* lmk_idx,
                  int32_t* lmk_ptr, int32_t* lmk_slots, int32_t* cam_tile_ptr, int32_t* cam_tiles) {
    if (!p) return fail(GBP_ERR_INVALID, "null plan");
    const GraphPlan& g = p->plan;
    auto cp = [](int32_t* dst, const std::vector<int>& v) { if (dst && !v.empty()) memcpy(dst, v.data(), v.size() * sizeof(int)); };
    if (tiles) for (size_t t = 0; t < g.tiles.size(); ++t) { tiles[2 * t] = g.tiles[t].cam; tiles[2 * t + 1] = g.tiles[t].count; }
    cp(slot_of_factor, g.slot_of_factor); cp(file_of_factor, g.file_of_factor); cp(adj, g.adj); cp(lmk_idx, g.lmk_idx);
    cp(lmk_ptr, g.lmk_ptr); cp(lmk_slots, g.lmk_slots); cp(cam_tile_ptr, g.cam_tile_ptr); cp(cam_tiles, g.cam_tiles);
    return GBP_OK;
}

int gbp_plan_chunks(gbp_plan p, int32_t* tile_chunk, int32_t* cam_chunk_ptr) {
    if (!p) return 0;
    const GraphPlan& g = p->plan;
    if (tile_chunk && !g.tile_chunk.empty()) memcpy(tile_chunk, g.tile_chunk.data(), g.tile_chunk.size() * sizeof(int));
    if (cam_chunk_ptr && !g.cam_chunk_ptr.empty()) memcpy(cam_chunk_ptr, g.cam_chunk_ptr.data(), g.cam_chunk_ptr.size() * sizeof(int));
    return g.n_chunks;
}

void gbp_plan_destroy(gbp_plan p) { delete p; }

}  // extern "C"

#include "gbp_dist.cu.inc"
#include "gbp_bal.cpp.inc"
#include "gbp_lin.cu.inc"
