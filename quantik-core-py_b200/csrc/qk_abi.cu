// qk_abi.cu -- the extern "C" boundary of libquantik_b200.so (include/quantik_b200.h):
// argument checks, per-device context, host<->device staging.  No rules live here.
#include "qk_common.cuh"

#include <atomic>
#include <mutex>
#include <string.h>
#include <stdio.h>

namespace qk {

static thread_local std::string g_error;
static Ctx g_ctx[64];
static std::mutex g_mu;
static std::atomic<long long> g_option[OPT_COUNT];
static std::atomic<int> g_stage_busy[64];

long long option(Option o) { return g_option[o].load(std::memory_order_relaxed); }

void set_error(const std::string &msg) { g_error = msg; }

int cuda_fail(cudaError_t e, const char *what)
{
    g_error = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    cudaGetLastError();   // clear the sticky-less error
    return QK_E_CUDA;
}

Ctx *get_ctx(int dev)
{
    if (dev < 0 || dev >= 64 || !g_ctx[dev].ok) {
        set_error("device " + std::to_string(dev) + " not initialised: call qk_init(dev) first");
        return nullptr;
    }
    return &g_ctx[dev];
}

// ------------------------------------------------------------------ Call
static bool is_device_pointer(const void *p)
{
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

Call::Call(int dev, void *stream) : stream_((cudaStream_t)stream)
{
    ctx_ = get_ctx(dev);
    if (!ctx_) { status_ = QK_E_NO_DEVICE; return; }
    cudaGetDevice(&prev_dev_);
    cudaError_t e = cudaSetDevice(dev);
    if (e != cudaSuccess) { status_ = cuda_fail(e, "cudaSetDevice"); ctx_ = nullptr; return; }
    // the small-call stage, if nobody else holds it; work an earlier holder left running on another stream goes first
    int expected = 0;
    if (ctx_->stage_dev && option(OPT_CALL_STAGE) != 1 && g_stage_busy[dev].compare_exchange_strong(expected, 1)) {
        staged_ = true;
        if (ctx_->stage_pending && ctx_->stage_stream != stream_) cudaStreamWaitEvent(stream_, ctx_->stage_ev, 0);
    }
}

Call::~Call()
{
    for (void *t : temps_) cudaFreeAsync(t, stream_);
    temps_.clear();
    if (staged_) {
        if (stage_used_ && !synced_) {
            if (stage_host_in_flight_) cudaStreamSynchronize(stream_);      // an error path left before finish(): the pinned block is still being read
            else { cudaEventRecord(ctx_->stage_ev, stream_); ctx_->stage_stream = stream_; ctx_->stage_pending = true; }
        } else if (stage_used_) ctx_->stage_pending = false;
        g_stage_busy[ctx_->dev].store(0);
    }
    if (prev_dev_ >= 0 && ctx_ && prev_dev_ != ctx_->dev) cudaSetDevice(prev_dev_);
}

char *Call::stage_alloc(size_t bytes)
{
    if (!staged_) return nullptr;
    const size_t need = (bytes + 255) & ~(size_t)255;
    if (need == 0 || need > Ctx::kStageBytes - stage_off_) return nullptr;
    char *d = ctx_->stage_dev + stage_off_;
    stage_off_ += need;
    stage_used_ = true;
    return d;
}

void *Call::alloc(size_t bytes)
{
    void *d = nullptr;
    cudaError_t e = cudaMallocAsync(&d, bytes ? bytes : 16, stream_);
    if (e != cudaSuccess) { status_ = cuda_fail(e, "cudaMallocAsync"); return nullptr; }
    temps_.push_back(d);
    return d;
}

const void *Call::in(const void *p, size_t bytes)
{
    if (!p || status_ != QK_OK) return nullptr;
    if (is_device_pointer(p)) return p;
    any_host_ = true;
    if (char *sd = stage_alloc(bytes)) {          // through the pinned block: one asynchronous copy, no allocation
        char *sh = ctx_->stage_host + (sd - ctx_->stage_dev);
        memcpy(sh, p, bytes);
        stage_host_in_flight_ = true;
        cudaError_t e = cudaMemcpyAsync(sd, sh, bytes, cudaMemcpyHostToDevice, stream_);
        if (e != cudaSuccess) { status_ = cuda_fail(e, "cudaMemcpyAsync(H2D)"); return nullptr; }
        return sd;
    }
    void *d = alloc(bytes);
    if (!d) return nullptr;
    cudaError_t e = cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, stream_);
    if (e != cudaSuccess) { status_ = cuda_fail(e, "cudaMemcpyAsync(H2D)"); return nullptr; }
    return d;
}

void *Call::out(void *p, size_t bytes, bool zero)
{
    if (!p || status_ != QK_OK) return nullptr;
    void *d = p;
    if (!is_device_pointer(p)) {
        any_host_ = true;
        const bool staged = (d = stage_alloc(bytes)) != nullptr;
        if (!staged) d = alloc(bytes);
        if (!d) return nullptr;
        backs_.push_back({ p, d, bytes, staged });
    }
    if (zero) {
        cudaError_t e = cudaMemsetAsync(d, 0, bytes, stream_);
        if (e != cudaSuccess) { status_ = cuda_fail(e, "cudaMemsetAsync"); return nullptr; }
    }
    return d;
}

void *Call::temp(size_t bytes, bool zero)
{
    if (status_ != QK_OK) return nullptr;
    void *d = stage_alloc(bytes);
    if (!d) d = alloc(bytes);
    if (d && zero) {
        cudaError_t e = cudaMemsetAsync(d, 0, bytes, stream_);
        if (e != cudaSuccess) { status_ = cuda_fail(e, "cudaMemsetAsync"); return nullptr; }
    }
    return d;
}

void Call::release(void *p)
{
    if (in_stage(p)) return;                  // a block of the stage: nothing to free
    for (size_t i = 0; i < temps_.size(); ++i)
        if (temps_[i] == p) {
            temps_[i] = temps_.back();
            temps_.pop_back();
            cudaFreeAsync(p, stream_);
            return;
        }
}

int Call::sync()
{
    QK_CUDA(cudaStreamSynchronize(stream_));
    return QK_OK;
}

int Call::finish()
{
    if (status_ != QK_OK) return status_;
    QK_CUDA(cudaGetLastError());
    std::vector<Back> backs;
    backs.swap(backs_);
    // staged outputs come back in ONE copy into the pinned block (everything between the first and the last of
    // them), and go to the caller's buffers after the synchronisation below
    size_t lo = Ctx::kStageBytes, hi = 0;
    for (const Back &b : backs) {
        if (!b.staged) { QK_CUDA(cudaMemcpyAsync(b.host, b.dev, b.bytes, cudaMemcpyDeviceToHost, stream_)); continue; }
        const size_t at = (size_t)((char *)b.dev - ctx_->stage_dev);
        lo = at < lo ? at : lo;
        hi = at + b.bytes > hi ? at + b.bytes : hi;
    }
    if (hi > lo) {
        stage_host_in_flight_ = true;
        QK_CUDA(cudaMemcpyAsync(ctx_->stage_host + lo, ctx_->stage_dev + lo, hi - lo, cudaMemcpyDeviceToHost, stream_));
    }
    std::vector<void *> temps;
    temps.swap(temps_);                       // whatever happens below, ~Call must not free these a second time
    cudaError_t first = cudaSuccess;
    for (void *t : temps) {
        const cudaError_t e = cudaFreeAsync(t, stream_);
        if (first == cudaSuccess) first = e;
    }
    QK_CUDA(first);
    if (any_host_) {
        QK_CUDA(cudaStreamSynchronize(stream_));
        synced_ = true;
        for (const Back &b : backs)
            if (b.staged) memcpy(b.host, ctx_->stage_host + ((char *)b.dev - ctx_->stage_dev), b.bytes);
    }
    return QK_OK;
}

}  // namespace qk

using namespace qk;

#define QK_REQUIRE(cond, msg)                     \
    do {                                          \
        if (!(cond)) { set_error(msg); return QK_E_BAD_ARG; } \
    } while (0)
#define QK_BEGIN(call)                            \
    Call call(dev, stream);                       \
    if (!call.valid()) return call.status()
#define QK_CHECK(call)                            \
    do { if (call.status() != QK_OK) return call.status(); } while (0)

extern "C" {

int qk_abi_version(void) { return QK_ABI_VERSION; }

const char *qk_last_error(void) { return g_error.c_str(); }

int qk_set_option(const char *name, long long value)
{
    static const char *const names[OPT_COUNT] = { "perft_kernel", "perft_min_items", "perft_tile_generic", "playout_variant",
                                                  "dedup_mode", "dedup_part_mb", "playout_grab", "grid_waves", "dedup_streams", "table_slots_x10", "guided_scoring", "call_stage", "perft_own_lanes" };
    QK_REQUIRE(name, "qk_set_option: name is NULL");
    for (int i = 0; i < OPT_COUNT; ++i)
        if (strcmp(name, names[i]) == 0) { g_option[i].store(value, std::memory_order_relaxed); return QK_OK; }
    set_error(std::string("qk_set_option: unknown option '") + name + "'");
    return QK_E_BAD_ARG;
}

int qk_init(int dev)
{
    std::lock_guard<std::mutex> lock(g_mu);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        set_error(std::string("no CUDA device available (") + (e != cudaSuccess ? cudaGetErrorString(e) : "count = 0") +
                  "); libquantik_b200 has no CPU fallback");
        return QK_E_NO_DEVICE;
    }
    QK_REQUIRE(dev >= 0 && dev < count && dev < 64, "qk_init: device ordinal out of range");
    if (g_ctx[dev].ok) return QK_OK;
    cudaDeviceProp prop;
    QK_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) {
        set_error(std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major * 10 + prop.minor) +
                  "; this library carries sm_100a code only");
        return QK_E_NO_DEVICE;
    }
    int prev = 0;
    cudaGetDevice(&prev);
    QK_CUDA(cudaSetDevice(dev));
    Ctx &c = g_ctx[dev];
    c.dev = dev;
    c.sms = prop.multiProcessorCount;
    QK_CUDA(cudaDeviceGetAttribute(&c.clock_khz, cudaDevAttrClockRate, dev));
    QK_CUDA(cudaDeviceGetDefaultMemPool(&c.pool, dev));
    uint64_t keep = UINT64_MAX;   // the pool is the scratch arena: never hand memory back between calls
    QK_CUDA(cudaMemPoolSetAttribute(c.pool, cudaMemPoolAttrReleaseThreshold, &keep));
    QK_CUDA(cudaMalloc((void **)&c.stage_dev, Ctx::kStageBytes));
    QK_CUDA(cudaHostAlloc((void **)&c.stage_host, Ctx::kStageBytes, cudaHostAllocDefault));
    QK_CUDA(cudaEventCreateWithFlags(&c.stage_ev, cudaEventDisableTiming));
    g_stage_busy[dev].store(0);
    c.ok = true;
    cudaSetDevice(prev);
    return QK_OK;
}

int qk_shutdown(int dev)
{
    std::lock_guard<std::mutex> lock(g_mu);
    QK_REQUIRE(dev >= 0 && dev < 64, "qk_shutdown: device ordinal out of range");
    if (!g_ctx[dev].ok) return QK_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(dev);
    cudaDeviceSynchronize();
    for (int i = 0; i < Ctx::kAux; ++i) {
        if (g_ctx[dev].aux[i]) cudaStreamDestroy(g_ctx[dev].aux[i]);
        if (g_ctx[dev].join_ev[i]) cudaEventDestroy(g_ctx[dev].join_ev[i]);
    }
    if (g_ctx[dev].fork_ev) cudaEventDestroy(g_ctx[dev].fork_ev);
    if (g_ctx[dev].stage_ev) cudaEventDestroy(g_ctx[dev].stage_ev);
    if (g_ctx[dev].stage_dev) cudaFree(g_ctx[dev].stage_dev);
    if (g_ctx[dev].stage_host) cudaFreeHost(g_ctx[dev].stage_host);
    cudaMemPoolTrimTo(g_ctx[dev].pool, 0);
    g_ctx[dev] = Ctx();
    cudaSetDevice(prev);
    return QK_OK;
}

int qk_device_info(int dev, int *sm_count, int *sm_clock_khz)
{
    Ctx *c = get_ctx(dev);
    if (!c) return QK_E_NO_DEVICE;
    if (sm_count) *sm_count = c->sms;
    if (sm_clock_khz) *sm_clock_khz = c->clock_khz;
    return QK_OK;
}

int qk_movegen_masks(const uint16_t *states, size_t n, uint64_t *action_mask, uint8_t *to_move, uint8_t *status,
                     int dev, void *stream)
{
    QK_REQUIRE(states || n == 0, "qk_movegen_masks: states is NULL");
    if (n == 0) return QK_OK;
    QK_BEGIN(call);
    const uint4 *d_states = (const uint4 *)call.in(states, n * 16);
    uint64_t *d_mask = (uint64_t *)call.out(action_mask, n * 8);
    uint8_t *d_tm = (uint8_t *)call.out(to_move, n);
    uint8_t *d_st = (uint8_t *)call.out(status, n);
    QK_CHECK(call);
    int rc = launch_movegen(d_states, n, d_mask, d_tm, d_st, call.ctx(), call.stream());
    if (rc != QK_OK) return rc;
    return call.finish();
}

int qk_apply_moves(const uint16_t *states, const uint8_t *action, size_t n, uint16_t *out, int dev, void *stream)
{
    QK_REQUIRE((states && action && out) || n == 0, "qk_apply_moves: NULL argument");
    if (n == 0) return QK_OK;
    QK_BEGIN(call);
    const uint4 *d_states = (const uint4 *)call.in(states, n * 16);
    const uint8_t *d_act = (const uint8_t *)call.in(action, n);
    uint4 *d_out = (uint4 *)call.out(out, n * 16);
    QK_CHECK(call);
    int rc = launch_apply(d_states, d_act, n, d_out, call.ctx(), call.stream());
    if (rc != QK_OK) return rc;
    return call.finish();
}

int qk_win_status(const uint16_t *states, size_t n, uint8_t *win_status, int dev, void *stream)
{
    QK_REQUIRE((states && win_status) || n == 0, "qk_win_status: NULL argument");
    if (n == 0) return QK_OK;
    QK_BEGIN(call);
    const uint4 *d_states = (const uint4 *)call.in(states, n * 16);
    uint8_t *d_out = (uint8_t *)call.out(win_status, n);
    QK_CHECK(call);
    int rc = launch_win(d_states, n, d_out, call.ctx(), call.stream());
    if (rc != QK_OK) return rc;
    return call.finish();
}

int qk_canonical(const uint16_t *states, size_t n, int color_swap, uint16_t *payload, uint16_t *transform,
                 int dev, void *stream)
{
    QK_REQUIRE((states && payload) || n == 0, "qk_canonical: NULL argument");
    if (n == 0) return QK_OK;
    QK_BEGIN(call);
    const uint4 *d_states = (const uint4 *)call.in(states, n * 16);
    uint4 *d_out = (uint4 *)call.out(payload, n * 16);
    uint16_t *d_tr = (uint16_t *)call.out(transform, n * 2);
    QK_CHECK(call);
    int rc = launch_canonical(d_states, n, color_swap, d_out, d_tr, call.ctx(), call.stream());
    if (rc != QK_OK) return rc;
    return call.finish();
}

int qk_orbit_size(const uint16_t *states, size_t n, uint8_t *orbit, int dev, void *stream)
{
    QK_REQUIRE((states && orbit) || n == 0, "qk_orbit_size: NULL argument");
    if (n == 0) return QK_OK;
    QK_BEGIN(call);
    const uint4 *d_states = (const uint4 *)call.in(states, n * 16);
    uint8_t *d_out = (uint8_t *)call.out(orbit, n);
    QK_CHECK(call);
    int rc = launch_orbit(d_states, n, d_out, call.ctx(), call.stream());
    if (rc != QK_OK) return rc;
    return call.finish();
}

int qk_apply_symmetry(const uint16_t *states, const uint16_t *transform, size_t n, uint16_t *out, int dev, void *stream)
{
    QK_REQUIRE((states && transform && out) || n == 0, "qk_apply_symmetry: NULL argument");
    if (n == 0) return QK_OK;
    QK_BEGIN(call);
    const uint4 *d_states = (const uint4 *)call.in(states, n * 16);
    const uint16_t *d_tr = (const uint16_t *)call.in(transform, n * 2);
    uint4 *d_out = (uint4 *)call.out(out, n * 16);
    QK_CHECK(call);
    int rc = launch_apply_symmetry(d_states, d_tr, n, d_out, call.ctx(), call.stream());
    if (rc != QK_OK) return rc;
    return call.finish();
}

int qk_apply_symmetry_to_moves(const uint8_t *players, const uint8_t *actions, const uint16_t *transform, size_t n, int inverse,
                               uint8_t *out_players, uint8_t *out_actions, int dev, void *stream)
{
    QK_REQUIRE((players && actions && transform && out_players && out_actions) || n == 0, "qk_apply_symmetry_to_moves: NULL argument");
    if (n == 0) return QK_OK;
    QK_BEGIN(call);
    const uint8_t *d_pl = (const uint8_t *)call.in(players, n);
    const uint8_t *d_ac = (const uint8_t *)call.in(actions, n);
    const uint16_t *d_tr = (const uint16_t *)call.in(transform, n * 2);
    uint8_t *d_opl = (uint8_t *)call.out(out_players, n);
    uint8_t *d_oac = (uint8_t *)call.out(out_actions, n);
    QK_CHECK(call);
    int rc = launch_apply_symmetry_to_moves(d_pl, d_ac, d_tr, n, inverse, d_opl, d_oac, call.ctx(), call.stream());
    if (rc != QK_OK) return rc;
    return call.finish();
}

int qk_canon_dedup(const uint16_t *states, const uint64_t *mult, size_t n, int color_swap, uint16_t *uniq,
                   uint64_t *uniq_mult, size_t cap, size_t *n_uniq, int dev, void *stream)
{
    QK_REQUIRE(n_uniq, "qk_canon_dedup: n_uniq is NULL");
    QK_REQUIRE(states || n == 0, "qk_canon_dedup: states is NULL");
    *n_uniq = 0;
    if (n == 0) return QK_OK;
    QK_BEGIN(call);
    const uint4 *d_states = (const uint4 *)call.in(states, n * 16);
    const uint64_t *d_mult = (const uint64_t *)call.in(mult, n * 8);
    QK_CHECK(call);
    int rc = run_dedup(call, d_states, d_mult, n, color_swap, uniq, uniq_mult, cap ? cap : n, n_uniq);
    if (rc != QK_OK) return rc;
    return call.finish();
}

int qk_bfs_level(const uint16_t *states, const uint64_t *mult, size_t n, int color_swap, int sort_output, uint16_t *uniq,
                 uint64_t *uniq_mult, size_t cap, size_t *n_uniq, uint64_t *stats, int dev, void *stream)
{
    QK_REQUIRE(n_uniq && stats, "qk_bfs_level: n_uniq / stats is NULL");
    QK_REQUIRE(states || n == 0, "qk_bfs_level: states is NULL");
    QK_REQUIRE(cap > 0, "qk_bfs_level: cap must be positive");
    *n_uniq = 0;
    for (int i = 0; i < 5; ++i) stats[i] = 0;
    if (n == 0) return QK_OK;
    QK_BEGIN(call);
    const uint4 *d_states = (const uint4 *)call.in(states, n * 16);
    const uint64_t *d_mult = (const uint64_t *)call.in(mult, n * 8);
    QK_CHECK(call);
    int rc = run_bfs_level(call, d_states, d_mult, n, color_swap, sort_output, uniq, uniq_mult, cap, n_uniq, stats);
    if (rc != QK_OK) return rc;
    return call.finish();
}

int qk_book_level(const uint16_t *states, size_t n, int color_swap, uint16_t *uniq, uint64_t *uniq_indegree, size_t cap,
                  size_t *n_uniq, uint64_t *stats, int dev, void *stream)
{
    QK_REQUIRE(n_uniq && stats && uniq && uniq_indegree, "qk_book_level: NULL output");
    QK_REQUIRE(states || n == 0, "qk_book_level: states is NULL");
    QK_REQUIRE(cap > 0, "qk_book_level: cap must be positive");
    *n_uniq = 0;
    stats[0] = stats[1] = 0;
    if (n == 0) return QK_OK;
    QK_BEGIN(call);
    const uint4 *d_states = (const uint4 *)call.in(states, n * 16);
    QK_CHECK(call);
    int rc = run_book_level(call, d_states, n, color_swap, uniq, uniq_indegree, cap, n_uniq, stats);
    if (rc != QK_OK) return rc;
    return call.finish();
}

static int make_peers(void *const *inbox, int world, int rank, size_t slots_per_src, Peers &p)
{
    QK_REQUIRE(inbox, "inbox pointer array is NULL");
    QK_REQUIRE(world >= 1 && world <= QK_MAX_WORLD, "world must be 1..QK_MAX_WORLD");
    QK_REQUIRE(rank >= 0 && rank < world, "rank out of range");
    QK_REQUIRE(slots_per_src > 0, "slots_per_src must be positive");
    for (int r = 0; r < QK_MAX_WORLD; ++r) p.inbox[r] = r < world ? inbox[r] : nullptr;
    for (int r = 0; r < world; ++r) QK_REQUIRE(inbox[r], "an inbox pointer is NULL");
    p.world = world; p.rank = rank; p.slots_per_src = slots_per_src;
    return QK_OK;
}

int qk_inbox_bytes(int world, size_t slots_per_src, size_t *bytes)
{
    QK_REQUIRE(bytes, "qk_inbox_bytes: bytes is NULL");
    QK_REQUIRE(world >= 1 && world <= QK_MAX_WORLD, "world must be 1..QK_MAX_WORLD");
    *bytes = inbox_bytes(world, slots_per_src);
    return QK_OK;
}

int qk_canon_dedup_scatter(const uint16_t *states, const uint64_t *mult, size_t n, int flags, void *const *inbox, int world,
                           int rank, size_t slots_per_src, uint64_t *sent, int dev, void *stream)
{
    QK_REQUIRE(states || n == 0, "qk_canon_dedup_scatter: states is NULL");
    Peers peers;
    int rc = make_peers(inbox, world, rank, slots_per_src, peers);
    if (rc != QK_OK) return rc;
    QK_BEGIN(call);
    const uint4 *d_states = (const uint4 *)call.in(states, n * 16);
    const uint64_t *d_mult = (const uint64_t *)call.in(mult, n * 8);
    QK_CHECK(call);
    rc = run_dedup_scatter(call, d_states, d_mult, n, flags, peers, sent);      // n == 0 still publishes zero counts
    if (rc != QK_OK) return rc;
    return call.finish();
}

int qk_bfs_level_scatter(const uint16_t *states, const uint64_t *mult, size_t n, int flags, size_t local_cap,
                         void *const *inbox, int world, int rank, size_t slots_per_src, uint64_t *stats, uint64_t *sent,
                         int dev, void *stream)
{
    QK_REQUIRE(stats, "qk_bfs_level_scatter: stats is NULL");
    QK_REQUIRE(states || n == 0, "qk_bfs_level_scatter: states is NULL");
    QK_REQUIRE(local_cap > 0, "qk_bfs_level_scatter: local_cap must be positive");
    for (int i = 0; i < 5; ++i) stats[i] = 0;
    Peers peers;
    int rc = make_peers(inbox, world, rank, slots_per_src, peers);
    if (rc != QK_OK) return rc;
    QK_BEGIN(call);
    const uint4 *d_states = (const uint4 *)call.in(states, n * 16);
    const uint64_t *d_mult = (const uint64_t *)call.in(mult, n * 8);
    QK_CHECK(call);
    rc = run_bfs_level_scatter(call, d_states, d_mult, n, flags, local_cap, peers, stats, sent);
    if (rc != QK_OK) return rc;
    return call.finish();
}

int qk_inbox_merge(const void *my_inbox, int world, size_t slots_per_src, int sort_output, uint16_t *uniq, uint64_t *uniq_mult,
                   size_t cap, size_t *n_uniq, int dev, void *stream)
{
    QK_REQUIRE(my_inbox && n_uniq && uniq && uniq_mult, "qk_inbox_merge: NULL argument");
    QK_REQUIRE(world >= 1 && world <= QK_MAX_WORLD, "world must be 1..QK_MAX_WORLD");
    QK_REQUIRE(slots_per_src > 0 && cap > 0, "qk_inbox_merge: slots_per_src and cap must be positive");
    *n_uniq = 0;
    QK_BEGIN(call);
    int rc = run_inbox_merge(call, my_inbox, world, slots_per_src, sort_output, uniq, uniq_mult, cap, n_uniq);
    if (rc != QK_OK) return rc;
    return call.finish();
}

int qk_playouts_ex(const uint16_t *roots, size_t n_roots, uint32_t rollouts_per_root, uint64_t seed,
                   uint64_t first_rollout, uint32_t root_index_base, int max_plies, uint32_t *wins_p0,
                   uint32_t *wins_p1, uint32_t *draws, int8_t *per_playout, uint8_t *per_plies, int dev, void *stream)
{
    QK_REQUIRE(roots || n_roots == 0, "qk_playouts: roots is NULL");
    QK_REQUIRE(n_roots < (1ull << 32), "qk_playouts: n_roots must be < 2^32");
    if (n_roots == 0) return QK_OK;
    QK_BEGIN(call);
    size_t total = n_roots * (size_t)rollouts_per_root;
    const uint4 *d_roots = (const uint4 *)call.in(roots, n_roots * 16);
    uint32_t *d_w0 = (uint32_t *)call.out(wins_p0, n_roots * 4, true);
    uint32_t *d_w1 = (uint32_t *)call.out(wins_p1, n_roots * 4, true);
    uint32_t *d_dr = (uint32_t *)call.out(draws, n_roots * 4, true);
    int8_t *d_pp = (int8_t *)call.out(per_playout, total);
    uint8_t *d_pl = (uint8_t *)call.out(per_plies, total);
    void *scratch = call.temp(playout_scratch_bytes(n_roots));
    QK_CHECK(call);
    if (rollouts_per_root == 0) return call.finish();          // tallies are zero
    int rc = launch_playouts(d_roots, n_roots, rollouts_per_root, seed, first_rollout, root_index_base, max_plies,
                             d_w0, d_w1, d_dr, d_pp, d_pl, scratch, call.ctx(), call.stream());
    if (rc != QK_OK) return rc;
    return call.finish();
}

int qk_playouts(const uint16_t *roots, size_t n_roots, uint32_t rollouts_per_root, uint64_t seed, int max_plies,
                uint32_t *wins_p0, uint32_t *wins_p1, uint32_t *draws, int8_t *per_playout, int dev, void *stream)
{
    return qk_playouts_ex(roots, n_roots, rollouts_per_root, seed, 0, 0, max_plies, wins_p0, wins_p1, draws,
                          per_playout, nullptr, dev, stream);
}

int qk_eval_features(const uint16_t *states, const uint8_t *player, size_t n, double *features, int dev, void *stream)
{
    QK_REQUIRE((states && player && features) || n == 0, "qk_eval_features: NULL argument");
    if (n == 0) return QK_OK;
    QK_BEGIN(call);
    const uint4 *d_states = (const uint4 *)call.in(states, n * 16);
    const uint8_t *d_pl = (const uint8_t *)call.in(player, n);
    double *d_out = (double *)call.out(features, n * 6 * sizeof(double));
    QK_CHECK(call);
    int rc = launch_eval_features(d_states, d_pl, n, d_out, call.ctx(), call.stream());
    if (rc != QK_OK) return rc;
    return call.finish();
}

int qk_playouts_guided(const uint16_t *roots, size_t n_roots, uint32_t rollouts_per_root, uint64_t seed, uint64_t first_rollout,
                       uint32_t root_index_base, int max_plies, double epsilon, const double *weights, uint32_t *wins_p0,
                       uint32_t *wins_p1, uint32_t *draws, int8_t *per_playout, uint8_t *per_plies, int dev, void *stream)
{
    QK_REQUIRE(roots || n_roots == 0, "qk_playouts_guided: roots is NULL");
    QK_REQUIRE(weights, "qk_playouts_guided: weights is NULL (6 doubles, host memory)");
    QK_REQUIRE(epsilon >= 0.0 && epsilon <= 1.0, "rollout_epsilon must be in [0.0, 1.0]");     // mcts.py:70-76
    QK_REQUIRE(n_roots < (1ull << 32), "qk_playouts_guided: n_roots must be < 2^32");
    if (n_roots == 0) return QK_OK;
    QK_BEGIN(call);
    size_t total = n_roots * (size_t)rollouts_per_root;
    const uint4 *d_roots = (const uint4 *)call.in(roots, n_roots * 16);
    uint32_t *d_w0 = (uint32_t *)call.out(wins_p0, n_roots * 4, true);
    uint32_t *d_w1 = (uint32_t *)call.out(wins_p1, n_roots * 4, true);
    uint32_t *d_dr = (uint32_t *)call.out(draws, n_roots * 4, true);
    int8_t *d_pp = (int8_t *)call.out(per_playout, total);
    uint8_t *d_pl = (uint8_t *)call.out(per_plies, total);
    void *d_lines = call.temp(guided_scratch_bytes(n_roots));
    QK_CHECK(call);
    if (rollouts_per_root == 0) return call.finish();
    int rc = launch_playouts_guided(d_roots, n_roots, rollouts_per_root, seed, first_rollout, root_index_base, max_plies, epsilon,
                                    weights, d_w0, d_w1, d_dr, d_pp, d_pl, d_lines, call.ctx(), call.stream());
    if (rc != QK_OK) return rc;
    return call.finish();
}

int qk_solve(const uint16_t *states, size_t n, int8_t *outcome, uint8_t *plies, uint8_t *best_action, int dev, void *stream)
{
    QK_REQUIRE((states && outcome && plies && best_action) || n == 0, "qk_solve: NULL argument");
    if (n == 0) return QK_OK;
    QK_BEGIN(call);
    const uint4 *d_states = (const uint4 *)call.in(states, n * 16);
    int8_t *d_oc = (int8_t *)call.out(outcome, n);
    uint8_t *d_pl = (uint8_t *)call.out(plies, n);
    uint8_t *d_ba = (uint8_t *)call.out(best_action, n);
    unsigned long long *counter = (unsigned long long *)call.temp(sizeof(unsigned long long));
    QK_CHECK(call);
    int rc = launch_solve(d_states, n, counter, d_oc, d_pl, d_ba, call.ctx(), call.stream());
    if (rc != QK_OK) return rc;
    return call.finish();
}

int qk_random_roots(uint64_t seed, uint64_t first, size_t n, int min_plies, int span, uint16_t *out_states,
                    int dev, void *stream)
{
    QK_REQUIRE(out_states || n == 0, "qk_random_roots: out_states is NULL");
    QK_REQUIRE(min_plies >= 0 && span >= 1 && min_plies + span - 1 <= 15, "qk_random_roots: plies must stay within 0..15");
    if (n == 0) return QK_OK;
    QK_BEGIN(call);
    uint4 *d_out = (uint4 *)call.out(out_states, n * 16);
    QK_CHECK(call);
    int rc = launch_random_roots(seed, first, n, min_plies, span, d_out, call.ctx(), call.stream());
    if (rc != QK_OK) return rc;
    return call.finish();
}

int qk_payload_owner(const uint16_t *payloads, size_t n, int world, uint8_t *owner, int dev, void *stream)
{
    QK_REQUIRE((payloads && owner) || n == 0, "qk_payload_owner: NULL argument");
    QK_REQUIRE(world >= 1 && world <= QK_MAX_WORLD, "qk_payload_owner: world out of range");
    if (n == 0) return QK_OK;
    QK_BEGIN(call);
    const uint4 *d_in = (const uint4 *)call.in(payloads, n * 16);
    uint8_t *d_out = (uint8_t *)call.out(owner, n);
    QK_CHECK(call);
    int rc = launch_payload_owner(d_in, n, world, d_out, call.ctx(), call.stream());
    if (rc != QK_OK) return rc;
    return call.finish();
}

int qk_perft_probe(const uint16_t *roots, size_t n_roots, uint32_t *cost, int dev, void *stream)
{
    QK_REQUIRE((roots && cost) || n_roots == 0, "qk_perft_probe: NULL argument");
    if (n_roots == 0) return QK_OK;
    QK_BEGIN(call);
    const uint4 *d_roots = (const uint4 *)call.in(roots, n_roots * 16);
    uint32_t *d_cost = (uint32_t *)call.out(cost, n_roots * 4);
    QK_CHECK(call);
    int rc = launch_perft_probe(d_roots, n_roots, d_cost, call.ctx(), call.stream());
    if (rc != QK_OK) return rc;
    return call.finish();
}

int qk_perft(const uint16_t *roots, const uint64_t *mult, size_t n_roots, int depth, uint64_t *nodes,
             uint64_t *wins_p0, uint64_t *wins_p1, uint64_t *blocked_p0_loses, uint64_t *blocked_p1_loses,
             int dev, void *stream)
{
    QK_REQUIRE(depth >= 0 && depth <= QK_MAX_DEPTH, "qk_perft: depth must be 0..16");
    QK_REQUIRE(nodes && wins_p0 && wins_p1 && blocked_p0_loses && blocked_p1_loses, "qk_perft: NULL output");
    QK_REQUIRE(roots || n_roots == 0, "qk_perft: roots is NULL");
    uint64_t host[5 * (QK_MAX_DEPTH + 1)] = { 0 };
    if (n_roots) {
        QK_BEGIN(call);
        const uint4 *d_roots = (const uint4 *)call.in(roots, n_roots * 16);
        const uint64_t *d_mult = (const uint64_t *)call.in(mult, n_roots * 8);
        QK_CHECK(call);
        int rc = run_perft(call, d_roots, d_mult, n_roots, depth, host);
        if (rc != QK_OK) return rc;
        rc = call.finish();
        if (rc != QK_OK) return rc;
    }
    for (int d = 0; d <= depth; ++d) {
        nodes[d] = host[0 * (QK_MAX_DEPTH + 1) + d];
        wins_p0[d] = host[1 * (QK_MAX_DEPTH + 1) + d];
        wins_p1[d] = host[2 * (QK_MAX_DEPTH + 1) + d];
        blocked_p0_loses[d] = host[3 * (QK_MAX_DEPTH + 1) + d];
        blocked_p1_loses[d] = host[4 * (QK_MAX_DEPTH + 1) + d];
    }
    return QK_OK;
}

int qk_int_issue_peak(int dev, double *alu_lane_ops_per_s, double *mixed_lane_ops_per_s)
{
    void *stream = nullptr;
    QK_BEGIN(call);
    double a = 0, m = 0;
    int rc = run_issue_peak(call.ctx(), &a, &m);
    if (rc != QK_OK) return rc;
    if (alu_lane_ops_per_s) *alu_lane_ops_per_s = a;
    if (mixed_lane_ops_per_s) *mixed_lane_ops_per_s = m;
    return QK_OK;
}

int qk_pipe_rates(int dev, double *lane_ops_per_s)
{
    void *stream = nullptr;
    QK_REQUIRE(lane_ops_per_s, "qk_pipe_rates: NULL output");
    QK_BEGIN(call);
    return run_pipe_rates(call.ctx(), lane_ops_per_s);
}

}  // extern "C"
