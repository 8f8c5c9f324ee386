// qk_common.cuh -- host-side plumbing shared by the translation units of
// libquantik_b200.so: per-device context, error text, host<->device staging.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <vector>

#include "../../include/quantik_b200.h"

namespace qk {

struct Ctx {
    bool ok = false;
    int dev = -1;
    int sms = 0;
    int clock_khz = 0;
    cudaMemPool_t pool = nullptr;
    // side streams of the two-level merge (created on first use, qk_dedup.cu: merge_buckets)
    static constexpr int kAux = 3;
    cudaStream_t aux[kAux] = { nullptr, nullptr, nullptr };
    cudaEvent_t fork_ev = nullptr, join_ev[kAux] = { nullptr, nullptr, nullptr };
    // small-call stage (qk_abi.cu: Call): one device buffer and one pinned host buffer of kStageBytes, handed to one
    // call at a time; a call whose host buffers and scratch fit makes no allocation and one copy each way
    static constexpr size_t kStageBytes = 256u << 10;
    char *stage_dev = nullptr, *stage_host = nullptr;
    cudaEvent_t stage_ev = nullptr;         // recorded by a call that left without synchronising
    cudaStream_t stage_stream = nullptr;    // ... on this stream
    bool stage_pending = false;
};

// Tuning / testing switches (qk_set_option in include/quantik_b200.h).  Process-wide, 0 = the default
// behaviour; the library never reads the environment.
enum Option { OPT_PERFT_KERNEL = 0, OPT_PERFT_MIN_ITEMS, OPT_PERFT_TILE_GENERIC, OPT_PLAYOUT_VARIANT, OPT_DEDUP_MODE, OPT_DEDUP_PART_MB, OPT_PLAYOUT_GRAB, OPT_GRID_WAVES, OPT_DEDUP_STREAMS, OPT_TABLE_SLOTS_X10, OPT_GUIDED_SCORING, OPT_CALL_STAGE, OPT_PERFT_OWN_LANES, OPT_COUNT };
long long option(Option o);

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what);          // records text, returns QK_E_CUDA
Ctx *get_ctx(int dev);                                   // nullptr (+error text) when not initialised

#define QK_CUDA(call)                                                     \
    do {                                                                  \
        cudaError_t e__ = (call);                                         \
        if (e__ != cudaSuccess) return ::qk::cuda_fail(e__, #call);       \
    } while (0)

// One API call: resolves every caller pointer to a device pointer, staging host
// buffers through the device's stream-ordered pool, and copies results back.
class Call {
public:
    Call(int dev, void *stream);
    ~Call();
    bool valid() const { return ctx_ != nullptr; }
    int status() const { return status_; }
    Ctx *ctx() const { return ctx_; }
    cudaStream_t stream() const { return stream_; }

    const void *in(const void *p, size_t bytes);          // nullptr stays nullptr
    void *out(void *p, size_t bytes, bool zero = false);  // nullptr stays nullptr
    void *temp(size_t bytes, bool zero = false);
    void release(void *temp_ptr);                         // early stream-ordered free of a temp()
    int finish();                                         // copy-backs, frees, sync if host buffers were used
    int sync();

private:
    struct Back { void *host; void *dev; size_t bytes; bool staged; };
    Ctx *ctx_ = nullptr;
    cudaStream_t stream_ = nullptr;
    int status_ = QK_OK;
    bool any_host_ = false;
    int prev_dev_ = -1;
    std::vector<void *> temps_;
    std::vector<Back> backs_;
    bool staged_ = false;            // this call holds the device's small-call stage
    bool stage_used_ = false, stage_host_in_flight_ = false, synced_ = false;
    size_t stage_off_ = 0;
    void *alloc(size_t bytes);
    char *stage_alloc(size_t bytes); // nullptr when the stage is not ours or the block does not fit
    bool in_stage(const void *p) const { return ctx_ && ctx_->stage_dev && (const char *)p >= ctx_->stage_dev && (const char *)p < ctx_->stage_dev + Ctx::kStageBytes; }
};

inline unsigned grid_for(size_t n, unsigned block, int sms, unsigned per_sm)
{
    size_t want = (n + block - 1) / block;
    const long long waves = option(OPT_GRID_WAVES);
    size_t cap = (size_t)sms * per_sm * (size_t)(waves > 0 ? waves : 16);   // 16 waves: CTAs finish at different times (unfair warp scheduling), later waves fill the gaps
    if (want < 1) want = 1;
    return (unsigned)(want < cap ? want : cap);
}

// launchers implemented in the kernel translation units (all asynchronous on `s`)
int launch_movegen(const uint4 *states, size_t n, uint64_t *mask, uint8_t *to_move, uint8_t *status, Ctx *c, cudaStream_t s);
int launch_apply(const uint4 *states, const uint8_t *action, size_t n, uint4 *out, Ctx *c, cudaStream_t s);
int launch_win(const uint4 *states, size_t n, uint8_t *win, Ctx *c, cudaStream_t s);
int launch_canonical(const uint4 *states, size_t n, int color_swap, uint4 *payload, uint16_t *transform, Ctx *c, cudaStream_t s);
int launch_orbit(const uint4 *states, size_t n, uint8_t *orbit, Ctx *c, cudaStream_t s);
int launch_apply_symmetry(const uint4 *states, const uint16_t *transform, size_t n, uint4 *out, Ctx *c, cudaStream_t s);
int launch_apply_symmetry_to_moves(const uint8_t *player, const uint8_t *action, const uint16_t *transform, size_t n, int inverse,
                                   uint8_t *out_player, uint8_t *out_action, Ctx *c, cudaStream_t s);
int launch_playouts(const uint4 *roots, size_t n_roots, uint32_t rollouts, uint64_t seed, uint64_t first_rollout,
                    uint32_t root_index_base, int max_plies, uint32_t *wins_p0, uint32_t *wins_p1, uint32_t *draws,
                    int8_t *per_playout, uint8_t *per_plies, void *rootrec_scratch, Ctx *c, cudaStream_t s);
size_t playout_scratch_bytes(size_t n_roots);
int launch_random_roots(uint64_t seed, uint64_t first, size_t n, int min_plies, int span, uint4 *out, Ctx *c, cudaStream_t s);
int launch_perft_probe(const uint4 *roots, size_t n, uint32_t *cost, Ctx *c, cudaStream_t s);
int run_perft(Call &call, const uint4 *roots, const uint64_t *mult, size_t n_roots, int depth, uint64_t *host_out /*5 x (depth+1)*/);
int run_dedup(Call &call, const uint4 *states, const uint64_t *mult, size_t n, int color_swap,
              uint16_t *uniq_host_or_dev, uint64_t *uniq_mult, size_t cap, size_t *n_uniq);
int run_bfs_level(Call &call, const uint4 *parents, const uint64_t *mult, size_t n, int color_swap, int sort_output,
                  uint16_t *uniq, uint64_t *uniq_mult, size_t cap, size_t *n_uniq, uint64_t *h_stats);
int run_book_level(Call &call, const uint4 *parents, size_t n, int color_swap, uint16_t *uniq, uint64_t *uniq_mult,
                   size_t cap, size_t *n_uniq, uint64_t *h_stats);
struct Peers { void *inbox[QK_MAX_WORLD]; int world; int rank; size_t slots_per_src; };
size_t inbox_bytes(int world, size_t slots_per_src);
int launch_payload_owner(const uint4 *payloads, size_t n, int world, uint8_t *owner, Ctx *c, cudaStream_t s);
int run_dedup_scatter(Call &call, const uint4 *states, const uint64_t *mult, size_t n, int flags, const Peers &peers, uint64_t *h_sent);
int run_bfs_level_scatter(Call &call, const uint4 *parents, const uint64_t *mult, size_t n, int flags, size_t local_cap,
                          const Peers &peers, uint64_t *h_stats, uint64_t *h_sent);
int run_inbox_merge(Call &call, const void *my_inbox, int world, size_t slots_per_src, int sort_output,
                    uint16_t *uniq, uint64_t *uniq_mult, size_t cap, size_t *n_uniq);
int launch_eval_features(const uint4 *states, const uint8_t *player, size_t n, double *features, Ctx *c, cudaStream_t s);
size_t guided_scratch_bytes(size_t n_roots);
int launch_playouts_guided(const uint4 *roots, size_t n_roots, uint32_t rollouts, uint64_t seed, uint64_t first_rollout,
                           uint32_t root_index_base, int max_plies, double epsilon, const double *weights, uint32_t *wins_p0,
                           uint32_t *wins_p1, uint32_t *draws, int8_t *per_playout, uint8_t *per_plies, void *line_scratch, Ctx *c,
                           cudaStream_t s);
int launch_solve(const uint4 *states, size_t n, unsigned long long *counter, int8_t *outcome, uint8_t *plies, uint8_t *best_action,
                 Ctx *c, cudaStream_t s);
int run_issue_peak(Ctx *c, double *alu, double *mixed);
int run_pipe_rates(Ctx *c, double *out8);

}  // namespace qk
