// qk_playout.cu -- batched uniform-random playouts (BASELINE configs 1 and 4).
//
// k_prepare_roots : root (16 B) -> RootRec (48 B): incremental rule state + result
//                   already known at the root (won / invalid).
// k_playout       : one playout per thread at a time, state in registers
//                   (qk_playout_core.cuh).  The flat playout space
//                   [0, n_roots * rollouts) is cut into one contiguous chunk per
//                   thread.  A thread whose game ends picks up its next playout at the
//                   next 4-ply boundary, where the whole warp computes one Philox4x32-10
//                   block (4 words = 4 plies) anyway -- so the refill never diverges the
//                   RNG code, and a playout's moves stay a pure function of
//                   (seed, root, rollout, ply).
//                   Tallies are kept per thread while the root stays the same and are
//                   flushed with warp-aggregated atomics.
// k_random_roots  : the config-4 mid-game root generator.
#include "qk_common.cuh"
#include "qk_playout_core.cuh"

#include <stdlib.h>

namespace qk {

static constexpr unsigned kBlock = 256;
static constexpr unsigned kFull = 0xFFFFFFFFu;

struct PlayoutArgs {
    const RootRec *roots;
    uint64_t total;          // n_roots * rollouts
    uint64_t chunk;          // playouts per thread
    uint64_t seed;
    uint64_t first_rollout;
    uint32_t rollouts;
    uint32_t root_index_base;
    int max_plies;
    uint32_t *wins_p0, *wins_p1, *draws;
    int8_t *per_playout;
    uint8_t *per_plies;
    uint32_t one, zero;      // 1 and 0 the compiler cannot see (k_playout_pred: keeps small additions on the FMA pipe)
    uint32_t *queue;         // k_playout_pred: next playout nobody has taken yet (zeroed before the launch)
    uint32_t grab;           // k_playout_pred: playouts a lane takes from the queue at a time
};

__global__ void __launch_bounds__(kBlock) k_prepare_roots(const uint4 *__restrict__ states, size_t n, RootRec *__restrict__ out)
{
    size_t i = blockIdx.x * (size_t)kBlock + threadIdx.x;
    if (i >= n) return;
    uint4 v = __ldg(states + i);
    State4 s = { v.x, v.y, v.z, v.w };
    out[i] = make_root(s);
}

__device__ __forceinline__ void fill_tables(SquareLut *lut, uint8_t *sel8)
{
    if (threadIdx.x < 16) lut[threadIdx.x] = square_lut((int)threadIdx.x);
    for (uint32_t i = threadIdx.x; i < 2048; i += blockDim.x) sel8[i] = (uint8_t)sel8_entry(i);
    __syncthreads();
}

// add a lane's tallies for `root`; lanes of the calling (converged or not) group that
// share a root are summed first so that one lane per root issues the atomics
__device__ __forceinline__ void flush_tallies(const PlayoutArgs &a, uint32_t root, uint32_t n, uint32_t p0, uint32_t dr)
{
    unsigned active = __activemask();
    unsigned peers = __match_any_sync(active, root);
    uint32_t sn = __reduce_add_sync(peers, n), sp0 = __reduce_add_sync(peers, p0), sdr = __reduce_add_sync(peers, dr);
    if ((unsigned)(__ffs(peers) - 1) == (threadIdx.x & 31u) && sn) {
        if (a.wins_p0 && sp0) atomicAdd(a.wins_p0 + root, sp0);
        if (a.wins_p1 && sn - sp0 - sdr) atomicAdd(a.wins_p1 + root, sn - sp0 - sdr);
        if (a.draws && sdr) atomicAdd(a.draws + root, sdr);
    }
}

#ifdef QK_PLAYOUT_MAXNREG      // experiment knob: cap registers to raise occupancy (see NOTES.md)
#define QK_PLAYOUT_BOUNDS __maxnreg__(QK_PLAYOUT_MAXNREG)
#else
#define QK_PLAYOUT_BOUNDS __launch_bounds__(kBlock)
#endif
template <bool CUTOFF, bool PER_PLAYOUT>
__global__ void QK_PLAYOUT_BOUNDS k_playout(const PlayoutArgs a)
{
    __shared__ __align__(16) SquareLut s_lut[16];
    __shared__ __align__(16) uint8_t s_sel8[2048];
    fill_tables(s_lut, s_sel8);

    const uint64_t tid = blockIdx.x * (uint64_t)kBlock + threadIdx.x;
    const uint64_t begin = tid * a.chunk;
    uint32_t remaining = 0;               // playouts this thread still has to START
    uint32_t root = 0, j = 0;             // next playout to start = (root, j)
    if (begin < a.total) {
        uint64_t end = begin + a.chunk < a.total ? begin + a.chunk : a.total;
        remaining = (uint32_t)(end - begin);
        root = (uint32_t)(begin / a.rollouts);
        j = (uint32_t)(begin % a.rollouts);
    }

    PlayState st = {};
    bool active = false;
    bool a_is_p0 = true;                   // role A (the root's mover) is player 0
    uint32_t ply = 0, block = 0;
    uint32_t cur_root = root, cur_j = 0;   // the playout in flight
    uint32_t acc_root = root, acc_n = 0, acc_p0 = 0, acc_dr = 0;

    for (;;) {
        // ---- 4-ply boundary: threads without a game in flight start their next one
        if (!active && remaining) {
            if (root != acc_root) {
                flush_tallies(a, acc_root, acc_n, acc_p0, acc_dr);
                acc_root = root; acc_n = acc_p0 = acc_dr = 0;
            }
            const uint4 *rp = reinterpret_cast<const uint4 *>(a.roots + root);
            uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1), r2 = __ldg(rp + 2);
            st.free2 = ~r0.x; st.ra_lo = r0.y; st.ra_hi = r0.z; st.rb_lo = r0.w;
            st.rb_hi = r1.x; st.ta_lo = r1.y; st.ta_hi = r1.z; st.tb_lo = r1.w;
            st.tb_hi = r2.x; st.lp_lo = r2.y; st.lp_hi = r2.z;
            uint32_t flags = r2.w;
            a_is_p0 = (flags & 1u) == 0;
            cur_root = root; cur_j = j;
            ply = 0; block = 0;
            --remaining;
            ++acc_n;                       // every started game finishes before its tallies are flushed
            if (++j == a.rollouts) { j = 0; ++root; }
            uint32_t immediate = (flags >> QK_ROOT_IMMEDIATE_SHIFT) & 3u;
            if ((CUTOFF && a.max_plies == 0) || immediate) {
                // nothing to play (mcts.py:412 cut-off first, then the root's own result)
                int res = (CUTOFF && a.max_plies == 0) ? 0 : (immediate == 1 ? 1 : -1);
                acc_p0 += res > 0; acc_dr += res == 0;
                if (PER_PLAYOUT) {
                    size_t o = (size_t)cur_root * a.rollouts + cur_j;
                    if (a.per_playout) a.per_playout[o] = (int8_t)res;
                    if (a.per_plies) a.per_plies[o] = 0;
                }
            } else {
                active = true;
            }
        }
        if (!__any_sync(kFull, active)) {
            if (!__any_sync(kFull, remaining != 0)) break;
            continue;
        }

        // ---- one Philox block = four plies, roles alternate A B A B
        const uint64_t rollout = a.first_rollout + cur_j;
        const Philox4 rnd = stream_block(a.seed, rollout, a.root_index_base + cur_root, block);
        ++block;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            bool blocked, win;
#ifdef QK_PLAIN_SELECT
            if (q & 1) ply_step<1>(st, rnd.v[q], s_lut, s_sel8, blocked, win);
            else       ply_step<0>(st, rnd.v[q], s_lut, s_sel8, blocked, win);
#else
            if (q & 1) ply_step_fma<1>(st, rnd.v[q], s_lut, s_sel8, blocked, win);
            else       ply_step_fma<0>(st, rnd.v[q], s_lut, s_sel8, blocked, win);
#endif
            // outcome bookkeeping without branches: a blocked mover loses (beam_search.py:641-642),
            // otherwise the move is made and may hit the cut-off (mcts.py:412,436-437) or complete a
            // line (beam_search.py:633-635)
            const bool mover_is_p0 = (q & 1) ? !a_is_p0 : a_is_p0;
            bool cut = false;
            if (CUTOFF || PER_PLAYOUT) ply += blocked ? 0u : 1u;
            if (CUTOFF) cut = !blocked && ply >= (uint32_t)a.max_plies;
            const bool fin = active && (blocked || win || cut);
            const bool p0_wins = (mover_is_p0 != blocked) && !cut;
            acc_p0 += (fin && p0_wins) ? 1u : 0u;
            if (CUTOFF) acc_dr += (fin && cut) ? 1u : 0u;
            if (PER_PLAYOUT && fin) {
                size_t o = (size_t)cur_root * a.rollouts + cur_j;
                if (a.per_playout) a.per_playout[o] = (int8_t)(cut ? 0 : (p0_wins ? 1 : -1));
                if (a.per_plies) a.per_plies[o] = (uint8_t)ply;
            }
            active = active && !fin;
        }
    }
    flush_tallies(a, acc_root, acc_n, acc_p0, acc_dr);
}

// BASELINE config 1 -- every playout from ONE root, to the end, tallies only -- is the headline workload, and there
// the refill path above (root / rollout bookkeeping, root-change flush, three global loads with address arithmetic)
// is 10 of the ~95 instructions a ply costs.  Same loop with the root record in shared memory and a plain counter.
// a shared-memory load the compiler may not hoist out of the loop (keeping the 11-word root record in registers
// would cost a CTA of occupancy)
__device__ __forceinline__ uint4 lds128_fresh(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return v;
}

template <bool FRESH>
__global__ void QK_PLAYOUT_BOUNDS k_playout_single(const PlayoutArgs a)
{
    __shared__ __align__(16) SquareLut s_lut[16];
    __shared__ __align__(16) uint8_t s_sel8[2048];
    __shared__ uint4 s_root[3];
    if (threadIdx.x < 3) s_root[threadIdx.x] = __ldg(reinterpret_cast<const uint4 *>(a.roots) + threadIdx.x);
    fill_tables(s_lut, s_sel8);

    const uint64_t tid = blockIdx.x * (uint64_t)kBlock + threadIdx.x;
    const uint64_t begin = tid * a.chunk;
    uint32_t remaining = 0, next_j = 0;    // playouts this thread still has to START, and the next one's index
    if (begin < a.total) {
        const uint64_t end = begin + a.chunk < a.total ? begin + a.chunk : a.total;
        remaining = (uint32_t)(end - begin);
        next_j = (uint32_t)begin;
    }
    const uint32_t flags = s_root[2].w;
    const bool a_is_p0 = (flags & 1u) == 0;
    const uint32_t immediate = (flags >> QK_ROOT_IMMEDIATE_SHIFT) & 3u;
    uint32_t acc_n = 0, acc_p0 = 0;
    if (immediate) {                       // the root is decided: every playout is that result (mcts.py:412 order)
        acc_n = remaining; acc_p0 = immediate == 1 ? remaining : 0u;
        remaining = 0;
    }

    PlayState st = {};
    bool active = false;
    uint32_t block = 0, cur_j = 0;
    for (;;) {
        // ---- 4-ply boundary: threads without a game in flight start their next one
        if (!active && remaining) {
            const uint4 r0 = FRESH ? lds128_fresh(s_root) : s_root[0], r1 = FRESH ? lds128_fresh(s_root + 1) : s_root[1],
                        r2 = FRESH ? lds128_fresh(s_root + 2) : s_root[2];
            st.free2 = ~r0.x; st.ra_lo = r0.y; st.ra_hi = r0.z; st.rb_lo = r0.w;
            st.rb_hi = r1.x; st.ta_lo = r1.y; st.ta_hi = r1.z; st.tb_lo = r1.w;
            st.tb_hi = r2.x; st.lp_lo = r2.y; st.lp_hi = r2.z;
            cur_j = next_j++;
            block = 0;
            --remaining;
            ++acc_n;
            active = true;
        }
        if (!__any_sync(kFull, active)) break;      // nobody plays and (checked above) nobody could start: done

        // ---- one Philox block = four plies, roles alternate A B A B
        const uint64_t rollout = a.first_rollout + cur_j;
        const Philox4 rnd = stream_block(a.seed, rollout, a.root_index_base, block);
        ++block;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            bool blocked, win;
            if (q & 1) ply_step_fma<1>(st, rnd.v[q], s_lut, s_sel8, blocked, win);
            else       ply_step_fma<0>(st, rnd.v[q], s_lut, s_sel8, blocked, win);
            // a blocked mover loses (beam_search.py:641-642), a completed line wins (:633-635)
            const bool mover_is_p0 = (q & 1) ? !a_is_p0 : a_is_p0;
            const bool fin = active && (blocked || win);
            acc_p0 += (fin && (mover_is_p0 != blocked)) ? 1u : 0u;
            active = active && !fin;
        }
    }
    flush_tallies(a, 0u, acc_n, acc_p0, 0u);
}

// Config 1 again, third version: the ply is ply_pred (predicated PTX, qk_playout_core.cuh) and the loop around it has
// no branch but the exit test -- the refill is three predicated LDS.128 of the root record, the game in flight is
// the all-ones / zero word `act`, and the tallies are one counter (games won by the root's mover; the number of
// games a thread plays is known in advance).
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// The queue's atomic in PTX: written as atomicAdd() the compiler aggregates it over the lanes of the warp that take a
// range in the same iteration and then needs the result at once (a shuffle from the leader) -- the whole warp would
// wait for the atomic's round trip, which taking a range one range ahead is meant to hide.
__device__ __forceinline__ uint32_t queue_take(uint32_t *queue, uint32_t n)
{
    uint32_t old;
    asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(queue), "r"(n) : "memory");
    return old;
}
// A value the compiler keeps in a register instead of recomputing it where it is used (shared-memory addresses cost
// an S2R + LEA each time otherwise -- five instructions per loop iteration of the playout kernels): it comes back
// from a volatile shared-memory load.
__device__ __forceinline__ uint32_t pinned_value(volatile uint32_t *slot, uint32_t v)
{
    *slot = v;
    return *slot;
}

#ifdef QK_PLAYOUT_TIMELINE      // experiment build (scripts/playout_timeline.py): when does each warp start and finish?
__device__ unsigned long long g_timeline[4 * 16384];
__device__ __forceinline__ unsigned long long globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif

template <int MIN_CTAS>
__global__ void __launch_bounds__(kBlock, MIN_CTAS) k_playout_pred(const PlayoutArgs a)
{
    __shared__ __align__(128) uint32_t s_lut[QK_MOVE_LUT_BYTES / 4];
    __shared__ uint8_t s_sel8r[4096];       // the byte table is the 2 KB-aligned half of it: the ply ORs (address / 8) into the index
    __shared__ __align__(16) uint4 s_root[3];
    if (threadIdx.x < 3) {
        uint4 v = __ldg(reinterpret_cast<const uint4 *>(a.roots) + threadIdx.x);
        if (threadIdx.x == 0) v.x = ~v.x;          // occupied -> free squares
        s_root[threadIdx.x] = v;
    }
    __shared__ uint32_t s_pin[2 * kBlock];
    const uint32_t lut_addr = smem_addr(s_lut), root_addr = pinned_value(&s_pin[threadIdx.x], smem_addr(s_root));
    const uint32_t sel_off = (0u - smem_addr(s_sel8r)) & 2047u;
    for (uint32_t i = threadIdx.x; i < QK_MOVE_LUT_BYTES / 4; i += kBlock) s_lut[i] = move_lut_image_word(i);
    for (uint32_t i = threadIdx.x; i < 2048; i += kBlock) s_sel8r[sel_off + i] = (uint8_t)sel8_row(i);
    __syncthreads();
#ifdef QK_PLAYOUT_TIMELINE
    const unsigned long long t_begin = globaltimer();
#endif
    PredConsts pc;
    pc.sel_addr8 = pinned_value(&s_pin[kBlock + threadIdx.x], (smem_addr(s_sel8r) + sel_off) >> 3);
    pc.col_lo = (threadIdx.x & 7u) * 16u + lut_addr; pc.col_hi = pc.col_lo + 4096u;
    pc.one = a.one; pc.zero = a.zero;

    // Work is not dealt out in advance: the warp scheduler favours some warps of an SM over others (with equal shares
    // the slowest quarter of the warps finished at 1.8x the time of the fastest, and the SM ran the last fifth of the
    // kernel with a quarter of its warps), so every lane takes `grab` playouts at a time from one queue, always
    // holding its current range [jn, jend) and the start of its next one (taken ahead of time, so that the atomic's
    // latency is never waited for).
    const uint32_t total = (uint32_t)a.total, grab = a.grab;
    uint32_t jn, jend, jnext;              // next playout this lane starts, one past its range, its next range
                                           // (a lane that found the queue empty keeps jn > jend: nothing to start, nothing to take)
    {
        uint32_t base = 0;
        if ((threadIdx.x & 31u) == 0) base = atomicAdd(a.queue, 64u * grab);
        base = __shfl_sync(kFull, base, 0);
        jn = base + (threadIdx.x & 31u) * grab;
        jnext = jn + 32u * grab;
        jend = jn + grab < total ? jn + grab : total;
        if (jn >= total) { jn = 1; jend = 0; }
    }
    const uint32_t flags = s_root[2].w;
    const bool a_is_p0 = (flags & 1u) == 0;
    const uint32_t immediate = (flags >> QK_ROOT_IMMEDIATE_SHIFT) & 3u;
    uint32_t acc = 0, games = jn < jend ? jend - jn : 0u;   // games won by role A; games played (each one started is finished)
    if (immediate) {                       // the root is decided: every playout is that result (mcts.py:412 order)
        games = (blockIdx.x == 0 && threadIdx.x == 0) ? total : 0u;
        acc = ((immediate == 1) == a_is_p0) ? games : 0u;
        jn = 1; jend = 0;
    }

    PlayState st = {};
    uint32_t act = 0, block = 0;
    const uint64_t rollout_base = a.first_rollout - 1;     // the game in flight is playout jn - 1
    for (;;) {
        // ---- a lane that has used up its range moves on to the one it holds in reserve and reserves another
        if ((act | (jn ^ jend)) == 0u) {
            jn = jnext;
            if (jn >= total) { jn = 1; jend = 0; }
            else {
                jend = jn + grab < total ? jn + grab : total;
                jnext = queue_take(a.queue, grab);
                games += jend - jn;
            }
        }
        // ---- 4-ply boundary: threads without a game in flight start their next one
        asm volatile(
            "{\n"
            " .reg .pred pi, pr;\n"
            " .reg .b32 fl;\n"
            " setp.eq.u32 pi, %13, 0;\n"
            " setp.lt.and.u32 pr, %11, %14, pi;\n"
            " @pr ld.shared.v4.u32 {%0, %1, %2, %3}, [%15];\n"
            " @pr ld.shared.v4.u32 {%4, %5, %6, %7}, [%15+16];\n"
            " @pr ld.shared.v4.u32 {%8, %9, %10, fl}, [%15+32];\n"
            " @pr add.u32 %11, %11, 1;\n"
            " @pr mov.b32 %12, 0;\n"
            " @pr mov.b32 %13, 0xFFFFFFFF;\n"
            "}\n"
            : "+r"(st.free2), "+r"(st.ra_lo), "+r"(st.ra_hi), "+r"(st.rb_lo), "+r"(st.rb_hi), "+r"(st.ta_lo), "+r"(st.ta_hi),
              "+r"(st.tb_lo), "+r"(st.tb_hi), "+r"(st.lp_lo), "+r"(st.lp_hi), "+r"(jn), "+r"(block), "+r"(act)
            : "r"(jend), "r"(root_addr));
        if (!__any_sync(kFull, act != 0u)) break;   // nobody plays and (checked above) nobody could start: done

        // ---- one Philox block = four plies, roles alternate A B A B
        const Philox4 rnd = stream_block(a.seed, rollout_base + jn, a.root_index_base, block);
        ++block;
        ply_pred<0>(st, rnd.v[0], pc, act, acc);
        ply_pred<1>(st, rnd.v[1], pc, act, acc);
        ply_pred<0>(st, rnd.v[2], pc, act, acc);
        ply_pred<1>(st, rnd.v[3], pc, act, acc);
    }
    flush_tallies(a, 0u, games, a_is_p0 ? acc : games - acc, 0u);
#ifdef QK_PLAYOUT_TIMELINE
    if ((threadIdx.x & 31u) == 0) {
        const uint32_t w = blockIdx.x * (kBlock / 32) + threadIdx.x / 32;
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        if (w < 16384) { g_timeline[4 * w] = t_begin; g_timeline[4 * w + 1] = globaltimer(); g_timeline[4 * w + 2] = smid; g_timeline[4 * w + 3] = games; }
    }
#endif
}
#ifdef QK_PLAYOUT_TIMELINE
}  // namespace qk
extern "C" __attribute__((visibility("default"))) int qk_debug_timeline(unsigned long long *out, size_t n)
{
    return (int)cudaMemcpyFromSymbol(out, qk::g_timeline, n * sizeof(unsigned long long));
}
namespace qk {
#endif

// Configs with many roots (BASELINE config 4, the batched engines' leaves), tallies only, played to the end: the same
// ply and the same queue, whose items are (root, slice of `grab` rollouts) -- a lane's range never straddles two roots,
// so the root's record address, its index in the random stream and the tallies' destination change only in the
// rarely taken range switch; a root that is already decided is accounted for there without being played.
// GENERAL adds what only some callers need, at a few instructions per ply: the MCTS cut-off after max_plies moves (a draw,
// tested before the win: mcts.py:412,436-437) and per-playout outcomes / lengths.
template <int MIN_CTAS, bool GENERAL>
__global__ void __launch_bounds__(kBlock, MIN_CTAS) k_playout_pred_roots(const PlayoutArgs a)
{
    __shared__ __align__(128) uint32_t s_lut[QK_MOVE_LUT_BYTES / 4];
    __shared__ uint8_t s_sel8r[4096];       // the byte table is the 2 KB-aligned half of it: the ply ORs (address / 8) into the index
    __shared__ __align__(16) uint4 s_rec[3][kBlock];   // every lane's copy of its current root's record (free squares first)
    const uint32_t lut_addr = smem_addr(s_lut), sel_off = (0u - smem_addr(s_sel8r)) & 2047u;
    for (uint32_t i = threadIdx.x; i < QK_MOVE_LUT_BYTES / 4; i += kBlock) s_lut[i] = move_lut_image_word(i);
    for (uint32_t i = threadIdx.x; i < 2048; i += kBlock) s_sel8r[sel_off + i] = (uint8_t)sel8_row(i);
    __syncthreads();
    PredConsts pc;
    __shared__ uint32_t s_pin[2 * kBlock];
    pc.sel_addr8 = pinned_value(&s_pin[threadIdx.x], (smem_addr(s_sel8r) + sel_off) >> 3);
    pc.col_lo = (threadIdx.x & 7u) * 16u + lut_addr; pc.col_hi = pc.col_lo + 4096u;
    pc.one = a.one; pc.zero = a.zero;

    const uint32_t grab = a.grab, slices = (a.rollouts + grab - 1) / grab;     // slices per root
    const uint32_t items = (uint32_t)(a.total);                                // here: number of (root, slice) items
    constexpr uint32_t kExhausted = 0xFFFFFFFFu;
    uint32_t qnext;                         // the item this lane plays next
    {
        uint32_t base = 0;
        if ((threadIdx.x & 31u) == 0) base = atomicAdd(a.queue, 32u);
        qnext = __shfl_sync(kFull, base, 0) + (threadIdx.x & 31u);
    }
    uint32_t j = 0, jend = 0;               // next rollout of the current root this lane starts, one past its slice
    uint32_t root = 0, acc = 0, games = 0;  // current root, games of it won by its mover, games of it played
    uint32_t accd = 0, moves = 0;           // GENERAL: games of it cut off (draws); moves made in the game in flight
    const bool cutoff = GENERAL && a.max_plies >= 0;
    bool a_is_p0 = true;
    const uint32_t rec_addr = pinned_value(&s_pin[kBlock + threadIdx.x], smem_addr(&s_rec[0][threadIdx.x]));

    PlayState st = {};
    uint32_t act = 0, block = 0;
    const uint64_t rollout_base = a.first_rollout - 1;     // the game in flight is rollout j - 1
    for (;;) {
        // ---- the lanes of a warp move on to their next items TOGETHER, when none of them has a game left to play or to
        //      start: the switch is a long divergent path, and taken lane by lane -- 30 % of all iterations, with one or
        //      two lanes in it -- it was 18 % of the executed instructions of config 4 (profiles/r02l_cfg4_playout_ncu);
        //      a lane now waits for the slowest lane's slice instead, 0.41 / sqrt(slice length) of its time (5 % at 64)
        //      (a lane whose item needs no playing -- its root is decided -- takes the next one at once, so that decided
        //      roots cost their lane a few dozen instructions each, not a round of waiting)
        if (!__any_sync(kFull, act != 0u || j != jend)) while (j == jend && qnext != kExhausted) {
            const uint32_t q = qnext;
            if (q >= items) qnext = kExhausted;
            else {
                qnext = queue_take(a.queue, 1u);
                const uint32_t r = q / slices, j0 = (q - r * slices) * grab;
                if (r != root) {
                    if (slices == 1u) {     // this lane played every rollout of the root: plain stores, nobody else adds
                        if (games) {
                            const uint32_t p0 = a_is_p0 ? acc : games - acc - accd;
                            if (a.wins_p0) a.wins_p0[root] = p0;
                            if (a.wins_p1) a.wins_p1[root] = games - p0 - accd;
                            if (a.draws) a.draws[root] = accd;
                        }
                    } else flush_tallies(a, root, games, a_is_p0 ? acc : games - acc - accd, accd);
                    acc = games = accd = 0;
                    root = r;
                }
                // the record moves into the lane's shared-memory slot: the refills below then cost what they cost
                // in the single-root kernel instead of three global loads each
                const uint4 *rp = reinterpret_cast<const uint4 *>(a.roots + r);
                uint4 r0 = __ldg(rp);
                const uint4 r1 = __ldg(rp + 1), r2 = __ldg(rp + 2);
                r0.x = ~r0.x;
                s_rec[0][threadIdx.x] = r0; s_rec[1][threadIdx.x] = r1; s_rec[2][threadIdx.x] = r2;
                const uint32_t flags = r2.w;
                a_is_p0 = (flags & 1u) == 0;
                const uint32_t len = a.rollouts - j0 < grab ? a.rollouts - j0 : grab;
                const uint32_t immediate = (flags >> QK_ROOT_IMMEDIATE_SHIFT) & 3u;
                games += len;
                j = j0; jend = j0 + len;
                if (cutoff && a.max_plies == 0) {           // nothing may be played: draws (the cut-off comes first)
                    accd += len;
                    if (GENERAL) for (uint32_t i = j0; i < j0 + len; ++i) {
                        if (a.per_playout) a.per_playout[(size_t)r * a.rollouts + i] = 0;
                        if (a.per_plies) a.per_plies[(size_t)r * a.rollouts + i] = 0;
                    }
                    j = jend;
                } else if (immediate) {     // decided at the root: that result for the whole slice, nothing to play
                    acc += ((immediate == 1) == a_is_p0) ? len : 0u;
                    if (GENERAL) for (uint32_t i = j0; i < j0 + len; ++i) {
                        if (a.per_playout) a.per_playout[(size_t)r * a.rollouts + i] = immediate == 1 ? 1 : -1;
                        if (a.per_plies) a.per_plies[(size_t)r * a.rollouts + i] = 0;
                    }
                    j = jend;
                }
            }
        }
        // ---- 4-ply boundary: threads without a game in flight start their next one
        asm volatile(
            "{\n"
            " .reg .pred pi, pr;\n"
            " .reg .b32 fl;\n"
            " setp.eq.u32 pi, %13, 0;\n"
            " setp.lt.and.u32 pr, %11, %14, pi;\n"
            " @pr ld.shared.v4.u32 {%0, %1, %2, %3}, [%15];\n"
            " @pr ld.shared.v4.u32 {%4, %5, %6, %7}, [%15+4096];\n"
            " @pr ld.shared.v4.u32 {%8, %9, %10, fl}, [%15+8192];\n"
            " @pr add.u32 %11, %11, 1;\n"
            " @pr mov.b32 %12, 0;\n"
            " @pr mov.b32 %13, 0xFFFFFFFF;\n"
            "}\n"
            : "+r"(st.free2), "+r"(st.ra_lo), "+r"(st.ra_hi), "+r"(st.rb_lo), "+r"(st.rb_hi), "+r"(st.ta_lo), "+r"(st.ta_hi),
              "+r"(st.tb_lo), "+r"(st.tb_hi), "+r"(st.lp_lo), "+r"(st.lp_hi), "+r"(j), "+r"(block), "+r"(act)
            : "r"(jend), "r"(rec_addr)
            : "memory");
        static_assert(kBlock * 16 == 4096, "the record planes' stride is spelled out in the PTX above");
        if (!__any_sync(kFull, act != 0u || qnext != kExhausted)) break;   // nobody plays, nobody can take another item

        // ---- one Philox block = four plies, roles alternate A B A B
        const Philox4 rnd = stream_block(a.seed, rollout_base + j, a.root_index_base + root, block);
        ++block;
        if (!GENERAL) {
            ply_pred<0>(st, rnd.v[0], pc, act, acc);
            ply_pred<1>(st, rnd.v[1], pc, act, acc);
            ply_pred<0>(st, rnd.v[2], pc, act, acc);
            ply_pred<1>(st, rnd.v[3], pc, act, acc);
        } else {
            if (block == 1u) moves = 0;                   // (a game starts with its first block)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (q & 1) ply_pred_move<1>(st, rnd.v[q], pc); else ply_pred_move<0>(st, rnd.v[q], pc);
                const int32_t t = ply_pred_outcome(st, act);
                if (act) {
                    moves += t >= 0;                       // a blocked mover makes no move
                    const bool cut = cutoff && t >= 0 && moves >= (uint32_t)a.max_plies;
                    if (cut || t != 0) {
                        const bool a_wins = !cut && ((q & 1) ? t < 0 : t > 0);
                        acc += a_wins; accd += cut;
                        const size_t o = (size_t)root * a.rollouts + (j - 1);
                        if (a.per_playout) a.per_playout[o] = (int8_t)(cut ? 0 : (a_wins == a_is_p0 ? 1 : -1));
                        if (a.per_plies) a.per_plies[o] = (uint8_t)moves;
                        act = 0;
                    }
                }
            }
        }
    }
    if (slices == 1u) {
        if (games) {
            const uint32_t p0 = a_is_p0 ? acc : games - acc - accd;
            if (a.wins_p0) a.wins_p0[root] = p0;
            if (a.wins_p1) a.wins_p1[root] = games - p0 - accd;
            if (a.draws) a.draws[root] = accd;
        }
    } else flush_tallies(a, root, games, a_is_p0 ? acc : games - acc - accd, accd);
}

size_t playout_scratch_bytes(size_t n_roots) { return n_roots * sizeof(RootRec) + 16; }   // + the work-queue counter

int launch_playouts(const uint4 *roots, size_t n_roots, uint32_t rollouts, uint64_t seed, uint64_t first_rollout,
                    uint32_t root_index_base, int max_plies, uint32_t *wins_p0, uint32_t *wins_p1, uint32_t *draws,
                    int8_t *per_playout, uint8_t *per_plies, void *rootrec_scratch, Ctx *c, cudaStream_t s)
{
    RootRec *recs = (RootRec *)rootrec_scratch;
    k_prepare_roots<<<(unsigned)((n_roots + kBlock - 1) / kBlock), kBlock, 0, s>>>(roots, n_roots, recs);
    QK_CUDA(cudaGetLastError());

    PlayoutArgs a;
    a.roots = recs;
    a.total = (uint64_t)n_roots * rollouts;
    a.seed = seed; a.first_rollout = first_rollout; a.rollouts = rollouts; a.root_index_base = root_index_base;
    a.max_plies = max_plies;
    a.wins_p0 = wins_p0; a.wins_p1 = wins_p1; a.draws = draws; a.per_playout = per_playout; a.per_plies = per_plies;
    a.one = 1u; a.zero = 0u;
    a.queue = reinterpret_cast<uint32_t *>(recs + n_roots); a.grab = 1;

    const bool cutoff = max_plies >= 0, pp = per_playout || per_plies;
    void (*kernel)(const PlayoutArgs) = cutoff ? (pp ? k_playout<true, true> : k_playout<true, false>)
                                               : (pp ? k_playout<false, true> : k_playout<false, false>);
    // Tallies only, played to the end: the predicated kernels, which take their work from a queue (a 32-bit counter:
    // every lane may over-shoot it once).  "playout_variant": 0 auto | 1, 6 the older single-root kernels (root record
    // hoisted / re-read) | 2 the general kernel | 3, 4 the predicated kernels at 5 / 6 resident CTAs per SM instead of 4
    const long long v = option(OPT_PLAYOUT_VARIANT);
    const bool queue_fits = a.total <= 0xFFFFFFFFull - (1ull << 27);
    bool pred = false, pred_roots = false;
    if (n_roots == 1 && !cutoff && !pp && queue_fits && v != 2) {
        pred = true;
        if (v == 0 || v == 5) kernel = k_playout_pred<4>;
        else if (v == 3) kernel = k_playout_pred<5>;
        else if (v == 4) kernel = k_playout_pred<6>;
        else {
            pred = false;
            if (v == 1) kernel = k_playout_single<false>;
            else if (v == 6) kernel = k_playout_single<true>;
        }
    } else if (queue_fits && v != 2 && (rollouts < 4 || rollouts >= 32 || v != 0)) {
        // (between 4 and 31 rollouts per root the range switch -- a divergent path with three global loads -- comes too
        // often: 2^20 roots x 16: 2.7e10 playouts/s against 3.3e10 for round 1's kernel, which refills from global memory
        // at every game anyway; x 1: 1.0e10 against 4.6e9, x 1024: 5.2e10 against 3.3e10)
        pred_roots = true;
        if (!cutoff && !pp) kernel = v == 3 ? k_playout_pred_roots<5, false> : v == 4 ? k_playout_pred_roots<6, false> : k_playout_pred_roots<4, false>;
        else kernel = k_playout_pred_roots<4, true>;
    }
    // persistent grid: exactly as many CTAs as are resident at once (no second wave), fewer only
    // when there are not enough playouts to give each thread one
    int per_sm = 0;
    QK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, (int)kBlock, 0));
    if (per_sm < 1) per_sm = 1;
    uint64_t max_threads = (uint64_t)c->sms * per_sm * kBlock;
    uint64_t threads = a.total < max_threads ? a.total : max_threads;
    unsigned grid = (unsigned)((threads + kBlock - 1) / kBlock);
    a.chunk = (a.total + (uint64_t)grid * kBlock - 1) / ((uint64_t)grid * kBlock);
    if (a.chunk >= (1ull << 32)) { set_error("qk_playouts: more than 2^32 playouts per thread"); return QK_E_BAD_ARG; }
    if (pred || pred_roots) {
        // lanes take their playouts from a queue, `grab` at a time: small enough that the lanes run dry together at the
        // end (about 1/16 of a lane's share), large enough that the counter is not a hot spot
        const long long g = option(OPT_PLAYOUT_GRAB);
        uint64_t share = a.total / ((uint64_t)grid * kBlock * 16);
        a.grab = g > 0 ? (uint32_t)(g < 256 ? g : 256) : (uint32_t)(share < 1 ? 1 : share > 64 ? 64 : share);
        if (pred_roots) {                               // the queue's items are (root, slice of `grab` rollouts)
            if (a.grab > rollouts) a.grab = rollouts;
            a.total = (uint64_t)n_roots * ((rollouts + a.grab - 1) / a.grab);
        }
        QK_CUDA(cudaMemsetAsync(a.queue, 0, 4, s));
    }
    kernel<<<grid, kBlock, 0, s>>>(a);
    QK_CUDA(cudaGetLastError());
    return QK_OK;
}

// ------------------------------------------------------------------ root generator
// benchmarks/dataset.py:38-51 semantics (rejection of finished lines); see
// include/quantik_b200.h qk_random_roots.  One thread per root; not a hot path.
__global__ void __launch_bounds__(kBlock) k_random_roots(uint64_t seed, uint64_t first, size_t n, int min_plies, int span,
                                                          uint4 *__restrict__ out)
{
    __shared__ __align__(16) SquareLut s_lut[16];
    __shared__ __align__(16) uint8_t s_sel8[2048];
    fill_tables(s_lut, s_sel8);
    size_t ii = blockIdx.x * (size_t)kBlock + threadIdx.x;
    if (ii >= n) return;
    const uint64_t i = first + ii;
    const int target = min_plies + (int)(i % (uint64_t)span);
    const uint64_t rseed = seed ^ 0x524F4F5453ull;
    State4 empty = { 0, 0, 0, 0 };
    const RootRec root = make_root(empty);
    for (uint32_t attempt = 0;; ++attempt) {
        PlayState st = load_play_state(root);
        State4 cur = empty;
        bool ok = true;
        Philox4 rnd = { { 0, 0, 0, 0 } };
        for (int ply = 0; ply < target && ok; ++ply) {
            if ((ply & 3) == 0)
                rnd = philox4x32_10((uint32_t)i, attempt, (uint32_t)(ply >> 2), 0u, (uint32_t)rseed, (uint32_t)(rseed >> 32));
            bool blocked, win;
            int action;
            if (ply & 1) ply_step<1>(st, rnd.v[ply & 3], s_lut, s_sel8, blocked, win, &action);
            else         ply_step<0>(st, rnd.v[ply & 3], s_lut, s_sel8, blocked, win, &action);
            if (blocked || win) { ok = false; break; }
            cur = apply_action(cur, ply & 1, action);
        }
        if (ok) {   // the side to move at the root must still have a move
            uint32_t r_lo = (target & 1) ? st.rb_lo : st.ra_lo, r_hi = (target & 1) ? st.rb_hi : st.ra_hi;
            ok = ((st.free2 & ~r_lo) | (st.free2 & ~r_hi)) != 0;
        }
        if (ok) { out[ii] = make_uint4(cur.x, cur.y, cur.z, cur.w); return; }
    }
}

int launch_random_roots(uint64_t seed, uint64_t first, size_t n, int min_plies, int span, uint4 *out, Ctx *c, cudaStream_t s)
{
    (void)c;
    k_random_roots<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, s>>>(seed, first, n, min_plies, span, out);
    QK_CUDA(cudaGetLastError());
    return QK_OK;
}

}  // namespace qk
