// qk_perft.cu -- exhaustive enumeration below a batch of roots (BASELINE configs 2
// and 5) with the counting rules of game_stats.SymmetryTable (game_stats.py:384-485).
//
// Plan of one qk_perft call (host driver run_perft at the bottom):
//   k_perft_filter_roots  roots -> frontier of valid, not-yet-won positions
//   k_perft_expand        level-synchronous widening, one position per thread, until the
//                         frontier has enough items to fill the machine
//   k_perft_dfs           one frontier item per WARP: depth-first walk with a
//                         warp-shared frame stack in shared memory.  Interior nodes are
//                         expanded cooperatively (lane l materialises children l and
//                         l+32 of the 64-bit action mask); positions one ply above the
//                         target depth ("leaf parents") are queued in shared memory and
//                         consumed 32 at a time, one per lane, so the dominant work runs
//                         with full warps.
//   k_perft_leaf          frontier already one ply above the target: one per thread.
// The last ply is never materialised: move counts and winning-move counts of a leaf
// parent come from popcounts (qk_perft_core.cuh).
#include "qk_common.cuh"
#include "qk_perft_core.cuh"
#include "qk_playout_core.cuh"     // SquareLut: scope / square masks of one square

#include <stdlib.h>

namespace qk {

static constexpr unsigned kBlock = 256;
static constexpr unsigned kWarps = kBlock / 32;
static constexpr unsigned kFull = 0xFFFFFFFFu;
static constexpr int kRows = QK_MAX_DEPTH + 1;        // counters[5][17]: nodes, wins_p0, wins_p1, blocked_p0, blocked_p1
typedef unsigned long long u64;

__device__ __forceinline__ u64 warp_sum(u64 v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ State4 to_state(const uint4 &v) { State4 s = { v.x, v.y, v.z, v.w }; return s; }
__device__ __forceinline__ uint4 to_u4(const State4 &s) { return make_uint4(s.x, s.y, s.z, s.w); }

// ------------------------------------------------------------------ roots -> frontier
__global__ void __launch_bounds__(kBlock) k_perft_filter_roots(const uint4 *__restrict__ roots, const u64 *__restrict__ mult,
                                                                size_t n, int depth, uint4 *__restrict__ out_states,
                                                                u64 *__restrict__ out_mult, u64 *__restrict__ out_count,
                                                                u64 *__restrict__ counters, unsigned *__restrict__ flags)
{
    size_t i = blockIdx.x * (size_t)kBlock + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    u64 m = 0, blocked = 0;
    bool keep = false;
    unsigned seen = 0, pieces = 0;
    State4 s = { 0, 0, 0, 0 };
    if (i < n) {
        s = to_state(__ldg(roots + i));
        m = mult ? mult[i] : 1ull;
        int player;
        if (win_status(s) == 0) {                       // a finished root has nothing below it
            if (validate(s, &player) != 0) blocked = m; // generate_legal_moves -> (0, {})  (move.py:223-225)
            else {
                keep = true; seen = (1u << player) | ((m >> 32) ? 4u : 0u);
                pieces = (unsigned)(popc(s.x) + popc(s.y) + popc(s.z) + popc(s.w));
            }
        }
    }
    // what the frontier looks like: k_perft_tile has a faster form for one side to move + 32-bit multiplicities,
    // and the choice tile / lane walk goes by the FULLEST root (flags[1] = most pieces on a kept root)
    seen = __reduce_or_sync(kFull, seen);
    pieces = __reduce_max_sync(kFull, pieces);
    if (lane == 0 && seen) { atomicOr(flags, seen); atomicMax(flags + 1, pieces); }
    u64 tot = warp_sum(m), blk = warp_sum(blocked);
    unsigned kmask = __ballot_sync(kFull, keep);
    u64 base = 0;
    if (lane == 0) {
        if (tot) atomicAdd(counters + 0 * kRows + 0, tot);
        if (blk && depth > 0) atomicAdd(counters + 3 * kRows + 0, blk);
        if (kmask) base = atomicAdd(out_count, (u64)__popc(kmask));
    }
    base = __shfl_sync(kFull, base, 0);
    if (keep) {
        u64 o = base + __popc(kmask & ((1u << lane) - 1u));
        out_states[o] = to_u4(s);
        out_mult[o] = m;
    }
}

// ------------------------------------------------------------------ one level, thread per position
// _process_parent_state (game_stats.py:384-424): total_moves, wins by the mover, children
// COUNT_ONLY: the sizing pass -- only the number of children the level will emit (the next frontier is then
// allocated exactly, not at 64 children per parent)
template <bool COUNT_ONLY>
__global__ void __launch_bounds__(kBlock) k_perft_expand(const uint4 *__restrict__ in_states, const u64 *__restrict__ in_mult,
                                                          size_t n, int d, uint4 *__restrict__ out_states,
                                                          u64 *__restrict__ out_mult, u64 *__restrict__ out_count,
                                                          u64 *__restrict__ counters, const u64 *__restrict__ n_dev = nullptr)
{
    // n_dev: the frontier's size is still on the device (small levels are chained without a host round trip, run_perft);
    // n is then only an upper bound, the one the grid was sized for
    if (n_dev) { const size_t n_real = (size_t)*n_dev; n = n_real < n ? n_real : n; }
    size_t i = blockIdx.x * (size_t)kBlock + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    State4 s = { 0, 0, 0, 0 };
    u64 m = 0;
    uint32_t nl = 0, nh = 0;
    int mover = 0, n_moves = 0, n_wins = 0;
    u64 c_nodes = 0, c_w0 = 0, c_w1 = 0, c_b0 = 0, c_b1 = 0;
    if (i < n) {
        s = to_state(__ldg(in_states + i));
        m = in_mult[i];
        mover = piece_balance(s);
        uint32_t lo, hi, wab, wcd;
        legal_masks(s, mover, lo, hi);
        winning_squares(s, wab, wcd);
        n_moves = popc(lo) + popc(hi);
        n_wins = popc(lo & wab) + popc(hi & wcd);
        nl = lo & ~wab; nh = hi & ~wcd;
        c_nodes = (u64)n_moves * m;
        if (mover == 0) { c_w0 = (u64)n_wins * m; c_b0 = n_moves ? 0 : m; }
        else            { c_w1 = (u64)n_wins * m; c_b1 = n_moves ? 0 : m; }
    }
    uint32_t cnt = (uint32_t)(popc(nl) + popc(nh)), incl = cnt;
    if (COUNT_ONLY) {
        const uint32_t total = (uint32_t)__reduce_add_sync(kFull, cnt);
        if (lane == 0 && total) atomicAdd(out_count, (u64)total);
        return;
    }
    c_nodes = warp_sum(c_nodes); c_w0 = warp_sum(c_w0); c_w1 = warp_sum(c_w1); c_b0 = warp_sum(c_b0); c_b1 = warp_sum(c_b1);
    // exclusive scan of the children counts inside the warp, one reservation per warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(kFull, incl, o); if (lane >= (unsigned)o) incl += t; }
    uint32_t total = __shfl_sync(kFull, incl, 31);
    u64 base = 0;
    if (lane == 0) {
        if (c_nodes) atomicAdd(counters + 0 * kRows + d + 1, c_nodes);
        if (c_w0) atomicAdd(counters + 1 * kRows + d + 1, c_w0);
        if (c_w1) atomicAdd(counters + 2 * kRows + d + 1, c_w1);
        if (c_b0) atomicAdd(counters + 3 * kRows + d, c_b0);
        if (c_b1) atomicAdd(counters + 4 * kRows + d, c_b1);
        if (total) base = atomicAdd(out_count, (u64)total);
    }
    // The children of the warp's 32 parents are dealt out: lane j of a round stores child base + j, whoever its parent
    // is -- one coalesced 512-byte store per round.  (Every lane writing its own children, 640 bytes apart from its
    // neighbours', ran at 0.8 TB/s: 0.20 of perft(6)'s 1.5 ms was this kernel's last level.)
    base = __shfl_sync(kFull, base, 0);
    const uint32_t start = incl - cnt;
    for (uint32_t r = 0; r < total; r += 32) {
        const uint32_t idx = r + lane;
        // owner of child idx: the first lane whose range ends beyond it = the number of lanes with incl <= idx
        int src = 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const uint32_t v = __shfl_sync(kFull, incl, src + o - 1); if (v <= idx) src += o; }
        const uint32_t k = idx - __shfl_sync(kFull, start, src);
        const uint32_t plo = __shfl_sync(kFull, nl, src), phi = __shfl_sync(kFull, nh, src);
        const State4 ps = { __shfl_sync(kFull, s.x, src), __shfl_sync(kFull, s.y, src), __shfl_sync(kFull, s.z, src), __shfl_sync(kFull, s.w, src) };
        const u64 pm = ((u64)__shfl_sync(kFull, (uint32_t)(m >> 32), src) << 32) | __shfl_sync(kFull, (uint32_t)m, src);
        const int pmover = __shfl_sync(kFull, mover, src);
        if (idx < total) {
            out_states[base + idx] = to_u4(apply_action(ps, pmover, select_bit64(plo, phi, popc(plo), (int)k)));
            out_mult[base + idx] = pm;
        }
    }
}

// ------------------------------------------------------------------ frontier one ply above the target
__global__ void __launch_bounds__(kBlock) k_perft_leaf(const uint4 *__restrict__ states, const u64 *__restrict__ mult,
                                                        size_t n, int d, u64 *__restrict__ counters)
{
    const unsigned lane = threadIdx.x & 31u;
    u64 c_nodes = 0, c_w0 = 0, c_w1 = 0, c_b0 = 0, c_b1 = 0;
    for (size_t i = blockIdx.x * (size_t)kBlock + threadIdx.x; i < n; i += (size_t)gridDim.x * kBlock) {
        State4 s = to_state(__ldg(states + i));
        u64 m = mult[i];
        int nm, nw;
        int mover = leaf_parent_counts(s, nm, nw);
        c_nodes += (u64)nm * m;
        if (mover == 0) { c_w0 += (u64)nw * m; c_b0 += nm ? 0 : m; }
        else            { c_w1 += (u64)nw * m; c_b1 += nm ? 0 : m; }
    }
    c_nodes = warp_sum(c_nodes); c_w0 = warp_sum(c_w0); c_w1 = warp_sum(c_w1); c_b0 = warp_sum(c_b0); c_b1 = warp_sum(c_b1);
    if (lane == 0) {
        if (c_nodes) atomicAdd(counters + 0 * kRows + d + 1, c_nodes);
        if (c_w0) atomicAdd(counters + 1 * kRows + d + 1, c_w0);
        if (c_w1) atomicAdd(counters + 2 * kRows + d + 1, c_w1);
        if (c_b0) atomicAdd(counters + 3 * kRows + d, c_b0);
        if (c_b1) atomicAdd(counters + 4 * kRows + d, c_b1);
    }
}

// ------------------------------------------------------------------ warp-cooperative DFS
// move / winning-move counts of a leaf parent from its PARENT's data: `legal` = the squares its mover could have
// played one ply earlier (legal_masks of the parent for the player NOT to move there), `action` = the move that
// was made.  The new piece takes its square away from every shape and its row / column / zone away from its own
// shape (move.py:169-196); nothing else changes for the other player.
__device__ __forceinline__ void leaf_counts_delta(const State4 &child, uint2 legal, int action, const SquareLut *lut,
                                                  int &n_moves, int &n_wins)
{
    const SquareLut e = lut[action & 15];
    const uint32_t lane = (action & 16) ? 0xFFFF0000u : 0x0000FFFFu;
    const uint32_t own = e.scope2 & lane;                                  // the square itself is inside its scope
    const uint32_t lo = legal.x & ~(e.bit2 | ((action & 32) ? 0u : own));
    const uint32_t hi = legal.y & ~(e.bit2 | ((action & 32) ? own : 0u));
    uint32_t wab, wcd;
    winning_squares(child, wab, wcd);
    n_moves = popc(lo) + popc(hi);
    n_wins = popc(lo & wab) + popc(hi & wcd);
}

struct WarpScratch {
    uint4 queue[96];              // leaf parents waiting for a full warp (<= 31 + 64)
    uint2 queue_legal[96];        // ... the legal squares of THEIR mover at their parent (before the parent's move)
    unsigned char queue_action[96];   // ... and the move that made them
    uint4 frame_state[QK_MAX_DEPTH];
    u64 frame_mask[QK_MAX_DEPTH]; // non-winning children of frame_state still to visit
    u64 cnt[5 * kRows];           // this warp's share of the counters, weighted by multiplicity
};

__global__ void __launch_bounds__(kBlock) k_perft_dfs(const uint4 *__restrict__ items, const u64 *__restrict__ mult, u64 n_items,
                                                       int d0, int depth, u64 *__restrict__ next_item, u64 *__restrict__ counters)
{
    __shared__ WarpScratch s_scratch[kWarps];
    __shared__ __align__(16) SquareLut s_lut[16];
    if (threadIdx.x < 16) s_lut[threadIdx.x] = square_lut((int)threadIdx.x);
    __syncthreads();
    WarpScratch &ws = s_scratch[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t lt = (1u << lane) - 1u;
    for (int k = lane; k < 5 * kRows; k += 32) ws.cnt[k] = 0;
    __syncwarp();

    const int rem0 = depth - d0;   // >= 2: plies left below an item
    for (;;) {
        u64 idx = 0;
        if (lane == 0) idx = atomicAdd(next_item, 1ull);
        idx = __shfl_sync(kFull, idx, 0);
        if (idx >= n_items) break;
        const State4 item = to_state(__ldg(items + idx));
        const u64 m = mult[idx];
        const int item_parity = piece_balance(item);                   // player to move at the item
        const int leaf_mover = (item_parity + rem0 - 1) & 1;           // who moves one ply above the target
        u64 leaf_nodes = 0, leaf_wins = 0, leaf_blocked = 0;           // per lane, unweighted

        int sp = 0, qn = 0;
        State4 node = item;
        int node_rem = rem0;
        bool expand = true;
        for (;;) {
            if (expand) {
                // ---- interior node (warp-uniform): its moves, its winning moves
                const int d = depth - node_rem;
                const int mover = (item_parity + rem0 - node_rem) & 1;
                uint32_t lo, hi, wab, wcd;
                legal_masks(node, mover, lo, hi);
                const int n_moves = popc(lo) + popc(hi);
                if (n_moves == 0) {
                    if (lane == 0) ws.cnt[(3 + mover) * kRows + d] += m;           // game_stats.py:426-443
                } else {
                    winning_squares(node, wab, wcd);
                    const uint32_t nl = lo & ~wab, nh = hi & ~wcd;
                    if (lane == 0) {
                        ws.cnt[0 * kRows + d + 1] += (u64)n_moves * m;             // :401
                        ws.cnt[(1 + mover) * kRows + d + 1] += (u64)(popc(lo & wab) + popc(hi & wcd)) * m;   // :411-415
                    }
                    if (node_rem == 2) {
                        // children are leaf parents: lane l writes children l and l+32 into the queue, together
                        // with what the OTHER player could play here -- a child's legal moves are those minus
                        // the new piece's square and scope (leaf_counts_delta), 8 operations instead of 46
                        uint32_t ylo, yhi;
                        legal_masks(node, mover ^ 1, ylo, yhi);
                        const uint2 ylegal = make_uint2(ylo, yhi);
                        const int c_lo = popc(nl);
                        if ((nl >> lane) & 1u) {
                            const int at = qn + popc(nl & lt);
                            ws.queue[at] = to_u4(apply_action(node, mover, (int)lane));
                            ws.queue_legal[at] = ylegal; ws.queue_action[at] = (unsigned char)lane;
                        }
                        if ((nh >> lane) & 1u) {
                            const int at = qn + c_lo + popc(nh & lt);
                            ws.queue[at] = to_u4(apply_action(node, mover, 32 + (int)lane));
                            ws.queue_legal[at] = ylegal; ws.queue_action[at] = (unsigned char)(32 + lane);
                        }
                        qn += c_lo + popc(nh);
                        __syncwarp();
                        while (qn >= 32) {
                            qn -= 32;
                            int nm, nw;
                            leaf_counts_delta(to_state(ws.queue[qn + lane]), ws.queue_legal[qn + lane], ws.queue_action[qn + lane],
                                              s_lut, nm, nw);
                            leaf_nodes += nm; leaf_wins += nw; leaf_blocked += nm == 0;
                            __syncwarp();
                        }
                    } else if (nl | nh) {
                        if (lane == 0) { ws.frame_state[sp] = to_u4(node); ws.frame_mask[sp] = ((u64)nh << 32) | nl; }
                        ++sp;
                        __syncwarp();
                    }
                }
            }
            if (sp == 0) break;
            // ---- next unvisited child of the top frame
            const u64 fm = ws.frame_mask[sp - 1];
            __syncwarp();
            if (fm == 0) { --sp; expand = false; continue; }
            if (lane == 0) ws.frame_mask[sp - 1] = fm & (fm - 1);
            const State4 parent = to_state(ws.frame_state[sp - 1]);
            node = apply_action(parent, (item_parity + sp - 1) & 1, __ffsll((long long)fm) - 1);
            node_rem = rem0 - sp;       // frame k sits k plies below the item
            expand = true;
            __syncwarp();
        }
        // ---- drain the queue, fold this item's leaf tallies into the warp counters
        if ((int)lane < qn) {
            int nm, nw;
            leaf_counts_delta(to_state(ws.queue[lane]), ws.queue_legal[lane], ws.queue_action[lane], s_lut, nm, nw);
            leaf_nodes += nm; leaf_wins += nw; leaf_blocked += nm == 0;
        }
        __syncwarp();
        leaf_nodes = warp_sum(leaf_nodes); leaf_wins = warp_sum(leaf_wins); leaf_blocked = warp_sum(leaf_blocked);
        if (lane == 0) {
            ws.cnt[0 * kRows + depth] += leaf_nodes * m;
            ws.cnt[(1 + leaf_mover) * kRows + depth] += leaf_wins * m;
            ws.cnt[(3 + leaf_mover) * kRows + depth - 1] += leaf_blocked * m;
        }
        __syncwarp();
    }
    __syncwarp();
    for (int k = lane; k < 5 * kRows; k += 32)
        if (ws.cnt[k]) atomicAdd(counters + k, ws.cnt[k]);
}

// ------------------------------------------------------------------ lane-parallel DFS
// One frontier item per LANE.  Every loop iteration each busy lane handles exactly one node
// (descend to the next unvisited child of its top frame, evaluate it, push or return), so all
// lanes run the same instruction stream no matter how deep they are: the right shape for the
// narrow, deep subtrees of the late game, where the warp-cooperative kernel above would spend a
// whole warp on nodes with two or three children.  Frames are just the 64-bit masks of unvisited
// non-winning children (the in-flight child is the lowest set bit of its parent's mask); the
// position itself lives in registers and is rebuilt on the way up by clearing the square.
// Idle lanes pull new items from a global counter (one aggregated atomic per warp and refill).
static constexpr int kLaneWarps = 4;                  // 128 threads per CTA
static constexpr int kLaneFrames = 13;                // plies below an item with their own frame
static constexpr int kLaneLevels = kLaneFrames + 2;   // counter rows, relative to the item depth

struct LaneScratch {
    u64 frame[kLaneFrames][32];                       // [level][lane]
    unsigned cnt[3][kLaneLevels][32];                 // nodes / wins / blocked per relative depth, unweighted
    u64 acc[5 * kRows];                               // this warp's weighted share of the counters
};

__device__ __forceinline__ State4 clear_square(State4 s, int action)
{
    const uint32_t b = ~(0x00010001u << (action & 15));
    s.x &= b; s.y &= b; s.z &= b; s.w &= b;
    return s;
}

__global__ void __launch_bounds__(kLaneWarps * 32) k_perft_lane(const uint4 *__restrict__ items, const u64 *__restrict__ mult,
                                                                 u64 n_items, int d0, int depth, u64 *__restrict__ next_item,
                                                                 u64 *__restrict__ counters)
{
    __shared__ LaneScratch s_scratch[kLaneWarps];
    LaneScratch &ws = s_scratch[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u;
    for (int k = lane; k < 5 * kRows; k += 32) ws.acc[k] = 0;
    for (int k = 0; k < 3 * kLaneLevels; ++k) (&ws.cnt[0][0][0])[k * 32 + lane] = 0;
    __syncwarp();

    const int rem0 = depth - d0;          // 1 <= rem0 <= kLaneFrames + 1
    State4 cur = { 0, 0, 0, 0 };          // position of the top frame
    int sp = -1;                          // top frame (= plies below the item); -1: no item
    int parity = 0;                       // player to move at the item
    u64 m = 0;                            // multiplicity of the item
    unsigned since_flush = 0;
    bool more = true;                     // the global queue may still have items

    for (;;) {
        // ---- idle lanes take the next items
        State4 node = cur;
        int level = sp + 1;
        bool have = sp >= 0;
        const unsigned idle = __ballot_sync(kFull, sp < 0);
        if (idle && more) {
            u64 base = 0;
            if (lane == (unsigned)(__ffs((int)idle) - 1)) base = atomicAdd(next_item, (u64)__popc(idle));
            base = __shfl_sync(kFull, base, __ffs((int)idle) - 1);
            if (sp < 0) {
                const u64 idx = base + __popc(idle & ((1u << lane) - 1u));
                if (idx < n_items) {
                    node = to_state(__ldg(items + idx));
                    m = mult[idx];
                    parity = piece_balance(node);
                    level = 0;
                    have = true;
                }
            }
            more = base + __popc(idle) < n_items;
        }
        if (!__any_sync(kFull, have)) break;

        if (have) {
            if (level > 0) {   // descend: the in-flight child is the lowest set bit of the parent's mask
                const u64 fm = ws.frame[sp][lane];
                node = apply_action(cur, (parity + sp) & 1, __ffsll((long long)fm) - 1);
            }
            // ---- evaluate the node: its moves, its winning moves (game_stats.py:384-424)
            const int mover = (parity + level) & 1;
            uint32_t lo, hi, wab, wcd;
            legal_masks(node, mover, lo, hi);
            winning_squares(node, wab, wcd);
            const int n_moves = popc(lo) + popc(hi);
            const int n_wins = popc(lo & wab) + popc(hi & wcd);
            const uint32_t nl = lo & ~wab, nh = hi & ~wcd;
            ws.cnt[0][level + 1][lane] += (unsigned)n_moves;
            ws.cnt[1][level + 1][lane] += (unsigned)n_wins;
            ws.cnt[2][level][lane] += n_moves == 0 ? 1u : 0u;
            since_flush += 64;
            bool done = false;                                   // item finished
            if (rem0 - level >= 2 && (nl | nh)) {
                cur = node; sp = level;                          // push
                ws.frame[level][lane] = ((u64)nh << 32) | nl;
            } else if (level == 0) {
                done = true;                                     // nothing to visit below the item
            } else {
                // return from the child, then from every frame it exhausts
                u64 fm = ws.frame[sp][lane];
                fm &= fm - 1;
                ws.frame[sp][lane] = fm;
                while (fm == 0) {
                    if (--sp < 0) { done = true; break; }
                    fm = ws.frame[sp][lane];
                    cur = clear_square(cur, __ffsll((long long)fm) - 1);
                    fm &= fm - 1;
                    ws.frame[sp][lane] = fm;
                }
            }
            // ---- fold the lane's unweighted tallies into the warp's weighted counters
            if (done || since_flush >= 0x70000000u) {
                for (int r = 0; r <= rem0; ++r) {
                    const unsigned c_nodes = ws.cnt[0][r][lane], c_wins = ws.cnt[1][r][lane], c_blk = ws.cnt[2][r][lane];
                    ws.cnt[0][r][lane] = 0; ws.cnt[1][r][lane] = 0; ws.cnt[2][r][lane] = 0;
                    const int e = d0 + r;                        // absolute depth; ply e was played by (parity + r - 1) & 1
                    if (c_nodes) atomicAdd(&ws.acc[0 * kRows + e], (u64)c_nodes * m);
                    if (c_wins) atomicAdd(&ws.acc[(1 + ((parity + r + 1) & 1)) * kRows + e], (u64)c_wins * m);
                    if (c_blk) atomicAdd(&ws.acc[(3 + ((parity + r) & 1)) * kRows + e], (u64)c_blk * m);
                }
                since_flush = 0;
            }
            if (done) sp = -1;
        }
    }
    __syncwarp();
    for (int k = lane; k < 5 * kRows; k += 32)
        if (ws.acc[k]) atomicAdd(counters + k, ws.acc[k]);
}

// ------------------------------------------------------------------ lane-parallel DFS with a dealt last level
// k_perft_lane's walk (one item per lane, frames = masks of unvisited children), except that the LAST interior
// level is not walked: when a lane evaluates a node two plies above the target, its non-winning children -- the
// leaf parents, 96 % of all nodes that are ever evaluated in a wide tree -- are listed in a warp-shared queue
// together with those of the other lanes and dealt out 32 per round.  So the dominant work runs with full warps
// whatever the fan-out of the individual nodes, and each leaf parent is evaluated from its parent's data
// (leaf_counts_delta) instead of from scratch.  Interior nodes count unweighted into one packed word per
// (level, lane); the words are folded into the warp's weighted counters by warp-wide sums when a lane finishes
// its item (no shared-memory atomics: 64-bit ones are CAS loops and every lane would hit the same row).
// The target depth accumulates weighted in registers.
static constexpr int kTileWarps = 4;                  // 128 threads per CTA
static constexpr int kTileMaxRem = kLaneFrames;       // 4 warps x (fixed part + 2 x (rem - 1) x 256 B) must stay below 48 KB
static constexpr unsigned kTileQueue = 1024;          // ring of dealt entries: < 32 left over + what one pass adds
static constexpr unsigned kTileParents = 64;          // ring of parent records: <= 31 still referenced + 32 new
static constexpr unsigned kTileOwnLanes = 24;         // givers in the warp from which on each evaluates the common part of its list itself
static constexpr unsigned kTileFlushEvery = 4000;     // evaluations per lane before its packed counters could overflow

// what a dealt leaf parent needs from its parent: one 48-byte record = three vector loads from one base address
struct __attribute__((aligned(16))) ParentRec {
    uint4 rz;                 // line presence: rows (A|B, C|D), zones (A|B, C|D)
    uint4 cl;                 // line presence: columns (A|B, C|D); what the player NOT to move at the parent could play (lo, hi)
    u64 mult;                 // the item's multiplicity
    u64 pad;
};
// per-warp shared memory; the two per-level arrays are sized for the job (dynamic shared memory), so shallow
// remainders -- the wide, early part of the game -- run at a higher occupancy
struct TileScratch {
    unsigned short queue[kTileQueue];                 // ring: (mover << 12 | parent record << 6 | action) of a dealt leaf parent
    ParentRec parent[kTileParents];                   // ring of parent records (nodes two plies above the target)
    u64 acc[5 * kRows];                               // this warp's weighted share of the counters
    u64 levels[1][32];                                // frame[rem0 - 1][32]: masks of unvisited children per level and lane,
                                                      // then cnt[rem0 - 1][32]: moves (bits 0-25) | winning moves (26-51) | blocked (52-63)
};
static size_t tile_warp_bytes(int rem0) { return (sizeof(TileScratch) - 256 + (size_t)(2 * rem0 - 2) * 256 + 15) / 16 * 16; }

// weighted tallies of the target depth, per lane.  FAST = every item has the same player to move and every
// multiplicity fits 32 bits (true for any frontier grown from one root, and for the depth-4 classes): one
// accumulator per counter and one IMAD.WIDE per addition instead of predicated 64 x 64-bit arithmetic.
template <bool FAST> struct LeafAcc;
template <> struct LeafAcc<false> {
    u64 nodes = 0, w0 = 0, w1 = 0, b0 = 0, b1 = 0;
    __device__ __forceinline__ void add(int nm, int nw, u64 pm, bool leaf_mover_p1)
    {
        nodes += (u64)nm * pm;
        const u64 w = (u64)nw * pm, b = nm ? 0ull : pm;
        if (leaf_mover_p1) { w1 += w; b1 += b; } else { w0 += w; b0 += b; }
    }
    __device__ __forceinline__ void add_sums(uint32_t nm, uint32_t nw, uint32_t blocked, u64 pm, bool leaf_mover_p1)
    {
        nodes += (u64)nm * pm;
        const u64 w = (u64)nw * pm, b = (u64)blocked * pm;
        if (leaf_mover_p1) { w1 += w; b1 += b; } else { w0 += w; b0 += b; }
    }
    __device__ __forceinline__ void finish(bool) {}
};
template <> struct LeafAcc<true> {
    u64 nodes = 0, w0 = 0, w1 = 0, b0 = 0, b1 = 0;         // w1 / b1 are filled by finish()
    __device__ __forceinline__ void add(int nm, int nw, u64 pm, bool)
    {
        const uint32_t m32 = (uint32_t)pm;
        nodes += (u64)((uint32_t)nm) * m32;
        w0 += (u64)((uint32_t)nw) * m32;
        b0 += nm ? 0u : m32;
    }
    __device__ __forceinline__ void add_sums(uint32_t nm, uint32_t nw, uint32_t blocked, u64 pm, bool)
    {
        const uint32_t m32 = (uint32_t)pm;
        nodes += (u64)nm * m32;
        w0 += (u64)nw * m32;
        b0 += (u64)blocked * m32;
    }
    __device__ __forceinline__ void finish(bool leaf_mover_p1)
    {
        if (leaf_mover_p1) { w1 = w0; b1 = b0; w0 = 0; b0 = 0; }
    }
};

// k_perft_tile, own lists: the children of shape S that this lane's node gives (bits of `word`'s S lane), evaluated by
// the lane itself; the loop's trip count is warp-uniform.  What is left in `word` afterwards goes to the dealt stage.
template <int S>
__device__ __forceinline__ void own_shape(uint32_t &word, uint32_t ylo, uint32_t yhi, const Presence &pres, const TileLut *lut,
                                          bool giver, unsigned mode, uint32_t n_givers, uint32_t &s_nm, uint32_t &s_nw, uint32_t &s_blk)
{
    constexpr int kShift = (S & 1) ? 16 : 0;
    uint32_t kids = giver ? (word >> kShift) & 0xFFFFu : 0u;
    const uint32_t mine = (uint32_t)popc(kids);
    // trip count: the shortest list of the warp (full lanes, the rest is dealt), the longest (nothing left to deal, shorter
    // lists idle), or the mean (most children covered, most lanes busy)
    uint32_t K;
    if (mode == 1) K = __reduce_max_sync(kFull, mine);
    else if (mode == 2) K = ((uint32_t)__reduce_add_sync(kFull, mine) + n_givers / 2) / n_givers;
    else K = __reduce_min_sync(kFull, giver ? mine : 0xFFFFFFFFu);
    if (K == 0) return;
    const ShapeLoop<S> loop(ylo, yhi, pres);
    for (uint32_t k = 0; k < K; ++k) {
        if (kids) {
            const int p = __ffs((int)kids) - 1;
            kids &= kids - 1;
            int nm, nw;
            loop.child(lut[p], nm, nw);
            s_nm += (uint32_t)nm; s_nw += (uint32_t)nw; s_blk += nm == 0 ? 1u : 0u;
        }
    }
    word = (word & ~(0xFFFFu << kShift)) | (kids << kShift);
}

// (8 resident CTAs = at most 64 registers: the own-list form took 68 and lost a CTA, perft(7) 24.1 vs 23.3 ms)
template <bool FAST, bool OWN>
__global__ void __launch_bounds__(kTileWarps * 32, 8) k_perft_tile(const uint4 *__restrict__ items, const u64 *__restrict__ mult,
                                                                 u64 n_items, int d0, int depth, int uniform_parity,
                                                                 u64 *__restrict__ next_item, u64 *__restrict__ counters, unsigned own_lanes)
{
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ __align__(16) TileLut s_lut[16];
    if (threadIdx.x < 16) s_lut[threadIdx.x] = tile_lut((int)threadIdx.x);
    const int rem0 = depth - d0;          // 2 <= rem0 <= kTileMaxRem
    const size_t warp_bytes = (sizeof(TileScratch) - 256 + (size_t)(2 * rem0 - 2) * 256 + 15) / 16 * 16;
    TileScratch &ws = *reinterpret_cast<TileScratch *>(s_dyn + (threadIdx.x >> 5) * warp_bytes);
    u64 (*const frame)[32] = ws.levels;                 // [rem0 - 1][32]
    u64 (*const cnt_packed)[32] = ws.levels + (rem0 - 1); // [rem0 - 1][32]
    const unsigned lane = threadIdx.x & 31u;
    for (int k = lane; k < 5 * kRows; k += 32) ws.acc[k] = 0;
    for (int r = 0; r <= rem0 - 2; ++r) cnt_packed[r][lane] = 0;
    __syncthreads();

    State4 cur = { 0, 0, 0, 0 };          // position of the top frame
    int sp = -1;                          // top frame (= plies below the item); -1: no item
    int parity = 0;                       // player to move at the item
    u64 m = 0;                            // multiplicity of the item
    bool more = true;                     // the global queue may still have items
    unsigned since_flush = 0;
    uint32_t q_head = 0, q_tail = 0, q_parents = 0;                       // rings (warp-uniform running counters)
    // Dealing costs the whole warp a fixed amount per iteration in which some lane gives (the other player's legal
    // squares, the scan, the list) and saves per leaf parent; late in the game nodes two plies above the target have
    // fewer than one child on average, and then the lanes walk the last level themselves (as k_perft_lane does).
    // Decided per warp every 32 iterations: deal while the busy lanes saw at least one such child per iteration each
    // (measured on the 10 946 depth-4 classes: to depth 12 dealing 11.5 s vs walking 15.6 s, to depth 16 165 s vs 128 s).
    bool deal = true;
    uint32_t seen_kids = 0, seen_busy = 0, seen_iters = 0;
    LeafAcc<FAST> acc;                                                    // the target depth, weighted, per lane

    for (;;) {
        // ---- idle lanes take the next items
        State4 node = cur;
        int level = sp + 1;
        bool have = sp >= 0;
        const unsigned idle = __ballot_sync(kFull, sp < 0);
        if (idle && more) {
            u64 base = 0;
            if (lane == (unsigned)(__ffs((int)idle) - 1)) base = atomicAdd(next_item, (u64)__popc(idle));
            base = __shfl_sync(kFull, base, __ffs((int)idle) - 1);
            if (sp < 0) {
                const u64 idx = base + __popc(idle & ((1u << lane) - 1u));
                if (idx < n_items) {
                    node = to_state(__ldg(items + idx));
                    m = mult[idx];
                    parity = piece_balance(node);
                    level = 0;
                    have = true;
                }
            }
            more = base + __popc(idle) < n_items;
        }
        if (!__any_sync(kFull, have)) break;

        uint32_t give_lo = 0, give_hi = 0;       // this lane's leaf parents for the dealt stage
        uint32_t give_ylo = 0, give_yhi = 0;
        Presence give_pres = { 0, 0, 0, 0, 0, 0 };
        int give_mover = 0;
        uint32_t kids = 0;                       // children of a node two plies above the target (either mode)
        bool flush = false;
        if (have) {
            if (level > 0) {   // descend: the in-flight child is the lowest set bit of the parent's mask
                const u64 fm = frame[sp][lane];
                node = apply_action(cur, (parity + sp) & 1, __ffsll((long long)fm) - 1);
            }
            // ---- evaluate the node: its moves, its winning moves (game_stats.py:384-424)
            const int mover = (parity + level) & 1;
            uint32_t lo, hi, wab, wcd;
            legal_masks(node, mover, lo, hi);
            give_pres = line_presence(node);
            winning_from_presence(give_pres, wab, wcd);
            const int n_moves = popc(lo) + popc(hi);
            const int n_wins = popc(lo & wab) + popc(hi & wcd);
            const uint32_t nl = lo & ~wab, nh = hi & ~wcd;
            const int rem = rem0 - level;
            bool ret = true;                                     // nothing (more) to walk below this node
            if (rem == 1) {
                // a leaf parent walked by its own lane (narrow trees): the target depth, weighted, in registers
                acc.add(n_moves, n_wins, m, mover != 0);
            } else {
                cnt_packed[level][lane] += (u64)n_moves | ((u64)n_wins << 26) | ((u64)(n_moves == 0) << 52);
                ++since_flush;
                if (rem == 2) kids = (uint32_t)(popc(nl) + popc(nh));
                if (nl | nh) {
                    if (rem >= 3 || !deal) {
                        cur = node; sp = level; ret = false;     // push
                        frame[level][lane] = ((u64)nh << 32) | nl;
                    } else {
                        // two plies above the target: hand the children to the dealt stage
                        legal_masks(node, mover ^ 1, give_ylo, give_yhi);
                        give_lo = nl; give_hi = nh; give_mover = mover;
                    }
                }
            }
            if (ret) {
                if (level == 0) {
                    sp = -1;                                     // item finished
                } else {
                    // return from the child, then from every frame it exhausts
                    u64 fm = frame[sp][lane];
                    fm &= fm - 1;
                    frame[sp][lane] = fm;
                    while (fm == 0) {
                        if (--sp < 0) break;
                        fm = frame[sp][lane];
                        cur = clear_square(cur, __ffsll((long long)fm) - 1);
                        fm &= fm - 1;
                        frame[sp][lane] = fm;
                    }
                }
            }
            flush = sp < 0 || since_flush >= kTileFlushEvery;
        }
        // ---- leaf parents, first part (wide trees only, see run_perft): when (nearly) every lane has children to give,
        //      each giver evaluates children of its own straight from its registers, shape by shape (own_shape: the
        //      lanes of a shape are compile-time masks, the other word pair's masks are hoisted out of the loop, the
        //      counts are summed unweighted and multiplied once per node -- about 50 instructions per leaf parent
        //      against 72 in the dealt stage, 83 for the generic evaluation in a loop).  The loop of a shape runs as
        //      often as the shortest / the longest / the mean list of that shape in the warp (own_lanes >> 6 = 0 / 1 / 2);
        //      what is left goes to the dealt stage.
        if (OWN) {
            const bool giver = (give_lo | give_hi) != 0;
            const uint32_t n_givers = (uint32_t)__popc(__ballot_sync(kFull, giver));
            if (n_givers >= (own_lanes & 63u) && n_givers) {
                uint32_t s_nm = 0, s_nw = 0, s_blk = 0;
                const unsigned mode = own_lanes >> 6;
                own_shape<0>(give_lo, give_ylo, give_yhi, give_pres, s_lut, giver, mode, n_givers, s_nm, s_nw, s_blk);
                own_shape<1>(give_lo, give_ylo, give_yhi, give_pres, s_lut, giver, mode, n_givers, s_nm, s_nw, s_blk);
                own_shape<2>(give_hi, give_ylo, give_yhi, give_pres, s_lut, giver, mode, n_givers, s_nm, s_nw, s_blk);
                own_shape<3>(give_hi, give_ylo, give_yhi, give_pres, s_lut, giver, mode, n_givers, s_nm, s_nw, s_blk);
                acc.add_sums(s_nm, s_nw, s_blk, m, give_mover == 0);      // the leaf parents' mover is the OTHER player
            }
        }
        seen_kids += (uint32_t)__reduce_add_sync(kFull, kids);
        seen_busy += (uint32_t)__popc(__ballot_sync(kFull, have));
        if (++seen_iters == 32) { deal = seen_kids >= seen_busy; seen_kids = 0; seen_busy = 0; seen_iters = 0; }
        // ---- the dealt stage: the givers' leaf parents join a ring shared by the warp; full rounds of 32 are
        //      evaluated at once, fewer than 32 stay for the next iteration.  Entries name a parent RECORD (not a
        //      lane), so they survive their lane moving on.
        const uint32_t cnt_all = (uint32_t)(popc(give_lo) + popc(give_hi));
        const unsigned givers = __ballot_sync(kFull, cnt_all != 0);
        if (givers) {
            const uint32_t rec = (q_parents + (uint32_t)__popc(givers & ((1u << lane) - 1u))) & (kTileParents - 1);
            if (cnt_all) {
                ws.parent[rec].rz = make_uint4(give_pres.r01, give_pres.r23, give_pres.z01, give_pres.z23);
                ws.parent[rec].cl = make_uint4(give_pres.c01, give_pres.c23, give_ylo, give_yhi);
                ws.parent[rec].mult = m;
            }
            q_parents += (uint32_t)__popc(givers);
            const uint32_t n_all = (uint32_t)__reduce_add_sync(kFull, cnt_all);
            const int n_pass = n_all + (q_tail - q_head) <= kTileQueue ? 1 : 4;     // 8 lanes x 64 children always fit
            for (int pass = 0; pass < n_pass; ++pass) {
                const uint32_t cnt = (n_pass == 1 || (int)(lane >> 3) == pass) ? cnt_all : 0u;
                uint32_t incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(kFull, incl, o); if (lane >= (unsigned)o) incl += t; }
                if (cnt) {
                    uint32_t at = q_tail + incl - cnt;
                    const uint32_t tag = ((uint32_t)give_mover << 12) | (rec << 6);
                    for (uint32_t w = give_lo; w; w &= w - 1) ws.queue[at++ & (kTileQueue - 1)] = (unsigned short)(tag | (uint32_t)(__ffs((int)w) - 1));
                    for (uint32_t w = give_hi; w; w &= w - 1) ws.queue[at++ & (kTileQueue - 1)] = (unsigned short)(tag | (uint32_t)(32 + __ffs((int)w) - 1));
                }
                q_tail += __shfl_sync(kFull, incl, 31);
                __syncwarp();
                while (q_tail - q_head >= 32) {
                    const uint32_t e = ws.queue[(q_head + lane) & (kTileQueue - 1)], pr = (e >> 6) & (kTileParents - 1), xm = e >> 12;
                    const int action = (int)(e & 63u);
                    const uint4 rz = ws.parent[pr].rz, cl = ws.parent[pr].cl;
                    const Presence pres = { rz.x, rz.y, rz.z, rz.w, cl.x, cl.y };
                    const u64 pm = FAST ? (u64)*reinterpret_cast<const uint32_t *>(&ws.parent[pr].mult) : ws.parent[pr].mult;
                    int nm, nw;
                    leaf_counts_from_parent(cl.z, cl.w, pres, action, s_lut, nm, nw);
                    acc.add(nm, nw, pm, xm == 0);                     // the leaf parent's mover is the OTHER player
                    q_head += 32;
                }
                __syncwarp();
            }
        }
        // ---- fold finished lanes' packed counters into the warp's weighted counters (warp-wide sums)
        if (__any_sync(kFull, flush)) {
            for (int r = 0; r <= rem0 - 2; ++r) {
                const u64 packed = flush ? cnt_packed[r][lane] : 0ull;
                if (!__any_sync(kFull, packed != 0)) continue;
                if (flush) cnt_packed[r][lane] = 0;
                const u64 c_nodes = (packed & 0x3FFFFFFull) * m, c_wins = ((packed >> 26) & 0x3FFFFFFull) * m, c_blk = (packed >> 52) * m;
                const bool p1 = ((parity + r) & 1) != 0;         // the player to move at this level
                const u64 s_nodes = warp_sum(c_nodes);
                const u64 s_w0 = warp_sum(p1 ? 0ull : c_wins), s_w1 = warp_sum(p1 ? c_wins : 0ull);
                const u64 s_b0 = warp_sum(p1 ? 0ull : c_blk), s_b1 = warp_sum(p1 ? c_blk : 0ull);
                if (lane == 0) {
                    const int d = d0 + r;
                    ws.acc[0 * kRows + d + 1] += s_nodes;                                      // :401
                    ws.acc[1 * kRows + d + 1] += s_w0; ws.acc[2 * kRows + d + 1] += s_w1;      // :411-415
                    ws.acc[3 * kRows + d] += s_b0; ws.acc[4 * kRows + d] += s_b1;              // :426-443
                }
            }
            if (flush) since_flush = 0;
        }
    }
    // ---- what is left in the ring (< 32 entries)
    if (lane < q_tail - q_head) {
        const uint32_t e = ws.queue[(q_head + lane) & (kTileQueue - 1)], pr = (e >> 6) & (kTileParents - 1), xm = e >> 12;
        const int action = (int)(e & 63u);
        const uint4 rz = ws.parent[pr].rz, cl = ws.parent[pr].cl;
        const Presence pres = { rz.x, rz.y, rz.z, rz.w, cl.x, cl.y };
        const u64 pm = ws.parent[pr].mult;
        int nm, nw;
        leaf_counts_from_parent(cl.z, cl.w, pres, action, s_lut, nm, nw);
        acc.add(nm, nw, pm, xm == 0);
    }
    acc.finish(((uniform_parity + rem0 - 1) & 1) != 0);     // who moves one ply above the target
    const u64 t_nodes = warp_sum(acc.nodes), t_w0 = warp_sum(acc.w0), t_w1 = warp_sum(acc.w1);
    const u64 t_b0 = warp_sum(acc.b0), t_b1 = warp_sum(acc.b1);
    if (lane == 0) {
        ws.acc[0 * kRows + depth] += t_nodes;
        ws.acc[1 * kRows + depth] += t_w0;
        ws.acc[2 * kRows + depth] += t_w1;
        ws.acc[3 * kRows + depth - 1] += t_b0;
        ws.acc[4 * kRows + depth - 1] += t_b1;
    }
    __syncwarp();
    for (int k = lane; k < 5 * kRows; k += 32)
        if (ws.acc[k]) atomicAdd(counters + k, ws.acc[k]);
}

// ------------------------------------------------------------------ cost probe for the static partitioning
// cost[i] = move sequences of length 2 below root i: what quantik_b200.dist sorts the roots by before it deals
// them to the ranks (SURVEY 8e "LPT/greedy by probe cost")
__global__ void __launch_bounds__(kBlock) k_perft_probe(const uint4 *__restrict__ roots, size_t n, uint32_t *__restrict__ cost)
{
    const size_t i = blockIdx.x * (size_t)kBlock + threadIdx.x;
    if (i >= n) return;
    const State4 s = to_state(__ldg(roots + i));
    uint32_t c = 0;
    int player;
    if (win_status(s) == 0 && validate(s, &player) == 0) {
        uint32_t lo, hi, wab, wcd;
        legal_masks(s, player, lo, hi);
        winning_squares(s, wab, wcd);
        for (uint32_t w = lo & ~wab; w; w &= w - 1) {
            int nm, nw;
            leaf_parent_counts_for(apply_action(s, player, __ffs((int)w) - 1), player ^ 1, nm, nw);
            c += (uint32_t)nm;
        }
        for (uint32_t w = hi & ~wcd; w; w &= w - 1) {
            int nm, nw;
            leaf_parent_counts_for(apply_action(s, player, 32 + __ffs((int)w) - 1), player ^ 1, nm, nw);
            c += (uint32_t)nm;
        }
    }
    cost[i] = c;
}

int launch_perft_probe(const uint4 *roots, size_t n, uint32_t *cost, Ctx *c, cudaStream_t s)
{
    (void)c;
    k_perft_probe<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, s>>>(roots, n, cost);
    QK_CUDA(cudaGetLastError());
    return QK_OK;
}

// ------------------------------------------------------------------ host driver
int run_perft(Call &call, const uint4 *roots, const uint64_t *mult, size_t n_roots, int depth, uint64_t *host_out)
{
    cudaStream_t s = call.stream();
    Ctx *c = call.ctx();
    // device block: counters[5][17], frontier size, dfs work counter, root flags (2 words)
    u64 *d_ctr = (u64 *)call.temp((5 * kRows + 3) * sizeof(u64), true);
    if (!d_ctr) return call.status();
    u64 *d_count = d_ctr + 5 * kRows, *d_next = d_count + 1;
    unsigned *d_flags = (unsigned *)(d_next + 1);

    uint4 *cur_s = (uint4 *)call.temp(n_roots * 16);
    u64 *cur_m = (u64 *)call.temp(n_roots * 8);
    if (!cur_s || !cur_m) return call.status();
    k_perft_filter_roots<<<(unsigned)((n_roots + kBlock - 1) / kBlock), kBlock, 0, s>>>(
        roots, (const u64 *)mult, n_roots, depth, cur_s, cur_m, d_count, d_ctr, d_flags);
    QK_CUDA(cudaGetLastError());
    struct { u64 count, next; unsigned flags, max_pieces; } head = { 0, 0, 0, 0 };      // = d_count, d_next, d_flags[0..1]
    QK_CUDA(cudaMemcpyAsync(&head, d_count, sizeof head, cudaMemcpyDeviceToHost, s));
    QK_CUDA(cudaStreamSynchronize(s));
    u64 n_cur = head.count;
    const unsigned root_flags = head.flags;
    const bool one_side = (root_flags & 3u) == 1u || (root_flags & 3u) == 2u, small_mult = (root_flags & 4u) == 0;
    const int root_parity = (root_flags & 2u) ? 1 : 0;

    // Widen level by level until every lane can have its own items, then walk depth-first with k_perft_tile
    // (one item per lane, last level dealt out to full warps).  Measured from the empty board, in ms, for
    // tile / warp-per-item / plain lane-per-item: depth 6 1.7 / 3.5 / 4.0, depth 7 38 / 105 / 62, depth 8 835 / 2688 / 1290.
    u64 want_warp_items = (u64)c->sms * 4 * kWarps * 8;
    // (r02c_experiments.json: the tile kernel is at its best with 4 plies below the items -- the share of one of 8
    // ranks, 1.2 M items: 23.6 ms started there vs 25.9 ms one level later -- unless the items are about as few as
    // the lanes: perft(7) from 167 K items 28.4 ms vs 25.0 ms from 6.7 M.  4 items per lane separates the cases.)
    //  Deep remainders are the opposite case: subtrees 10 plies deep differ in size by orders of magnitude, and only many
    //  items per lane even that out -- config 5 at full depth on 8 GPUs: 22.2 s started from 1.2 M items per rank, 16 s
    //  from 50 M -- so the low threshold applies to remainders of at most 4 plies.)
    const u64 want_lane_items_shallow = (u64)c->sms * 1024 * 4, want_lane_items_deep = (u64)c->sms * 2048 * 8;
    u64 want_lane_items = 0;
    const int force = (int)option(OPT_PERFT_KERNEL);                    // 1 tile / 2 lane / 3 warp: benchmarking aid
    if (const long long mi = option(OPT_PERFT_MIN_ITEMS))               // testing aid: depth-first kernels on small frontiers
        want_warp_items = want_lane_items = (u64)mi;
    int d = 0;
    // Small levels first, chained on the device: while even 64 children per position cannot reach the size at which
    // the depth-first kernel takes over, the next frontier is allocated at that bound and its size stays on the
    // device (the next expansion reads it there) -- no sizing pass, no host round trip per level.
    {
        u64 bound = n_cur;
        u64 *d_sizes = nullptr;
        const u64 *d_n = nullptr;                               // size of the current frontier on the device, once chained
        while (n_cur > 0 && depth - d >= 2 && bound <= 4096 && bound * 64 < want_lane_items_shallow && !force && !option(OPT_PERFT_MIN_ITEMS)) {
            if (!d_sizes) {
                d_sizes = (u64 *)call.temp((QK_MAX_DEPTH + 1) * sizeof(u64), true);
                if (!d_sizes) return call.status();
            }
            uint4 *nxt_s = (uint4 *)call.temp(bound * 64 * 16);
            u64 *nxt_m = (u64 *)call.temp(bound * 64 * 8);
            if (!nxt_s || !nxt_m) return call.status();
            k_perft_expand<false><<<(unsigned)((bound + kBlock - 1) / kBlock), kBlock, 0, s>>>(cur_s, cur_m, bound, d, nxt_s, nxt_m, d_sizes + d, d_ctr, d_n);
            QK_CUDA(cudaGetLastError());
            call.release(cur_s); call.release(cur_m);
            cur_s = nxt_s; cur_m = nxt_m;
            d_n = d_sizes + d;
            bound *= 64;
            ++d;
        }
        if (d_n) {                                              // back to exact sizes
            QK_CUDA(cudaMemcpyAsync(&n_cur, d_n, sizeof(u64), cudaMemcpyDeviceToHost, s));
            QK_CUDA(cudaStreamSynchronize(s));
        }
    }
    while (n_cur > 0 && d < depth) {
        const int rem = depth - d;
        if (rem == 1) {
            k_perft_leaf<<<grid_for(n_cur, kBlock, c->sms, 8), kBlock, 0, s>>>(cur_s, cur_m, n_cur, d, d_ctr);
            QK_CUDA(cudaGetLastError());
            break;
        }
        const bool can_lane = rem <= kLaneFrames + 1;
        const bool can_tile = rem >= 2 && rem <= kTileMaxRem;
        // As soon as every lane can have an item: the lane walk with the dealt last level (k_perft_tile) while the
        // target lies in the wide part of the game, the plain lane walk for the narrow end (measured below the
        // 10 946 depth-4 classes, tile / lane: to depth 12 11.7 / 15.6 s, to depth 14 108 / 91 s, to depth 16 147 / 128 s)
        bool go = n_cur >= (want_lane_items ? want_lane_items : (rem <= 4 ? want_lane_items_shallow : want_lane_items_deep));
        u64 n_next = 0;
        if (!go) {
            // sizing pass: how many children this level emits (the frontier is allocated exactly; a level that
            // would not be worth materialising is walked depth-first instead)
            QK_CUDA(cudaMemsetAsync(d_count, 0, sizeof(u64), s));
            k_perft_expand<true><<<(unsigned)((n_cur + kBlock - 1) / kBlock), kBlock, 0, s>>>(cur_s, cur_m, n_cur, d, nullptr, nullptr,
                                                                                           d_count, d_ctr);
            QK_CUDA(cudaGetLastError());
            QK_CUDA(cudaMemcpyAsync(&n_next, d_count, sizeof(u64), cudaMemcpyDeviceToHost, s));
            QK_CUDA(cudaStreamSynchronize(s));
            go = n_next > (1ull << 28);
        }
        bool use_tile = false, use_lane = false;
        if (go && !force) {
            const int target_pieces = (int)head.max_pieces + depth;      // the fullest root decides
            use_tile = can_tile && target_pieces <= 12;
            use_lane = !use_tile && can_lane;
            use_tile = use_tile || (!use_lane && can_tile);
        }
        bool use_warp = go && !use_tile && !use_lane;           // (cannot happen for depth <= 16; the warp kernel is kept for A/B runs)
        if (force) {                                            // benchmarking aid: the two older kernels
            use_tile = force == 1 && can_tile && go;
            use_lane = force == 2 && can_lane && go;
            use_warp = force == 3 && (n_cur >= want_warp_items || go);
        }
        if (use_tile || use_lane || use_warp) {
            QK_CUDA(cudaMemsetAsync(d_next, 0, sizeof(u64), s));
            if (use_tile) {
                int per_sm = 0;
                const size_t smem = kTileWarps * tile_warp_bytes(rem);
                const bool fast = one_side && small_mult && !option(OPT_PERFT_TILE_GENERIC);
                // Own lists (k_perft_tile<., true>) pay in the widest part of the game, where every lane gives in nearly every
                // iteration and the lists of a shape are about equally long.  Givers with at most 4 pieces: the loops run to
                // the LONGEST list of the warp (shorter ones idle) and nothing is left for the dealt stage; with 5 pieces: to
                // the MEAN length, the excess is dealt; beyond: not used.  From the empty board, never / longest / mean
                // (scripts/perft_own_lanes.py): perft(6) 1.28 / 0.83 / 0.87 ms, perft(7) 24.9 / 24.1 / 23.0 ms, perft(8)
                // 533 / 583 / 532 ms, the depth-4 classes + 6 plies 181 / 254 / 202 ms.
                // "perft_own_lanes": 33 never | n: always, from n giving lanes on, to the shortest list | 64 + n: ... to the
                // longest | 128 + n: ... to the mean length.
                const long long own_opt = option(OPT_PERFT_OWN_LANES);
                const int target_pieces_own = (int)head.max_pieces + depth;
                const bool own = own_opt == 0 ? target_pieces_own <= 7 : own_opt != 33;
                const unsigned own_arg = own_opt > 0 ? (unsigned)own_opt : (target_pieces_own <= 6 ? 64u : 128u) + kTileOwnLanes;
                typedef void (*TileFn)(const uint4 *, const u64 *, u64, int, int, int, u64 *, u64 *, unsigned);
                const TileFn tile = fast ? (own ? (TileFn)k_perft_tile<true, true> : (TileFn)k_perft_tile<true, false>)
                                         : (own ? (TileFn)k_perft_tile<false, true> : (TileFn)k_perft_tile<false, false>);
                QK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tile, kTileWarps * 32, smem));
                tile<<<c->sms * (per_sm > 0 ? per_sm : 1), kTileWarps * 32, smem, s>>>(cur_s, cur_m, n_cur, d, depth,
                                                                                     (root_parity + d) & 1, d_next, d_ctr,
                                                                                     own_arg);
            } else if (use_lane) {
                int per_sm = 0;
                QK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_perft_lane, kLaneWarps * 32, 0));
                k_perft_lane<<<c->sms * (per_sm > 0 ? per_sm : 1), kLaneWarps * 32, 0, s>>>(cur_s, cur_m, n_cur, d, depth, d_next, d_ctr);
            } else {
                k_perft_dfs<<<c->sms * 4, kBlock, 0, s>>>(cur_s, cur_m, n_cur, d, depth, d_next, d_ctr);
            }
            QK_CUDA(cudaGetLastError());
            break;
        }
        uint4 *nxt_s = (uint4 *)call.temp((n_next ? n_next : 1) * 16);
        u64 *nxt_m = (u64 *)call.temp((n_next ? n_next : 1) * 8);
        if (!nxt_s || !nxt_m) return call.status();
        QK_CUDA(cudaMemsetAsync(d_count, 0, sizeof(u64), s));
        k_perft_expand<false><<<(unsigned)((n_cur + kBlock - 1) / kBlock), kBlock, 0, s>>>(cur_s, cur_m, n_cur, d, nxt_s, nxt_m, d_count, d_ctr);
        QK_CUDA(cudaGetLastError());
        call.release(cur_s); call.release(cur_m);               // stream-ordered: freed after the expansion has read them
        n_cur = n_next;
        cur_s = nxt_s; cur_m = nxt_m;
        ++d;
    }
    QK_CUDA(cudaMemcpyAsync(host_out, d_ctr, 5 * kRows * sizeof(u64), cudaMemcpyDeviceToHost, s));
    QK_CUDA(cudaStreamSynchronize(s));
    return QK_OK;
}

}  // namespace qk
