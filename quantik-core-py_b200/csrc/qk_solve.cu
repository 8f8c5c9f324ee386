// qk_solve.cu -- exact game value with mate distance for a batch of (late-game) positions:
// the value MinimaxEngine.solve (minimax.py:182-204, _negamax :330-453) proves, i.e. the negamax of
//     terminal (line complete, or mover without a move)  ->  -(W - ply)
//     otherwise                                          ->  max over moves of -value(child)
// without the accelerators that never change it (alpha-beta, transposition table, sibling dedup).
// One position per LANE, depth-first with the frame layout of k_perft_lane: a frame is the 64-bit mask of
// unvisited moves (the move in flight is its lowest set bit) plus the best value / move so far; the position
// lives in registers and is rebuilt on the way up.  Two exact short-cuts keep the tree small:
//   * a position with a move that completes a line has value W - (ply + 1), the maximum any node at that
//     ply can have (winning_squares gives those moves without generating children);
//   * a frame stops as soon as its best value reaches that maximum (a move that leaves the opponent without a
//     reply is worth as much as completing a line).
// Intended for the endgame hand-off of HybridPlayer (hybrid.py:80-122; <= 8 empty cells); deeper positions
// are legal inputs but cost exponentially more.
#include "qk_common.cuh"
#include "qk_perft_core.cuh"

namespace qk {

static constexpr int kWarpsPerBlock = 4;
static constexpr unsigned kFull = 0xFFFFFFFFu;
static constexpr int W = 64;
typedef unsigned long long u64;

struct SolveScratch {
    u64 mask[16][32];          // [level][lane] unvisited moves of the frame
    int best[16][32];          // best value so far (side to move of the frame), low 8 bits: best action
};

__device__ __forceinline__ State4 clear_sq(State4 s, int action)
{
    const uint32_t b = ~(0x00010001u << (action & 15));
    s.x &= b; s.y &= b; s.z &= b; s.w &= b;
    return s;
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_solve_lane(const uint4 *__restrict__ states, u64 n, u64 *__restrict__ next_item,
                                                                     int8_t *__restrict__ outcome, uint8_t *__restrict__ plies,
                                                                     uint8_t *__restrict__ best_action)
{
    __shared__ SolveScratch s_scratch[kWarpsPerBlock];
    SolveScratch &ws = s_scratch[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u;
    State4 cur = { 0, 0, 0, 0 };
    int sp = -1, parity = 0;
    u64 item = 0;
    bool more = true;

    for (;;) {
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
                if (idx < n) {
                    node = to_state4(__ldg(states + idx));
                    item = idx;
                    level = 0;
                    have = true;
                }
            }
            more = base + __popc(idle) < n;
        }
        if (!__any_sync(kFull, have)) break;
        if (!have) continue;

        int value = 0, act = 255;          // value of `node` for its side to move, once known
        bool known = false;
        if (level == 0) {
            // the root may be anything: finished, invalid, or a live position
            int player = 0;
            const int code = validate(node, &player);
            parity = player;
            if (has_winning_line(node) || code != 0) { value = -(W - 0); known = true; }   // minimax.py:354-358; invalid -> no moves
        } else {
            const u64 fm = ws.mask[sp][lane];
            node = apply_action(cur, (parity + sp) & 1, __ffsll((long long)fm) - 1);
        }
        if (!known) {
            const int mover = (parity + level) & 1;
            uint32_t lo, hi, wab, wcd;
            legal_masks(node, mover, lo, hi);
            if ((lo | hi) == 0) { value = -(W - level); known = true; }                     // :366-369
            else {
                winning_squares(node, wab, wcd);
                const uint32_t wl = lo & wab, wh = hi & wcd;
                if (wl | wh) {                                                              // a move completes a line
                    value = W - (level + 1); known = true;
                    act = wl ? __ffs((int)wl) - 1 : 32 + __ffs((int)wh) - 1;
                    if (level == 0) {
                        // the reported move is the LOWEST optimal action: a lower move that leaves the
                        // opponent without a reply is worth exactly as much
                        const u64 lower = (((u64)hi << 32) | lo) & ((1ull << act) - 1ull);
                        for (u64 m = lower; m; m &= m - 1) {
                            const int a = __ffsll((long long)m) - 1;
                            uint32_t clo, chi;
                            legal_masks(apply_action(node, mover, a), 1 - mover, clo, chi);
                            if ((clo | chi) == 0) { act = a; break; }
                        }
                    }
                } else {
                    cur = node; sp = level;                                                 // push: every move must be searched
                    ws.mask[level][lane] = ((u64)hi << 32) | lo;
                    ws.best[level][lane] = (-1000) * 256 + 255;
                }
            }
        }
        // ---- fold known values upwards
        while (known) {
            if (level == 0) {                                                               // the root's value
                outcome[item] = (int8_t)(value > 0 ? 1 : -1);
                plies[item] = (uint8_t)(W - (value > 0 ? value : -value));
                best_action[item] = (uint8_t)act;
                sp = -1;
                break;
            }
            // `value` belongs to the child reached by the lowest set bit of frame sp (= level - 1)
            u64 fm = ws.mask[sp][lane];
            const int a = __ffsll((long long)fm) - 1;
            fm &= fm - 1;
            int best = ws.best[sp][lane] >> 8, best_a = ws.best[sp][lane] & 255;
            if (-value > best) { best = -value; best_a = a; }                               // :427-430 (first maximum)
            // W - (sp + 1) (the opponent is left without a reply) is the most this frame can be worth
            if (fm != 0 && best < W - (sp + 1)) {
                ws.mask[sp][lane] = fm;
                ws.best[sp][lane] = best * 256 + best_a;
                known = false;
            } else {                                                                        // frame finished: its value is known
                value = best; act = best_a; level = sp;
                --sp;
                if (sp >= 0) cur = clear_sq(cur, __ffsll((long long)ws.mask[sp][lane]) - 1);
            }
        }
    }
}

int launch_solve(const uint4 *states, size_t n, unsigned long long *counter, int8_t *outcome, uint8_t *plies, uint8_t *best_action,
                 Ctx *c, cudaStream_t s)
{
    QK_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), s));
    int per_sm = 0;
    QK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_lane, kWarpsPerBlock * 32, 0));
    if (per_sm < 1) per_sm = 1;
    size_t want = (n + kWarpsPerBlock * 32 - 1) / (kWarpsPerBlock * 32), cap = (size_t)c->sms * per_sm;
    k_solve_lane<<<(unsigned)(want < cap ? want : cap), kWarpsPerBlock * 32, 0, s>>>(states, n, counter, outcome, plies, best_action);
    QK_CUDA(cudaGetLastError());
    return QK_OK;
}

}  // namespace qk
