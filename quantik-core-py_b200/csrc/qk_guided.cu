// qk_guided.cu -- SURVEY 8f-4: the reference's fitted-linear evaluation (evaluation.py:130-212) in batch,
// and playouts under MCTSEngine._select_rollout_move's eval-guided epsilon-greedy policy (mcts.py:345-383).
// One position / one playout per thread.  A guided ply scores every child of the position (up to 64 evaluations);
// the evaluation looks at all 12 lines at once in packed 4-bit fields that the playout carries along
// (qk_eval_core.cuh: EvalPos), so a child costs about 100 integer instructions plus 11 float64 operations.
#include "qk_common.cuh"
#include "qk_eval_core.cuh"

namespace qk {

static constexpr unsigned kBlock = 128;

struct EvalWeights { double w[6]; };

__global__ void __launch_bounds__(kBlock) k_eval_features(const uint4 *__restrict__ states, const uint8_t *__restrict__ player,
                                                           size_t n, double *__restrict__ features)
{
    for (size_t i = blockIdx.x * (size_t)kBlock + threadIdx.x; i < n; i += (size_t)gridDim.x * kBlock) {
        const uint4 v = __ldg(states + i);
        const State4 s = { v.x, v.y, v.z, v.w };
        int f[6];
        eval_features(s, player[i] & 1, f);
#pragma unroll
        for (int k = 0; k < 6; ++k) features[6 * i + k] = (double)f[k];
    }
}

// the greedy line of every root (qk_eval_core.cuh: GuidedLine), one thread per root
__global__ void __launch_bounds__(kBlock) k_guided_lines(const uint4 *__restrict__ roots, uint64_t n_roots, const EvalWeights weights,
                                                          GuidedLine *__restrict__ lines)
{
    const uint64_t i = blockIdx.x * (uint64_t)kBlock + threadIdx.x;
    if (i >= n_roots) return;
    const uint4 v = __ldg(roots + i);
    const State4 s = { v.x, v.y, v.z, v.w };
    lines[i] = guided_line(s, weights.w);
}

// ---- the playout kernel: 32 playouts per warp, one ply per iteration for all of them.
// A lane whose playout is still on its root's greedy line takes the line's move; a lane that needs a greedy move of
// its own has to score every child of its position.  Unless nearly all lanes are in that situation with about the
// same number of children, the warp scores the children of all of them together, dealt out one per lane
// (dealt_greedy_actions; a lane-serial scan keeps the other lanes waiting for the longest one); otherwise every
// lane scans its own children.  Same moves either way (first decisive child, else first maximum).
static constexpr unsigned kFull = 0xFFFFFFFFu;
__device__ __forceinline__ uint64_t sortable(double d)      // order-preserving map double -> u64 (no NaNs, no -0.0 here)
{
    const long long bits = __double_as_longlong(d);
    return (uint64_t)(bits ^ ((bits >> 63) | (long long)0x8000000000000000ull));
}
// What a lane in need of a greedy move publishes for the warp: its position, the packed line fields, its legal
// moves and what the opponent could play now (the starting point of every child's reply count).
struct __align__(16) GuidedSlot { uint4 cur, ev0, ev1, misc; };     // 64 bytes

// The greedy moves of ALL lanes in need, worked out by the whole warp: their children -- n_mine each, `total` together
// -- are dealt out 32 per round, one per lane, whoever they belong to, so the scoring (about 115 instructions per
// child) runs with full warps however few lanes need a move and however different their child counts are.  Children
// of one lane are consecutive and in ascending order; the arg-max of each lane's children inside a round is a
// two-stage REDUX over the lanes holding them (first maximum: highest key, lowest lane among equals), a decisive
// child counts as the largest key of all (first decisive child, else first best child -- mcts.py:362-383), and a
// lane whose children span two rounds keeps the earlier round's winner unless the later one is strictly better.
__device__ __forceinline__ int dealt_greedy_actions(bool need, const State4 &cur, const EvalPos &ev, int player, uint64_t mask,
                                                    uint32_t n_mine, GuidedSlot *slots, const double *tab)
{
    const unsigned lane = threadIdx.x & 31u;
    if (need) {
        uint32_t opp_lo, opp_hi;
        legal_masks(cur, 1 - player, opp_lo, opp_hi);
        GuidedSlot &g = slots[lane];
        g.cur = make_uint4(cur.x, cur.y, cur.z, cur.w);
        g.ev0 = make_uint4((uint32_t)ev.occ, (uint32_t)(ev.occ >> 32), (uint32_t)ev.np, (uint32_t)(ev.np >> 32));
        g.ev1 = make_uint4((uint32_t)ev.pres, (uint32_t)(ev.pres >> 32), (uint32_t)mask, (uint32_t)(mask >> 32));
        g.misc = make_uint4(opp_lo, opp_hi, (uint32_t)player, 0u);
    }
    uint32_t incl = n_mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(kFull, incl, o); if (lane >= (unsigned)o) incl += t; }
    const uint32_t total = __shfl_sync(kFull, incl, 31), start = incl - n_mine;
    __syncwarp();
    uint64_t best_key = 0;
    int best_a = -1;
    for (uint32_t base = 0; base < total; base += 32) {
        const uint32_t idx = base + lane;
        const bool have = idx < total;
        // owner of child idx: the first lane whose range ends beyond it = the number of lanes with incl <= idx
        int src = 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const uint32_t v = __shfl_sync(kFull, incl, src + o - 1); if (v <= idx) src += o; }
        const uint32_t first = __shfl_sync(kFull, start, src);
        const GuidedSlot &g = slots[src];
        const uint4 gc = g.cur, e0 = g.ev0, e1 = g.ev1, gm = g.misc;
        const State4 c = { gc.x, gc.y, gc.z, gc.w };
        const uint32_t lo = e1.z, hi = e1.w;
        const int pl = (int)gm.z;
        const int a = select_bit64(lo, hi, __popc(lo), have ? (int)(idx - first) : 0);
        EvalPos ce = { ((uint64_t)e0.y << 32) | e0.x, ((uint64_t)e0.w << 32) | e0.z, ((uint64_t)e1.y << 32) | e1.x };
        eval_pos_place(ce, a >> 4, a & 15);
        uint32_t clo, chi;
        child_reply_masks(gm.x, gm.y, a, clo, chi);
        uint64_t key = ~0ull;                                            // decisive: a fourth shape on a line, or no reply
        if (!((ce.np & (kLineLsb << 2)) != 0 || (clo | chi) == 0)) {
            int f[6];
            eval_features_pos(apply_action(c, pl, a), ce, pl, 1 - pl, __popc(clo) + __popc(chi), f);
            key = sortable(eval_dot_tab(f, tab));
        }
        const unsigned seg = __match_any_sync(kFull, have ? src : 64);
        const uint32_t kh = (uint32_t)(key >> 32), mh = __reduce_max_sync(seg, kh);
        const bool cand = kh == mh;
        const uint32_t kl = cand ? (uint32_t)key : 0u, ml = __reduce_max_sync(seg, kl);
        const int wl = __ffs((int)(__ballot_sync(kFull, cand && kl == ml) & seg)) - 1;
        const int round_a = __shfl_sync(kFull, a, wl);
        // back to the owners: any lane that holds one of my children knows my round's winner
        const bool mine = need && start < base + 32 && incl > base;
        const int at = mine ? (int)((start > base ? start : base) - base) : 0;
        const uint32_t rh = __shfl_sync(kFull, mh, at), rl = __shfl_sync(kFull, ml, at);
        const int ra = __shfl_sync(kFull, round_a, at);
        const uint64_t rk = ((uint64_t)rh << 32) | rl;
        if (mine && (best_a < 0 || rk > best_key)) { best_key = rk; best_a = ra; }
    }
    __syncwarp();
    return best_a;
}

__global__ void __launch_bounds__(kBlock) k_playout_guided(const uint4 *__restrict__ roots, uint64_t n_roots, uint32_t rollouts,
                                                            uint64_t seed, uint64_t first_rollout, uint32_t root_index_base,
                                                            int max_plies, double epsilon, const EvalWeights weights,
                                                            const GuidedLine *__restrict__ lines, int scoring,
                                                            uint32_t *wins_p0, uint32_t *wins_p1, uint32_t *draws,
                                                            int8_t *per_playout, uint8_t *per_plies)
{
    __shared__ double s_tab[6 * kEvalTabStride];
    __shared__ GuidedSlot s_slots[kBlock / 32][32];
#pragma unroll
    for (int i = 0; i < 6; ++i)
        for (int v = threadIdx.x; v < kEvalTabStride; v += kBlock) s_tab[i * kEvalTabStride + v] = __dmul_rn(weights.w[i], (double)(v - 64));
    __syncthreads();
    const uint64_t total = n_roots * rollouts;
    const unsigned lane = threadIdx.x & 31u;
    const uint64_t warp0 = (blockIdx.x * (uint64_t)kBlock + threadIdx.x) >> 5, n_warps = ((uint64_t)gridDim.x * kBlock) >> 5;
    for (uint64_t base = warp0 * 32; base < total; base += n_warps * 32) {
        const uint64_t id = base + lane;
        bool done = id >= total;
        const uint32_t root = done ? 0u : (uint32_t)(id / rollouts), j = done ? 0u : (uint32_t)(id % rollouts);
        const uint4 v = __ldg(roots + root);
        State4 cur = { v.x, v.y, v.z, v.w };
        EvalPos ev = eval_pos(cur);
        const GuidedLine *line = lines ? lines + root : nullptr;
        bool on_line = line != nullptr;
        const uint64_t rollout = first_rollout + j;
        int result = 0, plies = 0;
        Philox4 rnd = { { 0, 0, 0, 0 } };
        for (int ply = 0;; ++ply) {
            // ---- is the game over?  (mcts.py:412 cut-off, :414-419 winner, :427-429 blocked mover)
            int player = 0;
            uint64_t mask = 0;
            if (!done) {
                if (max_plies >= 0 && ply >= max_plies) { result = 0; done = true; }
                else if (on_line) {
                    if (ply == (int)line->len) { result = line->result; done = true; }
                    else player = (line->player + ply) & 1;
                } else {
                    const int ws = win_status(cur);
                    if (ws) { result = ws == 1 ? 1 : -1; done = true; }
                    else {
                        int status;
                        mask = movegen(cur, &player, &status);
                        if (mask == 0) { result = player == 0 ? -1 : 1; done = true; }
                    }
                }
                if (done) plies = ply;
            }
            if (!__any_sync(kFull, !done)) break;
            if ((ply & 1) == 0)
                rnd = philox4x32_10((uint32_t)rollout, (uint32_t)(rollout >> 32), 0x80000000u | (uint32_t)(ply >> 1), root_index_base + root,
                                    (uint32_t)seed, (uint32_t)(seed >> 32));
            // ---- the move: exploring pick, the line's move, or a greedy move of its own
            int action = 0;
            bool need = false;
            if (!done) {
                const uint32_t r_eps = (ply & 1) ? rnd.v[2] : rnd.v[0], r_pick = (ply & 1) ? rnd.v[3] : rnd.v[1];
                if (epsilon > 0.0 && (double)r_eps * (1.0 / 4294967296.0) < epsilon) {
                    if (on_line) { int status; mask = movegen(cur, &player, &status); }
                    const uint32_t lo = (uint32_t)mask, hi = (uint32_t)(mask >> 32);
                    action = select_bit64(lo, hi, __popc(lo), (int)__umulhi(r_pick, (uint32_t)(__popc(lo) + __popc(hi))));
                    on_line = on_line && action == (int)line->action[ply];
                } else if (on_line) action = line->action[ply];
                else need = true;
            }
            if (__any_sync(kFull, need)) {
                const uint32_t n_mine = need ? (uint32_t)(__popc((uint32_t)mask) + __popc((uint32_t)(mask >> 32))) : 0u;
                const uint32_t n_max = __reduce_max_sync(kFull, n_mine), n_all = __reduce_add_sync(kFull, n_mine);
                // dealt: one round of ~160 instructions per 32 children of the warp; each lane on its own: the longest
                // scan, ~115 instructions per child
                if (scoring == 2 || (scoring == 0 && ((n_all + 31u) >> 5) * 160u < n_max * 115u)) {
                    const int a = dealt_greedy_actions(need, cur, ev, player, mask, n_mine, s_slots[threadIdx.x >> 5], s_tab);
                    if (need) action = a;
                } else if (need) action = guided_greedy_action(cur, ev, player, mask, weights.w, s_tab);
            }
            if (!done) {
                cur = apply_action(cur, player, action);
                eval_pos_place(ev, action >> 4, action & 15);
            }
        }
        // tallies: the lanes of the warp that played for the same root (consecutive ids: almost always all of them)
        // are summed first, one lane per root issues the atomics
        const bool valid = id < total;
        const unsigned active = __ballot_sync(kFull, valid);
        if (valid) {
            const unsigned peers = __match_any_sync(active, root);
            const unsigned p0 = __ballot_sync(active, result > 0) & peers, p1 = __ballot_sync(active, result < 0) & peers;
            if ((unsigned)(__ffs((int)peers) - 1) == lane) {
                const uint32_t n0 = (uint32_t)__popc(p0), n1 = (uint32_t)__popc(p1), nd = (uint32_t)__popc(peers) - n0 - n1;
                if (wins_p0 && n0) atomicAdd(wins_p0 + root, n0);
                if (wins_p1 && n1) atomicAdd(wins_p1 + root, n1);
                if (draws && nd) atomicAdd(draws + root, nd);
            }
            if (per_playout) per_playout[id] = (int8_t)result;
            if (per_plies) per_plies[id] = (uint8_t)plies;
        }
    }
}

int launch_eval_features(const uint4 *states, const uint8_t *player, size_t n, double *features, Ctx *c, cudaStream_t s)
{
    k_eval_features<<<grid_for(n, kBlock, c->sms, 16), kBlock, 0, s>>>(states, player, n, features);
    QK_CUDA(cudaGetLastError());
    return QK_OK;
}

size_t guided_scratch_bytes(size_t n_roots) { return n_roots * sizeof(GuidedLine); }

int launch_playouts_guided(const uint4 *roots, size_t n_roots, uint32_t rollouts, uint64_t seed, uint64_t first_rollout,
                           uint32_t root_index_base, int max_plies, double epsilon, const double *weights, uint32_t *wins_p0,
                           uint32_t *wins_p1, uint32_t *draws, int8_t *per_playout, uint8_t *per_plies, void *line_scratch, Ctx *c,
                           cudaStream_t s)
{
    EvalWeights w;
    for (int i = 0; i < 6; ++i) w.w[i] = weights[i];
    const size_t total = n_roots * (size_t)rollouts;
    // the greedy line pays off as soon as a root is played a few times (it costs about one non-exploring playout);
    // with epsilon = 1 no playout ever follows it
    GuidedLine *lines = (line_scratch && rollouts >= 4 && epsilon < 1.0) ? (GuidedLine *)line_scratch : nullptr;
    if (lines) {
        k_guided_lines<<<(unsigned)((n_roots + kBlock - 1) / kBlock), kBlock, 0, s>>>(roots, n_roots, w, lines);
        QK_CUDA(cudaGetLastError());
    }
    k_playout_guided<<<grid_for(total, kBlock, c->sms, 16), kBlock, 0, s>>>(roots, n_roots, rollouts, seed, first_rollout,
                                                                            root_index_base, max_plies, epsilon, w, lines,
                                                                            (int)option(OPT_GUIDED_SCORING), wins_p0,
                                                                            wins_p1, draws, per_playout, per_plies);
    QK_CUDA(cudaGetLastError());
    return QK_OK;
}

}  // namespace qk
