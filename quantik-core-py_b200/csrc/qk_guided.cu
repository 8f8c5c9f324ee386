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
// its own has to score every child of its position.  The warp scores the children of all such lanes together, dealt
// out one per lane (dealt_greedy_actions; a lane-serial scan keeps the other lanes waiting for the longest one and
// is kept as the second implementation the parity tests compare with: qk_set_option("guided_scoring", 1)).  Same
// moves either way (first decisive child, else first maximum).
static constexpr unsigned kFull = 0xFFFFFFFFu;
__device__ __forceinline__ uint64_t sortable(double d)      // order-preserving map double -> u64 (no NaNs, no -0.0 here)
{
    const long long bits = __double_as_longlong(d);
    return (uint64_t)(bits ^ ((bits >> 63) | (long long)0x8000000000000000ull));
}
// What a lane in need of a greedy move publishes for the warp: the packed line fields of its position, its legal
// moves, what the opponent could play now (the starting point of every child's reply count), and in `info` the
// occupied squares (bits 0-15), the mover's shapes with exactly one piece placed (16-19) and the mover (20).
struct __align__(16) GuidedSlot { uint32_t occ_lo, occ_hi, np_lo, np_hi, pres_lo, pres_hi, mask_lo, mask_hi, opp_lo, opp_hi, info, pad; };
struct GuidedWarp {
    GuidedSlot slot[32];
    uint64_t best_key[32];              // per owner lane: the best child so far, 0 = none yet (a key is never 0)
    unsigned short queue[2048];         // (owner lane << 6 | action) of every child to score, owner by owner, ascending actions
    unsigned char best_a[32];
};

// The greedy moves of ALL lanes in need, worked out by the whole warp: every such lane lists its children in the
// warp's queue, and the queue is scored 32 children per round, one per lane, whoever they belong to -- full warps
// however few lanes need a move and however different their child counts are.  Children of one lane are consecutive
// and in ascending order; the arg-max of each lane's children inside a round is a two-stage REDUX over the lanes
// holding them (first maximum: highest key, lowest lane among equals), a decisive child counts as the largest key
// of all (first decisive child, else first best child -- mcts.py:362-383), and across rounds an owner keeps the
// earlier winner unless a later one is strictly better.
__device__ __forceinline__ int dealt_greedy_actions(bool need, const State4 &cur, const EvalPos &ev, int player, uint64_t mask,
                                                    uint32_t n_mine, GuidedWarp &w, const uint32_t *s_line, const uint4 *s_sq, const double *tab)
{
    const unsigned lane = threadIdx.x & 31u;
    uint32_t incl = n_mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(kFull, incl, o); if (lane >= (unsigned)o) incl += t; }
    const uint32_t total = __shfl_sync(kFull, incl, 31);
    if (need) {
        uint32_t opp_lo, opp_hi;
        legal_masks(cur, 1 - player, opp_lo, opp_hi);
        const uint32_t all = cur.x | cur.y | cur.z | cur.w, once = placed_once(cur, player);
        GuidedSlot &g = w.slot[lane];
        g.occ_lo = (uint32_t)ev.occ; g.occ_hi = (uint32_t)(ev.occ >> 32); g.np_lo = (uint32_t)ev.np; g.np_hi = (uint32_t)(ev.np >> 32);
        g.pres_lo = (uint32_t)ev.pres; g.pres_hi = (uint32_t)(ev.pres >> 32); g.mask_lo = (uint32_t)mask; g.mask_hi = (uint32_t)(mask >> 32);
        g.opp_lo = opp_lo; g.opp_hi = opp_hi;
        g.info = ((all | (all >> 16)) & 0xFFFFu) | (once << 16) | ((uint32_t)player << 20);
        w.best_key[lane] = 0ull;
        uint32_t at = incl - n_mine;
        for (uint32_t v = (uint32_t)mask; v; v &= v - 1) w.queue[at++] = (unsigned short)((lane << 6) | (uint32_t)(__ffs((int)v) - 1));
        for (uint32_t v = (uint32_t)(mask >> 32); v; v &= v - 1) w.queue[at++] = (unsigned short)((lane << 6) | (uint32_t)(32 + __ffs((int)v) - 1));
    }
    __syncwarp();
    for (uint32_t base = 0; base < total; base += 32) {
        const uint32_t idx = base + lane;
        const bool have = idx < total;
        const uint32_t e = have ? w.queue[idx] : 0u;
        const int src = (int)(e >> 6), a = (int)(e & 63u), shape = a >> 4, sq = a & 15;
        const uint4 g0 = *reinterpret_cast<const uint4 *>(&w.slot[src].occ_lo), g1 = *reinterpret_cast<const uint4 *>(&w.slot[src].pres_lo),
                    g2 = *reinterpret_cast<const uint4 *>(&w.slot[src].opp_lo);
        // per-square table: the fields of the three lines through the square (line_fields), its scope, its bit in both lanes
        const uint4 q = s_sq[sq];
        const uint64_t lf = ((uint64_t)q.y << 32) | q.x;
        EvalPos ce = { ((uint64_t)g0.y << 32) | g0.x, ((uint64_t)g0.w << 32) | g0.z, ((uint64_t)g1.y << 32) | g1.x };
        ce.occ += lf; ce.np += lf & ~(ce.pres >> shape); ce.pres |= lf << shape;             // eval_pos_place
        // what the opponent may answer (child_reply_masks): not the square, not the shape inside the square's scope
        const uint32_t closed = q.z << ((shape & 1) << 4);
        const uint32_t clo = g2.x & ~(q.w | (shape < 2 ? closed : 0u)), chi = g2.y & ~(q.w | (shape < 2 ? 0u : closed));
        uint64_t key = have ? ~0ull : 0ull;                              // decisive: a fourth shape on a line, or no reply
        if (have && !((ce.np & (kLineLsb << 2)) != 0 || (clo | chi) == 0)) {
            // scored from the parent's data (qk_eval_core.cuh: eval_features_child); the child is never built.
            // The mover's own options (mover_options_in_child): the parent's without the square, and without the
            // shape if this was its second piece
            const uint32_t gone = ((g2.z >> (16 + shape)) & 1u) ? 0xFFFFu << ((shape & 1) << 4) : 0u;
            const uint32_t olo = g1.z & ~(q.w | (shape < 2 ? gone : 0u)), ohi = g1.w & ~(q.w | (shape < 2 ? 0u : gone));
            int f[6];
            eval_features_child(ce, ((uint64_t)ohi << 32) | olo, ((uint64_t)chi << 32) | clo, s_line, f);
            key = sortable(eval_dot_tab(f, tab));
        }
        const unsigned seg = __match_any_sync(kFull, have ? src : 64);
        const uint32_t kh = (uint32_t)(key >> 32), mh = __reduce_max_sync(seg, kh);
        const bool cand = kh == mh;
        const uint32_t kl = cand ? (uint32_t)key : 0u, ml = __reduce_max_sync(seg, kl);
        const int wl = __ffs((int)(__ballot_sync(kFull, cand && kl == ml) & seg)) - 1;
        if (have && (int)lane == wl && key > w.best_key[src]) { w.best_key[src] = key; w.best_a[src] = (unsigned char)a; }
        __syncwarp();
    }
    return need ? (int)w.best_a[lane] : 0;
}

__global__ void __launch_bounds__(kBlock) k_playout_guided(const uint4 *__restrict__ roots, uint64_t n_roots, uint32_t rollouts,
                                                            uint64_t seed, uint64_t first_rollout, uint32_t root_index_base,
                                                            int max_plies, double epsilon, const EvalWeights weights,
                                                            const GuidedLine *__restrict__ lines, int scoring,
                                                            uint32_t *wins_p0, uint32_t *wins_p1, uint32_t *draws,
                                                            int8_t *per_playout, uint8_t *per_plies)
{
    __shared__ double s_tab[6 * kEvalTabStride];
    __shared__ GuidedWarp s_warp[kBlock / 32];
    __shared__ uint32_t s_line[12];
    __shared__ uint4 s_sq[16];
    if (threadIdx.x < 12) s_line[threadIdx.x] = line_squares((int)threadIdx.x);
    if (threadIdx.x < 16) {
        const uint64_t lf = line_fields((int)threadIdx.x);
        s_sq[threadIdx.x] = make_uint4((uint32_t)lf, (uint32_t)(lf >> 32), scope_of_square((int)threadIdx.x), 0x00010001u << threadIdx.x);
    }
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
        int p0 = on_line ? (int)line->player : -1;      // the mover at the root, once known (the line of a finished or invalid root has no moves)
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
                } else if (p0 < 0) {
                    // first look at a position this playout did not reach by its own moves: the full rules
                    const int ws = win_status(cur);
                    if (ws) { result = ws == 1 ? 1 : -1; done = true; }
                    else {
                        int status;
                        mask = movegen(cur, &player, &status);
                        p0 = (player - ply) & 1;
                        if (mask == 0) { result = player == 0 ? -1 : 1; done = true; }
                    }
                } else {
                    // reached by legal moves from a valid position: it is valid, the players alternate, and a completed
                    // line shows as a fourth shape in the line fields the playout carries (the last mover made it)
                    player = (p0 + ply) & 1;
                    if (ev.np & (kLineLsb << 2)) { result = player ? 1 : -1; done = true; }
                    else {
                        uint32_t lo, hi;
                        legal_masks(cur, player, lo, hi);
                        mask = ((uint64_t)hi << 32) | lo;
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
                    if (on_line) { uint32_t lo, hi; legal_masks(cur, player, lo, hi); mask = ((uint64_t)hi << 32) | lo; }
                    const uint32_t lo = (uint32_t)mask, hi = (uint32_t)(mask >> 32);
                    action = select_bit64(lo, hi, __popc(lo), (int)__umulhi(r_pick, (uint32_t)(__popc(lo) + __popc(hi))));
                    on_line = on_line && action == (int)line->action[ply];
                } else if (on_line) action = line->action[ply];
                else need = true;
            }
            if (__any_sync(kFull, need)) {
                const uint32_t n_mine = need ? (uint32_t)(__popc((uint32_t)mask) + __popc((uint32_t)(mask >> 32))) : 0u;
                // dealt out over the warp (measured faster for every mix of lanes in need, 7.0e8 against 6.5e8 playouts/s
                // with a per-ply choice and 3.3e8 with every lane scanning its own children: the dealt child is also
                // the cheaper one to score); the lane-serial scan is kept for the parity tests
                if (scoring != 1) {
                    const int a = dealt_greedy_actions(need, cur, ev, player, mask, n_mine, s_warp[threadIdx.x >> 5], s_line, s_sq, s_tab);
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
