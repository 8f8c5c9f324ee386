// qk_eval_core.cuh -- the reference's handcrafted evaluation and its eval-guided rollout policy
// (SURVEY 8f-4): evaluation.features / evaluate (evaluation.py:130-212) and
// MCTSEngine._select_rollout_move (mcts.py:345-383), on the register-resident State4.
// __host__ __device__ like qk_bits.cuh, so tests/hostsim can run it on the CPU.
#pragma once
#include "qk_bits.cuh"
#include "qk_playout_core.cuh"

namespace qk {

QK_HD uint32_t board_word(const State4 &s, int i)      // word i of the <8H payload
{
    const uint32_t r = i < 2 ? s.x : (i < 4 ? s.y : (i < 6 ? s.z : s.w));
    return (i & 1) ? (r >> 16) : (r & 0xFFFFu);
}
QK_HD uint32_t scope_of_square(int p)
{
    return ((0xFu << (p & 12)) | (0x1111u << (p & 3)) | (0x33u << (p & 10))) & 0xFFFFu;
}

// ---------------------------------------------------------------- the 12 lines at once
// The reference walks the 12 lines one after the other (evaluation.py:147-189).  Here every line owns a 4-bit field
// of a 64-bit word (rows 0-3 -> fields 0-3, columns -> 4-7, zones -> 8-11, the order of commons.py:24-48), and a
// position carries, per line, the number of pieces on it (occ), the number of DISTINCT shapes on it (np) and which
// shapes those are (pres: bit `shape` of the line's field) -- all updated by one table word when a piece lands.  "Dead line" (a repeated shape,
// :164-165) is np < occ, occupied == 1/2/3 are bit patterns of the occ field, so build_one / build_two are two
// popcounts and only lines with three distinct shapes -- the threats, rarely more than two -- are looked at one by one.
struct EvalPos { uint64_t occ, np, pres; };
static constexpr uint64_t kLineLsb = 0x0000111111111111ull;

QK_HD uint64_t line_fields(int p)          // 1 in the fields of the three lines through square p
{
    const uint32_t lo = (1u << (p & 12)) | (0x10000u << ((p & 3) << 2));
    const uint32_t hi = 1u << ((p & 8) | ((p & 2) << 1));
    return ((uint64_t)hi << 32) | lo;
}
QK_HD void eval_pos_place(EvalPos &e, int shape, int p)
{
    const uint64_t t = line_fields(p);
    e.occ += t;
    e.np += t & ~(e.pres >> shape);
    e.pres |= t << shape;
}
QK_HD EvalPos eval_pos(const State4 &s)
{
    EvalPos e = { 0, 0, 0 };
    for (int i = 0; i < 8; ++i)
        for (uint32_t w = board_word(s, i); w; w &= w - 1) {
            int p = 0;
            while (!((w >> p) & 1u)) ++p;
            eval_pos_place(e, i & 3, p);
        }
    return e;
}
QK_HD int popc64(uint64_t v) { return popc((uint32_t)v) + popc((uint32_t)(v >> 32)); }
QK_HD uint32_t line_squares(int l)         // commons.py:24-48
{
    return l < 4 ? 0xFu << (4 * l) : (l < 8 ? 0x1111u << (l - 4) : 0x33u << ((((l - 8) >> 1) << 3) | (((l - 8) & 1) << 1)));
}

// features (evaluation.py:130-194) of a VALID position from `player`'s perspective, as small integers in
// FEATURE_NAMES order: threat_own, threat_opp, threat_shared, mobility_diff, build_two, build_one.
// n_legal = count_legal_moves of the side to move (:115-127).
QK_HD void eval_features_pos(const State4 &s, const EvalPos &e, int player, int side_to_move, int n_legal, int f[6])
{
    const int sign = side_to_move == player ? 1 : -1;
    const uint64_t diff = e.occ - e.np;                               // per field, 0..3, never a borrow
    const uint64_t live = ~(diff | (diff >> 1)) & kLineLsb;           // no repeated shape (:164-165)
    const uint64_t b0 = e.occ, b1 = e.occ >> 1;
    const int build_one = popc64(b0 & ~b1 & live), build_two = popc64(b1 & ~b0 & live);
    int threat_own = 0, threat_opp = 0, threat_shared = 0;
    uint64_t threats = b0 & b1 & live;                                // three pieces, three shapes
    if (threats) {
        const uint32_t all = (s.x | s.y | s.z | s.w), union_all = (all | (all >> 16)) & 0xFFFFu;
        while (threats) {
            int bit = 0;
            const uint32_t tlo = (uint32_t)threats;
#ifdef __CUDA_ARCH__
            bit = tlo ? __ffs((int)tlo) - 1 : 32 + __ffs((int)(uint32_t)(threats >> 32)) - 1;
#else
            bit = tlo ? __builtin_ctz(tlo) : 32 + __builtin_ctz((uint32_t)(threats >> 32));
#endif
            threats &= threats - 1;
            const int l = bit >> 2;
            const uint32_t absent = ~(uint32_t)(e.pres >> bit) & 0xFu;        // exactly one bit: the missing shape
            const int missing = (int)(absent >> 1) - (int)(absent >> 3);      // 1,2,4,8 -> 0,1,2,3
            const uint32_t empty = line_squares(l) & ~union_all;      // exactly one square
            int pos = 0;
#ifdef __CUDA_ARCH__
            pos = __ffs((int)empty) - 1;
#else
            pos = __builtin_ctz(empty);
#endif
            const uint32_t scope = scope_of_square(pos);
            const uint32_t w0 = board_word(s, missing), w1 = board_word(s, 4 + missing);
            const bool comp0 = popc(w0) < 2 && (w1 & scope) == 0;     // :170-175, :99-112
            const bool comp1 = popc(w1) < 2 && (w0 & scope) == 0;
            threat_own += (player ? comp1 : comp0) ? 1 : 0;
            threat_opp += (player ? comp0 : comp1) ? 1 : 0;
            if (comp0 && comp1) threat_shared += sign;
        }
    }
    f[0] = threat_own; f[1] = threat_opp; f[2] = threat_shared;
    f[3] = side_to_move == player ? n_legal : -n_legal;
    f[4] = sign * build_two; f[5] = sign * build_one;
}

QK_HD void eval_features(const State4 &s, int player, int f[6])
{
    const int side_to_move = piece_balance(s) == 0 ? 0 : 1;                                 // game_utils.py:226-231
    uint32_t lo, hi;
    legal_masks(s, side_to_move, lo, hi);
    eval_features_pos(s, eval_pos(s), player, side_to_move, popc(lo) + popc(hi), f);
}

// the same sum with the six products looked up: tab[i * kEvalTabStride + 64 + v] = w[i] * v for |v| <= 64 (every
// feature is a count of lines or moves), computed once per kernel with the same rounding.  Saves the six
// int -> float64 conversions and multiplications per position, the slowest instructions of the evaluation.
static constexpr int kEvalTabStride = 132;
QK_HD double eval_dot_tab(const int f[6], const double *tab)
{
    double acc = 0.0;
    for (int i = 0; i < 6; ++i) {
#ifdef __CUDA_ARCH__
        acc = __dadd_rn(acc, tab[i * kEvalTabStride + 64 + f[i]]);
#else
        acc += tab[i * kEvalTabStride + 64 + f[i]];
#endif
    }
    return acc;
}

QK_HD double eval_dot(const int f[6], const double w[6])
{
    double acc = 0.0;
    for (int i = 0; i < 6; ++i) {
#ifdef __CUDA_ARCH__
        acc = __dadd_rn(acc, __dmul_rn(w[i], (double)f[i]));
#else
        acc += w[i] * (double)f[i];
#endif
    }
    return acc;
}

// evaluate (:197-212): weights @ features, accumulated left to right in float64 without contraction
QK_HD double eval_score(const State4 &s, int player, const double w[6])
{
    int f[6];
    eval_features(s, player, f);
    return eval_dot(f, w);
}

// The greedy move of a position (the non-exploring branch of _select_rollout_move, mcts.py:362-383): the first
// decisive child, else the first child with the best evaluation.  -> action; cur must have a legal move.
// what the player NOT to move could play on `cur` -- the starting point of every child's reply count: after the
// mover puts shape s on square p the opponent loses p for every shape and the scope of p for shape s, nothing else
// (the opponent's own pieces, hence its exhausted shapes, are unchanged).
QK_HD void child_reply_masks(uint32_t opp_lo, uint32_t opp_hi, int action, uint32_t &clo, uint32_t &chi)
{
    const int p = action & 15, s = action >> 4;
    const uint32_t bit2 = 0x00010001u << p, closed = scope_of_square(p) << ((s & 1) << 4);
    clo = opp_lo & ~bit2 & ~(s < 2 ? closed : 0u);
    chi = opp_hi & ~bit2 & ~(s < 2 ? 0u : closed);
}

// ---- a child's features from its PARENT's data (k_playout_guided scores the children of a position without
// building them).  A threat line -- three pieces, three shapes -- has one empty square and one missing shape, and a
// player "can complete" it (evaluation.py:99-112, :170-175) exactly when that placement is among his legal moves;
// the legal moves of both players in the child follow from the parent's:
//   the mover's:    what he could play in the parent, minus the square just taken, minus the whole shape if this was
//                   its second piece (`once` = his shapes with exactly one piece placed);
//   the opponent's: child_reply_masks.
QK_HD uint32_t placed_once(const State4 &s, int player)
{
    const uint32_t m01 = player ? s.z : s.x, m23 = player ? s.w : s.y;
    return (popc(m01 & 0xFFFFu) == 1 ? 1u : 0u) | (popc(m01 >> 16) == 1 ? 2u : 0u) |
           (popc(m23 & 0xFFFFu) == 1 ? 4u : 0u) | (popc(m23 >> 16) == 1 ? 8u : 0u);
}
QK_HD uint64_t mover_options_in_child(uint64_t parent_mask, uint32_t once, int action)
{
    const int shape = action >> 4;
    uint64_t own = parent_mask & ~(0x0001000100010001ull << (action & 15));
    if ((once >> shape) & 1u) own &= ~(0xFFFFull << (shape << 4));
    return own;
}
// features from the MOVER's point of view (the side to move in the child is the opponent: sign = -1 in
// evaluation.py:130-194); ce = the child's line fields, own64 / opp64 = 64-bit action masks of what the mover /
// the opponent may play in the child, line_sq[l] = the squares of line l (line_squares)
QK_HD void eval_features_child(const EvalPos &ce, uint64_t own64, uint64_t opp64, const uint32_t *line_sq, int f[6])
{
    const uint64_t diff = ce.occ - ce.np;
    const uint64_t live = ~(diff | (diff >> 1)) & kLineLsb;
    const uint64_t b0 = ce.occ, b1 = ce.occ >> 1;
    f[5] = -popc64(b0 & ~b1 & live); f[4] = -popc64(b1 & ~b0 & live);
    int t_own = 0, t_opp = 0, t_both = 0;
    for (uint64_t threats = b0 & b1 & live; threats; threats &= threats - 1) {
        const uint32_t tlo = (uint32_t)threats;
        int bit;
#ifdef __CUDA_ARCH__
        bit = tlo ? __ffs((int)tlo) - 1 : 32 + __ffs((int)(uint32_t)(threats >> 32)) - 1;
#else
        bit = tlo ? __builtin_ctz(tlo) : 32 + __builtin_ctz((uint32_t)(threats >> 32));
#endif
        const uint32_t absent = ~(uint32_t)(ce.pres >> bit) & 0xFu;          // exactly one bit: the missing shape
        const int lane16 = (int)(((absent >> 1) - (absent >> 3)) << 4);       // 1,2,4,8 -> 0,16,32,48
        const uint32_t line = line_sq[bit >> 2];
        const uint32_t o = ((uint32_t)(own64 >> lane16) & line) ? 1u : 0u, q = ((uint32_t)(opp64 >> lane16) & line) ? 1u : 0u;
        t_own += (int)o; t_opp += (int)q; t_both += (int)(o & q);
    }
    f[0] = t_own; f[1] = t_opp; f[2] = -t_both; f[3] = -popc64(opp64);
}

QK_HD int guided_greedy_action(const State4 &cur, const EvalPos &ev, int player, uint64_t mask, const double w[6],
                               const double *tab = nullptr)
{
    int action = -1;
    double best = 0.0;
    uint32_t opp_lo, opp_hi;
    legal_masks(cur, 1 - player, opp_lo, opp_hi);
    for (uint64_t m = mask; m; m &= m - 1) {
        int a = 0;
        const uint32_t mlo = (uint32_t)m;
#ifdef __CUDA_ARCH__
        a = mlo ? __ffs((int)mlo) - 1 : 32 + __ffs((int)(uint32_t)(m >> 32)) - 1;
#else
        a = mlo ? __builtin_ctz(mlo) : 32 + __builtin_ctz((uint32_t)(m >> 32));
#endif
        EvalPos ce = ev;
        eval_pos_place(ce, a >> 4, a & 15);
        uint32_t clo, chi;
        child_reply_masks(opp_lo, opp_hi, a, clo, chi);
        // decisive move: a line with four shapes, or the opponent cannot answer -- take it
        if ((ce.np & (kLineLsb << 2)) != 0 || (clo | chi) == 0) return a;
        int f[6];
        eval_features_pos(apply_action(cur, player, a), ce, player, 1 - player, popc(clo) + popc(chi), f);
        const double sc = tab ? eval_dot_tab(f, tab) : eval_dot(f, w);
        if (action < 0 || sc > best) { best = sc; action = a; }                              // first maximum wins
    }
    return action;
}

// The greedy line of a root: what a playout does as long as it never explores.  Every playout of the root follows
// it up to its first exploring ply, so its moves are worked out ONCE per root (k_guided_lines) instead of once per
// playout -- from the empty board that is the 64-child scan of ply 0 for every playout, the 49-child scan of ply 1
// for 80 % of them, ...  len = greedy moves until the game ends; result = who wins there (+1 P0 / -1 P1).
struct __attribute__((aligned(32))) GuidedLine { uint8_t action[16]; uint8_t len; int8_t result; uint8_t player; uint8_t pad[13]; };

QK_HD GuidedLine guided_line(const State4 &root, const double w[6])
{
    GuidedLine g;
    for (int i = 0; i < 16; ++i) g.action[i] = 0;
    g.len = 0; g.result = 0; g.player = 0;
    State4 cur = root;
    EvalPos ev = eval_pos(root);
    for (int ply = 0;; ++ply) {
        const int ws = win_status(cur);
        if (ws) { g.result = (int8_t)(ws == 1 ? 1 : -1); break; }
        int player, status;
        const uint64_t mask = movegen(cur, &player, &status);
        if (ply == 0) g.player = (uint8_t)player;
        if (mask == 0) { g.result = (int8_t)(player == 0 ? -1 : 1); break; }
        const int a = guided_greedy_action(cur, ev, player, mask, w);
        g.action[ply] = (uint8_t)a;
        g.len = (uint8_t)(ply + 1);
        cur = apply_action(cur, player, a);
        eval_pos_place(ev, a >> 4, a & 15);
    }
    return g;
}

// One eval-guided playout (mcts.py:385-437 with _select_rollout_move :345-383).  Stream contract: ply p
// uses Philox block (rollout lo, rollout hi, 0x80000000 | p/2, root index); word 2*(p%2) is the epsilon
// draw (explore iff epsilon > 0 and word * 2^-32 < epsilon), word 2*(p%2)+1 the uniform pick.
QK_HD int guided_playout_one(const State4 &root, uint64_t seed, uint64_t rollout, uint32_t root_index, int max_plies,
                             double epsilon, const double w[6], int *plies_out, const GuidedLine *line = nullptr)
{
    State4 cur = root;
    EvalPos ev = eval_pos(root);          // carried along: one table word per move instead of a rescan of the board
    int ply = 0, result;
    bool on_line = line != nullptr;       // no exploring ply so far: the position is the line's
    Philox4 rnd = { { 0, 0, 0, 0 } };
    for (;;) {
        if (max_plies >= 0 && ply >= max_plies) { result = 0; break; }
        int player = 0;
        uint64_t mask = 0;
        if (on_line) {
            if (ply == (int)line->len) { result = line->result; break; }        // the line's own end: a win or a blocked mover
            player = (line->player + ply) & 1;
        } else {
            const int ws = win_status(cur);
            if (ws) { result = ws == 1 ? 1 : -1; break; }
            int status;
            mask = movegen(cur, &player, &status);
            if (mask == 0) { result = player == 0 ? -1 : 1; break; }
        }
        if ((ply & 1) == 0)
            rnd = philox4x32_10((uint32_t)rollout, (uint32_t)(rollout >> 32), 0x80000000u | (uint32_t)(ply >> 1), root_index,
                                (uint32_t)seed, (uint32_t)(seed >> 32));
        const uint32_t r_eps = (ply & 1) ? rnd.v[2] : rnd.v[0], r_pick = (ply & 1) ? rnd.v[3] : rnd.v[1];
        int action = -1;
        if (epsilon > 0.0 && (double)r_eps * (1.0 / 4294967296.0) < epsilon) {
            if (on_line) { int status; mask = movegen(cur, &player, &status); }
            const uint32_t lo = (uint32_t)mask, hi = (uint32_t)(mask >> 32);
            const int n = popc(lo) + popc(hi);
            action = select_bit64(lo, hi, popc(lo), (int)mulhi(r_pick, (uint32_t)n));
            on_line = on_line && action == (int)line->action[ply];              // an exploring ply may still pick the line's move
        } else if (on_line) {
            action = line->action[ply];
        } else {
            action = guided_greedy_action(cur, ev, player, mask, w);
        }
        cur = apply_action(cur, player, action);
        eval_pos_place(ev, action >> 4, action & 15);
        ++ply;
    }
    if (plies_out) *plies_out = ply;
    return result;
}

}  // namespace qk
