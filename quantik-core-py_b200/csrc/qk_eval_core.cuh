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

// features (evaluation.py:130-194) of a VALID position from `player`'s perspective, as small integers
// in FEATURE_NAMES order: threat_own, threat_opp, threat_shared, mobility_diff, build_two, build_one.
QK_HD void eval_features(const State4 &s, int player, int f[6])
{
    const uint32_t line_mask[12] = { 0x000F, 0x00F0, 0x0F00, 0xF000, 0x1111, 0x2222, 0x4444, 0x8888,
                                     0x0033, 0x00CC, 0x3300, 0xCC00 };                      // commons.py:24-48
    const int side_to_move = piece_balance(s) == 0 ? 0 : 1;                                 // game_utils.py:226-231
    const int sign = side_to_move == player ? 1 : -1;
    uint32_t su[4], union_all = 0;
    for (int sh = 0; sh < 4; ++sh) { su[sh] = board_word(s, sh) | board_word(s, sh + 4); union_all |= su[sh]; }
    int threat_own = 0, threat_opp = 0, threat_shared = 0, build_two = 0, build_one = 0;
    for (int l = 0; l < 12; ++l) {
        const uint32_t mask = line_mask[l];
        int n_present = 0, missing = 0;
        for (int sh = 3; sh >= 0; --sh) {
            if (su[sh] & mask) ++n_present; else missing = sh;                              // first absent shape
        }
        const int occupied = popc(union_all & mask);
        if (n_present < occupied) continue;                                                 // dead line (:164-165)
        if (occupied == 3) {
            const uint32_t e = mask & ~union_all;
            int pos = 0;
            while (!((e >> pos) & 1u)) ++pos;
            const uint32_t scope = scope_of_square(pos);
            bool comp[2];
            for (int side = 0; side < 2; ++side)                                            // :170-175, :99-112
                comp[side] = popc(board_word(s, side * 4 + missing)) < 2 &&
                             (board_word(s, (1 - side) * 4 + missing) & scope) == 0;
            if (comp[player]) ++threat_own;
            if (comp[1 - player]) ++threat_opp;
            if (comp[0] && comp[1]) threat_shared += sign;
        } else if (occupied == 2) build_two += sign;
        else if (occupied == 1) build_one += sign;
    }
    uint32_t lo, hi;
    legal_masks(s, side_to_move, lo, hi);
    const int n = popc(lo) + popc(hi);                                                      // count_legal_moves (:115-127)
    f[0] = threat_own; f[1] = threat_opp; f[2] = threat_shared;
    f[3] = side_to_move == player ? n : -n;
    f[4] = build_two; f[5] = build_one;
}

// evaluate (:197-212): weights @ features, accumulated left to right in float64 without contraction
QK_HD double eval_score(const State4 &s, int player, const double w[6])
{
    int f[6];
    eval_features(s, player, f);
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

// One eval-guided playout (mcts.py:385-437 with _select_rollout_move :345-383).  Stream contract: ply p
// uses Philox block (rollout lo, rollout hi, 0x80000000 | p/2, root index); word 2*(p%2) is the epsilon
// draw (explore iff epsilon > 0 and word * 2^-32 < epsilon), word 2*(p%2)+1 the uniform pick.
QK_HD int guided_playout_one(const State4 &root, uint64_t seed, uint64_t rollout, uint32_t root_index, int max_plies,
                             double epsilon, const double w[6], int *plies_out)
{
    State4 cur = root;
    int ply = 0, result;
    Philox4 rnd = { { 0, 0, 0, 0 } };
    for (;;) {
        if (max_plies >= 0 && ply >= max_plies) { result = 0; break; }
        const int ws = win_status(cur);
        if (ws) { result = ws == 1 ? 1 : -1; break; }
        int player, status;
        const uint64_t mask = movegen(cur, &player, &status);
        if (mask == 0) { result = player == 0 ? -1 : 1; break; }
        if ((ply & 1) == 0)
            rnd = philox4x32_10((uint32_t)rollout, (uint32_t)(rollout >> 32), 0x80000000u | (uint32_t)(ply >> 1), root_index,
                                (uint32_t)seed, (uint32_t)(seed >> 32));
        const uint32_t r_eps = rnd.v[2 * (ply & 1)], r_pick = rnd.v[2 * (ply & 1) + 1];
        const uint32_t lo = (uint32_t)mask, hi = (uint32_t)(mask >> 32);
        int action = -1;
        if (epsilon > 0.0 && (double)r_eps * (1.0 / 4294967296.0) < epsilon) {
            const int n = popc(lo) + popc(hi);
            action = select_bit64(lo, hi, popc(lo), (int)mulhi(r_pick, (uint32_t)n));
        } else {
            double best = 0.0;
            for (uint64_t m = mask; m; m &= m - 1) {
                int a = 0;
                const uint32_t mlo = (uint32_t)m;
#ifdef __CUDA_ARCH__
                a = mlo ? __ffs((int)mlo) - 1 : 32 + __ffs((int)(uint32_t)(m >> 32)) - 1;
#else
                a = mlo ? __builtin_ctz(mlo) : 32 + __builtin_ctz((uint32_t)(m >> 32));
#endif
                const State4 child = apply_action(cur, player, a);
                uint32_t clo, chi;
                legal_masks(child, 1 - player, clo, chi);
                if (has_winning_line(child) || (clo | chi) == 0) { action = a; break; }      // decisive move: take it
                const double sc = eval_score(child, player, w);
                if (action < 0 || sc > best) { best = sc; action = a; }                      // first maximum wins
            }
        }
        cur = apply_action(cur, player, action);
        ++ply;
    }
    if (plies_out) *plies_out = ply;
    return result;
}

}  // namespace qk
