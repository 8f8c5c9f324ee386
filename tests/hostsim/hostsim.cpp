// hostsim.cpp -- compiles the kernels' __host__ __device__ rule code (qk_bits.cuh,
// qk_playout_core.cuh) for the CPU so the exact bit tricks can be checked against the
// oracle in the CPU-only test tier.  TEST INFRASTRUCTURE; never part of the product.
#include <stdint.h>
#include <stddef.h>
#include "qk_bits.cuh"
#ifdef QK_HOSTSIM_PLAYOUT
#include "qk_playout_core.cuh"
#endif
#include "qk_perft_core.cuh"
#include "qk_eval_core.cuh"

using namespace qk;
static State4 ld(const uint16_t *p) { State4 s = { (uint32_t)p[0] | ((uint32_t)p[1] << 16), (uint32_t)p[2] | ((uint32_t)p[3] << 16),
                                                   (uint32_t)p[4] | ((uint32_t)p[5] << 16), (uint32_t)p[6] | ((uint32_t)p[7] << 16) }; return s; }
static void st(uint16_t *p, const State4 &s) { p[0] = s.x; p[1] = s.x >> 16; p[2] = s.y; p[3] = s.y >> 16; p[4] = s.z; p[5] = s.z >> 16; p[6] = s.w; p[7] = s.w >> 16; }

extern "C" {
void hs_movegen(const uint16_t *s, size_t n, uint64_t *mask, uint8_t *to_move, uint8_t *status)
{ for (size_t i = 0; i < n; ++i) { int p, c; mask[i] = movegen(ld(s + 8 * i), &p, &c); to_move[i] = p; status[i] = c; } }
void hs_win(const uint16_t *s, size_t n, uint8_t *w) { for (size_t i = 0; i < n; ++i) w[i] = win_status(ld(s + 8 * i)); }
void hs_canonical(const uint16_t *s, size_t n, int cs, uint16_t *out, uint32_t *tr)
{
    for (size_t i = 0; i < n; ++i) {
        State4 a = ld(s + 8 * i), c = cs ? canonical<true>(a) : canonical<false>(a);
        st(out + 8 * i, c);
        if (tr) { State4 c2 = canonical_with_transform(a, cs != 0, tr + i); if (c2.x != c.x || c2.y != c.y || c2.z != c.z || c2.w != c.w) tr[i] = 0xFFFFFFFFu; }
    }
}
void hs_orbit(const uint16_t *s, size_t n, uint8_t *o) { for (size_t i = 0; i < n; ++i) o[i] = orbit_size(ld(s + 8 * i)); }
void hs_apply_symmetry(const uint16_t *s, const uint32_t *tr, size_t n, uint16_t *out) { for (size_t i = 0; i < n; ++i) st(out + 8 * i, apply_symmetry(ld(s + 8 * i), tr[i])); }
uint16_t hs_d4(uint16_t v, int d4) { return (uint16_t)d4_apply((uint32_t)v, d4); }
void hs_apply(const uint16_t *s, const uint8_t *a, size_t n, uint16_t *out)
{ for (size_t i = 0; i < n; ++i) { State4 x = ld(s + 8 * i); int p = 0; if (validate(x, &p) != 0) p = 0; st(out + 8 * i, apply_action(x, p, a[i])); } }
void hs_leaf_parent(const uint16_t *s, size_t n, uint8_t *mover, uint8_t *moves, uint8_t *wins)
{ for (size_t i = 0; i < n; ++i) { int m, w; mover[i] = leaf_parent_counts(ld(s + 8 * i), m, w); moves[i] = m; wins[i] = w; } }
// every non-winning child of every position, evaluated from the parent's data (the perft tile kernel's way) and
// directly; -> children checked, mismatches
void hs_leaf_from_parent(const uint16_t *s, size_t n, uint64_t *checked, uint64_t *mismatch)
{
    TileLut lut[16];
    for (int p = 0; p < 16; ++p) lut[p] = tile_lut(p);
    *checked = 0; *mismatch = 0;
    for (size_t i = 0; i < n; ++i) {
        const State4 node = ld(s + 8 * i);
        const int mover = piece_balance(node);
        uint32_t lo, hi, wab, wcd, ylo, yhi;
        legal_masks(node, mover, lo, hi);
        const Presence pres = line_presence(node);
        winning_from_presence(pres, wab, wcd);
        legal_masks(node, mover ^ 1, ylo, yhi);
        const uint64_t kids = ((uint64_t)(hi & ~wcd) << 32) | (lo & ~wab);
        for (int a = 0; a < 64; ++a) {
            if (!((kids >> a) & 1)) continue;
            int m0, w0, m1, w1;
            leaf_parent_counts_for(apply_action(node, mover, a), mover ^ 1, m0, w0);
            leaf_counts_from_parent(ylo, yhi, pres, a, lut, m1, w1);
            // the shape-specialised form with the other word pair's masks hoisted (k_perft_tile's own lists)
            int m2 = -1, w2 = -1;
            switch (a >> 4) {
            case 0: ShapeLoop<0>(ylo, yhi, pres).child(lut[a & 15], m2, w2); break;
            case 1: ShapeLoop<1>(ylo, yhi, pres).child(lut[a & 15], m2, w2); break;
            case 2: ShapeLoop<2>(ylo, yhi, pres).child(lut[a & 15], m2, w2); break;
            default: ShapeLoop<3>(ylo, yhi, pres).child(lut[a & 15], m2, w2); break;
            }
            ++*checked;
            if (m0 != m1 || w0 != w1 || m0 != m2 || w0 != w2) ++*mismatch;
        }
    }
}
// every child of every position (valid, not won): the features the guided playout kernel derives from the PARENT's
// data against the evaluation of the child itself from its mover's point of view; -> children checked, mismatches
void hs_child_features(const uint16_t *s, size_t n, uint64_t *checked, uint64_t *mismatch)
{
    uint32_t line_sq[12];
    for (int l = 0; l < 12; ++l) line_sq[l] = line_squares(l);
    *checked = 0; *mismatch = 0;
    for (size_t i = 0; i < n; ++i) {
        const State4 node = ld(s + 8 * i);
        int player, status;
        const uint64_t mask = movegen(node, &player, &status);
        if (status != 0 || win_status(node) != 0) continue;
        const EvalPos ev = eval_pos(node);
        uint32_t opp_lo, opp_hi;
        legal_masks(node, 1 - player, opp_lo, opp_hi);
        const uint32_t once = placed_once(node, player);
        for (int a = 0; a < 64; ++a) {
            if (!((mask >> a) & 1)) continue;
            EvalPos ce = ev;
            eval_pos_place(ce, a >> 4, a & 15);
            uint32_t clo, chi;
            child_reply_masks(opp_lo, opp_hi, a, clo, chi);
            int f0[6], f1[6];
            eval_features_child(ce, mover_options_in_child(mask, once, a), ((uint64_t)chi << 32) | clo, line_sq, f0);
            eval_features(apply_action(node, player, a), player, f1);
            ++*checked;
            for (int k = 0; k < 6; ++k) if (f0[k] != f1[k]) { ++*mismatch; break; }
        }
    }
}
void hs_features(const uint16_t *s, const uint8_t *pl, size_t n, double *out)
{ for (size_t i = 0; i < n; ++i) { int f[6]; eval_features(ld(s + 8 * i), pl[i], f); for (int k = 0; k < 6; ++k) out[6 * i + k] = f[k]; } }
void hs_guided(const uint16_t *roots, size_t n_roots, uint32_t rollouts, uint64_t seed, uint64_t first, uint32_t root_base,
               int max_plies, double eps, const double *w, int8_t *outcome, uint8_t *plies)
{
    // same policy as launch_playouts_guided: the root's greedy line is worked out once and followed until the first exploring ply
    for (size_t r = 0; r < n_roots; ++r) {
        const bool use_line = rollouts >= 4 && eps < 1.0;
        GuidedLine line;
        if (use_line) line = guided_line(ld(roots + 8 * r), w);
        for (uint32_t j = 0; j < rollouts; ++j) {
            int pl;
            outcome[r * rollouts + j] = (int8_t)guided_playout_one(ld(roots + 8 * r), seed, first + j, root_base + (uint32_t)r, max_plies, eps, w, &pl,
                                                                   use_line ? &line : nullptr);
            plies[r * rollouts + j] = (uint8_t)pl;
        }
    }
}
int hs_select(uint32_t lo, uint32_t hi, int k) { return select_bit64(lo, hi, popc(lo), k); }
#ifdef QK_HOSTSIM_PLAYOUT
void hs_playouts(const uint16_t *roots, size_t n_roots, uint32_t rollouts, uint64_t seed, uint64_t first, uint32_t root_base,
                 int max_plies, int8_t *outcome, uint8_t *plies)
{
    for (size_t r = 0; r < n_roots; ++r) {
        RootRec rec = make_root(ld(roots + 8 * r));
        for (uint32_t j = 0; j < rollouts; ++j) {
            int pl;
            outcome[r * rollouts + j] = (int8_t)playout_one(rec, seed, first + j, root_base + (uint32_t)r, max_plies, &pl);
            plies[r * rollouts + j] = (uint8_t)pl;
        }
    }
}
// the same playouts through the C++ model of the predicated ply (move table image, 16-bit byte table, blocked entry, sign
// test) that k_playout_pred runs in PTX; outcome from player 0's point of view, no cut-off
void hs_playouts_pred(const uint16_t *roots, size_t n_roots, uint32_t rollouts, uint64_t seed, uint64_t first, uint32_t root_base,
                      int8_t *outcome, uint8_t *slots)
{
    static uint32_t lut[QK_MOVE_LUT_BYTES / 4];
    static uint8_t sel8r[2048];
    for (uint32_t i = 0; i < QK_MOVE_LUT_BYTES / 4; ++i) lut[i] = move_lut_image_word(i);
    for (uint32_t i = 0; i < 2048; ++i) sel8r[i] = (uint8_t)sel8_row(i);
    for (size_t r = 0; r < n_roots; ++r) {
        const RootRec rec = make_root(ld(roots + 8 * r));
        const uint32_t immediate = (rec.flags >> QK_ROOT_IMMEDIATE_SHIFT) & 3u;
        const bool a_is_p0 = (rec.flags & 1u) == 0;
        for (uint32_t j = 0; j < rollouts; ++j) {
            int res = 0, n_slots = 0;
            if (immediate) res = immediate == 1 ? 1 : -1;
            else {
                PlayState st = load_play_state(rec);
                uint32_t act = 0xFFFFFFFFu, acc = 0;
                const uint32_t col = (uint32_t)((r + j) & 7u) * 16u;               // any column must do
                Philox4 rnd = { { 0, 0, 0, 0 } };
                for (int ply = 0; act; ++ply) {
                    if ((ply & 3) == 0) rnd = stream_block(seed, first + j, root_base + (uint32_t)r, (uint32_t)(ply >> 2));
                    ply_pred_model(st, ply & 1, rnd.v[ply & 3], lut, sel8r, col, act, acc);
                    ++n_slots;
                }
                res = (acc != 0) == a_is_p0 ? 1 : -1;
            }
            outcome[r * rollouts + j] = (int8_t)res;
            slots[r * rollouts + j] = (uint8_t)n_slots;
        }
    }
}
#endif
}
