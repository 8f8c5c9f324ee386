/*
 * quantik_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * A plain-C restatement of the reference's (mberlanda/quantik-core-py 1.1.0)
 * algorithms for the batched-simulation hot path.  It deliberately follows the
 * reference's *loops* (12 line masks, 16 squares, 8x24(x2) symmetry candidates
 * compared as little-endian byte strings) instead of the bit tricks the CUDA
 * kernels use, so that a disagreement between the two is informative.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (libquantik_b200.so)
 * never links, loads or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function
 * below against golden vectors produced by importing the reference itself
 * (oracle/gen_golden.py -> tests/golden/), and against the reference's own
 * published known answers (GAME_TREE_ANALYSIS.md:4-11, tests/test_core.py,
 * tests/test_api_portability_report.py).
 *
 * All path:line citations are relative to /root/reference/.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define QO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------ */
/* constants: src/quantik_core/commons.py:24-52                              */
/* ------------------------------------------------------------------------ */
static const uint16_t WIN_MASKS[12] = {
    /* rows    commons.py:24-29 */ 0x000F, 0x00F0, 0x0F00, 0xF000,
    /* columns commons.py:32-37 */ 0x1111, 0x2222, 0x4444, 0x8888,
    /* zones   commons.py:40-45 */ 0x0033, 0x00CC, 0x3300, 0xCC00,
};
#define MAX_PIECES_PER_SHAPE 2 /* commons.py:20 */

/* ValidationResult codes: src/quantik_core/state_validator.py:16-27 */
enum { V_OK = 0, V_TURN_BALANCE = 1, V_SHAPE_COUNT = 2, V_NOT_PLAYER_TURN = 3,
       V_ILLEGAL_PLACEMENT = 4, V_PIECE_OVERLAP = 5 };

static int bit_count16(unsigned v) { int c = 0; while (v) { c += v & 1u; v >>= 1; } return c; }

/* ------------------------------------------------------------------------ */
/* state_validator._validate_game_state_single_pass  state_validator.py:126 */
/* ------------------------------------------------------------------------ */
QO_API int qo_validate(const uint16_t bb[8], int *player_out)
{
    int counts[8], p0 = 0, p1 = 0;
    unsigned all = 0;
    *player_out = -1;
    /* _validate_piece_counts_and_overlaps  state_validator.py:38-77 */
    for (int i = 0; i < 8; ++i) {
        unsigned v = bb[i];
        int c = bit_count16(v);
        if (c > MAX_PIECES_PER_SHAPE) return V_SHAPE_COUNT;
        if (i < 4) p0 += c; else p1 += c;
        if (all & v) return V_PIECE_OVERLAP;
        all |= v;
        counts[i] = c;
    }
    /* _validate_turn_balance  state_validator.py:80-98 */
    int diff = p0 - p1, next;
    if (diff == 0) next = 0; else if (diff == 1) next = 1; else return V_TURN_BALANCE;
    /* _validate_placement_legality  state_validator.py:101-123 */
    for (int s = 0; s < 4; ++s) {
        if (counts[s] == 0 || counts[4 + s] == 0) continue;
        for (int l = 0; l < 12; ++l)
            if ((bb[s] & WIN_MASKS[l]) && (bb[4 + s] & WIN_MASKS[l])) return V_ILLEGAL_PLACEMENT;
    }
    *player_out = next;
    return V_OK;
}

/* ------------------------------------------------------------------------ */
/* move.generate_legal_moves  move.py:199-268 (+ _is_move_legal_on_position  */
/* :169-196).  Result as the 64-bit action mask, bit = shape*16+position     */
/* (api_portability_report.py:75-86, docs/API_PORTABILITY.md:21).            */
/* Invalid state -> player 0, no moves (move.py:223-225).                    */
/* ------------------------------------------------------------------------ */
QO_API uint64_t qo_legal_mask(const uint16_t bb[8], int *player_out, int *status_out)
{
    int player;
    int st = qo_validate(bb, &player);
    if (status_out) *status_out = st;
    if (st != V_OK) { if (player_out) *player_out = 0; return 0; }
    if (player_out) *player_out = player;
    uint64_t mask = 0;
    for (int shape = 0; shape < 4; ++shape) {
        if (bit_count16(bb[player * 4 + shape]) >= MAX_PIECES_PER_SHAPE) continue;
        unsigned opp = bb[(1 - player) * 4 + shape];
        for (int pos = 0; pos < 16; ++pos) {
            unsigned pm = 1u << pos;
            int occupied = 0;
            for (int i = 0; i < 8; ++i) if (bb[i] & pm) occupied = 1; /* game_utils.py:398-415 */
            if (occupied) continue;
            int legal = 1;
            for (int l = 0; l < 12; ++l)
                if ((pm & WIN_MASKS[l]) && (opp & WIN_MASKS[l])) { legal = 0; break; }
            if (legal) mask |= 1ull << (shape * 16 + pos);
        }
    }
    return mask;
}

/* move.apply_move  move.py:135-166: OR the bit in, no legality check. */
QO_API void qo_apply(const uint16_t in[8], int player, int shape, int pos, uint16_t out[8])
{
    for (int i = 0; i < 8; ++i) out[i] = in[i];
    out[player * 4 + shape] |= (uint16_t)(1u << pos);
}

/* game_utils.has_winning_line  game_utils.py:123-158 */
QO_API int qo_has_winning_line(const uint16_t bb[8])
{
    unsigned u[4];
    for (int s = 0; s < 4; ++s) u[s] = bb[s] | bb[s + 4];
    for (int l = 0; l < 12; ++l) {
        int all = 1;
        for (int s = 0; s < 4; ++s) if (!(u[s] & WIN_MASKS[l])) { all = 0; break; }
        if (all) return 1;
    }
    return 0;
}

/* game_utils.check_game_winner  game_utils.py:161-189 -> WinStatus 0/1/2 */
QO_API int qo_check_game_winner(const uint16_t bb[8])
{
    if (!qo_has_winning_line(bb)) return 0;
    int t0 = 0, t1 = 0;
    for (int s = 0; s < 4; ++s) { t0 += bit_count16(bb[s]); t1 += bit_count16(bb[4 + s]); }
    return t0 > t1 ? 1 : 2;
}

/* ------------------------------------------------------------------------ */
/* symmetry.SymmetryHandler  symmetry.py:101-423                             */
/* ------------------------------------------------------------------------ */
static void d4_fn(int idx, int r, int c, int *r2, int *c2)
{   /* symmetry.py:101-110 */
    switch (idx) {
    case 0: *r2 = r;     *c2 = c;     break; /* id     */
    case 1: *r2 = c;     *c2 = 3 - r; break; /* rot90  */
    case 2: *r2 = 3 - r; *c2 = 3 - c; break; /* rot180 */
    case 3: *r2 = 3 - c; *c2 = r;     break; /* rot270 */
    case 4: *r2 = r;     *c2 = 3 - c; break; /* reflV  */
    case 5: *r2 = 3 - r; *c2 = c;     break; /* reflH  */
    case 6: *r2 = c;     *c2 = r;     break; /* reflD  */
    default:*r2 = 3 - c; *c2 = 3 - r; break; /* reflAD */
    }
}

/* permute16: scatter bit i to mapping[i]  (symmetry.py:159-175,199-204) */
QO_API uint16_t qo_permute16(uint16_t m, int d4)
{
    unsigned y = 0;
    for (int i = 0; i < 16; ++i) {
        if (!((m >> i) & 1u)) continue;
        int r2, c2;
        d4_fn(d4, i / 4, i % 4, &r2, &c2);
        y |= 1u << (r2 * 4 + c2);
    }
    return (uint16_t)y;
}

/* itertools.permutations(range(4)) in its lexicographic order (symmetry.py:113) */
static int PERMS[24][4];
static int perms_ready = 0;
static void build_perms(void)
{
    if (perms_ready) return;
    int n = 0;
    for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) { if (b == a) continue;
        for (int c = 0; c < 4; ++c) { if (c == a || c == b) continue;
            int d = 6 - a - b - c;
            PERMS[n][0] = a; PERMS[n][1] = b; PERMS[n][2] = c; PERMS[n][3] = d; ++n; } }
    perms_ready = 1;
}
QO_API void qo_perm(int index, int out[4]) { build_perms(); for (int i = 0; i < 4; ++i) out[i] = PERMS[index][i]; }

/* apply_symmetry  symmetry.py:252-287 : out[c][s] = C_c[perm[s]] */
QO_API void qo_apply_symmetry(const uint16_t bb[8], int d4, int color_swap, const int perm[4], uint16_t out[8])
{
    uint16_t g0[4], g1[4];
    for (int s = 0; s < 4; ++s) { g0[s] = qo_permute16(bb[s], d4); g1[s] = qo_permute16(bb[4 + s], d4); }
    const uint16_t *c0 = color_swap ? g1 : g0, *c1 = color_swap ? g0 : g1;
    for (int s = 0; s < 4; ++s) { out[s] = c0[perm[s]]; out[4 + s] = c1[perm[s]]; }
}

static void pack_le(const uint16_t bb[8], unsigned char p[16])
{   /* struct.pack("<8H", ...)  symmetry.py:341 */
    for (int i = 0; i < 8; ++i) { p[2 * i] = (unsigned char)(bb[i] & 0xFF); p[2 * i + 1] = (unsigned char)(bb[i] >> 8); }
}

/* find_canonical_form  symmetry.py:289-354.  Iteration order d4 -> colour ->
 * perm, strict '<' on the packed bytes, so the FIRST minimum wins.
 * transform_out = {d4, color_swap, perm_index, perm[0..3]}                   */
QO_API void qo_find_canonical(const uint16_t bb[8], int color_swap, uint16_t out[8], int transform_out[7])
{
    build_perms();
    unsigned char best[16], cand[16];
    int have = 0;
    for (int d4 = 0; d4 < 8; ++d4) {
        for (int cs = 0; cs <= (color_swap ? 1 : 0); ++cs) {
            for (int pi = 0; pi < 24; ++pi) {
                uint16_t c[8];
                qo_apply_symmetry(bb, d4, cs, PERMS[pi], c);
                pack_le(c, cand);
                if (!have || memcmp(cand, best, 16) < 0) {
                    have = 1;
                    memcpy(best, cand, 16);
                    for (int i = 0; i < 8; ++i) out[i] = c[i];
                    if (transform_out) {
                        transform_out[0] = d4; transform_out[1] = cs; transform_out[2] = pi;
                        for (int i = 0; i < 4; ++i) transform_out[3 + i] = PERMS[pi][i];
                    }
                }
            }
        }
    }
}

/* count_orbit_size  symmetry.py:356-396 : distinct images among the 192 */
QO_API int qo_orbit_size(const uint16_t bb[8])
{
    build_perms();
    unsigned char seen[192][16];
    int n = 0;
    for (int d4 = 0; d4 < 8; ++d4)
        for (int pi = 0; pi < 24; ++pi) {
            uint16_t c[8]; unsigned char p[16];
            qo_apply_symmetry(bb, d4, 0, PERMS[pi], c);
            pack_le(c, p);
            int dup = 0;
            for (int k = 0; k < n; ++k) if (!memcmp(seen[k], p, 16)) { dup = 1; break; }
            if (!dup) memcpy(seen[n++], p, 16);
        }
    return n;
}

/* apply_symmetry_to_move  symmetry.py:425-464: action = shape*16+pos of mover */
QO_API void qo_apply_symmetry_to_move(int player, int shape, int pos, int d4, int color_swap,
                                      const int perm[4], int out[3])
{
    int r2, c2;
    d4_fn(d4, pos / 4, pos % 4, &r2, &c2);
    int ns = 0;
    for (int i = 0; i < 4; ++i) if (perm[i] == shape) ns = i; /* perm.index(shape) :461 */
    out[0] = color_swap ? 1 - player : player; out[1] = ns; out[2] = r2 * 4 + c2;
}

/* ------------------------------------------------------------------------ */
/* batch wrappers (N x 8 uint16, the storage/compact_state.py:125-147 layout) */
/* ------------------------------------------------------------------------ */
QO_API void qo_movegen_batch(const uint16_t *states, size_t n, uint64_t *mask, uint8_t *to_move, uint8_t *status)
{
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; ++i) {
        int p, st;
        mask[i] = qo_legal_mask(states + 8 * i, &p, &st);
        to_move[i] = (uint8_t)p; status[i] = (uint8_t)st;
    }
}
QO_API void qo_win_batch(const uint16_t *states, size_t n, uint8_t *win)
{
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; ++i) win[i] = (uint8_t)qo_check_game_winner(states + 8 * i);
}
QO_API void qo_apply_batch(const uint16_t *states, const uint8_t *action, size_t n, uint16_t *out)
{   /* action = shape*16+pos for the side to move (state_validator), or | 0x80 | player << 6 */
    for (size_t i = 0; i < n; ++i) {
        int p; qo_validate(states + 8 * i, &p); if (p < 0) p = 0;
        if (action[i] & 0x80) p = (action[i] >> 6) & 1;
        qo_apply(states + 8 * i, p, (action[i] >> 4) & 3, action[i] & 15, out + 8 * i);
    }
}
QO_API void qo_canonical_batch(const uint16_t *states, size_t n, int color_swap, uint16_t *payload, int32_t *transform)
{
    build_perms();
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; ++i) {
        int t[7];
        qo_find_canonical(states + 8 * i, color_swap, payload + 8 * i, t);
        if (transform) for (int k = 0; k < 7; ++k) transform[7 * i + k] = t[k];
    }
}
QO_API void qo_orbit_batch(const uint16_t *states, size_t n, uint8_t *orbit)
{
    build_perms();
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; ++i) orbit[i] = (uint8_t)qo_orbit_size(states + 8 * i);
}

/* ------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., Random123; same generator as cuRAND's       */
/* curand_philox4x32_10).  Pinned by the Random123 known-answer vectors in    */
/* tests/test_oracle_golden.py.                                               */
/* ------------------------------------------------------------------------ */
QO_API void qo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* The shared random-stream contract (DESIGN.md "Random stream"):
 *   key     = (seed lo32, seed hi32)
 *   counter = (rollout lo32, rollout hi32, ply/4, root_index)
 *   ply p of the playout uses output word p%4;  move index = mulhi32(word, n_legal)
 *   the move is the idx-th set bit (ascending) of the 64-bit action mask.       */
static uint32_t stream_word(uint64_t seed, uint64_t rollout, uint32_t root_index, int ply)
{
    uint32_t ctr[4] = { (uint32_t)rollout, (uint32_t)(rollout >> 32), (uint32_t)(ply / 4), root_index };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) }, out[4];
    qo_philox4x32_10(ctr, key, out);
    return out[ply % 4];
}
static int kth_set_bit(uint64_t m, int k)
{
    for (int b = 0; b < 64; ++b) if ((m >> b) & 1ull) { if (k == 0) return b; --k; }
    return -1;
}
static int popcount64(uint64_t m) { int c = 0; while (m) { c += (int)(m & 1ull); m >>= 1; } return c; }

/* One random playout.  Loop shape of BeamSearchEngine._rollout
 * (beam_search.py:624-645) when max_plies < 0, and of MCTSEngine._simulate
 * (mcts.py:385-437; cutoff test `depth < max_depth` BEFORE the win test, and
 * 0.0 after the loop) when max_plies >= 0.  Returns +1 / -1 / 0 (P0 view). */
QO_API int qo_playout(const uint16_t root[8], uint64_t seed, uint64_t rollout, uint32_t root_index,
                      int max_plies, int *plies_out, uint16_t final_out[8])
{
    uint16_t cur[8], nxt[8];
    memcpy(cur, root, 16);
    int ply = 0, result;
    for (;;) {
        if (max_plies >= 0 && ply >= max_plies) { result = 0; break; }           /* mcts.py:412,436 */
        int w = qo_check_game_winner(cur);                                       /* beam :633 / mcts :414 */
        if (w) { result = (w == 1) ? 1 : -1; break; }
        int player, st;
        uint64_t mask = qo_legal_mask(cur, &player, &st);                        /* :637 / :422 */
        int n = popcount64(mask);
        if (n == 0) { result = (player == 0) ? -1 : 1; break; }                  /* :641-642 / :427-429 */
        uint32_t r = stream_word(seed, rollout, root_index, ply);
        int idx = (int)(((uint64_t)r * (uint64_t)n) >> 32);
        int a = kth_set_bit(mask, idx);
        qo_apply(cur, player, a >> 4, a & 15, nxt);                              /* :645 / :433 */
        memcpy(cur, nxt, 16);
        ++ply;
    }
    if (plies_out) *plies_out = ply;
    if (final_out) memcpy(final_out, cur, 16);
    return result;
}

/* Batched playouts: n_roots x rollouts, rollout ids first_rollout .. +rollouts-1,
 * root ids root_index_base + r.  Tallies per root; optional per-playout outcome
 * (int8) and ply count (uint8), index r*rollouts + j.                          */
QO_API void qo_playouts(const uint16_t *roots, size_t n_roots, uint32_t rollouts, uint64_t seed,
                        uint64_t first_rollout, uint32_t root_index_base, int max_plies,
                        uint32_t *wins_p0, uint32_t *wins_p1, uint32_t *draws,
                        int8_t *per_playout, uint8_t *per_plies)
{
    for (size_t r = 0; r < n_roots; ++r) {
        long w0 = 0, w1 = 0, dr = 0;
        #pragma omp parallel for schedule(static) reduction(+:w0,w1,dr)
        for (long j = 0; j < (long)rollouts; ++j) {
            int plies;
            int res = qo_playout(roots + 8 * r, seed, first_rollout + (uint64_t)j,
                                 root_index_base + (uint32_t)r, max_plies, &plies, NULL);
            if (res > 0) ++w0; else if (res < 0) ++w1; else ++dr;
            if (per_playout) per_playout[r * (size_t)rollouts + j] = (int8_t)res;
            if (per_plies) per_plies[r * (size_t)rollouts + j] = (uint8_t)plies;
        }
        if (wins_p0) wins_p0[r] = (uint32_t)w0;
        if (wins_p1) wins_p1[r] = (uint32_t)w1;
        if (draws) draws[r] = (uint32_t)dr;
    }
}

/* Mid-game root generator for BASELINE config 4 (rejection
 * rule of benchmarks/dataset.py:38-51): root i = empty board + (min_plies +
 * i % span) random plies, stream key seed ^ 0x524F4F5453 ("ROOTS"), counter
 * (i, attempt, ply/4, 0); a line that hits a finished game (line completed or
 * mover without moves) before or at the target ply count restarts with
 * attempt+1.                                                                    */
QO_API void qo_make_roots(uint64_t seed, size_t first, size_t n, int min_plies, int span, uint16_t *out)
{
    const uint64_t rseed = seed ^ 0x524F4F5453ull;
    #pragma omp parallel for schedule(dynamic, 64)
    for (long ii = 0; ii < (long)n; ++ii) {
        uint64_t i = (uint64_t)first + (uint64_t)ii;
        int target = min_plies + (int)(i % (uint64_t)span);
        for (uint32_t attempt = 0;; ++attempt) {
            uint16_t cur[8] = {0}, nxt[8];
            int ok = 1;
            for (int ply = 0; ply <= target; ++ply) {
                int player, st;
                if (qo_check_game_winner(cur)) { ok = 0; break; }
                uint64_t mask = qo_legal_mask(cur, &player, &st);
                int nl = popcount64(mask);
                if (nl == 0) { ok = 0; break; }
                if (ply == target) break;
                uint32_t ctr[4] = { (uint32_t)i, attempt, (uint32_t)(ply / 4), 0u };
                uint32_t key[2] = { (uint32_t)rseed, (uint32_t)(rseed >> 32) }, w[4];
                qo_philox4x32_10(ctr, key, w);
                int idx = (int)(((uint64_t)w[ply % 4] * (uint64_t)nl) >> 32);
                int a = kth_set_bit(mask, idx);
                qo_apply(cur, player, a >> 4, a & 15, nxt);
                memcpy(cur, nxt, 16);
            }
            if (ok) { memcpy(out + 8 * ii, cur, 16); break; }
        }
    }
}

/* ------------------------------------------------------------------------ */
/* perft: raw depth-first enumeration with the counting conventions of        */
/* game_stats.SymmetryTable (game_stats.py:359-485) but no symmetry merging:   */
/*   nodes[d]   = number of legal move sequences of length d (= the table's    */
/*                "Total Legal Moves" when all roots carry their multiplicity) */
/*   wins_pX[d] = sequences whose d-th move completed a line, X = the mover    */
/*   blocked_pX_loses[d] = non-won positions at depth d where X is to move and */
/*                has no legal move (the reference books these one ply later,  */
/*                as a win of the opponent at depth d+1 with 0 moves:          */
/*                game_stats.py:397-398,426-443)                               */
/* depth d is relative to the root (root = depth 0).  Arrays have depth+1      */
/* entries; the blocked arrays are filled for d < depth only (positions at the */
/* last depth are counted, not examined), entry [depth] stays 0.               */
/* ------------------------------------------------------------------------ */
typedef struct { uint64_t nodes[18], w0[18], w1[18], b0[18], b1[18]; } perft_acc;

static void perft_rec(const uint16_t cur[8], int d, int depth, uint64_t mult, perft_acc *acc)
{
    int player, st;
    if (d == depth) return;
    uint64_t mask = qo_legal_mask(cur, &player, &st);
    if (mask == 0) { if (player == 0) acc->b0[d] += mult; else acc->b1[d] += mult; return; }
    for (int a = 0; a < 64; ++a) {
        if (!((mask >> a) & 1ull)) continue;
        uint16_t nxt[8];
        qo_apply(cur, player, a >> 4, a & 15, nxt);
        acc->nodes[d + 1] += mult;
        if (qo_check_game_winner(nxt)) { if (player == 0) acc->w0[d + 1] += mult; else acc->w1[d + 1] += mult; continue; }
        perft_rec(nxt, d + 1, depth, mult, acc);
    }
}

QO_API void qo_perft(const uint16_t *roots, const uint64_t *mult, size_t n_roots, int depth,
                     uint64_t *nodes, uint64_t *wins_p0, uint64_t *wins_p1,
                     uint64_t *blocked_p0_loses, uint64_t *blocked_p1_loses)
{
    perft_acc total; memset(&total, 0, sizeof total);
    #pragma omp parallel
    {
        perft_acc acc; memset(&acc, 0, sizeof acc);
        #pragma omp for schedule(dynamic, 1)
        for (long r = 0; r < (long)n_roots; ++r) {
            uint64_t m = mult ? mult[r] : 1;
            acc.nodes[0] += m;
            if (qo_check_game_winner(roots + 8 * r)) continue; /* finished root: nothing below */
            perft_rec(roots + 8 * r, 0, depth, m, &acc);
        }
        #pragma omp critical
        for (int d = 0; d <= depth; ++d) {
            total.nodes[d] += acc.nodes[d]; total.w0[d] += acc.w0[d]; total.w1[d] += acc.w1[d];
            total.b0[d] += acc.b0[d]; total.b1[d] += acc.b1[d];
        }
    }
    for (int d = 0; d <= depth; ++d) {
        nodes[d] = total.nodes[d]; wins_p0[d] = total.w0[d]; wins_p1[d] = total.w1[d];
        blocked_p0_loses[d] = total.b0[d]; blocked_p1_loses[d] = total.b1[d];
    }
}

/* ------------------------------------------------------------------------ */
/* Level-synchronous canonical BFS with multiplicity = SymmetryTable itself    */
/* (game_stats.py:276-505): parents in insertion order, children in move       */
/* order, first-seen representative kept, multiplicities summed.               */
/* Per depth d (1-based like the reference's table): total_legal_moves,        */
/* unique_canonical, p0_wins, p1_wins, ongoing.                                */
/* ------------------------------------------------------------------------ */
typedef struct { uint16_t canon[8]; uint16_t rep[8]; uint64_t mult; } bfs_state;
typedef struct { bfs_state *v; size_t n, cap; int64_t *slots; size_t nslots; } bfs_level;

static uint64_t hash16(const uint16_t k[8])
{
    uint64_t a, b; memcpy(&a, k, 8); memcpy(&b, k + 4, 8);
    uint64_t h = a * 0x9E3779B97F4A7C15ull ^ (b + 0xD6E8FEB86659FD93ull) * 0xC2B2AE3D27D4EB4Full;
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    return h;
}
static void level_init(bfs_level *L, size_t expect)
{
    L->n = 0; L->cap = expect < 1024 ? 1024 : expect;
    L->v = (bfs_state *)malloc(L->cap * sizeof(bfs_state));
    L->nslots = 1; while (L->nslots < 2 * L->cap) L->nslots <<= 1;
    L->slots = (int64_t *)malloc(L->nslots * sizeof(int64_t));
    for (size_t i = 0; i < L->nslots; ++i) L->slots[i] = -1;
}
static void level_free(bfs_level *L) { free(L->v); free(L->slots); L->v = NULL; L->slots = NULL; }
static void level_grow(bfs_level *L)
{
    L->cap *= 2; L->v = (bfs_state *)realloc(L->v, L->cap * sizeof(bfs_state));
    free(L->slots); L->nslots *= 2; L->slots = (int64_t *)malloc(L->nslots * sizeof(int64_t));
    for (size_t i = 0; i < L->nslots; ++i) L->slots[i] = -1;
    for (size_t j = 0; j < L->n; ++j) {
        size_t s = hash16(L->v[j].canon) & (L->nslots - 1);
        while (L->slots[s] >= 0) s = (s + 1) & (L->nslots - 1);
        L->slots[s] = (int64_t)j;
    }
}
static void level_add(bfs_level *L, const uint16_t canon[8], const uint16_t rep[8], uint64_t mult)
{   /* _add_or_update_state  game_stats.py:487-505 */
    size_t s = hash16(canon) & (L->nslots - 1);
    while (L->slots[s] >= 0) {
        bfs_state *e = &L->v[L->slots[s]];
        if (!memcmp(e->canon, canon, 16)) { e->mult += mult; return; }
        s = (s + 1) & (L->nslots - 1);
    }
    if (L->n == L->cap) { level_grow(L); level_add(L, canon, rep, mult); return; }
    memcpy(L->v[L->n].canon, canon, 16); memcpy(L->v[L->n].rep, rep, 16); L->v[L->n].mult = mult;
    L->slots[s] = (int64_t)L->n; L->n++;
}

/* stats: [max_depth][5] = total_legal_moves, unique, p0_wins, p1_wins, ongoing.
 * If out_states != NULL the canonical states (payload order of insertion) of
 * depth `emit_depth` are copied there (cap entries) with multiplicities.       */
QO_API long qo_symmetry_table(int max_depth, uint64_t *stats, int emit_depth,
                              uint16_t *out_states, uint64_t *out_mult, size_t cap)
{
    build_perms();
    bfs_level cur; level_init(&cur, 1024);
    uint16_t empty[8] = {0};
    level_add(&cur, empty, empty, 1);                          /* _initialize_analysis */
    long emitted = -1;
    for (int d = 1; d <= max_depth; ++d) {
        bfs_level nxt; level_init(&nxt, cur.n * 8);
        uint64_t total = 0, w0 = 0, w1 = 0;
        /* canonicalise children in parallel (chunked), merge sequentially in order */
        const size_t CH = 4096;
        typedef struct { uint16_t canon[8], rep[8]; uint8_t kind; } child_t; /* kind 0 skip 1 state */
        child_t *buf = (child_t *)malloc(CH * 64 * sizeof(child_t));
        uint64_t *ptot = (uint64_t *)malloc(CH * 3 * sizeof(uint64_t));
        int *pcnt = (int *)malloc(CH * sizeof(int));
        for (size_t base = 0; base < cur.n; base += CH) {
            size_t m = cur.n - base < CH ? cur.n - base : CH;
            #pragma omp parallel for schedule(dynamic, 16)
            for (long j = 0; j < (long)m; ++j) {
                const bfs_state *ps = &cur.v[base + j];
                int player, st, cnt = 0;
                uint64_t mask = qo_legal_mask(ps->rep, &player, &st);   /* game_stats.py:391 */
                uint64_t t = 0, a0 = 0, a1 = 0;
                if (mask == 0) {                                        /* :397-398, :426-443 */
                    if (player == 0) a1 = ps->mult; else a0 = ps->mult;
                } else {
                    t = (uint64_t)popcount64(mask) * ps->mult;          /* :401 */
                    for (int a = 0; a < 64; ++a) {
                        if (!((mask >> a) & 1ull)) continue;
                        child_t *c = &buf[j * 64 + cnt];
                        qo_apply(ps->rep, player, a >> 4, a & 15, c->rep);   /* :453 */
                        if (qo_check_game_winner(c->rep)) {                  /* :464, :411-415 */
                            if (player == 0) a0 += ps->mult; else a1 += ps->mult;
                            continue;
                        }
                        qo_find_canonical(c->rep, 0, c->canon, NULL);        /* :468 */
                        c->kind = 1; ++cnt;
                    }
                }
                pcnt[j] = cnt; ptot[3 * j] = t; ptot[3 * j + 1] = a0; ptot[3 * j + 2] = a1;
            }
            for (size_t j = 0; j < m; ++j) {
                total += ptot[3 * j]; w0 += ptot[3 * j + 1]; w1 += ptot[3 * j + 2];
                for (int k = 0; k < pcnt[j]; ++k)
                    level_add(&nxt, buf[j * 64 + k].canon, buf[j * 64 + k].rep, cur.v[base + j].mult);
            }
        }
        free(buf); free(ptot); free(pcnt);
        uint64_t ongoing = 0;
        for (size_t j = 0; j < nxt.n; ++j) ongoing += nxt.v[j].mult;
        stats[(d - 1) * 5 + 0] = total; stats[(d - 1) * 5 + 1] = nxt.n;
        stats[(d - 1) * 5 + 2] = w0;    stats[(d - 1) * 5 + 3] = w1; stats[(d - 1) * 5 + 4] = ongoing;
        if (d == emit_depth && out_states) {
            emitted = (long)nxt.n;
            for (size_t j = 0; j < nxt.n && j < cap; ++j) {
                memcpy(out_states + 8 * j, nxt.v[j].canon, 16);
                if (out_mult) out_mult[j] = nxt.v[j].mult;
            }
        }
        level_free(&cur); cur = nxt;
    }
    level_free(&cur);
    return emitted;
}

/* ------------------------------------------------------------------------ */
/* Opening-book BFS (examples/generate_opening_book.py:185-280 _expand_chunk +   */
/* :458-496 main-process dedup) summarised like opening_book_summary.build_summary */
/* (opening_book_summary.py:36-93): per depth d                                   */
/*   positions = distinct canonical keys first reached at depth d (won ones too)  */
/*   terminal  = those that are won, or whose mover has no legal move             */
/*   edges     = distinct (parent key, child key) pairs of the expanded parents   */
/* out[(d)*3 + {0,1,2}].                                                          */
/* ------------------------------------------------------------------------ */
QO_API void qo_opening_book_summary(int max_depth, uint64_t *out)
{
    build_perms();
    bfs_level cur; level_init(&cur, 1024);
    uint16_t empty[8] = {0};
    level_add(&cur, empty, empty, 1);
    for (int d = 0; d <= max_depth; ++d) {
        bfs_level nxt; level_init(&nxt, cur.n * 8);
        uint64_t terminal = 0, edges = 0;
        for (size_t j = 0; j < cur.n; ++j) {
            const uint16_t *bb = cur.v[j].rep;
            if (qo_check_game_winner(bb)) { ++terminal; continue; }          /* :205-211 */
            int player, st;
            uint64_t mask = qo_legal_mask(bb, &player, &st);
            if (mask == 0) { ++terminal; continue; }                          /* :212-223 */
            uint16_t kids[64][8]; int nk = 0;
            for (int a = 0; a < 64; ++a) {
                if (!((mask >> a) & 1ull)) continue;
                uint16_t child[8], canon[8];
                qo_apply(bb, player, a >> 4, a & 15, child);                  /* :237-242 */
                qo_find_canonical(child, 0, canon, NULL);
                int dup = 0;
                for (int k = 0; k < nk; ++k) if (!memcmp(kids[k], canon, 16)) { dup = 1; break; }
                if (!dup) { memcpy(kids[nk++], canon, 16); }                  /* PRIMARY KEY (parent_key, child_key) */
                if (d < max_depth) level_add(&nxt, canon, child, 1);
            }
            edges += (uint64_t)nk;
        }
        out[d * 3 + 0] = cur.n; out[d * 3 + 1] = terminal; out[d * 3 + 2] = edges;
        level_free(&cur); cur = nxt;
    }
    level_free(&cur);
}

/* ------------------------------------------------------------------------ */
/* evaluation.features / evaluate  (evaluation.py:130-212) and the eval-guided  */
/* rollout policy MCTSEngine._select_rollout_move (mcts.py:345-383).            */
/* ------------------------------------------------------------------------ */
static int placement_is_legal(const uint16_t bb[8], int player, int shape, int position)
{   /* evaluation.py:99-112 */
    unsigned opp = bb[(1 - player) * 4 + shape], pm = 1u << position;
    for (int l = 0; l < 12; ++l) if ((pm & WIN_MASKS[l]) && (opp & WIN_MASKS[l])) return 0;
    return 1;
}
static int count_legal_moves_for(const uint16_t bb[8], int player)
{   /* evaluation.py:115-127: generate_legal_moves_list(bb, player) -> [] when it is not player's turn */
    int cur, st;
    uint64_t mask = qo_legal_mask(bb, &cur, &st);
    if (st != V_OK || cur != player) return 0;
    return popcount64(mask);
}
QO_API void qo_features(const uint16_t bb[8], int player, double out[6])
{
    int counts[2][4], t0 = 0, t1 = 0;
    for (int s = 0; s < 4; ++s) { counts[0][s] = bit_count16(bb[s]); counts[1][s] = bit_count16(bb[4 + s]); t0 += counts[0][s]; t1 += counts[1][s]; }
    int side_to_move = (t0 - t1 == 0) ? 0 : 1;                       /* game_utils.py:226-231 (valid states only) */
    double sign = side_to_move == player ? 1.0 : -1.0;
    unsigned union_all = 0, su[4];
    for (int i = 0; i < 8; ++i) union_all |= bb[i];
    for (int s = 0; s < 4; ++s) su[s] = bb[s] | bb[s + 4];
    double threat_own = 0, threat_opp = 0, threat_shared = 0, build_two = 0, build_one = 0;
    for (int l = 0; l < 12; ++l) {
        unsigned mask = WIN_MASKS[l];
        int n_present = 0, missing = -1;
        for (int s = 0; s < 4; ++s) { if (su[s] & mask) ++n_present; }
        int occupied = bit_count16(union_all & mask);
        if (n_present < occupied) continue;                          /* dead line */
        if (occupied == 3) {
            for (int s = 0; s < 4; ++s) if (!(su[s] & mask)) { missing = s; break; }
            unsigned e = mask & ~union_all;
            int empty_position = 0; while (!((e >> empty_position) & 1u)) ++empty_position;
            int comp[2];
            for (int side = 0; side < 2; ++side)
                comp[side] = counts[side][missing] < 2 && placement_is_legal(bb, side, missing, empty_position);
            if (comp[player]) threat_own += 1.0;
            if (comp[1 - player]) threat_opp += 1.0;
            if (comp[0] && comp[1]) threat_shared += sign;
        } else if (occupied == 2) build_two += sign;
        else if (occupied == 1) build_one += sign;
    }
    out[0] = threat_own; out[1] = threat_opp; out[2] = threat_shared;
    out[3] = (double)(count_legal_moves_for(bb, player) - count_legal_moves_for(bb, 1 - player));
    out[4] = build_two; out[5] = build_one;
}
QO_API double qo_evaluate(const uint16_t bb[8], int player, const double w[6])
{
    double f[6], acc = 0.0;
    qo_features(bb, player, f);
    for (int i = 0; i < 6; ++i) acc += w[i] * f[i];
    return acc;
}
QO_API void qo_features_batch(const uint16_t *states, const uint8_t *player, size_t n, double *out)
{
    for (size_t i = 0; i < n; ++i) qo_features(states + 8 * i, player[i], out + 6 * i);
}

/* Guided-stream contract: ply p of a guided playout uses Philox block
 * (rollout lo, rollout hi, 0x80000000 | p/2, root_index); word 2*(p%2) is the epsilon draw
 * (explore iff epsilon > 0 and word * 2^-32 < epsilon), word 2*(p%2)+1 the uniform pick. */
QO_API int qo_guided_playout(const uint16_t root[8], uint64_t seed, uint64_t rollout, uint32_t root_index,
                             int max_plies, double epsilon, const double weights[6], int *plies_out)
{
    uint16_t cur[8], nxt[8];
    memcpy(cur, root, 16);
    int ply = 0, result;
    for (;;) {
        if (max_plies >= 0 && ply >= max_plies) { result = 0; break; }
        int w = qo_check_game_winner(cur);
        if (w) { result = (w == 1) ? 1 : -1; break; }
        int player, st;
        uint64_t mask = qo_legal_mask(cur, &player, &st);
        int n = popcount64(mask);
        if (n == 0) { result = (player == 0) ? -1 : 1; break; }
        uint32_t ctr[4] = { (uint32_t)rollout, (uint32_t)(rollout >> 32), 0x80000000u | (uint32_t)(ply / 2), root_index };
        uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) }, r[4];
        qo_philox4x32_10(ctr, key, r);
        uint32_t r_eps = r[2 * (ply % 2)], r_pick = r[2 * (ply % 2) + 1];
        int a = -1;
        if (epsilon > 0 && (double)r_eps * (1.0 / 4294967296.0) < epsilon) {          /* mcts.py:371-374 */
            a = kth_set_bit(mask, (int)(((uint64_t)r_pick * (uint64_t)n) >> 32));
        } else {
            double best = 0; int have = 0;
            for (int m = 0; m < 64; ++m) {
                if (!((mask >> m) & 1ull)) continue;
                uint16_t child[8];
                qo_apply(cur, player, m >> 4, m & 15, child);
                int cp, cs;
                if (qo_has_winning_line(child) || qo_legal_mask(child, &cp, &cs) == 0) { a = m; break; }   /* :378-379 */
                double sc = qo_evaluate(child, player, weights);
                if (!have) { a = m; best = sc; have = 1; }           /* best_move = all_moves[0], best_score = -inf */
                else if (sc > best) { best = sc; a = m; }
            }
        }
        qo_apply(cur, player, a >> 4, a & 15, nxt);
        memcpy(cur, nxt, 16);
        ++ply;
    }
    if (plies_out) *plies_out = ply;
    return result;
}
QO_API void qo_guided_playouts(const uint16_t *roots, size_t n_roots, uint32_t rollouts, uint64_t seed, uint64_t first_rollout,
                               uint32_t root_index_base, int max_plies, double epsilon, const double *weights,
                               uint32_t *wins_p0, uint32_t *wins_p1, uint32_t *draws, int8_t *per_playout, uint8_t *per_plies)
{
    for (size_t r = 0; r < n_roots; ++r) {
        long w0 = 0, w1 = 0, dr = 0;
        #pragma omp parallel for schedule(dynamic, 16) reduction(+:w0,w1,dr)
        for (long j = 0; j < (long)rollouts; ++j) {
            int plies;
            int res = qo_guided_playout(roots + 8 * r, seed, first_rollout + (uint64_t)j, root_index_base + (uint32_t)r,
                                        max_plies, epsilon, weights, &plies);
            if (res > 0) ++w0; else if (res < 0) ++w1; else ++dr;
            if (per_playout) per_playout[r * (size_t)rollouts + j] = (int8_t)res;
            if (per_plies) per_plies[r * (size_t)rollouts + j] = (uint8_t)plies;
        }
        if (wins_p0) wins_p0[r] = (uint32_t)w0;
        if (wins_p1) wins_p1[r] = (uint32_t)w1;
        if (draws) draws[r] = (uint32_t)dr;
    }
}

/* ------------------------------------------------------------------------ */
/* Exact negamax with mate distance = MinimaxEngine._negamax (minimax.py:330-453) */
/* at max_depth 16 (MinimaxEngine.solve, :182-204) without its accelerators       */
/* (alpha-beta, transposition table, sibling dedup never change the value).       */
/* Returns the value from the side to move's perspective as  +-(W - ply_of_end),  */
/* W = 64;  *best = lowest action index among the optimal moves (-1 if terminal). */
/* ------------------------------------------------------------------------ */
static int negamax_rec(const uint16_t bb[8], int ply, int *best_action)
{
    if (best_action) *best_action = -1;
    if (qo_has_winning_line(bb)) return -(64 - ply);                 /* :354-358 */
    int player, st;
    uint64_t mask = qo_legal_mask(bb, &player, &st);
    if (mask == 0) return -(64 - ply);                               /* :366-369 */
    int best = -1000;
    for (int a = 0; a < 64; ++a) {
        if (!((mask >> a) & 1ull)) continue;
        uint16_t child[8];
        qo_apply(bb, player, a >> 4, a & 15, child);
        int v = -negamax_rec(child, ply + 1, NULL);                  /* :423-426 */
        if (v > best) { best = v; if (best_action) *best_action = a; }
    }
    return best;
}
QO_API void qo_solve_batch(const uint16_t *states, size_t n, int8_t *outcome, uint8_t *plies, uint8_t *best_action)
{
    #pragma omp parallel for schedule(dynamic, 1)
    for (long i = 0; i < (long)n; ++i) {
        int act;
        int v = negamax_rec(states + 8 * i, 0, &act);
        outcome[i] = (int8_t)(v > 0 ? 1 : -1);
        plies[i] = (uint8_t)(64 - (v > 0 ? v : -v));
        best_action[i] = (uint8_t)(act < 0 ? 255 : act);
    }
}

QO_API int qo_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
QO_API void qo_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
