// qk_playout_core.cuh -- one ply of a uniform-random Quantik playout on an
// INCREMENTAL state, plus the Philox4x32-10 counter stream that drives it.
//
// Restates, per ply, the loop body shared by BeamSearchEngine._rollout
// (beam_search.py:624-645) and MCTSEngine._simulate (mcts.py:385-437):
//     check_game_winner -> generate_legal_moves -> none: mover loses
//     -> uniform pick -> apply_move
// The rules are not re-derived from the 8 bitboards every ply.  Instead each
// playout carries 11 registers that are updated with a handful of LOP3s when a
// piece lands:
//     free2           EMPTY squares (replicated in both 16-bit lanes)
//     ra_lo/ra_hi     squares+shapes closed to role A (= the root's mover): the
//                     row/column/zone scope of B's pieces of that shape, plus
//                     whole lanes of shapes A has already placed twice
//     rb_lo/rb_hi     same for role B
//     ta_*/tb_*       lanes (shapes) of which the role has placed >= 1 piece
//     lp_lo/lp_hi     per shape (colour-blind) the 12-bit set of lines that
//                     contain it: rows 0-3, columns 4-7, zones 8-11
// lanes: lo = shape A | shape B << 16, hi = shape C | shape D << 16.
// Legal moves of the mover = free2 & ~r?_lo, free2 & ~r?_hi     (2 LOP3)
// A line is complete iff lp_A & lp_B & lp_C & lp_D != 0            (3 LOP3 + 1 SHF)
#pragma once
#include "qk_bits.cuh"

namespace qk {

// ---------------------------------------------------------------- Philox4x32-10
// Salmon et al. (Random123); identical to cuRAND's curand_philox4x32_10 block
// function.  Known-answer vectors are asserted in tests/.
struct Philox4 { uint32_t v[4]; };

QK_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0 = mulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = mulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Philox4 o = { { c0, c1, c2, c3 } };
    return o;
}

// Random-stream contract (DESIGN.md "Random stream"; shared with oracle/):
//   key = (seed lo, seed hi); counter = (rollout lo, rollout hi, ply/4, root index);
//   ply p uses word p % 4; move index = mulhi32(word, n_legal), ascending action order.
QK_HD Philox4 stream_block(uint64_t seed, uint64_t rollout, uint32_t root_index, uint32_t block)
{
    return philox4x32_10((uint32_t)rollout, (uint32_t)(rollout >> 32), block, root_index,
                         (uint32_t)seed, (uint32_t)(seed >> 32));
}

// ---------------------------------------------------------------- per-square table
// scope2 = scope of square p (its row | column | zone), lines2 = its three line bits,
// bit2 = the square itself; all replicated in the two 16-bit lanes so one AND with the
// lane mask routes them.  16 bytes: one LDS.128 per ply.
struct __attribute__((aligned(16))) SquareLut { uint32_t scope2, lines2, bit2, pad; };

QK_HD SquareLut square_lut(int p)
{
    uint32_t scope = (0xFu << (p & 12)) | (0x1111u << (p & 3)) | (0x33u << (p & 10));
    uint32_t lines = (1u << (p >> 2)) | (16u << (p & 3)) | (256u << (((p >> 3) << 1) | ((p >> 1) & 1)));
    SquareLut e = { (scope & 0xFFFFu) * 0x10001u, lines * 0x10001u, 0x00010001u << p, 0u };
    return e;
}
// sel8 table entry: byte offset (16 * position) into the SquareLut table of the k-th set bit
// of a byte (k < popcount), else 0
QK_HD uint32_t sel8_entry(uint32_t index) ;
// position of the k-th set bit of a byte (k < popcount), else 0
QK_HD uint32_t kth_in_byte(uint32_t byte, uint32_t k)
{
    for (uint32_t b = 0; b < 8; ++b)
        if ((byte >> b) & 1u) { if (k == 0) return b; --k; }
    return 0;
}

QK_HD uint32_t sel8_entry(uint32_t index) { return 16u * kth_in_byte(index >> 3, index & 7u); }

// ---------------------------------------------------------------- root record
// What k_prepare_roots writes and k_playout reloads at every refill (48 B = 3 x uint4).
struct __attribute__((aligned(16))) RootRec {
    uint32_t occ2, ra_lo, ra_hi, rb_lo;
    uint32_t rb_hi, ta_lo, ta_hi, tb_lo;
    uint32_t tb_hi, lp_lo, lp_hi, flags;
};
// flags: bit 0 = player to move at the root (role A); bits 8..9 = result known at
// the root without playing (0 = none, 1 = P0 wins, 2 = P1 wins); bits 16.. = pieces on board
enum { QK_ROOT_IMMEDIATE_SHIFT = 8 };

QK_HD uint32_t lanes_with_one(uint32_t u)
{
    return ((u & 0xFFFFu) ? 0x0000FFFFu : 0u) | ((u >> 16) ? 0xFFFF0000u : 0u);
}
QK_HD uint32_t line_presence2(uint32_t u)   // 12-bit line sets of the two lanes of u
{
    uint32_t r = 0;
    for (int p = 0; p < 16; ++p) {
        uint32_t e = square_lut(p).lines2;
        if ((u >> p) & 1u) r |= e & 0xFFFFu;
        if ((u >> (p + 16)) & 1u) r |= e & 0xFFFF0000u;
    }
    return r;
}

QK_HD RootRec make_root(const State4 &s)
{
    RootRec r;
    int player = 0;
    int code = validate(s, &player);
    int win = win_status(s);                 // checked first by both rollout loops
    uint32_t immediate = win ? (uint32_t)win : (code != 0 ? 2u : 0u);   // invalid: (0, {}) -> P0 cannot move
    uint32_t a_ab = player ? s.z : s.x, a_cd = player ? s.w : s.y;
    uint32_t b_ab = player ? s.x : s.z, b_cd = player ? s.y : s.w;
    uint32_t o = s.x | s.y | s.z | s.w;
    o = (o | (o >> 16)) & 0xFFFFu;
    r.occ2 = o * 0x10001u;
    r.ra_lo = scope_fill(b_ab) | lanes_with_two(a_ab);  r.ra_hi = scope_fill(b_cd) | lanes_with_two(a_cd);
    r.rb_lo = scope_fill(a_ab) | lanes_with_two(b_ab);  r.rb_hi = scope_fill(a_cd) | lanes_with_two(b_cd);
    r.ta_lo = lanes_with_one(a_ab); r.ta_hi = lanes_with_one(a_cd);
    r.tb_lo = lanes_with_one(b_ab); r.tb_hi = lanes_with_one(b_cd);
    r.lp_lo = line_presence2(s.x | s.z); r.lp_hi = line_presence2(s.y | s.w);
    r.flags = (uint32_t)player | (immediate << QK_ROOT_IMMEDIATE_SHIFT) | ((uint32_t)popc(o) << 16);
    return r;
}

// ---------------------------------------------------------------- one ply
struct PlayState {
    uint32_t free2, ra_lo, ra_hi, rb_lo, rb_hi, ta_lo, ta_hi, tb_lo, tb_hi, lp_lo, lp_hi;
};
QK_HD PlayState load_play_state(const RootRec &r)
{
    PlayState p = { ~r.occ2, r.ra_lo, r.ra_hi, r.rb_lo, r.rb_hi, r.ta_lo, r.ta_hi, r.tb_lo, r.tb_hi, r.lp_lo, r.lp_hi };
    return p;
}

// Plays one move for role ROLE (0 = A moves, 1 = B moves) using random word `rnd`.
// blocked: the mover had no legal move (mover loses; state untouched in a way that matters)
// win:     the move completed a line (mover wins)
// lut / sel8: the 16-entry square table and the 256x8 k-th-bit table (shared memory
// in the kernel, plain arrays in hostsim).  action (optional) = shape*16 + square.
template <int ROLE>
QK_HD void ply_step(PlayState &st, uint32_t rnd, const SquareLut *lut, const uint8_t *sel8,
                    bool &blocked, bool &win, int *action = nullptr)
{
    uint32_t &rm_lo = ROLE ? st.rb_lo : st.ra_lo, &rm_hi = ROLE ? st.rb_hi : st.ra_hi;   // mover
    uint32_t &ro_lo = ROLE ? st.ra_lo : st.rb_lo, &ro_hi = ROLE ? st.ra_hi : st.rb_hi;   // other
    uint32_t &tm_lo = ROLE ? st.tb_lo : st.ta_lo, &tm_hi = ROLE ? st.tb_hi : st.ta_hi;

    // generate_legal_moves: 64-bit action mask, bit = shape*16 + square
    uint32_t legal_lo = st.free2 & ~rm_lo, legal_hi = st.free2 & ~rm_hi;
    int nlo = popc(legal_lo), n = nlo + popc(legal_hi);
    blocked = (n == 0);

    // uniform pick: k-th legal action in ascending order
    int k = (int)mulhi(rnd, (uint32_t)n);
    bool up = k >= nlo;                                   // shapes C,D
    uint32_t w = up ? legal_hi : legal_lo;
    k -= up ? nlo : 0;
    int c = popc(w & 0xFFFFu);
    bool odd = k >= c;                                    // shapes B / D (upper lane)
    k -= odd ? c : 0;
    w = odd ? (w >> 16) : w;
    c = popc(w & 0xFFu);
    bool hib = k >= c;                                    // squares 8..15
    k -= hib ? c : 0;
    uint32_t byte = (hib ? (w >> 8) : w) & 0xFFu;
    // byte offset of the chosen square's entry in the SquareLut table (16 B per square)
    uint32_t off = (uint32_t)sel8[(byte << 3) | (uint32_t)(k & 7)] + (hib ? 128u : 0u);   // <= 240: always inside the table

    // apply_move + incremental rule state
    uint32_t lane = odd ? 0xFFFF0000u : 0x0000FFFFu;
    uint32_t l_lo = up ? 0u : lane, l_hi = up ? lane : 0u;
    const SquareLut e = *reinterpret_cast<const SquareLut *>(reinterpret_cast<const char *>(lut) + off);
    st.free2 &= ~e.bit2;
    ro_lo |= e.scope2 & l_lo;  ro_hi |= e.scope2 & l_hi;  // opponent may no longer use this shape in my scope
    rm_lo |= tm_lo & l_lo;     rm_hi |= tm_hi & l_hi;     // second piece of the shape: lane exhausted for me
    tm_lo |= l_lo;             tm_hi |= l_hi;
    st.lp_lo |= e.lines2 & l_lo; st.lp_hi |= e.lines2 & l_hi;

    // has_winning_line: a line holding all four shapes
    uint32_t all4 = st.lp_lo & st.lp_hi;
    win = (all4 & (all4 >> 16)) != 0;
    if (action) *action = (int)(off >> 4) + (odd ? 16 : 0) + (up ? 32 : 0);
}

// ---------------------------------------------------------------- tables of the predicated ply (ply_pred, below)
struct __attribute__((aligned(16))) MoveLut { uint32_t scope_lo, scope_hi, lines_lo, lines_hi, bit2, l_lo, l_hi, pad; };
// Shared-memory layout: two planes (entry words 0-3, words 4-7), each entry replicated into the eight 16-byte columns
// of a 128-byte row; a lane reads column lane % 8, so a 128-bit load of 32 different entries is conflict-free
// (four wavefronts, the minimum for 512 bytes) whatever the entries are.
enum { QK_MOVE_LUT_ENTRIES = 65, QK_MOVE_LUT_PLANE = QK_MOVE_LUT_ENTRIES * 128, QK_MOVE_LUT_BYTES = 2 * QK_MOVE_LUT_PLANE,
       QK_MOVE_LUT_BLOCKED_ROW = 4 * 1792 };
// word `wi` of entry `entry` = square | lane << 4 | half << 5; entry 64 = "the mover is blocked"
QK_HD uint32_t move_lut_word(uint32_t entry, uint32_t wi)
{
    if (entry >= 64u) return (wi == 2u || wi == 3u) ? 0x80008000u : 0u;
    const SquareLut q = square_lut((int)(entry & 15u));
    const uint32_t lane = (entry & 16u) ? 0xFFFF0000u : 0x0000FFFFu, hi = (entry >> 5) & 1u;
    const uint32_t v[8] = { hi ? 0u : q.scope2 & lane, hi ? q.scope2 & lane : 0u, hi ? 0u : q.lines2 & lane,
                            hi ? q.lines2 & lane : 0u, q.bit2, hi ? 0u : lane, hi ? lane : 0u, 0u };
    return v[wi & 7u];
}
// byte table of the predicated ply: row (= position) of the k-th set bit of a byte in the move table; the empty byte -- a
// mover without a legal action -- leads to the "blocked" entry (row 8 on top of the row offset every predicate adds)
QK_HD uint32_t sel8_row(uint32_t index)
{
    return (index >> 3) == 0u ? (64u * 128u - QK_MOVE_LUT_BLOCKED_ROW) / 128u : kth_in_byte(index >> 3, index & 7u);
}
// 32-bit word `w` of the shared-memory image of the move table
QK_HD uint32_t move_lut_image_word(uint32_t w)
{
    const uint32_t plane = w / (QK_MOVE_LUT_PLANE / 4u), rem = w % (QK_MOVE_LUT_PLANE / 4u);
    return move_lut_word(rem / 32u, (rem & 3u) + 4u * plane);
}

// C++ restatement of ply_pred (below): the same tables, the same three-level descent, the same sign test -- what the PTX
// computes, for the CPU test tier (tests/hostsim) and as the readable version of it.  `lut` = the shared-memory image
// of the move table (move_lut_image_word), `sel8r` = the byte table (sel8_row), `col` = the lane's
// column offset.  role 0: role A moves.
QK_HD void ply_pred_model(PlayState &st, int role, uint32_t rnd, const uint32_t *lut, const uint8_t *sel8r, uint32_t col,
                          uint32_t &act, uint32_t &acc)
{
    uint32_t &rm_lo = role ? st.rb_lo : st.ra_lo, &rm_hi = role ? st.rb_hi : st.ra_hi;   // mover
    uint32_t &ro_lo = role ? st.ra_lo : st.rb_lo, &ro_hi = role ? st.ra_hi : st.rb_hi;   // other
    uint32_t &tm_lo = role ? st.tb_lo : st.ta_lo, &tm_hi = role ? st.tb_hi : st.ta_hi;
    const uint32_t ll = st.free2 & ~rm_lo, lh = st.free2 & ~rm_hi;
    const uint32_t nl = (uint32_t)popc(ll), n = nl + (uint32_t)popc(lh);
    uint32_t k = mulhi(rnd, n), w = ll, row = col;
    if (k >= nl) { w = lh; row += 4096u; k -= nl; }                    // 64 -> 32
    uint32_t c = (uint32_t)popc(w << 16);
    if (k >= c) { w >>= 16; k -= c; row += 2048u; }                    // 32 -> 16
    c = (uint32_t)popc(w << 24);
    if (k >= c) { w >>= 8; k -= c; row += 1024u; }                     // 16 -> 8
    const uint32_t at = (uint32_t)sel8r[(w & 0xFFu) * 8u + k] * 128u + row;   // byte offset of the entry's first plane
    const uint32_t *e0 = lut + at / 4u, *e1 = lut + (at + QK_MOVE_LUT_PLANE) / 4u;
    ro_lo |= e0[0];            ro_hi |= e0[1];
    rm_lo |= tm_lo & e1[1];    rm_hi |= tm_hi & e1[2];
    tm_lo |= e1[1];            tm_hi |= e1[2];
    st.lp_lo |= e0[2];         st.lp_hi |= e0[3];
    st.free2 -= e1[0];
    const uint32_t a4 = st.lp_lo & st.lp_hi;
    const int32_t t = (int32_t)(a4 & (a4 << 16) & act);                // > 0: a line is complete, < 0: the mover was blocked
    if (role == 0 ? t > 0 : t < 0) ++acc;                              // role A wins
    if (t != 0) act = 0;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------- one ply, pipe-balanced (device only)
// Same function as ply_step, different instruction selection.  On sm_100 LOP3 / SEL / ISETP / SHF /
// IADD3 all issue to one 64-lane ALU pipe (0.5 warp-instructions per clock per scheduler) while IMAD
// goes to the FMA pipe, so a ply written with compares and selects is ALU-pipe bound at half the issue
// rate (ncu, profiles/r01b_playout_ncu.txt: ALU 72 %, FMA 15 %).  Here the popcount descent is integer
// multiply-add arithmetic instead: "k < c" is the sign bit of k - c (one SHF) and every conditional
// update is  x = flag * a + b  (one IMAD, 2 cycles on the FMA-heavy pipe; IMAD.HI / IMAD.WIDE cost 4,
// measured with qk_pipe_rates).  The split keeps both pipes near 75 cycles per warp-ply.
__device__ __forceinline__ uint32_t qk_mulhi(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t qk_mad(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t qk_madhi(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

template <int ROLE>
__device__ __forceinline__ void ply_step_fma(PlayState &st, uint32_t rnd, const SquareLut *lut, const uint8_t *sel8,
                                             bool &blocked, bool &win)
{
    uint32_t &rm_lo = ROLE ? st.rb_lo : st.ra_lo, &rm_hi = ROLE ? st.rb_hi : st.ra_hi;   // mover
    uint32_t &ro_lo = ROLE ? st.ra_lo : st.rb_lo, &ro_hi = ROLE ? st.ra_hi : st.rb_hi;   // other
    uint32_t &tm_lo = ROLE ? st.tb_lo : st.ta_lo, &tm_hi = ROLE ? st.tb_hi : st.ta_hi;

    const uint32_t legal_lo = st.free2 & ~rm_lo, legal_hi = st.free2 & ~rm_hi;
    const uint32_t nlo = (uint32_t)__popc(legal_lo), n = nlo + (uint32_t)__popc(legal_hi);
    blocked = (n == 0);
    uint32_t k = __umulhi(rnd, n);

    // 64 -> 32: lo if k < nlo
    uint32_t d = k - nlo;
    const uint32_t in_lo = d >> 31;                                // 1: shapes A,B   0: shapes C,D
    uint32_t w = qk_mad(in_lo, legal_lo - legal_hi, legal_hi);
    k = qk_mad(in_lo, nlo, d);
    // 32 -> 16: lower lane if k < c
    uint32_t c = (uint32_t)__popc(w << 16);
    d = k - c;
    const uint32_t even = d >> 31;                                 // 1: shape A / C   0: shape B / D
    k = qk_mad(even, c, d);
    w >>= qk_mad(even, 0xFFFFFFF0u, 16u);                          // 16 - 16 * even
    // 16 -> 8
    c = (uint32_t)__popc(w << 24);
    d = k - c;
    const uint32_t low8 = d >> 31;                                 // 1: squares 0..7
    k = qk_mad(low8, c, d);
    w >>= qk_mad(low8, 0xFFFFFFF8u, 8u);                           // 8 - 8 * low8
    // last three levels from the byte table; off = 16 * square
    const uint32_t idx = ((w & 0xFFu) << 3) | k;
    const uint32_t off = qk_mad(low8, 0xFFFFFF80u, 128u) + (uint32_t)sel8[idx];

    // lane masks: which 16-bit lane of which half receives the piece
    const uint32_t lane = qk_mad(even, 0x0000FFFFu - 0xFFFF0000u, 0xFFFF0000u);
    const uint32_t l_lo = qk_mad(in_lo, lane, 0u), l_hi = lane - l_lo;
    const SquareLut e = *reinterpret_cast<const SquareLut *>(reinterpret_cast<const char *>(lut) + off);
    st.free2 &= ~e.bit2;
    ro_lo |= e.scope2 & l_lo;  ro_hi |= e.scope2 & l_hi;
    rm_lo |= tm_lo & l_lo;     rm_hi |= tm_hi & l_hi;
    tm_lo |= l_lo;             tm_hi |= l_hi;
    st.lp_lo |= e.lines2 & l_lo; st.lp_hi |= e.lines2 & l_hi;
    const uint32_t all4 = st.lp_lo & st.lp_hi;
    win = (all4 & (all4 >> 16)) != 0;
}

// ---------------------------------------------------------------- one ply, predicated (device only)
// Third instruction selection of the same function.  The popcount descent is PREDICATED execution, written in
// PTX because the compiler turns the C++ form into compare + select pairs on the ALU pipe:
//       p = (k >= c);   @p k -= c;   @p w >>= 16;   @p move-table row += ...
// -- one ISETP and one SHF on the ALU pipe per level, the rest on the FMA pipe, where the multiply-add form above
// needs five instructions per level plus rematerialised constants.  The three predicates (which half / lane /
// byte holds the k-th legal action) only steer the ADDRESS of a 64-entry move table whose 32-byte entries are
// pre-masked for (half, lane, square): scope and line bits already routed to the lo / hi register, lane masks
// ready -- so the state update is eight unconditional LOP3s and no lane arithmetic.  A mover without a legal
// action (n = 0: k = 0 and all three predicates hold) lands, through the byte table's entries for the empty byte,
// on a 65th entry that changes nothing but sets bit 15 of all four shapes' line sets: "blocked" then comes out of
// the SAME test as a completed line, as the sign bit of  lines & lines << 16 & act  (act = all-ones while the
// game runs, zero afterwards).  Free squares are cleared by subtraction (the bit is set whenever the move is
// real) and the line test shifts LEFT: both forms the FMA pipe can take, which the ALU pipe -- the busier one,
// DESIGN.md section 5 -- cannot offload otherwise; `one` / `zero` are kernel parameters (so that ptxas keeps
// x = one * c + x  an IMAD instead of folding it back into an ALU-pipe add).
// `acc` counts the games role A wins (A = the root's mover): the mover of an even ply wins by completing a
// line, the mover of an odd ply makes A win by being blocked (beam_search.py:633-635,641-642).
// A lane whose game is over keeps executing plies on a meaningless state until the next 4-ply boundary:
// every quantity stays in range (k < popcount of the word it indexes), nothing is tallied (act == 0).
// (shared-memory address of the byte table, which is aligned to its 2 KB) / 8; the shared-memory address of the lane's
// column of the move table, and the same + the row offset of shapes C,D; the opaque 1 and 0
struct PredConsts { uint32_t sel_addr8, col_lo, col_hi, one, zero; };

// the move table and the byte table live in shared memory; ROLE 0: role A moves
// the move itself: pick + state update
template <int ROLE>
__device__ __forceinline__ void ply_pred_move(PlayState &st, uint32_t rnd, const PredConsts &pc)
{
    uint32_t &rm_lo = ROLE ? st.rb_lo : st.ra_lo, &rm_hi = ROLE ? st.rb_hi : st.ra_hi;   // mover
    uint32_t &ro_lo = ROLE ? st.ra_lo : st.rb_lo, &ro_hi = ROLE ? st.ra_hi : st.rb_hi;   // other
    uint32_t &tm_lo = ROLE ? st.tb_lo : st.ta_lo, &tm_hi = ROLE ? st.tb_hi : st.ta_hi;
    uint32_t e0x, e0y, e0z, e0w, e1x, e1y, e1z, e1w;
    asm("{\n"
        " .reg .pred p0, p1, p2;\n"
        " .reg .b32 ll, lh, nl, nh, n, k, w, t, c, row, ad;\n"
        " lop3.b32 ll, %8, %9, 0, 0x30;\n"              /* legal, shapes A,B: free & ~closed */
        " lop3.b32 lh, %8, %10, 0, 0x30;\n"             /* legal, shapes C,D */
        " popc.b32 nl, ll;\n"
        " popc.b32 nh, lh;\n"
        " add.u32 n, nl, nh;\n"
        " mul.hi.u32 k, %11, n;\n"                      /* uniform pick: k-th legal action, ascending */
        " setp.ge.u32 p0, k, nl;\n"                     /* 64 -> 32 */
        " selp.b32 w, lh, ll, p0;\n"
        " selp.b32 row, %15, %14, p0;\n"
        " @p0 sub.u32 k, k, nl;\n"
        " shl.b32 t, w, 16;\n"                          /* 32 -> 16 */
        " popc.b32 c, t;\n"
        " setp.ge.u32 p1, k, c;\n"
        " @p1 shr.u32 w, w, 16;\n"
        " @p1 sub.u32 k, k, c;\n"
        " @p1 mad.lo.u32 row, %13, 2048, row;\n"
        " shl.b32 t, w, 24;\n"                          /* 16 -> 8 */
        " popc.b32 c, t;\n"
        " setp.ge.u32 p2, k, c;\n"
        " @p2 shr.u32 w, w, 8;\n"
        " @p2 sub.u32 k, k, c;\n"
        " @p2 mad.lo.u32 row, %13, 1024, row;\n"
        " lop3.b32 t, w, 0xFF, %12, 0xEA;\n"            /* 8 -> 1 from the byte table: (byte | table / 8) * 8 + k */
        " mad.lo.u32 ad, t, 8, k;\n"
        " ld.shared.u8 ad, [ad];\n"
        " mad.lo.u32 ad, ad, 128, row;\n"              /* the entry: table row (position) on top of the row offset */
        " ld.shared.v4.u32 {%0, %1, %2, %3}, [ad];\n"
        " ld.shared.v4.u32 {%4, %5, %6, %7}, [ad+8320];\n"
        "}\n"
        : "=r"(e0x), "=r"(e0y), "=r"(e0z), "=r"(e0w), "=r"(e1x), "=r"(e1y), "=r"(e1z), "=r"(e1w)
        : "r"(st.free2), "r"(rm_lo), "r"(rm_hi), "r"(rnd), "r"(pc.sel_addr8), "r"(pc.one), "r"(pc.col_lo), "r"(pc.col_hi));
    static_assert(QK_MOVE_LUT_PLANE == 8320, "the second plane's offset is spelled out in the PTX above");
    ro_lo |= e0x;              ro_hi |= e0y;              // opponent may no longer use this shape in my scope
    rm_lo |= tm_lo & e1y;      rm_hi |= tm_hi & e1z;      // second piece of the shape: lane exhausted for me
    tm_lo |= e1y;              tm_hi |= e1z;
    st.lp_lo |= e0z;           st.lp_hi |= e0w;           // lines that now contain the shape
    st.free2 -= e1x;                                      // square no longer free
}
// > 0: the move completed a line (has_winning_line), < 0: there was no move, the mover is blocked; 0 otherwise and
// whenever the lane has no game running
__device__ __forceinline__ int32_t ply_pred_outcome(const PlayState &st, uint32_t act)
{
    const uint32_t a4 = st.lp_lo & st.lp_hi;
    return (int32_t)(a4 & (a4 << 16) & act);
}
template <int ROLE>
__device__ __forceinline__ void ply_pred(PlayState &st, uint32_t rnd, const PredConsts &pc, uint32_t &act, uint32_t &acc)
{
    ply_pred_move<ROLE>(st, rnd, pc);
    // has_winning_line (a line holding all four shapes) or blocked (the sign bit), while the game runs
    if (ROLE == 0)
        asm("{\n"
            " .reg .pred pf, pa;\n"
            " .reg .b32 a4, t;\n"
            " and.b32 a4, %2, %3;\n"
            " shl.b32 t, a4, 16;\n"
            " lop3.b32 t, a4, t, %0, 0x80;\n"
            " setp.ne.s32 pf, t, 0;\n"                  /* the game ends with this ply */
            " setp.gt.s32 pa, t, 0;\n"                  /* A completed a line */
            " @pa mad.lo.u32 %1, %4, %4, %1;\n"
            " @pf mul.lo.u32 %0, %0, %5;\n"
            "}\n"
            : "+r"(act), "+r"(acc)
            : "r"(st.lp_lo), "r"(st.lp_hi), "r"(pc.one), "r"(pc.zero));
    else
        asm("{\n"
            " .reg .pred pf, pa;\n"
            " .reg .b32 a4, t;\n"
            " and.b32 a4, %2, %3;\n"
            " shl.b32 t, a4, 16;\n"
            " lop3.b32 t, a4, t, %0, 0x80;\n"
            " setp.ne.s32 pf, t, 0;\n"
            " setp.lt.s32 pa, t, 0;\n"                  /* B cannot move: A wins */
            " @pa mad.lo.u32 %1, %4, %4, %1;\n"
            " @pf mul.lo.u32 %0, %0, %5;\n"
            "}\n"
            : "+r"(act), "+r"(acc)
            : "r"(st.lp_lo), "r"(st.lp_hi), "r"(pc.one), "r"(pc.zero));
}
#endif

// ---------------------------------------------------------------- scalar playout
// Whole playout on one thread, same ply_step as the kernel.  Used by
// k_playout's tail handling and by tests/hostsim.  Returns +1 / -1 / 0 from
// player 0's point of view (beam_search.py:635, mcts.py:417-419,437).
QK_HD int playout_one(const RootRec &root, uint64_t seed, uint64_t rollout, uint32_t root_index,
                      int max_plies, int *plies_out, const SquareLut *lut = nullptr, const uint8_t *sel8 = nullptr)
{
#ifndef __CUDA_ARCH__
    static SquareLut h_lut[16];
    static uint8_t h_sel8[2048];
    static bool ready = false;
    if (!ready) {
        for (int p = 0; p < 16; ++p) h_lut[p] = square_lut(p);
        for (uint32_t i = 0; i < 2048; ++i) h_sel8[i] = (uint8_t)sel8_entry(i);
        ready = true;
    }
    if (!lut) { lut = h_lut; sel8 = h_sel8; }
#endif
    int result = 0, ply = 0;
    uint32_t immediate = (root.flags >> QK_ROOT_IMMEDIATE_SHIFT) & 3u;
    int a_wins = (root.flags & 1u) ? -1 : 1;             // role A is the root's mover
    PlayState st = load_play_state(root);
    Philox4 rnd = { { 0, 0, 0, 0 } };
    for (;;) {
        if (max_plies >= 0 && ply >= max_plies) { result = 0; break; }      // mcts.py:412,436-437
        if (ply == 0 && immediate) { result = immediate == 1 ? 1 : -1; break; }
        if ((ply & 3) == 0) rnd = stream_block(seed, rollout, root_index, (uint32_t)(ply >> 2));
        bool blocked, win;
        if (ply & 1) ply_step<1>(st, rnd.v[ply & 3], lut, sel8, blocked, win);
        else         ply_step<0>(st, rnd.v[ply & 3], lut, sel8, blocked, win);
        int mover_wins = (ply & 1) ? -a_wins : a_wins;
        if (blocked) { result = -mover_wins; break; }                       // beam :641-642, mcts :427-429
        ++ply;
        if (max_plies >= 0 && ply >= max_plies) { result = 0; break; }      // cut-off precedes the win test
        if (win) { result = mover_wins; break; }
    }
    if (plies_out) *plies_out = ply;
    return result;
}

}  // namespace qk
