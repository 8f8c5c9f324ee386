// qk_bits.cuh -- Quantik rules on a register-resident 16-byte state (sm_100a).
//
// The state is the reference's `<8H` payload (commons.py:10-13,
// memory/bitboard_compact.py:36-37) held as one uint4:
//     x = P0.A | P0.B << 16      y = P0.C | P0.D << 16
//     z = P1.A | P1.B << 16      w = P1.C | P1.D << 16
// bit i of each 16-bit lane = square i, row-major 4x4.  All rule evaluation is
// SWAR arithmetic on these four registers: two boards per 32-bit ALU op, no
// loops over squares or lines.
//
// Every function is __host__ __device__ so that tests/hostsim can execute the
// exact same code on the CPU against the oracle before any GPU time is spent;
// the product only ever runs the __device__ instantiation.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define QK_HD __host__ __device__ __forceinline__
#else
#define QK_HD inline
#endif

namespace qk {

struct State4 { uint32_t x, y, z, w; };

// ---------------------------------------------------------------- intrinsics
QK_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
#ifdef __CUDA_ARCH__
    return __byte_perm(a, b, sel);
#else
    uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
#endif
}
QK_HD int popc(uint32_t v)
{
#ifdef __CUDA_ARCH__
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
QK_HD uint32_t brev(uint32_t v)
{
#ifdef __CUDA_ARCH__
    return __brev(v);
#else
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
    v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
    return (v >> 16) | (v << 16);
#endif
}
QK_HD uint32_t mulhi(uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// ---------------------------------------------------------------- right shift, optionally on the FMA pipe
// The canonicalisation and perft kernels are bound by the 64-lane ALU pipe (ncu: 96 % / 90 % busy, FMA pipe
// 6 % / 14 %).  A right shift by a constant is also  mul.hi(x, 2^(32-s)) , which issues to the FMA-heavy pipe
// (IMAD.HI; the multiplier must be opaque to ptxas, hence the __constant__ table).  MEASURED AND REJECTED
// (round 1): with all 48 transform shifts moved, canonical keys/s stayed at 4.49e10 and perft(6) went from
// 3.65 to 3.89 ms -- IMAD.HI is half rate (qk_pipe_rates) and the ALU pipe was not relieved enough to pay
// for it.  Kept behind QK_IMADHI_SHIFTS for the record; the default is the plain shift.
#if defined(__CUDACC__) && defined(QK_IMADHI_SHIFTS)
__constant__ static uint32_t qk_pow2[32] = {
    1u << 0, 1u << 1, 1u << 2, 1u << 3, 1u << 4, 1u << 5, 1u << 6, 1u << 7, 1u << 8, 1u << 9, 1u << 10, 1u << 11, 1u << 12,
    1u << 13, 1u << 14, 1u << 15, 1u << 16, 1u << 17, 1u << 18, 1u << 19, 1u << 20, 1u << 21, 1u << 22, 1u << 23, 1u << 24,
    1u << 25, 1u << 26, 1u << 27, 1u << 28, 1u << 29, 1u << 30, 1u << 31 };
#endif
template <int S> QK_HD uint32_t shr(uint32_t x)
{
#if defined(__CUDA_ARCH__) && defined(QK_IMADHI_SHIFTS)
    return __umulhi(x, qk_pow2[32 - S]);
#else
    return x >> S;
#endif
}

// ---------------------------------------------------------------- line fills
// For a pair of boards packed in one register: every square that shares a row /
// column / 2x2 zone with a set square.  (move.py:169-196 tests each square
// against the 12 WIN_MASKS of commons.py:24-48; this is the closed form.)
QK_HD uint32_t pair_or1(uint32_t u) { return u | shr<1>(u); }
QK_HD uint32_t row_presence(uint32_t t1) { return (t1 | shr<2>(t1)) & 0x11111111u; }   // bit 4r   of each lane
QK_HD uint32_t zone_presence(uint32_t t1) { return (t1 | shr<4>(t1)) & 0x05050505u; }  // bits 0,2,8,10
QK_HD uint32_t col_presence(uint32_t u) { uint32_t t = u | shr<4>(u); return (t | shr<8>(t)) & 0x000F000Fu; }  // bits 0..3

QK_HD uint32_t scope_fill(uint32_t u)
{
    uint32_t t1 = pair_or1(u);
    return row_presence(t1) * 0xFu | zone_presence(t1) * 0x33u | col_presence(u) * 0x1111u;
}

// 0xFFFF in every 16-bit lane that holds two or more set bits
// (MAX_PIECES_PER_SHAPE = 2, commons.py:20; move.py:247-249).
QK_HD uint32_t lanes_with_two(uint32_t u)
{
    // both lanes at once.  (u | H) - L never borrows across lanes, and u & that is non-zero in a lane exactly when
    // the lane holds >= 2 bits (bit 15 survives only next to another bit; the low bits lose their lowest one).
    const uint32_t H = 0x80008000u, L = 0x00010001u;
    const uint32_t t = u & ((u | H) - L);
    // non-zero lane -> bit 15 -> 0/1 -> 0xFFFF (the last step is an IMAD)
    const uint32_t nz = (((t & ~H) + ~H) | t) & H;
    return (nz >> 15) * 0xFFFFu;
}

// ---------------------------------------------------------------- win test
// has_winning_line (game_utils.py:123-158): some row, column or zone holds all
// four shapes, colours ignored.
QK_HD bool has_winning_line(const State4 &s)
{
    uint32_t u01 = s.x | s.z, u23 = s.y | s.w;           // shape unions A|B<<16, C|D<<16
    uint32_t t01 = pair_or1(u01), t23 = pair_or1(u23);
    uint32_t r = row_presence(t01) & row_presence(t23);
    uint32_t zn = zone_presence(t01) & zone_presence(t23);
    uint32_t c = col_presence(u01) & col_presence(u23);
    return (((r & (r >> 16)) | (zn & (zn >> 16)) | (c & (c >> 16))) & 0xFFFFu) != 0;
}

QK_HD int piece_balance(const State4 &s) { return popc(s.x) + popc(s.y) - popc(s.z) - popc(s.w); }

// check_game_winner (game_utils.py:161-189): 0 none, 1 P0, 2 P1 -- the winner is
// inferred from the piece counts (total0 > total1 ? P0 : P1).
QK_HD int win_status(const State4 &s)
{
    if (!has_winning_line(s)) return 0;
    return piece_balance(s) > 0 ? 1 : 2;
}

// ---------------------------------------------------------------- validation
// _validate_game_state_single_pass (state_validator.py:126-168), including the
// order in which the first failing check is reported: words 0..7, per word
// count>2 (2) before overlap-with-earlier-words (5); then balance (1); then an
// opposing pair of the same shape on one line (4).
// Returns the ValidationResult code; *player = side to move (0/1) when OK.
QK_HD int validate(const State4 &s, int *player)
{
    const uint32_t w[4] = { s.x, s.y, s.z, s.w };
    uint32_t seen = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t v = (w[i >> 1] >> ((i & 1) * 16)) & 0xFFFFu;
        if (popc(v) > 2) return 2;
        if (seen & v) return 5;
        seen |= v;
    }
    int diff = piece_balance(s);
    if (diff != 0 && diff != 1) return 1;
    // a P0 piece and a P1 piece of one shape share a line  <=>  P1's piece sits in
    // the scope fill of P0's pieces of that shape
    if ((scope_fill(s.x) & s.z) | (scope_fill(s.y) & s.w)) return 4;
    *player = diff;
    return 0;
}

// ---------------------------------------------------------------- move generation
// generate_legal_moves (move.py:199-268) as the 64-bit action mask of
// api_portability_report.py:75-86 (bit = shape*16 + square): for each shape the
// mover still holds (<2 on board), the empty squares outside the row/column/zone
// scope of the opponent's pieces of that shape.  Caller guarantees a valid state.
QK_HD void legal_masks(const State4 &s, int player, uint32_t &lo, uint32_t &hi)
{
    uint32_t own_ab = player ? s.z : s.x, own_cd = player ? s.w : s.y;
    uint32_t opp_ab = player ? s.x : s.z, opp_cd = player ? s.y : s.w;
    uint32_t o = s.x | s.y | s.z | s.w;
    o = (o | (o >> 16)) & 0xFFFFu;
    o |= o << 16;
    lo = ~(o | scope_fill(opp_ab) | lanes_with_two(own_ab));
    hi = ~(o | scope_fill(opp_cd) | lanes_with_two(own_cd));
}

// movegen with the reference's treatment of bad input: invalid state -> player 0,
// empty mask (move.py:223-225).
QK_HD uint64_t movegen(const State4 &s, int *player, int *status)
{
    int p = 0;
    int st = validate(s, &p);
    *status = st;
    if (st != 0) { *player = 0; return 0; }
    *player = p;
    uint32_t lo, hi;
    legal_masks(s, p, lo, hi);
    return ((uint64_t)hi << 32) | lo;
}

// apply_move (move.py:135-166): OR one bit into word player*4 + shape.
QK_HD State4 apply_action(State4 s, int player, int action)
{
    uint32_t bit = 1u << ((action & 15) | (action & 16));   // square + 16*(shape&1)
    bool cd = (action & 32) != 0;
    if (player == 0) { if (cd) s.y |= bit; else s.x |= bit; }
    else             { if (cd) s.w |= bit; else s.z |= bit; }
    return s;
}

// ---------------------------------------------------------------- D4 board symmetries
// permute16 (symmetry.py:159-250) without the 8 x 65536 LUT: every D4 element is a
// product of three bit-reversal style generators, applied to both lanes at once.
QK_HD uint32_t flip_cols(uint32_t v)   // reflV: (r,c) -> (r,3-c): reverse bits inside nibbles
{
    v = ((v & 0x55555555u) << 1) | (shr<1>(v) & 0x55555555u);
    return ((v & 0x33333333u) << 2) | (shr<2>(v) & 0x33333333u);
}
QK_HD uint32_t flip_rows(uint32_t v)   // reflH: (r,c) -> (3-r,c): reverse nibbles inside lanes
{
    v = ((v & 0x0F0F0F0Fu) << 4) | (shr<4>(v) & 0x0F0F0F0Fu);
    return prmt(v, 0, 0x2301);
}
// reflH WITHOUT its final byte swap inside the lanes: the canonicalisation only ever feeds such an image to
// byte-selecting PRMTs (sorted_image<true>), which absorb the pending swap for free.
QK_HD uint32_t flip_rows_pending(uint32_t v) { return ((v & 0x0F0F0F0Fu) << 4) | ((v >> 4) & 0x0F0F0F0Fu); }
QK_HD uint32_t transpose(uint32_t v)   // reflD: (r,c) -> (c,r)
{
    uint32_t t = (v ^ shr<3>(v)) & 0x0A0A0A0Au;
    v ^= t ^ (t << 3);
    t = (v ^ shr<6>(v)) & 0x00CC00CCu;
    return v ^ t ^ (t << 6);
}
// rot180: (r,c) -> (3-r,3-c) reverses the 16 bits of each lane = BREV (XU pipe, idle in these kernels) + lane swap
QK_HD uint32_t rot180(uint32_t v) { return prmt(brev(v), 0, 0x1032); }
// d4 index as in symmetry.py:101-110: id rot90 rot180 rot270 reflV reflH reflD reflAD
QK_HD uint32_t d4_apply(uint32_t v, int d4)
{
    switch (d4) {
    case 0: return v;
    case 1: return flip_cols(transpose(v));
    case 2: return rot180(v);
    case 3: return flip_rows(transpose(v));
    case 4: return flip_cols(v);
    case 5: return flip_rows(v);
    case 6: return transpose(v);
    default: return rot180(transpose(v));
    }
}
QK_HD State4 d4_apply(const State4 &s, int d4)
{
    State4 r = { d4_apply(s.x, d4), d4_apply(s.y, d4), d4_apply(s.z, d4), d4_apply(s.w, d4) };
    return r;
}

// ---------------------------------------------------------------- canonical form
// find_canonical_form (symmetry.py:289-354) = lexicographic minimum of the packed
// little-endian payload over 8 D4 maps x 24 shape relabelings (x colour swap).
// The same relabeling acts on both colours and the payload is [C0[pi], C1[pi]], so
// the minimum over the 24 relabelings of one board image is that image with its
// four shapes STABLY SORTED by the key  bswap16(C0[s]) << 16 | bswap16(C1[s])
// (bytes-lex order of an `<H` word = numeric order of its byte-swapped value).
// 192 candidates collapse to 8 four-element sorts (16 with colour swap).
struct Key128 { uint32_t a, b, c, d; };   // big-endian limbs of the byte-string being compared

QK_HD bool key_less(const Key128 &p, const Key128 &q)
{
    uint64_t ph = ((uint64_t)p.a << 32) | p.b, qh = ((uint64_t)q.a << 32) | q.b;
    uint64_t pl = ((uint64_t)p.c << 32) | p.d, ql = ((uint64_t)q.c << 32) | q.d;
    return ph < qh || (ph == qh && pl < ql);
}
// all-ones when p < q as 128-bit numbers, else 0: one borrow chain (4 subtracts) instead of 64-bit compare pairs
QK_HD uint32_t key_less_mask(const Key128 &p, const Key128 &q)
{
#ifdef __CUDA_ARCH__
    uint32_t m;
    asm("{\n\t.reg .u32 t;\n\t"
        "sub.cc.u32 t, %1, %5;\n\t"
        "subc.cc.u32 t, %2, %6;\n\t"
        "subc.cc.u32 t, %3, %7;\n\t"
        "subc.cc.u32 t, %4, %8;\n\t"
        "subc.u32 %0, 0, 0;\n\t}"
        : "=r"(m) : "r"(p.d), "r"(p.c), "r"(p.b), "r"(p.a), "r"(q.d), "r"(q.c), "r"(q.b), "r"(q.a));
    return m;
#else
    return key_less(p, q) ? 0xFFFFFFFFu : 0u;
#endif
}
QK_HD void key_min(Key128 &best, const Key128 &k)
{
#ifdef __CUDA_ARCH__
    // 0/1 flag from the borrow chain, then  best += flag * (k - best)  per limb: the selects run on the FMA pipe
    // (IMAD), which these ALU-bound kernels leave almost idle
    uint32_t f;
    asm("{\n\t.reg .u32 t;\n\t"
        "sub.cc.u32 t, %1, %5;\n\t"
        "subc.cc.u32 t, %2, %6;\n\t"
        "subc.cc.u32 t, %3, %7;\n\t"
        "subc.cc.u32 t, %4, %8;\n\t"
        "subc.u32 %0, 0, 0;\n\t}"
        : "=r"(f) : "r"(k.d), "r"(k.c), "r"(k.b), "r"(k.a), "r"(best.d), "r"(best.c), "r"(best.b), "r"(best.a));
    f = 0u - f;                                  // all-ones -> 1
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(best.a) : "r"(f), "r"(k.a - best.a));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(best.b) : "r"(f), "r"(k.b - best.b));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(best.c) : "r"(f), "r"(k.c - best.c));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(best.d) : "r"(f), "r"(k.d - best.d));
#else
    const uint32_t m = key_less_mask(k, best);
    best.a = (k.a & m) | (best.a & ~m); best.b = (k.b & m) | (best.b & ~m);
    best.c = (k.c & m) | (best.c & ~m); best.d = (k.d & m) | (best.d & ~m);
#endif
}
QK_HD bool key_equal(const Key128 &p, const Key128 &q) { return p.a == q.a && p.b == q.b && p.c == q.c && p.d == q.d; }

// compare-exchange.  On the device the maximum is  p + q - min  computed with two integer multiply-adds whose
// multipliers (1 and -1) come from constant memory, so that ptxas cannot fold them back into one IADD3: the
// canonicalisation kernels are bound by the ALU pipe (min/max, PRMT, LOP3, shifts and IADD3 all issue there: ncu
// 89 % busy with the FMA-heavy pipe at 33 %), and this moves one of the two VIMNMX of every comparator across.
#ifdef __CUDACC__
static __constant__ uint32_t qk_c_one = 1u, qk_c_minus_one = 0xFFFFFFFFu;
#endif
QK_HD void cswap(uint32_t &p, uint32_t &q)
{
#if defined(__CUDA_ARCH__) && !defined(QK_CSWAP_ALU)
    const uint32_t lo = p < q ? p : q;
    const uint32_t sum = p * qk_c_one + q;
    q = lo * qk_c_minus_one + sum;
    p = lo;
#else
    uint32_t lo = p < q ? p : q, hi = p < q ? q : p; p = lo; q = hi;
#endif
}

// sorted-shape arrangement of one board image (c0 = first colour's two registers).  PENDING = the registers
// still owe a byte swap inside each lane (flip_rows_pending): only the PRMT selectors change.
template <bool PENDING = false>
QK_HD Key128 sorted_image(uint32_t c0ab, uint32_t c0cd, uint32_t c1ab, uint32_t c1cd)
{
    uint32_t k0 = prmt(c0ab, c1ab, PENDING ? 0x1054 : 0x0145), k1 = prmt(c0ab, c1ab, PENDING ? 0x3276 : 0x2367);
    uint32_t k2 = prmt(c0cd, c1cd, PENDING ? 0x1054 : 0x0145), k3 = prmt(c0cd, c1cd, PENDING ? 0x3276 : 0x2367);
    cswap(k0, k1); cswap(k2, k3); cswap(k0, k2); cswap(k1, k3); cswap(k1, k2);
    Key128 r = { prmt(k0, k1, 0x3276), prmt(k2, k3, 0x3276), prmt(k0, k1, 0x1054), prmt(k2, k3, 0x1054) };
    return r;
}
QK_HD State4 key_to_state(const Key128 &k)
{
    State4 s = { prmt(k.a, 0, 0x0123), prmt(k.b, 0, 0x0123), prmt(k.c, 0, 0x0123), prmt(k.d, 0, 0x0123) };
    return s;
}

template <bool COLOR_SWAP>
QK_HD State4 canonical(const State4 &s)
{
    Key128 best;
    bool have = false;
    // Walk D4 with 20 ops per register instead of 50:
    //   id, rot180 = BREV(id) | H' = reflH with its byte swap pending, BREV(H') = reflV |
    //   T = reflD, BREV(T) = reflAD | H'(T) = rot270, BREV(H'(T)) = rot90.
    // BREV (XU pipe, otherwise idle here) reverses each lane = rot180 and also swaps the two lanes of a register,
    // i.e. relabels shapes A<->B and C<->D in both colours at once -- irrelevant, the four shape keys get sorted.
    // The pending byte swap of H' survives BREV as a pending byte swap and is absorbed by sorted_image<true>.
    State4 g = s;
#define QK_CANDIDATE(PENDING)                                                   \
    {                                                                           \
        Key128 k = sorted_image<PENDING>(g.x, g.y, g.z, g.w);                   \
        if (!have) { best = k; have = true; } else key_min(best, k);            \
        if (COLOR_SWAP) key_min(best, sorted_image<PENDING>(g.z, g.w, g.x, g.y)); \
    }
#define QK_STEP(F) { g.x = F(g.x); g.y = F(g.y); g.z = F(g.z); g.w = F(g.w); }
    QK_CANDIDATE(false);                              // id
    QK_STEP(brev);               QK_CANDIDATE(false); // rot180
    g = s;
    QK_STEP(flip_rows_pending);  QK_CANDIDATE(true);  // reflH
    QK_STEP(brev);               QK_CANDIDATE(true);  // reflV
    g = s;
    QK_STEP(transpose);          QK_CANDIDATE(false); // reflD
    const State4 t = g;
    QK_STEP(brev);               QK_CANDIDATE(false); // reflAD
    g = t;
    QK_STEP(flip_rows_pending);  QK_CANDIDATE(true);  // rot270
    QK_STEP(brev);               QK_CANDIDATE(true);  // rot90
#undef QK_STEP
#undef QK_CANDIDATE
    return key_to_state(best);
}

// Canonical form PLUS the transform the reference would report
// (symmetry.py:344-349: first strict minimum in d4 -> colour -> itertools
// permutation order).  transform = d4 | cs << 3 | perm[0..3] << (4 + 2 i).
// Slow path (used for move mapping, not for bulk keys).
QK_HD State4 canonical_with_transform(const State4 &s, bool color_swap, uint32_t *transform)
{
    Key128 best = { 0, 0, 0, 0 };
    int best_d4 = -1, best_cs = 0;
    for (int d4 = 0; d4 < 8; ++d4) {
        State4 g = d4_apply(s, d4);
        for (int cs = 0; cs <= (color_swap ? 1 : 0); ++cs) {
            Key128 k = cs ? sorted_image(g.z, g.w, g.x, g.y) : sorted_image(g.x, g.y, g.z, g.w);
            if (best_d4 < 0 || key_less(k, best)) { best = k; best_d4 = d4; best_cs = cs; }
        }
    }
    // recover the stable sorting permutation of the winning image: perm[i] = source
    // shape placed at slot i; ties keep ascending shape order = first permutation in
    // itertools.permutations order.
    State4 g = d4_apply(s, best_d4);
    uint32_t c0ab = best_cs ? g.z : g.x, c0cd = best_cs ? g.w : g.y, c1ab = best_cs ? g.x : g.z, c1cd = best_cs ? g.y : g.w;
    uint32_t key[4] = { prmt(c0ab, c1ab, 0x0145), prmt(c0ab, c1ab, 0x2367), prmt(c0cd, c1cd, 0x0145), prmt(c0cd, c1cd, 0x2367) };
    uint32_t t = (uint32_t)best_d4 | ((uint32_t)best_cs << 3);
    for (int sidx = 0; sidx < 4; ++sidx) {
        int rank = 0;
        for (int o = 0; o < 4; ++o) rank += (key[o] < key[sidx]) || (key[o] == key[sidx] && o < sidx);
        t |= (uint32_t)sidx << (4 + 2 * rank);
    }
    *transform = t;
    return key_to_state(best);
}

// count_orbit_size (symmetry.py:356-396): distinct images among the 8 x 24.
// By orbit-stabiliser it is 192 / |Stab|, and |Stab| = (#D4 maps whose sorted image
// equals the sorted original) x (prod over groups of identical shapes of size!).
QK_HD int orbit_size(const State4 &s)
{
    Key128 base = sorted_image(s.x, s.y, s.z, s.w);
    int fix = 0;
    for (int d4 = 0; d4 < 8; ++d4) {
        State4 g = d4_apply(s, d4);
        fix += key_equal(sorted_image(g.x, g.y, g.z, g.w), base) ? 1 : 0;
    }
    // sorted keys back out of `base`: k_i = (hi16 from a/b) << 16 | lo16 from c/d
    uint32_t k0 = (base.a & 0xFFFF0000u) | (base.c >> 16), k1 = (base.a << 16) | (base.c & 0xFFFFu);
    uint32_t k2 = (base.b & 0xFFFF0000u) | (base.d >> 16), k3 = (base.b << 16) | (base.d & 0xFFFFu);
    int m = 1, run = 1;
    run = (k1 == k0) ? run + 1 : 1; m *= run;
    run = (k2 == k1) ? run + 1 : 1; m *= run;
    run = (k3 == k2) ? run + 1 : 1; m *= run;
    return 192 / (fix * m);
}

// apply_symmetry (symmetry.py:252-287): out[c][i] = C_c[perm[i]] after the D4 map
// and the optional colour swap.  transform encoding as above.
QK_HD State4 apply_symmetry(const State4 &s, uint32_t transform)
{
    State4 g = d4_apply(s, (int)(transform & 7));
    if (transform & 8) { State4 t = { g.z, g.w, g.x, g.y }; g = t; }
    uint32_t c0[4] = { g.x & 0xFFFFu, g.x >> 16, g.y & 0xFFFFu, g.y >> 16 };
    uint32_t c1[4] = { g.z & 0xFFFFu, g.z >> 16, g.w & 0xFFFFu, g.w >> 16 };
    uint32_t p[4];
    for (int i = 0; i < 4; ++i) p[i] = (transform >> (4 + 2 * i)) & 3;
    State4 r;
    r.x = c0[p[0]] | (c0[p[1]] << 16); r.y = c0[p[2]] | (c0[p[3]] << 16);
    r.z = c1[p[0]] | (c1[p[1]] << 16); r.w = c1[p[2]] | (c1[p[3]] << 16);
    return r;
}

// ---------------------------------------------------------------- k-th set bit
// index of the k-th (0-based) set bit of a 64-bit mask given as two halves,
// k < popc(lo)+popc(hi).  Branch-free popcount descent.
QK_HD int select_bit64(uint32_t lo, uint32_t hi, int nlo, int k)
{
    bool up = k >= nlo;
    uint32_t w = up ? hi : lo;
    k -= up ? nlo : 0;
    int pos = up ? 32 : 0;
    int c = popc(w & 0xFFFFu); bool ge = k >= c; k -= ge ? c : 0; w = ge ? w >> 16 : w; pos += ge ? 16 : 0;
    c = popc(w & 0xFFu);       ge = k >= c;      k -= ge ? c : 0; w = ge ? w >> 8 : w;  pos += ge ? 8 : 0;
    c = popc(w & 0xFu);        ge = k >= c;      k -= ge ? c : 0; w = ge ? w >> 4 : w;  pos += ge ? 4 : 0;
    c = popc(w & 0x3u);        ge = k >= c;      k -= ge ? c : 0; w = ge ? w >> 2 : w;  pos += ge ? 2 : 0;
    pos += (k >= (int)(w & 1u)) ? 1 : 0;
    return pos;
}

}  // namespace qk
