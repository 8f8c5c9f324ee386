// qk_perft_core.cuh -- bulk counting of the last perft ply.
//
// For a non-won, valid position ("leaf parent") the enumerator of
// game_stats.py:384-424 needs, per parent: how many legal moves it has
// (total_moves, :401) and how many of them complete a line (:411-415, credited to
// the mover).  Neither needs the children to be materialised:
//   moves = popc(legal mask)
//   a placement of shape s on square p wins  <=>  p is legal for s and some line
//   through p already holds the three OTHER shapes
// so winning placements = popc(legal_s & (rows|cols|zones holding all shapes but s)).
#pragma once
#include "qk_bits.cuh"

namespace qk {

#ifdef __CUDACC__
__device__ __forceinline__ State4 to_state4(const uint4 &v) { State4 s = { v.x, v.y, v.z, v.w }; return s; }
#endif
QK_HD uint32_t swap_lanes(uint32_t v) { return prmt(v, 0, 0x1032); }

// which rows / zones / columns hold each shape, colours merged: one bit per line in each 16-bit lane
// (rows at bit 4r, zones at bits 0, 2, 8, 10, columns at bits 0..3); lanes (A|B<<16, C|D<<16).
// A new piece of shape s on square p adds bit (p & 12) / (p & 10) / (p & 3) to lane s of the r / z / c word.
struct Presence { uint32_t r01, r23, z01, z23, c01, c23; };

QK_HD Presence line_presence(const State4 &s)
{
    const uint32_t u01 = s.x | s.z, u23 = s.y | s.w;
    const uint32_t t01 = pair_or1(u01), t23 = pair_or1(u23);
    Presence p = { row_presence(t01), row_presence(t23), zone_presence(t01), zone_presence(t23),
                   col_presence(u01), col_presence(u23) };
    return p;
}

// squares where dropping shape s would complete a line; lanes (A|B<<16, C|D<<16)
QK_HD void winning_from_presence(const Presence &p, uint32_t &win_ab, uint32_t &win_cd)
{
    // "some piece of shape t shares this row / zone / column" as square masks
    uint32_t r01 = p.r01 * 0xFu, r23 = p.r23 * 0xFu;
    uint32_t z01 = p.z01 * 0x33u, z23 = p.z23 * 0x33u;
    uint32_t c01 = p.c01 * 0x1111u, c23 = p.c23 * 0x1111u;
    // x01 & both(x23) has, in lane A, has(A) & has(C) & has(D) -- the squares where the missing shape B wins -- and
    // in lane B the squares where A wins: the masks come out with their lanes swapped, so ONE swap of the combined
    // mask replaces a swap of each of the three ingredients.
    // both lanes of x & swap(x) hold lane0 & lane1: the AND of the word's two shapes, already replicated
    uint32_t rb01 = r01 & swap_lanes(r01), rb23 = r23 & swap_lanes(r23);
    uint32_t zb01 = z01 & swap_lanes(z01), zb23 = z23 & swap_lanes(z23);
    uint32_t cb01 = c01 & swap_lanes(c01), cb23 = c23 & swap_lanes(c23);
    win_ab = swap_lanes((r01 & rb23) | (z01 & zb23) | (c01 & cb23));
    win_cd = swap_lanes((r23 & rb01) | (z23 & zb01) | (c23 & cb01));
}

QK_HD void winning_squares(const State4 &s, uint32_t &win_ab, uint32_t &win_cd)
{
    winning_from_presence(line_presence(s), win_ab, win_cd);
}

// per-square masks for the evaluation of a leaf parent from its parent's data, replicated in both 16-bit lanes
struct __attribute__((aligned(16))) TileLut { uint32_t scope2, bit2, row2, zone2, col2, pad0, pad1, pad2; };
QK_HD TileLut tile_lut(int p)
{
    const uint32_t scope = ((0xFu << (p & 12)) | (0x1111u << (p & 3)) | (0x33u << (p & 10))) & 0xFFFFu;
    TileLut t = { scope * 0x10001u, 0x00010001u << p, 0x00010001u << (p & 12), 0x00010001u << (p & 10),
                  0x00010001u << (p & 3), 0u, 0u, 0u };
    return t;
}
// move / winning-move counts of a leaf parent from its PARENT: (legal_lo, legal_hi) = what its mover could have
// played one ply earlier (legal_masks of the parent for the player NOT to move there), `pres` = the parent's line
// presence, `action` = the move that was made.  The new piece takes its square away from every shape and its scope
// away from its own shape (move.py:169-196), and adds its shape to its row, zone and column; the child itself is
// never materialised.
QK_HD void leaf_counts_from_parent(uint32_t legal_lo, uint32_t legal_hi, Presence pres, int action, const TileLut *lut,
                                   int &n_moves, int &n_wins)
{
    const TileLut e = lut[action & 15];
    const uint32_t lane = (action & 16) ? 0xFFFF0000u : 0x0000FFFFu;
    const uint32_t in_lo = (action & 32) ? 0u : lane, in_hi = lane ^ in_lo;                 // lane of the new piece's shape
    const uint32_t lo = legal_lo & ~(e.bit2 | (e.scope2 & in_lo)), hi = legal_hi & ~(e.bit2 | (e.scope2 & in_hi));
    pres.r01 |= e.row2 & in_lo; pres.r23 |= e.row2 & in_hi;
    pres.z01 |= e.zone2 & in_lo; pres.z23 |= e.zone2 & in_hi;
    pres.c01 |= e.col2 & in_lo; pres.c23 |= e.col2 & in_hi;
    uint32_t wab, wcd;
    winning_from_presence(pres, wab, wcd);
    n_moves = popc(lo) + popc(hi);
    n_wins = popc(lo & wab) + popc(hi & wcd);
}

// The same counts for the children of ONE shape of a node, with everything that does not depend on the square hoisted:
// a child of shape S changes the line presence of S's word pair only (A,B live in r01/z01/c01, C,D in r23/z23/c23), so
// the other pair's expanded masks and its "both shapes" masks are loop invariants, and the lane of S inside its word is
// a compile-time mask.  About 40 instructions per child against 83 for leaf_counts_from_parent in a loop
// (k_perft_tile: a lane evaluating its own children).
template <int S> struct ShapeLoop {
    static constexpr uint32_t kLane = (S & 1) ? 0xFFFF0000u : 0x0000FFFFu;
    uint32_t legal_own, legal_oth;            // what the child's mover could play one ply earlier: the pair holding S / the other pair
    uint32_t r_own, z_own, c_own;             // presence words of S's pair
    uint32_t R_oth, Z_oth, C_oth;             // the other pair's presence, expanded to squares
    uint32_t rb_oth, zb_oth, cb_oth;          // ... and where BOTH its shapes are present, replicated in both lanes
    QK_HD ShapeLoop(uint32_t legal_lo, uint32_t legal_hi, const Presence &p)
    {
        legal_own = S < 2 ? legal_lo : legal_hi; legal_oth = S < 2 ? legal_hi : legal_lo;
        r_own = S < 2 ? p.r01 : p.r23; z_own = S < 2 ? p.z01 : p.z23; c_own = S < 2 ? p.c01 : p.c23;
        R_oth = (S < 2 ? p.r23 : p.r01) * 0xFu; Z_oth = (S < 2 ? p.z23 : p.z01) * 0x33u; C_oth = (S < 2 ? p.c23 : p.c01) * 0x1111u;
        rb_oth = R_oth & swap_lanes(R_oth); zb_oth = Z_oth & swap_lanes(Z_oth); cb_oth = C_oth & swap_lanes(C_oth);
    }
    // the child that puts shape S on the square of table entry e
    QK_HD void child(const TileLut &e, int &n_moves, int &n_wins) const
    {
        const uint32_t own = legal_own & ~(e.bit2 | (e.scope2 & kLane)), oth = legal_oth & ~e.bit2;
        const uint32_t R = (r_own | (e.row2 & kLane)) * 0xFu, Z = (z_own | (e.zone2 & kLane)) * 0x33u, C = (c_own | (e.col2 & kLane)) * 0x1111u;
        // winning squares as in winning_from_presence: own pair needs both shapes of the other pair plus its partner lane
        const uint32_t win_own = swap_lanes((R & rb_oth) | (Z & zb_oth) | (C & cb_oth));
        const uint32_t win_oth = swap_lanes((R_oth & R & swap_lanes(R)) | (Z_oth & Z & swap_lanes(Z)) | (C_oth & C & swap_lanes(C)));
        n_moves = popc(own) + popc(oth);
        n_wins = popc(own & win_own) + popc(oth & win_oth);
    }
};

// same, when the caller already knows who is to move (depth parity): saves the four popcounts of piece_balance
QK_HD void leaf_parent_counts_for(const State4 &s, int mover, int &n_moves, int &n_wins)
{
    uint32_t lo, hi, wab, wcd;
    legal_masks(s, mover, lo, hi);
    winning_squares(s, wab, wcd);
    n_moves = popc(lo) + popc(hi);
    n_wins = popc(lo & wab) + popc(hi & wcd);
}

// state must be valid and not already won.  Returns the mover; n_moves / n_wins
// as described above.
QK_HD int leaf_parent_counts(const State4 &s, int &n_moves, int &n_wins)
{
    int mover = piece_balance(s);           // 0 -> P0 to move, 1 -> P1 (state_validator.py:82-98)
    uint32_t lo, hi, wab, wcd;
    legal_masks(s, mover, lo, hi);
    winning_squares(s, wab, wcd);
    n_moves = popc(lo) + popc(hi);
    n_wins = popc(lo & wab) + popc(hi & wcd);
    return mover;
}

}  // namespace qk
