// qk_dedup.cu -- canonicalise + merge equal canonical payloads + sum multiplicities:
// the dictionary step of game_stats.py:487-505, beam_search.py:526-546 and
// examples/generate_opening_book.py:458-496, done in HBM.
//
//   k_dedup_insert   one state per thread: canonical form in registers, then insert into
//                    an open-addressing table keyed by the 16-byte payload with a single
//                    128-bit compare-and-swap (atom.cas.b128, sm_90+), multiplicity added
//                    with a 64-bit atomic.
//   k_dedup_compact  table -> dense (key, multiplicity) arrays, one cursor atomic per 1 024 slots
//   k_bitonic_step   sorts the distinct keys in payload byte order (deterministic output)
#include "qk_common.cuh"
#include <math.h>
#include "qk_bits.cuh"
#include "qk_perft_core.cuh"

namespace qk {

static constexpr unsigned kBlock = 256;
typedef unsigned long long u64;

struct __align__(16) Slot { u64 hi, lo; };     // Key128 (a:b, c:d); all ones = empty
// One table entry = one 32-byte sector: the key and its multiplicity travel together, so the probe's load,
// the claim's CAS and the multiplicity's RED all hit the same sector (the table is far larger than L2 and the
// probes are random: with the counts in a second array every insert cost two DRAM row activations).
struct __align__(32) TSlot { u64 hi, lo, count, pad; };
static constexpr size_t kMaxProbe = 8192;      // longer chains mean the table is as good as full

__device__ __forceinline__ void cas128(void *addr, u64 cmp_hi, u64 cmp_lo, u64 new_hi, u64 new_lo, u64 &old_hi, u64 &old_lo)
{
    // memory layout of Slot is {hi, lo}; the b128 register packs {low 64 bits, high 64 bits}
    asm volatile(
        "{\n\t"
        ".reg .b128 c, s, o;\n\t"
        "mov.b128 c, {%2, %3};\n\t"
        "mov.b128 s, {%4, %5};\n\t"
        "atom.global.cas.b128 o, [%6], c, s;\n\t"
        "mov.b128 {%0, %1}, o;\n\t"
        "}\n"
        : "=l"(old_hi), "=l"(old_lo)
        : "l"(cmp_hi), "l"(cmp_lo), "l"(new_hi), "l"(new_lo), "l"(addr)
        : "memory");
}

__device__ __forceinline__ u64 mix(u64 h, u64 l)
{
    u64 x = h * 0x9E3779B97F4A7C15ull ^ (l + 0xD6E8FEB86659FD93ull) * 0xC2B2AE3D27D4EB4Full;
    x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
    return x;
}

// empty table: key halves all ones, multiplicity and pad zero.  Written as 16-byte halves, consecutive lanes to
// consecutive halves (one 32-byte entry per lane -- two 16-byte stores 32 bytes apart -- left every store instruction
// with half-written sectors: 3.8 TB/s; this form writes at 7.2 TB/s, the rate of a plain fill)
__global__ void __launch_bounds__(kBlock) k_dedup_fill(TSlot *table, size_t slots)
{
    uint4 *halves = reinterpret_cast<uint4 *>(table);
    const size_t n = 2 * slots;
    const uint4 key = make_uint4(~0u, ~0u, ~0u, ~0u), zero = make_uint4(0u, 0u, 0u, 0u);
    const size_t stride = (size_t)gridDim.x * kBlock;            // even: the parity of a thread's halves never changes
    const bool odd = threadIdx.x & 1u;
    for (size_t i = blockIdx.x * (size_t)kBlock + threadIdx.x; i < n; i += stride) halves[i] = odd ? zero : key;
}

// Adds multiplicity m to key (khi_new, klo_new), creating the entry when it is new, probing linearly from `slot`.
// -> true when THIS call created it.  Distinct keys are counted by the callers in registers and added to the global
// counter once per warp at the end of the kernel (flush_claims): a per-insert atomic on one address serialises in
// one L2 slice.
__device__ __forceinline__ bool table_insert_at(TSlot *table, size_t slots, size_t slot, u64 khi_new, u64 klo_new, u64 m, int *overflow)
{
    const size_t max_probe = slots - 1 < kMaxProbe ? slots - 1 : kMaxProbe;
    for (size_t probe = 0; probe <= max_probe; ++probe, slot = slot + 1 == slots ? 0 : slot + 1) {
        // fast path: one 128-bit load; a slot caught half-way through another thread's
        // CAS can only look like "some half is all ones", which falls through to the CAS
        const ulonglong2 raw = __ldcv(reinterpret_cast<const ulonglong2 *>(table + slot));
        u64 khi = raw.x, klo = raw.y;
        bool claimed = false;
        if (khi == ~0ull || klo == ~0ull) {
            cas128(table + slot, ~0ull, ~0ull, khi_new, klo_new, khi, klo);
            if (khi == ~0ull && klo == ~0ull) { claimed = true; khi = khi_new; klo = klo_new; }
        }
        if (khi == khi_new && klo == klo_new) { atomicAdd(&table[slot].count, m); return claimed; }   // game_stats.py:487-505
    }
    *overflow = 1;      // no room within kMaxProbe slots: the host retries with a larger table
    return false;
}
__device__ __forceinline__ bool table_insert(TSlot *table, size_t slots, u64 khi_new, u64 klo_new, u64 m,
                                             u64 *all_ones_count, int *overflow)
{
    if ((khi_new & klo_new) == ~0ull) { atomicAdd(all_ones_count, m); return false; }   // the payload that equals the empty marker
    // any table size (not only powers of two: the table is filled and scanned once per call, so its size is time)
    return table_insert_at(table, slots, (size_t)__umul64hi(mix(khi_new, klo_new), (u64)slots), khi_new, klo_new, m, overflow);
}

// every lane of the warp must call this (after its loops): one atomic per warp and kernel
__device__ __forceinline__ void flush_claims(unsigned claims, u64 *n_unique, int *overflow, u64 max_unique)
{
    const unsigned total = __reduce_add_sync(0xFFFFFFFFu, claims);
    if ((threadIdx.x & 31u) == 0 && total) {
        const u64 before = atomicAdd(n_unique, (u64)total);
        if (before + total > max_unique) *overflow = 1;
    }
}

template <bool COLOR_SWAP>
__global__ void __launch_bounds__(kBlock) k_dedup_insert(const uint4 *__restrict__ states, const u64 *__restrict__ mult, size_t n,
                                                          TSlot *table, size_t slots,
                                                          u64 *n_unique, u64 *all_ones_count, int *overflow, u64 max_unique)
{
    unsigned claims = 0;
    for (size_t i = blockIdx.x * (size_t)kBlock + threadIdx.x; i < n; i += (size_t)gridDim.x * kBlock) {
        if (*reinterpret_cast<volatile int *>(overflow)) break;     // table too small: give up early
        uint4 v = __ldg(states + i);
        State4 s = { v.x, v.y, v.z, v.w };
        State4 c = canonical<COLOR_SWAP>(s);
        // comparable big-endian limbs of the payload bytes
        u64 hi = ((u64)prmt(c.x, 0, 0x0123) << 32) | prmt(c.y, 0, 0x0123);
        u64 lo = ((u64)prmt(c.z, 0, 0x0123) << 32) | prmt(c.w, 0, 0x0123);
        claims += table_insert(table, slots, hi, lo, mult ? mult[i] : 1ull, all_ones_count, overflow) ? 1u : 0u;
    }
    flush_claims(claims, n_unique, overflow, max_unique);
}

__device__ __forceinline__ TSlot load_entry(const TSlot *table, size_t i)
{
    const ulonglong4 e = *reinterpret_cast<const ulonglong4 *>(table + i);
    TSlot t = { e.x, e.y, e.z, e.w };
    return t;
}

// table -> dense output.  A block takes runs of kBlock x kCompactItems slots, counts the run's entries and reserves
// their place with ONE atomic.  (One reservation per warp made the single cursor the bottleneck -- 2.6 M atomics on
// one address for the largest BFS level: 5.1 ms for a 4.8 GB table that is read in 0.85 ms.)  `slots` is a
// multiple of 256.  EMIT: straight into the caller's payload / multiplicity arrays (unsorted output), else into
// the (key, multiplicity) arrays the sort works on.
static constexpr int kCompactItems = 4;
template <bool EMIT>
__global__ void __launch_bounds__(kBlock) k_dedup_compact(const TSlot *table, size_t slots, Slot *keys, u64 *mults,
                                                           uint4 *uniq, u64 *uniq_mult, u64 *cursor)
{
    __shared__ unsigned s_warp[kBlock / 32];
    __shared__ u64 s_base;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (size_t run = blockIdx.x * (size_t)(kBlock * kCompactItems); run < slots; run += (size_t)gridDim.x * kBlock * kCompactItems) {
        TSlot e[kCompactItems];
        unsigned mask[kCompactItems], mine = 0;
#pragma unroll
        for (int k = 0; k < kCompactItems; ++k) {
            const size_t i = run + (size_t)k * kBlock + threadIdx.x;
            e[k].hi = ~0ull; e[k].lo = ~0ull; e[k].count = 0;
            if (i < slots) e[k] = load_entry(table, i);
            mask[k] = __ballot_sync(0xFFFFFFFFu, !(e[k].hi == ~0ull && e[k].lo == ~0ull));
            mine += (unsigned)__popc(mask[k]);
        }
        if (lane == 0) s_warp[warp] = mine;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned total = 0;
            for (unsigned w = 0; w < kBlock / 32; ++w) { const unsigned c = s_warp[w]; s_warp[w] = total; total += c; }
            s_base = total ? atomicAdd(cursor, (u64)total) : 0ull;
        }
        __syncthreads();
        u64 o = s_base + s_warp[warp];
#pragma unroll
        for (int k = 0; k < kCompactItems; ++k) {
            if ((mask[k] >> lane) & 1u) {
                const u64 at = o + (u64)__popc(mask[k] & ((1u << lane) - 1u));
                if (EMIT) {
                    uniq[at] = make_uint4(prmt((uint32_t)(e[k].hi >> 32), 0, 0x0123), prmt((uint32_t)e[k].hi, 0, 0x0123),
                                          prmt((uint32_t)(e[k].lo >> 32), 0, 0x0123), prmt((uint32_t)e[k].lo, 0, 0x0123));
                    uniq_mult[at] = e[k].count;
                } else {
                    keys[at].hi = e[k].hi; keys[at].lo = e[k].lo; mults[at] = e[k].count;
                }
            }
            o += (u64)__popc(mask[k]);
        }
        __syncthreads();
    }
}
static unsigned compact_grid(size_t slots, int sms) { return grid_for((slots + kCompactItems - 1) / kCompactItems, kBlock, sms, 8); }

__global__ void __launch_bounds__(kBlock) k_pad(Slot *keys, u64 *mults, size_t from, size_t to)
{
    for (size_t i = from + blockIdx.x * (size_t)kBlock + threadIdx.x; i < to; i += (size_t)gridDim.x * kBlock) {
        keys[i].hi = ~0ull; keys[i].lo = ~0ull; mults[i] = 0;
    }
}

__global__ void __launch_bounds__(kBlock) k_bitonic_step(Slot *keys, u64 *mults, size_t n, size_t j, size_t k)
{
    for (size_t i = blockIdx.x * (size_t)kBlock + threadIdx.x; i < n; i += (size_t)gridDim.x * kBlock) {
        size_t p = i ^ j;
        if (p <= i) continue;
        Slot a = keys[i], b = keys[p];
        bool a_gt_b = a.hi > b.hi || (a.hi == b.hi && a.lo > b.lo);
        bool up = (i & k) == 0;
        if (a_gt_b == up) {
            keys[i] = b; keys[p] = a;
            u64 t = mults[i]; mults[i] = mults[p]; mults[p] = t;
        }
    }
}

// every compare-exchange step of distance < kTile of the stages k_lo..k_hi, on a tile held in shared memory
// (one global round trip instead of up to 55)
static constexpr unsigned kTile = 1024;
__global__ void __launch_bounds__(kBlock) k_bitonic_tile(Slot *keys, u64 *mults, size_t n, size_t k_lo, size_t k_hi)
{
    __shared__ Slot s_key[kTile];
    __shared__ u64 s_mult[kTile];
    for (size_t base = blockIdx.x * (size_t)kTile; base < n; base += (size_t)gridDim.x * kTile) {
        for (unsigned t = threadIdx.x; t < kTile; t += kBlock) { s_key[t] = keys[base + t]; s_mult[t] = mults[base + t]; }
        __syncthreads();
        for (size_t k = k_lo; k <= k_hi; k <<= 1) {
            unsigned j = (k >> 1) < (kTile >> 1) ? (unsigned)(k >> 1) : (kTile >> 1);
            for (; j > 0; j >>= 1) {
                for (unsigned t = threadIdx.x; t < kTile / 2; t += kBlock) {
                    const unsigned i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), p = i | j;
                    const Slot a = s_key[i], b = s_key[p];
                    const bool a_gt_b = a.hi > b.hi || (a.hi == b.hi && a.lo > b.lo);
                    const bool up = ((base + i) & k) == 0;
                    if (a_gt_b == up) {
                        s_key[i] = b; s_key[p] = a;
                        const u64 m = s_mult[i]; s_mult[i] = s_mult[p]; s_mult[p] = m;
                    }
                }
                __syncthreads();
            }
        }
        for (unsigned t = threadIdx.x; t < kTile; t += kBlock) { keys[base + t] = s_key[t]; mults[base + t] = s_mult[t]; }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kBlock) k_emit(const Slot *keys, const u64 *mults, size_t n, uint4 *uniq, u64 *uniq_mult)
{
    for (size_t i = blockIdx.x * (size_t)kBlock + threadIdx.x; i < n; i += (size_t)gridDim.x * kBlock) {
        Slot k = keys[i];
        if (uniq) uniq[i] = make_uint4(prmt((uint32_t)(k.hi >> 32), 0, 0x0123), prmt((uint32_t)k.hi, 0, 0x0123),
                                       prmt((uint32_t)(k.lo >> 32), 0, 0x0123), prmt((uint32_t)k.lo, 0, 0x0123));
        if (uniq_mult) uniq_mult[i] = mults[i];
    }
}

// ---------------------------------------------------------------------------------------------
// One BFS level of game_stats.SymmetryTable._process_depth (game_stats.py:359-485), fused: expand the
// parents, drop line-completing moves, canonicalise the other children in registers and insert them into
// the table with the parent's multiplicity.  Children are never written to HBM.
// stats: [0] total_moves  [1] line wins by P0  [2] by P1  [3] parents with P0 blocked  [4] P1 blocked
// Lanes first play "one parent each" (masks, counts), then the children of the 32 parents are dealt out
// evenly: every lane canonicalises and inserts one child per round, so the expensive part (330 instructions
// of canonicalisation, the random table probe) runs with full warps and 32 probes in flight per warp --
// a warp-per-parent version kept ~8 of 32 lanes busy (late-game parents have few children).
struct BfsWarpScratch {
    uint4 state[32];               // the tile's parents
    u64 mult[32];
    unsigned short queue[2048];    // (parent lane << 6 | action) of every non-winning child of the tile
    unsigned char mover[32];
};

template <bool COLOR_SWAP>
__global__ void __launch_bounds__(kBlock) k_bfs_expand_insert(const uint4 *__restrict__ parents, const u64 *__restrict__ mult,
                                                               size_t n, TSlot *table, size_t slots,
                                                               u64 *n_unique, u64 *all_ones_count, int *overflow,
                                                               u64 max_unique, u64 *stats)
{
    __shared__ BfsWarpScratch s_scratch[kBlock / 32];
    BfsWarpScratch &ws = s_scratch[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u;
    const size_t warp = (blockIdx.x * (size_t)kBlock + threadIdx.x) >> 5, n_warps = ((size_t)gridDim.x * kBlock) >> 5;
    u64 s_moves = 0, s_w0 = 0, s_w1 = 0, s_b0 = 0, s_b1 = 0;      // per lane
    unsigned claims = 0;
    for (size_t tile = warp * 32; tile < n; tile += n_warps * 32) {
        if (__any_sync(0xFFFFFFFFu, *reinterpret_cast<volatile int *>(overflow))) break;   // table too small: the host retries
        // ---- one parent per lane: moves, winning moves, the children worth keeping
        uint32_t nl = 0, nh = 0;
        const size_t i = tile + lane;
        if (i < n) {
            const uint4 v = __ldg(parents + i);
            const State4 parent = { v.x, v.y, v.z, v.w };
            const u64 m = mult ? mult[i] : 1ull;
            const int mover = piece_balance(parent);
            uint32_t lo, hi, wab, wcd;
            legal_masks(parent, mover, lo, hi);
            const int n_moves = popc(lo) + popc(hi);
            if (n_moves == 0) { if (mover == 0) s_b0 += m; else s_b1 += m; }           // game_stats.py:426-443
            else {
                winning_squares(parent, wab, wcd);
                s_moves += (u64)n_moves * m;                                            // :401
                const u64 wins = (u64)(popc(lo & wab) + popc(hi & wcd)) * m;            // :411-415
                if (mover == 0) s_w0 += wins; else s_w1 += wins;
                nl = lo & ~wab; nh = hi & ~wcd;
            }
            ws.state[lane] = v; ws.mult[lane] = m; ws.mover[lane] = (unsigned char)mover;
        }
        // ---- deal the children out: exclusive scan of the counts, then every lane lists its own
        const uint32_t cnt = (uint32_t)(popc(nl) + popc(nh));
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (unsigned)o) incl += t; }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        uint32_t at = incl - cnt;
        for (uint32_t w = nl; w; w &= w - 1) ws.queue[at++] = (unsigned short)((lane << 6) | (uint32_t)(__ffs((int)w) - 1));
        for (uint32_t w = nh; w; w &= w - 1) ws.queue[at++] = (unsigned short)((lane << 6) | (uint32_t)(32 + __ffs((int)w) - 1));
        __syncwarp();
        // ---- full warps: one child per lane per round, never written to HBM
        for (uint32_t base = 0; base < total; base += 32) {
            const uint32_t idx = base + lane;
            if (idx < total) {
                const uint32_t e = ws.queue[idx], pl = e >> 6;
                const uint4 pv = ws.state[pl];
                const State4 parent = { pv.x, pv.y, pv.z, pv.w };
                const State4 c = canonical<COLOR_SWAP>(apply_action(parent, ws.mover[pl], (int)(e & 63u)));   // :453,:468
                claims += table_insert(table, slots, ((u64)prmt(c.x, 0, 0x0123) << 32) | prmt(c.y, 0, 0x0123),
                                       ((u64)prmt(c.z, 0, 0x0123) << 32) | prmt(c.w, 0, 0x0123), ws.mult[pl], all_ones_count,
                                       overflow) ? 1u : 0u;
            }
        }
        __syncwarp();
    }
    flush_claims(claims, n_unique, overflow, max_unique);
    // per-lane sums -> one atomic per warp and counter
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s_moves += __shfl_xor_sync(0xFFFFFFFFu, s_moves, o); s_w0 += __shfl_xor_sync(0xFFFFFFFFu, s_w0, o);
        s_w1 += __shfl_xor_sync(0xFFFFFFFFu, s_w1, o); s_b0 += __shfl_xor_sync(0xFFFFFFFFu, s_b0, o);
        s_b1 += __shfl_xor_sync(0xFFFFFFFFu, s_b1, o);
    }
    if (lane == 0) {
        if (s_moves) atomicAdd(stats + 0, s_moves);
        if (s_w0) atomicAdd(stats + 1, s_w0);
        if (s_w1) atomicAdd(stats + 2, s_w1);
        if (s_b0) atomicAdd(stats + 3, s_b0);
        if (s_b1) atomicAdd(stats + 4, s_b1);
    }
}

// ---------------------------------------------------------------------------------------------
// One depth of the opening-book BFS (examples/generate_opening_book.py:185-280,458-496) with the
// counting of opening_book_summary.build_summary (opening_book_summary.py:36-93): positions are
// canonical classes INCLUDING finished ones; a finished position (line complete, or mover without a
// move) is terminal and not expanded; an edge is a distinct (parent class, child class) pair.
// Warp per parent; the distinct children of a parent are found with match.any on both key halves
// (lanes l / l+32 hold children l / l+32; the second batch is checked against the first through
// shared memory).  Only one lane per distinct child inserts, so the stored multiplicity of a class is
// its in-degree in position_edges.    stats: [0] terminal parents  [1] edges
template <bool COLOR_SWAP>
__global__ void __launch_bounds__(kBlock) k_book_expand_insert(const uint4 *__restrict__ parents, size_t n, TSlot *table,
                                                                size_t slots, u64 *n_unique, u64 *all_ones_count, int *overflow,
                                                                u64 max_unique, u64 *stats)
{
    __shared__ Slot s_first[kBlock / 32][32];
    Slot *first = s_first[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u;
    const size_t warp = (blockIdx.x * (size_t)kBlock + threadIdx.x) >> 5, n_warps = ((size_t)gridDim.x * kBlock) >> 5;
    u64 s_terminal = 0, s_edges = 0;
    unsigned claims = 0;
    for (size_t i = warp; i < n; i += n_warps) {
        if (__any_sync(0xFFFFFFFFu, *reinterpret_cast<volatile int *>(overflow))) break;
        const uint4 v = __ldg(parents + i);
        const State4 parent = { v.x, v.y, v.z, v.w };
        if (has_winning_line(parent)) { ++s_terminal; continue; }                       // generate_opening_book.py:205-211
        const int mover = piece_balance(parent);
        uint32_t lo, hi;
        legal_masks(parent, mover, lo, hi);
        if ((lo | hi) == 0) { ++s_terminal; continue; }                                 // :212-223
        Slot key[2];
        bool leader[2];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const bool has = ((half ? hi : lo) >> lane) & 1u;
            key[half].hi = ~0ull; key[half].lo = ~0ull;
            leader[half] = false;
            const unsigned mask = __ballot_sync(0xFFFFFFFFu, has);
            if (has) {
                const State4 c = canonical<COLOR_SWAP>(apply_action(parent, mover, half * 32 + (int)lane));   // :237-242
                key[half].hi = ((u64)prmt(c.x, 0, 0x0123) << 32) | prmt(c.y, 0, 0x0123);
                key[half].lo = ((u64)prmt(c.z, 0, 0x0123) << 32) | prmt(c.w, 0, 0x0123);
                const unsigned same = __match_any_sync(mask, key[half].hi) & __match_any_sync(mask, key[half].lo);
                leader[half] = (unsigned)(__ffs((int)same) - 1) == lane;
            }
            if (half == 0) { first[lane] = key[0]; __syncwarp(); }
        }
        if (leader[1])                                                   // already among the first 32 children?
            for (int j = 0; j < 32; ++j)
                if (first[j].hi == key[1].hi && first[j].lo == key[1].lo) { leader[1] = false; break; }
        __syncwarp();
        s_edges += (u64)(__popc(__ballot_sync(0xFFFFFFFFu, leader[0])) + __popc(__ballot_sync(0xFFFFFFFFu, leader[1])));
#pragma unroll
        for (int half = 0; half < 2; ++half)
            if (leader[half])
                claims += table_insert(table, slots, key[half].hi, key[half].lo, 1ull, all_ones_count, overflow) ? 1u : 0u;
    }
    flush_claims(claims, n_unique, overflow, max_unique);
    if (lane == 0) {
        if (s_terminal) atomicAdd(stats + 0, s_terminal);
        if (s_edges) atomicAdd(stats + 1, s_edges);
    }
}

// ---------------------------------------------------------------------------------------------
// Two-level insert.  A table for tens of millions of classes is far larger than the 126 MB L2, and every insert
// into it is a random DRAM sector: read for the probe, written back after the multiplicity's RED, read again by the
// next duplicate (ncu, profiles/r01g_*: 80 B of DRAM reads per key against 24 B of key + multiplicity).  Here the
// keys are first ROUTED by the high hash bits into `parts` buckets (pass A: streaming writes), then the buckets are
// merged one at a time into ONE table that is small enough to stay in L2 and is reused for every bucket
// (pass B: the probes, the CAS and the RED are L2 hits; the table never travels to DRAM), and each bucket's classes
// are appended to the output as soon as it is merged.  DRAM sees every key twice, sequentially.
//   Buckets are cut per producing block -- sub-bucket (part, block) is a private range with its cursor in shared
// memory -- so routing needs no global atomics; a block's share of a part is binomial, the ranges are sized at the
// mean + 6 sigma, and an overflowing range sends the whole call back to the one-level path.
//   What it buys (profiles/r02_atomics_rate.txt, r02_insert_rate.txt): random probe / RED / 128-bit CAS run at
// 2.2 / 1.8 / 1.0 x 10^11 per second on an L2-resident table against 4.1 / 2.2 / 2.2 x 10^10 on a 1 GiB one, and an
// insert is a probe, a RED and (first arrival, or a duplicate racing it) a CAS: about 2.7 x 10^10 inserts/s in L2.
// The merge is bound by the L2's atomic units either way, not by DRAM bytes; the two-level form is 1.5x the
// one-level one on 2^24 keys, and is NOT used for the BFS expansion, whose 2.7 x 10^10 children/s already equal it.
static constexpr int kMaxParts = 1024;
// one part's table.  Measured on 2^24 keys (scripts/dedup_case.py, whole call): 24 MiB parts 2.01 ms, 40 MiB 1.64 ms,
// 64 MiB 1.36 ms, 96 MiB 1.23 ms, 112 MiB 1.15 ms, 120-128 MiB 1.14 ms (one-level table: 1.85 ms) -- fewer, larger parts
// win as long as the table still fits the 126 MB L2 next to the bucket streams.
static constexpr size_t kL2TableBytes = 112u << 20;
static size_t l2_table_bytes() { const long long mb = option(OPT_DEDUP_PART_MB); return mb > 0 ? (size_t)mb << 20 : kL2TableBytes; }

__device__ __forceinline__ void route(u64 hi, u64 lo, int parts, uint32_t part_slots, int &part, uint32_t &slot)
{
    const u64 h = mix(hi, lo);
    part = (int)__umulhi((uint32_t)(h >> 32), (uint32_t)parts);
    slot = __umulhi((uint32_t)h, part_slots);
}

struct Buckets {
    Slot *keys;              // [parts][blocks][sub_cap]
    u64 *mults;              // same shape, or null when every multiplicity is 1
    unsigned *counts;        // [parts][blocks]
    int parts;
    unsigned blocks;
    uint32_t sub_cap, part_slots;
};

__device__ __forceinline__ void bucket_put(const Buckets &b, unsigned *s_cur, u64 hi, u64 lo, u64 m, u64 *all_ones_count, int *overflow)
{
    if ((hi & lo) == ~0ull) { atomicAdd(all_ones_count, m); return; }       // the payload that equals the empty marker
    int part;
    uint32_t slot;
    route(hi, lo, b.parts, b.part_slots, part, slot);
    const unsigned pos = atomicAdd(&s_cur[part], 1u);
    if (pos >= b.sub_cap) { *overflow = 1; return; }
    const size_t at = ((size_t)part * b.blocks + blockIdx.x) * b.sub_cap + pos;
    b.keys[at].hi = hi; b.keys[at].lo = lo;
    if (b.mults) b.mults[at] = m;
}

template <bool COLOR_SWAP>
__global__ void __launch_bounds__(kBlock) k_bucket_states(const uint4 *__restrict__ states, const u64 *__restrict__ mult, size_t n,
                                                           size_t per_block, Buckets b, u64 *all_ones_count, int *overflow)
{
    extern __shared__ unsigned s_cur[];
    for (int k = threadIdx.x; k < b.parts; k += kBlock) s_cur[k] = 0;
    __syncthreads();
    const size_t begin = blockIdx.x * per_block, end = begin + per_block < n ? begin + per_block : n;
    for (size_t i = begin + threadIdx.x; i < end; i += kBlock) {
        const uint4 v = __ldg(states + i);
        const State4 st = { v.x, v.y, v.z, v.w };
        const State4 c = canonical<COLOR_SWAP>(st);
        bucket_put(b, s_cur, ((u64)prmt(c.x, 0, 0x0123) << 32) | prmt(c.y, 0, 0x0123), ((u64)prmt(c.z, 0, 0x0123) << 32) | prmt(c.w, 0, 0x0123),
                   mult ? mult[i] : 1ull, all_ones_count, overflow);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < b.parts; k += kBlock)
        b.counts[(size_t)k * b.blocks + blockIdx.x] = s_cur[k] < b.sub_cap ? s_cur[k] : b.sub_cap;
}

// ---- the BFS level in two-level form (k_bfs_expand_insert's table is DRAM-sized at the large levels -- 4.8 GB for 55 M
// classes -- and every child is a random DRAM sector there: 3.7 x 10^10 random reads per second on this machine,
// profiles/r02n_atomics_sweep.txt, against 2 x 10^11 on an L2-resident table).  Same plan as the dedup above: a
// count pass (how many children each block of parents will emit: the sub-bucket ranges are sized from it), the
// expansion with the children ROUTED into (part, block) ranges instead of inserted, then merge_buckets.
__global__ void __launch_bounds__(kBlock) k_bfs_count_children(const uint4 *__restrict__ parents, size_t n, size_t per_block,
                                                                unsigned *__restrict__ block_children)
{
    const size_t begin = blockIdx.x * per_block, end = begin + per_block < n ? begin + per_block : n;
    unsigned cnt = 0;
    for (size_t i = begin + threadIdx.x; i < end; i += kBlock) {
        const uint4 v = __ldg(parents + i);
        const State4 parent = { v.x, v.y, v.z, v.w };
        uint32_t lo, hi, wab, wcd;
        legal_masks(parent, piece_balance(parent), lo, hi);
        winning_squares(parent, wab, wcd);
        cnt += (unsigned)(popc(lo & ~wab) + popc(hi & ~wcd));
    }
    cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
    if ((threadIdx.x & 31u) == 0 && cnt) atomicAdd(block_children + blockIdx.x, cnt);
}

template <bool COLOR_SWAP>
__global__ void __launch_bounds__(kBlock) k_bfs_expand_bucket(const uint4 *__restrict__ parents, const u64 *__restrict__ mult,
                                                               size_t n, size_t per_block, Buckets b, u64 *all_ones_count,
                                                               int *overflow, u64 *stats)
{
    __shared__ BfsWarpScratch s_scratch[kBlock / 32];
    extern __shared__ unsigned s_cur[];
    for (int k = threadIdx.x; k < b.parts; k += kBlock) s_cur[k] = 0;
    __syncthreads();
    BfsWarpScratch &ws = s_scratch[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u;
    const size_t begin = blockIdx.x * per_block, end = begin + per_block < n ? begin + per_block : n;
    u64 s_moves = 0, s_w0 = 0, s_w1 = 0, s_b0 = 0, s_b1 = 0;      // per lane
    for (size_t tile = begin + (threadIdx.x >> 5) * 32; tile < end; tile += kBlock) {
        // ---- one parent per lane: moves, winning moves, the children worth keeping (as k_bfs_expand_insert)
        uint32_t nl = 0, nh = 0;
        const size_t i = tile + lane;
        if (i < end) {
            const uint4 v = __ldg(parents + i);
            const State4 parent = { v.x, v.y, v.z, v.w };
            const u64 m = mult ? mult[i] : 1ull;
            const int mover = piece_balance(parent);
            uint32_t lo, hi, wab, wcd;
            legal_masks(parent, mover, lo, hi);
            const int n_moves = popc(lo) + popc(hi);
            if (n_moves == 0) { if (mover == 0) s_b0 += m; else s_b1 += m; }           // game_stats.py:426-443
            else {
                winning_squares(parent, wab, wcd);
                s_moves += (u64)n_moves * m;                                            // :401
                const u64 wins = (u64)(popc(lo & wab) + popc(hi & wcd)) * m;            // :411-415
                if (mover == 0) s_w0 += wins; else s_w1 += wins;
                nl = lo & ~wab; nh = hi & ~wcd;
            }
            ws.state[lane] = v; ws.mult[lane] = m; ws.mover[lane] = (unsigned char)mover;
        }
        const uint32_t cnt = (uint32_t)(popc(nl) + popc(nh));
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (unsigned)o) incl += t; }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        uint32_t at = incl - cnt;
        for (uint32_t w = nl; w; w &= w - 1) ws.queue[at++] = (unsigned short)((lane << 6) | (uint32_t)(__ffs((int)w) - 1));
        for (uint32_t w = nh; w; w &= w - 1) ws.queue[at++] = (unsigned short)((lane << 6) | (uint32_t)(32 + __ffs((int)w) - 1));
        __syncwarp();
        // ---- full warps: one child per lane per round, canonical form -> its (part, block) range
        for (uint32_t base = 0; base < total; base += 32) {
            const uint32_t idx = base + lane;
            if (idx < total) {
                const uint32_t e = ws.queue[idx], pl = e >> 6;
                const uint4 pv = ws.state[pl];
                const State4 parent = { pv.x, pv.y, pv.z, pv.w };
                const State4 c = canonical<COLOR_SWAP>(apply_action(parent, ws.mover[pl], (int)(e & 63u)));   // :453,:468
                bucket_put(b, s_cur, ((u64)prmt(c.x, 0, 0x0123) << 32) | prmt(c.y, 0, 0x0123),
                           ((u64)prmt(c.z, 0, 0x0123) << 32) | prmt(c.w, 0, 0x0123), ws.mult[pl], all_ones_count, overflow);
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s_moves += __shfl_xor_sync(0xFFFFFFFFu, s_moves, o); s_w0 += __shfl_xor_sync(0xFFFFFFFFu, s_w0, o);
        s_w1 += __shfl_xor_sync(0xFFFFFFFFu, s_w1, o); s_b0 += __shfl_xor_sync(0xFFFFFFFFu, s_b0, o);
        s_b1 += __shfl_xor_sync(0xFFFFFFFFu, s_b1, o);
    }
    if (lane == 0) {
        if (s_moves) atomicAdd(stats + 0, s_moves);
        if (s_w0) atomicAdd(stats + 1, s_w0);
        if (s_w1) atomicAdd(stats + 2, s_w1);
        if (s_b0) atomicAdd(stats + 3, s_b0);
        if (s_b1) atomicAdd(stats + 4, s_b1);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < b.parts; k += kBlock)
        b.counts[(size_t)k * b.blocks + blockIdx.x] = s_cur[k] < b.sub_cap ? s_cur[k] : b.sub_cap;
}

// pass B, first half: every sub-bucket of one part into the L2-resident table.  Four keys per thread at a time: the
// four key loads, then the four probe loads are in flight together (a probe read early may be stale, which is safe:
// a slot only ever changes from empty to a key, an "empty" that is no longer true fails its CAS and a key stays).
static constexpr int kInsertBatch = 4;
__global__ void __launch_bounds__(kBlock) k_part_insert(Buckets b, int part, TSlot *table, u64 *n_unique, int *overflow, u64 max_unique)
{
    unsigned claims = 0;
    for (unsigned blk = blockIdx.x; blk < b.blocks; blk += gridDim.x) {
        const size_t sub = (size_t)part * b.blocks + blk, base = sub * b.sub_cap;
        const unsigned cnt = b.counts[sub];
        for (unsigned i0 = threadIdx.x; i0 < cnt; i0 += kBlock * kInsertBatch) {
            ulonglong2 key[kInsertBatch], raw[kInsertBatch];
            u64 m[kInsertBatch];
            uint32_t slot[kInsertBatch];
#pragma unroll
            for (int k = 0; k < kInsertBatch; ++k) {
                const unsigned i = i0 + k * kBlock;
                key[k] = make_ulonglong2(~0ull, ~0ull);
                m[k] = 1ull;
                if (i < cnt) {
                    key[k] = *reinterpret_cast<const ulonglong2 *>(b.keys + base + i);
                    if (b.mults) m[k] = b.mults[base + i];
                }
            }
            unsigned pending = 0;
#pragma unroll
            for (int k = 0; k < kInsertBatch; ++k) {
                slot[k] = __umulhi((uint32_t)mix(key[k].x, key[k].y), b.part_slots);
                if (i0 + k * kBlock < cnt) pending |= 1u << k;
            }
            // probe rounds in step for the whole batch: the loads of a round go out together, then its claims, then
            // the answers are looked at; a key whose slot holds another key moves one slot on and joins the next round
            // (finishing each key's chain on its own made the warp wait for 4 chains one after the other)
            for (uint32_t round = 0; pending; ++round) {
                if (round > kMaxProbe || round >= b.part_slots) { *overflow = 1; break; }
#pragma unroll
                for (int k = 0; k < kInsertBatch; ++k)
                    if ((pending >> k) & 1u) raw[k] = __ldcv(reinterpret_cast<const ulonglong2 *>(table + slot[k]));
                unsigned tried = 0;
#pragma unroll
                for (int k = 0; k < kInsertBatch; ++k)
                    if (((pending >> k) & 1u) && (raw[k].x == ~0ull || raw[k].y == ~0ull)) {
                        tried |= 1u << k;
                        cas128(table + slot[k], ~0ull, ~0ull, key[k].x, key[k].y, raw[k].x, raw[k].y);
                    }
#pragma unroll
                for (int k = 0; k < kInsertBatch; ++k) {
                    if (!((pending >> k) & 1u)) continue;
                    const bool claimed = ((tried >> k) & 1u) && raw[k].x == ~0ull && raw[k].y == ~0ull;
                    if (claimed || (raw[k].x == key[k].x && raw[k].y == key[k].y)) {
                        atomicAdd(&table[slot[k]].count, m[k]);                    // game_stats.py:487-505
                        claims += claimed ? 1u : 0u;
                        pending &= ~(1u << k);
                    } else {
                        slot[k] = slot[k] + 1 == b.part_slots ? 0 : slot[k] + 1;
                    }
                }
            }
        }
    }
    flush_claims(claims, n_unique, overflow, max_unique);
}

// pass B, second half: the part's classes are appended to the dense (key, multiplicity) arrays and the table is
// left empty for the next part.  A block takes a contiguous run of kDrainItems x 256 slots, counts its entries and
// reserves their place with ONE atomic (a warp-level reservation made the single cursor the bottleneck: 37 000
// same-address atomics per part).  part_slots is a multiple of 256.
static constexpr int kDrainItems = 4;
__global__ void __launch_bounds__(kBlock) k_part_drain(TSlot *table, uint32_t part_slots, Slot *keys, u64 *mults, u64 *cursor, u64 max_out)
{
    __shared__ unsigned s_warp[kBlock / 32];
    __shared__ u64 s_base;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (size_t run = blockIdx.x * (size_t)(kBlock * kDrainItems); run < part_slots; run += (size_t)gridDim.x * kBlock * kDrainItems) {
        TSlot e[kDrainItems];
        unsigned mask[kDrainItems], mine = 0;
#pragma unroll
        for (int k = 0; k < kDrainItems; ++k) {
            const size_t i = run + (size_t)k * kBlock + threadIdx.x;
            e[k].hi = ~0ull; e[k].lo = ~0ull; e[k].count = 0;
            if (i < part_slots) e[k] = load_entry(table, i);
            mask[k] = __ballot_sync(0xFFFFFFFFu, !(e[k].hi == ~0ull && e[k].lo == ~0ull));
            mine += (unsigned)__popc(mask[k]);
        }
        if (lane == 0) s_warp[warp] = mine;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned total = 0;
            for (unsigned w = 0; w < kBlock / 32; ++w) { const unsigned c = s_warp[w]; s_warp[w] = total; total += c; }
            s_base = total ? atomicAdd(cursor, (u64)total) : 0ull;
        }
        __syncthreads();
        u64 o = s_base + s_warp[warp];
#pragma unroll
        for (int k = 0; k < kDrainItems; ++k) {
            if ((mask[k] >> lane) & 1u) {
                const u64 at = o + (u64)__popc(mask[k] & ((1u << lane) - 1u));
                if (at < max_out) { keys[at].hi = e[k].hi; keys[at].lo = e[k].lo; mults[at] = e[k].count; }
                const ulonglong4 empty = { ~0ull, ~0ull, 0ull, 0ull };
                *reinterpret_cast<ulonglong4 *>(table + run + (size_t)k * kBlock + threadIdx.x) = empty;
            }
            o += (u64)__popc(mask[k]);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ host side
struct DedupTable { TSlot *table; u64 *meta; size_t slots; u64 max_unique; };

static int dedup_begin(Call &call, size_t expected_unique, DedupTable &t)
{
    cudaStream_t s = call.stream();
    // 1.5 slots per expected key (linear probing stays short up to ~2/3 full), a multiple of 256 so that warps
    // scan it converged
    const long long x10 = option(OPT_TABLE_SLOTS_X10);
    t.slots = x10 >= 11 ? (size_t)((double)expected_unique * (double)x10 / 10.0) + 2 : expected_unique + expected_unique / 2 + 2;
    if (t.slots < 1024) t.slots = 1024;
    t.slots = (t.slots + 255) / 256 * 256;
    t.table = (TSlot *)call.temp(t.slots * sizeof(TSlot));
    t.meta = (u64 *)call.temp(10 * sizeof(u64), true);   // n_unique, all-ones mult, cursor, overflow, stats[5]
    if (!t.table || !t.meta) return call.status();
    t.max_unique = (u64)(expected_unique < 512 ? 512 : expected_unique);
    k_dedup_fill<<<grid_for(2 * t.slots, kBlock, call.ctx()->sms, 8), kBlock, 0, s>>>(t.table, t.slots);
    QK_CUDA(cudaGetLastError());
    return QK_OK;
}

// dense (keys, mults)[0, nu) in arrays with room for `padded` + 1 entries (padded = a power of two >= nu and >= kTile when
// `sort`) -> the caller's buffers, in payload byte order when `sort`; the all-ones class (kept outside the tables) last
static int dedup_emit_dense(Call &call, Slot *keys, u64 *mults, size_t padded, size_t nu, bool special, u64 special_mult,
                            uint16_t *uniq, uint64_t *uniq_mult, size_t *n_uniq, bool sort)
{
    cudaStream_t s = call.stream();
    Ctx *c = call.ctx();
    if (sort) {
        k_pad<<<grid_for(padded - nu + 1, kBlock, c->sms, 8), kBlock, 0, s>>>(keys, mults, nu, padded);
        const unsigned tile_grid = grid_for(padded / kTile * kBlock, kBlock, c->sms, 8);
        k_bitonic_tile<<<tile_grid, kBlock, 0, s>>>(keys, mults, padded, 2, kTile);
        for (size_t k = 2 * (size_t)kTile; k <= padded; k <<= 1) {
            for (size_t j = k >> 1; j >= kTile; j >>= 1)
                k_bitonic_step<<<grid_for(padded, kBlock, c->sms, 8), kBlock, 0, s>>>(keys, mults, padded, j, k);
            k_bitonic_tile<<<tile_grid, kBlock, 0, s>>>(keys, mults, padded, k, k);
        }
        QK_CUDA(cudaGetLastError());
    }
    if (special) {   // the all-ones payload sorts last by construction
        Slot ones = { ~0ull, ~0ull };
        QK_CUDA(cudaMemcpyAsync(keys + nu, &ones, sizeof ones, cudaMemcpyHostToDevice, s));
        QK_CUDA(cudaMemcpyAsync(mults + nu, &special_mult, sizeof(u64), cudaMemcpyHostToDevice, s));
        QK_CUDA(cudaStreamSynchronize(s));
        ++nu;
    }
    *n_uniq = nu;
    if (nu && (uniq || uniq_mult)) {
        uint4 *d_uniq = (uint4 *)call.out(uniq, nu * 16);
        u64 *d_um = (u64 *)call.out(uniq_mult, nu * 8);
        if (call.status() != QK_OK) return call.status();
        k_emit<<<grid_for(nu, kBlock, c->sms, 8), kBlock, 0, s>>>(keys, mults, nu, d_uniq, d_um);
        QK_CUDA(cudaGetLastError());
    }
    return QK_OK;
}

// table -> sorted distinct payloads + multiplicities in the caller's buffers; h_stats (5 x u64, may be null)
static int dedup_finish(Call &call, DedupTable &t, uint16_t *uniq, uint64_t *uniq_mult, size_t cap, size_t *n_uniq,
                        uint64_t *h_stats, const char *who, bool sort_output = true)
{
    cudaStream_t s = call.stream();
    Ctx *c = call.ctx();
    u64 h_meta[10];
    QK_CUDA(cudaMemcpyAsync(h_meta, t.meta, sizeof h_meta, cudaMemcpyDeviceToHost, s));
    QK_CUDA(cudaStreamSynchronize(s));
    if (h_stats) for (int i = 0; i < 5; ++i) h_stats[i] = h_meta[4 + i];
    size_t nu = (size_t)h_meta[0];
    const bool special = h_meta[1] != 0;
    if ((int)h_meta[3] != 0 || (nu + (special ? 1 : 0) > cap && (uniq || uniq_mult))) {
        *n_uniq = nu + (special ? 1 : 0);
        set_error(std::string(who) + ": more distinct canonical states than `cap`");
        return QK_E_NOMEM;
    }
    if (!sort_output && uniq && uniq_mult) {
        // arbitrary order: straight from the table into the caller's arrays
        uint4 *d_uniq = (uint4 *)call.out(uniq, (nu + 1) * 16);
        u64 *d_um = (u64 *)call.out(uniq_mult, (nu + 1) * 8);
        if (call.status() != QK_OK) return call.status();
        k_dedup_compact<true><<<compact_grid(t.slots, c->sms), kBlock, 0, s>>>(t.table, t.slots, nullptr, nullptr, d_uniq, d_um, t.meta + 2);
        QK_CUDA(cudaGetLastError());
        if (special) {
            const uint32_t ones[4] = { ~0u, ~0u, ~0u, ~0u };
            QK_CUDA(cudaMemcpyAsync(d_uniq + nu, ones, 16, cudaMemcpyHostToDevice, s));
            QK_CUDA(cudaMemcpyAsync(d_um + nu, &h_meta[1], sizeof(u64), cudaMemcpyHostToDevice, s));
            QK_CUDA(cudaStreamSynchronize(s));
            ++nu;
        }
        *n_uniq = nu;
        return QK_OK;
    }
    size_t padded = kTile;
    while (padded < nu) padded <<= 1;
    Slot *keys = (Slot *)call.temp((padded + 1) * sizeof(Slot));
    u64 *mults = (u64 *)call.temp((padded + 1) * sizeof(u64));
    if (!keys || !mults) return call.status();
    k_dedup_compact<false><<<compact_grid(t.slots, c->sms), kBlock, 0, s>>>(t.table, t.slots, keys, mults, nullptr, nullptr, t.meta + 2);
    QK_CUDA(cudaGetLastError());
    return dedup_emit_dense(call, keys, mults, padded, nu, special, h_meta[1], uniq, uniq_mult, n_uniq, true);
}

// Plan of a two-level merge of `n_keys` keys with at most `expected_unique` classes (see the kernels above).
static bool plan_two_level(size_t n_keys, size_t expected_unique, int sms, int streams, Buckets &b, size_t &per_block)
{
    const size_t slots = expected_unique + expected_unique / 2 + 2;
    (void)streams;
    const size_t part_bytes = l2_table_bytes();
    size_t parts = (slots * sizeof(TSlot) + part_bytes - 1) / part_bytes;
    if (parts < 2) parts = 2;
    if (parts > (size_t)kMaxParts) return false;
    // hash-uniform shares: mean + 6 sigma of the classes of one part, 1.5 slots per class, a multiple of 256
    const double mean = (double)expected_unique / (double)parts;
    size_t part_slots = (size_t)(1.5 * (mean + 6.0 * sqrt(mean) + 64.0));
    part_slots = (part_slots + 255) / 256 * 256;
    if (part_slots >= (1ull << 32)) return false;
    b.parts = (int)parts;
    b.part_slots = (uint32_t)part_slots;
    b.blocks = (unsigned)sms * 4;
    per_block = (n_keys + b.blocks - 1) / b.blocks;
    per_block = (per_block + kBlock - 1) / kBlock * kBlock;
    const double share = (double)per_block / (double)parts;
    b.sub_cap = (uint32_t)(share + 6.0 * sqrt(share) + 32.0);
    return true;
}

// side streams of the merge: created once per device, non-blocking (they are ordered against the caller's stream
// by events only, whatever kind of stream that is)
static int merge_streams(Ctx *c, int n_aux)
{
    if (!c->fork_ev) QK_CUDA(cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming));
    for (int i = 0; i < n_aux; ++i) {
        if (!c->aux[i]) QK_CUDA(cudaStreamCreateWithFlags(&c->aux[i], cudaStreamNonBlocking));
        if (!c->join_ev[i]) QK_CUDA(cudaEventCreateWithFlags(&c->join_ev[i], cudaEventDisableTiming));
    }
    return QK_OK;
}
static int dedup_streams()
{
    const long long v = option(OPT_DEDUP_STREAMS);
    return v <= 0 ? 3 : (v > Ctx::kAux + 1 ? Ctx::kAux + 1 : (int)v);
}

// pass B for every part + the tail.  meta = n_unique, all-ones mult, cursor, overflow (as DedupTable::meta).
// One part's kernels are short (about 75 + 19 us at 2^24 keys, a fifth of it ramp-up and tail), so `streams` parts
// are in flight at a time, each in its own full-size table on its own stream: part k + 1 starts on the SMs that part
// k's tail and drain leave idle.  An insert kernel fills the machine, so the parts mostly run one after the other and
// the L2 holds the table that is being filled.  Measured (2^24 keys, whole call): one stream 1.15 ms, two 1.07 ms, three 1.04 ms
// (the default);
// splitting the L2 between the tables instead (2 x 48 MiB: 1.31 ms, 4 x 24 MiB: 1.55 ms) loses -- twice the parts,
// twice the launches.  The parts are independent -- different keys, different tables; they share only the append
// cursor and the counters, which are atomics.
struct DenseClasses { Slot *keys; u64 *mults; size_t padded, n; u64 all_ones_mult; };   // n classes + the all-ones payload's multiplicity (0 = absent)

static int merge_buckets_dense(Call &call, const Buckets &b, int streams, u64 *meta, size_t expected_unique, bool sort_output,
                               DenseClasses &out, bool *fallback, uint64_t *h_stats = nullptr)
{
    cudaStream_t s = call.stream();
    Ctx *c = call.ctx();
    *fallback = false;
    if (streams > b.parts) streams = b.parts;
    int rc = merge_streams(c, streams - 1);
    if (rc != QK_OK) return rc;
    TSlot *table = (TSlot *)call.temp((size_t)streams * b.part_slots * sizeof(TSlot));
    size_t padded = kTile;
    while (padded < expected_unique) padded <<= 1;
    if (!sort_output) padded = expected_unique;
    Slot *keys = (Slot *)call.temp((padded + 1) * sizeof(Slot));
    u64 *mults = (u64 *)call.temp((padded + 1) * sizeof(u64));
    if (!table || !keys || !mults) return call.status();
    k_dedup_fill<<<grid_for(2 * (size_t)streams * b.part_slots, kBlock, c->sms, 8), kBlock, 0, s>>>(table, (size_t)streams * b.part_slots);
    const u64 max_unique = (u64)(expected_unique < 512 ? 512 : expected_unique);
    const unsigned insert_grid = b.blocks < (unsigned)c->sms * 8 ? b.blocks : (unsigned)c->sms * 8;
    const unsigned drain_grid = grid_for((b.part_slots + kDrainItems - 1) / kDrainItems, kBlock, c->sms, 8);
    if (streams > 1) {
        QK_CUDA(cudaEventRecord(c->fork_ev, s));
        for (int i = 0; i + 1 < streams; ++i) QK_CUDA(cudaStreamWaitEvent(c->aux[i], c->fork_ev, 0));
    }
    for (int part = 0; part < b.parts; ++part) {
        const int lane = part % streams;
        cudaStream_t st = lane == 0 ? s : c->aux[lane - 1];
        TSlot *t = table + (size_t)lane * b.part_slots;
        k_part_insert<<<insert_grid, kBlock, 0, st>>>(b, part, t, meta, (int *)(meta + 3), max_unique);
        k_part_drain<<<drain_grid, kBlock, 0, st>>>(t, b.part_slots, keys, mults, meta + 2, (u64)padded);
    }
    const cudaError_t launch_err = cudaGetLastError();
    for (int i = 0; i + 1 < streams; ++i) {             // join even after a failed launch: `s` frees the tables
        cudaEventRecord(c->join_ev[i], c->aux[i]);
        cudaStreamWaitEvent(s, c->join_ev[i], 0);
    }
    QK_CUDA(launch_err);
    QK_CUDA(cudaGetLastError());
    u64 h_meta[10];
    QK_CUDA(cudaMemcpyAsync(h_meta, meta, sizeof h_meta, cudaMemcpyDeviceToHost, s));
    QK_CUDA(cudaStreamSynchronize(s));
    if ((int)h_meta[3] != 0) { *fallback = true; return QK_OK; }         // a range or the part table was too small
    if (h_stats) for (int i = 0; i < 5; ++i) h_stats[i] = h_meta[4 + i];
    call.release(table);
    out.keys = keys; out.mults = mults; out.padded = padded; out.n = (size_t)h_meta[0]; out.all_ones_mult = h_meta[1];
    return QK_OK;
}

// ... and into the caller's buffers
static int merge_buckets(Call &call, const Buckets &b, int streams, u64 *meta, size_t expected_unique, uint16_t *uniq,
                         uint64_t *uniq_mult, size_t cap, size_t *n_uniq, bool sort_output, const char *who, bool *fallback,
                         uint64_t *h_stats = nullptr)
{
    DenseClasses d;
    const int rc = merge_buckets_dense(call, b, streams, meta, expected_unique, sort_output, d, fallback, h_stats);
    if (rc != QK_OK || *fallback) return rc;
    const bool special = d.all_ones_mult != 0;
    if (d.n + (special ? 1 : 0) > cap && (uniq || uniq_mult)) {
        *n_uniq = d.n + (special ? 1 : 0);
        set_error(std::string(who) + ": more distinct canonical states than `cap`");
        return QK_E_NOMEM;
    }
    return dedup_emit_dense(call, d.keys, d.mults, d.padded, d.n, special, d.all_ones_mult, uniq, uniq_mult, n_uniq, sort_output);
}

static int alloc_buckets(Call &call, Buckets &b, bool with_mults)
{
    const size_t entries = (size_t)b.parts * b.blocks * b.sub_cap;
    b.keys = (Slot *)call.temp(entries * sizeof(Slot));
    b.mults = with_mults ? (u64 *)call.temp(entries * sizeof(u64)) : nullptr;
    b.counts = (unsigned *)call.temp((size_t)b.parts * b.blocks * sizeof(unsigned));
    return (!b.keys || !b.counts || (with_mults && !b.mults)) ? call.status() : QK_OK;
}

int run_dedup(Call &call, const uint4 *states, const uint64_t *mult, size_t n, int color_swap,
              uint16_t *uniq, uint64_t *uniq_mult, size_t cap, size_t *n_uniq)
{
    cudaStream_t s = call.stream();
    Ctx *c = call.ctx();
    const bool unsorted = (color_swap & QK_DEDUP_UNSORTED) != 0;
    const size_t expected = n < cap ? n : cap;
    // two-level merge for all but small inputs (qk_set_option "dedup_mode": 1 = never, 2 = always)
    const long long mode = option(OPT_DEDUP_MODE);
    Buckets b;
    size_t per_block = 0;
    // measured (scripts/dedup_threshold.py, whole call, one table / two-level): 2^18 keys 0.098 / 0.113 ms, 2^19 0.134 / 0.123,
    // 2^20 0.207 / 0.148, 2^22 0.575 / 0.347, 2^24 1.48 / 1.06, 2^25 2.48 / 2.26 -- the routed form wins from half a million keys on
    if (mode != 1 && (mode == 2 || (expected + expected / 2) * sizeof(TSlot) > (24u << 20)) &&
        plan_two_level(n, expected, c->sms, dedup_streams(), b, per_block)) {
        int rc = alloc_buckets(call, b, mult != nullptr);
        if (rc != QK_OK) return rc;
        u64 *meta = (u64 *)call.temp(10 * sizeof(u64), true);
        if (!meta) return call.status();
        const size_t smem = (size_t)b.parts * sizeof(unsigned);
        if (color_swap & 1)
            k_bucket_states<true><<<b.blocks, kBlock, smem, s>>>(states, (const u64 *)mult, n, per_block, b, meta + 1, (int *)(meta + 3));
        else
            k_bucket_states<false><<<b.blocks, kBlock, smem, s>>>(states, (const u64 *)mult, n, per_block, b, meta + 1, (int *)(meta + 3));
        QK_CUDA(cudaGetLastError());
        bool fallback = false;
        rc = merge_buckets(call, b, dedup_streams(), meta, expected, uniq, uniq_mult, cap, n_uniq, !unsorted, "qk_canon_dedup", &fallback);
        if (rc != QK_OK || !fallback) return rc;
        call.release(b.keys); call.release(b.counts);
        if (b.mults) call.release(b.mults);
    }
    DedupTable t;
    int rc = dedup_begin(call, expected, t);
    if (rc != QK_OK) return rc;
    if (color_swap & 1)
        k_dedup_insert<true><<<grid_for(n, kBlock, c->sms, 8), kBlock, 0, s>>>(states, (const u64 *)mult, n, t.table, t.slots,
                                                                             t.meta, t.meta + 1, (int *)(t.meta + 3), t.max_unique);
    else
        k_dedup_insert<false><<<grid_for(n, kBlock, c->sms, 8), kBlock, 0, s>>>(states, (const u64 *)mult, n, t.table, t.slots,
                                                                              t.meta, t.meta + 1, (int *)(t.meta + 3), t.max_unique);
    QK_CUDA(cudaGetLastError());
    return dedup_finish(call, t, uniq, uniq_mult, cap, n_uniq, nullptr, "qk_canon_dedup", !unsorted);
}

// the two-level form of one BFS level; *fallback = a range or a part table was too small (the caller runs the one-level form)
static int bfs_level_two_level(Call &call, const uint4 *parents, const uint64_t *mult, size_t n, int color_swap, int sort_output,
                               uint16_t *uniq, uint64_t *uniq_mult, size_t cap, size_t *n_uniq, uint64_t *h_stats, bool *fallback)
{
    cudaStream_t s = call.stream();
    Ctx *c = call.ctx();
    *fallback = true;
    Buckets b;
    b.blocks = (unsigned)c->sms * 4;
    size_t per_block = (n + b.blocks - 1) / b.blocks;
    per_block = (per_block + 31) / 32 * 32;
    unsigned *d_children = (unsigned *)call.temp(b.blocks * sizeof(unsigned), true);
    if (!d_children) return call.status();
    k_bfs_count_children<<<b.blocks, kBlock, 0, s>>>(parents, n, per_block, d_children);
    QK_CUDA(cudaGetLastError());
    std::vector<unsigned> h_children(b.blocks);
    QK_CUDA(cudaMemcpyAsync(h_children.data(), d_children, b.blocks * sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    QK_CUDA(cudaStreamSynchronize(s));
    call.release(d_children);
    unsigned most = 0;
    for (unsigned v : h_children) most = v > most ? v : most;
    // parts: the table of `cap` classes cut into L2-sized pieces; a part's classes and a block's share of a part are
    // hash-uniform, sized at mean + 6 sigma as in plan_two_level
    const size_t slots = cap + cap / 2 + 2;
    size_t parts = (slots * sizeof(TSlot) + l2_table_bytes() - 1) / l2_table_bytes();
    if (parts < 2) parts = 2;
    if (parts > (size_t)kMaxParts) return QK_OK;
    const double mean = (double)cap / (double)parts, share = (double)most / (double)parts;
    size_t part_slots = (size_t)(1.5 * (mean + 6.0 * sqrt(mean) + 64.0));
    part_slots = (part_slots + 255) / 256 * 256;
    if (part_slots >= (1ull << 32)) return QK_OK;
    b.parts = (int)parts;
    b.part_slots = (uint32_t)part_slots;
    b.sub_cap = (uint32_t)(share + 6.0 * sqrt(share) + 32.0);
    int rc = alloc_buckets(call, b, true);
    if (rc != QK_OK) return rc;
    u64 *meta = (u64 *)call.temp(10 * sizeof(u64), true);      // n_unique, all-ones mult, cursor, overflow, stats[5]
    if (!meta) return call.status();
    const size_t smem = (size_t)b.parts * sizeof(unsigned);
    if (color_swap)
        k_bfs_expand_bucket<true><<<b.blocks, kBlock, smem, s>>>(parents, (const u64 *)mult, n, per_block, b, meta + 1, (int *)(meta + 3), meta + 4);
    else
        k_bfs_expand_bucket<false><<<b.blocks, kBlock, smem, s>>>(parents, (const u64 *)mult, n, per_block, b, meta + 1, (int *)(meta + 3), meta + 4);
    QK_CUDA(cudaGetLastError());
    rc = merge_buckets(call, b, dedup_streams(), meta, cap, uniq, uniq_mult, cap, n_uniq, sort_output != 0, "qk_bfs_level", fallback, h_stats);
    call.release(b.keys); call.release(b.mults); call.release(b.counts);
    return rc;
}

int run_bfs_level(Call &call, const uint4 *parents, const uint64_t *mult, size_t n, int color_swap, int sort_output,
                  uint16_t *uniq, uint64_t *uniq_mult, size_t cap, size_t *n_uniq, uint64_t *h_stats)
{
    cudaStream_t s = call.stream();
    Ctx *c = call.ctx();
    // The routed (two-level) form is built and parity-tested but NOT the default here: on the largest level (4.1 x 10^8
    // children, 55 M classes) it takes 10.3 ms to expand + route and 10.3 + 1.3 ms to merge 42 parts, against 16.0 +
    // 0.85 + 1.2 ms for this one table (complete game tree 0.096 s against 0.063 s; profiles/r02n_bfs_two_level.txt).
    // qk_set_option("dedup_mode", 2) selects it.
    const long long mode = option(OPT_DEDUP_MODE);
    if (mode == 2 && n > 0) {
        bool fallback = false;
        const int rc = bfs_level_two_level(call, parents, mult, n, color_swap, sort_output, uniq, uniq_mult, cap, n_uniq, h_stats, &fallback);
        if (rc != QK_OK || !fallback) return rc;
    }
    DedupTable t;
    int rc = dedup_begin(call, cap, t);
    if (rc != QK_OK) return rc;
    // one parent per lane (tiles of 32 per warp), grid-stride
    const unsigned grid = grid_for(n, kBlock, c->sms, 5);
    if (color_swap)
        k_bfs_expand_insert<true><<<grid, kBlock, 0, s>>>(parents, (const u64 *)mult, n, t.table, t.slots, t.meta, t.meta + 1,
                                                         (int *)(t.meta + 3), t.max_unique, t.meta + 4);
    else
        k_bfs_expand_insert<false><<<grid, kBlock, 0, s>>>(parents, (const u64 *)mult, n, t.table, t.slots, t.meta, t.meta + 1,
                                                          (int *)(t.meta + 3), t.max_unique, t.meta + 4);
    QK_CUDA(cudaGetLastError());
    return dedup_finish(call, t, uniq, uniq_mult, cap, n_uniq, h_stats, "qk_bfs_level", sort_output != 0);
}

// ------------------------------------------------------------------ exchange step over peer memory
// Inbox of one rank: [0, 128) u64 counts[source] | Slot keys[world][slots_per_src] | u64 mults[world][slots_per_src]
struct InboxView { u64 *counts; Slot *keys; u64 *mults; };
static constexpr size_t kInboxHeader = 128;

size_t inbox_bytes(int world, size_t slots_per_src)
{
    return kInboxHeader + (size_t)world * slots_per_src * (sizeof(Slot) + sizeof(u64));
}
__host__ __device__ __forceinline__ InboxView inbox_view(void *base, int world, size_t slots_per_src)
{
    char *b = (char *)base;
    InboxView v = { (u64 *)b, (Slot *)(b + kInboxHeader), (u64 *)(b + kInboxHeader + (size_t)world * slots_per_src * sizeof(Slot)) };
    return v;
}
// owner rank of a canonical payload: high bits of a second mix, so it is independent of the slot index bits
__device__ __forceinline__ int owner_of(u64 hi, u64 lo, int world)
{
    return (int)__umulhi((uint32_t)(mix(lo, hi) >> 32), (uint32_t)world);
}

__global__ void __launch_bounds__(kBlock) k_payload_owner(const uint4 *__restrict__ payloads, size_t n, int world, uint8_t *__restrict__ owner)
{
    const size_t i = blockIdx.x * (size_t)kBlock + threadIdx.x;
    if (i >= n) return;
    const uint4 v = __ldg(payloads + i);
    // Key128 limbs = big-endian words of the payload bytes (qk_bits.cuh sorted_image / key_to_state)
    const u64 hi = ((u64)prmt(v.x, 0, 0x0123) << 32) | prmt(v.y, 0, 0x0123), lo = ((u64)prmt(v.z, 0, 0x0123) << 32) | prmt(v.w, 0, 0x0123);
    owner[i] = (uint8_t)owner_of(hi, lo, world);
}

int launch_payload_owner(const uint4 *payloads, size_t n, int world, uint8_t *owner, Ctx *c, cudaStream_t s)
{
    (void)c;
    k_payload_owner<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, s>>>(payloads, n, world, owner);
    QK_CUDA(cudaGetLastError());
    return QK_OK;
}

// local table -> the owners' inboxes.  Per 256 slots a block counts its entries per destination in shared
// memory, reserves one range per destination in the SOURCE-side cursor (regions are per source, so no remote
// atomics), then every thread stores its (key, multiplicity) straight into the peer's memory.
__global__ void __launch_bounds__(kBlock) k_dedup_scatter(const TSlot *table, size_t slots, Peers p,
                                                           u64 *cursor, int *overflow)
{
    __shared__ unsigned s_cnt[QK_MAX_WORLD];
    __shared__ u64 s_base[QK_MAX_WORLD];
    const size_t stride = (size_t)gridDim.x * kBlock;
    const size_t rounds = (slots + stride - 1) / stride;
    for (size_t round = 0; round < rounds; ++round) {
        const size_t i = round * stride + blockIdx.x * (size_t)kBlock + threadIdx.x;
        if (threadIdx.x < QK_MAX_WORLD) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        TSlot e = { ~0ull, ~0ull, 0ull, 0ull };
        if (i < slots) e = load_entry(table, i);
        const Slot k = { e.hi, e.lo };
        const bool have = !(k.hi == ~0ull && k.lo == ~0ull);
        int dst = 0;
        unsigned r = 0;
        const u64 m = e.count;
        if (have) {
            dst = owner_of(k.hi, k.lo, p.world);
            r = atomicAdd(&s_cnt[dst], 1u);
        }
        __syncthreads();
        if (threadIdx.x < (unsigned)p.world && s_cnt[threadIdx.x])
            s_base[threadIdx.x] = atomicAdd(cursor + threadIdx.x, (u64)s_cnt[threadIdx.x]);
        __syncthreads();
        if (have) {
            const u64 pos = s_base[dst] + r;
            if (pos < p.slots_per_src) {
                const InboxView v = inbox_view(p.inbox[dst], p.world, p.slots_per_src);
                const size_t at = (size_t)p.rank * p.slots_per_src + pos;
                v.keys[at] = k;
                v.mults[at] = m;
            } else {
                *overflow = 1;
            }
        }
    }
}

// the same from a dense list of classes (the two-level merge's output): every entry is a class
__global__ void __launch_bounds__(kBlock) k_dense_scatter(const Slot *__restrict__ keys, const u64 *__restrict__ mults, size_t n, Peers p,
                                                           u64 *cursor, int *overflow)
{
    __shared__ unsigned s_cnt[QK_MAX_WORLD];
    __shared__ u64 s_base[QK_MAX_WORLD];
    const size_t stride = (size_t)gridDim.x * kBlock;
    const size_t rounds = (n + stride - 1) / stride;
    for (size_t round = 0; round < rounds; ++round) {
        const size_t i = round * stride + blockIdx.x * (size_t)kBlock + threadIdx.x;
        if (threadIdx.x < QK_MAX_WORLD) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        const bool have = i < n;
        Slot k = { 0, 0 };
        u64 m = 0;
        int dst = 0;
        unsigned r = 0;
        if (have) {
            k = keys[i]; m = mults[i];
            dst = owner_of(k.hi, k.lo, p.world);
            r = atomicAdd(&s_cnt[dst], 1u);
        }
        __syncthreads();
        if (threadIdx.x < (unsigned)p.world && s_cnt[threadIdx.x])
            s_base[threadIdx.x] = atomicAdd(cursor + threadIdx.x, (u64)s_cnt[threadIdx.x]);
        __syncthreads();
        if (have) {
            const u64 pos = s_base[dst] + r;
            if (pos < p.slots_per_src) {
                const InboxView v = inbox_view(p.inbox[dst], p.world, p.slots_per_src);
                const size_t at = (size_t)p.rank * p.slots_per_src + pos;
                v.keys[at] = k;
                v.mults[at] = m;
            } else {
                *overflow = 1;
            }
        }
    }
}

// after the scatter: the all-ones payload (kept outside the table) joins its owner's region, and every
// destination learns how many entries this source wrote
__global__ void k_scatter_publish(Peers p, u64 *cursor, u64 all_ones_mult)
{
    const int dst = (int)threadIdx.x;
    if (dst >= p.world) return;
    const InboxView v = inbox_view(p.inbox[dst], p.world, p.slots_per_src);
    u64 n = cursor[dst];
    if (all_ones_mult && owner_of(~0ull, ~0ull, p.world) == dst) {
        if (n < p.slots_per_src) {
            const size_t at = (size_t)p.rank * p.slots_per_src + n;
            v.keys[at].hi = ~0ull; v.keys[at].lo = ~0ull;
            v.mults[at] = all_ones_mult;
        }
        ++n;
        cursor[dst] = n;
    }
    v.counts[p.rank] = n;
}

// the owner's side: entries of one inbox region (canonical already) into the merge table
__global__ void __launch_bounds__(kBlock) k_inbox_insert(const Slot *keys, const u64 *mults, size_t n, TSlot *table,
                                                          size_t slots, u64 *n_unique, u64 *all_ones_count, int *overflow,
                                                          u64 max_unique)
{
    unsigned claims = 0;
    for (size_t i = blockIdx.x * (size_t)kBlock + threadIdx.x; i < n; i += (size_t)gridDim.x * kBlock) {
        if (*reinterpret_cast<volatile int *>(overflow)) break;
        const Slot key = keys[i];
        claims += table_insert(table, slots, key.hi, key.lo, mults[i], all_ones_count, overflow) ? 1u : 0u;
    }
    flush_claims(claims, n_unique, overflow, max_unique);
}

// classes (a local table, or a dense list) -> peers; h_sent (world x u64, may be null)
static int deliver(Call &call, const DedupTable *t, const DenseClasses *d, u64 all_ones_mult, const Peers &peers, uint64_t *h_sent,
                   const char *who)
{
    cudaStream_t s = call.stream();
    Ctx *c = call.ctx();
    u64 *cursor = (u64 *)call.temp((QK_MAX_WORLD + 1) * sizeof(u64), true);     // [QK_MAX_WORLD] = overflow flag
    if (!cursor) return call.status();
    int *overflow = (int *)(cursor + QK_MAX_WORLD);
    if (t) k_dedup_scatter<<<grid_for(t->slots, kBlock, c->sms, 8), kBlock, 0, s>>>(t->table, t->slots, peers, cursor, overflow);
    else if (d->n) k_dense_scatter<<<grid_for(d->n, kBlock, c->sms, 8), kBlock, 0, s>>>(d->keys, d->mults, d->n, peers, cursor, overflow);
    k_scatter_publish<<<1, 32, 0, s>>>(peers, cursor, all_ones_mult);
    QK_CUDA(cudaGetLastError());
    u64 h_cursor[QK_MAX_WORLD + 1];
    QK_CUDA(cudaMemcpyAsync(h_cursor, cursor, sizeof h_cursor, cudaMemcpyDeviceToHost, s));
    QK_CUDA(cudaStreamSynchronize(s));                                          // peer stores are complete here
    bool too_small = (int)h_cursor[QK_MAX_WORLD] != 0;
    for (int r = 0; r < peers.world; ++r) {
        if (h_sent) h_sent[r] = h_cursor[r];
        too_small |= h_cursor[r] > peers.slots_per_src;
    }
    if (too_small) {
        set_error(std::string(who) + ": an inbox region (slots_per_src) is too small for this rank's share");
        return QK_E_NOMEM;
    }
    return QK_OK;
}

// local table -> peers; h_stats (5 x u64, may be null), h_sent (world x u64, may be null)
static int scatter_finish(Call &call, DedupTable &t, const Peers &peers, uint64_t *h_stats, uint64_t *h_sent, const char *who)
{
    cudaStream_t s = call.stream();
    u64 h_meta[10];
    QK_CUDA(cudaMemcpyAsync(h_meta, t.meta, sizeof h_meta, cudaMemcpyDeviceToHost, s));
    QK_CUDA(cudaStreamSynchronize(s));
    if (h_stats) for (int i = 0; i < 5; ++i) h_stats[i] = h_meta[4 + i];
    if ((int)h_meta[3] != 0) {
        set_error(std::string(who) + ": more distinct canonical states than the local table holds");
        return QK_E_NOMEM;
    }
    return deliver(call, &t, nullptr, h_meta[1], peers, h_sent, who);
}

int run_dedup_scatter(Call &call, const uint4 *states, const uint64_t *mult, size_t n, int flags, const Peers &peers, uint64_t *h_sent)
{
    cudaStream_t s = call.stream();
    Ctx *c = call.ctx();
    // the local merge in two-level form (as run_dedup chooses it), its dense output delivered directly
    const long long mode = option(OPT_DEDUP_MODE);
    Buckets b;
    size_t per_block = 0;
    if (n && mode != 1 && (mode == 2 || (n + n / 2) * sizeof(TSlot) > (24u << 20)) && plan_two_level(n, n, c->sms, dedup_streams(), b, per_block)) {
        int rc = alloc_buckets(call, b, mult != nullptr);
        if (rc != QK_OK) return rc;
        u64 *meta = (u64 *)call.temp(10 * sizeof(u64), true);
        if (!meta) return call.status();
        const size_t smem = (size_t)b.parts * sizeof(unsigned);
        if (flags & 1)
            k_bucket_states<true><<<b.blocks, kBlock, smem, s>>>(states, (const u64 *)mult, n, per_block, b, meta + 1, (int *)(meta + 3));
        else
            k_bucket_states<false><<<b.blocks, kBlock, smem, s>>>(states, (const u64 *)mult, n, per_block, b, meta + 1, (int *)(meta + 3));
        QK_CUDA(cudaGetLastError());
        DenseClasses d;
        bool fallback = false;
        rc = merge_buckets_dense(call, b, dedup_streams(), meta, n, false, d, &fallback);
        if (rc != QK_OK) return rc;
        call.release(b.keys); call.release(b.counts);
        if (b.mults) call.release(b.mults);
        if (!fallback) return deliver(call, nullptr, &d, d.all_ones_mult, peers, h_sent, "qk_canon_dedup_scatter");
    }
    DedupTable t;
    int rc = dedup_begin(call, n, t);
    if (rc != QK_OK) return rc;
    if (n) {
        if (flags & 1)
            k_dedup_insert<true><<<grid_for(n, kBlock, c->sms, 8), kBlock, 0, s>>>(states, (const u64 *)mult, n, t.table, t.slots,
                                                                                 t.meta, t.meta + 1, (int *)(t.meta + 3), t.max_unique);
        else
            k_dedup_insert<false><<<grid_for(n, kBlock, c->sms, 8), kBlock, 0, s>>>(states, (const u64 *)mult, n, t.table, t.slots,
                                                                                  t.meta, t.meta + 1, (int *)(t.meta + 3), t.max_unique);
        QK_CUDA(cudaGetLastError());
    }
    return scatter_finish(call, t, peers, nullptr, h_sent, "qk_canon_dedup_scatter");
}

int run_bfs_level_scatter(Call &call, const uint4 *parents, const uint64_t *mult, size_t n, int flags, size_t local_cap,
                          const Peers &peers, uint64_t *h_stats, uint64_t *h_sent)
{
    cudaStream_t s = call.stream();
    Ctx *c = call.ctx();
    DedupTable t;
    int rc = dedup_begin(call, local_cap, t);
    if (rc != QK_OK) return rc;
    if (n) {
        const unsigned grid = grid_for(n, kBlock, c->sms, 5);
        if (flags & 1)
            k_bfs_expand_insert<true><<<grid, kBlock, 0, s>>>(parents, (const u64 *)mult, n, t.table, t.slots, t.meta, t.meta + 1,
                                                             (int *)(t.meta + 3), t.max_unique, t.meta + 4);
        else
            k_bfs_expand_insert<false><<<grid, kBlock, 0, s>>>(parents, (const u64 *)mult, n, t.table, t.slots, t.meta, t.meta + 1,
                                                              (int *)(t.meta + 3), t.max_unique, t.meta + 4);
        QK_CUDA(cudaGetLastError());
    }
    return scatter_finish(call, t, peers, h_stats, h_sent, "qk_bfs_level_scatter");
}

int run_inbox_merge(Call &call, const void *my_inbox, int world, size_t slots_per_src, int sort_output,
                    uint16_t *uniq, uint64_t *uniq_mult, size_t cap, size_t *n_uniq)
{
    cudaStream_t s = call.stream();
    Ctx *c = call.ctx();
    const InboxView v = inbox_view(const_cast<void *>(my_inbox), world, slots_per_src);
    u64 h_counts[QK_MAX_WORLD];
    QK_CUDA(cudaMemcpyAsync(h_counts, v.counts, world * sizeof(u64), cudaMemcpyDeviceToHost, s));
    QK_CUDA(cudaStreamSynchronize(s));
    size_t total = 0;
    for (int r = 0; r < world; ++r) {
        if (h_counts[r] > slots_per_src) {
            set_error("qk_inbox_merge: source rank " + std::to_string(r) + " overflowed its region");
            return QK_E_NOMEM;
        }
        total += (size_t)h_counts[r];
    }
    DedupTable t;
    int rc = dedup_begin(call, total < cap ? total : cap, t);
    if (rc != QK_OK) return rc;
    for (int r = 0; r < world; ++r) {
        if (!h_counts[r]) continue;
        k_inbox_insert<<<grid_for((size_t)h_counts[r], kBlock, c->sms, 8), kBlock, 0, s>>>(
            v.keys + (size_t)r * slots_per_src, v.mults + (size_t)r * slots_per_src, (size_t)h_counts[r], t.table, t.slots,
            t.meta, t.meta + 1, (int *)(t.meta + 3), t.max_unique);
    }
    QK_CUDA(cudaGetLastError());
    return dedup_finish(call, t, uniq, uniq_mult, cap, n_uniq, nullptr, "qk_inbox_merge", sort_output != 0);
}

int run_book_level(Call &call, const uint4 *parents, size_t n, int color_swap, uint16_t *uniq, uint64_t *uniq_mult,
                   size_t cap, size_t *n_uniq, uint64_t *h_stats)
{
    cudaStream_t s = call.stream();
    Ctx *c = call.ctx();
    DedupTable t;
    int rc = dedup_begin(call, cap, t);
    if (rc != QK_OK) return rc;
    const unsigned grid = grid_for(n * 32, kBlock, c->sms, 8);
    if (color_swap)
        k_book_expand_insert<true><<<grid, kBlock, 0, s>>>(parents, n, t.table, t.slots, t.meta, t.meta + 1,
                                                          (int *)(t.meta + 3), t.max_unique, t.meta + 4);
    else
        k_book_expand_insert<false><<<grid, kBlock, 0, s>>>(parents, n, t.table, t.slots, t.meta, t.meta + 1,
                                                           (int *)(t.meta + 3), t.max_unique, t.meta + 4);
    QK_CUDA(cudaGetLastError());
    uint64_t st[5] = { 0, 0, 0, 0, 0 };
    rc = dedup_finish(call, t, uniq, uniq_mult, cap, n_uniq, st, "qk_book_level", false);
    h_stats[0] = st[0]; h_stats[1] = st[1];
    return rc;
}

}  // namespace qk
