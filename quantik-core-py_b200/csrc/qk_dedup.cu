// qk_dedup.cu -- canonicalise + merge equal canonical payloads + sum multiplicities:
// the dictionary step of game_stats.py:487-505, beam_search.py:526-546 and
// examples/generate_opening_book.py:458-496, done in HBM.
//
//   k_dedup_insert   one state per thread: canonical form in registers, then insert into
//                    an open-addressing table keyed by the 16-byte payload with a single
//                    128-bit compare-and-swap (atom.cas.b128, sm_90+), multiplicity added
//                    with a 64-bit atomic.
//   k_dedup_compact  table -> dense (key, multiplicity) arrays
//   k_bitonic_step   sorts the distinct keys in payload byte order (deterministic output)
#include "qk_common.cuh"
#include "qk_bits.cuh"

namespace qk {

static constexpr unsigned kBlock = 256;
typedef unsigned long long u64;

struct __align__(16) Slot { u64 hi, lo; };     // Key128 (a:b, c:d); all ones = empty

__device__ __forceinline__ void cas128(Slot *addr, u64 cmp_hi, u64 cmp_lo, u64 new_hi, u64 new_lo, u64 &old_hi, u64 &old_lo)
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

__global__ void __launch_bounds__(kBlock) k_dedup_fill(Slot *table, u64 *counts, size_t slots)
{
    for (size_t i = blockIdx.x * (size_t)kBlock + threadIdx.x; i < slots; i += (size_t)gridDim.x * kBlock) {
        table[i].hi = ~0ull; table[i].lo = ~0ull; counts[i] = 0;
    }
}

template <bool COLOR_SWAP>
__global__ void __launch_bounds__(kBlock) k_dedup_insert(const uint4 *__restrict__ states, const u64 *__restrict__ mult, size_t n,
                                                          Slot *table, u64 *counts, size_t slot_mask,
                                                          u64 *n_unique, u64 *all_ones_count, int *overflow, u64 max_unique)
{
    for (size_t i = blockIdx.x * (size_t)kBlock + threadIdx.x; i < n; i += (size_t)gridDim.x * kBlock) {
        uint4 v = __ldg(states + i);
        State4 s = { v.x, v.y, v.z, v.w };
        State4 c = canonical<COLOR_SWAP>(s);
        // comparable big-endian limbs of the payload bytes
        u64 hi = ((u64)prmt(c.x, 0, 0x0123) << 32) | prmt(c.y, 0, 0x0123);
        u64 lo = ((u64)prmt(c.z, 0, 0x0123) << 32) | prmt(c.w, 0, 0x0123);
        u64 m = mult ? mult[i] : 1ull;
        if ((hi & lo) == ~0ull) { atomicAdd(all_ones_count, m); continue; }   // the one payload that equals the empty marker
        size_t slot = (size_t)mix(hi, lo) & slot_mask;
        for (size_t probe = 0; probe <= slot_mask; ++probe, slot = (slot + 1) & slot_mask) {
            // fast path: one 128-bit load; a slot caught half-way through another thread's
            // CAS can only look like "some half is all ones", which falls through to the CAS
            const ulonglong2 raw = __ldcv(reinterpret_cast<const ulonglong2 *>(table + slot));
            u64 khi = raw.x, klo = raw.y;
            if (khi == ~0ull || klo == ~0ull) {
                cas128(table + slot, ~0ull, ~0ull, hi, lo, khi, klo);
                if (khi == ~0ull && klo == ~0ull) {            // we claimed the slot
                    if (atomicAdd(n_unique, 1ull) >= max_unique) *overflow = 1;
                    khi = hi; klo = lo;
                }
            }
            if (khi == hi && klo == lo) { atomicAdd(counts + slot, m); break; }
        }
    }
}

__global__ void __launch_bounds__(kBlock) k_dedup_compact(const Slot *table, const u64 *counts, size_t slots,
                                                           Slot *keys, u64 *mults, u64 *cursor)
{
    for (size_t i = blockIdx.x * (size_t)kBlock + threadIdx.x; i < slots; i += (size_t)gridDim.x * kBlock) {
        Slot k = table[i];
        if (k.hi == ~0ull && k.lo == ~0ull) continue;
        u64 o = atomicAdd(cursor, 1ull);
        keys[o] = k; mults[o] = counts[i];
    }
}

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

__global__ void __launch_bounds__(kBlock) k_emit(const Slot *keys, const u64 *mults, size_t n, uint4 *uniq, u64 *uniq_mult)
{
    for (size_t i = blockIdx.x * (size_t)kBlock + threadIdx.x; i < n; i += (size_t)gridDim.x * kBlock) {
        Slot k = keys[i];
        if (uniq) uniq[i] = make_uint4(prmt((uint32_t)(k.hi >> 32), 0, 0x0123), prmt((uint32_t)k.hi, 0, 0x0123),
                                       prmt((uint32_t)(k.lo >> 32), 0, 0x0123), prmt((uint32_t)k.lo, 0, 0x0123));
        if (uniq_mult) uniq_mult[i] = mults[i];
    }
}

int run_dedup(Call &call, const uint4 *states, const uint64_t *mult, size_t n, int color_swap,
              uint16_t *uniq, uint64_t *uniq_mult, size_t cap, size_t *n_uniq)
{
    cudaStream_t s = call.stream();
    Ctx *c = call.ctx();
    size_t want = n < cap || cap == 0 ? n : cap;
    size_t slots = 1024;
    while (slots < 2 * want + 2) slots <<= 1;
    Slot *table = (Slot *)call.temp(slots * sizeof(Slot));
    u64 *counts = (u64 *)call.temp(slots * sizeof(u64));
    u64 *meta = (u64 *)call.temp(4 * sizeof(u64), true);   // n_unique, all-ones multiplicity, cursor, overflow flag
    if (!table || !counts || !meta) return call.status();
    k_dedup_fill<<<grid_for(slots, kBlock, c->sms, 8), kBlock, 0, s>>>(table, counts, slots);
    const u64 max_unique = (u64)(slots / 2);
    if (color_swap)
        k_dedup_insert<true><<<grid_for(n, kBlock, c->sms, 8), kBlock, 0, s>>>(states, (const u64 *)mult, n, table, counts, slots - 1,
                                                                             meta, meta + 1, (int *)(meta + 3), max_unique);
    else
        k_dedup_insert<false><<<grid_for(n, kBlock, c->sms, 8), kBlock, 0, s>>>(states, (const u64 *)mult, n, table, counts, slots - 1,
                                                                              meta, meta + 1, (int *)(meta + 3), max_unique);
    QK_CUDA(cudaGetLastError());
    u64 h_meta[4];
    QK_CUDA(cudaMemcpyAsync(h_meta, meta, sizeof h_meta, cudaMemcpyDeviceToHost, s));
    QK_CUDA(cudaStreamSynchronize(s));
    size_t nu = (size_t)h_meta[0];
    bool special = h_meta[1] != 0;
    if ((int)h_meta[3] != 0 || (cap && nu + (special ? 1 : 0) > cap && (uniq || uniq_mult))) {
        set_error("qk_canon_dedup: more distinct canonical states than `cap`");
        return QK_E_BAD_ARG;
    }
    size_t padded = 1;
    while (padded < nu) padded <<= 1;
    Slot *keys = (Slot *)call.temp((padded + 1) * sizeof(Slot));
    u64 *mults = (u64 *)call.temp((padded + 1) * sizeof(u64));
    if (!keys || !mults) return call.status();
    k_dedup_compact<<<grid_for(slots, kBlock, c->sms, 8), kBlock, 0, s>>>(table, counts, slots, keys, mults, meta + 2);
    k_pad<<<grid_for(padded - nu + 1, kBlock, c->sms, 8), kBlock, 0, s>>>(keys, mults, nu, padded);
    for (size_t k = 2; k <= padded; k <<= 1)
        for (size_t j = k >> 1; j > 0; j >>= 1)
            k_bitonic_step<<<grid_for(padded, kBlock, c->sms, 8), kBlock, 0, s>>>(keys, mults, padded, j, k);
    QK_CUDA(cudaGetLastError());
    if (special) {   // the all-ones payload sorts last by construction
        Slot ones = { ~0ull, ~0ull };
        QK_CUDA(cudaMemcpyAsync(keys + nu, &ones, sizeof ones, cudaMemcpyHostToDevice, s));
        QK_CUDA(cudaMemcpyAsync(mults + nu, &h_meta[1], sizeof(u64), cudaMemcpyHostToDevice, s));
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

}  // namespace qk
