// Micro-benchmark of the L2-resident insert: 800 K keys (3 copies of each of 270 K) into a 1.2 M-slot table,
// 592 x 256 threads, batches of 4, keys (a) computed, (b) loaded from per-block ranges.  Isolates where
// k_part_insert's 50 us go.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o insert_rate insert_rate.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
struct __align__(32) TSlot { u64 hi, lo, count, pad; };
struct __align__(16) Slot { u64 hi, lo; };
__device__ __forceinline__ u64 mix(u64 h, u64 l)
{
    u64 x = h * 0x9E3779B97F4A7C15ull ^ (l + 0xD6E8FEB86659FD93ull) * 0xC2B2AE3D27D4EB4Full;
    x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
    return x;
}
__device__ __forceinline__ void cas128(void *addr, u64 ch, u64 cl, u64 nh, u64 nl, u64 &oh, u64 &ol)
{
    asm volatile("{\n\t.reg .b128 c, s, o;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 s, {%4, %5};\n\tatom.global.cas.b128 o, [%6], c, s;\n\tmov.b128 {%0, %1}, o;\n\t}\n"
                 : "=l"(oh), "=l"(ol) : "l"(ch), "l"(cl), "l"(nh), "l"(nl), "l"(addr) : "memory");
}
__device__ __forceinline__ Slot make_key(unsigned id) { Slot k = { mix(id, 77) | 1ull, mix(id, 99) & ~1ull }; return k; }

template <bool LOAD, int BATCH>
__global__ void __launch_bounds__(256) k_insert(const Slot *keys, unsigned per_block, unsigned cap, unsigned uniq, TSlot *table, uint32_t slots, u64 *claims_out)
{
    unsigned claims = 0;
    const size_t base = (size_t)blockIdx.x * cap;
    for (unsigned i0 = threadIdx.x; i0 < per_block; i0 += 256 * BATCH) {
        Slot key[BATCH]; ulonglong2 raw[BATCH]; uint32_t slot[BATCH]; unsigned pending = 0;
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {
            const unsigned i = i0 + k * 256;
            key[k].hi = key[k].lo = ~0ull;
            if (i < per_block) { key[k] = LOAD ? keys[base + i] : make_key((blockIdx.x * per_block + i) % uniq); pending |= 1u << k; }
        }
#pragma unroll
        for (int k = 0; k < BATCH; ++k) slot[k] = __umulhi((uint32_t)mix(key[k].hi, key[k].lo), slots);
        for (uint32_t round = 0; pending && round < 8192; ++round) {
#pragma unroll
            for (int k = 0; k < BATCH; ++k) if ((pending >> k) & 1u) raw[k] = __ldcv(reinterpret_cast<const ulonglong2 *>(table + slot[k]));
            unsigned tried = 0;
#pragma unroll
            for (int k = 0; k < BATCH; ++k)
                if (((pending >> k) & 1u) && (raw[k].x == ~0ull || raw[k].y == ~0ull)) { tried |= 1u << k; cas128(table + slot[k], ~0ull, ~0ull, key[k].hi, key[k].lo, raw[k].x, raw[k].y); }
#pragma unroll
            for (int k = 0; k < BATCH; ++k) {
                if (!((pending >> k) & 1u)) continue;
                const bool claimed = ((tried >> k) & 1u) && raw[k].x == ~0ull && raw[k].y == ~0ull;
                if (claimed || (raw[k].x == key[k].hi && raw[k].y == key[k].lo)) { atomicAdd(&table[slot[k]].count, 1ull); claims += claimed; pending &= ~(1u << k); }
                else slot[k] = slot[k] + 1 == slots ? 0 : slot[k] + 1;
            }
        }
    }
    claims = __reduce_add_sync(0xFFFFFFFFu, claims);
    if ((threadIdx.x & 31) == 0 && claims) atomicAdd(claims_out, (u64)claims);
}
__global__ void k_fill_keys(Slot *keys, unsigned per_block, unsigned cap, unsigned uniq)
{
    for (unsigned i = threadIdx.x; i < per_block; i += blockDim.x) keys[(size_t)blockIdx.x * cap + i] = make_key((blockIdx.x * per_block + i) % uniq);
}
int main(int argc, char **argv)
{
    // insert_rate [inserts per block]: 1350 = the size of one part of qk_canon_dedup's two-level merge; larger values show
    // the rate the table sustains once the kernel is long enough to amortise its launch and its first probes
    const unsigned blocks = 592, per_block = argc > 1 ? (unsigned)atoi(argv[1]) : 1350, cap = per_block + 250, uniq = 270000;
    const uint32_t slots = 1200128;
    TSlot *table; cudaMalloc(&table, (size_t)slots * 32);
    Slot *keys; cudaMalloc(&keys, (size_t)blocks * cap * 16);
    u64 *claims; cudaMalloc(&claims, 8);
    k_fill_keys<<<blocks, 256>>>(keys, per_block, cap, uniq);
    for (int variant = 0; variant < 6; ++variant) {
        float best = 1e9; u64 h = 0;
        for (int rep = 0; rep < 4; ++rep) {
            cudaMemset(table, 0xFF, (size_t)slots * 32); cudaMemset(claims, 0, 8);
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            cudaEventRecord(a);
            switch (variant) {
            case 0: k_insert<false, 4><<<blocks, 256>>>(keys, per_block, cap, uniq, table, slots, claims); break;
            case 1: k_insert<true, 4><<<blocks, 256>>>(keys, per_block, cap, uniq, table, slots, claims); break;
            case 2: k_insert<true, 1><<<blocks, 256>>>(keys, per_block, cap, uniq, table, slots, claims); break;
            case 3: k_insert<true, 2><<<blocks, 256>>>(keys, per_block, cap, uniq, table, slots, claims); break;
            case 4: k_insert<true, 6><<<blocks, 256>>>(keys, per_block, cap, uniq, table, slots, claims); break;
            case 5: k_insert<true, 4><<<blocks * 2, 256>>>(keys, per_block / 2, cap / 2, uniq, table, slots, claims); break;
            }
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
            cudaMemcpy(&h, claims, 8, cudaMemcpyDeviceToHost);
        }
        const char *names[6] = { "computed keys, batch 4", "loaded keys, batch 4", "loaded, batch 1", "loaded, batch 2", "loaded, batch 6", "loaded, batch 4, 2x blocks of half size" };
        printf("%-42s %8.1f us   claims %llu   %.3e inserts/s\n", names[variant], best * 1e3, h, (double)blocks * per_block / (best * 1e-3));
    }
    return 0;
}
