// Micro-benchmark: throughput of random 16-byte loads, 64-bit REDs, 64-bit CAS and 128-bit CAS on a table that
// fits L2 (32 MiB) and one that does not (1 GiB).  nvcc -arch=sm_100a -O3 -o atomics_rate atomics_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
struct __align__(32) TSlot { u64 hi, lo, count, pad; };
__device__ __forceinline__ u64 mix(u64 x) { x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32; x *= 0x94D049BB133111EBull; x ^= x >> 29; return x; }
__device__ __forceinline__ void cas128(void *addr, u64 ch, u64 cl, u64 nh, u64 nl, u64 &oh, u64 &ol)
{
    asm volatile("{\n\t.reg .b128 c, s, o;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 s, {%4, %5};\n\tatom.global.cas.b128 o, [%6], c, s;\n\tmov.b128 {%0, %1}, o;\n\t}\n"
                 : "=l"(oh), "=l"(ol) : "l"(ch), "l"(cl), "l"(nh), "l"(nl), "l"(addr) : "memory");
}
template <int MODE>
__global__ void k(TSlot *t, size_t slots, size_t n, u64 *sink)
{
    u64 acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const u64 h = mix(i * 0x9E3779B97F4A7C15ull + 12345);
        TSlot *p = t + (size_t)__umul64hi(h, (u64)slots);
        if (MODE == 0) { const ulonglong2 v = __ldcv(reinterpret_cast<const ulonglong2 *>(p)); acc += v.x ^ v.y; }
        if (MODE == 1) atomicAdd(&p->count, 1ull);
        if (MODE == 2) acc += atomicCAS(&p->hi, ~0ull, h | 1);
        if (MODE == 3) { u64 oh, ol; cas128(p, ~0ull, ~0ull, h | 1, h, oh, ol); acc += oh ^ ol; }
        if (MODE == 4) acc += atomicAdd(&p->count, 1ull);       // ATOM (value returned) instead of RED
        if (MODE == 5) {                                        // the insert's pattern: probe, then RED on the same sector
            const ulonglong2 v = __ldcv(reinterpret_cast<const ulonglong2 *>(p));
            atomicAdd(&p->count, 1ull + ((v.x ^ v.y) == 0x1234567ull ? 1ull : 0ull));
        }
    }
    if (acc == 0x1234567) *sink = acc;
}
int main(int argc, char **argv)
{
    const size_t n = 1u << 24;
    if (argc > 1) {     // size sweep: where does the rate fall -- at the L2 capacity, or further out (TLB reach)?
        u64 *sink2; cudaMalloc(&sink2, 8);
        for (size_t mb : { (size_t)32, (size_t)64, (size_t)96, (size_t)128, (size_t)192, (size_t)256, (size_t)512, (size_t)1024, (size_t)2048, (size_t)4096, (size_t)8192 }) {
            const size_t slots = (mb << 20) / sizeof(TSlot);
            TSlot *t; if (cudaMalloc(&t, slots * sizeof(TSlot)) != cudaSuccess) break;
            float best[3] = { 1e9f, 1e9f, 1e9f };
            for (int mode = 0; mode < 3; ++mode)
                for (int rep = 0; rep < 3; ++rep) {
                    cudaMemset(t, 0xFF, slots * sizeof(TSlot));
                    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
                    cudaEventRecord(a);
                    if (mode == 0) k<0><<<148 * 8, 256>>>(t, slots, n, sink2);
                    if (mode == 1) k<1><<<148 * 8, 256>>>(t, slots, n, sink2);
                    if (mode == 2) k<5><<<148 * 8, 256>>>(t, slots, n, sink2);
                    cudaEventRecord(b); cudaEventSynchronize(b);
                    float ms; cudaEventElapsedTime(&ms, a, b);
                    if (ms < best[mode]) best[mode] = ms;
                }
            printf("table %5zu MiB  ld.16B %.3e/s   red.64 %.3e/s   ld+red same sector %.3e/s\n", mb, n / (best[0] * 1e-3), n / (best[1] * 1e-3), n / (best[2] * 1e-3));
            cudaFree(t);
        }
        return 0;
    }
    u64 *sink; cudaMalloc(&sink, 8);
    for (size_t mb : { (size_t)32, (size_t)1024 }) {
        const size_t slots = (mb << 20) / sizeof(TSlot);
        TSlot *t; cudaMalloc(&t, slots * sizeof(TSlot));
        const char *names[5] = { "ld.16B", "red.64", "cas.64", "cas.128", "atom.add.64" };
        for (int mode = 0; mode < 5; ++mode) {
            float best = 1e9;
            for (int rep = 0; rep < 3; ++rep) {
                cudaMemset(t, 0xFF, slots * sizeof(TSlot));
                cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
                cudaEventRecord(a);
                switch (mode) {
                case 0: k<0><<<148 * 8, 256>>>(t, slots, n, sink); break;
                case 1: k<1><<<148 * 8, 256>>>(t, slots, n, sink); break;
                case 2: k<2><<<148 * 8, 256>>>(t, slots, n, sink); break;
                case 3: k<3><<<148 * 8, 256>>>(t, slots, n, sink); break;
                case 4: k<4><<<148 * 8, 256>>>(t, slots, n, sink); break;
                }
                cudaEventRecord(b); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b);
                if (ms < best) best = ms;
            }
            printf("table %4zu MiB  %-12s %8.3f ms  %.3e ops/s\n", mb, names[mode], best, n / (best * 1e-3));
        }
        cudaFree(t);
    }
    return 0;
}
