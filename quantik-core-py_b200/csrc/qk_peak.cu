// qk_peak.cu -- measured integer-issue roofline denominators: dependency-free
// instruction streams timed with CUDA events, reported as lane-operations per second
// (warp instructions x 32).  Modes: 0 LOP3  1 IMAD (mad.lo)  2 IMAD.HI (mul.hi)  3 IMAD.WIDE
// 4 POPC  5 LOP3+IMAD 1:1  6 LOP3+IMAD.HI 1:1  7 SHF+LOP3 1:1 (both on the ALU pipe)
#include "qk_common.cuh"

namespace qk {

static constexpr int kIters = 4096;

template <int MODE>
__device__ __forceinline__ void op(uint32_t &a, uint64_t &wide, uint32_t b, uint32_t c, int slot)
{
    const int m = MODE <= 4 ? MODE : (MODE == 5 ? (slot & 1 ? 1 : 0) : (MODE == 6 ? (slot & 1 ? 2 : 0) : (slot & 1 ? 8 : 0)));
    if (m == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
    else if (m == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    else if (m == 2) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a) : "r"(b));
    else if (m == 3) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(wide) : "r"((uint32_t)wide), "r"(b));
    else if (m == 4) asm volatile("popc.b32 %0, %0;" : "+r"(a));
    else asm volatile("shf.r.clamp.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
}

template <int MODE>
__global__ void __launch_bounds__(256) k_issue(uint32_t *out, uint32_t b, uint32_t c)
{
    uint32_t a[8];
    uint64_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x + i; w[i] = 0; }
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) op<MODE>(a[i], w[i], b, c, i);
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= a[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
static int time_issue(Ctx *c, uint32_t *buf, double *lane_ops_per_s)
{
    const unsigned grid = (unsigned)c->sms * 8, block = 256;
    cudaEvent_t e0, e1;
    QK_CUDA(cudaEventCreate(&e0));
    QK_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        QK_CUDA(cudaEventRecord(e0, 0));
        k_issue<MODE><<<grid, block>>>(buf, 0x9E3779B9u + rep, 3u);
        QK_CUDA(cudaEventRecord(e1, 0));
        QK_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        QK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    double ops = (double)grid * block * (double)kIters * 4 * 8;
    *lane_ops_per_s = ops / (best * 1e-3);
    return QK_OK;
}

int run_pipe_rates(Ctx *c, double *out8)
{
    uint32_t *buf = nullptr;
    QK_CUDA(cudaMalloc(&buf, (size_t)c->sms * 8 * 256 * sizeof(uint32_t)));
    int rc = time_issue<0>(c, buf, out8 + 0);
    if (rc == QK_OK) rc = time_issue<1>(c, buf, out8 + 1);
    if (rc == QK_OK) rc = time_issue<2>(c, buf, out8 + 2);
    if (rc == QK_OK) rc = time_issue<3>(c, buf, out8 + 3);
    if (rc == QK_OK) rc = time_issue<4>(c, buf, out8 + 4);
    if (rc == QK_OK) rc = time_issue<5>(c, buf, out8 + 5);
    if (rc == QK_OK) rc = time_issue<6>(c, buf, out8 + 6);
    if (rc == QK_OK) rc = time_issue<7>(c, buf, out8 + 7);
    cudaFree(buf);
    return rc;
}

int run_issue_peak(Ctx *c, double *alu, double *mixed)
{
    uint32_t *buf = nullptr;
    QK_CUDA(cudaMalloc(&buf, (size_t)c->sms * 8 * 256 * sizeof(uint32_t)));
    int rc = time_issue<0>(c, buf, alu);
    if (rc == QK_OK) rc = time_issue<5>(c, buf, mixed);
    cudaFree(buf);
    return rc;
}

}  // namespace qk
