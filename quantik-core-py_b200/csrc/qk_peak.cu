// qk_peak.cu -- measured integer-issue roofline denominators: a
// dependency-free stream of (a) LOP3 only, (b) LOP3 + IMAD interleaved, timed with CUDA
// events.  Reported as lane-operations per second (warp instructions x 32).
#include "qk_common.cuh"

namespace qk {

static constexpr int kIters = 4096;
static constexpr int kUnroll = 8;        // 8 independent chains per thread

template <bool MIXED>
__global__ void __launch_bounds__(256) k_issue(uint32_t *out, uint32_t b, uint32_t c)
{
    uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
#pragma unroll 1
    for (int i = 0; i < kIters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(b), "r"(c));
            if (MIXED) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a1) : "r"(b), "r"(c));
            else       asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a1) : "r"(b), "r"(c));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a2) : "r"(b), "r"(c));
            if (MIXED) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a3) : "r"(b), "r"(c));
            else       asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a3) : "r"(b), "r"(c));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a4) : "r"(b), "r"(c));
            if (MIXED) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a5) : "r"(b), "r"(c));
            else       asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a5) : "r"(b), "r"(c));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a6) : "r"(b), "r"(c));
            if (MIXED) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a7) : "r"(b), "r"(c));
            else       asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a7) : "r"(b), "r"(c));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}

template <bool MIXED>
static int time_issue(Ctx *c, uint32_t *buf, double *lane_ops_per_s)
{
    const unsigned grid = (unsigned)c->sms * 8, block = 256;
    cudaEvent_t e0, e1;
    QK_CUDA(cudaEventCreate(&e0));
    QK_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        QK_CUDA(cudaEventRecord(e0, 0));
        k_issue<MIXED><<<grid, block>>>(buf, 0x9E3779B9u + rep, 0x7F4A7C15u);
        QK_CUDA(cudaEventRecord(e1, 0));
        QK_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        QK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    double ops = (double)grid * block * (double)kIters * 4 * kUnroll;
    *lane_ops_per_s = ops / (best * 1e-3);
    return QK_OK;
}

int run_issue_peak(Ctx *c, double *alu, double *mixed)
{
    uint32_t *buf = nullptr;
    QK_CUDA(cudaMalloc(&buf, (size_t)c->sms * 8 * 256 * sizeof(uint32_t)));
    int rc = time_issue<false>(c, buf, alu);
    if (rc == QK_OK) rc = time_issue<true>(c, buf, mixed);
    cudaFree(buf);
    return rc;
}

}  // namespace qk
