// ubench.cu -- integer-pipe micro-benchmarks for the popcount roofline denominator.
// "P_popc ... verify for cc 10.0 by micro-benchmark before fixing the
// roofline denominator".  Every mode keeps 8 independent dependency chains per thread.
#include "common.cuh"

namespace bnn {

constexpr int UB_CHAINS = 8;
constexpr int UB_INNER = 64;

template <int MODE>
__global__ void __launch_bounds__(256) ubench_kernel(uint32_t* sink, int iters, uint32_t seed) {
    uint32_t a[UB_CHAINS], s[UB_CHAINS];
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t m = 0xffff0f0fu ^ seed, t = 0x9e3779b9u * (tid | 1u);
#pragma unroll
    for (int i = 0; i < UB_CHAINS; ++i) { a[i] = tid * 2654435761u + i * 40503u + seed; s[i] = a[i] ^ 0x5bd1e995u; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < UB_INNER; ++k) {
            if (MODE == 0) {            // POPC only
#pragma unroll
                for (int i = 0; i < UB_CHAINS; ++i) a[i] = __popc(a[i]);
            } else if (MODE == 1) {     // LOP3 only
#pragma unroll
                for (int i = 0; i < UB_CHAINS; ++i) a[i] = m & (a[i] ^ t);
            } else if (MODE == 2) {     // one word: LOP3 + POPC + IADD
#pragma unroll
                for (int i = 0; i < UB_CHAINS; ++i) { a[i] += __popc(m & (s[i] ^ t)); s[i] = s[i] * 3u + a[i]; }
            } else if (MODE == 3) {     // 3 words: 3 LOP3 + 2 LOP3 (3:2 CSA) + 2 POPC
#pragma unroll
                for (int i = 0; i < UB_CHAINS; ++i) {
                    const uint32_t x0 = m & (s[i] ^ t), x1 = m & (s[i] ^ a[i]), x2 = t & (s[i] ^ m);
                    a[i] += __popc(x0 ^ x1 ^ x2) + 2 * __popc((x0 & x1) | (x2 & (x0 ^ x1)));
                    s[i] = s[i] * 3u + a[i];
                }
            } else {                    // 7 words: 7 LOP3 + 4 CSA (8 LOP3) + 3 POPC
#pragma unroll
                for (int i = 0; i < UB_CHAINS; ++i) {
                    const uint32_t v = s[i];
                    const uint32_t x0 = m & (v ^ t), x1 = m & (v ^ a[i]), x2 = t & (v ^ m), x3 = a[i] & (v ^ t),
                                   x4 = m & (v ^ ~t), x5 = t & (v ^ a[i]), x6 = ~m & (v ^ t);
                    const uint32_t s1 = x0 ^ x1 ^ x2, c1 = (x0 & x1) | (x2 & (x0 ^ x1));
                    const uint32_t s2 = x3 ^ x4 ^ x5, c2 = (x3 & x4) | (x5 & (x3 ^ x4));
                    const uint32_t s3 = s1 ^ s2 ^ x6, c3 = (s1 & s2) | (x6 & (s1 ^ s2));
                    const uint32_t tw = c1 ^ c2 ^ c3, fo = (c1 & c2) | (c3 & (c1 ^ c2));
                    a[i] += __popc(s3) + 2 * __popc(tw) + 4 * __popc(fo);
                    s[i] = v * 3u + a[i];
                }
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < UB_CHAINS; ++i) r ^= a[i] ^ s[i];
    if (r == 0x12345678u) sink[0] = r;   // never true in practice; keeps the chains alive
}

template <int MODE>
static int run_mode(int iters, double words_per_inner, double* gops) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint32_t* sink = nullptr;
    cudaError_t e = cudaMalloc(&sink, 4);
    if (e != cudaSuccess) return (int)e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 8, threads = 256;
    ubench_kernel<MODE><<<blocks, threads>>>(sink, 4, 1u);      // warm-up
    cudaEventRecord(e0);
    ubench_kernel<MODE><<<blocks, threads>>>(sink, iters, 2u);
    cudaEventRecord(e1);
    e = cudaEventSynchronize(e1);
    count_launch(2);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(sink);
    if (e != cudaSuccess) return (int)e;
    const double ops = (double)blocks * threads * (double)iters * UB_INNER * UB_CHAINS * words_per_inner;
    *gops = ops / (ms * 1e-3) * 1e-9;
    return (int)cudaGetLastError();
}


// Floating-point / legacy tensor-path rates (what a faster fp32 stem could be built on): MODE 5 = mma.sync m16n8k8
// tf32, 6 = mma.sync m16n8k16 f16, 7 = mma.sync m16n8k16 bf16 (all fp32 accumulate, SASS HMMA), 8 = fma.rn.f32x2
// (FFMA2), 9 = scalar FFMA.  8 independent accumulator sets per thread.
template <int MODE>
__global__ void __launch_bounds__(256) ubench_fp_kernel(float* sink, int iters, float seed) {
    float acc[UB_CHAINS][4];
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < UB_CHAINS; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = seed * (float)(i + j);
    uint32_t a0 = tid * 2654435761u, a1 = a0 ^ 0x3c003c00u, a2 = a0 + 77u, a3 = a1 + 99u, b0 = 0x3c003800u, b1 = 0x38003c00u;
    if (MODE == 5) { a0 = __float_as_uint(1.0f); a1 = __float_as_uint(0.5f); a2 = a0; a3 = a1; b0 = a0; b1 = a1; }
    const float fa = 1.0f + seed, fb = 0.25f * seed;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < UB_INNER; ++k) {
#pragma unroll
            for (int i = 0; i < UB_CHAINS; ++i) {
                if (MODE == 5) {
                    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
                } else if (MODE == 6) {
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
                } else if (MODE == 7) {
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
                } else if (MODE == 8) {
                    unsigned long long x, y, w, z;
                    asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(acc[i][0]), "f"(acc[i][1]));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(acc[i][2]), "f"(acc[i][3]));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(w) : "f"(fa), "f"(fa));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(z) : "f"(fb), "f"(fb));
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(w), "l"(z));
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(y) : "l"(w), "l"(z));
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[i][0]), "=f"(acc[i][1]) : "l"(x));
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[i][2]), "=f"(acc[i][3]) : "l"(y));
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[i][j]) : "f"(fa), "f"(fb));
                }
            }
        }
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < UB_CHAINS; ++i) r += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
    if (r == 12345.678f) sink[0] = r;
}

template <int MODE>
static int run_fp_mode(int iters, double fma_per_thread_inner, double* gops) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    float* sink = nullptr;
    cudaError_t e = cudaMalloc(&sink, 4);
    if (e != cudaSuccess) return (int)e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 8, threads = 256;
    ubench_fp_kernel<MODE><<<blocks, threads>>>(sink, 4, 1.0f);
    cudaEventRecord(e0);
    ubench_fp_kernel<MODE><<<blocks, threads>>>(sink, iters, 0.5f);
    cudaEventRecord(e1);
    e = cudaEventSynchronize(e1);
    count_launch(2);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(sink);
    if (e != cudaSuccess) return (int)e;
    const double ops = (double)blocks * threads * (double)iters * UB_INNER * UB_CHAINS * fma_per_thread_inner;
    *gops = ops / (ms * 1e-3) * 1e-9;      // G multiply-adds / s
    return (int)cudaGetLastError();
}

}  // namespace bnn

extern "C" int bnn_ubench(int32_t which, int32_t iters, double* gops) {
    if (!gops) return BNN_E_NULL;
    if (iters <= 0) return BNN_E_SHAPE;
    switch (which) {
        case 0: return bnn::run_mode<0>(iters, 1.0, gops);
        case 1: return bnn::run_mode<1>(iters, 1.0, gops);
        case 2: return bnn::run_mode<2>(iters, 1.0, gops);
        case 3: return bnn::run_mode<3>(iters, 3.0, gops);
        case 4: return bnn::run_mode<4>(iters, 7.0, gops);
        case 5: return bnn::run_fp_mode<5>(iters, 32.0, gops);     // 16*8*8 fma per warp instruction
        case 6: return bnn::run_fp_mode<6>(iters, 64.0, gops);
        case 7: return bnn::run_fp_mode<7>(iters, 64.0, gops);
        case 8: return bnn::run_fp_mode<8>(iters, 4.0, gops);
        case 9: return bnn::run_fp_mode<9>(iters, 4.0, gops);
        default: return BNN_E_SHAPE;
    }
}
