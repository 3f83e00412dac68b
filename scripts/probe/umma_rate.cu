// umma_rate.cu -- stand-alone: cycles per tcgen05.mma (M=128, K=16, fp16) for the operand layouts stem_tc.cu uses.
//   layout 0: plain K-major no-swizzle (LBO 128, SBO 256)      layout 1: Hankel view (LBO 16, SBO 128)
//   N = 128 and N = 64; one issuing thread per CTA, one CTA per SM, ITERS instructions accumulating into one tile
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../binary-networks-pytorch_b200/csrc/tc05.cuh"
using namespace bnn;
namespace bnn { void count_launch(int) {} }

__global__ void __launch_bounds__(128, 1) rate_kernel(int layout, int n, int iters, int pairs, long long* cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + 64);
    unsigned char* a_s = smem + 1024;             // 64 KB of operand space
    unsigned char* b_s = smem + 1024 + 65536;     // 48 KB
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (65536 + 49152) / 4; i += 128) reinterpret_cast<uint32_t*>(a_s)[i] = 0x3c003c00u;   // fp16 1.0
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (warp == 0) tc05::tmem_alloc<256>(tslot);
    fence_proxy_async();
    tc05::fence_before_sync();
    __syncthreads();
    tc05::fence_after_sync();
    const uint32_t tmem = *tslot;
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        const bool leader = tc05::elect_one();
        const uint32_t idesc = tc05::idesc_f16_f32(128, n), idesc64 = tc05::idesc_f16_f32(128, 64);
        const uint32_t au = smem_u32(a_s) >> 4, bu = smem_u32(b_s) >> 4;
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int ks = 0; ks < 12; ++ks) {
                const uint64_t ad = layout ? tc05::smem_desc_units(au + (ks % 6) * 384 + 2 * (ks & 1), 16, 128)
                                           : tc05::smem_desc_units(au + ks * 256, 128, 256);
                const uint64_t ad2 = layout ? tc05::smem_desc_units(au + (ks % 6) * 384 + 128 + 2 * (ks & 1), 16, 128)
                                            : tc05::smem_desc_units(au + ks * 256 + 2048, 128, 256);
                const uint64_t bd = tc05::smem_desc_units(bu + ks * 256, 128, 256);
                if (leader) {
                    tc05::mma_f16_ss(tmem, ad, bd, idesc, 1);
                    if (pairs) tc05::mma_f16_ss(tmem + 64, ad2, bd, idesc64, 1);
                }
            }
        }
        if (leader) tc05::commit(bar);
        __syncwarp();
        mbar_wait(bar, 0);
        t1 = clock64();
    }
    tc05::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc<256>(tmem);
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    long long* d; cudaMalloc(&d, 148 * 8);
    const int smem = 1024 + 65536 + 49152;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int ctas : {1, 148})
        for (int layout = 0; layout < 2; ++layout)
            for (int cfg = 0; cfg < 3; ++cfg) {
                const int n = cfg == 1 ? 64 : 128, pairs = cfg == 2;
                const int iters = 200;
                rate_kernel<<<ctas, 128, smem>>>(layout, n, iters, pairs, d);
                cudaError_t e = cudaDeviceSynchronize();
                std::vector<long long> h(ctas);
                cudaMemcpy(h.data(), d, ctas * 8, cudaMemcpyDeviceToHost);
                long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
                const double per = (double)mx / (iters * 12);
                printf("{\"probe\": \"umma_rate\", \"ctas\": %d, \"layout\": \"%s\", \"what\": \"%s\", \"cycles_per_kstep\": %.1f, \"cuda\": \"%s\"}\n",
                       ctas, layout ? "hankel" : "plain", pairs ? "N128 + N64 pair" : (n == 128 ? "N128" : "N64"), per, cudaGetErrorString(e));
            }
    return 0;
}
