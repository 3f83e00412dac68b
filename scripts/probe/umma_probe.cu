// umma_probe.cu -- stand-alone check of the tcgen05 wrappers in csrc/tc05.cuh on a B200 (not part of the library).
//   test 0: M=128 N=128 K=32 fp16 MMA (two K=16 instructions), A and B in the plain K-major no-swizzle canonical layout
//   test 1: the same product with A addressed as a Hankel view  A[m][16-byte chunk j] = X[m + j]  of a 1-D array of
//           16-byte units (LBO = 16 B, SBO = 128 B: overlapping core matrices) -- the implicit-im2col trick of stem_tc.cu
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu ; run: ./umma_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <vector>
#include "../../binary-networks-pytorch_b200/csrc/tc05.cuh"

using namespace bnn;

namespace bnn { void count_launch(int) {} }

__global__ void __launch_bounds__(128, 1) probe_kernel(const __half* xa, const __half* xb, float* d, int hankel) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + 64);
    __half* a_s = reinterpret_cast<__half*>(smem + 1024);            // 16 KB region
    __half* b_s = reinterpret_cast<__half*>(smem + 1024 + 16384);    // 128 rows x 32 halfs = 8 KB
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (warp == 0) tc05::tmem_alloc<128>(tslot);
    // A: hankel == 0: canonical [k16 step][m/8][chunk j][m%8][8 halfs]: LBO = 128, SBO = 256, 4 KB per k16 step
    //    hankel == 1: units X[u] (u < 160), A[m][k16 step s][chunk j] = X[m + 2 s + j]
    if (!hankel) {
        for (int i = tid; i < 128 * 32; i += 128) {
            const int m = i / 32, k = i % 32, s = k / 16, j = (k % 16) / 8, e = k % 8;
            a_s[s * 2048 + (m / 8) * 128 + j * 64 + (m % 8) * 8 + e] = xa[m * 32 + k];
        }
    } else {
        for (int i = tid; i < 160 * 8; i += 128) a_s[i] = xa[i];
    }
    for (int i = tid; i < 128 * 32; i += 128) {
        const int n = i / 32, k = i % 32, s = k / 16, j = (k % 16) / 8, e = k % 8;
        b_s[s * 2048 + (n / 8) * 128 + j * 64 + (n % 8) * 8 + e] = xb[n * 32 + k];
    }
    fence_proxy_async();
    tc05::fence_before_sync();
    __syncthreads();
    tc05::fence_after_sync();
    const uint32_t tmem = *tslot;
    if (tid == 0) {
        const uint32_t idesc = tc05::idesc_f16_f32(128, 128);
        for (int s = 0; s < 2; ++s) {
            const uint64_t ad = hankel ? tc05::smem_desc(smem_u32(a_s) + s * 32, 16, 128)
                                       : tc05::smem_desc(smem_u32(a_s) + s * 4096, 128, 256);
            const uint64_t bd = tc05::smem_desc(smem_u32(b_s) + s * 4096, 128, 256);
            tc05::mma_f16_ss(tmem, ad, bd, idesc, s > 0);
        }
        tc05::commit(bar);
    }
    mbar_wait(bar, 0);
    tc05::fence_after_sync();
    for (int c = 0; c < 128; c += 32) {
        float v0[16], v1[16];
        const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + c;
        tc05::tmem_ld16x2_sync(ta, ta + 16, v0, v1);
        for (int i = 0; i < 16; ++i) { d[(warp * 32 + lane) * 128 + c + i] = v0[i]; d[(warp * 32 + lane) * 128 + c + 16 + i] = v1[i]; }
    }
    tc05::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc<128>(tmem);
}

int main() {
    int bad = 0;
    for (int hankel = 0; hankel < 2; ++hankel) {
        std::vector<__half> xa(hankel ? 160 * 8 : 128 * 32), xb(128 * 32);
        srand(7 + hankel);
        for (auto& v : xa) v = __float2half((float)(rand() % 2001 - 1000) / 500.0f);
        for (auto& v : xb) v = __float2half((float)(rand() % 2001 - 1000) / 500.0f);
        __half *dxa, *dxb; float* dd;
        cudaMalloc(&dxa, xa.size() * 2); cudaMalloc(&dxb, xb.size() * 2); cudaMalloc(&dd, 128 * 128 * 4);
        cudaMemcpy(dxa, xa.data(), xa.size() * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(dxb, xb.data(), xb.size() * 2, cudaMemcpyHostToDevice);
        cudaMemset(dd, 0xff, 128 * 128 * 4);
        const int smem = 1024 + 16384 + 8192;
        cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        probe_kernel<<<1, 128, smem>>>(dxa, dxb, dd, hankel);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> d(128 * 128);
        cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 128; ++n) {
                double ref = 0;
                for (int k = 0; k < 32; ++k) {
                    const float a = hankel ? __half2float(xa[(m + k / 8) * 8 + k % 8]) : __half2float(xa[m * 32 + k]);
                    ref += (double)a * (double)__half2float(xb[n * 32 + k]);
                }
                const double err = fabs(ref - (double)d[m * 128 + n]);
                if (!(err <= maxerr)) maxerr = err;
                if (fabs(ref) > maxref) maxref = fabs(ref);
            }
        printf("{\"probe\": \"umma\", \"hankel\": %d, \"cuda\": \"%s\", \"max_abs_err\": %.3e, \"max_ref\": %.3e, \"d00\": %.4f}\n",
               hankel, cudaGetErrorString(e), maxerr, maxref, d[0]);
        if (e != cudaSuccess || !(maxerr <= 1e-3 * maxref)) ++bad;
        if (e != cudaSuccess) break;
        cudaFree(dxa); cudaFree(dxb); cudaFree(dd);
    }
    return bad;
}
