// umma_row.cu -- stand-alone: cycles per "conv row" of stem_tc.cu's MMA stream (25 instructions, three accumulators) with
// 0 / 1 / 2 tcgen05.commit per row, and with the Hankel A operands rotating through 6 ring slots as in the kernel.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../binary-networks-pytorch_b200/csrc/tc05.cuh"
using namespace bnn;
namespace bnn { void count_launch(int) {} }

constexpr int NU = 128, SLOT = 3 * 2 * NU * 16, BSTEP = 4096;

__global__ void __launch_bounds__(128, 1) row_kernel(int commits, int rotate, int stages, int rows, long long* cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);       // [8]
    uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + 128);
    unsigned char* b_s = smem + 2048;
    unsigned char* ring = b_s + 12 * BSTEP;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (12 * BSTEP + 6 * SLOT) / 4; i += 128) reinterpret_cast<uint32_t*>(b_s)[i] = 0x3c003c00u;
    if (tid == 0) { for (int i = 0; i < 8; ++i) mbar_init(bars + i, 1); fence_mbar_init(); }
    if (warp == 0) tc05::tmem_alloc<512>(tslot);
    fence_proxy_async();
    tc05::fence_before_sync();
    __syncthreads();
    tc05::fence_after_sync();
    const uint32_t tmem = *tslot;
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        const bool leader = tc05::elect_one();
        const uint32_t idesc128 = tc05::idesc_f16_f32(128, 128), idesc64 = tc05::idesc_f16_f32(128, 64);
        const uint32_t ring_lo = tc05::desc_lo(smem_u32(ring) >> 4, 16), b_lo0 = tc05::desc_lo(smem_u32(b_s) >> 4, 128);
        constexpr uint32_t A_HI = tc05::desc_hi(128), B_HI = tc05::desc_hi(256);
        t0 = clock64();
        for (int r = 0; r < rows; ++r) {
            const int s0 = rotate ? r % 6 : 0, s2 = rotate ? (r + 2) % 6 : 2;
            const uint32_t a0 = ring_lo + (uint32_t)s0 * (SLOT / 16), a2 = ring_lo + (uint32_t)s2 * (SLOT / 16);
            const uint32_t dst = tmem + (uint32_t)((r % stages) * 256);
#pragma unroll
            for (int ks = 0; ks < 12; ++ks) {
                const int grp = ks / 6, ci = (ks >> 1) % 3, pp = ks & 1;
                const uint32_t au = (grp ? a2 : a0) + (uint32_t)(2 * pp);
                const uint32_t a_hi = au + (uint32_t)(ci * 2 + 0) * NU, a_lo = au + (uint32_t)(ci * 2 + 1) * NU;
                const uint32_t b_all = b_lo0 + (uint32_t)ks * (BSTEP / 16), b_wh = b_all + (grp ? 128u : 0u);
                if (leader) {
                    if (ks == 6) {
                        tc05::mma_f16_ss_w(dst + 128u, a_hi, A_HI, b_wh, B_HI, idesc64, 0);
                        tc05::mma_f16_ss_w(dst + 64u, a_hi, A_HI, b_all, B_HI, idesc64, 1);
                    } else {
                        tc05::mma_f16_ss_w(dst + (grp ? 64u : 0u), a_hi, A_HI, b_all, B_HI, idesc128, ks > 0);
                    }
                    tc05::mma_f16_ss_w(dst + 64u, a_lo, A_HI, b_wh, B_HI, idesc64, 1);
                }
            }
            if (leader) {
                if (commits >= 1) tc05::commit(bars + 1 + (r & 1));
                if (commits >= 2) tc05::commit(bars + 3 + (r % 3));
            }
            __syncwarp();
        }
        if (leader) tc05::commit(bars);
        __syncwarp();
        mbar_wait(bars, 0);
        t1 = clock64();
    }
    tc05::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc<512>(tmem);
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    long long* d; cudaMalloc(&d, 148 * 8);
    const int smem = 2048 + 12 * BSTEP + 6 * SLOT + 1024;
    cudaFuncSetAttribute(row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int rows = 194;
    for (int commits = 0; commits < 3; ++commits)
        for (int rotate = 0; rotate < 2; ++rotate)
            for (int stages = 1; stages < 3; ++stages) {
                row_kernel<<<148, 128, smem>>>(commits, rotate, stages, rows, d);
                cudaError_t e = cudaDeviceSynchronize();
                std::vector<long long> h(148);
                cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost);
                long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
                printf("{\"probe\": \"umma_row\", \"commits_per_row\": %d, \"rotate_slots\": %d, \"stages\": %d, \"cycles_per_row\": %.0f, \"cuda\": \"%s\"}\n",
                       commits, rotate, stages, (double)mx / rows, cudaGetErrorString(e));
            }
    return 0;
}
