// stem.cu -- the fp32 stem of bnn.models.resnet in one kernel:
//     conv 7x7 / stride 2 / pad 3 (3 -> 64)  ->  eval BatchNorm  ->  ReLU  ->  MaxPool 3x3 / 2 / pad 1
// (reference bnn/models/resnet.py:85-92,147-153; plain nn.Conv2d there, cuDNN + three elementwise
// passes on a GPU).  Output: the residual stream in NHWC fp32 plus the sign/mask planes the first
// binarized conv needs, so the 112x112x64 conv output never goes to HBM.
//
// CTA = 4x8 pooled pixels x 64 channels, two CTAs per SM.  The repacked [tap][64] weights arrive by one TMA
// bulk copy; the 23x39x3 input window is staged with coalesced loads (zero fill = the conv's padding).  Lanes <->
// output channels: a weight is a conflict-free per-lane LDS, an input row segment is broadcast and kept
// in registers for the 7 horizontal taps, 17 conv pixels x 1 channel of fp32 accumulators per thread
// (sequential fma chain in (c_in, kh, kw) order -- the oracle restates exactly this order).  The
// 9x17x64 conv tile goes through shared memory to the 3x3 max, and the pooled pixel's planes are two
// ballots.  FP32-FMA bound: 118 M multiply-adds per 224x224 image, ~20 % recomputed halo.
#include "common.cuh"

#include <cudaTypedefs.h>
#include <mutex>

namespace bnn {

constexpr int ST_PH = 4, ST_PW = 8;                       // pooled tile
constexpr int ST_CR = 2 * ST_PH + 1, ST_CC = 2 * ST_PW + 1;   // conv tile 9 x 17
constexpr int ST_IR = 2 * ST_CR + 5, ST_IC = 2 * ST_CC + 5;   // input tile 23 x 39
constexpr int ST_ICP = 40;                                // input row pitch (floats), 160 B
constexpr int ST_CO = 64, ST_CI = 3, ST_K = 7;
constexpr int ST_WARPS = 8;                                // = ST_CR - 1, see the conv phase
constexpr size_t ST_IN_BYTES = (size_t)ST_CI * ST_IR * ST_ICP * 4;            // 16800
constexpr size_t ST_W_BYTES = (size_t)ST_CI * ST_K * ST_K * ST_CO * 4;        // 37632
constexpr size_t ST_CONV_BYTES = (size_t)ST_CR * ST_CC * ST_CO * 4;          // 65280
constexpr size_t ST_SMEM = 128 + ((ST_IN_BYTES + 127) & ~(size_t)127) + ST_W_BYTES + ST_CONV_BYTES;

struct StemArgs {
    const float* x;           // [n,3,h,w] contiguous
    const float* wt;          // [3][7][7][32][2]: (w[l], w[l + 32]) pairs
    const float *bn_scale, *bn_shift, *nx_scale, *nx_shift;
    float* out;               // [n,hp,wp,64]
    uint4* obits;             // [n][1][hp][wp]
    int N, H, W, Hc, Wc, Hp, Wp, tiles_h, tiles_w, stage;
};

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));    // SASS: FFMA2 (sm_100)
    return r;
}

__global__ void __launch_bounds__(ST_WARPS * 32, 2)
stem_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ StemArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    float* in_s = reinterpret_cast<float*>(smem + 128);
    float* w_s = reinterpret_cast<float*>(smem + 128 + ((ST_IN_BYTES + 127) & ~(size_t)127));
    float* conv_s = w_s + ST_CI * ST_K * ST_K * ST_CO;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int t = blockIdx.x;
    const int tw = t % a.tiles_w; t /= a.tiles_w;
    const int th = t % a.tiles_h;
    const int n = t / a.tiles_h;
    const int ph0 = th * ST_PH, pw0 = tw * ST_PW;
    const int cr0 = 2 * ph0 - 1, cc0 = 2 * pw0 - 1;            // first conv row / col of the tile
    const int hi0 = 2 * cr0 - 3, wi0 = 2 * cc0 - 3;            // first input row / col

    // staging: bit 0 of `stage` = input window by TMA tensor load, bit 1 = weights by TMA bulk copy
    const bool tma_in = (a.stage & 1) != 0, tma_w = (a.stage & 2) != 0;
    if (tma_in || tma_w) {
        if (threadIdx.x == 0) {
            if (tma_in) prefetch_tensormap(&tmap);
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, (unsigned)((tma_in ? ST_IN_BYTES : 0) + (tma_w ? ST_W_BYTES : 0)));
            if (tma_in) tma_load_4d(in_s, &tmap, bar, wi0, hi0, 0, n);
            if (tma_w) bulk_load_1d(w_s, a.wt, (unsigned)ST_W_BYTES, bar);
        }
    }
    if (!tma_in) {
        for (int i = threadIdx.x; i < ST_CI * ST_IR * ST_ICP; i += blockDim.x) {
            const int c = i % ST_ICP, r = (i / ST_ICP) % ST_IR, ci = i / (ST_ICP * ST_IR);
            const int hi = hi0 + r, wi = wi0 + c;
            float v = 0.0f;
            if ((unsigned)hi < (unsigned)a.H && (unsigned)wi < (unsigned)a.W)
                v = __ldg(a.x + (((size_t)n * ST_CI + ci) * a.H + hi) * a.W + wi);
            in_s[i] = v;
        }
    }
    if (!tma_w) {
        const float4* src = reinterpret_cast<const float4*>(a.wt);
        float4* dst = reinterpret_cast<float4*>(w_s);
        for (int i = threadIdx.x; i < ST_CI * ST_K * ST_K * ST_CO / 4; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    if (tma_in || tma_w) mbar_wait(bar, 0);

    // ---------------- conv + BN + ReLU into the shared conv tile ----------------
    // one warp = one conv row of the tile, both 32-channel blocks: the broadcast input row feeds 2 x 17 fma chains
    // 8 warps: warp w owns conv row w of the tile; the ninth row is shared out, three pixels per warp
    // ({2w, 2w+1, 2w+2}: neighbours overlap by one pixel and write the same value) so that every warp -- and
    // therefore every scheduler of the SM -- carries the same 20 fma chains.
    {
        const int r = warp;                      // ST_WARPS == ST_CR - 1
        const int e0 = 2 * warp;                 // first pixel of this warp's share of row ST_CR - 1
        // packed accumulators: (lo, hi) = (channel lane, channel lane + 32) of conv pixel c.  One FFMA2
        // (fma.rn.f32x2: input broadcast to both halves, weight pair from one LDS.64) does both channels --
        // the loop is issue-bound, so halving the FMA instruction count is what matters.  Each half is an
        // ordinary IEEE fma, so the result is bit-identical to the scalar chain the oracle restates.
        unsigned long long acc[ST_CC], acx[3];
#pragma unroll
        for (int c = 0; c < ST_CC; ++c) acc[c] = 0ull;
        acx[0] = acx[1] = acx[2] = 0ull;
        for (int ci = 0; ci < ST_CI; ++ci) {
#pragma unroll 1
            for (int kh = 0; kh < ST_K; ++kh) {
                const float4* irow = reinterpret_cast<const float4*>(in_s + (ci * ST_IR + 2 * r + kh) * ST_ICP);
                float iv[ST_ICP];
#pragma unroll
                for (int q = 0; q < ST_ICP / 4; ++q) {
                    const float4 v = irow[q];              // warp-uniform address: broadcast
                    iv[4 * q] = v.x; iv[4 * q + 1] = v.y; iv[4 * q + 2] = v.z; iv[4 * q + 3] = v.w;
                }
                // last conv row, pixels e0..e0+2: input columns 2*e0 .. 2*e0+10 (16-byte aligned start)
                const float4* xrow = reinterpret_cast<const float4*>(in_s + (ci * ST_IR + 2 * (ST_CR - 1) + kh) * ST_ICP + 2 * e0);
                float xv[12];
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const float4 v = xrow[q];
                    xv[4 * q] = v.x; xv[4 * q + 1] = v.y; xv[4 * q + 2] = v.z; xv[4 * q + 3] = v.w;
                }
                const float2* wrow = reinterpret_cast<const float2*>(w_s) + ((ci * ST_K + kh) * ST_K) * 32 + lane;
#pragma unroll
                for (int kw = 0; kw < ST_K; ++kw) {
                    const float2 w2 = wrow[kw * 32];
                    const unsigned long long ww = pack2(w2.x, w2.y);
#pragma unroll
                    for (int c = 0; c < ST_CC; ++c) acc[c] = fma2(pack2(iv[2 * c + kw], iv[2 * c + kw]), ww, acc[c]);
#pragma unroll
                    for (int c = 0; c < 3; ++c) acx[c] = fma2(pack2(xv[2 * c + kw], xv[2 * c + kw]), ww, acx[c]);
                }
            }
        }
        const float g0 = __ldg(a.bn_scale + lane), h0 = __ldg(a.bn_shift + lane);
        const float g1 = __ldg(a.bn_scale + 32 + lane), h1 = __ldg(a.bn_shift + 32 + lane);
        const bool row_ok = (unsigned)(cr0 + r) < (unsigned)a.Hc;
#pragma unroll
        for (int c = 0; c < ST_CC; ++c) {
            const bool ok = row_ok && (unsigned)(cc0 + c) < (unsigned)a.Wc;
            float lo, hi;
            unpack2(acc[c], lo, hi);
            // positions outside the conv output are max-pool padding: 0 is neutral after the ReLU
            conv_s[(r * ST_CC + c) * ST_CO + lane] = ok ? fmaxf(__fmaf_rn(lo, g0, h0), 0.0f) : 0.0f;
            conv_s[(r * ST_CC + c) * ST_CO + 32 + lane] = ok ? fmaxf(__fmaf_rn(hi, g1, h1), 0.0f) : 0.0f;
        }
        const bool last_ok = (unsigned)(cr0 + ST_CR - 1) < (unsigned)a.Hc;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int cc = e0 + c;
            const bool ok = last_ok && (unsigned)(cc0 + cc) < (unsigned)a.Wc;
            float lo, hi;
            unpack2(acx[c], lo, hi);
            conv_s[((ST_CR - 1) * ST_CC + cc) * ST_CO + lane] = ok ? fmaxf(__fmaf_rn(lo, g0, h0), 0.0f) : 0.0f;
            conv_s[((ST_CR - 1) * ST_CC + cc) * ST_CO + 32 + lane] = ok ? fmaxf(__fmaf_rn(hi, g1, h1), 0.0f) : 0.0f;
        }
    }
    __syncthreads();

    // ---------------- 3x3 / stride 2 max, NHWC store, planes for the first binarized conv ----------------
    for (int task = warp; task < ST_PH * ST_PW; task += ST_WARPS) {
        const int pr = task / ST_PW, pc = task - pr * ST_PW;
        const int ph = ph0 + pr, pw = pw0 + pc;
        if (ph >= a.Hp || pw >= a.Wp) continue;              // warp-uniform
        uint32_t sw[2], mw[2];
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
            const int ch = cb * 32 + lane;
            float m = 0.0f;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) m = fmaxf(m, conv_s[((2 * pr + i) * ST_CC + 2 * pc + j) * ST_CO + ch]);
            a.out[(((size_t)n * a.Hp + ph) * a.Wp + pw) * ST_CO + ch] = m;
            const float b = a.nx_scale ? __fmaf_rn(__ldg(a.nx_scale + ch), m, __ldg(a.nx_shift + ch)) : m;
            sw[cb] = __ballot_sync(0xffffffffu, b > 0.0f);
            mw[cb] = __ballot_sync(0xffffffffu, b > 0.0f || b < 0.0f);
        }
        if (lane == 0 && a.obits) a.obits[((size_t)n * a.Hp + ph) * a.Wp + pw] = make_uint4(sw[0], sw[1], mw[0], mw[1]);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn();      // bconv.cu

}  // namespace bnn

using namespace bnn;

extern "C" int bnn_stem_out_hw(int32_t h, int32_t w, int32_t* hp, int32_t* wp) {
    if (!hp || !wp) return BNN_E_NULL;
    if (h <= 0 || w <= 0) return BNN_E_SHAPE;
    const int hc = (h + 6 - 7) / 2 + 1, wc = (w + 6 - 7) / 2 + 1;
    *hp = (hc + 2 - 3) / 2 + 1;
    *wp = (wc + 2 - 3) / 2 + 1;
    return 0;
}

extern "C" int bnn_stem_fwd(const float* x, int32_t n, int32_t h, int32_t w, const float* w_t,
                            const float* bn_scale, const float* bn_shift, const float* nx_scale,
                            const float* nx_shift, float* out, void* out_bits, uint32_t flags, void* stream_) {
    if (!x || !w_t || !bn_scale || !bn_shift || !out) return BNN_E_NULL;
    if ((nx_scale == nullptr) != (nx_shift == nullptr)) return BNN_E_NULL;
    if (n <= 0 || h < 7 || w < 7) return BNN_E_SHAPE;
    if (((uintptr_t)w_t & 15) || ((uintptr_t)out_bits & 15)) return BNN_E_ALIGN;
    StemArgs a{};
    a.x = x; a.wt = w_t; a.bn_scale = bn_scale; a.bn_shift = bn_shift; a.nx_scale = nx_scale; a.nx_shift = nx_shift;
    a.out = out; a.obits = (uint4*)out_bits;
    a.N = n; a.H = h; a.W = w;
    a.Hc = (h + 6 - 7) / 2 + 1; a.Wc = (w + 6 - 7) / 2 + 1;
    a.Hp = (a.Hc + 2 - 3) / 2 + 1; a.Wp = (a.Wc + 2 - 3) / 2 + 1;
    a.tiles_h = (a.Hp + ST_PH - 1) / ST_PH; a.tiles_w = (a.Wp + ST_PW - 1) / ST_PW;
    // staging mode: weights by TMA bulk copy; the input window by TMA only on request (BNN_F_STEM_TMA_IN) and
    // only when rows are 16-byte aligned -- plain coalesced loads otherwise
    a.stage = (flags & BNN_F_STAGE_LDG) ? 0 : 2;
    if ((flags & BNN_F_STEM_TMA_IN) && (w % 4) == 0 && ((uintptr_t)x & 15) == 0) a.stage |= 1;
    if (flags & BNN_F_STEM_NO_BULK) a.stage &= ~2;

    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if (a.stage & 1) {
        EncodeTiledFn enc = encode_tiled_fn();
        if (!enc) return BNN_E_DRIVER;
        const cuuint64_t gdim[4] = {(cuuint64_t)w, (cuuint64_t)h, 3, (cuuint64_t)n};
        const cuuint64_t gstr[3] = {(cuuint64_t)w * 4, (cuuint64_t)w * h * 4, (cuuint64_t)w * h * 3 * 4};
        const cuuint32_t box[4] = {ST_ICP, ST_IR, ST_CI, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return BNN_E_DRIVER;
    }
    cudaError_t ce = cudaFuncSetAttribute((const void*)stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM);
    if (ce != cudaSuccess) return (int)ce;
    const long long ctas = (long long)n * a.tiles_h * a.tiles_w;
    if (ctas > 0x7fffffffLL) return BNN_E_UNSUPPORTED;
    stem_kernel<<<(unsigned)ctas, ST_WARPS * 32, ST_SMEM, (cudaStream_t)stream_>>>(tmap, a);
    count_launch(1);
    return (int)cudaGetLastError();
}
