// shortcut.cu -- the down-sampling shortcut of bnn.models.resnet in one kernel:
//     AvgPool2d(k, stride k, ceil_mode, count_include_pad=False)  ->  sign()  ->  binarized conv 1x1  ->  eval BatchNorm
// (reference bnn/models/resnet.py:129-133; the conv is a bnn.layers.Conv2d, bnn/layers/conv.py:90-97).
// The two-launch form (bnn_avgpool_pack_f32 + bnn_bconv2d_fused_fwd) writes the pooled planes to HBM and reads them
// back; here they live in shared memory for the few microseconds between the two phases of a CTA, so the kernel
// reads the fp32 residual stream once and writes the fp32 shortcut once -- HBM-bound by construction:
//     bytes = 4 * C_in * H * W  +  4 * C_out * H_o * W_o   per image.
// Phase 1 (warp = pooled pixel, lanes <-> input channels): k x k average in the pack kernel's exact operation order,
// planes by ballot, all loads of two pixels in flight per warp.  Phase 2 (lanes <-> output channels): XNOR/AND +
// POPC against the packed weights, folded epilogue z = fma(k0, dot, k1) with k0/k1 exactly as bconv_kernel's EPI 1
// folds alpha, bias, post-scale and BatchNorm; NHWC store, 128 B per warp.  In the compiled-chunk-count instances a
// warp takes one 32-channel block at a time: the block's weight words for ALL chunks sit in registers (one L2 round
// trip per block), the CTA's 32 pixels stream past them as broadcast LDS.128, and chunk triples go through the same
// 3:2 carry-save adder as bconv_kernel (3 words, 2 POPC).  Layers with few pixels and many output channels (ResNet-50
// layer3/4: 25 k / 6 k pixels, 1024 / 2048 channels) split the channel blocks over gridDim.y so every SM has work.
// Bit-identical to the two-launch form (tests/test_gpu_fused.py).
#include "common.cuh"

namespace bnn {

constexpr int SC_WARPS = 8, SC_PIX = 32, SC_PPW = SC_PIX / SC_WARPS;     // pooled pixels per CTA / per warp (phase 1)
constexpr int SC_GRP = 8, SC_NGRP = SC_PIX / SC_GRP;                     // phase 2: pixel groups of 8
constexpr int SC_WREG = 2;                                               // weight chunks held in registers per pass
static_assert(SC_PPW % 2 == 0, "phase 1 packs pixel pairs");

struct ShortcutArgs {
    const float* x;                // channels-last: element (n, c, h, w) at n*sn + h*sh + w*sw + c
    long long sn, sh, sw;
    const uint2* wbits;            // [c_out/32][chunk][32] x {lo, hi}
    const float *scale, *bias, *post, *bn_scale, *bn_shift;
    float* out;                    // [n, ho, wo, c_out] contiguous
    int N, C, H, W, Ho, Wo, pool, Cout, nch, nblk32;
    int pixels;                    // n * ho * wo
    int parts;                     // phase 2 (compiled chunk counts): pixel parts per channel block, 1 / 2 / 4 / 8
    int ysplit;                    // gridDim.y: the 32-channel output blocks are divided among the CTAs of a pixel tile
};

// k x k average of one channel at one pooled pixel, the pack kernel's operation order (pack.cu)
__device__ __forceinline__ float pooled_value(const float* base, int sh, int sw, int k, int hmax, int wmax) {
    if (k == 1) return __ldg(base);
    if (k == 2 && hmax == 2 && wmax == 2) {
        const float t00 = __ldg(base), t01 = __ldg(base + sw), t10 = __ldg(base + sh), t11 = __ldg(base + sh + sw);
        return __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(t00, t01), t10), t11), 4.0f);
    }
    float sum = 0.0f;
    for (int i = 0; i < hmax; ++i)
        for (int j = 0; j < wmax; ++j) sum = __fadd_rn(sum, __ldg(base + i * sh + j * sw));
    return __fdiv_rn(sum, (float)(hmax * wmax));
}

// NCH > 0: chunks known at compile time.  POOL = 1 / 2: no pooling / 2x2 windows that are always full, over a multiple
// of 64 channels (every ResNet shape): straight-line code, 8 / 32 loads in flight per lane.  POOL == 0: any geometry.
template <int NCH, int POOL>
__global__ void __launch_bounds__(SC_WARPS * 32, 3)
shortcut_kernel(const __grid_constant__ ShortcutArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int nch = NCH > 0 ? NCH : a.nch;
    uint4* bits = reinterpret_cast<uint4*>(smem);                            // [SC_PIX][nch]
    int* msum = reinterpret_cast<int*>(smem + (size_t)SC_PIX * nch * 16);    // [SC_PIX]
    float* k0s = reinterpret_cast<float*>(msum + SC_PIX);                    // [nblk32 * 32]
    float* k1s = k0s + a.nblk32 * 32;

    pdl_launch_dependents();      // programmatic dependent launch (common.cuh): no global access before pdl_wait()
    pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pix0 = blockIdx.x * SC_PIX;
    constexpr bool FAST = POOL > 0;
    const int k = FAST ? POOL : (a.pool > 1 ? a.pool : 1);
    // in-image offsets fit 32 bits (host-checked); only the image base is a 64-bit product
    const int sh = (int)a.sh, sw = (int)a.sw;

    // ---------------- phase 1: pool + sign -> planes in shared memory ----------------
    // a warp packs SC_PPW consecutive pooled pixels; (n, h, w) by one division pair, then incrementally
    {
        const int first = pix0 + warp * SC_PPW;
        int w = first % a.Wo, r = first / a.Wo;
        int h = r % a.Ho, n = r / a.Ho;
        const float* base[SC_PPW];
        int hmax[SC_PPW], wmax[SC_PPW];
        bool ok[SC_PPW];
#pragma unroll
        for (int u = 0; u < SC_PPW; ++u) {
            ok[u] = first + u < a.pixels;                      // warp-uniform
            base[u] = a.x + (ok[u] ? n : 0) * a.sn + (h * k * sh + w * k * sw) + lane;
            hmax[u] = min(k, a.H - h * k); wmax[u] = min(k, a.W - w * k);
            if (++w == a.Wo) { w = 0; if (++h == a.Ho) { h = 0; ++n; } }
        }
        int cnt[SC_PPW];
#pragma unroll
        for (int u = 0; u < SC_PPW; ++u) cnt[u] = 0;
#pragma unroll 1
        for (int ch = 0; ch < nch; ++ch) {                     // one chunk = 32 loads per lane in flight
            float v[SC_PPW][2];
            if constexpr (POOL == 2) {
                float t[SC_PPW][2][4];
#pragma unroll
                for (int u = 0; u < SC_PPW; ++u)
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        const float* q = base[u] + ch * 64 + b * 32;
                        const bool on = ok[u];
                        t[u][b][0] = on ? __ldg(q) : 0.0f; t[u][b][1] = on ? __ldg(q + sw) : 0.0f;
                        t[u][b][2] = on ? __ldg(q + sh) : 0.0f; t[u][b][3] = on ? __ldg(q + sh + sw) : 0.0f;
                    }
#pragma unroll
                for (int u = 0; u < SC_PPW; ++u)
#pragma unroll
                    for (int b = 0; b < 2; ++b)
                        v[u][b] = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(t[u][b][0], t[u][b][1]), t[u][b][2]), t[u][b][3]), 4.0f);
            } else if constexpr (POOL == 1) {
#pragma unroll
                for (int u = 0; u < SC_PPW; ++u)
#pragma unroll
                    for (int b = 0; b < 2; ++b) v[u][b] = ok[u] ? __ldg(base[u] + ch * 64 + b * 32) : 0.0f;
            } else {
#pragma unroll
                for (int u = 0; u < SC_PPW; ++u)
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        const int c = ch * 64 + b * 32 + lane;
                        v[u][b] = (ok[u] && c < a.C) ? pooled_value(base[u] + ch * 64 + b * 32, sh, sw, k, hmax[u], wmax[u]) : 0.0f;
                    }
            }
#pragma unroll
            for (int u = 0; u < SC_PPW; ++u) {
                uint32_t s_[2], m_[2];
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    s_[b] = __ballot_sync(0xffffffffu, v[u][b] > 0.0f);
                    m_[b] = __ballot_sync(0xffffffffu, v[u][b] > 0.0f || v[u][b] < 0.0f);
                }
                cnt[u] += __popc(m_[0]) + __popc(m_[1]);
                if (lane == 0) bits[(warp * SC_PPW + u) * nch + ch] = make_uint4(s_[0], s_[1], m_[0], m_[1]);
            }
        }
        if (lane < SC_PPW) {
            int mine = 0;
#pragma unroll
            for (int u = 0; u < SC_PPW; ++u) mine = (lane == u) ? cnt[u] : mine;
            msum[warp * SC_PPW + lane] = mine;
        }
    }

    // folded per-channel constants (same operations as bconv_kernel EPI 1); after phase 1 so that these loads do not
    // sit in front of the input loads
    for (int i = threadIdx.x; i < a.nblk32 * 32; i += blockDim.x) {
        const bool ok = i < a.Cout;
        float k0 = (ok && a.scale) ? __ldg(a.scale + i) : 1.0f;
        float k1 = (ok && a.bias) ? __ldg(a.bias + i) : 0.0f;
        const float post = (ok && a.post) ? __ldg(a.post + i) : 1.0f;
        k0 = __fmul_rn(k0, post); k1 = __fmul_rn(k1, post);
        if (a.bn_scale) {
            const float g = ok ? __ldg(a.bn_scale + i) : 1.0f, h = ok ? __ldg(a.bn_shift + i) : 0.0f;
            k0 = __fmul_rn(k0, g);
            k1 = __fadd_rn(__fmul_rn(k1, g), h);
        }
        k0s[i] = k0; k1s[i] = k1;
    }
    __syncthreads();

    // ---------------- phase 2: 1x1 binary conv + folded epilogue ----------------
    const int blk_lo = (int)(((long long)a.nblk32 * blockIdx.y) / gridDim.y);
    const int blk_hi = (int)(((long long)a.nblk32 * (blockIdx.y + 1)) / gridDim.y);
    if constexpr (NCH > 0) {
        // task = (32-channel block, pixel part): the block's weights of all chunks in registers, the part's pixels streamed
        // past them.  parts > 1 only when there are fewer blocks than warps (host-chosen so that no warp idles).
        const int nb = blk_hi - blk_lo, ppp = SC_PIX / a.parts;
        for (int task = warp; task < nb * a.parts; task += SC_WARPS) {
            const int part = task / nb, blk = blk_lo + (task - part * nb);
            const int c = blk * 32 + lane;
            const uint2* wrow = a.wbits + (size_t)blk * NCH * 32 + lane;
            uint2 t[NCH];
#pragma unroll
            for (int u = 0; u < NCH; ++u) t[u] = __ldg(wrow + u * 32);
            const float k0 = k0s[c], k1 = k1s[c];          // c < nblk32 * 32: the tables are padded
            const bool c_ok = c < a.Cout;
            const int p_lo = part * ppp;
            const int p_hi = min(p_lo + ppp, a.pixels - pix0);           // warp-uniform
            float* op = a.out + (size_t)pix0 * a.Cout + c;
            constexpr int NTRI = NCH / 3;
            // one pixel at a time: the dot of a (pixel, channel) is a scalar per lane, so nothing but the weights stays live
#pragma unroll 2
            for (int p = p_lo; p < p_hi; ++p) {
                const uint4* vp = bits + p * NCH;
                int dis = 0;
#pragma unroll
                for (int q = 0; q < NTRI; ++q) {
                    const uint4 v0 = vp[3 * q], v1 = vp[3 * q + 1], v2 = vp[3 * q + 2];
                    const uint32_t x0 = v0.z & (v0.x ^ t[3 * q].x), x1 = v1.z & (v1.x ^ t[3 * q + 1].x),
                                   x2 = v2.z & (v2.x ^ t[3 * q + 2].x);
                    const uint32_t y0 = v0.w & (v0.y ^ t[3 * q].y), y1 = v1.w & (v1.y ^ t[3 * q + 1].y),
                                   y2 = v2.w & (v2.y ^ t[3 * q + 2].y);
                    const int ones = __popc(x0 ^ x1 ^ x2) + __popc(y0 ^ y1 ^ y2);
                    const int twos = __popc((x0 & x1) | (x2 & (x0 ^ x1))) + __popc((y0 & y1) | (y2 & (y0 ^ y1)));
                    dis += ones + 2 * twos;
                }
#pragma unroll
                for (int u = 3 * NTRI; u < NCH; ++u) {
                    const uint4 v = vp[u];
                    dis += __popc(v.z & (v.x ^ t[u].x)) + __popc(v.w & (v.y ^ t[u].y));
                }
                if (c_ok) op[(size_t)p * a.Cout] = __fmaf_rn(k0, (float)(msum[p] - 2 * dis), k1);
            }
        }
        return;
    }
    // any chunk count: task = (32-channel output block, group of 8 pixels), round-robin over the warps; the block's
    // weight words sit in registers (SC_WREG chunks per pass), the planes are broadcast LDS.128.
    const int ntask = (blk_hi - blk_lo) * SC_NGRP;
    for (int task = warp; task < ntask; task += SC_WARPS) {
        const int blk = blk_lo + task / SC_NGRP, i0 = (task % SC_NGRP) * SC_GRP;
        const int npix = min(SC_GRP, a.pixels - (pix0 + i0));  // <= 0 beyond the last pixel: warp-uniform
        if (npix <= 0) continue;
        const int c = blk * 32 + lane;
        const uint2* wrow = a.wbits + (size_t)blk * nch * 32 + lane;
        int dis[SC_GRP];
#pragma unroll
        for (int p = 0; p < SC_GRP; ++p) dis[p] = 0;
        for (int ch0 = 0; ch0 < nch; ch0 += SC_WREG) {
            uint2 t[SC_WREG];
#pragma unroll
            for (int u = 0; u < SC_WREG; ++u) t[u] = (ch0 + u < nch) ? __ldg(wrow + (ch0 + u) * 32) : make_uint2(0u, 0u);
#pragma unroll
            for (int u = 0; u < SC_WREG; ++u) {
                if (ch0 + u < nch) {
#pragma unroll
                    for (int p = 0; p < SC_GRP; ++p) {
                        const uint4 v = bits[(i0 + p) * nch + ch0 + u];     // rows beyond npix: stale, never stored
                        dis[p] += __popc(v.z & (v.x ^ t[u].x)) + __popc(v.w & (v.y ^ t[u].y));
                    }
                }
            }
        }
        if (c < a.Cout) {
            const float k0 = k0s[c], k1 = k1s[c];
            float* op = a.out + (size_t)(pix0 + i0) * a.Cout + c;
#pragma unroll
            for (int p = 0; p < SC_GRP; ++p)
                if (p < npix) op[p * a.Cout] = __fmaf_rn(k0, (float)(msum[i0 + p] - 2 * dis[p]), k1);
        }
    }
}

template <int NCH, int POOL>
static cudaError_t launch_shortcut(const ShortcutArgs& a, size_t smem, unsigned ctas, cudaStream_t stream) {
    cudaError_t ce = cudaFuncSetAttribute((const void*)shortcut_kernel<NCH, POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce != cudaSuccess) return ce;
    ce = launch_pdl(shortcut_kernel<NCH, POOL>, dim3(ctas, (unsigned)a.ysplit, 1), dim3(SC_WARPS * 32), smem, stream, a);
    return ce != cudaSuccess ? ce : cudaGetLastError();
}

}  // namespace bnn

using namespace bnn;

extern "C" int bnn_shortcut_fwd(const float* x, int64_t xs_n, int64_t xs_h, int64_t xs_w, int32_t n, int32_t c_in,
                                int32_t h, int32_t w, int32_t pool, int32_t ceil_mode, const void* wbits,
                                int32_t c_out, const float* scale, const float* bias, const float* post,
                                const float* bn_scale, const float* bn_shift, float* out, uint32_t flags,
                                void* stream_) {
    if (!x || !wbits || !out) return BNN_E_NULL;
    if ((bn_scale == nullptr) != (bn_shift == nullptr)) return BNN_E_NULL;
    if (n <= 0 || c_in <= 0 || h <= 0 || w <= 0 || c_out <= 0 || pool < 1) return BNN_E_SHAPE;
    const int k = pool;
    const int ho = ceil_mode ? (h + k - 1) / k : h / k, wo = ceil_mode ? (w + k - 1) / k : w / k;
    if (ho <= 0 || wo <= 0) return BNN_E_SHAPE;
    const long long pixels = (long long)n * ho * wo;
    if (pixels > 0x7fffffffLL - SC_PIX) return BNN_E_UNSUPPORTED;
    ShortcutArgs a{};
    a.x = x; a.sn = xs_n; a.sh = xs_h; a.sw = xs_w;
    a.wbits = (const uint2*)wbits;
    a.scale = scale; a.bias = bias; a.post = post; a.bn_scale = bn_scale; a.bn_shift = bn_shift;
    a.out = out;
    a.N = n; a.C = c_in; a.H = h; a.W = w; a.Ho = ho; a.Wo = wo; a.pool = pool; a.Cout = c_out;
    a.nch = (c_in + 63) / 64; a.nblk32 = (c_out + 31) / 32;
    a.pixels = (int)pixels;
    const size_t smem = (size_t)SC_PIX * a.nch * 16 + SC_PIX * 4 + (size_t)a.nblk32 * 32 * 8;
    if (smem > 200 * 1024) return BNN_E_UNSUPPORTED;
    // 32-bit offsets inside one image and inside the output
    if ((long long)(h - 1) * xs_h + (long long)(w - 1) * xs_w + c_in >= 0x7fffffffLL || xs_h < 0 || xs_w < 0) return BNN_E_UNSUPPORTED;
    if ((long long)SC_PIX * c_out >= 0x7fffffffLL) return BNN_E_UNSUPPORTED;
    const unsigned ctas = (unsigned)((pixels + SC_PIX - 1) / SC_PIX);
    // few pixel tiles, many channel blocks: divide the blocks over gridDim.y (each CTA repeats the cheap pool + pack) so
    // that the SMs get equal numbers of CTAs; every CTA keeps at least one block per warp
    a.ysplit = 1;
    {
        int sms = 148, dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (ctas < 4u * (unsigned)sms) {
            double best = 1e30;
            for (int ys = 1; ys <= 4 && a.nblk32 / ys >= SC_WARPS; ++ys) {
                const double per_sm = (double)ctas * ys / sms;
                const double cost = (double)((long long)(per_sm + 0.999999)) / per_sm * (1.0 + 0.04 * (ys - 1));
                if (cost < best) { best = cost; a.ysplit = ys; }
            }
        }
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool fast = !(flags & BNN_F_STAGE_LDG) && c_in % 64 == 0 &&
                      ((pool == 2 && h % 2 == 0 && w % 2 == 0) || pool == 1);
    cudaError_t ce;
    {   // pixel parts per block so that (blocks of a CTA) x parts fills the 8 warps evenly
        const int nb = a.nblk32 / a.ysplit;               // smallest share of a CTA
        double best = 1e30;
        a.parts = 1;
        for (int parts = 1; parts <= SC_WARPS; parts *= 2) {
            const double rounds = (double)nb * parts / SC_WARPS;
            const double cost = (double)((long long)(rounds + 0.999999)) / rounds * (1.0 + 0.02 * (parts - 1));
            if (cost < best) { best = cost; a.parts = parts; }
        }
    }
#define BNN_SC(nch_, pool_) if (fast && a.nch == nch_ && pool == pool_) ce = launch_shortcut<nch_, pool_>(a, smem, ctas, stream); else
    BNN_SC(1, 2) BNN_SC(2, 2) BNN_SC(4, 2) BNN_SC(8, 2) BNN_SC(16, 2)
    BNN_SC(1, 1) BNN_SC(2, 1) BNN_SC(4, 1) BNN_SC(8, 1) BNN_SC(16, 1)
    ce = launch_shortcut<0, 0>(a, smem, ctas, stream);
#undef BNN_SC
    count_launch(1);
    return (int)ce;
}
