// pack.cu -- the two binarizers of the reference as bit-pack kernels.
//
//   bnn_pack_act_f32     BasicInputBinarizer / SignActivation.forward
//                        (reference bnn/ops.py:151-152, 63-66)
//   bnn_pack_weight_f32  XNORWeightBinarizer.forward (reference bnn/ops.py:116-140)
//
// Both are HBM-streaming kernels: the activation pack reads 4 B/element once,
// coalesced along w, and writes 2 bits/element; the weight pack runs once at
// prepare time.
#include "common.cuh"

namespace bnn {

// ---------------------------------------------------------------------------
// activations: fp32 (element strides) -> abits[n][chunk][h][w]
// one thread = one (n, chunk, h, w) unit = 64 channels of one pixel; adjacent
// lanes are adjacent w, so every one of the 64 loads of a warp is one 128-B line
// for NCHW input.  POOL > 0 averages a POOL x POOL window first (AvgPool2d with
// kernel = stride, count_include_pad = False): sum in row-major order, divide by
// the number of in-bounds elements, exactly like torch's CPU kernel.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float load_pooled(const float* p, long long sh, long long sw, int k, int hmax, int wmax) {
    if (k <= 1) return __ldg(p);
    float sum = 0.0f;
    int cnt = 0;
    for (int i = 0; i < k && i < hmax; ++i)
        for (int j = 0; j < k && j < wmax; ++j) { sum = __fadd_rn(sum, __ldg(p + i * sh + j * sw)); ++cnt; }
    return __fdiv_rn(sum, (float)cnt);
}

__global__ void __launch_bounds__(256)
pack_act_kernel(const float* __restrict__ x, long long sn, long long sc, long long sh, long long sw,
                int N, int C, int H, int W, int nch, int pool, int Hin, int Win,
                const float* __restrict__ pre_scale, const float* __restrict__ pre_shift, int pre_relu,
                uint4* __restrict__ abits) {
    pdl_launch_dependents();      // programmatic dependent launch (common.cuh): no global access before pdl_wait()
    pdl_wait();
    const long long total = (long long)N * nch * H * W;       // H, W: OUTPUT (pooled) plane size
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int w = (int)(idx % W);
    long long r = idx / W;
    const int h = (int)(r % H);
    r /= H;
    const int ch = (int)(r % nch);
    const int n = (int)(r / nch);

    const int k = pool > 1 ? pool : 1;
    const float* base = x + n * sn + (long long)h * k * sh + (long long)w * k * sw + (long long)ch * 64 * sc;
    const int hmax = Hin - h * k, wmax = Win - w * k;
    const int cmax = min(64, C - ch * 64);
    uint32_t s[2] = {0u, 0u}, m[2] = {0u, 0u};
    if (cmax == 64 && pool <= 1 && pre_scale == nullptr && !pre_relu) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int b0 = 0; b0 < 32; b0 += 16) {
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __ldg(base + (long long)(half * 32 + b0 + i) * sc);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const uint32_t pos = v[i] > 0.0f, neg = v[i] < 0.0f;  // NaN, +-0 -> neither
                    s[half] |= pos << (b0 + i);
                    m[half] |= (pos | neg) << (b0 + i);
                }
            }
        }
    } else {
#pragma unroll 4
        for (int b = 0; b < cmax; ++b) {
            float v = load_pooled(base + (long long)b * sc, sh, sw, k, hmax, wmax);
            if (pre_scale != nullptr) {
                const int c = ch * 64 + b;
                v = __fadd_rn(__fmul_rn(v, __ldg(pre_scale + c)), __ldg(pre_shift + c));
            }
            const uint32_t pos = v > 0.0f, neg = (v < 0.0f) && !pre_relu;     // relu(v) < 0 never happens
            s[b >> 5] |= pos << (b & 31);
            m[b >> 5] |= (pos | neg) << (b & 31);
        }
    }
    abits[idx] = make_uint4(s[0], s[1], m[0], m[1]);
}

// ---------------------------------------------------------------------------
// channel-contiguous input (stride_c == 1: torch channels_last, or Linear's [rows, features]):
// one warp per output pixel, lanes <-> channels, so every load is one 128-byte line and the
// 32 sign / mask bits of a block are one ballot each.  Same pooling / affine semantics as above.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_act_cl_kernel(const float* __restrict__ x, long long sn, long long sh, long long sw,
                   int N, int C, int H, int W, int nch, int pool, int Hin, int Win,
                   const float* __restrict__ pre_scale, const float* __restrict__ pre_shift, int pre_relu,
                   uint32_t* __restrict__ abits) {
    pdl_launch_dependents();      // programmatic dependent launch (common.cuh): no global access before pdl_wait()
    pdl_wait();
    const long long pixels = (long long)N * H * W;
    const long long pix = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pix >= pixels) return;                               // warp-uniform
    const int lane = threadIdx.x & 31;
    const int w = (int)(pix % W);
    const long long r = pix / W;
    const int h = (int)(r % H), n = (int)(r / H);
    const int k = pool > 1 ? pool : 1;
    const float* base = x + n * sn + (long long)h * k * sh + (long long)w * k * sw;
    const int hmax = min(k, Hin - h * k), wmax = min(k, Win - w * k);
    const float cnt_f = (float)(hmax * wmax);
    // two 32-channel blocks (one 64-channel unit) per iteration: all loads of the unit are issued before the
    // first ballot, and lane 0 writes the unit with one 16-byte store
    for (int ch = 0; ch < nch; ++ch) {
        float v[2];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int c = ch * 64 + b * 32 + lane;
            float t = 0.0f;
            if (c < C) {
                if (k == 1) t = __ldg(base + c);
                else if (k == 2 && hmax == 2 && wmax == 2) {
                    const float t00 = __ldg(base + c), t01 = __ldg(base + sw + c);
                    const float t10 = __ldg(base + sh + c), t11 = __ldg(base + sh + sw + c);
                    t = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(t00, t01), t10), t11), 4.0f);
                } else {
                    float sum = 0.0f;
                    for (int i = 0; i < hmax; ++i)
                        for (int j = 0; j < wmax; ++j) sum = __fadd_rn(sum, __ldg(base + i * sh + j * sw + c));
                    t = __fdiv_rn(sum, cnt_f);
                }
                if (pre_scale != nullptr) t = __fadd_rn(__fmul_rn(t, __ldg(pre_scale + c)), __ldg(pre_shift + c));
            }
            v[b] = t;
        }
        uint32_t s_[2], m_[2];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const bool ok = ch * 64 + b * 32 + lane < C;
            s_[b] = __ballot_sync(0xffffffffu, ok && v[b] > 0.0f);
            m_[b] = __ballot_sync(0xffffffffu, ok && (v[b] > 0.0f || (v[b] < 0.0f && !pre_relu)));
        }
        if (lane == 0)
            reinterpret_cast<uint4*>(abits)[(((size_t)n * nch + ch) * H + h) * W + w] = make_uint4(s_[0], s_[1], m_[0], m_[1]);
    }
}

// Dense NHWC fast path (pixel p at x + p * C, C = 64 * NCH, no pooling): a warp packs PL_PPW consecutive pixels with
// every load of all of them in flight (4 * 2 * NCH independent 128-byte lines per warp), the optional pre-sign affine
// held in registers, pixel -> (image, position) by one division per warp.  HBM-streaming: 4 B/element in, 2 bits out.
constexpr int PL_PPW = 4;

template <int NCH>
__global__ void __launch_bounds__(256)
pack_act_cl_dense_kernel(const float* __restrict__ x, int pixels, int HW, const float* __restrict__ pre_scale,
                         const float* __restrict__ pre_shift, int pre_relu, uint4* __restrict__ abits) {
    pdl_launch_dependents();      // programmatic dependent launch (common.cuh): no global access before pdl_wait()
    pdl_wait();
    constexpr int C = 64 * NCH, NB = 2 * NCH;
    const int lane = threadIdx.x & 31;
    const int pix0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * PL_PPW;
    if (pix0 >= pixels) return;                              // warp-uniform
    float v[PL_PPW][NB];
#pragma unroll
    for (int u = 0; u < PL_PPW; ++u) {
        const float* px = x + (size_t)(pix0 + u) * C + lane;
        const bool on = pix0 + u < pixels;
#pragma unroll
        for (int b = 0; b < NB; ++b) v[u][b] = on ? __ldg(px + b * 32) : 0.0f;
    }
    const bool pre = pre_scale != nullptr;
    float sc[NB], sf[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        sc[b] = pre ? __ldg(pre_scale + b * 32 + lane) : 1.0f;
        sf[b] = pre ? __ldg(pre_shift + b * 32 + lane) : 0.0f;
    }
    int n = pix0 / HW, hw = pix0 - n * HW;
#pragma unroll
    for (int u = 0; u < PL_PPW; ++u) {
        if (pix0 + u >= pixels) break;                       // warp-uniform
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            uint32_t s_[2], m_[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float t = v[u][2 * ch + h];
                if (pre) t = __fadd_rn(__fmul_rn(t, sc[2 * ch + h]), sf[2 * ch + h]);      // same two roundings as the generic kernels
                s_[h] = __ballot_sync(0xffffffffu, t > 0.0f);
                m_[h] = __ballot_sync(0xffffffffu, t > 0.0f || (t < 0.0f && !pre_relu));
            }
            if (lane == 0) abits[((size_t)n * NCH + ch) * HW + hw] = make_uint4(s_[0], s_[1], m_[0], m_[1]);
        }
        if (++hw == HW) { hw = 0; ++n; }
    }
}

template <int NCH>
static int launch_pack_dense(const float* x, long long pixels, int HW, const float* pre_scale, const float* pre_shift,
                             int pre_relu, void* abits, cudaStream_t stream) {
    const long long warps = (pixels + PL_PPW - 1) / PL_PPW, blocks = (warps + 7) / 8;
    if (blocks > 0x7fffffffLL || pixels > 0x7fffffffLL - PL_PPW) return BNN_E_UNSUPPORTED;
    launch_pdl(pack_act_cl_dense_kernel<NCH>, dim3((unsigned)blocks), dim3(256), 0, stream, x, (int)pixels, HW, pre_scale, pre_shift, pre_relu, (uint4*)abits);
    count_launch(1);
    return (int)cudaGetLastError();
}

// AvgPool2d(2) (kernel = stride = 2, floor mode) of a dense NHWC tensor, fused with the bit-pack of the NEXT binarized
// layer: one pass reads the fp32 tensor, writes the pooled fp32 tensor (the residual stream the next block adds its
// convolutions to) and the planes of sign(pooled * pre_scale + pre_shift).  The Hierarchical-Block harness pools a
// 1 GB NHWC tensor between two blocks (hierarchical_block.py:38-60 stacked as in SURVEY.md A.1.4); as a torch AvgPool2d
// followed by a pack that was 1.6 + 0.06 ms, this is one HBM pass.  Window sum in row-major order, divided by 4: the pack
// kernels' (= torch's CPU kernel's) operation order.  A warp owns AP_PPW pooled pixels, lanes <-> channels.
constexpr int AP_PPW = 2;

template <int NCH>
__global__ void __launch_bounds__(256)
avgpool2_pack_cl_kernel(const float* __restrict__ x, long long pooled_pixels, int H, int W, int Ho, int Wo,
                        const float* __restrict__ pre_scale, const float* __restrict__ pre_shift, int pre_relu,
                        float* __restrict__ pooled, uint4* __restrict__ abits) {
    pdl_launch_dependents();      // programmatic dependent launch (common.cuh): no global access before pdl_wait()
    pdl_wait();
    constexpr int C = 64 * NCH, NB = 2 * NCH;
    const int lane = threadIdx.x & 31;
    const long long pp0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * AP_PPW;
    if (pp0 >= pooled_pixels) return;                        // warp-uniform
    float v[AP_PPW][4][NB];
    int n_[AP_PPW], ho_[AP_PPW], wo_[AP_PPW];
#pragma unroll
    for (int u = 0; u < AP_PPW; ++u) {
        const long long pp = pp0 + u < pooled_pixels ? pp0 + u : pooled_pixels - 1;
        wo_[u] = (int)(pp % Wo);
        const long long t = pp / Wo;
        ho_[u] = (int)(t % Ho);
        n_[u] = (int)(t / Ho);
        const float* px = x + (((size_t)n_[u] * H + 2 * ho_[u]) * W + 2 * wo_[u]) * C + lane;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float* pq = px + ((size_t)(q >> 1) * W + (q & 1)) * C;
#pragma unroll
            for (int b = 0; b < NB; ++b) v[u][q][b] = __ldg(pq + b * 32);
        }
    }
    const bool pre = pre_scale != nullptr;
#pragma unroll
    for (int u = 0; u < AP_PPW; ++u) {
        if (pp0 + u >= pooled_pixels) break;                 // warp-uniform
        const size_t opix = ((size_t)n_[u] * Ho + ho_[u]) * Wo + wo_[u];
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            uint32_t s_[2], m_[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int b = 2 * ch + h;
                float t = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(v[u][0][b], v[u][1][b]), v[u][2][b]), v[u][3][b]), 4.0f);
                if (pooled != nullptr) pooled[opix * C + b * 32 + lane] = t;
                if (pre) t = __fadd_rn(__fmul_rn(t, __ldg(pre_scale + b * 32 + lane)), __ldg(pre_shift + b * 32 + lane));
                s_[h] = __ballot_sync(0xffffffffu, t > 0.0f);
                m_[h] = __ballot_sync(0xffffffffu, t > 0.0f || (t < 0.0f && !pre_relu));
            }
            if (lane == 0) abits[((size_t)n_[u] * NCH + ch) * Ho * Wo + (size_t)ho_[u] * Wo + wo_[u]] = make_uint4(s_[0], s_[1], m_[0], m_[1]);
        }
    }
}

template <int NCH>
static int launch_avgpool2_pack(const float* x, int n, int h, int w, const float* pre_scale, const float* pre_shift,
                                int pre_relu, float* pooled, void* abits, cudaStream_t stream) {
    const int ho = h / 2, wo = w / 2;
    const long long pp = (long long)n * ho * wo, warps = (pp + AP_PPW - 1) / AP_PPW, blocks = (warps + 7) / 8;
    if (blocks > 0x7fffffffLL) return BNN_E_UNSUPPORTED;
    launch_pdl(avgpool2_pack_cl_kernel<NCH>, dim3((unsigned)blocks), dim3(256), 0, stream, x, pp, h, w, ho, wo, pre_scale, pre_shift, pre_relu, pooled, (uint4*)abits);
    count_launch(1);
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------
// weights: one CTA per output channel.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    // deterministic: fixed shuffle tree, then warp partials added in warp order
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double tot = 0.0;
    for (int i = 0; i < nw; ++i) tot += scratch[i];
    return tot;
}

__global__ void __launch_bounds__(128)
pack_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, int nch,
                   int center, int compute_alpha, uint32_t* __restrict__ wbits, uint32_t* __restrict__ wzero,
                   float* __restrict__ alpha, int* __restrict__ n_zero) {
    extern __shared__ double sm_d[];
    double* scratch = sm_d;                                 // [4]
    float* mean = reinterpret_cast<float*>(sm_d + 4);       // [taps]
    const int co = blockIdx.x;
    const float* wc = w + (long long)co * Cin * taps;
    const int nk = nch * taps;

    // mean over c_in per tap (reference ops.py:130-132), summed in fp64, rounded once
    for (int t = 0; t < taps; ++t) {
        double part = 0.0;
        if (center)
            for (int ci = threadIdx.x; ci < Cin; ci += blockDim.x) part += (double)wc[(long long)ci * taps + t];
        const double tot = center ? block_sum(part, scratch) : 0.0;
        if (threadIdx.x == 0) mean[t] = center ? (float)(tot / (double)Cin) : 0.0f;
    }
    __syncthreads();

    // alpha = mean |centred w| over (c_in, kh, kw)  (ops.py:116-123)
    if (compute_alpha) {
        double part = 0.0;
        const int per_out = Cin * taps;
        for (int i = threadIdx.x; i < per_out; i += blockDim.x) {
            const float v = center ? wc[i] - mean[i % taps] : wc[i];
            part += fabs((double)v);
        }
        const double tot = block_sum(part, scratch);
        if (threadIdx.x == 0) alpha[co] = (float)(tot / (double)per_out);
    } else if (threadIdx.x == 0) {
        alpha[co] = 1.0f;
    }

    // sign bits: item = (kstep, word) -> 32 channels
    int zeros = 0;
    for (int item = threadIdx.x; item < nk * 2; item += blockDim.x) {
        const int ks = item >> 1, word = item & 1;
        const int ch = ks / taps, t = ks - ch * taps;
        uint32_t bits = 0u, zbits = 0u;
        for (int b = 0; b < 32; ++b) {
            const int ci = ch * 64 + word * 32 + b;
            if (ci >= Cin) break;
            const float raw = wc[(long long)ci * taps + t];
            const float v = center ? raw - mean[t] : raw;
            const bool pos = v > 0.0f, neg = v < 0.0f;
            bits |= (uint32_t)pos << b;
            zbits |= (uint32_t)(!pos && !neg) << b;
            zeros += (!pos && !neg);
        }
        const long long widx = ((((long long)(co >> 5) * nk + ks) * 32) + (co & 31)) * 2 + word;
        wbits[widx] = bits;
        if (wzero != nullptr) wzero[widx] = zbits;       // plane of exactly-zero (centred) weights: sign 0 in the reference
    }
    if (n_zero != nullptr) {
        for (int o = 16; o > 0; o >>= 1) zeros += __shfl_down_sync(0xffffffffu, zeros, o);
        if ((threadIdx.x & 31) == 0 && zeros) atomicAdd(n_zero, zeros);
    }
}

}  // namespace bnn

using namespace bnn;

extern "C" size_t bnn_act_bits_bytes(int32_t n, int32_t c, int32_t h, int32_t w) {
    if (n <= 0 || c <= 0 || h <= 0 || w <= 0) return 0;
    return (size_t)n * ((c + 63) / 64) * h * w * 16;
}
extern "C" size_t bnn_weight_bits_bytes(int32_t c_out, int32_t c_in, int32_t kh, int32_t kw) {
    if (c_out <= 0 || c_in <= 0 || kh <= 0 || kw <= 0) return 0;
    return (size_t)((c_out + 31) / 32) * ((c_in + 63) / 64) * kh * kw * 32 * 8;
}

static int launch_pack(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, int32_t n, int32_t c,
                       int32_t h, int32_t w, int32_t pool, int32_t ceil_mode, const float* pre_scale,
                       const float* pre_shift, int32_t pre_relu, void* abits, void* stream_) {
    if (!x || !abits) return BNN_E_NULL;
    if ((pre_scale == nullptr) != (pre_shift == nullptr)) return BNN_E_NULL;
    if (n <= 0 || c <= 0 || h <= 0 || w <= 0 || pool < 0) return BNN_E_SHAPE;
    if (((uintptr_t)abits & 15) != 0) return BNN_E_ALIGN;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int nch = (c + 63) / 64;
    int ho = h, wo = w;
    if (pool > 1) {
        ho = ceil_mode ? (h + pool - 1) / pool : h / pool;
        wo = ceil_mode ? (w + pool - 1) / pool : w / pool;
        if (ho <= 0 || wo <= 0) return BNN_E_SHAPE;
    }
    const int threads = 256;
    if (sc == 1 && pool <= 1 && c % 64 == 0 && sw == c && sh == (int64_t)w * c && sn == (int64_t)h * w * c) {
        // dense NHWC: the streaming fast path
        const long long pixels = (long long)n * h * w;
        switch (c / 64) {
            case 1: return launch_pack_dense<1>(x, pixels, h * w, pre_scale, pre_shift, pre_relu, abits, stream);
            case 2: return launch_pack_dense<2>(x, pixels, h * w, pre_scale, pre_shift, pre_relu, abits, stream);
            case 4: return launch_pack_dense<4>(x, pixels, h * w, pre_scale, pre_shift, pre_relu, abits, stream);
            case 8: return launch_pack_dense<8>(x, pixels, h * w, pre_scale, pre_shift, pre_relu, abits, stream);
            default: break;
        }
    }
    if (sc == 1 && c >= 32) {
        // channels are contiguous: warp-per-pixel ballot kernel
        const long long pixels = (long long)n * ho * wo;
        const long long blocks = (pixels + 7) / 8;
        if (blocks > 0x7fffffffLL) return BNN_E_UNSUPPORTED;
        launch_pdl(pack_act_cl_kernel, dim3((unsigned)blocks), dim3(threads), 0, stream, x, sn, sh, sw, n, c, ho, wo, nch, pool, h, w,
                   pre_scale, pre_shift, pre_relu, (uint32_t*)abits);
        count_launch(1);
        return (int)cudaGetLastError();
    }
    const long long total = (long long)n * nch * ho * wo;
    const long long blocks = (total + threads - 1) / threads;
    if (blocks > 0x7fffffffLL) return BNN_E_UNSUPPORTED;
    launch_pdl(pack_act_kernel, dim3((unsigned)blocks), dim3(threads), 0, stream, x, sn, sc, sh, sw, n, c, ho, wo, nch, pool, h, w,
               pre_scale, pre_shift, pre_relu, (uint4*)abits);
    count_launch(1);
    return (int)cudaGetLastError();
}

extern "C" int bnn_pack_act_f32(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw,
                                int32_t n, int32_t c, int32_t h, int32_t w, const float* pre_scale,
                                const float* pre_shift, int32_t pre_relu, void* abits, void* stream) {
    return launch_pack(x, sn, sc, sh, sw, n, c, h, w, 0, 0, pre_scale, pre_shift, pre_relu, abits, stream);
}

extern "C" int bnn_avgpool_pack_f32(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw,
                                    int32_t n, int32_t c, int32_t h, int32_t w, int32_t k, int32_t ceil_mode,
                                    const float* pre_scale, const float* pre_shift, int32_t pre_relu, void* abits,
                                    void* stream) {
    if (k < 1) return BNN_E_SHAPE;
    return launch_pack(x, sn, sc, sh, sw, n, c, h, w, k, ceil_mode, pre_scale, pre_shift, pre_relu, abits, stream);
}

extern "C" int bnn_avgpool2_pack_cl_f32(const float* x, int32_t n, int32_t c, int32_t h, int32_t w, const float* pre_scale,
                                        const float* pre_shift, int32_t pre_relu, float* pooled_out, void* abits, void* stream_) {
    if (!x || !abits) return BNN_E_NULL;
    if ((pre_scale == nullptr) != (pre_shift == nullptr)) return BNN_E_NULL;
    if (n <= 0 || c <= 0 || h < 2 || w < 2) return BNN_E_SHAPE;
    if (((uintptr_t)abits & 15) != 0) return BNN_E_ALIGN;
    cudaStream_t stream = (cudaStream_t)stream_;
    switch (c) {
        case 64: return launch_avgpool2_pack<1>(x, n, h, w, pre_scale, pre_shift, pre_relu, pooled_out, abits, stream);
        case 128: return launch_avgpool2_pack<2>(x, n, h, w, pre_scale, pre_shift, pre_relu, pooled_out, abits, stream);
        case 256: return launch_avgpool2_pack<4>(x, n, h, w, pre_scale, pre_shift, pre_relu, pooled_out, abits, stream);
        case 512: return launch_avgpool2_pack<8>(x, n, h, w, pre_scale, pre_shift, pre_relu, pooled_out, abits, stream);
        default: return BNN_E_UNSUPPORTED;
    }
}

extern "C" int bnn_pack_weight_f32(const float* w, int32_t c_out, int32_t c_in, int32_t kh, int32_t kw,
                                   int32_t center, int32_t compute_alpha, void* wbits, float* alpha,
                                   int32_t* n_zero, void* stream_) {
    return bnn_pack_weight_ternary_f32(w, c_out, c_in, kh, kw, center, compute_alpha, wbits, nullptr, alpha, n_zero, stream_);
}

extern "C" int bnn_pack_weight_ternary_f32(const float* w, int32_t c_out, int32_t c_in, int32_t kh, int32_t kw,
                                           int32_t center, int32_t compute_alpha, void* wbits, void* wzero, float* alpha,
                                           int32_t* n_zero, void* stream_) {
    if (!w || !wbits || !alpha) return BNN_E_NULL;
    if (c_out <= 0 || c_in <= 0 || kh <= 0 || kw <= 0) return BNN_E_SHAPE;
    cudaStream_t stream = (cudaStream_t)stream_;
    cudaError_t e;
    // padded output channels / channels beyond c_in keep all-zero bits
    e = cudaMemsetAsync(wbits, 0, bnn_weight_bits_bytes(c_out, c_in, kh, kw), stream);
    if (e != cudaSuccess) return (int)e;
    if (wzero) {
        e = cudaMemsetAsync(wzero, 0, bnn_weight_bits_bytes(c_out, c_in, kh, kw), stream);
        if (e != cudaSuccess) return (int)e;
    }
    if (n_zero) {
        e = cudaMemsetAsync(n_zero, 0, sizeof(int32_t), stream);
        if (e != cudaSuccess) return (int)e;
    }
    const int taps = kh * kw, nch = (c_in + 63) / 64;
    const size_t smem = 4 * sizeof(double) + (size_t)taps * sizeof(float);
    if (smem > 48 * 1024) return BNN_E_UNSUPPORTED;
    pack_weight_kernel<<<c_out, 128, smem, stream>>>(w, c_out, c_in, taps, nch, center, compute_alpha,
                                                     (uint32_t*)wbits, (uint32_t*)wzero, alpha, n_zero);
    count_launch(1);
    return (int)cudaGetLastError();
}
