// bconv.cu -- XNOR/AND + popcount binary convolution with fused epilogue.
//
// Replaces, for packed operands, the body of bnn.layers.Conv2d.forward /
// Linear.forward (reference bnn/layers/conv.py:90-97, bnn/layers/linear.py:22-27):
//     y = ( alpha_w[co] * sum_k sign(x)_k * sign(w)_k + bias[co] ) * alpha_post[co]
// with sign(x) ternary (bnn/ops.py:66) and zero padding applied after sign().
//
// Mapping (B200-first, not a translation of anything in the reference -- the
// reference calls F.conv2d on dense fp32):
//   * CTA  = one "unit" of output pixels (TH rows x TW cols of one image) x one
//            tile of 32*C output channels, full K reduction.
//   * The input window of the unit (all 64-channel chunks, with halo) is staged
//     into shared memory by ONE 5-D TMA tensor load; out-of-bounds rows/cols
//     are zero-filled by the TMA unit, and a zero {s,m} pair has m = 0, i.e.
//     contributes nothing: the convolution's zero padding costs no instruction.
//     The weight tile arrives as 1-D TMA bulk copies on the same mbarrier.
//   * lanes <-> output channels, so a weight word is a per-lane LDS.64 and an
//     activation unit {s_lo,s_hi,m_lo,m_hi} is a warp-uniform (broadcast)
//     LDS.128: no bank conflicts for any stride / dilation.
//   * each warp owns groups of P consecutive output pixels of a row and keeps
//     a sliding window of input units in registers, P x C accumulators/thread.
//   * inner op per 32 bit-MACs: LOP3 (m & (s ^ t)) + POPC; the CSA mode folds
//     the three taps of a 3-wide kernel row with a 3:2 carry-save adder
//     (2 more LOP3) so that 3 words cost 2 POPC -- POPC is the slow pipe.
#include "common.cuh"

#include <cudaTypedefs.h>
#include <mutex>

namespace bnn {

struct ConvArgs {
    const uint4* abits;
    const uint32_t* cnt;
    const uint2* wbits;
    const float* scale;
    const float* bias;
    const float* post;
    float* out;
    long long on, oc, oh, ow;
    int N, Cin, H, W, Cout, KH, KW, SH, SW, PH, PW, DH, DW, Ho, Wo;
    int nch, nk, nblk32;          // 64-ch chunks, k-steps, 32-channel output blocks
    int TH, TW, BH, BW;           // output tile, input box
    int gpr, G;                   // pixel groups per tile row, per unit
    int tiles_h, tiles_w;
    unsigned act_bytes, w_bytes;  // bytes per staged activation box / per 32-channel weight block
    int stage_ldg;
};

__device__ __forceinline__ int word_dis(uint32_t m, uint32_t s, uint32_t t) { return __popc(m & (s ^ t)); }
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (c & (a ^ b)); }

template <int P, int C, int KWT, int SWT, int MODE>
__global__ void __launch_bounds__(256, 2)
bconv_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ ConvArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    uint4* act = reinterpret_cast<uint4*>(smem + 128);
    uint2* wsm = reinterpret_cast<uint2*>(smem + 128 + ((a.act_bytes + 127u) & ~127u));
    // per-warp transpose buffer for the epilogue: 32 channels x PITCH pixels (odd pitch: no conflicts)
    constexpr int PITCH = P | 1;
    float* stage_all = reinterpret_cast<float*>(smem + 128 + ((a.act_bytes + 127u) & ~127u) + (size_t)C * a.w_bytes);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    int unit = blockIdx.x;
    const int tw_i = unit % a.tiles_w;
    unit /= a.tiles_w;
    const int th_i = unit % a.tiles_h;
    const int n = unit / a.tiles_h;
    const int ho0 = th_i * a.TH, wo0 = tw_i * a.TW;
    const int hi0 = ho0 * a.SH - a.PH, wi0 = wo0 * a.SW - a.PW;
    const int blk0 = blockIdx.y * C;                 // first 32-channel block of this CTA
    const int nk32 = a.nk * 32;

    // ---------------- stage the activation window + weight tile ----------------
    if (!a.stage_ldg) {
        if (threadIdx.x == 0) {
            prefetch_tensormap(&tmap);
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const int nvalid = min(C, a.nblk32 - blk0);
            mbar_expect_tx(bar, a.act_bytes + (unsigned)nvalid * a.w_bytes);
            tma_load_5d(act, &tmap, bar, 0, wi0, hi0, 0, n);
            for (int j = 0; j < nvalid; ++j)
                bulk_load_1d(wsm + (size_t)j * nk32, a.wbits + (size_t)(blk0 + j) * nk32, a.w_bytes, bar);
        }
        mbar_wait(bar, 0);
    } else {
        const int units = a.nch * a.BH * a.BW;
        for (int i = threadIdx.x; i < units; i += blockDim.x) {
            const int c = i % a.BW;
            const int rr = (i / a.BW) % a.BH;
            const int ch = i / (a.BW * a.BH);
            const int hi = hi0 + rr, wi = wi0 + c;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if ((unsigned)hi < (unsigned)a.H && (unsigned)wi < (unsigned)a.W)
                v = a.abits[(((size_t)n * a.nch + ch) * a.H + hi) * a.W + wi];
            act[i] = v;
        }
        for (int j = 0; j < C; ++j) {
            if (blk0 + j >= a.nblk32) break;
            const uint2* src = a.wbits + (size_t)(blk0 + j) * nk32;
            for (int i = threadIdx.x; i < nk32; i += blockDim.x) wsm[(size_t)j * nk32 + i] = src[i];
        }
        __syncthreads();
    }

    // ---------------- per-lane epilogue constants ----------------
    float e_scale[C], e_bias[C], e_post[C];
    bool c_ok[C];
#pragma unroll
    for (int j = 0; j < C; ++j) {
        const int c = (blk0 + j) * 32 + lane;
        c_ok[j] = c < a.Cout;
        e_scale[j] = (c_ok[j] && a.scale) ? __ldg(a.scale + c) : 1.0f;
        e_bias[j] = (c_ok[j] && a.bias) ? __ldg(a.bias + c) : 0.0f;
        e_post[j] = (c_ok[j] && a.post) ? __ldg(a.post + c) : 1.0f;
    }
    const int SW = (KWT > 0) ? SWT : a.SW;
    const int KW = (KWT > 0) ? KWT : a.KW;
    const int DW = (KWT > 0) ? 1 : a.DW;

    // ---------------- pixel groups ----------------
    for (int g = warp; g < a.G; g += nwarps) {
        const int r = g / a.gpr;
        const int wq = (g - r * a.gpr) * P;      // first output column inside the tile
        const int ho = ho0 + r;
        const int wo_first = wo0 + wq;
        if (ho >= a.Ho || wo_first >= a.Wo) continue;   // warp-uniform

        int acc[P][C];
#pragma unroll
        for (int p = 0; p < P; ++p)
#pragma unroll
            for (int j = 0; j < C; ++j) acc[p][j] = 0;

        for (int ch = 0; ch < a.nch; ++ch) {
            for (int kh = 0; kh < a.KH; ++kh) {
                const uint4* arow = act + (size_t)(ch * a.BH + r * a.SH + kh * a.DH) * a.BW + wq * SW;
                const uint2* wrow = wsm + (size_t)((ch * a.KH + kh) * KW) * 32 + lane;
                if constexpr (KWT > 0) {
                    constexpr int U = (P - 1) * SWT + KWT;
                    constexpr bool WINDOW = (U <= 12);
                    uint4 u[WINDOW ? U : 1];
                    if constexpr (WINDOW) {
#pragma unroll
                        for (int i = 0; i < U; ++i) u[i] = arow[i];
                    }
                    if constexpr (MODE == 1 && KWT == 3) {
                        // 3:2 carry-save over the three taps of this kernel row
                        uint2 t[3][C];
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw)
#pragma unroll
                            for (int j = 0; j < C; ++j) t[kw][j] = wrow[(size_t)j * nk32 + kw * 32];
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            uint4 v0, v1, v2;
                            if constexpr (WINDOW) {
                                v0 = u[p * SWT]; v1 = u[p * SWT + 1]; v2 = u[p * SWT + 2];
                            } else {
                                v0 = arow[p * SWT]; v1 = arow[p * SWT + 1]; v2 = arow[p * SWT + 2];
                            }
#pragma unroll
                            for (int j = 0; j < C; ++j) {
                                const uint32_t x0 = v0.z & (v0.x ^ t[0][j].x), x1 = v1.z & (v1.x ^ t[1][j].x),
                                               x2 = v2.z & (v2.x ^ t[2][j].x);
                                const uint32_t y0 = v0.w & (v0.y ^ t[0][j].y), y1 = v1.w & (v1.y ^ t[1][j].y),
                                               y2 = v2.w & (v2.y ^ t[2][j].y);
                                const int ones = __popc(x0 ^ x1 ^ x2) + __popc(y0 ^ y1 ^ y2);
                                const int twos = __popc(maj3(x0, x1, x2)) + __popc(maj3(y0, y1, y2));
                                acc[p][j] += ones + 2 * twos;
                            }
                        }
                    } else {
#pragma unroll
                        for (int kw = 0; kw < KWT; ++kw) {
                            uint2 t[C];
#pragma unroll
                            for (int j = 0; j < C; ++j) t[j] = wrow[(size_t)j * nk32 + kw * 32];
#pragma unroll
                            for (int p = 0; p < P; ++p) {
                                uint4 v;
                                if constexpr (WINDOW) v = u[p * SWT + kw];
                                else v = arow[p * SWT + kw];
#pragma unroll
                                for (int j = 0; j < C; ++j)
                                    acc[p][j] += word_dis(v.z, v.x, t[j].x) + word_dis(v.w, v.y, t[j].y);
                            }
                        }
                    }
                } else {
                    for (int kw = 0; kw < KW; ++kw) {
                        uint2 t[C];
#pragma unroll
                        for (int j = 0; j < C; ++j) t[j] = wrow[(size_t)j * nk32 + kw * 32];
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            const uint4 v = arow[p * SW + kw * DW];
#pragma unroll
                            for (int j = 0; j < C; ++j)
                                acc[p][j] += word_dis(v.z, v.x, t[j].x) + word_dis(v.w, v.y, t[j].y);
                        }
                    }
                }
            }
        }

        // number of non-zero inputs under each pixel's receptive field (lane p <-> pixel p)
        int msum = 0;
        if (lane < P) {
            const int wo = wo_first + lane;
            if (wo < a.Wo) {
                for (int kh = 0; kh < a.KH; ++kh) {
                    const int hi = ho * a.SH - a.PH + kh * a.DH;
                    if ((unsigned)hi >= (unsigned)a.H) continue;
                    for (int kw = 0; kw < KW; ++kw) {
                        const int wi = wo * SW - a.PW + kw * DW;
                        if ((unsigned)wi < (unsigned)a.W) msum += (int)__ldg(a.cnt + ((size_t)n * a.H + hi) * a.W + wi);
                    }
                }
            }
        }

        // fused epilogue: y = (alpha_w * dot + bias) * alpha_post, same order as the reference
        float* obase = a.out + (long long)n * a.on + (long long)ho * a.oh;
        int ms[P];
#pragma unroll
        for (int p = 0; p < P; ++p) ms[p] = __shfl_sync(0xffffffffu, msum, p);
        if (a.ow != 1) {
            // channel-contiguous output (Linear's [rows, out]): lanes <-> channels is already coalesced
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const int wo = wo_first + p;
                if (wo >= a.Wo) break;
#pragma unroll
                for (int j = 0; j < C; ++j) {
                    if (!c_ok[j]) continue;
                    const int c = (blk0 + j) * 32 + lane;
                    float y = __fmul_rn(e_scale[j], (float)(ms[p] - 2 * acc[p][j]));
                    if (a.bias) y = __fadd_rn(y, e_bias[j]);
                    if (a.post) y = __fmul_rn(y, e_post[j]);
                    obase[(long long)c * a.oc + (long long)wo * a.ow] = y;
                }
            }
        } else {
            // pixel-contiguous output (NCHW): transpose each 32-channel block through shared memory so a
            // store instruction covers whole 32-byte runs of P pixels instead of 32 scattered words
            float* stg = stage_all + warp * (32 * PITCH);
            constexpr int PW = (P > 4) ? 8 : 4;          // lanes per channel row when reading back
            constexpr int ROWS = 32 / PW;                // channel rows per store instruction
            const int pr = lane % PW, rr = lane / PW;
#pragma unroll
            for (int j = 0; j < C; ++j) {
                __syncwarp();
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    float y = __fmul_rn(e_scale[j], (float)(ms[p] - 2 * acc[p][j]));
                    if (a.bias) y = __fadd_rn(y, e_bias[j]);
                    if (a.post) y = __fmul_rn(y, e_post[j]);
                    stg[lane * PITCH + p] = y;
                }
                __syncwarp();
                const int cblk = (blk0 + j) * 32;
#pragma unroll
                for (int r0 = 0; r0 < 32; r0 += ROWS) {
                    const int cl = r0 + rr;
                    const int c = cblk + cl, wo = wo_first + pr;
                    if (pr < P && wo < a.Wo && c < a.Cout)
                        obase[(long long)c * a.oc + wo] = stg[cl * PITCH + pr];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// host side: tensor map, planner, dispatch
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

struct Plan {
    int P, C, kwt, swt, mode;
    int TH, TW, BH, BW, NW, gpr, G, tiles_h, tiles_w;
    size_t smem;
};

typedef void (*KernelFn)(const CUtensorMap, const ConvArgs);

template <int P, int C, int KWT, int SWT, int MODE>
static KernelFn kernel_ptr() { return bconv_kernel<P, C, KWT, SWT, MODE>; }

template <int KWT, int SWT, int MODE>
static KernelFn pick_pc(int P, int C) {
#define BNN_PC(p, c) if (P == p && C == c) return kernel_ptr<p, c, KWT, SWT, MODE>();
    BNN_PC(8, 4) BNN_PC(8, 2) BNN_PC(8, 1)
    BNN_PC(7, 4) BNN_PC(7, 2) BNN_PC(7, 1)
    BNN_PC(4, 4) BNN_PC(4, 2) BNN_PC(4, 1)
#undef BNN_PC
    return nullptr;
}

static KernelFn pick_kernel(const Plan& p) {
    if (p.kwt == 3 && p.swt == 1) return p.mode ? pick_pc<3, 1, 1>(p.P, p.C) : pick_pc<3, 1, 0>(p.P, p.C);
    if (p.kwt == 3 && p.swt == 2) return p.mode ? pick_pc<3, 2, 1>(p.P, p.C) : pick_pc<3, 2, 0>(p.P, p.C);
    if (p.kwt == 1 && p.swt == 1) return pick_pc<1, 1, 0>(p.P, p.C);
    return pick_pc<0, 0, 0>(p.P, p.C);
}

static int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Choose tile shape for one layer.  Small search, host only, microseconds.
static int make_plan(const bnn_conv_geom& g, int Ho, int Wo, uint32_t flags, int sms, Plan* out) {
    Plan pl{};
    const int nch = ceil_div(g.c_in, 64), nk = nch * g.kh * g.kw;
    if (nch > 256) return BNN_E_UNSUPPORTED;
    // specialisation on kernel width / horizontal stride (dilation_w must be 1)
    pl.kwt = 0; pl.swt = 0;
    if (g.dil_w == 1) {
        if (g.kw == 3 && (g.stride_w == 1 || g.stride_w == 2)) { pl.kwt = 3; pl.swt = g.stride_w; }
        else if (g.kw == 1 && g.stride_w == 1) { pl.kwt = 1; pl.swt = 1; }
    }
    pl.mode = (pl.kwt == 3 && !(flags & BNN_F_NO_CSA)) ? 1 : 0;

    // output channels per lane
    const int nblk32 = ceil_div(g.c_out, 32);
    int C = nblk32 >= 4 ? 4 : (nblk32 >= 2 ? 2 : 1);
    const size_t smem_cap = 200 * 1024;
    while (C > 1 && (size_t)C * nk * 256 > 96 * 1024) C >>= 1;
    const size_t wbytes = (size_t)C * nk * 256;
    if (wbytes + 8192 > smem_cap) return BNN_E_UNSUPPORTED;
    const size_t act_budget = (wbytes <= 64 * 1024 ? 110 * 1024 : smem_cap) - wbytes - 256 - 8 * 32 * 9 * 4;

    // pixels per group: least padding waste, then prefer a register window, then larger
    const int cand[3] = {8, 7, 4};
    int bestP = 0; double bestW = 1e9;
    for (int i = 0; i < 3; ++i) {
        const int P = cand[i];
        const int sw = pl.kwt ? pl.swt : 1, kw = pl.kwt ? pl.kwt : 1;
        const bool window = ((P - 1) * sw + kw) <= 12;
        double waste = (double)ceil_div(Wo, P) * P / Wo;
        waste += window ? 0.0 : 0.02;
        waste -= 0.001 * P;
        if (waste < bestW) { bestW = waste; bestP = P; }
    }
    pl.P = bestP; pl.C = C;

    // tile width: whole row if the input box fits the 256-element TMA box limit
    const int wo_pad = ceil_div(Wo, pl.P) * pl.P;
    int TW = wo_pad;
    while ((TW - 1) * g.stride_w + (g.kw - 1) * g.dil_w + 1 > 256 && TW > pl.P) TW -= pl.P;
    if ((TW - 1) * g.stride_w + (g.kw - 1) * g.dil_w + 1 > 256) return BNN_E_UNSUPPORTED;
    pl.TW = TW;
    pl.BW = (TW - 1) * g.stride_w + (g.kw - 1) * g.dil_w + 1;
    pl.gpr = TW / pl.P;
    pl.tiles_w = ceil_div(Wo, TW);

    // tile height / warps per CTA: minimise (row padding) x (idle warps in the last round)
    double best = 1e18; int bTH = 0, bNW = 0;
    const long long cout_tiles = ceil_div(nblk32, C);
    for (int NW = 8; NW >= 7; --NW) {
        for (int TH = 1; TH <= Ho; ++TH) {
            const int BH = (TH - 1) * g.stride_h + (g.kh - 1) * g.dil_h + 1;
            if (BH > 256) break;
            if ((size_t)nch * BH * pl.BW * 16 > act_budget) break;
            const int G = TH * pl.gpr;
            const int rounds = ceil_div(G, NW);
            if (rounds > 8 && TH > 1) break;
            const double row_waste = (double)ceil_div(Ho, TH) * TH / Ho;
            const double warp_waste = (double)rounds * NW / G;
            const double halo = (double)BH / ((TH - 1) * g.stride_h + 1);     // staging re-reads
            const long long ctas = (long long)g.n * ceil_div(Ho, TH) * pl.tiles_w * cout_tiles;
            const double fill = ctas < 2LL * sms ? (double)(2LL * sms) / (double)ctas : 1.0;
            const double score = row_waste * warp_waste * (1.0 + 0.01 * halo) * (1.0 + 0.25 * (fill - 1.0)) *
                                 (1.0 + 0.02 / rounds) * (NW == 8 ? 1.0 : 1.01);
            if (score < best) { best = score; bTH = TH; bNW = NW; }
        }
    }
    if (bTH == 0) return BNN_E_UNSUPPORTED;
    pl.TH = bTH; pl.NW = bNW;
    pl.BH = (bTH - 1) * g.stride_h + (g.kh - 1) * g.dil_h + 1;
    pl.G = bTH * pl.gpr;
    pl.tiles_h = ceil_div(Ho, bTH);
    const size_t act_bytes = (size_t)nch * pl.BH * pl.BW * 16;
    pl.smem = 128 + ((act_bytes + 127) & ~(size_t)127) + wbytes + (size_t)pl.NW * 32 * (pl.P | 1) * 4;
    *out = pl;
    return 0;
}

static int launch_bconv(const void* abits, const uint32_t* cnt, const void* wbits, const float* scale,
                        const float* bias, const float* post, float* out, int64_t on, int64_t oc, int64_t oh,
                        int64_t ow, const bnn_conv_geom& g, uint32_t flags, cudaStream_t stream) {
    if (!abits || !cnt || !wbits || !out) return BNN_E_NULL;
    if (g.n <= 0 || g.c_in <= 0 || g.h <= 0 || g.w <= 0 || g.c_out <= 0 || g.kh <= 0 || g.kw <= 0 ||
        g.stride_h <= 0 || g.stride_w <= 0 || g.pad_h < 0 || g.pad_w < 0 || g.dil_h <= 0 || g.dil_w <= 0)
        return BNN_E_SHAPE;
    if (((uintptr_t)abits & 15) || ((uintptr_t)wbits & 15)) return BNN_E_ALIGN;
    const int Ho = out_dim(g.h, g.kh, g.stride_h, g.pad_h, g.dil_h);
    const int Wo = out_dim(g.w, g.kw, g.stride_w, g.pad_w, g.dil_w);
    if (Ho <= 0 || Wo <= 0) return BNN_E_SHAPE;

    int dev = 0, sms = 148;
    cudaError_t ce = cudaGetDevice(&dev);
    if (ce != cudaSuccess) return (int)ce;
    static int cached_sms[64] = {0};
    if (dev < 64 && cached_sms[dev]) sms = cached_sms[dev];
    else {
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (dev < 64) cached_sms[dev] = sms;
    }

    Plan pl;
    int rc = make_plan(g, Ho, Wo, flags, sms, &pl);
    if (rc) return rc;
    KernelFn fn = pick_kernel(pl);
    if (!fn) return BNN_E_UNSUPPORTED;

    ConvArgs a{};
    a.abits = (const uint4*)abits; a.cnt = cnt; a.wbits = (const uint2*)wbits;
    a.scale = scale; a.bias = bias; a.post = post; a.out = out;
    a.on = on; a.oc = oc; a.oh = oh; a.ow = ow;
    a.N = g.n; a.Cin = g.c_in; a.H = g.h; a.W = g.w; a.Cout = g.c_out; a.KH = g.kh; a.KW = g.kw;
    a.SH = g.stride_h; a.SW = g.stride_w; a.PH = g.pad_h; a.PW = g.pad_w; a.DH = g.dil_h; a.DW = g.dil_w;
    a.Ho = Ho; a.Wo = Wo;
    a.nch = ceil_div(g.c_in, 64); a.nk = a.nch * g.kh * g.kw; a.nblk32 = ceil_div(g.c_out, 32);
    a.TH = pl.TH; a.TW = pl.TW; a.BH = pl.BH; a.BW = pl.BW; a.gpr = pl.gpr; a.G = pl.G;
    a.tiles_h = pl.tiles_h; a.tiles_w = pl.tiles_w;
    a.act_bytes = (unsigned)((size_t)a.nch * pl.BH * pl.BW * 16);
    a.w_bytes = (unsigned)((size_t)a.nk * 256);
    a.stage_ldg = (flags & BNN_F_STAGE_LDG) ? 1 : 0;

    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if (!a.stage_ldg) {
        EncodeTiledFn enc = encode_tiled_fn();
        if (!enc) return BNN_E_DRIVER;
        // abits as a 5-D u32 tensor, fastest first: {4 words, W, H, chunks, N}
        const cuuint64_t gdim[5] = {4, (cuuint64_t)g.w, (cuuint64_t)g.h, (cuuint64_t)a.nch, (cuuint64_t)g.n};
        const cuuint64_t gstr[4] = {16, (cuuint64_t)g.w * 16, (cuuint64_t)g.w * g.h * 16,
                                    (cuuint64_t)g.w * g.h * a.nch * 16};
        const cuuint32_t box[5] = {4, (cuuint32_t)pl.BW, (cuuint32_t)pl.BH, (cuuint32_t)a.nch, 1};
        const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 5, const_cast<void*>(abits), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return BNN_E_DRIVER;
    }

    if (pl.smem > 48 * 1024) {
        ce = cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
        if (ce != cudaSuccess) return (int)ce;
    }
    const long long units = (long long)g.n * pl.tiles_h * pl.tiles_w;
    const int cout_tiles = ceil_div(a.nblk32, pl.C);
    if (units > 0x7fffffffLL || cout_tiles > 65535) return BNN_E_UNSUPPORTED;
    dim3 grid((unsigned)units, (unsigned)cout_tiles, 1);
    fn<<<grid, pl.NW * 32, pl.smem, stream>>>(tmap, a);
    count_launch(1);
    return (int)cudaGetLastError();
}

}  // namespace bnn

using namespace bnn;

extern "C" int bnn_bconv2d_fwd(const void* abits, const uint32_t* cnt, const void* wbits, const float* scale,
                               const float* bias, const float* post, float* out, int64_t on, int64_t oc,
                               int64_t oh, int64_t ow, const bnn_conv_geom* geom, uint32_t flags, void* stream) {
    if (!geom) return BNN_E_NULL;
    return launch_bconv(abits, cnt, wbits, scale, bias, post, out, on, oc, oh, ow, *geom, flags,
                        (cudaStream_t)stream);
}

extern "C" int bnn_blinear_fwd(const void* abits, const uint32_t* cnt, const void* wbits, const float* scale,
                               const float* bias, const float* post, float* out, int32_t rows,
                               int32_t in_features, int32_t out_features, uint32_t flags, void* stream) {
    // rows play the role of the image width: x is [1, in, 1, rows], out is stored [rows, out]
    bnn_conv_geom g{1, in_features, 1, rows, out_features, 1, 1, 1, 1, 0, 0, 1, 1};
    return launch_bconv(abits, cnt, wbits, scale, bias, post, out, 0, 1, 0, (int64_t)out_features, g, flags,
                        (cudaStream_t)stream);
}
