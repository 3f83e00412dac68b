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
#pragma once
#include <type_traits>
#include "common.cuh"

namespace bnn {

struct Epi {                       // device view of bnn_epilogue
    const float *scale, *bias, *post, *bn_scale, *bn_shift, *slope, *nx_scale, *nx_shift;
    const float* res;
    long long rn, rc, rh, rw;
    float* out;
    long long on, oc, oh, ow;
    uint4* obits;
    int act, res_after_act, ochunks, nx_relu, bits_pre_res;
};

struct ConvArgs {
    const uint4* abits;
    const uint2* wbits;
    Epi e;
    int N, Cin, H, W, Cout, KH, KW, SH, SW, PH, PW, DH, DW, Ho, Wo;
    int nch, nk, nblk32;          // 64-ch chunks, k-steps, 32-channel output blocks
    int TH, TW, BH, BW;           // output tile, input box
    int gpr, G;                   // pixel groups per tile row, per unit
    int tiles_h, tiles_w;
    unsigned act_bytes, w_bytes;  // bytes per staged activation box / per 32-channel weight block
    int stage_ldg;
    // split-K launches contract a chunk range of a wider layer: the packed tensors keep the FULL layer's strides
    int nch_total;                // chunks per image of the abits tensor (>= nch)
    unsigned w_blk_stride;        // uint2 elements between consecutive 32-channel weight blocks (nk_total * 32 >= nk * 32)
};

// per-CTA table of per-channel epilogue constants in shared memory: [EP_N][32*C]
//   EPI == 0 (reference epilogue, exact order):  y = (k0 * dot + k1) * k2         k = scale, bias, post
//   EPI == 1 (cross-module fusion):              z = fma(k0, dot, k1)            k0/k1 fold scale, bias, post, BN
//                                                k2 = PReLU slope, k3/k4 = next layer's pre-sign affine
enum { EP_N = 5 };

// (a dropped read-only load instead of the prefetch instruction was measured slower: the register it names is waited on)
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__device__ __forceinline__ int word_dis(uint32_t m, uint32_t s, uint32_t t) { return __popc(m & (s ^ t)); }
// (ms + magic) - 2 * acc as one integer multiply-add
__device__ __forceinline__ int msm2(int msm, int acc) {
    int r;
    asm("mad.lo.s32 %0, %1, -2, %2;" : "=r"(r) : "r"(acc), "r"(msm));
    return r;
}
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (c & (a ^ b)); }

// CTAs per SM the instance is compiled for.  1x1 kernels with small tiles keep few values live and their layers are
// latency-bound (short K loops between a TMA prologue and an HBM-fed epilogue): three resident CTAs (80 registers)
// overlap those phases better than two; everything else gets the full 128 registers.
#ifndef BNN_CONV_REGS_C2
#define BNN_CONV_REGS_C2 96        // registers per thread of the C <= 2 instances: 5 resident 4-warp CTAs = 20 warps per SM (128: 16)
#endif
__host__ __device__ constexpr int bconv_min_ctas(int P, int C, int KWT) { return (KWT == 1 && P * C <= 16) ? 3 : 2; }
// registers per thread an instance is compiled for, and the warps per SM the register file then holds
__host__ __device__ constexpr int bconv_regs(int P, int C, int KWT) {
    return (KWT == 1 && P * C <= 16) ? 80 : (C <= 2 ? BNN_CONV_REGS_C2 : 128);
}
__host__ __device__ constexpr int bconv_warps_per_sm(int P, int C, int KWT) {
    return (KWT == 1 && P * C <= 16) ? 24 : 2048 / bconv_regs(P, C, KWT);
}

template <int P, int C, int KWT, int SWT, int MODE, int EPI>
#if BNN_CONV_REGS_C2 == 128
__global__ void __launch_bounds__(256, bconv_min_ctas(P, C, KWT))
#else
__global__ void __maxnreg__(bconv_regs(P, C, KWT))
#endif
bconv_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ ConvArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    // EPI 0: reference epilogue.  EPI 1: fused epilogue, any strides.  EPI 2: fused epilogue with channel-contiguous
    // (NHWC) residual and output -- what the fused engine always uses: the pixel-contiguous transposes are compiled out
    // EPI 3: the lean NHWC epilogue of residual blocks, launched only when every pixel group and every channel block is
    // complete -- no bounds predicates, no stride arithmetic beyond one multiply per access, planes through the warp's
    // staging area.  EPI 3, fast form: BatchNorm, optional shortcut add BEFORE a ReLU, planes where "non-zero" ==
    // "positive".  EPI 4, general form: any activation, shortcut before or after it, optional affine in front of the
    // next sign() (its own instance so that the fast form stays small).
    constexpr bool FUSED = EPI >= 1, CL = EPI >= 2, LEAN = EPI >= 3, LEAN_GENERAL = EPI == 4;
    // int -> float without the conversion pipe (I2F shares the quarter-rate XU pipe with POPC): for |k| < 2^22 the bit
    // pattern 0x4B400000 + k IS the float 12582912 + k, and subtracting 12582912.0f is exact -- the same value I2F gives
    constexpr int MAGIC_I = 0x4B400000;
    constexpr float MAGIC_F = 12582912.0f;
    constexpr int PITCH = P | 1;      // odd pitch: conflict-free transposes
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    uint4* act = reinterpret_cast<uint4*>(smem + 128);
    unsigned char* after_act = smem + 128 + ((a.act_bytes + 127u) & ~127u);
    uint2* wsm = reinterpret_cast<uint2*>(after_act);
    float* stage_all = reinterpret_cast<float*>(after_act + (size_t)C * a.w_bytes);
    float* epc = stage_all + (blockDim.x >> 5) * (32 * PITCH);      // [EP_N][32*C]
    int* ms_s = reinterpret_cast<int*>(epc + EP_N * 32 * C);         // [TH][TW] non-zero inputs per window

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
    // Programmatic dependent launch: this CTA may become resident while the previous kernel still drains its last wave.
    // Up to pdl_wait() it touches no global memory at all (barrier init, tensor-map prefetch from the parameter space,
    // index arithmetic): weights and per-channel constants may have been written by the kernel just before this one.
    pdl_launch_dependents();
    if (!a.stage_ldg) {
        if (threadIdx.x == 0) {
            prefetch_tensormap(&tmap);
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
        pdl_wait();
        if (threadIdx.x == 0) {
            const int nvalid = min(C, a.nblk32 - blk0);
            mbar_expect_tx(bar, a.act_bytes + (unsigned)nvalid * a.w_bytes);
            tma_load_5d(act, &tmap, bar, 0, wi0, hi0, 0, n);
            for (int j = 0; j < nvalid; ++j)
                bulk_load_1d(wsm + (size_t)j * nk32, a.wbits + (size_t)(blk0 + j) * a.w_blk_stride, a.w_bytes, bar);
        }
    } else {
        pdl_wait();
        const int units = a.nch * a.BH * a.BW;
        for (int i = threadIdx.x; i < units; i += blockDim.x) {
            const int c = i % a.BW;
            const int rr = (i / a.BW) % a.BH;
            const int ch = i / (a.BW * a.BH);
            const int hi = hi0 + rr, wi = wi0 + c;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if ((unsigned)hi < (unsigned)a.H && (unsigned)wi < (unsigned)a.W)
                v = a.abits[(((size_t)n * a.nch_total + ch) * a.H + hi) * a.W + wi];
            act[i] = v;
        }
        for (int j = 0; j < C; ++j) {
            if (blk0 + j >= a.nblk32) break;
            const uint2* src = a.wbits + (size_t)(blk0 + j) * a.w_blk_stride;
            for (int i = threadIdx.x; i < nk32; i += blockDim.x) wsm[(size_t)j * nk32 + i] = src[i];
        }
    }
    // per-channel epilogue constants -> shared memory (overlaps the TMA flight time)
    for (int i = threadIdx.x; i < 32 * C; i += blockDim.x) {
        const int c = blk0 * 32 + i;
        const bool ok = c < a.Cout;
        float k0 = (ok && a.e.scale) ? __ldg(a.e.scale + c) : 1.0f;
        float k1 = (ok && a.e.bias) ? __ldg(a.e.bias + c) : 0.0f;
        const float post = (ok && a.e.post) ? __ldg(a.e.post + c) : 1.0f;
        if constexpr (EPI == 0) {
            epc[0 * 32 * C + i] = k0; epc[1 * 32 * C + i] = k1; epc[2 * 32 * C + i] = post;
        } else {
            // fold (scale*dot + bias)*post and the eval BatchNorm into one multiply-add (fixed order, see oracle)
            k0 = __fmul_rn(k0, post); k1 = __fmul_rn(k1, post);
            if (a.e.bn_scale) {
                const float g = ok ? __ldg(a.e.bn_scale + c) : 1.0f, h = ok ? __ldg(a.e.bn_shift + c) : 0.0f;
                k0 = __fmul_rn(k0, g);
                k1 = __fadd_rn(__fmul_rn(k1, g), h);
            }
            epc[0 * 32 * C + i] = k0; epc[1 * 32 * C + i] = k1;
            epc[2 * 32 * C + i] = (ok && a.e.slope) ? __ldg(a.e.slope + c) : 0.0f;
            epc[3 * 32 * C + i] = (ok && a.e.nx_scale) ? __ldg(a.e.nx_scale + c) : 1.0f;
            epc[4 * 32 * C + i] = (ok && a.e.nx_shift) ? __ldg(a.e.nx_shift + c) : 0.0f;
        }
    }
    __syncthreads();
    if (!a.stage_ldg) mbar_wait(bar, 0);

    const int SW = (KWT > 0) ? SWT : a.SW;
    const int KW = (KWT > 0) ? KWT : a.KW;
    const int DW = (KWT > 0) ? 1 : a.DW;
    // number of non-zero inputs under every output pixel's window (popc of the m planes already staged):
    // once per CTA, one or two pixels per thread, instead of POPCs per output channel
    for (int i = threadIdx.x; i < a.TH * a.TW; i += blockDim.x) {
        const int r = i / a.TW, q = i - r * a.TW;
        int cnt = 0;
        for (int ch = 0; ch < a.nch; ++ch)
            for (int kh = 0; kh < a.KH; ++kh) {
                const uint4* arow = act + (size_t)(ch * a.BH + r * a.SH + kh * a.DH) * a.BW + q * SW;
                for (int kw = 0; kw < KW; ++kw) {
                    const uint4 v = arow[kw * DW];
                    cnt += __popc(v.z) + __popc(v.w);
                }
            }
        ms_s[i] = cnt + MAGIC_I;               // every epilogue converts (ms + magic) - 2 acc by bit pattern, see MAGIC_I
    }
    __syncthreads();

    float* stg = stage_all + warp * (32 * PITCH);
    constexpr int PW = (P > 4) ? 8 : 4;          // lanes per channel row in the transposed phases
    constexpr int ROWS = 32 / PW;                // channel rows per load/store instruction
    const int pr = lane % PW, rr = lane / PW;

    // ---------------- pixel groups ----------------
    // per-image bases (the image is fixed for the CTA); inside an image 32-bit element offsets suffice (host-checked)
    const float* res_n = (FUSED && a.e.res) ? a.e.res + (long long)n * a.e.rn : nullptr;
    float* out_n = a.e.out ? a.e.out + (long long)n * a.e.on : nullptr;
    const int e_rc = CL ? 1 : (int)a.e.rc, e_rh = (int)a.e.rh, e_rw = (int)a.e.rw;
    const int e_oc = CL ? 1 : (int)a.e.oc, e_oh = (int)a.e.oh, e_ow = (int)a.e.ow;
    const long long rwb = (long long)e_rw * 4, owb = (long long)e_ow * 4;        // byte steps between pixels
    const int a_chstep = a.BH * a.BW, a_khstep = a.DH * a.BW;     // activation rows: per chunk, per kernel row
    // lean plane store: lane l owns unit (chunk l / P, pixel l % P) of every group -- per-lane part of the unit index
    size_t obits_base = 0;
    if constexpr (EPI == 3) {
        const int lj = lane / P, lp = lane - lj * P;
        obits_base = ((size_t)n * a.e.ochunks + (blk0 >> 1) + lj) * ((size_t)a.Ho * a.Wo) + lp;
    }
    // shortcut prefetch (channels-last): lane -> (32-channel block lane / P, pixel lane % P) of every group
    int pf_off = -1, pf_p = 0;
    if (FUSED && lane < P * C) {
        const int j = lane / P;
        pf_p = lane - j * P;
        if ((blk0 + j) * 32 < a.Cout) pf_off = (blk0 + j) * 32 * e_rc + pf_p * e_rw;
    }
    int g_row = warp / a.gpr, g_col = warp - g_row * a.gpr;      // one division per warp, then incremental
    const int step_row = nwarps / a.gpr, step_col = nwarps - step_row * a.gpr;
    for (int g = warp; g < a.G; g += nwarps) {
        const int r = g_row;
        const int wq = g_col * P;                // first output column inside the tile
        g_row += step_row; g_col += step_col;
        if (g_col >= a.gpr) { g_col -= a.gpr; ++g_row; }
        const int ho = ho0 + r;
        const int wo_first = wo0 + wq;
        if (ho >= a.Ho || wo_first >= a.Wo) continue;   // warp-uniform

        if constexpr (FUSED) {
            // the residual tile is needed only after the K loop: start pulling its lines toward the SM now so the
            // epilogue does not sit on DRAM latency
            if (a.e.res != nullptr) {
                const float* rb = res_n + ho * e_rh;
                if (!CL && e_rw == 1) {       // NCHW: one 32-byte pixel run per channel row
#pragma unroll
                    for (int j = 0; j < C; ++j) {
                        const int c = (blk0 + j) * 32 + lane;
                        if (c < a.Cout) prefetch_l1(rb + c * e_rc + wo_first);
                    }
                } else if (pf_off >= 0 && (LEAN || wo_first + pf_p < a.Wo)) {
                    // channels-last: one 128-byte line per (pixel, 32-channel block); the lane's part of the offset is hoisted
                    prefetch_l1(rb + (wo_first * e_rw + pf_off));
                    if constexpr (CL && KWT == 1) {
                        // 1x1 kernels: the K loop of a group is far shorter than a DRAM round trip, so also request the lines
                        // of this warp's NEXT group now (g_row / g_col already point at it)
                        const int ho2 = ho0 + g_row, wo2 = wo0 + g_col * P;
                        if (g + nwarps < a.G && ho2 < a.Ho && wo2 + pf_p < a.Wo)
                            prefetch_l1(res_n + (ho2 * e_rh + wo2 * e_rw + pf_off));
                    }
                }
            }
        }

        int acc[P][C];
#pragma unroll
        for (int p = 0; p < P; ++p)
#pragma unroll
            for (int j = 0; j < C; ++j) acc[p][j] = 0;

        // row pointers advance by additions: (chunk, kernel row) is a running k-step for the weights
        const uint4* arow_c = act + (r * a.SH) * a.BW + wq * SW;
        const uint2* wrow = wsm + lane - KW * 32;
        if constexpr (KWT == 1 && MODE == 1) {
            // 1x1 kernels (and Linear): the same 3:2 carry-save adder, over three consecutive 64-channel CHUNKS instead of
            // three taps -- 3 words cost 2 POPC.  KH == 1 here (the host selects this instance only for kh == 1).
            const uint2* wr = wsm + lane;
            int ch = 0;
            for (; ch + 3 <= a.nch; ch += 3, arow_c += 3 * a_chstep, wr += 3 * 32) {
                uint2 t[3][C];
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < C; ++j) t[i][j] = wr[(size_t)j * nk32 + i * 32];
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const uint4 v0 = arow_c[p], v1 = arow_c[a_chstep + p], v2 = arow_c[2 * a_chstep + p];
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
            }
            for (; ch < a.nch; ++ch, arow_c += a_chstep, wr += 32) {        // one or two chunks left: one POPC per word
                uint2 t[C];
#pragma unroll
                for (int j = 0; j < C; ++j) t[j] = wr[(size_t)j * nk32];
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const uint4 v = arow_c[p];
#pragma unroll
                    for (int j = 0; j < C; ++j) acc[p][j] += word_dis(v.z, v.x, t[j].x) + word_dis(v.w, v.y, t[j].y);
                }
            }
        } else
        for (int ch = 0; ch < a.nch; ++ch, arow_c += a_chstep) {
            const uint4* arow = arow_c - a_khstep;
            for (int kh = 0; kh < a.KH; ++kh) {
                arow += a_khstep;
                wrow += KW * 32;
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

        // ---------------- epilogue ----------------
        int ms[P];
        if constexpr (P % 4 == 0) {              // tile widths are multiples of P: 16-byte aligned broadcast loads
            const int4* m4 = reinterpret_cast<const int4*>(ms_s + r * a.TW + wq);
#pragma unroll
            for (int q = 0; q < P / 4; ++q) {
                const int4 t = m4[q];
                ms[4 * q] = t.x; ms[4 * q + 1] = t.y; ms[4 * q + 2] = t.z; ms[4 * q + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int p = 0; p < P; ++p) ms[p] = ms_s[r * a.TW + wq + p];      // broadcast loads
        }
        if constexpr (LEAN) {
            const bool has_res = a.e.res != nullptr, has_out = a.e.out != nullptr;
            const int cch = blk0 * 32 + lane;
            const float* rp = res_n + (ho * e_rh + wo_first * e_rw + cch);      // dereferenced only if has_res
            float* op = out_n + (ho * e_oh + wo_first * e_ow + cch);            // dereferenced only if has_out
            float res[C][P];
            if (has_res) {                       // every shortcut line of the group in flight before the first use
                const float* rq = rp;
#pragma unroll
                for (int p = 0; p < P; ++p) {
#pragma unroll
                    for (int j = 0; j < C; ++j) res[j][p] = __ldg(rq + j * 32);
                    rq = reinterpret_cast<const float*>(reinterpret_cast<const char*>(rq) + rwb);
                }
            }
            if constexpr (LEAN_GENERAL) {
                const bool res_after = a.e.res_after_act != 0, nx = a.e.nx_scale != nullptr;
                // general form (pre-activation blocks: PReLU, shortcut after the activation, the next layer's
                // BatchNorm in front of its sign): same operations in the same order as the EPI 1 / 2 epilogue
                const int act = a.e.act;
                const bool want_bits = a.e.obits != nullptr;
                uint4* sb = reinterpret_cast<uint4*>(stg);
                if (want_bits) __syncwarp();
#pragma unroll
                for (int j = 0; j < C; j += 2) {
                    uint32_t sw[2][P], mw[2][P];
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        const int cl = (j + jj) * 32 + lane;
                        const float k0 = epc[cl], k1 = epc[32 * C + cl], k2 = epc[2 * 32 * C + cl];
                        const float k3 = epc[3 * 32 * C + cl], k4 = epc[4 * 32 * C + cl];
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            float v = __fmaf_rn(k0, __fadd_rn(__int_as_float(msm2(ms[p], acc[p][j + jj])), -MAGIC_F), k1);
                            if (has_res && !res_after) v = __fadd_rn(v, res[j + jj][p]);
                            if (act == BNN_ACT_RELU) v = fmaxf(v, 0.0f);
                            else if (act == BNN_ACT_PRELU) v = (v > 0.0f) ? v : __fmul_rn(k2, v);
                            if (has_res && res_after) v = __fadd_rn(v, res[j + jj][p]);
                            if (has_out) op[p * e_ow + (j + jj) * 32] = v;
                            const float b = nx ? __fmaf_rn(k3, v, k4) : v;
                            sw[jj][p] = __ballot_sync(0xffffffffu, b > 0.0f);
                            mw[jj][p] = __ballot_sync(0xffffffffu, b > 0.0f || b < 0.0f);
                        }
                    }
                    if (want_bits && lane == 0) {
#pragma unroll
                        for (int p = 0; p < P; ++p) sb[p * (C / 2) + j / 2] = make_uint4(sw[0][p], sw[1][p], mw[0][p], mw[1][p]);
                    }
                }
                if (want_bits) {
                    __syncwarp();
                    if (lane < P) {
                        const size_t unit0 = (((size_t)n * a.e.ochunks + (blk0 >> 1)) * a.Ho + ho) * a.Wo + wo_first + lane;
                        const size_t ustep = (size_t)a.Ho * a.Wo;
#pragma unroll
                        for (int j = 0; j < C; j += 2) a.e.obits[unit0 + (j / 2) * ustep] = sb[lane * (C / 2) + j / 2];
                    }
                    __syncwarp();
                }
                continue;
            }
            // fast form (EPI 3).  Per output: IMAD (ms + magic - 2 acc), FADD (- magic), FFMA, [FADD shortcut], FMNMX,
            // [STG], FSETP, VOTE -- addresses are running pointers (one 64-bit add per pixel), no conversion instruction
            // Ballots are warp-uniform, so every lane parks the same finished words {s_lo, s_hi | m_lo, m_hi} (m == s: ReLU
            // output) in the warp's staging area -- no divergence -- and lane l later stores unit l (chunk-major, P
            // consecutive units per 64 channels).
            uint2* sb = reinterpret_cast<uint2*>(stg);
            const bool want_bits = a.e.obits != nullptr;
            float k0[C], k1[C];
#pragma unroll
            for (int j = 0; j < C; ++j) { k0[j] = epc[j * 32 + lane]; k1[j] = epc[32 * C + j * 32 + lane]; }
            // the four (shortcut, fp32 output) combinations as straight-line code each
            auto finish = [&](auto HAS_RES, auto HAS_OUT) {
                float* oq = op;
#pragma unroll
                for (int p = 0; p < P; ++p) {
#pragma unroll
                    for (int j = 0; j < C; j += 2) {
                        uint32_t w2[2];
#pragma unroll
                        for (int jj = 0; jj < 2; ++jj) {
                            float v = __fmaf_rn(k0[j + jj], __fadd_rn(__int_as_float(msm2(ms[p], acc[p][j + jj])), -MAGIC_F), k1[j + jj]);
                            if constexpr (HAS_RES.value) v = __fadd_rn(v, res[j + jj][p]);
                            if constexpr (HAS_OUT.value) {
                                v = fmaxf(v, 0.0f);
                                oq[(j + jj) * 32] = v;
                            }
                            w2[jj] = __ballot_sync(0xffffffffu, v > 0.0f);       // max(v, 0) > 0 == v > 0
                        }
                        if (!HAS_OUT.value || want_bits) {
                            sb[((j / 2) * P + p) * 2] = make_uint2(w2[0], w2[1]);
                            sb[((j / 2) * P + p) * 2 + 1] = make_uint2(w2[0], w2[1]);
                        }
                    }
                    if constexpr (HAS_OUT.value) oq = reinterpret_cast<float*>(reinterpret_cast<char*>(oq) + owb);
                }
            };
            using T_ = std::true_type;
            using F_ = std::false_type;
            if (has_res) { if (has_out) finish(T_{}, T_{}); else finish(T_{}, F_{}); }
            else         { if (has_out) finish(F_{}, T_{}); else finish(F_{}, F_{}); }
            if (want_bits) {
                __syncwarp();
                if (lane < P * (C / 2))
                    a.e.obits[obits_base + (size_t)(ho * a.Wo + wo_first)] = reinterpret_cast<const uint4*>(stg)[lane];
                __syncwarp();
            }
            continue;
        }
        const bool transposed = CL ? false : ((a.e.ow == 1) || (a.e.out == nullptr));
        const bool has_res = FUSED && a.e.res != nullptr;
        const bool want_bits = FUSED && a.e.obits != nullptr;
        // ReLU output with no affine in front of the next sign(): "non-zero" and "positive" coincide
        const bool bits_pre = has_res && a.e.res_after_act && a.e.bits_pre_res;
        const bool relu_bits = a.e.nx_relu || (a.e.act == BNN_ACT_RELU && a.e.nx_scale == nullptr &&
                                               !(has_res && a.e.res_after_act && !bits_pre));
        const bool full = (wo_first + P <= a.Wo) && ((blk0 + C) * 32 <= a.Cout);
        uint32_t sbits[C], mbits[C];     // lane p keeps the packed words of pixel p
#pragma unroll
        for (int j = 0; j < C; ++j) { sbits[j] = 0u; mbits[j] = 0u; }

#pragma unroll
        for (int j = 0; j < C; ++j) {
            const int cl = j * 32 + lane;                // channel inside the CTA tile
            const int cblk = (blk0 + j) * 32;
            const bool c_ok = cblk + lane < a.Cout;
            const float k0 = epc[0 * 32 * C + cl], k1 = epc[1 * 32 * C + cl], k2 = epc[2 * 32 * C + cl];
            float v[P];
            if constexpr (EPI == 0) {
                // reference order (conv.py:92-97, ops.py:136,202): (alpha*dot + bias) * alpha_post
#pragma unroll
                for (int p = 0; p < P; ++p)
                    v[p] = __fmul_rn(__fadd_rn(__fmul_rn(k0, __fadd_rn(__int_as_float(msm2(ms[p], acc[p][j])), -MAGIC_F)), k1), k2);
            } else {
                // ---- fused epilogue.  `full` groups (all P pixels and all 32 channels valid) take the
                //      predicate-free path; strides are 32-bit here (the host checked the tensors fit)
                float res[P];
                const bool res_direct = CL ? has_res : (has_res && a.e.rw != 1);
                if (has_res && !res_direct) {
                    // NCHW residual: tile [32 ch][P px] through shared memory, coalesced along pixels
                    const float* rbase = res_n + ho * e_rh;
                    __syncwarp();
#pragma unroll
                    for (int r0 = 0; r0 < 32; r0 += ROWS) {
                        const int rl = r0 + rr, c = cblk + rl, wo = wo_first + pr;
                        float t = 0.0f;
                        if (pr < P && wo < a.Wo && c < a.Cout) t = __ldg(rbase + c * e_rc + wo);
                        if (pr < P) stg[rl * PITCH + pr] = t;
                    }
                    __syncwarp();
#pragma unroll
                    for (int p = 0; p < P; ++p) res[p] = stg[lane * PITCH + p];
                } else if (res_direct) {
                    // channel-contiguous residual (NHWC): lanes <-> channels reads whole 128-byte lines
                    const float* rp = res_n + (ho * e_rh + wo_first * e_rw + (cblk + lane) * e_rc);
                    const int rw = e_rw;
                    if (full) {
#pragma unroll
                        for (int p = 0; p < P; ++p) res[p] = __ldg(rp + p * rw);
                    } else {
#pragma unroll
                        for (int p = 0; p < P; ++p) res[p] = (c_ok && wo_first + p < a.Wo) ? __ldg(rp + p * rw) : 0.0f;
                    }
                }
#pragma unroll
                for (int p = 0; p < P; ++p) v[p] = __fmaf_rn(k0, __fadd_rn(__int_as_float(msm2(ms[p], acc[p][j])), -MAGIC_F), k1);
                if (has_res && !a.e.res_after_act) {
#pragma unroll
                    for (int p = 0; p < P; ++p) v[p] = __fadd_rn(v[p], res[p]);
                }
                if (a.e.act == BNN_ACT_RELU) {
#pragma unroll
                    for (int p = 0; p < P; ++p) v[p] = fmaxf(v[p], 0.0f);
                } else if (a.e.act == BNN_ACT_PRELU) {
#pragma unroll
                    for (int p = 0; p < P; ++p) v[p] = (v[p] > 0.0f) ? v[p] : __fmul_rn(k2, v[p]);
                }
                if (has_res && a.e.res_after_act && !bits_pre) {
#pragma unroll
                    for (int p = 0; p < P; ++p) v[p] = __fadd_rn(v[p], res[p]);
                }
                if (want_bits) {
                    float b[P];
                    if (a.e.nx_scale) {
                        const float k3 = epc[3 * 32 * C + cl], k4 = epc[4 * 32 * C + cl];
#pragma unroll
                        for (int p = 0; p < P; ++p) b[p] = __fmaf_rn(k3, v[p], k4);
                    } else {
#pragma unroll
                        for (int p = 0; p < P; ++p) b[p] = v[p];
                    }
                    if (relu_bits) {
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            const uint32_t sw = __ballot_sync(0xffffffffu, c_ok && b[p] > 0.0f);
                            if (lane == p) { sbits[j] = sw; mbits[j] = sw; }
                        }
                    } else {
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            const uint32_t sw = __ballot_sync(0xffffffffu, c_ok && b[p] > 0.0f);
                            const uint32_t mw = __ballot_sync(0xffffffffu, c_ok && (b[p] > 0.0f || b[p] < 0.0f));
                            if (lane == p) { sbits[j] = sw; mbits[j] = mw; }
                        }
                    }
                }
                if (bits_pre) {            // the planes were taken before the shortcut is added
#pragma unroll
                    for (int p = 0; p < P; ++p) v[p] = __fadd_rn(v[p], res[p]);
                }
            }
            if (a.e.out != nullptr) {
                float* obase = out_n + ho * e_oh;
                if (!transposed) {
                    // channel-contiguous output (Linear's [rows, out]): lanes <-> channels is already coalesced
                    float* op = obase + (wo_first * e_ow + (cblk + lane) * e_oc);
                    const int ow = e_ow;
                    if (full) {
#pragma unroll
                        for (int p = 0; p < P; ++p) op[p * ow] = v[p];
                    } else {
#pragma unroll
                        for (int p = 0; p < P; ++p)
                            if (c_ok && wo_first + p < a.Wo) op[p * ow] = v[p];
                    }
                } else {
                    // pixel-contiguous output (NCHW): transpose through shared memory so one store instruction
                    // writes whole 32-byte pixel runs instead of 32 scattered words
                    __syncwarp();
#pragma unroll
                    for (int p = 0; p < P; ++p) stg[lane * PITCH + p] = v[p];
                    __syncwarp();
                    const bool lane_ok = pr < P && wo_first + pr < a.Wo;
                    float* optr = obase + ((cblk + rr) * e_oc + wo_first + pr);
                    const int ostep = ROWS * e_oc;
#pragma unroll
                    for (int r0 = 0; r0 < 32; r0 += ROWS) {
                        if (lane_ok && cblk + r0 + rr < a.Cout) *optr = stg[(r0 + rr) * PITCH + pr];
                        optr += ostep;
                    }
                }
            }
        }
        if (want_bits && lane < P && wo_first + lane < a.Wo) {
            // lane p writes the 16-byte units of pixel p: consecutive lanes -> consecutive units (coalesced)
            uint32_t* ob = reinterpret_cast<uint32_t*>(a.e.obits);
#pragma unroll
            for (int j = 0; j < C; ++j) {
                const int blk = blk0 + j;
                if (blk >= a.nblk32) break;
                const size_t unit_idx = (((size_t)n * a.e.ochunks + (blk >> 1)) * a.Ho + ho) * a.Wo + wo_first + lane;
                if (C >= 2) {
                    if ((j & 1) == 0) {       // blk0 is even when C >= 2: (j, j+1) form one 64-channel unit
                        constexpr int JH = (C >= 2) ? 1 : 0;
                        reinterpret_cast<uint4*>(ob)[unit_idx] = make_uint4(sbits[j], sbits[j + JH], mbits[j], mbits[j + JH]);
                    }
                } else {
                    ob[unit_idx * 4 + (blk & 1)] = sbits[j];
                    ob[unit_idx * 4 + 2 + (blk & 1)] = mbits[j];
                    if ((blk & 1) == 0 && blk + 1 >= a.nblk32) {   // no odd partner: its half of the unit is zero
                        ob[unit_idx * 4 + 1] = 0u;
                        ob[unit_idx * 4 + 3] = 0u;
                    }
                }
            }
        }
    }
}

struct Plan {
    int P, C, kwt, swt, mode;
    int TH, TW, BH, BW, NW, gpr, G, tiles_h, tiles_w;
    size_t smem;
};

typedef void (*KernelFn)(const CUtensorMap, const ConvArgs);

template <int KWT, int SWT, int MODE, int EPI>
inline KernelFn pick_pc(int P, int C) {
#define BNN_PC(p, c) if (P == p && C == c) return bconv_kernel<p, c, KWT, SWT, MODE, EPI>;
    BNN_PC(8, 4) BNN_PC(8, 2)
    BNN_PC(7, 4) BNN_PC(7, 2)
    BNN_PC(4, 4) BNN_PC(4, 2)
    if constexpr (EPI < 3) {           // the lean epilogues write whole 64-channel units: C >= 2 only
        BNN_PC(8, 1) BNN_PC(7, 1) BNN_PC(4, 1)
    }
#undef BNN_PC
    return nullptr;
}


// one translation unit per epilogue kind (bconv_inst.cu, compiled with -DBNN_EPI=0..4): the instances build in parallel
KernelFn pick_kernel_epi0(const Plan& p);
KernelFn pick_kernel_epi1(const Plan& p);
KernelFn pick_kernel_epi2(const Plan& p);
KernelFn pick_kernel_epi3(const Plan& p);
KernelFn pick_kernel_epi4(const Plan& p);

}  // namespace bnn
