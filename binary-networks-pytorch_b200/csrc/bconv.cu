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

#include <algorithm>
#include <cudaTypedefs.h>
#include <map>
#include <mutex>
#include <vector>

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
};

// per-CTA table of per-channel epilogue constants in shared memory: [EP_N][32*C]
//   EPI == 0 (reference epilogue, exact order):  y = (k0 * dot + k1) * k2         k = scale, bias, post
//   EPI == 1 (cross-module fusion):              z = fma(k0, dot, k1)            k0/k1 fold scale, bias, post, BN
//                                                k2 = PReLU slope, k3/k4 = next layer's pre-sign affine
enum { EP_N = 5 };

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__device__ __forceinline__ int word_dis(uint32_t m, uint32_t s, uint32_t t) { return __popc(m & (s ^ t)); }
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (c & (a ^ b)); }

template <int P, int C, int KWT, int SWT, int MODE, int EPI>
__global__ void __launch_bounds__(256, 2)
bconv_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ ConvArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    // EPI 0: reference epilogue.  EPI 1: fused epilogue, any strides.  EPI 2: fused epilogue with channel-contiguous
    // (NHWC) residual and output -- what the fused engine always uses: the pixel-contiguous transposes are compiled out
    // EPI 3: the lean NHWC epilogue of residual blocks, launched only when every pixel group and every channel block is
    // complete -- no bounds predicates, no stride arithmetic beyond one multiply per access, planes through the warp's
    // staging area.  Fast form: BatchNorm, optional shortcut add BEFORE a ReLU, planes where "non-zero" == "positive";
    // general form: any activation, shortcut before or after it, optional affine in front of the next sign().
    constexpr bool FUSED = EPI >= 1, CL = EPI >= 2, LEAN = EPI == 3;
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
        ms_s[i] = cnt;
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
    const int a_chstep = a.BH * a.BW, a_khstep = a.DH * a.BW;     // activation rows: per chunk, per kernel row
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
                } else if (lane < P * C) {    // channels-last: one 128-byte line per (pixel, 32-channel block)
                    const int j = lane / P, p = lane - j * P;
                    if ((blk0 + j) * 32 < a.Cout && wo_first + p < a.Wo)
                        prefetch_l1(rb + (blk0 + j) * 32 * e_rc + (wo_first + p) * e_rw);
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
#pragma unroll
        for (int p = 0; p < P; ++p) ms[p] = ms_s[r * a.TW + wq + p];      // broadcast loads
        if constexpr (LEAN) {
            const bool has_res = a.e.res != nullptr, has_out = a.e.out != nullptr;
            const int cch = blk0 * 32 + lane;
            const float* rp = res_n + (ho * e_rh + wo_first * e_rw + cch);      // dereferenced only if has_res
            float* op = out_n + (ho * e_oh + wo_first * e_ow + cch);            // dereferenced only if has_out
            float res[C][P];
            if (has_res) {                       // every shortcut line of the group in flight before the first use
#pragma unroll
                for (int j = 0; j < C; ++j)
#pragma unroll
                    for (int p = 0; p < P; ++p) res[j][p] = __ldg(rp + p * e_rw + j * 32);
            }
            const bool res_after = a.e.res_after_act != 0, nx = a.e.nx_scale != nullptr;
            if (a.e.act != BNN_ACT_RELU || nx || (has_res && res_after)) {
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
                            float v = __fmaf_rn(k0, (float)(ms[p] - 2 * acc[p][j + jj]), k1);
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
            uint32_t sw[P][C];                   // ballots are warp-uniform: every lane holds every word
#pragma unroll
            for (int j = 0; j < C; ++j) {
                const float k0 = epc[j * 32 + lane], k1 = epc[32 * C + j * 32 + lane];
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    float v = __fmaf_rn(k0, (float)(ms[p] - 2 * acc[p][j]), k1);
                    if (has_res) v = __fadd_rn(v, res[j][p]);
                    v = fmaxf(v, 0.0f);
                    if (has_out) op[p * e_ow + j * 32] = v;
                    sw[p][j] = __ballot_sync(0xffffffffu, v > 0.0f);
                }
            }
            if (a.e.obits != nullptr) {
                // lane 0 parks finished 16-byte units {s_lo, s_hi, m_lo, m_hi} (m == s: ReLU output) in the warp's
                // staging area; lane p then stores pixel p's units: consecutive lanes -> consecutive units
                uint4* sb = reinterpret_cast<uint4*>(stg);
                __syncwarp();
                if (lane == 0) {
#pragma unroll
                    for (int p = 0; p < P; ++p)
#pragma unroll
                        for (int j = 0; j < C; j += 2) sb[p * (C / 2) + j / 2] = make_uint4(sw[p][j], sw[p][j + 1], sw[p][j], sw[p][j + 1]);
                }
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
                    v[p] = __fmul_rn(__fadd_rn(__fmul_rn(k0, (float)(ms[p] - 2 * acc[p][j])), k1), k2);
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
                for (int p = 0; p < P; ++p) v[p] = __fmaf_rn(k0, (float)(ms[p] - 2 * acc[p][j]), k1);
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

// ---------------------------------------------------------------------------
// host side: tensor map, planner, dispatch
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
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

template <int KWT, int SWT, int MODE, int EPI>
static KernelFn pick_pc(int P, int C) {
#define BNN_PC(p, c) if (P == p && C == c) return bconv_kernel<p, c, KWT, SWT, MODE, EPI>;
    BNN_PC(8, 4) BNN_PC(8, 2)
    BNN_PC(7, 4) BNN_PC(7, 2)
    BNN_PC(4, 4) BNN_PC(4, 2)
    if constexpr (EPI != 3) {          // the lean epilogue writes whole 64-channel units: C >= 2 only
        BNN_PC(8, 1) BNN_PC(7, 1) BNN_PC(4, 1)
    }
#undef BNN_PC
    return nullptr;
}

// EPI 0 = reference epilogue (both inner-loop modes, for A/B runs); EPI 1 / 2 = fused epilogue (carry-save only)
static KernelFn pick_kernel(const Plan& p, int epi) {
    if (epi == 0) {
        if (p.kwt == 3 && p.swt == 1) return p.mode ? pick_pc<3, 1, 1, 0>(p.P, p.C) : pick_pc<3, 1, 0, 0>(p.P, p.C);
        if (p.kwt == 3 && p.swt == 2) return p.mode ? pick_pc<3, 2, 1, 0>(p.P, p.C) : pick_pc<3, 2, 0, 0>(p.P, p.C);
        if (p.kwt == 1 && p.swt == 1) return pick_pc<1, 1, 0, 0>(p.P, p.C);
        return pick_pc<0, 0, 0, 0>(p.P, p.C);
    }
    if (epi == 3) {
        if (p.kwt == 3 && p.swt == 1) return pick_pc<3, 1, 1, 3>(p.P, p.C);
        if (p.kwt == 3 && p.swt == 2) return pick_pc<3, 2, 1, 3>(p.P, p.C);
        if (p.kwt == 1 && p.swt == 1) return pick_pc<1, 1, 0, 3>(p.P, p.C);
        return pick_pc<0, 0, 0, 3>(p.P, p.C);
    }
    if (epi == 2) {
        if (p.kwt == 3 && p.swt == 1) return pick_pc<3, 1, 1, 2>(p.P, p.C);
        if (p.kwt == 3 && p.swt == 2) return pick_pc<3, 2, 1, 2>(p.P, p.C);
        if (p.kwt == 1 && p.swt == 1) return pick_pc<1, 1, 0, 2>(p.P, p.C);
        return pick_pc<0, 0, 0, 2>(p.P, p.C);
    }
    if (p.kwt == 3 && p.swt == 1) return pick_pc<3, 1, 1, 1>(p.P, p.C);
    if (p.kwt == 3 && p.swt == 2) return pick_pc<3, 2, 1, 1>(p.P, p.C);
    if (p.kwt == 1 && p.swt == 1) return pick_pc<1, 1, 0, 1>(p.P, p.C);
    return pick_pc<0, 0, 0, 1>(p.P, p.C);
}

// 0 = reference epilogue, 1 = fused, 2 = fused with channel-contiguous residual / output (or neither)
static int epilogue_kind(const bnn_epilogue& ep) {
    const bool fused = ep.bn_scale || ep.residual || ep.act != BNN_ACT_NONE || ep.out_bits || ep.nx_scale;
    if (!fused) return 0;
    const bool cl = (!ep.out || ep.ostride_c == 1) && (!ep.residual || ep.rstride_c == 1);
    if (!cl) return 1;
    // 3: residual blocks of either flavour (post-activation: BatchNorm folded, shortcut before the ReLU; pre-activation:
    // PReLU, shortcut after it, the next BatchNorm in front of the sign) -- everything but the Hierarchical-Block forms;
    // whether the lean instance can actually run also depends on the tile plan, see lean_plan_ok()
    const bool lean = !ep.nx_relu && !ep.bits_before_residual;
    return lean ? 3 : 2;
}

// the lean epilogue has no bounds predicates: every pixel group and every channel block must be complete
static bool lean_plan_ok(const Plan& pl, const bnn_conv_geom& g, int Wo) {
    return pl.C >= 2 && Wo % pl.P == 0 && g.c_out % (32 * pl.C) == 0;
}

static int ceil_div(int a, int b) { return (a + b - 1) / b; }

static size_t plan_smem(int nch, int BH, int BW, int C, int nk, int NW, int P, int TH, int TW) {
    const size_t act_bytes = (size_t)nch * BH * BW * 16;
    return 128 + ((act_bytes + 127) & ~(size_t)127) + (size_t)C * nk * 256 + (size_t)NW * 32 * (P | 1) * 4 +
           (size_t)EP_N * 32 * C * 4 + (size_t)TH * TW * 4;
}

// Candidate tile shapes for one layer, ranked by a small cost model (host only, microseconds):
//   time ~ waves x (rounds x work_per_round x warps_on_fullest_scheduler + staging),
//   waves = ceil(CTAs / (SMs x CTAs_per_SM)).
// It captures the measured loss terms: idle lanes of partial pixel groups / idle warps in the last round,
// the tail of the last wave, uneven warps per scheduler, and the staging latency of very small CTAs.
// The model only ranks; bnn_conv_tune times the best few on the device and caches the winner.
struct Cand { Plan pl; double cost; };

static int enumerate_plans(const bnn_conv_geom& g, int Ho, int Wo, uint32_t flags, int sms, std::vector<Cand>& out) {
    const int nch = ceil_div(g.c_in, 64), nk = nch * g.kh * g.kw;
    if (nch > 256) return BNN_E_UNSUPPORTED;
    int kwt = 0, swt = 0;            // unrolled instances: kernel width x horizontal stride, dilation_w == 1
    if (g.dil_w == 1) {
        if (g.kw == 3 && (g.stride_w == 1 || g.stride_w == 2)) { kwt = 3; swt = g.stride_w; }
        else if (g.kw == 1 && g.stride_w == 1) { kwt = 1; swt = 1; }
    }
    const int mode = (kwt == 3 && !(flags & BNN_F_NO_CSA)) ? 1 : 0;
    const int nblk32 = ceil_div(g.c_out, 32);
    const size_t smem_cap = 220 * 1024;
    const int candP[3] = {8, 7, 4}, candC[3] = {4, 2, 1};
    for (int ci = 0; ci < 3; ++ci) {
        const int C = candC[ci];
        if (C > 1 && C / 2 >= nblk32) continue;                  // would only add idle channel blocks
        const size_t wbytes = (size_t)C * nk * 256;
        if (wbytes + 16 * 1024 > smem_cap) continue;
        const int cout_tiles = ceil_div(nblk32, C);
        for (int pi = 0; pi < 3; ++pi) {
            const int P = candP[pi];
            const int usw = kwt ? swt : 1, ukw = kwt ? kwt : 1;
            const bool window = ((P - 1) * usw + ukw) <= 12;
            int TW = ceil_div(Wo, P) * P;
            while ((TW - 1) * g.stride_w + (g.kw - 1) * g.dil_w + 1 > 256 && TW > P) TW -= P;
            const int BW = (TW - 1) * g.stride_w + (g.kw - 1) * g.dil_w + 1;
            if (BW > 256) continue;
            const int gpr = TW / P, tiles_w = ceil_div(Wo, TW);
            // cycles per 32-bit word per warp: fewer channels per lane -> more shared-memory loads per word
            double cpw = (C == 4 ? 1.0 : C == 2 ? 1.25 : 1.5) * (window ? 1.0 : 1.12) * (P == 4 ? 1.06 : 1.0);
            const double round_work = (double)P * C * nk * 2.0 * cpw + 60.0 * P * C;   // main loop + epilogue
            for (int NW = 8; NW >= 7; --NW) {
                for (int TH = 1; TH <= Ho; ++TH) {
                    const int BH = (TH - 1) * g.stride_h + (g.kh - 1) * g.dil_h + 1;
                    if (BH > 256) break;
                    const size_t smem = plan_smem(nch, BH, BW, C, nk, NW, P, TH, TW);
                    if (smem > smem_cap) break;
                    const int G = TH * gpr, rounds = ceil_div(G, NW);
                    if (rounds > 12 && TH > 1) break;
                    const int occ = smem * 2 <= 226 * 1024 ? 2 : 1;
                    const long long ctas = (long long)g.n * ceil_div(Ho, TH) * tiles_w * cout_tiles;
                    const long long slots = (long long)sms * occ;
                    const double waves = (double)((ctas + slots - 1) / slots);
                    const double share = (double)((NW * occ + 3) / 4);       // warps on the fullest scheduler
                    const double stage = 1500.0 + 0.02 * (double)(smem);
                    Cand c{};
                    c.cost = waves * (rounds * round_work * share + stage);
                    c.pl.P = P; c.pl.C = C; c.pl.kwt = kwt; c.pl.swt = swt; c.pl.mode = mode;
                    c.pl.TH = TH; c.pl.TW = TW; c.pl.BH = BH; c.pl.BW = BW; c.pl.NW = NW; c.pl.gpr = gpr; c.pl.G = G;
                    c.pl.tiles_h = ceil_div(Ho, TH); c.pl.tiles_w = tiles_w; c.pl.smem = smem;
                    out.push_back(c);
                }
            }
        }
    }
    if (out.empty()) return BNN_E_UNSUPPORTED;
    std::sort(out.begin(), out.end(), [](const Cand& a, const Cand& b) { return a.cost < b.cost; });
    return 0;
}

// tuned plans, keyed by geometry + epilogue kind + flags that change the kernel
struct PlanKey {
    int v[16];
    bool operator<(const PlanKey& o) const { return memcmp(v, o.v, sizeof(v)) < 0; }
};
static std::map<PlanKey, Plan> g_tuned;
static std::mutex g_tuned_mu;

static PlanKey plan_key(const bnn_conv_geom& g, int epi, uint32_t flags) {
    PlanKey k{};
    const int vals[16] = {g.n, g.c_in, g.h, g.w, g.c_out, g.kh, g.kw, g.stride_h, g.stride_w, g.pad_h, g.pad_w,
                          g.dil_h, g.dil_w, epi, (int)(flags & BNN_F_NO_CSA), 0};
    memcpy(k.v, vals, sizeof(vals));
    return k;
}

static int make_plan(const bnn_conv_geom& g, int Ho, int Wo, uint32_t flags, int sms, Plan* out, int epi = 0) {
    {
        std::lock_guard<std::mutex> lock(g_tuned_mu);
        auto it = g_tuned.find(plan_key(g, epi, flags));
        if (it != g_tuned.end()) { *out = it->second; return 0; }
    }
    std::vector<Cand> cands;
    int rc = enumerate_plans(g, Ho, Wo, flags, sms, cands);
    if (rc) return rc;
    *out = cands[0].pl;
    return 0;
}

static int launch_bconv(const void* abits, const void* wbits, const bnn_conv_geom& g, const bnn_epilogue& ep,
                        uint32_t flags, cudaStream_t stream, const Plan* forced = nullptr) {
    if (!abits || !wbits || (!ep.out && !ep.out_bits)) return BNN_E_NULL;
    if (g.n <= 0 || g.c_in <= 0 || g.h <= 0 || g.w <= 0 || g.c_out <= 0 || g.kh <= 0 || g.kw <= 0 ||
        g.stride_h <= 0 || g.stride_w <= 0 || g.pad_h < 0 || g.pad_w < 0 || g.dil_h <= 0 || g.dil_w <= 0)
        return BNN_E_SHAPE;
    if (ep.act < BNN_ACT_NONE || ep.act > BNN_ACT_PRELU || (ep.act == BNN_ACT_PRELU && !ep.act_slope)) return BNN_E_SHAPE;
    if ((ep.bn_scale == nullptr) != (ep.bn_shift == nullptr) || (ep.nx_scale == nullptr) != (ep.nx_shift == nullptr))
        return BNN_E_NULL;
    if (((uintptr_t)abits & 15) || ((uintptr_t)wbits & 15) || ((uintptr_t)ep.out_bits & 15)) return BNN_E_ALIGN;
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

    const int epi = epilogue_kind(ep);
    {   // the kernel indexes inside one image row with 32-bit element offsets
        const long long lim = 0x7fffffffLL;
        auto ab = [](int64_t v) { return (long long)(v < 0 ? -v : v); };
        auto span = [&](int64_t sc_, int64_t sh_, int64_t sw_) {
            return (long long)g.c_out * ab(sc_) + (long long)Ho * ab(sh_) + (long long)Wo * ab(sw_);
        };
        if ((ep.out && span(ep.ostride_c, ep.ostride_h, ep.ostride_w) > lim) ||
            (ep.residual && span(ep.rstride_c, ep.rstride_h, ep.rstride_w) > lim))
            return BNN_E_UNSUPPORTED;
    }
    if (epi) flags &= ~BNN_F_NO_CSA;
    Plan pl;
    int rc = 0;
    if (forced) pl = *forced;
    else rc = make_plan(g, Ho, Wo, flags, sms, &pl, epi);
    if (rc) return rc;
    KernelFn fn = pick_kernel(pl, (epi == 3 && !lean_plan_ok(pl, g, Wo)) ? 2 : epi);
    if (!fn) return BNN_E_UNSUPPORTED;

    ConvArgs a{};
    a.abits = (const uint4*)abits; a.wbits = (const uint2*)wbits;
    a.e.scale = ep.scale; a.e.bias = ep.bias; a.e.post = ep.post;
    a.e.bn_scale = ep.bn_scale; a.e.bn_shift = ep.bn_shift; a.e.slope = ep.act_slope;
    a.e.nx_scale = ep.nx_scale; a.e.nx_shift = ep.nx_shift;
    a.e.res = ep.residual; a.e.rn = ep.rstride_n; a.e.rc = ep.rstride_c; a.e.rh = ep.rstride_h; a.e.rw = ep.rstride_w;
    a.e.out = ep.out; a.e.on = ep.ostride_n; a.e.oc = ep.ostride_c; a.e.oh = ep.ostride_h; a.e.ow = ep.ostride_w;
    a.e.obits = (uint4*)ep.out_bits; a.e.act = ep.act; a.e.res_after_act = ep.residual_after_act;
    a.e.nx_relu = ep.nx_relu; a.e.bits_pre_res = ep.bits_before_residual;
    a.e.ochunks = ceil_div(g.c_out, 64);
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

extern "C" int bnn_bconv2d_tune(const void* abits, const void* wbits, const bnn_conv_geom* geom,
                                const bnn_epilogue* epilogue, uint32_t flags, int32_t top_k, void* stream_) {
    if (!geom || !epilogue) return BNN_E_NULL;
    const bnn_conv_geom& g = *geom;
    const bnn_epilogue& ep = *epilogue;
    const int Ho = out_dim(g.h, g.kh, g.stride_h, g.pad_h, g.dil_h);
    const int Wo = out_dim(g.w, g.kw, g.stride_w, g.pad_w, g.dil_w);
    if (Ho <= 0 || Wo <= 0) return BNN_E_SHAPE;
    const int epi = epilogue_kind(ep);
    if (epi) flags &= ~BNN_F_NO_CSA;
    const PlanKey key = plan_key(g, epi, flags);
    {
        std::lock_guard<std::mutex> lock(g_tuned_mu);
        if (g_tuned.count(key)) return 0;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    std::vector<Cand> cands;
    int rc = enumerate_plans(g, Ho, Wo, flags, sms, cands);
    if (rc) return rc;
    // keep the model's best few, but make sure different (P, C, warps) families are represented
    std::vector<Plan> tries;
    for (const Cand& c : cands) {
        int same = 0;
        for (const Plan& t : tries) same += (t.P == c.pl.P && t.C == c.pl.C && t.NW == c.pl.NW);
        if (same < 2) tries.push_back(c.pl);
        if ((int)tries.size() >= (top_k > 0 ? top_k : 8)) break;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    Plan best_pl = tries[0];
    for (const Plan& pl : tries) {
        rc = launch_bconv(abits, wbits, g, ep, flags, stream, &pl);            // warm-up
        if (rc) continue;
        float ms_min = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0, stream);
            launch_bconv(abits, wbits, g, ep, flags, stream, &pl);
            cudaEventRecord(e1, stream);
            if (cudaEventSynchronize(e1) != cudaSuccess) { ms_min = 1e30f; break; }
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            ms_min = ms < ms_min ? ms : ms_min;
        }
        if (ms_min < best) { best = ms_min; best_pl = pl; }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) return (int)ce;
    std::lock_guard<std::mutex> lock(g_tuned_mu);
    g_tuned[key] = best_pl;
    return 0;
}

extern "C" int bnn_conv_plan(const bnn_conv_geom* g, uint32_t flags, int32_t sms, int32_t* plan) {
    if (!g || !plan) return BNN_E_NULL;
    const int Ho = out_dim(g->h, g->kh, g->stride_h, g->pad_h, g->dil_h);
    const int Wo = out_dim(g->w, g->kw, g->stride_w, g->pad_w, g->dil_w);
    if (Ho <= 0 || Wo <= 0) return BNN_E_SHAPE;
    Plan pl;
    int rc = make_plan(*g, Ho, Wo, flags, sms > 0 ? sms : 148, &pl);
    if (rc) return rc;
    const int v[12] = {pl.P, pl.C, pl.kwt, pl.swt, pl.mode, pl.TH, pl.TW, pl.NW, pl.tiles_h * pl.tiles_w * g->n,
                       ceil_div(ceil_div(g->c_out, 32), pl.C), (int)pl.smem, pl.G};
    for (int i = 0; i < 12; ++i) plan[i] = v[i];
    return 0;
}

extern "C" int bnn_bconv2d_fused_fwd(const void* abits, const void* wbits, const bnn_conv_geom* geom,
                                     const bnn_epilogue* epilogue, uint32_t flags, void* stream) {
    if (!geom || !epilogue) return BNN_E_NULL;
    return launch_bconv(abits, wbits, *geom, *epilogue, flags, (cudaStream_t)stream);
}

extern "C" int bnn_bconv2d_fwd(const void* abits, const void* wbits, const float* scale, const float* bias,
                               const float* post, float* out, int64_t on, int64_t oc, int64_t oh, int64_t ow,
                               const bnn_conv_geom* geom, uint32_t flags, void* stream) {
    if (!geom) return BNN_E_NULL;
    if (!out) return BNN_E_NULL;
    bnn_epilogue ep{};
    ep.scale = scale; ep.bias = bias; ep.post = post;
    ep.out = out; ep.ostride_n = on; ep.ostride_c = oc; ep.ostride_h = oh; ep.ostride_w = ow;
    return launch_bconv(abits, wbits, *geom, ep, flags, (cudaStream_t)stream);
}

extern "C" int bnn_blinear_fwd(const void* abits, const void* wbits, const float* scale, const float* bias,
                               const float* post, float* out, int32_t rows, int32_t in_features,
                               int32_t out_features, uint32_t flags, void* stream) {
    if (!out) return BNN_E_NULL;
    // rows play the role of the image width: x is [1, in, 1, rows], out is stored [rows, out]
    bnn_conv_geom g{1, in_features, 1, rows, out_features, 1, 1, 1, 1, 0, 0, 1, 1};
    bnn_epilogue ep{};
    ep.scale = scale; ep.bias = bias; ep.post = post;
    ep.out = out; ep.ostride_n = 0; ep.ostride_c = 1; ep.ostride_h = 0; ep.ostride_w = (int64_t)out_features;
    return launch_bconv(abits, wbits, g, ep, flags, (cudaStream_t)stream);
}
