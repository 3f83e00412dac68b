// stem_mma.cu -- the fp32 stem of bnn.models.resnet on the legacy tensor path (mma.sync, SASS HMMA):
//     conv 7x7 / stride 2 / pad 3 (3 -> 64)  ->  eval BatchNorm  ->  ReLU  ->  MaxPool 3x3 / 2 / pad 1
// (reference bnn/models/resnet.py:85-92,147-153).  Same contract as stem.cu (NHWC fp32 + the first
// binarized conv's planes); this kernel trades the fp32 FMA pipe (37 T fma/s measured, bnn_ubench 8)
// for mma.sync.m16n8k16 f16 (278 T fma/s measured, bnn_ubench 6) WITHOUT giving up fp32 accuracy:
//
//   x*2^sx = xh + xl,  w*2^sw = wh + wl     (xh, wh = fp16 round-to-nearest; xl, wl = fp16 of the exact
//                                            fp32 remainder: 22 significand bits per operand)
//   x*w ~= xh*wh + xh*wl + xl*wh            (every product of two fp16 values is exact in fp32; the dropped
//                                            xl*wl term is <= 2^-24 relative)
//
// Per 16-wide k step and accumulator tile the three MMAs run as one short chain from a ZERO accumulator
// (small terms first) and the result is added to the running sum with a round-to-nearest FADD, so the tensor
// core's truncating accumulate only ever sees a 16-term partial sum: the total error stays at the level of an
// fp32 fma chain over the 147 taps (tests/test_gpu_fused.py measures both against a float64 convolution).
//
// Implicit GEMM: M = conv pixels of the CTA's tile (17 x 15 = 255 -> 16 m16 tiles, 2 per warp), N = 64
// channels (8 n8 tiles), K = 21 kernel rows (c_in, kh) x 8 (kw padded with one zero tap) = 11 k16 steps of two
// kernel rows.  With kw padded to 8 the two fp16 values a thread needs for an A fragment register are one aligned
// 32-bit word of the staged input window, at word address  row(c_in, 2r + kh) + c + t  -- consecutive across the
// warp, so the im2col gather is conflict-free LDS.32 and needs no materialised A tile.  B fragments (hi and lo,
// 16 B per lane per (k step, n tile)) are pre-ordered by bnn_stem_mma_pack_weight and arrive by one TMA bulk copy.
// CTA = 8 x 7 pooled pixels (56 = 7 * 8: no ragged tiles at 224 x 224), 8 warps, two CTAs per SM; the BN+ReLU'd
// conv tile aliases the operand buffers in shared memory for the 3x3 max.
#include "common.cuh"

#include <cuda_fp16.h>

namespace bnn {

constexpr int SM_PH = 8, SM_PW = 7;                            // pooled tile
constexpr int SM_CR = 2 * SM_PH + 1, SM_CC = 2 * SM_PW + 1;    // conv tile 17 x 15
constexpr int SM_NPIX = SM_CR * SM_CC;                         // 255
constexpr int SM_IR = 2 * SM_CR + 5, SM_IC = 2 * SM_CC + 5;    // input window 39 x 35
constexpr int SM_IPW = 18;                                     // window row pitch in 32-bit words (36 halves)
constexpr int SM_WARPS = 8, SM_MT = 2;                         // m16 tiles per warp
constexpr int SM_KSTEPS = 11, SM_NT = 8;
constexpr int SM_CPITCH = 72;                                  // conv tile pixel pitch (floats): conflict-free float2 stores
constexpr int SM_IN_WORDS = 3 * SM_IR * SM_IPW;                // 2106
constexpr int SM_IN_WORDS_PAD = (SM_IN_WORDS + 3) & ~3;        // 2108
constexpr size_t SM_W_BYTES = (size_t)SM_KSTEPS * SM_NT * 32 * 16;          // 45056
constexpr size_t SM_CONV_BYTES = (size_t)(SM_NPIX + 1) * SM_CPITCH * 4;     // 73728
constexpr size_t SM_OPER_BYTES = SM_W_BYTES + 2 * (size_t)SM_IN_WORDS_PAD * 4;
constexpr size_t SM_SMEM = 128 + (SM_CONV_BYTES > SM_OPER_BYTES ? SM_CONV_BYTES : SM_OPER_BYTES);
static_assert(SM_WARPS * SM_MT * 16 >= SM_NPIX, "m tiles must cover the conv tile");
static_assert(SM_PH == SM_WARPS, "one warp per pooled row of the tile");

struct StemMmaArgs {
    const float* x;           // [n,3,h,w] contiguous
    const uint4* wfrag;       // [11][8][32] x {b0_hi, b1_hi, b0_lo, b1_lo}
    const float *bn_scale, *bn_shift, *nx_scale, *nx_shift;
    float* out;               // [n,hp,wp,64]
    uint4* obits;             // [n][1][hp][wp]
    float x_scale, inv_scale; // 2^sx, 2^-(sx+sw)
    const float* x_amax;      // device scalar max|x|: when given, sx is chosen in the kernel (max|x| * 2^sx in [2^14, 2^15))
    int w_log2_scale;
    int N, H, W, Hc, Wc, Hp, Wp, tiles_h, tiles_w;
};

__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, const float (&c)[4]) {
    // not volatile: a pure function of its operands, so ptxas may interleave independent accumulator chains
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
}

// (hi, lo) fp16 split of two adjacent scaled inputs, packed as the two halves of a word each
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(v0, v1);               // low half = v0 (the even column / lower k index)
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

template <bool CHAIN>
__global__ void __launch_bounds__(SM_WARPS * 32, 2)
stem_mma_kernel(const __grid_constant__ StemMmaArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    const uint4* w_s = reinterpret_cast<const uint4*>(smem + 128);
    uint32_t* in_hi = reinterpret_cast<uint32_t*>(smem + 128 + SM_W_BYTES);
    uint32_t* in_lo = in_hi + SM_IN_WORDS_PAD;
    float* conv_s = reinterpret_cast<float*>(smem + 128);        // aliases the operands after the MMA phase

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    int tile = blockIdx.x;
    const int tw = tile % a.tiles_w; tile /= a.tiles_w;
    const int th = tile % a.tiles_h;
    const int n = tile / a.tiles_h;
    const int ph0 = th * SM_PH, pw0 = tw * SM_PW;
    const int cr0 = 2 * ph0 - 1, cc0 = 2 * pw0 - 1;            // first conv row / col of the tile
    const int hi0 = 2 * cr0 - 3, wi0 = 2 * cc0 - 3;            // first input row / col of the window

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        mbar_expect_tx(bar, (unsigned)SM_W_BYTES);
        bulk_load_1d(const_cast<uint4*>(w_s), a.wfrag, (unsigned)SM_W_BYTES, bar);
    }
    float x_scale = a.x_scale, inv_scale = a.inv_scale;
    if (a.x_amax != nullptr) {
        // guarded input range: the power-of-two scale follows the measured max|x| (bnn_amax_f32), so no input can
        // leave the fp16 range whatever its magnitude
        const int e = (int)((__float_as_uint(__ldg(a.x_amax)) >> 23) & 0xffu) - 126;
        int sx = 15 - e;
        sx = sx < -60 ? -60 : (sx > 60 ? 60 : sx);
        x_scale = __int_as_float((127 + sx) << 23);
        inv_scale = __int_as_float((127 - (sx + a.w_log2_scale)) << 23);
    }
    // input window: fp32 -> scaled (hi, lo) fp16 pairs; zero fill = the convolution's padding.  Column 35 of a
    // row is the zero-weight eighth tap of the last pixel: it must be finite, so it is zero as well.
    {
        // thread -> word i = tid + 256 k of the window, (rf, pc) = divmod(i, 18) kept incrementally (256 = 14 * 18 + 4);
        // rf = c_in * 39 + window row.  32-bit offsets inside the image (3 * H * W < 2^31, checked by the caller).
        constexpr int ITERS = (SM_IN_WORDS + SM_WARPS * 32 - 1) / (SM_WARPS * 32);      // 9
        const float* xn = a.x + (size_t)n * 3 * a.H * a.W;
        int rf = threadIdx.x / SM_IPW, pc = threadIdx.x - rf * SM_IPW;
        float v0[ITERS], v1[ITERS];
#pragma unroll
        for (int k = 0; k < ITERS; ++k) {                      // all loads in flight before the first conversion
            const int ci = (rf >= SM_IR) + (rf >= 2 * SM_IR);
            const int hi = hi0 + rf - ci * SM_IR, wi = wi0 + 2 * pc;
            const bool rowok = (unsigned)hi < (unsigned)a.H && (k + 1 < ITERS || rf < 3 * SM_IR);
            const int off = (ci * a.H + hi) * a.W + wi;
            const bool ok0 = rowok && (unsigned)wi < (unsigned)a.W;
            const bool ok1 = rowok && pc < SM_IPW - 1 && (unsigned)(wi + 1) < (unsigned)a.W;
            v0[k] = ok0 ? __ldg(xn + off) : 0.0f;
            v1[k] = ok1 ? __ldg(xn + off + 1) : 0.0f;
            rf += (SM_WARPS * 32) / SM_IPW; pc += (SM_WARPS * 32) % SM_IPW;
            if (pc >= SM_IPW) { pc -= SM_IPW; rf += 1; }
        }
#pragma unroll
        for (int k = 0; k < ITERS; ++k) {
            const int i = threadIdx.x + k * SM_WARPS * 32;
            uint32_t h, l;
            split2(v0[k] * x_scale, v1[k] * x_scale, h, l);
            if (k + 1 < ITERS || i < SM_IN_WORDS) { in_hi[i] = h; in_lo[i] = l; }
        }
    }
    __syncthreads();
    mbar_wait(bar, 0);

    // ---------------- implicit GEMM on mma.sync ----------------
    float acc[SM_MT][SM_NT][4];
#pragma unroll
    for (int mt = 0; mt < SM_MT; ++mt)
#pragma unroll
        for (int j = 0; j < SM_NT; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][j][i] = 0.0f;
    int pb[SM_MT][2];                                          // word offset of (pixel, kw pair t) inside a kernel row
#pragma unroll
    for (int mt = 0; mt < SM_MT; ++mt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            int p = (warp * SM_MT + mt) * 16 + g + 8 * hh;
            p = p < SM_NPIX ? p : SM_NPIX - 1;                 // the one pad pixel reads a valid address, result unused
            const int r = p / SM_CC, c = p - r * SM_CC;
            pb[mt][hh] = 2 * r * SM_IPW + c + t;
        }
    const float zero[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int s = 0; s < SM_KSTEPS; ++s) {
        const int rowA = 2 * s, rowB = (2 * s + 1 < 21) ? 2 * s + 1 : 2 * s;      // row 21 does not exist: zero weights
        const int offA = ((rowA / 7) * SM_IR + rowA % 7) * SM_IPW, offB = ((rowB / 7) * SM_IR + rowB % 7) * SM_IPW;
        uint32_t ah[SM_MT][4], al[SM_MT][4];
#pragma unroll
        for (int mt = 0; mt < SM_MT; ++mt) {
            ah[mt][0] = in_hi[offA + pb[mt][0]]; ah[mt][1] = in_hi[offA + pb[mt][1]];
            ah[mt][2] = in_hi[offB + pb[mt][0]]; ah[mt][3] = in_hi[offB + pb[mt][1]];
            al[mt][0] = in_lo[offA + pb[mt][0]]; al[mt][1] = in_lo[offA + pb[mt][1]];
            al[mt][2] = in_lo[offB + pb[mt][0]]; al[mt][3] = in_lo[offB + pb[mt][1]];
        }
#pragma unroll
        for (int jp = 0; jp < SM_NT; jp += 2) {                // two n tiles x two m tiles = four independent chains
            uint4 b[2];
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) b[jj] = w_s[(s * SM_NT + jp + jj) * 32 + lane];    // {b0_hi, b1_hi, b0_lo, b1_lo}
            if (CHAIN) {
                // experiment: every product accumulates inside the tensor core (no fp32 adds outside)
#pragma unroll
                for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                    for (int mt = 0; mt < SM_MT; ++mt) mma_f16(acc[mt][jp + jj], al[mt], b[jj].x, b[jj].y, acc[mt][jp + jj]);
#pragma unroll
                for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                    for (int mt = 0; mt < SM_MT; ++mt) mma_f16(acc[mt][jp + jj], ah[mt], b[jj].z, b[jj].w, acc[mt][jp + jj]);
#pragma unroll
                for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                    for (int mt = 0; mt < SM_MT; ++mt) mma_f16(acc[mt][jp + jj], ah[mt], b[jj].x, b[jj].y, acc[mt][jp + jj]);
            } else {
            float d[2][SM_MT][4];
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                for (int mt = 0; mt < SM_MT; ++mt) mma_f16(d[jj][mt], al[mt], b[jj].x, b[jj].y, zero);          // xl * wh
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                for (int mt = 0; mt < SM_MT; ++mt) mma_f16(d[jj][mt], ah[mt], b[jj].z, b[jj].w, d[jj][mt]);     // xh * wl
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                for (int mt = 0; mt < SM_MT; ++mt) mma_f16(d[jj][mt], ah[mt], b[jj].x, b[jj].y, d[jj][mt]);     // xh * wh
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                for (int mt = 0; mt < SM_MT; ++mt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[mt][jp + jj][i] += d[jj][mt][i];
            }
        }
    }
    __syncthreads();                                           // every warp is done with the operand buffers

    // ---------------- BN + ReLU into the shared conv tile ----------------
#pragma unroll
    for (int j = 0; j < SM_NT; ++j) {
        const int ch = 8 * j + 2 * t;
        const float2 gs = __ldg(reinterpret_cast<const float2*>(a.bn_scale + ch));
        const float2 hs = __ldg(reinterpret_cast<const float2*>(a.bn_shift + ch));
#pragma unroll
        for (int mt = 0; mt < SM_MT; ++mt)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int p = (warp * SM_MT + mt) * 16 + g + 8 * hh;
                const int r = p / SM_CC, c = p - r * SM_CC;
                // positions outside the conv output are max-pool padding: 0 is neutral after the ReLU
                const bool ok = (unsigned)(cr0 + r) < (unsigned)a.Hc && (unsigned)(cc0 + c) < (unsigned)a.Wc && p < SM_NPIX;
                float2 v;
                v.x = ok ? fmaxf(__fmaf_rn(acc[mt][j][2 * hh] * inv_scale, gs.x, hs.x), 0.0f) : 0.0f;
                v.y = ok ? fmaxf(__fmaf_rn(acc[mt][j][2 * hh + 1] * inv_scale, gs.y, hs.y), 0.0f) : 0.0f;
                *reinterpret_cast<float2*>(conv_s + p * SM_CPITCH + ch) = v;
            }
    }
    __syncthreads();

    // ---------------- 3x3 / stride 2 max, NHWC store, planes for the first binarized conv ----------------
    // warp = pooled row of the tile, its 7 pooled pixels unrolled (constant shared-memory offsets), lanes <-> channels
    const int ph = ph0 + warp;
    if (ph < a.Hp) {
        const bool has_nx = a.nx_scale != nullptr;
        float nxs[2] = {1.0f, 1.0f}, nxh[2] = {0.0f, 0.0f};
        if (has_nx) {
            nxs[0] = __ldg(a.nx_scale + lane); nxs[1] = __ldg(a.nx_scale + 32 + lane);
            nxh[0] = __ldg(a.nx_shift + lane); nxh[1] = __ldg(a.nx_shift + 32 + lane);
        }
        const size_t pix0 = ((size_t)n * a.Hp + ph) * a.Wp + pw0;
        float* orow = a.out + pix0 * 64 + lane;
        const float* cbase = conv_s + (2 * warp * SM_CC) * SM_CPITCH + lane;
#pragma unroll
        for (int pc = 0; pc < SM_PW; ++pc) {
            if (pw0 + pc < a.Wp) {                             // warp-uniform
                uint32_t sw[2], mw[2];
#pragma unroll
                for (int cb = 0; cb < 2; ++cb) {
                    float m = 0.0f;
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll
                        for (int j = 0; j < 3; ++j) m = fmaxf(m, cbase[(i * SM_CC + 2 * pc + j) * SM_CPITCH + cb * 32]);
                    orow[pc * 64 + cb * 32] = m;
                    const float b = has_nx ? __fmaf_rn(nxs[cb], m, nxh[cb]) : m;
                    sw[cb] = __ballot_sync(0xffffffffu, b > 0.0f);
                    mw[cb] = __ballot_sync(0xffffffffu, b > 0.0f || b < 0.0f);
                }
                if (lane == 0 && a.obits) a.obits[pix0 + pc] = make_uint4(sw[0], sw[1], mw[0], mw[1]);
            }
        }
    }
}

// conv weight [64,3,7,7] fp32 -> B fragments of mma.m16n8k16 (col-major k x n), hi and lo halves of w * 2^sw
__global__ void stem_mma_pack_weight_kernel(const float* __restrict__ w, float w_scale, uint4* __restrict__ frag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= SM_KSTEPS * SM_NT * 32) return;
    const int lane = i & 31, j = (i >> 5) % SM_NT, s = i / (32 * SM_NT);
    const int g = lane >> 2, t = lane & 3, ch = 8 * j + g;
    uint32_t hi[2], lo[2];
#pragma unroll
    for (int half = 0; half < 2; ++half) {                     // b0: kernel row 2s, b1: kernel row 2s + 1
        const int row = 2 * s + half;                          // (c_in, kh) = divmod(row, 7)
        float v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int kw = 2 * t + e;
            v[e] = (row < 21 && kw < 7) ? w[(ch * 21 + row) * 7 + kw] * w_scale : 0.0f;
        }
        split2(v[0], v[1], hi[half], lo[half]);
    }
    frag[i] = make_uint4(hi[0], hi[1], lo[0], lo[1]);
}

}  // namespace bnn

using namespace bnn;

extern "C" size_t bnn_stem_mma_weight_bytes(void) { return SM_W_BYTES; }

extern "C" int bnn_stem_mma_pack_weight(const float* w, int32_t w_log2_scale, void* w_frag, void* stream_) {
    if (!w || !w_frag) return BNN_E_NULL;
    if (w_log2_scale < -60 || w_log2_scale > 60) return BNN_E_SHAPE;
    if ((uintptr_t)w_frag & 15) return BNN_E_ALIGN;
    const int total = SM_KSTEPS * SM_NT * 32;
    stem_mma_pack_weight_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(w, ldexpf(1.0f, w_log2_scale), (uint4*)w_frag);
    count_launch(1);
    return (int)cudaGetLastError();
}

extern "C" int bnn_stem_mma_fwd(const float* x, int32_t n, int32_t h, int32_t w, const void* w_frag,
                                int32_t x_log2_scale, const float* x_amax, int32_t w_log2_scale, const float* bn_scale,
                                const float* bn_shift, const float* nx_scale, const float* nx_shift, float* out,
                                void* out_bits, uint32_t flags, void* stream_) {
    if (!x || !w_frag || !bn_scale || !bn_shift || !out) return BNN_E_NULL;
    if ((nx_scale == nullptr) != (nx_shift == nullptr)) return BNN_E_NULL;
    if (n <= 0 || h < 7 || w < 7) return BNN_E_SHAPE;
    if ((long long)h * w * 3 >= 0x7fffffffLL) return BNN_E_UNSUPPORTED;
    if (x_log2_scale < -60 || x_log2_scale > 60 || w_log2_scale < -60 || w_log2_scale > 60) return BNN_E_SHAPE;
    if (((uintptr_t)w_frag & 15) || ((uintptr_t)out_bits & 15) || ((uintptr_t)bn_scale & 7) || ((uintptr_t)bn_shift & 7))
        return BNN_E_ALIGN;
    StemMmaArgs a{};
    a.x = x; a.wfrag = (const uint4*)w_frag; a.bn_scale = bn_scale; a.bn_shift = bn_shift;
    a.nx_scale = nx_scale; a.nx_shift = nx_shift; a.out = out; a.obits = (uint4*)out_bits;
    a.x_scale = ldexpf(1.0f, x_log2_scale);
    a.inv_scale = ldexpf(1.0f, -(x_log2_scale + w_log2_scale));
    a.x_amax = x_amax; a.w_log2_scale = w_log2_scale;
    a.N = n; a.H = h; a.W = w;
    a.Hc = (h + 6 - 7) / 2 + 1; a.Wc = (w + 6 - 7) / 2 + 1;
    a.Hp = (a.Hc + 2 - 3) / 2 + 1; a.Wp = (a.Wc + 2 - 3) / 2 + 1;
    a.tiles_h = (a.Hp + SM_PH - 1) / SM_PH; a.tiles_w = (a.Wp + SM_PW - 1) / SM_PW;
    const bool chain = (flags & BNN_F_STEM_MMA_CHAIN) != 0;
    const void* fn = chain ? (const void*)stem_mma_kernel<true> : (const void*)stem_mma_kernel<false>;
    cudaError_t ce = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_SMEM);
    if (ce != cudaSuccess) return (int)ce;
    const long long ctas = (long long)n * a.tiles_h * a.tiles_w;
    if (ctas > 0x7fffffffLL) return BNN_E_UNSUPPORTED;
    if (chain) stem_mma_kernel<true><<<(unsigned)ctas, SM_WARPS * 32, SM_SMEM, (cudaStream_t)stream_>>>(a);
    else stem_mma_kernel<false><<<(unsigned)ctas, SM_WARPS * 32, SM_SMEM, (cudaStream_t)stream_>>>(a);
    count_launch(1);
    return (int)cudaGetLastError();
}
