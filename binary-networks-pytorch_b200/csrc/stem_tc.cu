// stem_tc.cu -- the fp32 stem of bnn.models.resnet on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM):
//     conv 7x7 / stride 2 / pad 3 (3 -> 64)  ->  eval BatchNorm  ->  ReLU  ->  MaxPool 3x3 / 2 / pad 1
// (reference bnn/models/resnet.py:85-92,147-153).  Same contract and the same split-fp16 arithmetic as stem_mma.cu
// (x*2^sx = xh + xl, w*2^sw = wh + wl, x*w ~= xh*wh + xh*wl + xl*wh: 22 significand bits per operand, products exact in
// the fp32 accumulator), re-designed around what tcgen05 wants:
//
// * Implicit GEMM per CONV ROW: M = 128 consecutive conv columns of one conv row (lane m <-> column 2*pc0 - 1 + m, so a
//   tile of up to 63 pooled columns has its 3-wide pooling windows inside one M tile), N = 64 channels, K = 192.
// * No im2col tile is ever materialised.  The input is staged as 16-byte UNITS  U[ci][u] = {x[ci][y0+i][c0+2u+b]},
//   i = 0..3 (four input rows of a "group"), b = 0,1 (a column pair): 8 fp16 values.  Conv column m needs, for its kernel
//   columns (2p, 2p+1), p = 0..3, exactly unit m + p -- a Hankel matrix A[m][chunk j] = U[m + j].  The K-major no-swizzle
//   shared-memory descriptor expresses that directly: rows of a core matrix 16 bytes apart (fixed by the layout),
//   8-row groups 128 bytes apart (SBO), the second 16-byte K chunk 16 bytes further (LBO = 16: overlapping core
//   matrices).  One MMA (K = 16) covers kernel rows 4g..4g+3 x kernel columns 4pp..4pp+3 of one input channel; a conv row
//   is 2 groups x 3 channels x 2 column halves = 12 K steps (K = 192 for the 147 real taps: row 7 and column 7 are
//   zero weights).  Conv row r uses groups r and r + 2 (group k = input rows 2k-3 .. 2k), so consecutive conv rows
//   share three quarters of their operands: a group is converted once and read by two conv rows.
// * Per K step two instructions: A_hi x [wh | wl] (N = 128: xh*wh into columns 0..63, xh*wl into 64..127) and
//   A_lo x wh (N = 64, accumulating into columns 0..63); the epilogue adds the two column halves with a rounded fp32
//   add.  14 KB of shared-memory operand reads per 96 tensor cycles.
// * Warp-specialised persistent CTA, one per SM, walking a contiguous range of (image, pooled row):
//     warps 0-3   convert fp32 input rows to (hi, lo) fp16 units into a 6-deep ring of groups (mbarrier full / empty)
//     warp 19     one thread issues the MMAs of a conv row into one of 2 TMEM accumulator stages, tcgen05.commit
//                 releases the stage to the accumulator warps and the oldest group back to the converters
//     warps 4-11  tcgen05.ld the accumulators (warp % 4 = TMEM lane quarter, two channel halves), BatchNorm + ReLU, keep
//                 the running vertical max of the 3 conv rows of a pooled row in REGISTERS (thread = conv column) and
//                 park it in shared memory once per pooled row (double-buffered, mbarrier full / empty)
//     warps 12-18 take the horizontal 3-max of a parked row with lanes <-> channels, store NHWC lines and ballot the
//                 first binarized layer's planes.
#include "tc05.cuh"

#include <cuda_fp16.h>
#include <type_traits>

namespace bnn {

constexpr int TC_NU = 128;                               // 16-byte units per (group, channel, hi/lo) row = converter threads
constexpr int TC_D = 6;                                  // groups in the ring
constexpr int TC_SLOT = 3 * 2 * TC_NU * 16;              // 12288 bytes per group
constexpr int TC_KSTEPS = 12;
constexpr int TC_BSTEP = 4096;                           // [wh (64 rows) | wl (64 rows)] x 32 bytes per K step
constexpr int TC_B_BYTES = TC_KSTEPS * TC_BSTEP;         // 49152
constexpr int TC_VPITCH = 68;                            // floats per parked conv column (conflict-free STS.128)
constexpr int TC_VBUF = 128 * TC_VPITCH * 4;             // 34816
constexpr int TC_STAGES = 2;                             // TMEM accumulator stages: [hh0 | sm | hh1] (192 of 256 columns)
constexpr int TC_STAGE_COLS = 256;
constexpr int TC_CONV_WARPS = 4, TC_EPI_WARPS = 8, TC_OUT_WARPS = 7;
constexpr int TC_MMA_WARP = TC_CONV_WARPS + TC_EPI_WARPS + TC_OUT_WARPS;  // 19
constexpr int TC_THREADS = (TC_MMA_WARP + 1) * 32;                        // 640: 96 registers per thread
constexpr int TC_OFF_CONST = 256, TC_OFF_B = 2048, TC_OFF_RING = TC_OFF_B + TC_B_BYTES;
constexpr int TC_OFF_VBUF = TC_OFF_RING + TC_D * TC_SLOT;
constexpr int TC_SMEM = TC_OFF_VBUF + 2 * TC_VBUF;                        // 194560
constexpr int TC_MAX_PT = 62;                            // pooled columns per M tile: conv columns m <= 124, units m + p <= 127
constexpr int TC_MAX_CT = 125;                           // conv columns per M tile without pooling
static_assert(TC_CONV_WARPS * 32 == TC_NU, "one converter thread per unit");

struct StemTcArgs {
    const void* x;            // IN 0: fp32 [n,3,h,w];  IN 1: uint8 [n,h,w,3]
    const void* wops;         // bnn_stem_tc_pack_weight output
    const float *bn_scale, *bn_shift, *nx_scale, *nx_shift, *nx2_scale, *nx2_shift;
    const float* x_amax;      // device scalar max|x| (NULL: x_log2_scale is used as given)
    float* out;               // pool: [n,hp,wp,64]; no pool: [n,hc,wc,64]
    uint4 *obits, *obits2;    // [n][1][rows][cols] planes of sign(out*nx + nx_shift) / sign(out*nx2 + nx2_shift)
    float u8_mean[3], u8_istd[3];
    int x_log2_scale, w_log2_scale, nx_relu, nx2_relu, vec2;
    int dbg;                  // timing experiments only (flags >> 8 of bnn_stem_tc_run): 1 accumulator warps idle, 2 output warps idle,
                              // 4 converters idle, 8 no MMAs -- results are then garbage
    int N, H, W, Hc, Wc, Ho, Wo, tiles_w, TW;   // Ho x Wo: output rows / cols (pooled or conv); TW: output cols per tile
    long long total_rows;     // N * tiles_w * Ho output rows, split evenly over the CTAs
};

struct Seg { int n, cbase, col0, row_a, r_first, nrows; };

// next run of output rows of one (image, column tile) inside [pos, hi)
template <bool POOL>
__device__ __forceinline__ bool next_seg(long long& pos, long long hi, const StemTcArgs& a, Seg& s) {
    if (pos >= hi) return false;
    const long long img = pos / a.Ho;
    const int row_a = (int)(pos - img * a.Ho);
    const long long end = (img + 1) * a.Ho < hi ? (img + 1) * a.Ho : hi;
    const int row_b = row_a + (int)(end - pos);
    s.n = (int)(img / a.tiles_w);
    s.col0 = (int)(img % a.tiles_w) * a.TW;
    s.row_a = row_a;
    if (POOL) {
        s.cbase = 2 * s.col0 - 1;                                  // lane m <-> conv column 2*pc0 - 1 + m
        s.r_first = 2 * row_a - 1 > 0 ? 2 * row_a - 1 : 0;
        const int r_last = 2 * row_b - 1 < a.Hc - 1 ? 2 * row_b - 1 : a.Hc - 1;
        s.nrows = r_last - s.r_first + 1;
    } else {
        s.cbase = s.col0;
        s.r_first = row_a;
        s.nrows = row_b - row_a;
    }
    pos = end;
    return true;
}

__device__ __forceinline__ float pow2f(int e) { return __int_as_float((127 + e) << 23); }      // e in [-126, 127]

// power-of-two input scale: the immediate, or from max|x| so that max|x| * 2^sx lies in [2^14, 2^15)
__device__ __forceinline__ int input_log2_scale(const StemTcArgs& a) {
    if (a.x_amax == nullptr) return a.x_log2_scale;
    const int e = (int)((__float_as_uint(__ldg(a.x_amax)) >> 23) & 0xffu) - 126;       // amax = m * 2^e, m in [0.5, 1)
    const int sx = 15 - e;
    return sx < -60 ? -60 : (sx > 60 ? 60 : sx);
}

// two input rows (y, y + 1) x 3 channels x the unit's column pair, as they come from memory
template <int IN> struct Raw2 { float2 v[2][3]; };
template <> struct Raw2<1> { unsigned short v[2][3]; };              // uint8 NHWC: 6 bytes per row = three 16-bit loads

template <int IN>
__device__ __forceinline__ void load2(const StemTcArgs& a, int n, int y, int colL, Raw2<IN>& r) {
    const bool cok = (unsigned)colL < (unsigned)a.W;                 // colL is even and (vector path) W is even: pair in or out
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int yy = y + i;
        const bool ok = cok && (unsigned)yy < (unsigned)a.H;
        if constexpr (IN == 0) {
            const float* x = reinterpret_cast<const float*>(a.x);
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                const float* p = x + (((size_t)n * 3 + ci) * a.H + (ok ? yy : 0)) * a.W + (ok ? colL : 0);
                if (a.vec2) {
                    r.v[i][ci] = ok ? __ldg(reinterpret_cast<const float2*>(p)) : make_float2(0.0f, 0.0f);
                } else {                                             // odd W: element-wise, the right column may be outside
                    const bool ok1 = ok && colL + 1 < a.W;
                    r.v[i][ci].x = ok ? __ldg(p) : 0.0f;
                    r.v[i][ci].y = ok1 ? __ldg(p + 1) : 0.0f;
                }
            }
        } else {
            const unsigned short* p = reinterpret_cast<const unsigned short*>(
                reinterpret_cast<const unsigned char*>(a.x) + (((size_t)n * a.H + (ok ? yy : 0)) * a.W + (ok ? colL : 0)) * 3);
#pragma unroll
            for (int k = 0; k < 3; ++k) r.v[i][k] = ok ? __ldg(p + k) : (unsigned short)0;        // cvt2 zeroes outside positions
        }
    }
}

// -> scaled (hi, lo) fp16 pairs: hi[ci][row], lo[ci][row], each the half2 {left column, right column}
template <int IN>
__device__ __forceinline__ void cvt2(const StemTcArgs& a, const Raw2<IN>& r, float xs, bool in0, bool in1, uint32_t (&hi)[3][2], uint32_t (&lo)[3][2]) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
            float l, rr;
            if constexpr (IN == 0) {
                l = r.v[i][ci].x; rr = r.v[i][ci].y;
            } else {
                // bytes of a row: [c0 ch0, c0 ch1, c0 ch2, c1 ch0, c1 ch1, c1 ch2]; normalisation = two rounded fp32 operations
                const uint32_t w01 = r.v[i][0], w23 = r.v[i][1], w45 = r.v[i][2];
                const uint32_t b0 = ci == 0 ? (w01 & 0xff) : ci == 1 ? (w01 >> 8) : (w23 & 0xff);
                const uint32_t b1 = ci == 0 ? (w23 >> 8) : ci == 1 ? (w45 & 0xff) : (w45 >> 8);
                l = __fmul_rn(__fsub_rn((float)b0, a.u8_mean[ci]), a.u8_istd[ci]);
                rr = __fmul_rn(__fsub_rn((float)b1, a.u8_mean[ci]), a.u8_istd[ci]);
                if (!(i ? in1 : in0)) { l = 0.0f; rr = 0.0f; }       // zero padding of the NORMALISED image
            }
            const float sa = l * xs, sb = rr * xs;
            const __half2 hh = __floats2half2_rn(sa, sb);
            const float2 hf = __half22float2(hh);
            const __half2 ll = __floats2half2_rn(sa - hf.x, sb - hf.y);
            hi[ci][i] = *reinterpret_cast<const uint32_t*>(&hh);
            lo[ci][i] = *reinterpret_cast<const uint32_t*>(&ll);
        }
}

// timing experiments (dbg & 16): CTA 0 records clock64() stamps of its roles into the buffer passed as out_bits2:
// tl[role][event index][slot], role 0 converter warp 0, 1 MMA issuer, 2 accumulator warp 4, 3 output warp 12; 4 slots per event
#define TC_TL(role, idx, slot)                                                                                         \
    do {                                                                                                               \
        if (tl != nullptr && lane == 0 && (idx) < 512) tl[((role) * 512 + (idx)) * 4 + (slot)] = clock64();            \
    } while (0)

template <bool POOL, int IN>
__global__ void __launch_bounds__(TC_THREADS, 1) stem_tc_kernel(const __grid_constant__ StemTcArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);             // [TC_D]   converters -> MMA
    uint64_t* empty = full + TC_D;                                  // [TC_D]   MMA (commit) -> converters
    uint64_t* acc_full = empty + TC_D;                              // [TC_STAGES] MMA (commit) -> epilogue
    uint64_t* acc_empty = acc_full + TC_STAGES;                     // [TC_STAGES] epilogue -> MMA
    uint64_t* bbar = acc_empty + TC_STAGES;                         // weights landed
    uint64_t* vfull = bbar + 1;                                     // [2] accumulator warps -> output warps (parked rows)
    uint64_t* vempty = vfull + 2;                                   // [2] output warps -> accumulator warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(vempty + 2);
    float* consts = reinterpret_cast<float*>(smem + TC_OFF_CONST);  // [6][64]: bn_scale * 2^-(sx+sw), bn_shift, nx, nx2
    unsigned char* b_s = smem + TC_OFF_B;
    unsigned char* ring = smem + TC_OFF_RING;
    float* vbuf = reinterpret_cast<float*>(smem + TC_OFF_VBUF);

    pdl_launch_dependents();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    long long* const tl = ((a.dbg & 16) && blockIdx.x == 0) ? reinterpret_cast<long long*>(a.obits2) : nullptr;
    const long long lo_row = a.total_rows * blockIdx.x / gridDim.x, hi_row = a.total_rows * (blockIdx.x + 1) / gridDim.x;

    if (warp == TC_MMA_WARP) {
        if (lane == 0) {
            for (int i = 0; i < TC_D; ++i) { mbar_init(full + i, TC_CONV_WARPS); mbar_init(empty + i, 1); }
            for (int i = 0; i < TC_STAGES; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, TC_EPI_WARPS); }
            for (int i = 0; i < 2; ++i) { mbar_init(vfull + i, TC_EPI_WARPS); mbar_init(vempty + i, TC_OUT_WARPS); }
            mbar_init(bbar, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tc05::tmem_alloc<512>(tmem_slot);
    }
    // programmatic dependent launch (common.cuh): barrier init and the TMEM allocation above overlap the previous kernel's
    // tail; nothing before this line touches global memory
    pdl_wait();
    if (warp == TC_MMA_WARP && lane == 0) {
        mbar_expect_tx(bbar, (unsigned)TC_B_BYTES);
        bulk_load_1d(b_s, a.wops, (unsigned)TC_B_BYTES, bbar);
    }
    const int sx = input_log2_scale(a);
    for (int i = tid; i < 384; i += TC_THREADS) {
        const int which = i >> 6, c = i & 63;
        const float* src = which == 0 ? a.bn_scale : which == 1 ? a.bn_shift : which == 2 ? a.nx_scale
                         : which == 3 ? a.nx_shift : which == 4 ? a.nx2_scale : a.nx2_shift;
        float v = src ? __ldg(src + c) : ((which & 1) ? 0.0f : 1.0f);
        // the accumulators hold conv * 2^(sx+sw): a power of two folds into the BatchNorm scale exactly
        if (which == 0) v *= pow2f(-(sx + a.w_log2_scale));
        consts[i] = v;
    }
    tc05::fence_before_sync();
    __syncthreads();
    tc05::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp < TC_CONV_WARPS) {
        // =================== converters: input rows -> (hi, lo) fp16 units; thread u owns unit u of every group ==========
        // group k = input rows 2(r_first+k)-3 .. +3; consecutive groups share two rows, so per group two NEW rows are
        // loaded (prefetched one group ahead), converted once and kept for the next group
        const float xs = pow2f(sx);
        const int u = tid;
        long long pos = lo_row, K = 0;
        Seg s;
        while (next_seg<POOL>(pos, hi_row, a, s)) {
            const int colL = 2 * (s.cbase - 2 + u);                 // unit u <-> input columns (colL, colL + 1)
            const bool cin = (unsigned)colL < (unsigned)a.W;
            const int ngroups = s.nrows + 2;
            int y = 2 * s.r_first - 3;
            Raw2<IN> ra, rb;
            load2<IN>(a, s.n, y, colL, ra);
            load2<IN>(a, s.n, y + 2, colL, rb);
            uint32_t phi[3][2], plo[3][2];
            cvt2<IN>(a, ra, xs, cin && (unsigned)y < (unsigned)a.H, cin && (unsigned)(y + 1) < (unsigned)a.H, phi, plo);
            for (int k = 0; k < ngroups; ++k, ++K) {
                // rb holds rows y + 2, y + 3 (the new half of group k); prefetch the new half of group k + 1
                Raw2<IN> rn;
                if (k + 1 < ngroups && !(a.dbg & 4)) load2<IN>(a, s.n, y + 4, colL, rn);
                uint32_t chi[3][2], clo[3][2];
                cvt2<IN>(a, rb, xs, cin && (unsigned)(y + 2) < (unsigned)a.H, cin && (unsigned)(y + 3) < (unsigned)a.H, chi, clo);
                const int slot = (int)(K % TC_D);
                const long long use = K / TC_D;
                if (warp == 0) TC_TL(0, (int)K, 0);
                if (use > 0) mbar_wait(empty + slot, (uint32_t)((use - 1) & 1));
                if (warp == 0) TC_TL(0, (int)K, 1);
                unsigned char* sb = ring + (size_t)slot * TC_SLOT + (size_t)u * 16;
                if (!(a.dbg & 4))
#pragma unroll
                for (int ci = 0; ci < 3; ++ci) {
                    *reinterpret_cast<uint4*>(sb + (size_t)(ci * 2 + 0) * TC_NU * 16) = make_uint4(phi[ci][0], phi[ci][1], chi[ci][0], chi[ci][1]);
                    *reinterpret_cast<uint4*>(sb + (size_t)(ci * 2 + 1) * TC_NU * 16) = make_uint4(plo[ci][0], plo[ci][1], clo[ci][0], clo[ci][1]);
                }
                fence_proxy_async();                  // generic-proxy stores -> visible to the tensor core's async proxy
                __syncwarp();
                if (lane == 0) tc05::mbar_arrive(full + slot);
                if (warp == 0) TC_TL(0, (int)K, 2);
#pragma unroll
                for (int ci = 0; ci < 3; ++ci) { phi[ci][0] = chi[ci][0]; phi[ci][1] = chi[ci][1]; plo[ci][0] = clo[ci][0]; plo[ci][1] = clo[ci][1]; }
                if (k + 1 < ngroups) rb = rn;
                y += 2;
            }
        }
    } else if (warp == TC_MMA_WARP) {
        // =================== MMA issuer ===================
        // The tensor core's fp32 accumulate truncates; the error grows with the length of an accumulation chain.  So the
        // large products xh*wh of the two groups go to two separate accumulators hh0 / hh1 (chains of 6 instead of 12) and
        // the small terms (xh*wl, xl*wh: 2^-11 of the result) to their own columns sm; the epilogue adds the three with
        // rounded fp32 adds (max error vs float64: 7e-7 -> 2.3e-7 of max|y|, rms 1.1e-7 -> 3e-8).
        mbar_wait(bbar, 0);
        const uint32_t idesc128 = tc05::idesc_f16_f32(128, 128), idesc64 = tc05::idesc_f16_f32(128, 64);
        // descriptor low words (address in 16-byte units | LBO field); all shared addresses are below 2^14 units, so adding a
        // unit offset never carries into the LBO field
        const uint32_t ring_lo = tc05::desc_lo(smem_u32(ring) >> 4, 16), b_lo0 = tc05::desc_lo(smem_u32(b_s) >> 4, 128);
        constexpr uint32_t A_HI = tc05::desc_hi(128), B_HI = tc05::desc_hi(256);
        // The issue loop is software-pipelined: the barrier waits of row j + 1 (free accumulator stage, next group landed)
        // are executed between the two halves of row j's instructions, so their wake-up latency is covered by tensor work
        // that is already queued instead of leaving the pipe idle between rows.
        struct RowIt {
            long long pos, K, J;       // next segment start, first group of the segment, global row counter
            Seg s;
            int j;                     // row inside the segment
            bool valid;
        } cur, nxt;
        auto first_row = [&](RowIt& it) {
            it.pos = lo_row; it.K = 0; it.J = 0; it.j = 0;
            it.valid = next_seg<POOL>(it.pos, hi_row, a, it.s);
        };
        auto next_row = [&](const RowIt& c, RowIt& n) {
            n = c;
            ++n.J;
            if (++n.j == c.s.nrows) {
                n.K = c.K + c.s.nrows + 2;
                n.j = 0;
                n.valid = next_seg<POOL>(n.pos, hi_row, a, n.s);
            }
        };
        auto wait_row = [&](const RowIt& it) {
            const int stage = (int)(it.J % TC_STAGES);
            const long long ause = it.J / TC_STAGES;
            TC_TL(1, (int)it.J, 0);
            if (ause > 0) mbar_wait(acc_empty + stage, (uint32_t)((ause - 1) & 1));
            const long long k0 = it.K + it.j, k2 = k0 + 2;
            if (it.j < 2) mbar_wait(full + (int)(k0 % TC_D), (uint32_t)((k0 / TC_D) & 1));    // later rows saw it as their k2 two rows ago
            mbar_wait(full + (int)(k2 % TC_D), (uint32_t)((k2 / TC_D) & 1));
            tc05::fence_after_sync();
            TC_TL(1, (int)it.J, 1);
        };
        first_row(cur);
        if (cur.valid) wait_row(cur);
        while (cur.valid) {
            const int stage = (int)(cur.J % TC_STAGES);
            const long long k0 = cur.K + cur.j, k2 = k0 + 2;
            const int s0 = (int)(k0 % TC_D), s2 = (int)(k2 % TC_D);
            // everything below is warp-uniform; only the tcgen05 instructions themselves are predicated on one lane
            const bool leader = tc05::elect_one();
            const uint32_t a0 = ring_lo + (uint32_t)s0 * (TC_SLOT / 16), a2 = ring_lo + (uint32_t)s2 * (TC_SLOT / 16);
            const uint32_t dst = tmem + (uint32_t)(stage * TC_STAGE_COLS);
            const bool run_mma = leader && !(a.dbg & 8);
            TC_TL(1, (int)cur.J, 2);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
#pragma unroll
                for (int kk = 0; kk < 6; ++kk) {
                    const int ks = half * 6 + kk;
                    const int grp = ks / 6, ci = (ks >> 1) % 3, pp = ks & 1;
                    const uint32_t au = (grp ? a2 : a0) + (uint32_t)(2 * pp);
                    const uint32_t a_hi = au + (uint32_t)(ci * 2 + 0) * TC_NU, a_lo = au + (uint32_t)(ci * 2 + 1) * TC_NU;
                    // group 0 blocks are packed [wh | wl], group 1 blocks [wl | wh]: the small-term accumulator sits between
                    // the two large-term accumulators, so one N = 128 instruction covers (hh0, sm) or (sm, hh1)
                    const uint32_t b_all = b_lo0 + (uint32_t)ks * (TC_BSTEP / 16), b_wh = b_all + (grp ? 128u : 0u);
                    if (run_mma) {
                        if (ks == 6) {
                            // first instruction into hh1 must overwrite it, but may not reset sm: three N = 64 instructions
                            tc05::mma_f16_ss_w(dst + 128u, a_hi, A_HI, b_wh, B_HI, idesc64, 0);      // xh*wh -> hh1
                            tc05::mma_f16_ss_w(dst + 64u, a_hi, A_HI, b_all, B_HI, idesc64, 1);      // xh*wl -> sm
                        } else {
                            tc05::mma_f16_ss_w(dst + (grp ? 64u : 0u), a_hi, A_HI, b_all, B_HI, idesc128, ks > 0);   // (hh0, sm) / (sm, hh1)
                        }
                        tc05::mma_f16_ss_w(dst + 64u, a_lo, A_HI, b_wh, B_HI, idesc64, 1);           // xl*wh -> sm
                    }
                }
                if (half == 0) {
                    next_row(cur, nxt);
                    if (nxt.valid) wait_row(nxt);
                }
            }
            if (leader) {
                tc05::commit(acc_full + stage);
                tc05::commit(empty + s0);                                      // group j is not needed any more
                if (cur.j == cur.s.nrows - 1) {
                    tc05::commit(empty + (int)((cur.K + cur.s.nrows) % TC_D));
                    tc05::commit(empty + (int)((cur.K + cur.s.nrows + 1) % TC_D));
                }
            }
            __syncwarp();
            TC_TL(1, (int)cur.J, 3);
            cur = nxt;
        }
    } else if (warp < TC_CONV_WARPS + TC_EPI_WARPS) {
        // =================== accumulator warps: TMEM -> BN (+ ReLU) -> [vertical max in registers] -> parked row in smem
        const int e = warp - TC_CONV_WARPS;            // 0..7
        const int lq = warp & 3;                       // TMEM lane quarter this warp may read
        const int cb = 32 * (e >> 2);                  // channel half
        const int m = 32 * lq + lane;                  // conv column inside the M tile
        long long pos = lo_row, J = 0;
        int emits = 0;
        Seg s;
        while (next_seg<POOL>(pos, hi_row, a, s)) {
            const int c = s.cbase + m;
            const bool col_ok = (unsigned)c < (unsigned)a.Wc;
            const bool all_ok = __all_sync(0xffffffffu, col_ok);
            float state[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) state[i] = 0.0f;            // 0 = max-pool padding after the ReLU (and the ReLU itself)
            for (int j = 0; j < s.nrows; ++j, ++J) {
                const int r = s.r_first + j;
                const int stage = (int)(J % TC_STAGES);
                bool emit = true;
                if (POOL) emit = ((r & 1) || r == a.Hc - 1) && ((r & 1) ? (r - 1) >> 1 : r >> 1) >= s.row_a;
                float* vb = vbuf + (size_t)(emits & 1) * (TC_VBUF / 4) + m * TC_VPITCH + cb;
                if (e == 0) TC_TL(2, (int)J, 0);
                if (emit && emits >= 2) mbar_wait(vempty + (emits & 1), (uint32_t)(((emits >> 1) - 1) & 1));
                mbar_wait(acc_full + stage, (uint32_t)((J / TC_STAGES) & 1));
                tc05::fence_after_sync();
                if (e == 0) TC_TL(2, (int)J, 1);
                const uint32_t taddr = tmem + ((uint32_t)(32 * lq) << 16) + (uint32_t)(stage * TC_STAGE_COLS + cb);
                // MODE 0: no emit (state = max(state, v));  1: emit max(state, v), an odd conv row also opens the next pooled
                // row (state = relu(v));  2: pool-less, emit relu(v).  One instance of the row body per mode: no per-element selects
                auto row = [&](auto mode_tag) {
                    constexpr int MODE = decltype(mode_tag)::value;
#pragma unroll
                    for (int hh = 0; hh < 4; ++hh) {                 // 8 channels at a time
                        float p0[8], p1[8], q[8];
                        tc05::tmem_ld8x3_sync(taddr + 8 * hh, taddr + 128 + 8 * hh, taddr + 64 + 8 * hh, p0, p1, q);   // hh0, hh1, sm
                        if (hh == 3) {                               // every column of this stage has been read
                            tc05::fence_before_sync();
                            __syncwarp();
                            if (lane == 0) tc05::mbar_arrive(acc_empty + stage);
                        }
#pragma unroll
                        for (int i4 = 0; i4 < 8; i4 += 4) {
                            const float4 g = *reinterpret_cast<const float4*>(consts + cb + 8 * hh + i4);
                            const float4 h = *reinterpret_cast<const float4*>(consts + 64 + cb + 8 * hh + i4);
                            const float gg[4] = {g.x, g.y, g.z, g.w}, hv[4] = {h.x, h.y, h.z, h.w};
                            float o[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int k = 8 * hh + i4 + i;
                                const float conv = __fadd_rn(__fadd_rn(p0[i4 + i], p1[i4 + i]), q[i4 + i]);
                                const float v = __fmaf_rn(conv, gg[i], hv[i]);
                                if (MODE == 0) {
                                    state[k] = fmaxf(state[k], v);       // state >= 0: the ReLU is implied
                                } else if (MODE == 1) {
                                    o[i] = fmaxf(state[k], v);
                                    state[k] = fmaxf(v, 0.0f);
                                } else {
                                    o[i] = fmaxf(v, 0.0f);
                                }
                            }
                            if (MODE != 0) {                         // columns outside the conv output: max-pool padding (0)
                                float4 ov = make_float4(o[0], o[1], o[2], o[3]);
                                if (!all_ok && !col_ok) ov = make_float4(0.f, 0.f, 0.f, 0.f);
                                *reinterpret_cast<float4*>(vb + 8 * hh + i4) = ov;
                            }
                        }
                    }
                };
                if (a.dbg & 1) {
                    tc05::fence_before_sync();
                    __syncwarp();
                    if (lane == 0) tc05::mbar_arrive(acc_empty + stage);
                } else if (!POOL) row(std::integral_constant<int, 2>{});
                else if (emit) row(std::integral_constant<int, 1>{});
                else row(std::integral_constant<int, 0>{});
                if (emit) {
                    __syncwarp();
                    if (lane == 0) tc05::mbar_arrive(vfull + (emits & 1));
                    ++emits;
                }
                if (e == 0) TC_TL(2, (int)J, 2);
            }
        }
    } else {
        // =================== output warps: parked rows -> [horizontal 3-max] -> NHWC lines + planes, lanes <-> channels
        const int e = warp - TC_CONV_WARPS - TC_EPI_WARPS;           // 0..6
        const bool has_nx = a.nx_scale != nullptr, has_nx2 = a.nx2_scale != nullptr && a.obits2 != nullptr;
        const bool nx_relu = a.nx_relu != 0, nx2_relu = a.nx2_relu != 0;
        const float k2[2] = {consts[128 + lane], consts[160 + lane]}, k3[2] = {consts[192 + lane], consts[224 + lane]};
        const float k4[2] = {consts[256 + lane], consts[288 + lane]}, k5[2] = {consts[320 + lane], consts[352 + lane]};
        float* const outp = a.out;
        uint4* const ob1 = a.obits;
        uint4* const ob2 = a.obits2;
        const int Hc = a.Hc, Ho = a.Ho, Wo = a.Wo, TW = a.TW;
        long long pos = lo_row;
        int emits = 0;
        Seg s;
        while (next_seg<POOL>(pos, hi_row, a, s)) {
            const int ncols = Wo - s.col0 < TW ? Wo - s.col0 : TW;  // output columns of this tile
            for (int j = 0; j < s.nrows; ++j) {
                const int r = s.r_first + j;
                int orow = r;
                if (POOL) {
                    orow = (r & 1) ? (r - 1) >> 1 : r >> 1;
                    if (!(((r & 1) || r == Hc - 1) && orow >= s.row_a)) continue;
                }
                const float* vb = vbuf + (size_t)(emits & 1) * (TC_VBUF / 4) + lane;
                if (e == 0) TC_TL(3, emits, 0);
                mbar_wait(vfull + (emits & 1), (uint32_t)((emits >> 1) & 1));
                if (e == 0) TC_TL(3, emits, 1);
                const size_t prow = ((size_t)s.n * Ho + orow) * Wo + s.col0;
                float* po = outp + (prow + e) * 64 + lane;
                const float* v0 = vb + (POOL ? 2 * e : e) * TC_VPITCH;
                for (int jp = (a.dbg & 2) ? ncols : e; jp < ncols; jp += TC_OUT_WARPS, po += TC_OUT_WARPS * 64, v0 += (POOL ? 2 : 1) * TC_OUT_WARPS * TC_VPITCH) {
                    float mx0 = v0[0], mx1 = v0[32];
                    if (POOL) {
                        mx0 = fmaxf(fmaxf(mx0, v0[TC_VPITCH]), v0[2 * TC_VPITCH]);
                        mx1 = fmaxf(fmaxf(mx1, v0[TC_VPITCH + 32]), v0[2 * TC_VPITCH + 32]);
                    }
                    po[0] = mx0;
                    po[32] = mx1;
                    const float b0 = has_nx ? __fmaf_rn(k2[0], mx0, k3[0]) : mx0, b1 = has_nx ? __fmaf_rn(k2[1], mx1, k3[1]) : mx1;
                    const uint32_t s0 = __ballot_sync(0xffffffffu, b0 > 0.0f), s1 = __ballot_sync(0xffffffffu, b1 > 0.0f);
                    uint32_t m0 = s0, m1 = s1;
                    if (!nx_relu) {
                        m0 = __ballot_sync(0xffffffffu, b0 != 0.0f && b0 == b0);          // non-zero and not NaN
                        m1 = __ballot_sync(0xffffffffu, b1 != 0.0f && b1 == b1);
                    }
                    if (lane == 0 && ob1 != nullptr) ob1[prow + jp] = make_uint4(s0, s1, m0, m1);
                    if (has_nx2) {
                        const float c0 = __fmaf_rn(k4[0], mx0, k5[0]), c1 = __fmaf_rn(k4[1], mx1, k5[1]);
                        const uint32_t t0 = __ballot_sync(0xffffffffu, c0 > 0.0f), t1 = __ballot_sync(0xffffffffu, c1 > 0.0f);
                        uint32_t n0 = t0, n1 = t1;
                        if (!nx2_relu) {
                            n0 = __ballot_sync(0xffffffffu, c0 != 0.0f && c0 == c0);
                            n1 = __ballot_sync(0xffffffffu, c1 != 0.0f && c1 == c1);
                        }
                        if (lane == 0) ob2[prow + jp] = make_uint4(t0, t1, n0, n1);
                    }
                }
                __syncwarp();
                if (lane == 0) tc05::mbar_arrive(vempty + (emits & 1));
                if (e == 0) TC_TL(3, emits, 2);
                ++emits;
            }
        }
    }
    tc05::fence_before_sync();
    __syncthreads();
    if (warp == TC_MMA_WARP) tc05::tmem_dealloc<512>(tmem);
}

// conv weight [64,3,7,7] fp32 -> the B operand image: 12 K steps x ([wh | wl] for group 0, [wl | wh] for group 1) rows x 16 K elements, K-major no-swizzle core
// matrices (8 rows x 16 bytes, the two K chunks 128 bytes apart, 8-row groups 256 bytes apart).
// K step ks = (grp * 3 + ci) * 2 + pp; K element jj * 8 + i * 2 + b  <->  w[n][ci][kh = 4 grp + i][kw = 4 pp + 2 jj + b - 1]
// (units are ALIGNED input column pairs (2q, 2q+1); the conv's left padding of 3 puts the zero tap at kw = -1).
__global__ void stem_tc_pack_weight_kernel(const float* __restrict__ w, float w_scale, __half* __restrict__ ops) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;       // one fp16 element of the image
    if (idx >= TC_B_BYTES / 2) return;
    const int ks = idx / (TC_BSTEP / 2), rem = idx % (TC_BSTEP / 2);
    const int n8 = rem / 128, jj = (rem % 128) / 64, nr = (rem % 64) / 8, el = rem % 8;
    const int grp = ks / 6, ci = (ks >> 1) % 3, pp = ks & 1;
    int n = n8 * 8 + nr;                                         // group 0 blocks: rows 0..63 wh, 64..127 wl
    if (grp) n ^= 64;                                            // group 1 blocks: rows 0..63 wl, 64..127 wh
    const int kh = 4 * grp + (el >> 1), kw = 4 * pp + 2 * jj + (el & 1) - 1;
    float v = 0.0f;
    if (kh < 7 && kw >= 0 && kw < 7) v = w[(((n & 63) * 3 + ci) * 7 + kh) * 7 + kw] * w_scale;
    const __half hi = __float2half_rn(v);
    ops[idx] = n < 64 ? hi : __float2half_rn(v - __half2float(hi));
}

// max |x| over a tensor (NaN ignored), atomically merged into *amax (which the caller zeroes first)
__global__ void __launch_bounds__(256) amax_kernel(const float* __restrict__ x, long long count, float* __restrict__ amax) {
    pdl_launch_dependents();      // programmatic dependent launch (common.cuh): no global access before pdl_wait()
    pdl_wait();
    float m = 0.0f;
    const long long n4 = count >> 2, stride = (long long)gridDim.x * blockDim.x;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {               // four independent 16-byte loads in flight per thread
        const float4 v0 = __ldg(x4 + i), v1 = __ldg(x4 + i + stride), v2 = __ldg(x4 + i + 2 * stride), v3 = __ldg(x4 + i + 3 * stride);
        m = fmaxf(m, fmaxf(fmaxf(fmaxf(fabsf(v0.x), fabsf(v0.y)), fmaxf(fabsf(v0.z), fabsf(v0.w))),
                           fmaxf(fmaxf(fabsf(v1.x), fabsf(v1.y)), fmaxf(fabsf(v1.z), fabsf(v1.w)))));
        m = fmaxf(m, fmaxf(fmaxf(fmaxf(fabsf(v2.x), fabsf(v2.y)), fmaxf(fabsf(v2.z), fabsf(v2.w))),
                           fmaxf(fmaxf(fabsf(v3.x), fabsf(v3.y)), fmaxf(fabsf(v3.z), fabsf(v3.w)))));
    }
    for (; i < n4; i += stride) {
        const float4 v = __ldg(x4 + i);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    for (long long t = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) m = fmaxf(m, fabsf(__ldg(x + t)));
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(reinterpret_cast<unsigned int*>(amax), __float_as_uint(m));
}

}  // namespace bnn

using namespace bnn;

extern "C" size_t bnn_stem_tc_weight_bytes(void) { return TC_B_BYTES; }

extern "C" int bnn_stem_tc_pack_weight(const float* w, int32_t w_log2_scale, void* w_ops, void* stream_) {
    if (!w || !w_ops) return BNN_E_NULL;
    if (w_log2_scale < -60 || w_log2_scale > 60) return BNN_E_SHAPE;
    if ((uintptr_t)w_ops & 15) return BNN_E_ALIGN;
    const int total = TC_B_BYTES / 2;
    stem_tc_pack_weight_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(w, ldexpf(1.0f, w_log2_scale), (__half*)w_ops);
    count_launch(1);
    return (int)cudaGetLastError();
}

extern "C" int bnn_amax_f32(const float* x, int64_t count, float* amax, void* stream_) {
    if (!x || !amax) return BNN_E_NULL;
    if (count <= 0) return BNN_E_SHAPE;
    if ((uintptr_t)x & 15) return BNN_E_ALIGN;
    cudaStream_t stream = (cudaStream_t)stream_;
    cudaError_t ce = cudaMemsetAsync(amax, 0, sizeof(float), stream);
    if (ce != cudaSuccess) return (int)ce;
    const long long want = (count / 16 + 255) / 256;
    const int blocks = (int)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
    launch_pdl(amax_kernel, dim3(blocks), dim3(256), 0, stream, x, (long long)count, amax);
    count_launch(1);
    return (int)cudaGetLastError();
}

extern "C" int bnn_stem_tc_run(const bnn_stem_tc_params* p, uint32_t flags, void* stream_) {
    if (!p) return BNN_E_NULL;
    if (!p->x || !p->w_ops || !p->bn_scale || !p->bn_shift || !p->out) return BNN_E_NULL;
    if ((p->nx_scale == nullptr) != (p->nx_shift == nullptr) || (p->nx2_scale == nullptr) != (p->nx2_shift == nullptr)) return BNN_E_NULL;
    if (p->nx2_scale && !p->out_bits2) return BNN_E_NULL;
    if (p->n <= 0 || p->h < 7 || p->w < 7) return BNN_E_SHAPE;
    if (p->x_dtype != 0 && p->x_dtype != 1) return BNN_E_SHAPE;
    if ((long long)p->h * p->w * 3 >= 0x7fffffffLL) return BNN_E_UNSUPPORTED;
    if (p->x_log2_scale < -60 || p->x_log2_scale > 60 || p->w_log2_scale < -60 || p->w_log2_scale > 60) return BNN_E_SHAPE;
    if (((uintptr_t)p->w_ops & 15) || ((uintptr_t)p->out_bits & 15) || ((uintptr_t)p->out_bits2 & 15)) return BNN_E_ALIGN;
    if (p->x_dtype == 1 && (p->w & 1)) return BNN_E_UNSUPPORTED;            // uint8 rows are read as 16-bit pairs
    if (p->x_dtype == 1 && ((uintptr_t)p->x & 1)) return BNN_E_ALIGN;
    StemTcArgs a{};
    a.x = p->x; a.wops = p->w_ops; a.bn_scale = p->bn_scale; a.bn_shift = p->bn_shift;
    a.nx_scale = p->nx_scale; a.nx_shift = p->nx_shift; a.nx2_scale = p->nx2_scale; a.nx2_shift = p->nx2_shift;
    a.nx_relu = p->nx_relu != 0; a.nx2_relu = p->nx2_relu != 0;
    a.x_amax = p->x_dtype == 0 ? p->x_amax : nullptr; a.out = p->out; a.obits = (uint4*)p->out_bits; a.obits2 = (uint4*)p->out_bits2;
    for (int i = 0; i < 3; ++i) { a.u8_mean[i] = p->u8_mean[i]; a.u8_istd[i] = p->u8_istd[i]; }
    a.x_log2_scale = p->x_log2_scale; a.w_log2_scale = p->w_log2_scale;
    a.dbg = (int)((flags >> 8) & 31u);
    a.vec2 = (p->x_dtype == 0 && (p->w & 1) == 0 && ((uintptr_t)p->x & 7) == 0) ? 1 : 0;
    a.N = p->n; a.H = p->h; a.W = p->w;
    a.Hc = (p->h + 6 - 7) / 2 + 1; a.Wc = (p->w + 6 - 7) / 2 + 1;
    const bool pool = p->pool != 0;
    a.Ho = pool ? (a.Hc + 2 - 3) / 2 + 1 : a.Hc;
    a.Wo = pool ? (a.Wc + 2 - 3) / 2 + 1 : a.Wc;
    const int maxt = pool ? TC_MAX_PT : TC_MAX_CT;
    a.tiles_w = (a.Wo + maxt - 1) / maxt;
    a.TW = (a.Wo + a.tiles_w - 1) / a.tiles_w;
    a.total_rows = (long long)p->n * a.tiles_w * a.Ho;
    int dev = 0, sms = 148;
    cudaError_t ce = cudaGetDevice(&dev);
    if (ce != cudaSuccess) return (int)ce;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const void* fn = pool ? (p->x_dtype ? (const void*)stem_tc_kernel<true, 1> : (const void*)stem_tc_kernel<true, 0>)
                          : (p->x_dtype ? (const void*)stem_tc_kernel<false, 1> : (const void*)stem_tc_kernel<false, 0>);
    ce = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM);
    if (ce != cudaSuccess) return (int)ce;
    const unsigned ctas = (unsigned)(a.total_rows < sms ? a.total_rows : sms);
    cudaStream_t st = (cudaStream_t)stream_;
    if (pool) {
        if (p->x_dtype) launch_pdl(stem_tc_kernel<true, 1>, dim3(ctas), dim3(TC_THREADS), TC_SMEM, st, a);
        else launch_pdl(stem_tc_kernel<true, 0>, dim3(ctas), dim3(TC_THREADS), TC_SMEM, st, a);
    } else {
        if (p->x_dtype) launch_pdl(stem_tc_kernel<false, 1>, dim3(ctas), dim3(TC_THREADS), TC_SMEM, st, a);
        else launch_pdl(stem_tc_kernel<false, 0>, dim3(ctas), dim3(TC_THREADS), TC_SMEM, st, a);
    }
    count_launch(1);
    return (int)cudaGetLastError();
}

extern "C" int bnn_stem_tc_fwd(const float* x, int32_t n, int32_t h, int32_t w, const void* w_ops, int32_t x_log2_scale,
                               const float* x_amax, int32_t w_log2_scale, const float* bn_scale, const float* bn_shift,
                               const float* nx_scale, const float* nx_shift, float* out, void* out_bits, uint32_t flags,
                               void* stream_) {
    bnn_stem_tc_params p{};
    p.x = x; p.x_dtype = 0; p.n = n; p.h = h; p.w = w; p.w_ops = w_ops; p.w_log2_scale = w_log2_scale;
    p.x_log2_scale = x_log2_scale; p.x_amax = x_amax; p.bn_scale = bn_scale; p.bn_shift = bn_shift; p.pool = 1;
    p.nx_scale = nx_scale; p.nx_shift = nx_shift; p.out_bits = out_bits; p.out = out;
    return bnn_stem_tc_run(&p, flags, stream_);
}
