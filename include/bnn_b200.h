/*
 * bnn_b200.h -- C ABI of the B200-native binary-convolution forward path.
 *
 * This is the drop-in boundary for the hot path of 1adrianb/binary-networks-pytorch
 * (`bnn` 0.1.2): everything a binarized `bnn.layers.Conv2d` / `Linear` does in
 * `forward` (reference bnn/layers/conv.py:90-97, bnn/layers/linear.py:22-27) is
 * reachable through the five compute entry points below.  The reference has no
 * FFI of its own (it is pure Python over torch); the binding a maintainer adds
 * is the ctypes stub shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory owned by the
 *     caller unless stated otherwise; no allocation, no hidden global state
 *     besides a lazily resolved driver entry point, no device synchronisation;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - return value: 0 = ok, >0 = cudaError_t, <0 = BNN_E_* argument error;
 *     `bnn_strerror` renders either;
 *   - re-entrant across streams and devices (one call = current device, one stream).
 *
 * Packed layouts (shared with oracle/bnn_oracle.c)
 *   activation planes  abits[n][chunk][h][w]  = {s_lo, s_hi, m_lo, m_hi}  4 x u32 (16 B)
 *       chunk = 64 input channels; s bit <=> x > 0; m bit <=> x != 0 (false for +-0, NaN):
 *       the reference's sign() is ternary (bnn/ops.py:66), and zero padding is applied
 *       after binarisation (bnn/layers/conv.py:91-92) -- padded taps simply have m = 0.
 *   weight planes      wbits[c_out/32][kstep][c_out%32][2]  u32,
 *       kstep = (chunk*kh + i)*kw + j, word 0/1 = channels chunk*64 + 0..31 / 32..63,
 *       bit <=> centred weight > 0 (bnn/ops.py:130-136)
 *   dot[n,co,ho,wo] = sum_k popc(m) - 2 * sum_k popc(m & (s ^ t))     (k over the receptive field)
 *   y = (scale[co]*dot + bias[co]) * post[co]          (conv.py:92-97, ops.py:136,202)
 */
#ifndef BNN_B200_H
#define BNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BNN_B200_ABI_VERSION 4

/* argument errors (negative); positive codes are cudaError_t */
#define BNN_E_NULL        (-1)  /* required pointer is NULL                         */
#define BNN_E_SHAPE       (-2)  /* non-positive / inconsistent dimension            */
#define BNN_E_UNSUPPORTED (-3)  /* geometry does not fit this build (smem, box)      */
#define BNN_E_DRIVER      (-4)  /* cuTensorMapEncodeTiled unavailable or failed      */
#define BNN_E_ALIGN       (-5)  /* pointer not 16-byte aligned                       */

/* Geometry of one binarized convolution; mirrors torch.nn.Conv2d's attributes
 * read by bnn.layers.Conv2d.from_module (bnn/layers/conv.py:107-110); groups == 1. */
typedef struct bnn_conv_geom {
    int32_t n, c_in, h, w;        /* input  [n, c_in, h, w]                 */
    int32_t c_out, kh, kw;        /* weight [c_out, c_in, kh, kw]           */
    int32_t stride_h, stride_w;
    int32_t pad_h, pad_w;         /* zero padding, applied AFTER sign()     */
    int32_t dil_h, dil_w;
} bnn_conv_geom;

/* bnn_query selectors */
#define BNN_Q_ABI_VERSION   0
#define BNN_Q_SM_ARCH       1   /* 100 => compiled for sm_100a                */
#define BNN_Q_DEVICE_SMS    2   /* multiprocessor count of the current device */
#define BNN_Q_LAUNCH_COUNT  3   /* kernels launched by this library so far (process-wide) */

/* flags for bnn_bconv2d_fwd / bnn_blinear_fwd */
#define BNN_F_STAGE_LDG   1u    /* stage tiles with plain loads instead of TMA (debug / A-B) */
#define BNN_F_NO_CSA      2u    /* force the one-POPC-per-word inner loop                     */
#define BNN_F_STEM_TMA_IN 4u    /* bnn_stem_fwd: stage the input window with a 4-D TMA tensor load (experimental) */
#define BNN_F_STEM_NO_BULK 8u   /* bnn_stem_fwd: stage the weights with plain loads instead of a TMA bulk copy   */
#define BNN_F_STEM_MMA_CHAIN 16u /* bnn_stem_mma_fwd: accumulate every product inside the tensor core (A-B experiment) */

int         bnn_query(int what, int64_t *value);
const char *bnn_strerror(int code);

/* sizes (bytes) of the packed buffers the caller must allocate */
size_t bnn_act_bits_bytes(int32_t n, int32_t c, int32_t h, int32_t w);
size_t bnn_weight_bits_bytes(int32_t c_out, int32_t c_in, int32_t kh, int32_t kw);

/*
 * BasicInputBinarizer / SignActivation.forward (bnn/ops.py:151-152, 63-66) as a
 * bit-pack: fp32 activations, addressed with ELEMENT strides (so NCHW,
 * channels_last and Linear's [rows,in] viewed as n=1,h=1,w=rows all fit),
 * -> abits (16-byte aligned).  pre_scale/pre_shift ([c], both or neither) fold an
 * eval-mode BatchNorm that sits in front of the layer: sign(x*pre_scale + pre_shift);
 * pre_relu != 0 additionally puts a ReLU between that BatchNorm and the sign (bn -> relu -> conv,
 * bnn/models/layers/hierarchical_block.py:41-43): negative values become 0, i.e. m = s.
 */
int bnn_pack_act_f32(const float *x, int64_t stride_n, int64_t stride_c, int64_t stride_h,
                     int64_t stride_w, int32_t n, int32_t c, int32_t h, int32_t w,
                     const float *pre_scale, const float *pre_shift, int32_t pre_relu,
                     void *abits, void *stream);

/*
 * AvgPool2d(kernel = stride = k, ceil_mode, count_include_pad = False) followed by the
 * bit-pack above, in one pass: the shortcut branch of bnn.models.resnet
 * (bnn/models/resnet.py:129-133) feeds a binarized 1x1 conv from an average-pooled map.
 * Output planes have ceil(h/k) x ceil(w/k) pixels (ceil_mode) or floor (otherwise).
 */
int bnn_avgpool_pack_f32(const float *x, int64_t stride_n, int64_t stride_c, int64_t stride_h,
                         int64_t stride_w, int32_t n, int32_t c, int32_t h, int32_t w,
                         int32_t k, int32_t ceil_mode, const float *pre_scale,
                         const float *pre_shift, int32_t pre_relu, void *abits, void *stream);

/*
 * nn.AvgPool2d(2) (kernel = stride = 2, floor mode) of a DENSE channels-last tensor x[n, h, w, c] fused with the bit-pack
 * above, in one pass over x: pooled_out (may be NULL) receives the pooled fp32 tensor [n, h/2, w/2, c] (channels-last,
 * the residual stream of the next block), abits the planes of sign(pooled * pre_scale + pre_shift) (pre_relu as above).
 * Same operation order as bnn_avgpool_pack_f32.  c in {64, 128, 256, 512}, else BNN_E_UNSUPPORTED.
 */
int bnn_avgpool2_pack_cl_f32(const float *x, int32_t n, int32_t c, int32_t h, int32_t w, const float *pre_scale,
                             const float *pre_shift, int32_t pre_relu, float *pooled_out, void *abits, void *stream);

/*
 * XNORWeightBinarizer.forward (bnn/ops.py:129-140) as a prepare-time pack:
 * optional centring over c_in (ops.py:130-132), alpha[co] = mean |w| after
 * centring (ops.py:116-127; 1.0 if !compute_alpha), sign bits -> wbits.
 * w is contiguous [c_out, c_in, kh, kw] (Linear: kh = kw = 1).
 * n_zero (device int32, may be NULL) receives the number of centred weights
 * that are exactly 0 -- their sign is 0 in the reference, which one bit cannot
 * hold; when it is non-zero the layer has to run the ternary form below.
 */
int bnn_pack_weight_f32(const float *w, int32_t c_out, int32_t c_in, int32_t kh, int32_t kw,
                        int32_t center_weights, int32_t compute_alpha,
                        void *wbits, float *alpha, int32_t *n_zero, void *stream);

/*
 * Ternary weights.  The reference evaluates sign(0) = 0 for weights too (bnn/ops.py:66,136): an exactly-zero (centred)
 * weight contributes nothing.  bnn_pack_weight_ternary_f32 is bnn_pack_weight_f32 plus `wzero` (same layout and size as
 * wbits, may be NULL): the plane of those exactly-zero weights.  With  t- = wbits  (zeros packed as -1) and
 * t+ = wbits | wzero  (zeros packed as +1) the ternary dot is the mean of two binary dots,
 *     dot = (dot(t-) + dot(t+)) / 2      (exact: the sum is always even),
 * which bnn_bconv2d_partial_fwd (twice) + bnn_dot_finish_f32(divisor = 2) compute.
 */
int bnn_pack_weight_ternary_f32(const float *w, int32_t c_out, int32_t c_in, int32_t kh, int32_t kw,
                                int32_t center_weights, int32_t compute_alpha, void *wbits, void *wzero,
                                float *alpha, int32_t *n_zero, void *stream);

/*
 * Split-K.  bnn_bconv2d_fwd keeps a CTA's whole reduction (all input chunks of the kernel window) in shared memory; a
 * layer that cannot (Linear(25088, 4096): 392 chunks; convolutions over more than 16384 channels) is contracted chunk
 * range by chunk range.  bnn_conv_split says how: *nparts == 1 means the plain entry points work; otherwise the layer
 * is cut into *nparts ranges of *chunks_per_part 64-channel chunks (the last one may be shorter).
 * bnn_bconv2d_partial_fwd contracts input chunks [chunk0, chunk0 + nchunks) of the FULL packed tensors abits / wbits
 * (geom describes the full layer) and writes the integer dots of that range, as exact fp32 values, to
 * part[n, c_out, ho, wo] (contiguous).  bnn_dot_finish_f32 adds `nparts` such buffers (parts = [nparts][n*c_out*ho*wo]),
 * divides by `divisor` (1, or 2 for the two launches of a ternary weight tensor) and applies the reference epilogue
 *     y = (scale[co] * dot + bias[co]) * post[co]
 * with element strides for the output, like bnn_bconv2d_fwd.
 */
int bnn_conv_split(const bnn_conv_geom *geom, uint32_t flags, int32_t *chunks_per_part, int32_t *nparts);
int bnn_bconv2d_partial_fwd(const void *abits, const void *wbits, const bnn_conv_geom *geom, int32_t chunk0,
                            int32_t nchunks, float *part, uint32_t flags, void *stream);
int bnn_dot_finish_f32(const float *parts, int32_t nparts, int32_t divisor, const float *scale, const float *bias,
                       const float *post, float *out, int64_t ostride_n, int64_t ostride_c, int64_t ostride_h,
                       int64_t ostride_w, int32_t n, int32_t c_out, int32_t ho, int32_t wo, void *stream);

/*
 * bnn.layers.Conv2d.forward (bnn/layers/conv.py:90-97) on packed operands:
 * XNOR/AND + popcount contraction with the fused epilogue
 *     y = (scale[co] * dot + bias[co]) * post[co]
 * scale / bias / post may each be NULL (1, 0, 1).  out is fp32, written with
 * element strides (NCHW contiguous: c_out*ho*wo, ho*wo, wo, 1).
 */
int bnn_bconv2d_fwd(const void *abits, const void *wbits,
                    const float *scale, const float *bias, const float *post,
                    float *out, int64_t ostride_n, int64_t ostride_c, int64_t ostride_h,
                    int64_t ostride_w, const bnn_conv_geom *geom, uint32_t flags, void *stream);

/*
 * Cross-module fusion: the modules that follow a binarized conv
 * inside the reference's blocks (bnn/models/layers/res_block.py:40-56,152-167,
 * hierarchical_block.py:38-60) folded into the same kernel.  With y as above:
 *     z  = y * bn_scale[co] + bn_shift[co]                (eval BatchNorm; skipped if NULL)
 *     z += residual                 if residual && !residual_after_act
 *     v  = act(z)                   0 none, 1 ReLU, 2 PReLU(act_slope[co])
 *     v += residual                 if residual &&  residual_after_act
 *     out      <- v                 fp32, element strides; skipped if out == NULL
 *     out_bits <- planes of sign(v * nx_scale[co] + nx_shift[co])   (next layer's input, in the
 *                 abits layout for [n, c_out, ho, wo]; nx_* NULL = identity; skipped if NULL)
 *                 nx_relu: a ReLU sits between that BatchNorm and the sign (m = s);
 *                 bits_before_residual: take v before the after-activation residual add (HBlock:
 *                 the next conv sees the conv output, the block output adds the shortcut)
 * In this fused mode the per-channel constants are folded once per launch --
 *     k0 = scale*post*bn_scale,  k1 = (bias*post)*bn_scale + bn_shift,  z = fma(k0, dot, k1)
 * -- the next-layer affine is one fma, and ReLU is fmaxf(z, 0); oracle/bnn_oracle.c restates it
 * operation for operation.  bnn_bconv2d_fwd (no fusion field set) keeps the reference's exact order.
 */
typedef struct bnn_epilogue {
    const float *scale, *bias, *post;
    const float *bn_scale, *bn_shift;
    const float *residual;
    int64_t rstride_n, rstride_c, rstride_h, rstride_w;
    int32_t residual_after_act;
    int32_t act;
    const float *act_slope;
    float *out;
    int64_t ostride_n, ostride_c, ostride_h, ostride_w;
    void *out_bits;
    const float *nx_scale, *nx_shift;
    int32_t nx_relu;
    int32_t bits_before_residual;
} bnn_epilogue;

#define BNN_ACT_NONE  0
#define BNN_ACT_RELU  1
#define BNN_ACT_PRELU 2

int bnn_bconv2d_fused_fwd(const void *abits, const void *wbits, const bnn_conv_geom *geom,
                          const bnn_epilogue *epilogue, uint32_t flags, void *stream);

/*
 * Autotune: time the cost model's best `top_k` tile plans for this geometry / epilogue kind with real
 * launches on the caller's buffers (same arguments as bnn_bconv2d_fused_fwd; results are the same for
 * every plan, so the buffers end up holding the correct output), and remember the fastest in a
 * process-wide cache that later launches of the same geometry use.  Synchronises `stream`; call it at
 * warm-up, never inside a CUDA-graph capture.  A geometry that is already tuned returns immediately.
 */
int bnn_bconv2d_tune(const void *abits, const void *wbits, const bnn_conv_geom *geom,
                     const bnn_epilogue *epilogue, uint32_t flags, int32_t top_k, void *stream);

/*
 * Introspection: the tile plan the library would use for a geometry (host only, no GPU needed).
 * plan[12] = {P pixels/group, C 32-channel blocks/lane, kw instance, stride instance, carry-save mode,
 *             TH, TW, warps per CTA, pixel units (grid.x), channel tiles (grid.y), smem bytes, groups/unit}.
 * sms <= 0 assumes 148.
 */
int bnn_conv_plan(const bnn_conv_geom *geom, uint32_t flags, int32_t sms, int32_t *plan);

/*
 * Test / tuning hooks for the tile planner.  bnn_conv_plan_list enumerates every feasible tile plan of a geometry
 * (rows of 12 ints in bnn_conv_plan's format, cost-model order, at most `cap` rows written, *count = how many exist).
 * bnn_bconv2d_fused_fwd_plan is bnn_bconv2d_fused_fwd with the plan forced to the candidate whose (P, C, TH, warps)
 * match (TH <= 0 / warps <= 0: best-ranked candidate of that (P, C) family); BNN_E_UNSUPPORTED if there is none.
 * bnn_conv_instance reports which kernel instance that launch runs: inst[6] = {P, C, kw instance, stride instance,
 * carry-save mode, epilogue instance 0..4}.  The parity suite sweeps every instance with these (tests/test_gpu_plans.py).
 */
int bnn_conv_plan_list(const bnn_conv_geom *geom, uint32_t flags, int32_t *plans, int32_t cap, int32_t *count);
int bnn_bconv2d_fused_fwd_plan(const void *abits, const void *wbits, const bnn_conv_geom *geom,
                               const bnn_epilogue *epilogue, uint32_t flags, int32_t P, int32_t C, int32_t TH,
                               int32_t warps, void *stream);
int bnn_conv_instance(const bnn_conv_geom *geom, const bnn_epilogue *epilogue, uint32_t flags, int32_t P, int32_t C,
                      int32_t TH, int32_t warps, int32_t *inst);

/*
 * bnn.layers.Linear.forward (bnn/layers/linear.py:22-27): rows x in_features
 * packed as n=1,h=1,w=rows (bnn_pack_act_f32 with stride_w = in_features,
 * stride_c = 1), weight packed with kh = kw = 1; out is [rows, out_features].
 */
int bnn_blinear_fwd(const void *abits, const void *wbits,
                    const float *scale, const float *bias, const float *post,
                    float *out, int32_t rows, int32_t in_features, int32_t out_features,
                    uint32_t flags, void *stream);

/*
 * The fp32 stem of bnn.models.resnet (bnn/models/resnet.py:85-92,147-153) as one kernel:
 *   conv 7x7 / 2 / pad 3 (3 -> 64 channels) -> eval BatchNorm (folded: v*bn_scale + bn_shift)
 *   -> ReLU -> MaxPool 3x3 / 2 / pad 1.
 * x: [n,3,h,w] contiguous fp32.  w_t: the conv weight [64,3,7,7] repacked to [3][7][7][32][2] with
 * w_t[ci][kh][kw][l][b] = w[b*32 + l][ci][kh][kw] (16-byte aligned): a lane's two channels are one 8-byte load.
 * out: [n,hp,wp,64] fp32 (NHWC = torch channels_last), (hp,wp) from bnn_stem_out_hw.
 * out_bits (may be NULL): planes of sign(out*nx_scale + nx_shift) in the abits layout [n][1][hp][wp]
 * for the first binarized conv (nx_* NULL = identity).  The conv accumulates one fma chain per output
 * in (c_in, kh, kw) order; oracle/bnn_oracle.c restates the same order.
 */
int bnn_stem_out_hw(int32_t h, int32_t w, int32_t *hp, int32_t *wp);
int bnn_stem_fwd(const float *x, int32_t n, int32_t h, int32_t w, const float *w_t,
                 const float *bn_scale, const float *bn_shift, const float *nx_scale,
                 const float *nx_shift, float *out, void *out_bits, uint32_t flags, void *stream);

/*
 * The same stem on the legacy tensor path (mma.sync m16n8k16 f16, fp32 accumulate) with fp32-level accuracy:
 * both operands are split into (hi, lo) fp16 pairs of x*2^x_log2_scale and w*2^w_log2_scale (22 significand bits
 * each), the products hi*hi + hi*lo + lo*hi are exact in fp32, and each 16-tap partial sum is added to the running
 * sum with a round-to-nearest fp32 add -- total error at the level of an fp32 fma chain over the 147 taps, but NOT
 * bit-identical to bnn_stem_fwd (different summation order).  Operands must stay inside the fp16 range:
 * |x| * 2^x_log2_scale < 65504 and |w| * 2^w_log2_scale < 65504 (choose w_log2_scale from max|w|).
 * w_frag: bnn_stem_mma_weight_bytes() bytes, 16-byte aligned, written by bnn_stem_mma_pack_weight from the plain
 * [64,3,7,7] fp32 conv weight (mma B fragments in k-step / n-tile / lane order).  bn_scale / bn_shift 8-byte aligned.
 * Everything else as bnn_stem_fwd.
 */
size_t bnn_stem_mma_weight_bytes(void);
int bnn_stem_mma_pack_weight(const float *w, int32_t w_log2_scale, void *w_frag, void *stream);
int bnn_stem_mma_fwd(const float *x, int32_t n, int32_t h, int32_t w, const void *w_frag,
                     int32_t x_log2_scale, const float *x_amax, int32_t w_log2_scale, const float *bn_scale,
                     const float *bn_shift, const float *nx_scale, const float *nx_shift, float *out, void *out_bits,
                     uint32_t flags, void *stream);

/*
 * Guarded input range for the split-fp16 stems: x_amax (device scalar, may be NULL) holds max|x| of the batch, written by
 * bnn_amax_f32 on the same stream (one extra pass over the input, ~25 us for 154 MB).  When it is given the kernels pick
 * the power-of-two input scale themselves (max|x| * 2^sx in [2^14, 2^15)), so no finite input can overflow fp16 and
 * x_log2_scale is ignored; with x_amax == NULL the caller guarantees |x| * 2^x_log2_scale < 65504.
 * bnn_amax_f32: x 16-byte aligned, `count` elements; *amax = max |x[i]| (NaN ignored); no host synchronisation.
 */
int bnn_amax_f32(const float *x, int64_t count, float *amax, void *stream);

/*
 * The same stem on the 5th-generation tensor cores (tcgen05.mma, accumulators in tensor memory; csrc/stem_tc.cu): same
 * contract, same split-fp16 arithmetic and accuracy class as bnn_stem_mma_fwd, NOT bit-identical to it (the tensor core
 * sums the 147 taps of an output in one accumulator chain).  Implicit GEMM per conv row, the im2col view expressed by
 * the shared-memory matrix descriptor itself (no A tile is materialised), persistent warp-specialised CTAs.
 * w_ops: bnn_stem_tc_weight_bytes() bytes, 16-byte aligned, written by bnn_stem_tc_pack_weight from the plain
 * [64,3,7,7] fp32 conv weight.  x_amax as above.
 */
size_t bnn_stem_tc_weight_bytes(void);
int bnn_stem_tc_pack_weight(const float *w, int32_t w_log2_scale, void *w_ops, void *stream);

/*
 * bnn_stem_tc_run: every form of the tcgen05 stem, described by one struct (bnn_stem_tc_fwd is the fp32 / max-pool form
 * with positional arguments).
 *   x_dtype 0: x = fp32 [n,3,h,w] contiguous.
 *   x_dtype 1: x = uint8 [n,h,w,3] (decoded image bytes, w even); the kernel normalises while it stages the window,
 *              x' = (float(x) - u8_mean[c]) * u8_istd[c]  (two rounded fp32 operations, i.e. exactly
 *              `(x.float() - mean) * istd` in torch), then proceeds as for fp32 -- a quarter of the upload bytes.
 *              x_amax is ignored: choose x_log2_scale from max_c max(|0 - mean|, |255 - mean|) * istd.
 *   pool 1:    conv7x7/2 -> BatchNorm -> ReLU -> MaxPool 3x3/2/1   (bnn/models/resnet.py:85-92,147-153); out [n,hp,wp,64]
 *   pool 0:    conv7x7/2 -> BatchNorm -> ReLU                      (the Hierarchical-Block harness stem); out [n,hc,wc,64]
 *   out_bits  (may be NULL): planes of sign(out * nx_scale + nx_shift), nx_* NULL = identity; nx_relu: a ReLU sits in
 *              front of the sign (m = s).  out_bits2 / nx2_*: a second set of planes from the same tensor (the HBlock
 *              harness feeds its first block's conv1 and the shortcut conv from two different BatchNorms).
 */
typedef struct bnn_stem_tc_params {
    const void *x;
    int32_t x_dtype, n, h, w;
    const void *w_ops;
    int32_t w_log2_scale, x_log2_scale;
    const float *x_amax;
    float u8_mean[3], u8_istd[3];
    const float *bn_scale, *bn_shift;
    int32_t pool;
    const float *nx_scale, *nx_shift;
    int32_t nx_relu;
    void *out_bits;
    const float *nx2_scale, *nx2_shift;
    int32_t nx2_relu;
    void *out_bits2;
    float *out;
} bnn_stem_tc_params;
int bnn_stem_tc_run(const bnn_stem_tc_params *params, uint32_t flags, void *stream);
#define BNN_F_STEM_TC_DBG_SHIFT 8   /* bits 8..11 of flags idle one role of the kernel each (timing experiments; garbage results) */
int bnn_stem_tc_fwd(const float *x, int32_t n, int32_t h, int32_t w, const void *w_ops, int32_t x_log2_scale,
                    const float *x_amax, int32_t w_log2_scale, const float *bn_scale, const float *bn_shift,
                    const float *nx_scale, const float *nx_shift, float *out, void *out_bits, uint32_t flags,
                    void *stream);

/*
 * The down-sampling shortcut of bnn.models.resnet (bnn/models/resnet.py:129-133) as one kernel:
 *   AvgPool2d(pool, stride pool, ceil_mode, count_include_pad=False) -> sign() -> binarized conv 1x1 -> eval BatchNorm,
 * i.e. bnn_avgpool_pack_f32 followed by bnn_bconv2d_fused_fwd (1x1, stride 1, no padding, epilogue with scale / bias /
 * post / bn_* only), bit-identical to that pair but without the round trip of the planes through HBM.
 * x: channels-last fp32, element (n, c, h, w) at n*xs_n + h*xs_h + w*xs_w + c (channel stride 1).
 * wbits: bnn_pack_weight_f32 output for the [c_out, c_in, 1, 1] weight.  scale (= alpha), bias, post, bn_* may be NULL.
 * out: [n, ho, wo, c_out] contiguous fp32 (torch channels_last), ho = ceil or floor of h / pool.
 */
int bnn_shortcut_fwd(const float *x, int64_t xs_n, int64_t xs_h, int64_t xs_w, int32_t n, int32_t c_in,
                     int32_t h, int32_t w, int32_t pool, int32_t ceil_mode, const void *wbits, int32_t c_out,
                     const float *scale, const float *bias, const float *post, const float *bn_scale,
                     const float *bn_shift, float *out, uint32_t flags, void *stream);

/*
 * Integer-pipe micro-benchmarks used for the popcount roofline denominator
 * (bench.py): runs `which` on every SM and returns achieved giga-operations/s
 * (warp-lane operations) in *gops.  Synchronises the device.  which:
 *   0 POPC only   1 LOP3 only   2 LOP3+POPC+IADD (one word per POPC)
 *   3 3:2 carry-save (5 LOP3 + 2 POPC per 3 words)   4 7:3 carry-save
 * For 2..4 the figure is 32-bit XNOR-popcount WORDS per second.
 * Floating-point rates behind the stem kernels, in giga multiply-adds / s:
 *   5 mma.sync m16n8k8 tf32   6 mma.sync m16n8k16 f16   7 mma.sync m16n8k16 bf16   8 fma.rn.f32x2   9 fma.rn.f32
 */
int bnn_ubench(int32_t which, int32_t iters, double *gops);

#ifdef __cplusplus
}
#endif
#endif /* BNN_B200_H */
