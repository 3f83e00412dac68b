/*
 * oracle/bnn_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's binarized Conv2d / Linear forward
 * path (1adrianb/binary-networks-pytorch, `bnn` 0.1.2).  Nothing outside
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * this file; the shipped CUDA path never touches it.
 *
 * Two independent restatements live here:
 *
 *  (A) orc_floatsim_*  -- the reference's own arithmetic, in float:
 *        sign(x)                         bnn/ops.py:63-66   (sign(+-0)=0, sign(nan)=0)
 *        W - mean_{C_in}(W)              bnn/ops.py:130-132 (center_weights)
 *        alpha = mean|W| per c_out       bnn/ops.py:116-127 (after centering)
 *        sign(W) * alpha                 bnn/ops.py:136
 *        conv / linear, zero padding of the *binarized* input, + bias
 *                                        bnn/layers/conv.py:91-92, linear.py:24-25
 *        out *= alpha_post               bnn/ops.py:200-202
 *      The contraction itself is torch's F.conv2d in the reference (third
 *      party, ATen/oneDNN); here it is a plain fp32 accumulation loop.
 *
 *  (B) orc_pack_* + orc_bconv2d[_fused] -- the integer formulation the CUDA kernels
 *      implement, on exactly the packed layouts of
 *      include/bnn_b200.h:  dot = popc(m) - 2*popc(m & (s ^ t)),
 *      y = (alpha_w*dot + bias) * alpha_post, plus the cross-module fusion epilogue
 *      (eval BatchNorm, residual, ReLU/PReLU, emission of the next layer's planes).
 *
 * Parity pin: tests/test_oracle.py checks (A) against the reference's own
 * golden vectors (test/test_layers.py:22-66, test/test_binarize.py:118-120)
 * and against outputs of the real reference generated in the build container
 * (tests/golden/make_golden.py -> tests/golden/*.npz); (B) is checked against (A).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t n, c_in, h, w, c_out, kh, kw;
    int32_t stride_h, stride_w, pad_h, pad_w, dil_h, dil_w;
} orc_geom;

static int out_dim(int in, int k, int s, int p, int d) {
    return (in + 2 * p - d * (k - 1) - 1) / s + 1;
}

int orc_out_h(const orc_geom *g) { return out_dim(g->h, g->kh, g->stride_h, g->pad_h, g->dil_h); }
int orc_out_w(const orc_geom *g) { return out_dim(g->w, g->kw, g->stride_w, g->pad_w, g->dil_w); }

/* bnn/ops.py:66 -- torch.sign: -1, 0, +1; zero for +-0 and NaN. */
static float sgn(float v) { return (float)((v > 0.0f) - (v < 0.0f)); }

void orc_sign_f32(const float *x, float *y, int64_t count) {
    for (int64_t i = 0; i < count; ++i) y[i] = sgn(x[i]);
}

/*
 * XNORWeightBinarizer.forward, bnn/ops.py:129-140, for a weight of shape
 * [c_out, c_in, kh, kw] (Linear: kh=kw=1; Conv1d: kh=1).
 *   centered[co][ci][t] = w - mean_ci(w[co][:][t])      (ops.py:130-132)
 *   alpha[co] = sum |centered[co]| / (c_in*kh*kw)        (ops.py:116-123)
 * Sums are taken in double and rounded once to float (torch reduces in fp32
 * with a blocked order; both are within 1 ulp-ish of the exact value).
 */
void orc_weight_center_alpha(const float *w, int c_out, int c_in, int kh, int kw,
                             int center, int compute_alpha,
                             float *centered, float *alpha) {
    const int taps = kh * kw;
    const int64_t per_out = (int64_t)c_in * taps;
    for (int co = 0; co < c_out; ++co) {
        const float *wc = w + co * per_out;
        float *cc = centered + co * per_out;
        for (int t = 0; t < taps; ++t) {
            float mean = 0.0f;
            if (center) {
                double s = 0.0;
                for (int ci = 0; ci < c_in; ++ci) s += (double)wc[(int64_t)ci * taps + t];
                mean = (float)(s / (double)c_in);
            }
            for (int ci = 0; ci < c_in; ++ci)
                cc[(int64_t)ci * taps + t] = center ? wc[(int64_t)ci * taps + t] - mean
                                                   : wc[(int64_t)ci * taps + t];
        }
        if (compute_alpha) {
            double s = 0.0;
            for (int64_t i = 0; i < per_out; ++i) s += fabs((double)cc[i]);
            alpha[co] = (float)(s / (double)per_out);
        } else {
            alpha[co] = 1.0f;
        }
    }
}

/*
 * (A) float simulation of bnn.layers.Conv2d.forward (bnn/layers/conv.py:90-97)
 * with BasicInputBinarizer / XNORWeightBinarizer / optional BasicScaleBinarizer.
 * x: [n,c_in,h,w] contiguous; w: [c_out,c_in,kh,kw]; bias/post may be NULL;
 * out: [n,c_out,ho,wo].
 */
int orc_floatsim_conv2d(const float *x, const float *w, const float *bias, const float *post,
                        const orc_geom *g, int center, int compute_alpha, float *out) {
    const int ho_n = orc_out_h(g), wo_n = orc_out_w(g);
    const int taps = g->kh * g->kw;
    const int64_t per_out = (int64_t)g->c_in * taps;
    float *centered = (float *)malloc(sizeof(float) * per_out * g->c_out);
    float *alpha = (float *)malloc(sizeof(float) * g->c_out);
    float *wb = (float *)malloc(sizeof(float) * per_out * g->c_out);
    float *xb = (float *)malloc(sizeof(float) * (int64_t)g->n * g->c_in * g->h * g->w);
    if (!centered || !alpha || !wb || !xb) return -1;
    orc_weight_center_alpha(w, g->c_out, g->c_in, g->kh, g->kw, center, compute_alpha, centered, alpha);
    for (int co = 0; co < g->c_out; ++co)
        for (int64_t i = 0; i < per_out; ++i)
            wb[co * per_out + i] = compute_alpha ? sgn(centered[co * per_out + i]) * alpha[co]
                                                 : sgn(centered[co * per_out + i]);
    orc_sign_f32(x, xb, (int64_t)g->n * g->c_in * g->h * g->w);
    for (int n = 0; n < g->n; ++n)
        for (int co = 0; co < g->c_out; ++co)
            for (int ho = 0; ho < ho_n; ++ho)
                for (int wo = 0; wo < wo_n; ++wo) {
                    float acc = 0.0f;
                    for (int ci = 0; ci < g->c_in; ++ci)
                        for (int kh = 0; kh < g->kh; ++kh) {
                            const int hi = ho * g->stride_h - g->pad_h + kh * g->dil_h;
                            if (hi < 0 || hi >= g->h) continue; /* zero padding of sign(x) */
                            for (int kw = 0; kw < g->kw; ++kw) {
                                const int wi = wo * g->stride_w - g->pad_w + kw * g->dil_w;
                                if (wi < 0 || wi >= g->w) continue;
                                acc += xb[(((int64_t)n * g->c_in + ci) * g->h + hi) * g->w + wi] *
                                       wb[co * per_out + ((int64_t)ci * g->kh + kh) * g->kw + kw];
                            }
                        }
                    if (bias) acc += bias[co];
                    if (post) acc *= post[co];
                    out[(((int64_t)n * g->c_out + co) * ho_n + ho) * wo_n + wo] = acc;
                }
    free(centered); free(alpha); free(wb); free(xb);
    return 0;
}

/* (A) bnn.layers.Linear.forward (bnn/layers/linear.py:22-27): x [rows,in], w [out,in]. */
int orc_floatsim_linear(const float *x, const float *w, const float *bias, const float *post,
                        int rows, int in_f, int out_f, int center, int compute_alpha, float *out) {
    orc_geom g = {1, in_f, 1, rows, out_f, 1, 1, 1, 1, 0, 0, 1, 1};
    /* view x as [1,in,1,rows] (transposed), out as [1,out,1,rows] (transposed) */
    float *xt = (float *)malloc(sizeof(float) * (int64_t)rows * in_f);
    float *ot = (float *)malloc(sizeof(float) * (int64_t)rows * out_f);
    if (!xt || !ot) return -1;
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < in_f; ++c) xt[(int64_t)c * rows + r] = x[(int64_t)r * in_f + c];
    int rc = orc_floatsim_conv2d(xt, w, bias, post, &g, center, compute_alpha, ot);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < out_f; ++c) out[(int64_t)r * out_f + c] = ot[(int64_t)c * rows + r];
    free(xt); free(ot);
    return rc;
}

/* ------------------------------------------------------------------ */
/* (B) packed integer path -- same layouts as include/bnn_b200.h.      */
/* ------------------------------------------------------------------ */

static int popc32(uint32_t v) { return __builtin_popcount(v); }

int64_t orc_act_units(int n, int c, int h, int w) { return (int64_t)n * ((c + 63) / 64) * h * w; }
int64_t orc_weight_words(int c_out, int c_in, int kh, int kw) {
    return (int64_t)((c_out + 31) / 32) * ((c_in + 63) / 64) * kh * kw * 32 * 2;
}

/*
 * Activation planes: abits[n][chunk][h][w] = {s_lo, s_hi, m_lo, m_hi} (4 x u32),
 * chunk = 64 channels; s bit <=> v>0, m bit <=> v>0 || v<0  (bnn/ops.py:66:
 * sign(0)=sign(-0)=sign(nan)=0 has m=0).  x is addressed with element strides so NCHW,
 * channels_last and Linear's [rows,in] (as n=1,h=1,w=rows) all go through the same routine.
 * pool > 1: v is first the AvgPool2d(kernel=stride=pool, ceil_mode, count_include_pad=False)
 * of x (bnn/models/resnet.py:129-133), summed row-major and divided by the in-bounds count;
 * pre_scale/pre_shift (may be NULL): v = v*pre_scale[c] + pre_shift[c] (an eval BatchNorm in
 * front of the binarized layer, e.g. PreBasicBlock.bn1, res_block.py:152-154); pre_relu: a ReLU
 * between that BatchNorm and the sign (HBlock, hierarchical_block.py:41-43).
 * h, w are the INPUT plane size; the output plane is ho x wo.
 */
void orc_pack_act(const float *x, int64_t sn, int64_t sc, int64_t sh, int64_t sw,
                  int n, int c, int h, int w, int pool, int ceil_mode,
                  const float *pre_scale, const float *pre_shift, int pre_relu, uint32_t *abits) {
    const int nch = (c + 63) / 64;
    const int k = pool > 1 ? pool : 1;
    const int ho = pool > 1 ? (ceil_mode ? (h + k - 1) / k : h / k) : h;
    const int wo = pool > 1 ? (ceil_mode ? (w + k - 1) / k : w / k) : w;
    for (int in = 0; in < n; ++in)
        for (int ih = 0; ih < ho; ++ih)
            for (int iw = 0; iw < wo; ++iw)
                for (int ch = 0; ch < nch; ++ch) {
                    uint32_t u[4] = {0, 0, 0, 0};
                    for (int b = 0; b < 64; ++b) {
                        const int ci = ch * 64 + b;
                        if (ci >= c) break;
                        float v;
                        if (k == 1) {
                            v = x[in * sn + ci * sc + ih * sh + iw * sw];
                        } else {
                            float sum = 0.0f;
                            int cnt = 0;
                            for (int i = 0; i < k && ih * k + i < h; ++i)
                                for (int j = 0; j < k && iw * k + j < w; ++j) {
                                    sum += x[in * sn + ci * sc + (ih * k + i) * sh + (iw * k + j) * sw];
                                    ++cnt;
                                }
                            v = sum / (float)cnt;
                        }
                        if (pre_scale) v = v * pre_scale[ci] + pre_shift[ci];
                        const uint32_t pos = v > 0.0f, neg = (v < 0.0f) && !pre_relu;   /* bn -> relu -> sign */
                        u[b >> 5] |= pos << (b & 31);
                        u[2 + (b >> 5)] |= (pos | neg) << (b & 31);
                    }
                    memcpy(abits + ((((int64_t)in * nch + ch) * ho + ih) * wo + iw) * 4, u, 16);
                }
}

/*
 * Weight planes: wbits[co/32][kstep][co%32][2] (u32), kstep = (chunk*kh + i)*kw + j,
 * word 0 = channels chunk*64+0..31, word 1 = +32..63; bit <=> centered w > 0.
 * alpha as orc_weight_center_alpha.  *n_zero = number of exactly-zero
 * (centered) weights: those have sign 0 in the reference (ops.py:66) which a
 * 1-bit plane cannot express -- the host refuses the packed path if n_zero != 0.
 */
int orc_pack_weight(const float *w, int c_out, int c_in, int kh, int kw, int center,
                    int compute_alpha, uint32_t *wbits, float *alpha, int32_t *n_zero) {
    const int taps = kh * kw, nch = (c_in + 63) / 64, nk = nch * taps;
    const int64_t per_out = (int64_t)c_in * taps;
    float *centered = (float *)malloc(sizeof(float) * per_out * c_out);
    if (!centered) return -1;
    orc_weight_center_alpha(w, c_out, c_in, kh, kw, center, compute_alpha, centered, alpha);
    memset(wbits, 0, sizeof(uint32_t) * orc_weight_words(c_out, c_in, kh, kw));
    int32_t zeros = 0;
    for (int co = 0; co < c_out; ++co)
        for (int ci = 0; ci < c_in; ++ci)
            for (int t = 0; t < taps; ++t) {
                const float v = centered[co * per_out + (int64_t)ci * taps + t];
                if (!(v > 0.0f) && !(v < 0.0f)) ++zeros;
                if (v > 0.0f) {
                    const int ks = (ci / 64) * taps + t, b = ci % 64;
                    wbits[((((int64_t)(co / 32) * nk + ks) * 32) + (co % 32)) * 2 + (b >> 5)] |= 1u << (b & 31);
                }
            }
    *n_zero = zeros;
    free(centered);
    return 0;
}

/*
 * Packed binary convolution with the fused epilogue of include/bnn_b200.h (struct bnn_epilogue):
 *   dot = sum popc(m) - 2 * sum popc(m & (s ^ t))   over the in-bounds taps
 *   y = (scale*dot + bias) * post                        reference conv.py:92-97, ops.py:136,202
 *   z = y*bn_scale + bn_shift ; z += res (pre) ; v = act(z) ; v += res (post)     
 *   out <- v ;  out_bits <- planes of sign(v*nx_scale + nx_shift)
 * When any fusion field is set, the per-channel constants are folded first (k0, k1 below) and the
 * affine steps are single fused multiply-adds; ReLU is fmaxf(z, 0).
 * Out-of-bounds taps contribute nothing (their m would be 0): zero padding is applied after
 * binarization, bnn/layers/conv.py:91-92.  Every float op is separately rounded (-ffp-contract=off).
 */
typedef struct {
    const float *scale, *bias, *post;
    const float *bn_scale, *bn_shift;
    const float *residual;
    int64_t rn, rc, rh, rw;
    int32_t residual_after_act, act;
    const float *act_slope;
    float *out;
    int64_t on, oc, oh, ow;
    uint32_t *out_bits;
    const float *nx_scale, *nx_shift;
    int32_t nx_relu, bits_before_residual;
} orc_epilogue;

void orc_bconv2d_fused(const uint32_t *abits, const uint32_t *wbits, const orc_geom *g, const orc_epilogue *e) {
    const int ho_n = orc_out_h(g), wo_n = orc_out_w(g);
    const int taps = g->kh * g->kw, nch = (g->c_in + 63) / 64, nk = nch * taps;
    const int ochunks = (g->c_out + 63) / 64;
    if (e->out_bits) memset(e->out_bits, 0, sizeof(uint32_t) * 4 * (size_t)g->n * ochunks * ho_n * wo_n);
    for (int n = 0; n < g->n; ++n)
        for (int ho = 0; ho < ho_n; ++ho)
            for (int wo = 0; wo < wo_n; ++wo)
                for (int co = 0; co < g->c_out; ++co) {
                    int msum = 0, dis = 0;
                    for (int ch = 0; ch < nch; ++ch)
                        for (int kh = 0; kh < g->kh; ++kh)
                            for (int kw = 0; kw < g->kw; ++kw) {
                                const int hi = ho * g->stride_h - g->pad_h + kh * g->dil_h;
                                const int wi = wo * g->stride_w - g->pad_w + kw * g->dil_w;
                                if (hi < 0 || hi >= g->h || wi < 0 || wi >= g->w) continue;
                                const uint32_t *u = abits + ((((int64_t)n * nch + ch) * g->h + hi) * g->w + wi) * 4;
                                const int ks = (ch * g->kh + kh) * g->kw + kw;
                                const uint32_t *t = wbits + ((((int64_t)(co / 32) * nk + ks) * 32) + (co % 32)) * 2;
                                msum += popc32(u[2]) + popc32(u[3]);
                                dis += popc32(u[2] & (u[0] ^ t[0])) + popc32(u[3] & (u[1] ^ t[1]));
                            }
                    const int fused = e->bn_scale || e->residual || e->act != 0 || e->out_bits || e->nx_scale;
                    float y, ybits = 0.0f;
                    if (!fused) {
                        /* reference order, conv.py:92-97 + ops.py:136,202 */
                        y = (e->scale ? e->scale[co] : 1.0f) * (float)(msum - 2 * dis);
                        y = y + (e->bias ? e->bias[co] : 0.0f);
                        y = y * (e->post ? e->post[co] : 1.0f);
                    } else {
                        /* fused mode folds scale/bias/post/BatchNorm into one multiply-add per channel:
                         *   k0 = scale*post*bn_scale, k1 = (bias*post)*bn_scale + bn_shift, z = fma(k0, dot, k1) */
                        const float post = e->post ? e->post[co] : 1.0f;
                        float k0 = (e->scale ? e->scale[co] : 1.0f) * post;
                        float k1 = (e->bias ? e->bias[co] : 0.0f) * post;
                        if (e->bn_scale) {
                            k0 = k0 * e->bn_scale[co];
                            k1 = k1 * e->bn_scale[co] + e->bn_shift[co];
                        }
                        y = fmaf(k0, (float)(msum - 2 * dis), k1);
                        const float r = e->residual ? e->residual[n * e->rn + co * e->rc + ho * e->rh + wo * e->rw] : 0.0f;
                        y = y + (e->residual_after_act ? 0.0f : r);
                        if (e->act == 1) y = fmaxf(y, 0.0f);
                        else if (e->act == 2) y = (y > 0.0f) ? y : e->act_slope[co] * y;
                        ybits = y;            /* value before the after-activation shortcut add */
                        y = y + (e->residual_after_act ? r : 0.0f);
                        if (!(e->residual && e->residual_after_act && e->bits_before_residual)) ybits = y;
                    }
                    if (e->out) e->out[n * e->on + co * e->oc + ho * e->oh + wo * e->ow] = y;
                    if (e->out_bits) {
                        float b = fused ? ybits : y;
                        if (e->nx_scale) b = fmaf(e->nx_scale[co], b, e->nx_shift[co]);
                        if (e->nx_relu && b < 0.0f) b = 0.0f;
                        uint32_t *u = e->out_bits + ((((int64_t)n * ochunks + co / 64) * ho_n + ho) * wo_n + wo) * 4;
                        const int bit = co % 64;
                        if (b > 0.0f) u[bit >> 5] |= 1u << (bit & 31);
                        if (b > 0.0f || b < 0.0f) u[2 + (bit >> 5)] |= 1u << (bit & 31);
                    }
                }
}

void orc_bconv2d(const uint32_t *abits, const uint32_t *wbits,
                 const float *scale, const float *bias, const float *post, const orc_geom *g,
                 float *out, int64_t on, int64_t oc, int64_t oh, int64_t ow) {
    orc_epilogue e;
    memset(&e, 0, sizeof(e));
    e.scale = scale; e.bias = bias; e.post = post;
    e.out = out; e.on = on; e.oc = oc; e.oh = oh; e.ow = ow;
    orc_bconv2d_fused(abits, wbits, g, &e);
}

/* raw integer dot products (scale=1, no bias/post) for bit-exact comparison */
void orc_bconv2d_dot(const uint32_t *abits, const uint32_t *wbits, const orc_geom *g, int32_t *dot) {
    const int ho_n = orc_out_h(g), wo_n = orc_out_w(g);
    const int64_t total = (int64_t)g->n * g->c_out * ho_n * wo_n;
    float *tmp = (float *)malloc(sizeof(float) * total);
    if (!tmp) return;
    orc_bconv2d(abits, wbits, NULL, NULL, NULL, g, tmp,
                (int64_t)g->c_out * ho_n * wo_n, (int64_t)ho_n * wo_n, wo_n, 1);
    for (int64_t i = 0; i < total; ++i) dot[i] = (int32_t)tmp[i];
    free(tmp);
}

/*
 * Stem of bnn.models.resnet (bnn/models/resnet.py:85-92,147-153): conv 7x7/2/pad 3 (3->64, no bias)
 * -> eval BatchNorm (folded affine) -> ReLU -> MaxPool 3x3/2/pad 1.  x [n,3,h,w], w [64,3,7,7] (torch
 * layout), out [n,hp,wp,64] (NHWC).  The conv is one fma chain per output in (c_in, kh, kw) order,
 * the order the CUDA kernel uses, so the comparison is bit-exact.
 */
void orc_stem(const float *x, int n, int h, int w, const float *wt, const float *bn_scale,
              const float *bn_shift, const float *nx_scale, const float *nx_shift, float *out,
              uint32_t *out_bits) {
    const int hc = (h + 6 - 7) / 2 + 1, wc = (w + 6 - 7) / 2 + 1;
    const int hp = (hc + 2 - 3) / 2 + 1, wp = (wc + 2 - 3) / 2 + 1;
    float *conv = (float *)malloc(sizeof(float) * (size_t)hc * wc * 64);
    if (!conv) return;
    for (int in = 0; in < n; ++in) {
        for (int r = 0; r < hc; ++r)
            for (int c = 0; c < wc; ++c)
                for (int co = 0; co < 64; ++co) {
                    float acc = 0.0f;
                    for (int ci = 0; ci < 3; ++ci)
                        for (int kh = 0; kh < 7; ++kh)
                            for (int kw = 0; kw < 7; ++kw) {
                                const int hi = 2 * r - 3 + kh, wi = 2 * c - 3 + kw;
                                const float v = (hi < 0 || hi >= h || wi < 0 || wi >= w)
                                                    ? 0.0f : x[(((size_t)in * 3 + ci) * h + hi) * w + wi];
                                acc = fmaf(v, wt[((co * 3 + ci) * 7 + kh) * 7 + kw], acc);
                            }
                    conv[((size_t)r * wc + c) * 64 + co] = fmaxf(fmaf(acc, bn_scale[co], bn_shift[co]), 0.0f);
                }
        for (int pr = 0; pr < hp; ++pr)
            for (int pc = 0; pc < wp; ++pc) {
                uint32_t u[4] = {0, 0, 0, 0};
                for (int co = 0; co < 64; ++co) {
                    float m = 0.0f;   /* ReLU output is >= 0, so 0 is neutral for the padded window */
                    for (int i = 0; i < 3; ++i)
                        for (int j = 0; j < 3; ++j) {
                            const int r = 2 * pr - 1 + i, c = 2 * pc - 1 + j;
                            if (r < 0 || r >= hc || c < 0 || c >= wc) continue;
                            m = fmaxf(m, conv[((size_t)r * wc + c) * 64 + co]);
                        }
                    out[(((size_t)in * hp + pr) * wp + pc) * 64 + co] = m;
                    const float b = nx_scale ? fmaf(nx_scale[co], m, nx_shift[co]) : m;
                    if (b > 0.0f) u[co >> 5] |= 1u << (co & 31);
                    if (b > 0.0f || b < 0.0f) u[2 + (co >> 5)] |= 1u << (co & 31);
                }
                if (out_bits) memcpy(out_bits + (((size_t)in * hp + pr) * wp + pc) * 4, u, 16);
            }
    }
    free(conv);
}
