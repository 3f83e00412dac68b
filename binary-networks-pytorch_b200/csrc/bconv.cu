// bconv.cu -- host side of the XNOR/AND + popcount binary convolution: tensor map, tile planner, autotuner, dispatch.
// The kernel template lives in bconv_kernel.cuh, its instances in bconv_inst.cu (one translation unit per epilogue kind).
//
// XNOR/AND + popcount binary convolution with fused epilogue.
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
#include <cstdio>
#include <cstdlib>
#include "bconv_kernel.cuh"

#include <algorithm>
#include <cudaTypedefs.h>
#include <map>
#include <mutex>
#include <vector>

namespace bnn {

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

static KernelFn pick_kernel(const Plan& p, int epi) {
    switch (epi) {
        case 0: return pick_kernel_epi0(p);
        case 1: return pick_kernel_epi1(p);
        case 2: return pick_kernel_epi2(p);
        case 3: return pick_kernel_epi3(p);
        default: return pick_kernel_epi4(p);
    }
}

// 0 = reference epilogue, 1 = fused, 2 = fused with channel-contiguous residual / output (or neither), 3 / 4 = lean forms
static int epilogue_kind(const bnn_epilogue& ep) {
    const bool fused = ep.bn_scale || ep.residual || ep.act != BNN_ACT_NONE || ep.out_bits || ep.nx_scale;
    if (!fused) return 0;
    const bool cl = (!ep.out || ep.ostride_c == 1) && (!ep.residual || ep.rstride_c == 1);
    if (!cl) return 1;
    // 3: residual blocks of either flavour (post-activation: BatchNorm folded, shortcut before the ReLU; pre-activation:
    // PReLU, shortcut after it, the next BatchNorm in front of the sign) -- everything but the Hierarchical-Block forms;
    // whether the lean instance can actually run also depends on the tile plan, see lean_plan_ok()
    if (ep.nx_relu || ep.bits_before_residual) return 2;
    const bool fast = ep.act == BNN_ACT_RELU && !ep.nx_scale && !(ep.residual && ep.residual_after_act);
    return fast ? 3 : 4;
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
    // carry-save front end: over the three taps of a 3-wide kernel row, or over chunk triples of a 1x1 kernel
    const int mode = ((kwt == 3 || (kwt == 1 && g.kh == 1)) && !(flags & BNN_F_NO_CSA)) ? 1 : 0;
    const int nblk32 = ceil_div(g.c_out, 32);
    const size_t smem_cap = 220 * 1024;
    const int candP[3] = {8, 7, 4}, candC[3] = {4, 2, 1};
    // Tiles span the whole output row whenever that fits (shrink == 0).  Only if NO shape fits shared memory that way
    // (many input chunks x a large kernel x a wide image) are narrower tiles tried: the row is cut into 2, 4, 8, ...
    // pieces until something fits, so the plans of every geometry that already had one are unchanged.
    for (int shrink = 0; shrink < 8 && out.empty(); ++shrink)
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
            int TW = ceil_div(ceil_div(Wo, 1 << shrink), P) * P;
            while ((TW - 1) * g.stride_w + (g.kw - 1) * g.dil_w + 1 > 256 && TW > P) TW -= P;
            const int BW = (TW - 1) * g.stride_w + (g.kw - 1) * g.dil_w + 1;
            if (BW > 256) continue;
            const int gpr = TW / P, tiles_w = ceil_div(Wo, TW);
            // cycles per 32-bit word per warp: fewer channels per lane -> more shared-memory loads per word
            double cpw = (C == 4 ? 1.0 : C == 2 ? 1.25 : 1.5) * (window ? 1.0 : 1.12) * (P == 4 ? 1.06 : 1.0);
            const double round_work = (double)P * C * nk * 2.0 * cpw + 60.0 * P * C;   // main loop + epilogue
            // warps per CTA: 8 / 7 (two or three resident CTAs), or fewer -- smaller CTAs, more of them resident: about the
            // same number of warps per SM, but the prologues, epilogue stalls and tails of four to eight CTAs interleave
            // instead of two (registers per thread are those of the instance; shared memory decides whether they all fit)
            const int nws[4] = {8, 7, 4, 2};        // (5 and 3 warps were measured too: never the best, r02al tuner log)
            for (int nwi = 0; nwi < 4; ++nwi) {
                const int NW = nws[nwi];
                for (int TH = 1; TH <= Ho; ++TH) {
                    const int BH = (TH - 1) * g.stride_h + (g.kh - 1) * g.dil_h + 1;
                    if (BH > 256) break;
                    const size_t smem = plan_smem(nch, BH, BW, C, nk, NW, P, TH, TW);
                    if (smem > smem_cap) break;
                    const int G = TH * gpr, rounds = ceil_div(G, NW);
                    if (rounds > 12 && TH > 1) break;
                    // resident CTAs: what the instance is compiled for (register budget), capped by shared memory
                    // resident CTAs by registers (the instance is compiled for bconv_min_ctas CTAs of 256 threads)
                    const int by_regs = NW >= 7 ? bconv_min_ctas(P, C, kwt) : bconv_warps_per_sm(P, C, kwt) / NW;
                    int occ = by_regs < 32 ? by_regs : 32;
                    while (occ > 1 && smem * occ > 226 * 1024) --occ;
                    if (NW < 7 && occ < by_regs) continue;       // small CTAs only pay when all of them fit
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

// a launch over a chunk range of a wider layer (split-K): strides of the full packed tensors
struct Slice { int nch_total, nk_total; };

static int launch_bconv(const void* abits, const void* wbits, const bnn_conv_geom& g, const bnn_epilogue& ep,
                        uint32_t flags, cudaStream_t stream, const Plan* forced = nullptr, const Slice* slice = nullptr) {
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
    KernelFn fn = pick_kernel(pl, (epi >= 3 && !lean_plan_ok(pl, g, Wo)) ? 2 : epi);
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
    if ((long long)a.nk * 64 >= (1LL << 22)) return BNN_E_UNSUPPORTED;     // epilogues convert |dot| < 2^22 by bit pattern
    a.TH = pl.TH; a.TW = pl.TW; a.BH = pl.BH; a.BW = pl.BW; a.gpr = pl.gpr; a.G = pl.G;
    a.tiles_h = pl.tiles_h; a.tiles_w = pl.tiles_w;
    a.act_bytes = (unsigned)((size_t)a.nch * pl.BH * pl.BW * 16);
    a.w_bytes = (unsigned)((size_t)a.nk * 256);
    a.stage_ldg = (flags & BNN_F_STAGE_LDG) ? 1 : 0;
    a.nch_total = slice ? slice->nch_total : a.nch;
    a.w_blk_stride = (unsigned)((slice ? slice->nk_total : a.nk) * 32);

    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if (!a.stage_ldg) {
        EncodeTiledFn enc = encode_tiled_fn();
        if (!enc) return BNN_E_DRIVER;
        // abits as a 5-D u32 tensor, fastest first: {4 words, W, H, chunks, N}
        const cuuint64_t gdim[5] = {4, (cuuint64_t)g.w, (cuuint64_t)g.h, (cuuint64_t)a.nch, (cuuint64_t)g.n};
        const cuuint64_t gstr[4] = {16, (cuuint64_t)g.w * 16, (cuuint64_t)g.w * g.h * 16,
                                    (cuuint64_t)g.w * g.h * a.nch_total * 16};
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
    ce = launch_pdl(fn, grid, dim3(pl.NW * 32), pl.smem, stream, tmap, a);
    count_launch(1);
    return (int)(ce != cudaSuccess ? ce : cudaGetLastError());
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
    // the model's best candidate of every (P, C) family first (the families differ in registers / resident CTAs, which
    // the model only approximates), then its best few overall with at most two per (P, C, warps) family
    std::vector<Plan> tries;
    for (const Cand& c : cands) {
        bool seen = false;
        for (const Plan& t : tries) seen |= (t.P == c.pl.P && t.C == c.pl.C && (t.NW >= 7 ? 8 : t.NW) == (c.pl.NW >= 7 ? 8 : c.pl.NW));
        if (!seen) tries.push_back(c.pl);
    }
    const size_t families = tries.size();
    for (const Cand& c : cands) {
        if (tries.size() >= families + (size_t)(top_k > 0 ? top_k : 6)) break;
        int same = 0;
        bool dup = false;
        for (const Plan& t : tries) {
            same += (t.P == c.pl.P && t.C == c.pl.C && t.NW == c.pl.NW);
            dup |= (t.P == c.pl.P && t.C == c.pl.C && t.NW == c.pl.NW && t.TH == c.pl.TH);
        }
        if (same < 2 && !dup) tries.push_back(c.pl);
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    // BNN_B200_TUNE_LOG=1: one stderr line per timed candidate (which plans exist for a layer and how far apart they are)
    static const bool tune_log = [] { const char* e = getenv("BNN_B200_TUNE_LOG"); return e && e[0] && e[0] != '0'; }();
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
        if (tune_log)
            fprintf(stderr, "[bnn tune] n%d c%d->%d %dx%d k%d s%d epi%d | P%d C%d TH%d NW%d ctas %lld smem %zu | %.4f ms\n", g.n,
                    g.c_in, g.c_out, g.h, g.w, g.kh, g.stride_h, epi, pl.P, pl.C, pl.TH, pl.NW,
                    (long long)g.n * pl.tiles_h * pl.tiles_w * ceil_div(ceil_div(g.c_out, 32), pl.C), pl.smem, ms_min);
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

// the candidate of enumerate_plans() whose (P, C, TH, warps) match; TH <= 0 / warps <= 0 match the best-ranked one
static int find_plan(const bnn_conv_geom& g, uint32_t flags, int P, int C, int TH, int warps, Plan* out) {
    const int Ho = out_dim(g.h, g.kh, g.stride_h, g.pad_h, g.dil_h);
    const int Wo = out_dim(g.w, g.kw, g.stride_w, g.pad_w, g.dil_w);
    if (Ho <= 0 || Wo <= 0) return BNN_E_SHAPE;
    std::vector<Cand> cands;
    int rc = enumerate_plans(g, Ho, Wo, flags, 148, cands);
    if (rc) return rc;
    for (const Cand& c : cands)
        if (c.pl.P == P && c.pl.C == C && (TH <= 0 || c.pl.TH == TH) && (warps <= 0 || c.pl.NW == warps)) {
            *out = c.pl;
            return 0;
        }
    return BNN_E_UNSUPPORTED;
}

extern "C" int bnn_bconv2d_fused_fwd_plan(const void* abits, const void* wbits, const bnn_conv_geom* geom,
                                          const bnn_epilogue* epilogue, uint32_t flags, int32_t P, int32_t C,
                                          int32_t TH, int32_t warps, void* stream) {
    if (!geom || !epilogue) return BNN_E_NULL;
    if (epilogue_kind(*epilogue)) flags &= ~BNN_F_NO_CSA;
    Plan pl;
    int rc = find_plan(*geom, flags, P, C, TH, warps, &pl);
    if (rc) return rc;
    return launch_bconv(abits, wbits, *geom, *epilogue, flags, (cudaStream_t)stream, &pl);
}

// ---------------------------------------------------------------------------
// split-K: layers whose full reduction does not fit one CTA's shared memory (Linear(25088, 4096), > 16384 input
// channels) are contracted chunk range by chunk range; every launch writes exact integer dots (as fp32) of its range
// and bnn_dot_finish_f32 sums them and applies the reference epilogue.  The same finish kernel averages the two
// launches of a TERNARY weight tensor (exact-zero weights, sign(0) = 0): zeros packed once as +1 and once as -1.
// ---------------------------------------------------------------------------
static bool plans_exist(const bnn_conv_geom& g, uint32_t flags) {
    const int Ho = out_dim(g.h, g.kh, g.stride_h, g.pad_h, g.dil_h);
    const int Wo = out_dim(g.w, g.kw, g.stride_w, g.pad_w, g.dil_w);
    std::vector<Cand> cands;
    return Ho > 0 && Wo > 0 && enumerate_plans(g, Ho, Wo, flags, 148, cands) == 0;
}

extern "C" int bnn_conv_split(const bnn_conv_geom* g, uint32_t flags, int32_t* chunks_per_part, int32_t* nparts) {
    if (!g || !chunks_per_part || !nparts) return BNN_E_NULL;
    if (g->c_in <= 0 || g->kh <= 0 || g->kw <= 0) return BNN_E_SHAPE;
    const int nch = ceil_div(g->c_in, 64);
    if (plans_exist(*g, flags)) { *chunks_per_part = nch; *nparts = 1; return 0; }
    // largest chunk range that has a plan (feasibility is monotone in the number of chunks: smaller tiles always fit)
    int lo = 0, hi = nch;                      // lo: feasible (0 = none found yet), hi: infeasible
    while (hi - lo > 1) {
        const int mid = (lo + hi) / 2;
        bnn_conv_geom sub = *g;
        sub.c_in = mid * 64;
        if (plans_exist(sub, flags)) lo = mid; else hi = mid;
    }
    if (lo == 0) return BNN_E_UNSUPPORTED;
    const int parts = ceil_div(nch, lo);
    *chunks_per_part = ceil_div(nch, parts);   // balanced ranges
    *nparts = ceil_div(nch, *chunks_per_part);
    return 0;
}

extern "C" int bnn_bconv2d_partial_fwd(const void* abits, const void* wbits, const bnn_conv_geom* geom, int32_t chunk0,
                                       int32_t nchunks, float* part, uint32_t flags, void* stream) {
    if (!geom || !part) return BNN_E_NULL;
    const bnn_conv_geom& g = *geom;
    if (g.c_in <= 0 || g.h <= 0 || g.w <= 0 || g.kh <= 0 || g.kw <= 0) return BNN_E_SHAPE;
    const int nch_total = ceil_div(g.c_in, 64);
    if (chunk0 < 0 || nchunks <= 0 || chunk0 + nchunks > nch_total) return BNN_E_SHAPE;
    const int Ho = out_dim(g.h, g.kh, g.stride_h, g.pad_h, g.dil_h);
    const int Wo = out_dim(g.w, g.kw, g.stride_w, g.pad_w, g.dil_w);
    if (Ho <= 0 || Wo <= 0) return BNN_E_SHAPE;
    bnn_conv_geom sub = g;
    sub.c_in = std::min(g.c_in - chunk0 * 64, nchunks * 64);
    const Slice sl{nch_total, nch_total * g.kh * g.kw};
    bnn_epilogue ep{};
    ep.out = part;
    ep.ostride_n = (int64_t)g.c_out * Ho * Wo; ep.ostride_c = (int64_t)Ho * Wo; ep.ostride_h = Wo; ep.ostride_w = 1;
    const unsigned char* a0 = abits ? (const unsigned char*)abits + (size_t)chunk0 * g.h * g.w * 16 : nullptr;
    const unsigned char* w0 = wbits ? (const unsigned char*)wbits + (size_t)chunk0 * g.kh * g.kw * 256 : nullptr;
    return launch_bconv(a0, w0, sub, ep, flags, (cudaStream_t)stream, nullptr, &sl);
}

namespace bnn {
__global__ void __launch_bounds__(256)
dot_finish_kernel(const float* __restrict__ parts, int nparts, long long count, float inv_div, const float* __restrict__ scale,
                  const float* __restrict__ bias, const float* __restrict__ post, float* __restrict__ out,
                  long long on, long long oc, long long oh, long long ow, int C, int H, int W) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float dot = 0.0f;
    for (int r = 0; r < nparts; ++r) dot = __fadd_rn(dot, parts[(long long)r * count + i]);      // exact: small integers
    dot = __fmul_rn(dot, inv_div);                                                              // exact: even sums, 1 or 1/2
    const int w = (int)(i % W);
    long long t = i / W;
    const int h = (int)(t % H);
    t /= H;
    const int c = (int)(t % C);
    const long long n = t / C;
    const float k0 = scale ? __ldg(scale + c) : 1.0f, k1 = bias ? __ldg(bias + c) : 0.0f, k2 = post ? __ldg(post + c) : 1.0f;
    // reference order (conv.py:92-97, ops.py:136,202): (alpha*dot + bias) * alpha_post, three rounded operations
    out[n * on + c * oc + h * oh + w * ow] = __fmul_rn(__fadd_rn(__fmul_rn(k0, dot), k1), k2);
}
}  // namespace bnn

extern "C" int bnn_dot_finish_f32(const float* parts, int32_t nparts, int32_t divisor, const float* scale, const float* bias,
                                  const float* post, float* out, int64_t on, int64_t oc, int64_t oh, int64_t ow, int32_t n,
                                  int32_t c_out, int32_t ho, int32_t wo, void* stream) {
    if (!parts || !out) return BNN_E_NULL;
    if (nparts <= 0 || (divisor != 1 && divisor != 2) || n <= 0 || c_out <= 0 || ho <= 0 || wo <= 0) return BNN_E_SHAPE;
    const long long count = (long long)n * c_out * ho * wo;
    const long long blocks = (count + 255) / 256;
    if (blocks > 0x7fffffffLL) return BNN_E_UNSUPPORTED;
    dot_finish_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(parts, nparts, count, 1.0f / (float)divisor, scale, bias,
                                                                         post, out, on, oc, oh, ow, c_out, ho, wo);
    count_launch(1);
    return (int)cudaGetLastError();
}

extern "C" int bnn_conv_plan_list(const bnn_conv_geom* g, uint32_t flags, int32_t* plans, int32_t cap, int32_t* count) {
    if (!g || !count || (cap > 0 && !plans)) return BNN_E_NULL;
    const int Ho = out_dim(g->h, g->kh, g->stride_h, g->pad_h, g->dil_h);
    const int Wo = out_dim(g->w, g->kw, g->stride_w, g->pad_w, g->dil_w);
    if (Ho <= 0 || Wo <= 0) return BNN_E_SHAPE;
    std::vector<Cand> cands;
    int rc = enumerate_plans(*g, Ho, Wo, flags, 148, cands);
    if (rc) return rc;
    *count = (int32_t)cands.size();
    for (int i = 0; i < (int)cands.size() && i < cap; ++i) {
        const Plan& pl = cands[i].pl;
        const int v[12] = {pl.P, pl.C, pl.kwt, pl.swt, pl.mode, pl.TH, pl.TW, pl.NW, pl.tiles_h * pl.tiles_w * g->n,
                           ceil_div(ceil_div(g->c_out, 32), pl.C), (int)pl.smem, pl.G};
        for (int k = 0; k < 12; ++k) plans[i * 12 + k] = v[k];
    }
    return 0;
}

extern "C" int bnn_conv_instance(const bnn_conv_geom* g, const bnn_epilogue* ep, uint32_t flags, int32_t P, int32_t C,
                                 int32_t TH, int32_t warps, int32_t* inst) {
    if (!g || !ep || !inst) return BNN_E_NULL;
    int epi = epilogue_kind(*ep);
    if (epi) flags &= ~BNN_F_NO_CSA;
    Plan pl;
    int rc = find_plan(*g, flags, P, C, TH, warps, &pl);
    if (rc) return rc;
    const int Wo = out_dim(g->w, g->kw, g->stride_w, g->pad_w, g->dil_w);
    if (epi >= 3 && !lean_plan_ok(pl, *g, Wo)) epi = 2;
    if (!pick_kernel(pl, epi)) return BNN_E_UNSUPPORTED;
    const int v[6] = {pl.P, pl.C, pl.kwt, pl.swt, pl.mode, epi};
    for (int k = 0; k < 6; ++k) inst[k] = v[k];
    return 0;
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
