// bconv_inst.cu -- the bconv_kernel instances of ONE epilogue kind (-DBNN_EPI=0..4, see the Makefile): splitting the
// ~200 instances over five translation units lets them compile in parallel.
#include "bconv_kernel.cuh"

#ifndef BNN_EPI
#error "compile with -DBNN_EPI=<0..4>"
#endif
#define BNN_CAT2(a, b) a##b
#define BNN_CAT(a, b) BNN_CAT2(a, b)

namespace bnn {

// EPI 0 = reference epilogue: both inner-loop modes (A/B runs with BNN_F_NO_CSA); EPI >= 1 = fused epilogues
KernelFn BNN_CAT(pick_kernel_epi, BNN_EPI)(const Plan& p) {
    constexpr int E = BNN_EPI;
    if (p.kwt == 3 && p.swt == 1) {
        if constexpr (E == 0) { if (!p.mode) return pick_pc<3, 1, 0, E>(p.P, p.C); }
        return pick_pc<3, 1, 1, E>(p.P, p.C);
    }
    if (p.kwt == 3 && p.swt == 2) {
        if constexpr (E == 0) { if (!p.mode) return pick_pc<3, 2, 0, E>(p.P, p.C); }
        return pick_pc<3, 2, 1, E>(p.P, p.C);
    }
    if (p.kwt == 1 && p.swt == 1) return p.mode ? pick_pc<1, 1, 1, E>(p.P, p.C) : pick_pc<1, 1, 0, E>(p.P, p.C);
    return pick_pc<0, 0, 0, E>(p.P, p.C);
}

}  // namespace bnn
