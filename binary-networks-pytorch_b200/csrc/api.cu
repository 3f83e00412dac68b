// api.cu -- bnn_query / bnn_strerror and the launch counter.
#include "common.cuh"

#include <atomic>
#include <cstdio>
#include <cstdlib>

namespace bnn {
static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("BNN_B200_NO_PDL"); return !(e && e[0] && e[0] != '0'); }();
    return on;
}
}  // namespace bnn

extern "C" int bnn_query(int what, int64_t* value) {
    if (!value) return BNN_E_NULL;
    switch (what) {
        case BNN_Q_ABI_VERSION: *value = BNN_B200_ABI_VERSION; return 0;
        case BNN_Q_SM_ARCH: *value = 100; return 0;
        case BNN_Q_DEVICE_SMS: {
            int dev = 0, sms = 0;
            cudaError_t e = cudaGetDevice(&dev);
            if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            *value = sms;
            return (int)e;
        }
        case BNN_Q_LAUNCH_COUNT: *value = bnn::g_launches.load(std::memory_order_relaxed); return 0;
        default: return BNN_E_SHAPE;
    }
}

extern "C" const char* bnn_strerror(int code) {
    switch (code) {
        case 0: return "ok";
        case BNN_E_NULL: return "bnn_b200: required pointer is NULL";
        case BNN_E_SHAPE: return "bnn_b200: invalid dimension";
        case BNN_E_UNSUPPORTED: return "bnn_b200: geometry not supported by this build";
        case BNN_E_DRIVER: return "bnn_b200: cuTensorMapEncodeTiled unavailable or failed";
        case BNN_E_ALIGN: return "bnn_b200: buffer not 16-byte aligned";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "bnn_b200: unknown error";
}
