// Version / error-string entry points of the C ABI (include/desenet_b200.h).
#include "common.cuh"

extern "C" __attribute__((visibility("default"))) int dsn_version(void) { return DSN_VERSION; }

extern "C" __attribute__((visibility("default"))) const char *dsn_error_string(int code) {
    switch (code) {
        case DSN_OK: return "ok";
        case DSN_E_INVALID: return "invalid argument (null pointer, non-positive size or out-of-range value)";
        case DSN_E_UNSUPPORTED: return "shape not supported by the sm_100a kernels";
        case DSN_E_ALIGN: return "pointer is not aligned as the interface requires";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown desenet_b200 error code";
}

#ifdef DSN_TIMING
namespace dsn { __device__ unsigned long long dsn_timing[64]; }
extern "C" __attribute__((visibility("default"))) int dsn_debug_timing(unsigned long long *host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, dsn::dsn_timing, sizeof(unsigned long long) * 64);
}
#endif
