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

// Native serving loop: launches n captured post-process steps back to back (step i = graph_execs[i] on streams[i]).
// One Python call per burst instead of one per step: at ~30 us per step the host side of a per-step Python replay
// (stream context + graph launch) is the bottleneck as soon as several ranks share the host cores.
extern "C" __attribute__((visibility("default"))) int dsn_graph_launch_many(void *const *graph_execs, void *const *streams, int n) {
    if (n < 0 || (n > 0 && (!graph_execs || !streams))) return DSN_E_INVALID;
    for (int i = 0; i < n; ++i) {
        if (!graph_execs[i]) return DSN_E_INVALID;
        DSN_CUDA(cudaGraphLaunch((cudaGraphExec_t)graph_execs[i], (cudaStream_t)streams[i]));
    }
    return DSN_OK;
}

#ifdef DSN_TIMING
namespace dsn { __device__ unsigned long long dsn_timing[64]; __device__ unsigned long long dsn_log[2048]; }
extern "C" __attribute__((visibility("default"))) int dsn_debug_log(unsigned long long *host_out) {
    cudaError_t e = cudaMemcpyFromSymbol(host_out, dsn::dsn_log, sizeof(unsigned long long) * 2048);
    if (e != cudaSuccess) return (int)e;
    static unsigned long long zeros[2048];
    return (int)cudaMemcpyToSymbol(dsn::dsn_log, zeros, sizeof(zeros));
}
extern "C" __attribute__((visibility("default"))) int dsn_debug_timing(unsigned long long *host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, dsn::dsn_timing, sizeof(unsigned long long) * 64);
}
#endif
