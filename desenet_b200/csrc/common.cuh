// Shared device helpers for the DeSeNet post-process kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/desenet_b200.h"

#define DSN_CHECK_LAUNCH()                         \
    do {                                           \
        cudaError_t e__ = cudaGetLastError();      \
        if (e__ != cudaSuccess) return (int)e__;   \
    } while (0)

#define DSN_CUDA(expr)                             \
    do {                                           \
        cudaError_t e__ = (expr);                  \
        if (e__ != cudaSuccess) return (int)e__;   \
    } while (0)

namespace dsn {

// ---- per-device one-time kernel configuration (host) --------------------------------------------------------
// cudaFuncSetAttribute is per device: remember, per kernel and per device, that the opt-in dynamic shared memory
// size was raised.  A relaxed atomic bitmask per call site keeps the entry points re-entrant.
constexpr int kMaxDevices = 64;
struct DeviceOnce {
    unsigned long long done = 0ull;
    template <class F>
    int run(F &&f) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return (int)e;
        const unsigned long long bit = 1ull << (dev % kMaxDevices);
        if (__atomic_load_n(&done, __ATOMIC_RELAXED) & bit) return 0;
        const int rc = f();
        if (rc == 0) __atomic_fetch_or(&done, bit, __ATOMIC_RELAXED);
        return rc;
    }
};
template <class Kernel>
static inline int raise_dynamic_smem(DeviceOnce &once, Kernel kernel, int bytes) {
    return once.run([&]() { return (int)cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); });
}
static inline int sm_count_current_device() {
    static int cache[kMaxDevices] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    int &n = cache[dev % kMaxDevices];
    if (n == 0 && (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)) n = 148;
    return n;
}

// Optional phase timeline for kernel tuning (compiled in only with -DDSN_TIMING; never in the shipped .so)
#ifdef DSN_TIMING
extern __device__ unsigned long long dsn_timing[64];
__device__ __forceinline__ void dsn_mark(int slot) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        dsn_timing[slot] = t;
    }
}
#define DSN_MARK(slot) dsn::dsn_mark(slot)
// accumulate the time since the previous DSN_LAP / DSN_LAP_START of this thread into dsn_timing[slot]
// (a log of (slot, time) pairs written with plain stores - a read-modify-write of global memory per lap would cost the
// timed thread a memory round trip each time; the host sums the differences per slot)
extern __device__ unsigned long long dsn_log[2048];
#define DSN_LAP_START() unsigned int dsn_lap_n = 0; DSN_LAP(0)
#define DSN_LAP(slot) do { if (threadIdx.x == 0 && blockIdx.x == 0 && dsn_lap_n < 2048u) { unsigned long long t__; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__)); dsn::dsn_log[dsn_lap_n++] = ((unsigned long long)(slot) << 56) | (t__ & 0x00FFFFFFFFFFFFFFull); } } while (0)
#define DSN_COUNT(slot, v) do { if (threadIdx.x == 0 && blockIdx.x == 0) dsn::dsn_timing[slot] += (unsigned long long)(v); } while (0)
#define DSN_COUNT_ALL(slot_sum, slot_max, v) do { if (threadIdx.x == 0) { atomicAdd(&dsn::dsn_timing[slot_sum], (unsigned long long)(v)); atomicMax(&dsn::dsn_timing[slot_max], (unsigned long long)(v)); } } while (0)
#else
#define DSN_COUNT_ALL(slot_sum, slot_max, v) ((void)0)
#define DSN_COUNT(slot, v) ((void)0)
#define DSN_MARK(slot) ((void)0)
#define DSN_LAP_START() ((void)0)
#define DSN_LAP(slot) ((void)0)
#endif

// Streaming 128-bit loads that do not pollute L1 (inputs are read exactly once).
__device__ __forceinline__ float4 ldg_stream_f4(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream_f1(const float *p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ longlong2 ldg_stream_ll2(const long long *p) {
    longlong2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    return r;
}
// Streaming stores (outputs are never re-read by the producing kernel).
__device__ __forceinline__ void stg_stream_f4(float *p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// ---- element types of the raw network outputs (the `dtype` argument of the C ABI) ----------------------------------------
// The kernels compute in fp32 whatever the network's precision: fp16 / bf16 head outputs (`--half`, scripts/val.py:144,
// 178-179) are converted as they are loaded (exact), so the results are those of the fp32 path on the upcast values, with
// half the input bytes and no separate conversion pass.
//   ld1 / ld4   streaming loads of 1 / 4 consecutive elements from global memory (ld4: 4 * sizeof(T) aligned)
//   ld1c        cached load (values that are re-read)
//   lds4 / cvt  4 consecutive elements / one element already in hand (shared memory, registers)
template <typename T>
struct Elem;
template <>
struct Elem<float> {
    static __device__ __forceinline__ float ld1(const float *p) { return ldg_stream_f1(p); }
    static __device__ __forceinline__ float4 ld4(const float *p) { return ldg_stream_f4(p); }
    static __device__ __forceinline__ float ld1c(const float *p) { return __ldg(p); }
    static __device__ __forceinline__ float4 lds4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
    static __device__ __forceinline__ float cvt(float v) { return v; }
};
template <>
struct Elem<__half> {
    static __device__ __forceinline__ float cvt(__half v) { return __half2float(v); }
    static __device__ __forceinline__ float4 unpack(uint2 r) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&r.x)), b = __half22float2(*reinterpret_cast<const __half2 *>(&r.y));
        return make_float4(a.x, a.y, b.x, b.y);
    }
    static __device__ __forceinline__ float ld1(const __half *p) {
        unsigned short r;
        asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(r) : "l"(p));
        return __half2float(__ushort_as_half(r));
    }
    static __device__ __forceinline__ float4 ld4(const __half *p) {
        uint2 r;
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
        return unpack(r);
    }
    static __device__ __forceinline__ float ld1c(const __half *p) { return __half2float(__ldg(p)); }
    static __device__ __forceinline__ float4 lds4(const __half *p) { return unpack(*reinterpret_cast<const uint2 *>(p)); }
};
template <>
struct Elem<__nv_bfloat16> {
    static __device__ __forceinline__ float cvt(__nv_bfloat16 v) { return __bfloat162float(v); }
    static __device__ __forceinline__ float4 unpack(uint2 r) {          // bf16 -> fp32 is a 16-bit shift
        return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xFFFF0000u), __uint_as_float(r.y << 16),
                           __uint_as_float(r.y & 0xFFFF0000u));
    }
    static __device__ __forceinline__ float ld1(const __nv_bfloat16 *p) {
        unsigned short r;
        asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(r) : "l"(p));
        return __uint_as_float((unsigned int)r << 16);
    }
    static __device__ __forceinline__ float4 ld4(const __nv_bfloat16 *p) {
        uint2 r;
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
        return unpack(r);
    }
    static __device__ __forceinline__ float ld1c(const __nv_bfloat16 *p) { return __bfloat162float(__ldg(p)); }
    static __device__ __forceinline__ float4 lds4(const __nv_bfloat16 *p) { return unpack(*reinterpret_cast<const uint2 *>(p)); }
};
static inline int dtype_size(int dtype) { return dtype == DSN_F32 ? 4 : ((dtype == DSN_F16 || dtype == DSN_BF16) ? 2 : 0); }
// runs `...` with `T` bound to the element type of `dtype`
#define DSN_DISPATCH_DTYPE(dtype, T, ...)                                  \
    switch (dtype) {                                                       \
        case DSN_F32: { using T = float; __VA_ARGS__; break; }             \
        case DSN_F16: { using T = __half; __VA_ARGS__; break; }            \
        case DSN_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }    \
        default: return DSN_E_INVALID;                                     \
    }

// sigmoid as the reference computes it (torch.sigmoid, core/models/yolo.py:266): 1/(1+exp(-x)).
// expf / IEEE division (no fast-math): within a few ulp of ATen's CPU kernel; the north-star
// tolerance for decoded boxes and scores is 1e-5 relative.
// __frcp_rn is the correctly rounded reciprocal, i.e. exactly __fdiv_rn(1.0f, d), in fewer instructions.
__device__ __forceinline__ float sigmoidf_ref(float x) { return __frcp_rn(__fadd_rn(1.0f, expf(-x))); }

// monotone key of a confidence: larger conf -> smaller key.  conf > 0 always (it passed
// `conf > conf_thres >= 0`), so the raw IEEE bits are already ordered.
__device__ __forceinline__ uint32_t conf_key(float conf) { return ~__float_as_uint(conf); }
__device__ __forceinline__ float key_conf(uint64_t key) { return __uint_as_float(~(uint32_t)(key >> 32)); }

// ATen align_corners=True source index + lambdas, in fp32 (see oracle/c/oracle.c: coeffs()).
struct Lerp {
    int i0;
    float l0, l1;
};
__device__ __forceinline__ Lerp lerp_coeff(float scale, int dst, int n_in) {
    float src = __fmul_rn(scale, (float)dst);
    int a = (int)src;
    a = a > n_in - 1 ? n_in - 1 : a;
    Lerp r;
    r.i0 = a;
    r.l1 = __fsub_rn(src, (float)a);
    r.l0 = __fsub_rn(1.0f, r.l1);
    return r;
}

// argmax update, torch semantics: first maximum wins, NaN counts as maximal (first NaN wins).
__device__ __forceinline__ bool argmax_better(float cur, float best) {
    return !(cur <= best) && (best == best);
}

}  // namespace dsn
