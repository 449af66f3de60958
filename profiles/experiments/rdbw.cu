// pure-read bandwidth of an int64 target batch with the seg kernel's access pattern (contiguous run of rows per CTA,
// 160-thread CTAs, 8 x 16 B loads in flight per thread) vs a grid-interleaved pattern
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ longlong2 ld2(const long long *p) { longlong2 r; asm volatile("ld.global.nc.L1::no_allocate.v2.s64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p)); return r; }
// mode 0: CTA owns rows [b*rpc, (b+1)*rpc), thread g owns 4 px of every row (W = 4*G), 4 rows in flight
template <int ROWS_IN_FLIGHT>
__global__ void runs(const long long *t, long long total_rows, int rows_per_cta, int W, unsigned long long *out) {
    const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(total_rows, r0 + rows_per_cta);
    long long acc = 0;
    for (int g = threadIdx.x; g < W / 4; g += blockDim.x) {
        const long long *p = t + r0 * W + g * 4;
        long long r = r0;
        for (; r + ROWS_IN_FLIGHT <= r1; r += ROWS_IN_FLIGHT) {
            longlong2 a[ROWS_IN_FLIGHT], b[ROWS_IN_FLIGHT];
#pragma unroll
            for (int q = 0; q < ROWS_IN_FLIGHT; ++q) { a[q] = ld2(p + (size_t)q * W); b[q] = ld2(p + (size_t)q * W + 2); }
#pragma unroll
            for (int q = 0; q < ROWS_IN_FLIGHT; ++q) acc += a[q].x + a[q].y + b[q].x + b[q].y;
            p += (size_t)ROWS_IN_FLIGHT * W;
        }
        for (; r < r1; ++r) { longlong2 a = ld2(p), b = ld2(p + 2); acc += a.x + a.y + b.x + b.y; p += W; }
    }
    if (acc == 0x123456789LL) out[0] = acc;
}
// same as runs<4>, but only the LOW 32 bits of every int64 are used (what the seg kernel needs): ptxas then narrows
// each 16-byte load into two 4-byte loads of the used words
__global__ void runs_low(const long long *t, long long total_rows, int rows_per_cta, int W, unsigned long long *out) {
    const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(total_rows, r0 + rows_per_cta);
    unsigned int acc = 0;
    for (int g = threadIdx.x; g < W / 4; g += blockDim.x) {
        const long long *p = t + r0 * W + g * 4;
        long long r = r0;
        for (; r + 4 <= r1; r += 4) {
            longlong2 a[4], b[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { a[q] = ld2(p + (size_t)q * W); b[q] = ld2(p + (size_t)q * W + 2); }
#pragma unroll
            for (int q = 0; q < 4; ++q) acc += (unsigned)a[q].x + (unsigned)a[q].y + (unsigned)b[q].x + (unsigned)b[q].y;
            p += (size_t)4 * W;
        }
        for (; r < r1; ++r) { longlong2 a = ld2(p), b = ld2(p + 2); acc += (unsigned)a.x + (unsigned)a.y + (unsigned)b.x + (unsigned)b.y; p += W; }
    }
    if (acc == 0x12345678u) out[0] = acc;
}
// mode 1: flat grid-stride over 16 B units, U loads in flight
template <int U>
__global__ void flat(const long long *t, long long n16, unsigned long long *out) {
    long long acc = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n16; i += U * stride) {
        longlong2 a[U];
#pragma unroll
        for (int q = 0; q < U; ++q) a[q] = ld2(t + (i + q * stride) * 2);
#pragma unroll
        for (int q = 0; q < U; ++q) acc += a[q].x + a[q].y;
    }
    for (; i < n16; i += stride) { longlong2 a = ld2(t + i * 2); acc += a.x + a.y; }
    if (acc == 0x123456789LL) out[0] = acc;
}
int main() {
    const int bs = 64, H = 640, W = 640;
    const size_t n = (size_t)bs * H * W;
    long long *buf[4]; for (auto &b : buf) { cudaMalloc(&b, n * 8); cudaMemset(b, 0, n * 8); }
    unsigned long long *out; cudaMalloc(&out, 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto time = [&](auto fn, const char *name) {
        for (int i = 0; i < 8; ++i) fn(i);
        cudaDeviceSynchronize();
        cudaEventRecord(a);
        for (int i = 0; i < 40; ++i) fn(i);
        cudaEventRecord(b); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("%-44s %7.1f us  %6.0f GB/s\n", name, ms * 1e3 / 40, n * 8 / (ms / 40 * 1e-3) / 1e9);
    };
    const long long rows = (long long)bs * H;
    for (int ctas : {148 * 2, 148 * 3, 148 * 4, 148 * 6, 148 * 8}) {
        const int rpc = (int)((rows + ctas - 1) / ctas);
        char nm[96];
        snprintf(nm, 96, "runs  %4d CTAs x160, 4 rows in flight", ctas);
        time([&](int i) { runs<4><<<ctas, 160>>>(buf[i % 4], rows, rpc, W, out); }, nm);
        snprintf(nm, 96, "runs  %4d CTAs x160, 4 rows, LOW words only", ctas);
        time([&](int i) { runs_low<<<ctas, 160>>>(buf[i % 4], rows, rpc, W, out); }, nm);
        snprintf(nm, 96, "runs  %4d CTAs x160, 8 rows in flight", ctas);
        time([&](int i) { runs<8><<<ctas, 160>>>(buf[i % 4], rows, rpc, W, out); }, nm);
    }
    for (int ctas : {148 * 2, 148 * 4, 148 * 8}) {
        char nm[96];
        snprintf(nm, 96, "flat  %4d CTAs x256, 8 loads in flight", ctas);
        time([&](int i) { flat<8><<<ctas, 256>>>(buf[i % 4], (long long)n / 2, out); }, nm);
        snprintf(nm, 96, "flat  %4d CTAs x1024, 4 loads in flight", ctas);
        time([&](int i) { flat<4><<<ctas, 1024>>>(buf[i % 4], (long long)n / 2, out); }, nm);
    }
    return 0;
}
