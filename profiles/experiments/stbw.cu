// Pure-write bandwidth of the detect-mode mask (bs x H x W fp32, 105 MB at bs 64 / 640 px) under different store
// strategies — the ceiling the fused seg kernel can reach.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
// -cudart static -o stbw stbw.cu ; run on the B200: ./stbw
//   rows<P,CS>   : the seg kernel's decomposition (persistent grid, contiguous run of rows per CTA, P px per thread,
//                  one 16 B store per 4 px per row), st.global.cs or default stores
//   flat<U>      : grid-stride 16 B stores, U per thread per iteration (what a memset-style kernel does)
//   tma<R>       : rows staged in shared memory (STS.128), R rows at a time, written with ONE cp.async.bulk
//                  shared->global per R rows (UBLKCP), two stages
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void st_cs(float *p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int P, bool CS>
__global__ void rows(float *out, long long total_rows, int rows_per_cta, int W) {
    const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(total_rows, r0 + rows_per_cta);
    const int X = threadIdx.x * P;
    if (X >= W) return;
    float *p = out + r0 * W + X;
    for (long long r = r0; r < r1; ++r, p += W) {
        const float v = (float)(r & 1);
#pragma unroll
        for (int k = 0; k < P; k += 4) {
            if (CS) st_cs(p + k, make_float4(v, v, v, v));
            else *reinterpret_cast<float4 *>(p + k) = make_float4(v, v, v, v);
        }
    }
}

template <int U, bool CS>
__global__ void flat(float *out, long long n16) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride * U) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long j = i + u * stride;
            if (j < n16) {
                if (CS) st_cs(out + 4 * j, make_float4(1.f, 0.f, 1.f, 0.f));
                else *reinterpret_cast<float4 *>(out + 4 * j) = make_float4(1.f, 0.f, 1.f, 0.f);
            }
        }
    }
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// R rows per stage, 2 stages; blockDim.x * 4 == W
template <int R>
__global__ void tma(float *out, long long total_rows, int rows_per_cta, int W) {
    extern __shared__ __align__(128) unsigned char sm[];
    const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(total_rows, r0 + rows_per_cta);
    const size_t row_bytes = (size_t)W * 4;
    int stage = 0;
    for (long long r = r0; r < r1; r += R) {
        const int n = (int)min((long long)R, r1 - r);
        float *st = reinterpret_cast<float *>(sm + (size_t)stage * R * row_bytes);
        for (int q = 0; q < n; ++q) {
            const float v = (float)((r + q) & 1);
            *reinterpret_cast<float4 *>(st + (size_t)q * W + threadIdx.x * 4) = make_float4(v, v, v, v);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + r * W), "r"(smem_u32(st)),
                         "r"((unsigned)(n * row_bytes))
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");     // the OTHER stage may be overwritten now
        }
        __syncthreads();
        stage ^= 1;
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <class F>
static float time_us(F launch, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int i = 0; i < 5; ++i) launch(i);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) launch(i);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
    return ms * 1000.f / reps;
}

int main(int argc, char **argv) {
    const int bs = argc > 1 ? atoi(argv[1]) : 64, H = 640, W = 640, NB = 4, reps = 200;
    const long long total_rows = (long long)bs * H;
    const size_t bytes = (size_t)total_rows * W * 4;
    float *buf[NB];
    for (int i = 0; i < NB; ++i) cudaMalloc(&buf[i], bytes);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto report = [&](const char *name, float us) { printf("%-44s %8.2f us  %7.1f GB/s\n", name, us, bytes / us * 1e-3); };
    report("cudaMemsetAsync", time_us([&](int i) { cudaMemsetAsync(buf[i % NB], 0, bytes); }, reps));
    for (int occ : {1, 2, 3, 4, 6, 8}) {
        const long long grid = (long long)sms * occ;
        const int rpc = (int)((total_rows + grid - 1) / grid);
        const int g = (int)((total_rows + rpc - 1) / rpc);
        char nm[96];
        snprintf(nm, sizeof nm, "rows<4,cs>  %d CTAs/SM x160 thr", occ);
        report(nm, time_us([&](int i) { rows<4, true><<<g, 160>>>(buf[i % NB], total_rows, rpc, W); }, reps));
        snprintf(nm, sizeof nm, "rows<4,def> %d CTAs/SM x160 thr", occ);
        report(nm, time_us([&](int i) { rows<4, false><<<g, 160>>>(buf[i % NB], total_rows, rpc, W); }, reps));
        snprintf(nm, sizeof nm, "rows<8,cs>  %d CTAs/SM x80 thr", occ);
        report(nm, time_us([&](int i) { rows<8, true><<<g, 80>>>(buf[i % NB], total_rows, rpc, W); }, reps));
    }
    for (int per_sm : {2, 4, 8, 16, 32}) {
        char nm[96];
        snprintf(nm, sizeof nm, "flat<4,cs>  %d x256 thr per SM", per_sm);
        report(nm, time_us([&](int i) { flat<4, true><<<sms * per_sm, 256>>>(buf[i % NB], (long long)(bytes / 16)); }, reps));
        snprintf(nm, sizeof nm, "flat<4,def> %d x256 thr per SM", per_sm);
        report(nm, time_us([&](int i) { flat<4, false><<<sms * per_sm, 256>>>(buf[i % NB], (long long)(bytes / 16)); }, reps));
    }
    {
        const long long n16 = bytes / 16;
        const int g = (int)((n16 + 256 * 4 - 1) / (256 * 4));
        report("flat<4,def> one pass (non-persistent)", time_us([&](int i) { flat<4, false><<<g, 256>>>(buf[i % NB], n16); }, reps));
        report("flat<4,cs>  one pass (non-persistent)", time_us([&](int i) { flat<4, true><<<g, 256>>>(buf[i % NB], n16); }, reps));
    }
    for (int occ : {1, 2, 3, 4}) {
        const long long grid = (long long)sms * occ;
        const int rpc = (int)((total_rows + grid - 1) / grid);
        const int g = (int)((total_rows + rpc - 1) / rpc);
        char nm[96];
        cudaFuncSetAttribute(tma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        cudaFuncSetAttribute(tma<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        cudaFuncSetAttribute(tma<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        snprintf(nm, sizeof nm, "tma<4 rows>  %d CTAs/SM", occ);
        report(nm, time_us([&](int i) { tma<4><<<g, 160, 2 * 4 * W * 4>>>(buf[i % NB], total_rows, rpc, W); }, reps));
        snprintf(nm, sizeof nm, "tma<8 rows>  %d CTAs/SM", occ);
        report(nm, time_us([&](int i) { tma<8><<<g, 160, 2 * 8 * W * 4>>>(buf[i % NB], total_rows, rpc, W); }, reps));
        if (occ <= 2) {
            snprintf(nm, sizeof nm, "tma<16 rows> %d CTAs/SM", occ);
            report(nm, time_us([&](int i) { tma<16><<<g, 160, 2 * 16 * W * 4>>>(buf[i % NB], total_rows, rpc, W); }, reps));
        }
    }
    // correctness spot check of the bulk-store variant
    {
        const long long grid = (long long)sms * 2;
        const int rpc = (int)((total_rows + grid - 1) / grid);
        cudaMemset(buf[0], 0xff, bytes);
        tma<8><<<(int)((total_rows + rpc - 1) / rpc), 160, 2 * 8 * W * 4>>>(buf[0], total_rows, rpc, W);
        float *h = (float *)malloc(bytes);
        cudaMemcpy(h, buf[0], bytes, cudaMemcpyDeviceToHost);
        long long bad = 0;
        for (long long r = 0; r < total_rows; ++r)
            for (int x = 0; x < W; x += 37) bad += h[r * W + x] != (float)(r & 1);
        printf("tma<8> check: %lld mismatches\n", bad);
        free(h);
    }
    return 0;
}
