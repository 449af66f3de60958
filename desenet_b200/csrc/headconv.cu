// Producer of the seg tail's input (first slice): the seg head's last convolution,
// nn.Conv2d(c_hid, n_segcls, kernel_size=1) with bias (core/models/yolo.py:182; same tail in the other seg heads :71,
// :122, :148, :228), as a streaming kernel that emits the LOW-RES fp32 logits the fused upsample + argmax kernels read.
// It takes the feature map in the network's precision (fp32 / fp16 / bf16, converted on load), accumulates in fp32 in
// channel order with one FMA per term, and writes fp32 - under --half the reference's cuDNN conv rounds the logits to fp16
// before the argmax, this path does not.  No tensor cores: K = c_hid (128 for DeSeNet-s) x C = 2..4 outputs per pixel is a
// GEMV per pixel, 2*C flops per loaded element - HBM-bound by two orders of magnitude.
//   one thread = 4 consecutive pixels, all C outputs; weights in shared memory; 8 channel loads in flight per thread
#include <algorithm>

#include "common.cuh"

namespace dsn {

template <typename T, int C, int V>
__global__ void __launch_bounds__(256) head_conv1x1_kernel(const T *__restrict__ x, const float *__restrict__ weight,
                                                           const float *__restrict__ bias, int K, long long plane,
                                                           float *__restrict__ out) {
    extern __shared__ float sw[];                      // [C][K]
    for (int i = threadIdx.x; i < C * K; i += blockDim.x) sw[i] = weight[i];
    __syncthreads();
    const int b = blockIdx.y;
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;          // unit of V pixels
    if (u * V >= plane) return;
    const T *px = x + (size_t)b * K * plane + u * V;
    float acc[C][V];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
        for (int j = 0; j < V; ++j) acc[c][j] = bias ? bias[c] : 0.f;
#pragma unroll 8
    for (int k = 0; k < K; ++k) {
        float v[V];
        if constexpr (V == 4) {
            const float4 t = Elem<T>::ld4(px + (size_t)k * plane);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
            v[0] = Elem<T>::ld1(px + (size_t)k * plane);
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float wk = sw[c * K + k];
#pragma unroll
            for (int j = 0; j < V; ++j) acc[c][j] = __fmaf_rn(wk, v[j], acc[c][j]);
        }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
        float *po = out + ((size_t)b * C + c) * plane + u * V;
        if constexpr (V == 4) *reinterpret_cast<float4 *>(po) = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
        else po[0] = acc[c][0];
    }
}

template <typename T, int C>
static int launch_head_conv(const void *x, const float *w, const float *bias, int bs, int K, long long plane, float *out, bool vec,
                            cudaStream_t st) {
    const size_t smem = (size_t)C * K * sizeof(float);
    if (smem > 48 * 1024) return DSN_E_UNSUPPORTED;
    const long long units = vec ? plane / 4 : plane;
    dim3 grid((unsigned)((units + 255) / 256), bs);
    if (vec) head_conv1x1_kernel<T, C, 4><<<grid, 256, smem, st>>>(static_cast<const T *>(x), w, bias, K, plane, out);
    else head_conv1x1_kernel<T, C, 1><<<grid, 256, smem, st>>>(static_cast<const T *>(x), w, bias, K, plane, out);
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}

}  // namespace dsn

using namespace dsn;

extern "C" __attribute__((visibility("default"))) int dsn_seg_head_conv1x1(const void *features, int dtype, const float *weight, const float *bias,
                                                                           int bs, int K, int C, int h, int w, float *logits, void *stream) {
    if (!features || !weight || !logits || bs < 0 || K < 1 || C < 1 || h < 1 || w < 1 || dtype_size(dtype) == 0) return DSN_E_INVALID;
    if (C > 4 || bs > 65535) return DSN_E_UNSUPPORTED;
    if (((uintptr_t)features & (uintptr_t)(dtype_size(dtype) - 1)) != 0 || ((uintptr_t)logits & 3) != 0) return DSN_E_ALIGN;
    if (bs == 0) return DSN_OK;
    const long long plane = (long long)h * w;
    const bool vec = plane % 4 == 0 && (((uintptr_t)features & (uintptr_t)(4 * dtype_size(dtype) - 1)) == 0) && (((uintptr_t)logits & 15) == 0);
    cudaStream_t st = (cudaStream_t)stream;
    DSN_DISPATCH_DTYPE(dtype, T, {
        switch (C) {
            case 1: return launch_head_conv<T, 1>(features, weight, bias, bs, K, plane, logits, vec, st);
            case 2: return launch_head_conv<T, 2>(features, weight, bias, bs, K, plane, logits, vec, st);
            case 3: return launch_head_conv<T, 3>(features, weight, bias, bs, K, plane, logits, vec, st);
            default: return launch_head_conv<T, 4>(features, weight, bias, bs, K, plane, logits, vec, st);
        }
    })
    return DSN_OK;
}
