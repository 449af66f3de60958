// scale_coords + clip_coords (core/utils/general.py:598-623 of the reference): the step every caller runs
// right after non_max_suppression (scripts/detect.py:218, scripts/val.py:250-256, core/models/common.py:778) —
// map xyxy boxes from the letterboxed network input back to the original image and clamp them to it.
// One thread per box; applied in place on the (bs, max_det, 6) detection buffer (row stride 6) or on any
// (n, 4+) row-strided box tensor.  Arithmetic as the reference's fp32 tensor ops: (x - pad) / gain with IEEE
// subtraction and division, clamp to [0, w0] / [0, h0], optional round-half-even (detect.py:218 `.round()`).
#include "common.cuh"

namespace dsn {

// params (bs, 5) fp32: pad_x, pad_y, gain, w0, h0 of every image
__global__ void __launch_bounds__(256) scale_coords_kernel(float *__restrict__ boxes, long long row_stride, long long image_stride,
                                                          const int *__restrict__ counts, int rows_per_image,
                                                          const float *__restrict__ params, int do_round) {
    const int b = blockIdx.y;
    const int n = counts ? min(counts[b], rows_per_image) : rows_per_image;
    const float px = params[b * 5 + 0], py = params[b * 5 + 1], gain = params[b * 5 + 2];
    const float w0 = params[b * 5 + 3], h0 = params[b * 5 + 4];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float *r = boxes + (size_t)b * image_stride + (size_t)i * row_stride;
        float x1 = __fdiv_rn(__fsub_rn(r[0], px), gain), y1 = __fdiv_rn(__fsub_rn(r[1], py), gain);      // general.py:607-609
        float x2 = __fdiv_rn(__fsub_rn(r[2], px), gain), y2 = __fdiv_rn(__fsub_rn(r[3], py), gain);
        x1 = fminf(fmaxf(x1, 0.f), w0); y1 = fminf(fmaxf(y1, 0.f), h0);                                  // general.py:614-619
        x2 = fminf(fmaxf(x2, 0.f), w0); y2 = fminf(fmaxf(y2, 0.f), h0);
        if (do_round) { x1 = rintf(x1); y1 = rintf(y1); x2 = rintf(x2); y2 = rintf(y2); }                // detect.py:218
        r[0] = x1; r[1] = y1; r[2] = x2; r[3] = y2;
    }
}

// val.py:227,243-255 for all labels of a batch in one launch: targets rows [image, cls, x, y, w, h] (normalised xywh) ->
// pixels (`targets[:, 2:] *= (w, h, w, h)`), xywh2xyxy (general.py:523-530), scale_coords + clip_coords into the
// image's native space (general.py:598-623) -> labelsn rows [cls, x1, y1, x2, y2].  One thread per label.
__global__ void __launch_bounds__(256) labels_to_native_kernel(const float *__restrict__ targets, long long n, float width, float height,
                                                              const float *__restrict__ params, int bs, float *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *t = targets + i * 6;
    const int b = (int)t[0];
    float *o = out + i * 5;
    o[0] = t[1];
    const float x = __fmul_rn(t[2], width), y = __fmul_rn(t[3], height), w = __fmul_rn(t[4], width), h = __fmul_rn(t[5], height);
    float x1 = __fsub_rn(x, __fdiv_rn(w, 2.f)), y1 = __fsub_rn(y, __fdiv_rn(h, 2.f));
    float x2 = __fadd_rn(x, __fdiv_rn(w, 2.f)), y2 = __fadd_rn(y, __fdiv_rn(h, 2.f));
    if (b >= 0 && b < bs) {
        const float px = params[b * 5 + 0], py = params[b * 5 + 1], gain = params[b * 5 + 2];
        const float w0 = params[b * 5 + 3], h0 = params[b * 5 + 4];
        x1 = __fdiv_rn(__fsub_rn(x1, px), gain); y1 = __fdiv_rn(__fsub_rn(y1, py), gain);
        x2 = __fdiv_rn(__fsub_rn(x2, px), gain); y2 = __fdiv_rn(__fsub_rn(y2, py), gain);
        x1 = fminf(fmaxf(x1, 0.f), w0); y1 = fminf(fmaxf(y1, 0.f), h0);
        x2 = fminf(fmaxf(x2, 0.f), w0); y2 = fminf(fmaxf(y2, 0.f), h0);
    }
    o[1] = x1; o[2] = y1; o[3] = x2; o[4] = y2;
}

}  // namespace dsn

extern "C" __attribute__((visibility("default"))) int dsn_labels_to_native(const float *targets, int64_t n, float width, float height,
                                                                           const float *params, int bs, float *labelsn, void *stream) {
    if (n < 0 || bs < 0) return DSN_E_INVALID;
    if (n == 0) return DSN_OK;
    if (!targets || !params || !labelsn) return DSN_E_INVALID;
    dsn::labels_to_native_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(targets, n, width, height, params, bs, labelsn);
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}

extern "C" __attribute__((visibility("default"))) int dsn_scale_coords(float *boxes, int64_t row_stride, int64_t image_stride,
                                                                       const int *counts, int bs, int rows_per_image,
                                                                       const float *params, int do_round, void *stream) {
    if (!boxes || !params || bs < 0 || rows_per_image < 0 || row_stride < 4) return DSN_E_INVALID;
    if (bs > 65535) return DSN_E_UNSUPPORTED;
    if (bs == 0 || rows_per_image == 0) return DSN_OK;
    const int gx = (rows_per_image + 255) / 256 > 64 ? 64 : (rows_per_image + 255) / 256;
    dsn::scale_coords_kernel<<<dim3(gx, bs), 256, 0, (cudaStream_t)stream>>>(boxes, row_stride, image_stride, counts,
                                                                             rows_per_image, params, do_round);
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}
