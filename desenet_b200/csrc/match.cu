// Detection-metric step right after NMS:
//   * box_iou        core/utils/general.py:626-648   (N x M IoU matrix)
//   * process_batch  scripts/val.py:101-122          (labels x detections matching -> `correct` (N, niou) bools)
// Both are tiny (<= 300 detections x <= ~100 labels per image) and latency-bound; the point of having them on the
// device is that val.run (scripts/val.py:235-270) no longer copies every image's detections to the host.
#include "common.cuh"

namespace dsn {

// torch.min / torch.max (elementwise) propagate NaN; fminf / fmaxf do not.
__device__ __forceinline__ float t_min(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fminf(a, b); }
__device__ __forceinline__ float t_max(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b); }
__device__ __forceinline__ float t_clamp0(float a) { return a != a ? a : fmaxf(a, 0.f); }

// iou of box a (x1,y1,x2,y2) and box b exactly as general.py:640-648 evaluates it with fp32 tensor ops
__device__ __forceinline__ float box_iou_ref(const float *a, const float *b) {
    const float area1 = __fmul_rn(__fsub_rn(a[2], a[0]), __fsub_rn(a[3], a[1]));
    const float area2 = __fmul_rn(__fsub_rn(b[2], b[0]), __fsub_rn(b[3], b[1]));
    const float w = t_clamp0(__fsub_rn(t_min(a[2], b[2]), t_max(a[0], b[0])));
    const float h = t_clamp0(__fsub_rn(t_min(a[3], b[3]), t_max(a[1], b[1])));
    const float inter = __fmul_rn(w, h);
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area1, area2), inter));
}

__global__ void __launch_bounds__(256) box_iou_kernel(const float *__restrict__ box1, long long stride1, int n,
                                                      const float *__restrict__ box2, long long stride2, int m,
                                                      float *__restrict__ out) {
    const long long total = (long long)n * m;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t / m), j = (int)(t - (long long)i * m);
        out[t] = box_iou_ref(box1 + (size_t)i * stride1, box2 + (size_t)j * stride2);
    }
}

// One CTA per image.  labels rows: [cls, x1, y1, x2, y2]; detection rows: [x1, y1, x2, y2, conf, cls].
// val.py:111-121: candidate pairs = (iou >= iouv[0]) & class match; sort by IoU descending; per detection keep its
// best label (np.unique on the detection column, first occurrence); then per label keep the FIRST detection in
// detection order (np.unique on the label column of the detection-ordered array — the upstream re-sort is commented
// out, val.py:117); correct[d] = iou >= iouv.  Ties between equal IoUs of one detection are resolved towards the
// lowest label index (the reference's order there depends on numpy's unstable argsort).
__global__ void __launch_bounds__(256) process_batch_kernel(const float *__restrict__ det, long long det_image_stride,
                                                            long long det_row_stride, const int *__restrict__ det_count,
                                                            int max_rows, const float *__restrict__ labels,
                                                            const int *__restrict__ label_off, const float *__restrict__ iouv,
                                                            int niou, unsigned char *__restrict__ correct) {
    extern __shared__ int win[];                  // per label: lowest detection index whose best label it is
    const int b = blockIdx.x;
    const int n = det_count ? min(det_count[b], max_rows) : max_rows;
    const int l0 = label_off[b], m = label_off[b + 1] - l0;
    const float *lab = labels + (size_t)l0 * 5;
    const float *d0 = det + (size_t)b * det_image_stride;
    unsigned char *c0 = correct + (size_t)b * max_rows * niou;
    for (int i = threadIdx.x; i < m; i += blockDim.x) win[i] = 0x7fffffff;
    __syncthreads();
    const float thr0 = iouv[0];
    // registers hold the state of up to kPer detections per thread
    constexpr int kPer = 8;
    int best_l[kPer];
    float best_iou[kPer];
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        best_l[q] = -1;
        best_iou[q] = 0.f;
        const int d = threadIdx.x + q * blockDim.x;
        if (d < n) {
            const float *dr = d0 + (size_t)d * det_row_stride;
            const float box[4] = {dr[0], dr[1], dr[2], dr[3]};
            const float cls = dr[5];
            for (int l = 0; l < m; ++l) {
                if (lab[l * 5] != cls) continue;                                  // val.py:111 class match
                const float iou = box_iou_ref(lab + l * 5 + 1, box);             // box_iou(labels, detections)
                if (iou >= thr0 && (best_l[q] < 0 || iou > best_iou[q])) { best_l[q] = l; best_iou[q] = iou; }
            }
            if (best_l[q] >= 0) atomicMin(&win[best_l[q]], d);
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        const int d = threadIdx.x + q * blockDim.x;
        if (d < n) {
            const bool matched = best_l[q] >= 0 && win[best_l[q]] == d;
            for (int k = 0; k < niou; ++k) c0[(size_t)d * niou + k] = (matched && best_iou[q] >= iouv[k]) ? 1 : 0;   // val.py:121
        }
    }
}

}  // namespace dsn

extern "C" __attribute__((visibility("default"))) int dsn_box_iou(const float *box1, int64_t stride1, int n, const float *box2,
                                                                  int64_t stride2, int m, float *out, void *stream) {
    if (n < 0 || m < 0 || stride1 < 4 || stride2 < 4) return DSN_E_INVALID;
    if (n == 0 || m == 0) return DSN_OK;
    if (!box1 || !box2 || !out) return DSN_E_INVALID;
    const long long total = (long long)n * m;
    const int grid = (int)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
    dsn::box_iou_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(box1, stride1, n, box2, stride2, m, out);
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}

extern "C" __attribute__((visibility("default"))) int dsn_process_batch(const float *det, int64_t det_image_stride, int64_t det_row_stride,
                                                                        const int *det_count, int bs, int max_rows,
                                                                        const float *labels, const int *label_off,
                                                                        int max_labels_per_image, const float *iouv, int niou,
                                                                        unsigned char *correct, void *stream) {
    if (bs < 0 || max_rows < 0 || niou < 1 || max_labels_per_image < 0 || det_row_stride < 6) return DSN_E_INVALID;
    if (bs == 0 || max_rows == 0) return DSN_OK;
    if (!det || !label_off || !iouv || !correct || (max_labels_per_image > 0 && !labels)) return DSN_E_INVALID;
    if (max_rows > 8 * 256 || max_labels_per_image > 12000) return DSN_E_UNSUPPORTED;
    const size_t smem = sizeof(int) * (size_t)(max_labels_per_image > 0 ? max_labels_per_image : 1);
    dsn::process_batch_kernel<<<bs, 256, smem, (cudaStream_t)stream>>>(det, det_image_stride, det_row_stride, det_count, max_rows,
                                                                       labels, label_off, iouv, niou, correct);
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}
