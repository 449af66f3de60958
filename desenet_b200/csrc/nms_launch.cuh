// The NMS kernel, its launcher and the argument plumbing of the two entry families, for ONE CTA size (DSN_NMS_THREADS,
// inline namespace DSN_NMS_NS): included by nms.cu (512 threads) and nms_small.cu (256 threads).
#pragma once
#include "nms_core.cuh"

namespace dsn {
inline namespace DSN_NMS_NS {

template <class BoxSrc>
__global__ void __launch_bounds__(kNmsThreads, 1024 / kNmsThreads) nms_kernel(const NmsParams p, const BoxSrc src) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NmsSmem &s = *reinterpret_cast<NmsSmem *>(smem_raw);
    const int b = blockIdx.x;
    const int total = p.counts[b];
    if (total > p.cap) {                     // keys were dropped: the caller must retry with a larger cap
        if (threadIdx.x == 0) p.det_count[b] = -1;
        return;
    }
    if (total <= 0 || p.max_det <= 0 || p.max_nms <= 0) {
        const bool starved = total <= 0 && p.max_det > 0 && p.max_nms > 0 && p.cut_dropped && p.cut_dropped[b];
        if (threadIdx.x == 0) p.det_count[b] = starved ? -2 : 0;
        return;
    }
    nms_image(s, p, src, b, p.keys + (size_t)b * p.cap, total);
}

template <class BoxSrc>
static int launch_nms(const NmsParams &p, const BoxSrc &src, int bs, cudaStream_t st) {
    if (bs == 0) return DSN_OK;
    const size_t smem = sizeof(NmsSmem);
    static DeviceOnce once;           // per template instantiation
    const int rc = raise_dynamic_smem(once, nms_kernel<BoxSrc>, (int)smem);
    if (rc) return rc;
    nms_kernel<BoxSrc><<<bs, kNmsThreads, smem, st>>>(p, src);
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}


static inline int nms_levels_impl(const void *const *raw, int dtype, const int *ny, const int *nx, const float *stride, const float *anchor_grid, int nl,
                           int bs, int na, int no, const uint64_t *keys, int64_t cap, const int *counts, float conf_thres,
                           double iou_thres, int agnostic, int max_nms, int max_det, float max_wh, const unsigned int *cut_hi,
                           const int *cut_dropped, const float *labels, int max_labels, float *det, int *det_count, void *stream) {
    if (!stride || !anchor_grid) return DSN_E_INVALID;
    if (max_labels < 0 || (max_labels > 0 && !labels)) return DSN_E_INVALID;
    HeadGeom g;
    int rc = make_geom(g, raw, dtype, ny, nx, stride, anchor_grid, nl, bs, na, no);
    if (rc) return rc;
    if ((unsigned long long)(g.A + max_labels) * (unsigned long long)(no - 5) >= (1ull << 32)) return DSN_E_UNSUPPORTED;
    NmsParams p;
    rc = make_nms_params(p, keys, cap, counts, bs, no - 5, conf_thres, iou_thres, agnostic, max_nms, max_det, max_wh, det, det_count);
    if (rc) return rc;
    p.cut_dropped = cut_dropped;
    p.cut_hi = cut_dropped ? cut_hi : nullptr;
    DSN_DISPATCH_DTYPE(dtype, T, {
        BoxFromLevels<T> src;
        src.g = g;
        src.extra = labels;
        src.extra_rows = max_labels;
        return launch_nms(p, src, bs, (cudaStream_t)stream);
    })
    return DSN_OK;
}


static inline int nms_z_impl(const void *z, int dtype, int bs, int64_t A, int no, const uint64_t *keys, int64_t cap, const int *counts,
                             float conf_thres, double iou_thres, int agnostic, int max_nms, int max_det, float max_wh, float *det,
                             int *det_count, void *stream) {
    if (!z || A < 1 || no < 6 || dtype_size(dtype) == 0) return DSN_E_INVALID;
    if (((uintptr_t)z & (uintptr_t)(dtype_size(dtype) - 1)) != 0) return DSN_E_ALIGN;
    NmsParams p;
    int rc = make_nms_params(p, keys, cap, counts, bs, no - 5, conf_thres, iou_thres, agnostic, max_nms, max_det, max_wh, det, det_count);
    if (rc) return rc;
    DSN_DISPATCH_DTYPE(dtype, T, {
        BoxFromZ<T> src{static_cast<const T *>(z), A, no};
        return launch_nms(p, src, bs, (cudaStream_t)stream);
    })
    return DSN_OK;
}

}  // inline namespace DSN_NMS_NS
}  // namespace dsn
