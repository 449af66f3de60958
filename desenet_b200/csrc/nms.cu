// Per-image NMS over the 64-bit candidate keys (core/utils/general.py:728-745 of the reference,
// torchvision.ops.nms CPU arithmetic at the call site general.py:734).
//
// One CTA per image.  torchvision evaluates all n^2/2 pairs and the reference then keeps the first
// max_det survivors; greedy NMS in descending-score order never needs candidates that come after
// the max_det-th survivor, so this kernel consumes the candidates LAZILY in score order:
//   1. histogram-select the next <= kChunk best remaining keys (radix refinement on the 64-bit key,
//      one pass over the image's keys per level),
//   2. gather + bitonic-sort them in shared memory (ascending key = descending confidence, ties by
//      reference candidate order = torchvision's stable sort),
//   3. fetch their boxes (decode on the fly from the raw head planes, or from z), add the class
//      offset (general.py:732-733),
//   4. suppress against the survivors so far, resolve the chunk in 64-wide sub-blocks with a
//      bitmask + serial scan, append survivors to the output (general.py:745),
// and stops at max_det survivors (general.py:735-736), after max_nms candidates
// (general.py:728-729) or when the keys are exhausted.  Latency-bound, not HBM-bound: it runs next
// to the streaming kernels on otherwise idle SMs.
#include "nms_core.cuh"

namespace dsn {

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

}  // namespace dsn

using namespace dsn;

extern "C" __attribute__((visibility("default"))) int dsn_nms_levels(const float *const *raw, const int *ny, const int *nx, const float *stride,
                              const float *anchor_grid, int nl, int bs, int na, int no, const uint64_t *keys,
                              int64_t cap, const int *counts, float conf_thres, double iou_thres, int agnostic,
                              int max_nms, int max_det, float max_wh, const int *cut_dropped, float *det, int *det_count,
                              void *stream) {
    if (!stride || !anchor_grid) return DSN_E_INVALID;
    BoxFromLevels src;
    int rc = make_geom(src.g, raw, ny, nx, stride, anchor_grid, nl, bs, na, no);
    if (rc) return rc;
    NmsParams p;
    rc = make_nms_params(p, keys, cap, counts, bs, no - 5, conf_thres, iou_thres, agnostic, max_nms, max_det, max_wh, det, det_count);
    if (rc) return rc;
    p.cut_dropped = cut_dropped;
    return launch_nms(p, src, bs, (cudaStream_t)stream);
}

extern "C" __attribute__((visibility("default"))) int dsn_nms_z(const float *z, int bs, int64_t A, int no, const uint64_t *keys, int64_t cap,
                         const int *counts, float conf_thres, double iou_thres, int agnostic, int max_nms, int max_det,
                         float max_wh, float *det, int *det_count, void *stream) {
    if (!z || A < 1 || no < 6) return DSN_E_INVALID;
    NmsParams p;
    int rc = make_nms_params(p, keys, cap, counts, bs, no - 5, conf_thres, iou_thres, agnostic, max_nms, max_det, max_wh, det, det_count);
    if (rc) return rc;
    BoxFromZ src{z, A, no};
    return launch_nms(p, src, bs, (cudaStream_t)stream);
}
