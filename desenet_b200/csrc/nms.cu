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
//   4. suppress against the survivors so far, resolve the chunk in rounds of 256 candidates with a
//      bitmask + serial scan, append survivors to the output (general.py:745),
// and stops at max_det survivors (general.py:735-736), after max_nms candidates
// (general.py:728-729) or when the keys are exhausted.  Latency-bound, not HBM-bound.  Two instantiations of the kernel
// (nms_launch.cuh): 512-thread CTAs here, 256-thread CTAs in nms_small.cu; dsn_nms_levels_cta picks.
#include "nms_launch.cuh"

// the 256-thread instantiation (nms_small.cu)
int dsn_nms_levels_small(const void *const *raw, int dtype, const int *ny, const int *nx, const float *stride, const float *anchor_grid, int nl,
                         int bs, int na, int no, const uint64_t *keys, int64_t cap, const int *counts, float conf_thres, double iou_thres,
                         int agnostic, int max_nms, int max_det, float max_wh, float *det, int *det_count, void *stream);

namespace dsn {

struct FilterClassMask {
    unsigned long long m[4];
    int use;
};

// A-priori labels as candidates (general.py:690-696, val.py:228 save_hybrid): label j of image b is prediction row A + j
// with obj = 1 and a one-hot class, i.e. ONE candidate of confidence 1.0 for its class in both selection modes.
__global__ void append_label_keys_kernel(const float *__restrict__ labels, const int *__restrict__ label_count, int max_labels, int A,
                                         int nc, float conf_thres, FilterClassMask cm, unsigned long long *__restrict__ keys,
                                         long long cap, int *__restrict__ counts) {
    const int b = blockIdx.x;
    const int n = min(label_count[b], max_labels);
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const float c = labels[((size_t)b * max_labels + j) * 5];
        const int cls = (int)c;                                              // .long() of the class column (general.py:695)
        if (!(1.0f > conf_thres) || cls < 0 || cls >= nc) continue;          // conf = 1 * 1 must exceed the threshold (general.py:710-714)
        if (cm.use && !((cm.m[(cls >> 6) & 3] >> (cls & 63)) & 1ull)) continue;   // general.py:717-718
        const int slot = atomicAdd(counts + b, 1);
        if (slot < cap)
            keys[(size_t)b * cap + slot] = ((unsigned long long)conf_key(1.0f) << 32) | (unsigned long long)((unsigned)(A + j) * (unsigned)nc + (unsigned)cls);
    }
}

}  // namespace dsn

using namespace dsn;

extern "C" __attribute__((visibility("default"))) int dsn_nms_levels_cta(const void *const *raw, int dtype, const int *ny, const int *nx,
                              const float *stride, const float *anchor_grid, int nl, int bs, int na, int no, const uint64_t *keys,
                              int64_t cap, const int *counts, float conf_thres, double iou_thres, int agnostic,
                              int max_nms, int max_det, float max_wh, const unsigned int *cut_hi, const int *cut_dropped,
                              float *det, int *det_count, int cta_threads, void *stream) {
    if (cta_threads != 0 && cta_threads != 256 && cta_threads != 512) return DSN_E_INVALID;
    // Default rule: no confidence cut = the detect setting (a few hundred candidates per image, the seg mask kernel of another
    // batch running next to it): 256-thread CTAs, which find room between that kernel's short-lived CTAs (see nms_small.cu).
    // With a cut: thousands of candidates per image, 512.  The 256-thread instantiation carries no cut plumbing.
    const bool small = cta_threads == 256 || (cta_threads == 0 && !cut_dropped);
    if (small && !cut_dropped)
        return dsn_nms_levels_small(raw, dtype, ny, nx, stride, anchor_grid, nl, bs, na, no, keys, cap, counts, conf_thres, iou_thres, agnostic,
                                    max_nms, max_det, max_wh, det, det_count, stream);
    return nms_levels_impl(raw, dtype, ny, nx, stride, anchor_grid, nl, bs, na, no, keys, cap, counts, conf_thres, iou_thres, agnostic, max_nms,
                           max_det, max_wh, cut_hi, cut_dropped, nullptr, 0, det, det_count, stream);
}

extern "C" __attribute__((visibility("default"))) int dsn_nms_levels(const void *const *raw, int dtype, const int *ny, const int *nx, const float *stride,
                              const float *anchor_grid, int nl, int bs, int na, int no, const uint64_t *keys,
                              int64_t cap, const int *counts, float conf_thres, double iou_thres, int agnostic,
                              int max_nms, int max_det, float max_wh, const unsigned int *cut_hi, const int *cut_dropped,
                              float *det, int *det_count, void *stream) {
    return dsn_nms_levels_cta(raw, dtype, ny, nx, stride, anchor_grid, nl, bs, na, no, keys, cap, counts, conf_thres, iou_thres, agnostic, max_nms,
                              max_det, max_wh, cut_hi, cut_dropped, det, det_count, 0, stream);
}

extern "C" __attribute__((visibility("default"))) int dsn_nms_levels_labels(const void *const *raw, int dtype, const int *ny, const int *nx, const float *stride,
                              const float *anchor_grid, int nl, int bs, int na, int no, const uint64_t *keys,
                              int64_t cap, const int *counts, float conf_thres, double iou_thres, int agnostic,
                              int max_nms, int max_det, float max_wh, const float *labels, int max_labels, float *det,
                              int *det_count, void *stream) {
    return nms_levels_impl(raw, dtype, ny, nx, stride, anchor_grid, nl, bs, na, no, keys, cap, counts, conf_thres, iou_thres, agnostic, max_nms,
                           max_det, max_wh, nullptr, nullptr, labels, max_labels, det, det_count, stream);
}

extern "C" __attribute__((visibility("default"))) int dsn_append_label_keys(const float *labels, const int *label_count, int bs, int max_labels,
                                                                            int64_t A, int nc, float conf_thres, const uint64_t *class_mask,
                                                                            uint64_t *keys, int64_t cap, int *counts, void *stream) {
    if (bs < 0 || max_labels < 0 || A < 0 || nc < 1 || cap < 1 || !keys || !counts) return DSN_E_INVALID;
    if (!(conf_thres >= 0.f && conf_thres <= 1.f)) return DSN_E_INVALID;
    if (bs == 0 || max_labels == 0) return DSN_OK;
    if (!labels || !label_count) return DSN_E_INVALID;
    if ((unsigned long long)(A + max_labels) * (unsigned long long)nc >= (1ull << 32) || bs > 65535) return DSN_E_UNSUPPORTED;
    FilterClassMask cm;
    cm.use = class_mask != nullptr;
    for (int i = 0; i < 4; ++i) cm.m[i] = class_mask ? class_mask[i] : ~0ull;
    append_label_keys_kernel<<<bs, 128, 0, (cudaStream_t)stream>>>(labels, label_count, max_labels, (int)A, nc, conf_thres, cm,
                                                                   (unsigned long long *)keys, cap, counts);
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}

extern "C" __attribute__((visibility("default"))) int dsn_nms_z(const void *z, int dtype, int bs, int64_t A, int no, const uint64_t *keys, int64_t cap,
                         const int *counts, float conf_thres, double iou_thres, int agnostic, int max_nms, int max_det,
                         float max_wh, float *det, int *det_count, void *stream) {
    return nms_z_impl(z, dtype, bs, A, no, keys, cap, counts, conf_thres, iou_thres, agnostic, max_nms, max_det, max_wh, det, det_count, stream);
}
