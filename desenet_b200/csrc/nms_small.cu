// Second instantiation of the NMS kernel: 256-thread CTAs (16 K registers instead of 32 K per image).
// A smaller CTA finds room on an SM next to the CTAs of the HBM-bound seg / filter kernels of the other batches in
// flight, where the 512-thread CTA waits for half an SM to drain: at the detect setting (a few hundred candidates per
// image, no confidence cut) the pipelined step is 3-4 % faster with it although the kernel alone is slower
// (cfg 2: NMS alone 17.6 -> 20.7 us, step 28.8 -> 27.8 us).  With the confidence cut in force (eval setting, thousands of
// candidates per image) the 512-thread CTA wins on both counts and stays (profiles/experiments/x7_nms256.sh).
#define DSN_NMS_THREADS 256
#define DSN_NMS_NS nms_t256
#include "nms_launch.cuh"

using namespace dsn;

int dsn_nms_levels_small(const void *const *raw, int dtype, const int *ny, const int *nx, const float *stride, const float *anchor_grid, int nl,
                         int bs, int na, int no, const uint64_t *keys, int64_t cap, const int *counts, float conf_thres, double iou_thres,
                         int agnostic, int max_nms, int max_det, float max_wh, float *det, int *det_count, void *stream) {
    return nms_levels_impl(raw, dtype, ny, nx, stride, anchor_grid, nl, bs, na, no, keys, cap, counts, conf_thres, iou_thres, agnostic, max_nms,
                           max_det, max_wh, nullptr, nullptr, nullptr, 0, det, det_count, stream);
}
