/*
 * desenet_b200.h — C ABI of the B200-native (sm_100a) DeSeNet post-network inference path.
 *
 * The reference (splwany/DeSeNet) exposes no FFI/plugin layer for this path: its boundary is a
 * set of Python callables.  This header is the boundary a binding would bind:
 * plain pointers and sizes, no torch types.  Every entry point
 *   - returns int: 0 = ok, >0 = a cudaError_t raised while enqueueing, <0 = argument error
 *     (DSN_E_*), never throws;
 *   - only ENQUEUES work on `stream` (a cudaStream_t passed as void*; NULL = legacy default
 *     stream) and never synchronises the device;
 *   - owns no caller-visible memory: every buffer is allocated by the caller (device memory
 *     unless the parameter says "host").
 *
 * Reference file:line each entry point replaces is given at its declaration
 * (paths are into the reference tree).
 */
#ifndef DESENET_B200_H_
#define DESENET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSN_VERSION 200          /* major*10000 + minor*100 + patch */
#define DSN_MAX_LEVELS 8         /* detection levels (reference: nl = 3, core/models/yolo.py:244) */
#define DSN_MAX_ANCHORS 16       /* anchors per level (reference: na = 3, yolo.py:245) */
#define DSN_MAX_CM_ROWS 8        /* confusion-matrix rows handled by the packed-counter path */

/* element type of the raw network outputs (`dtype` arguments): the kernels always compute in fp32; fp16 / bf16 inputs
 * (the reference's --half, scripts/val.py:144,178-179) are converted as they are loaded */
#define DSN_F32 0
#define DSN_F16 1
#define DSN_BF16 2

/* argument-error codes (negative); positive return values are cudaError_t */
#define DSN_OK 0
#define DSN_E_INVALID (-1)       /* null pointer / non-positive size / out-of-range value */
#define DSN_E_UNSUPPORTED (-2)   /* shape outside what the kernels implement */
#define DSN_E_ALIGN (-3)         /* pointer not aligned as documented */

int dsn_version(void);
/* Static string for a return code of any function below (cudaGetErrorString for codes > 0). */
const char *dsn_error_string(int code);

/* Native serving loop: launch n captured steps back to back, step i = cudaGraphExec_t graph_execs[i] on cudaStream_t
 * streams[i] (handles of graphs the caller captured from the enqueue-only entry points below, e.g.
 * torch.cuda.CUDAGraph.raw_cuda_graph_exec()).  Enqueue-only. */
int dsn_graph_launch_many(void *const *graph_execs, void *const *streams, int n);

/* ------------------------------------------------------------------------------------------
 * Detection head geometry, shared by the calls below.
 *   raw[l]      device pointer, level l conv output, NCHW (bs, na*no, ny[l], nx[l]), elements of type `dtype`
 *               (DSN_F32 / DSN_F16 / DSN_BF16); channel = a*no + o  (core/models/yolo.py:258-260)
 *   stride[l]   host, pixels per cell (8,16,32)                      (yolo.py:315)
 *   anchor_grid host, nl*na*2 floats, PIXEL anchors (w,h) = the `anchor_grid` buffer (yolo.py:251)
 *   no = nc + 5; rows of the decoded prediction are ordered level-major, then anchor, gy, gx
 *   (yolo.py:275-277), A = sum_l na*ny[l]*nx[l].
 * ------------------------------------------------------------------------------------------ */

/* Detect.forward, eval branch, everything after the 1x1 convs (core/models/yolo.py:259-277).
 * z      (bs, A, no) fp32 or NULL:  sigmoid + grid/anchor xywh decode, levels concatenated.
 * xperm  host array of nl device pointers or NULL: xperm[l] (bs, na, ny, nx, no) = the permuted
 *        raw logits the reference also returns (yolo.py:260). */
int dsn_detect_decode(const void *const *raw, int dtype, const int *ny, const int *nx, const float *stride,
                      const float *anchor_grid, int nl, int bs, int na, int no,
                      float *z, float *const *xperm, void *stream);

/* Candidate keys.  A candidate is one (row, class) pair that passes non_max_suppression's
 * filters (core/utils/general.py:668,687,703,710-718).  It is emitted as ONE 64-bit key
 *     key = (~float_bits(conf) << 32) | (row * nc + cls)
 * so ascending key order == descending confidence, ties by reference candidate order
 * (general.py:710-714 row-major) — the order torchvision's stable sort sees (general.py:734).
 * keys    (bs, cap) uint64, written in arbitrary order within an image
 * counts  (bs) int32: total candidates of the image (may exceed cap: then only the first cap
 *         arrivals were stored and the caller must retry with a larger cap)
 * class_mask host, 4 x uint64 bitmask of allowed classes (general.py:717-718) or NULL = all.
 */

/* Fused decode + filter straight from the raw head outputs: sigmoid(obj) > conf_thres,
 * conf = obj*cls, multi-label / best-class selection, class filter, warp-aggregated compaction.
 * Replaces yolo.py:259-277 + general.py:668-718 without materialising z. */
int dsn_filter_levels(const void *const *raw, int dtype, const int *ny, const int *nx, int nl, int bs, int na,
                      int no, float conf_thres, int multi_label, const uint64_t *class_mask,
                      uint64_t *keys, int64_t cap, int *counts, void *stream);

/* Confidence cut for huge candidate sets (conf 0.001 + multi_label: up to A*nc candidates per image, while the
 * greedy NMS stops after max_det survivors and never looks past the max_nms best, general.py:728-736).
 * dsn_conf_cut_levels histograms the confidences of a 1/64 sample of every image (one thread-block cluster per
 * image, histograms merged over distributed shared memory) and writes, per image, the confidence-key bound
 * cut_hi[b] above which about 2*target candidates are expected, and cut_dropped[b] = 1 when such a bound exists
 * (0: the image is expected to have fewer candidates, cut_hi[b] = 0xFFFFFFFF keeps everything).
 * dsn_filter_levels_cut then emits only candidates at or above that bound - exactly the best ones.  dsn_nms_levels
 * takes cut_dropped: an image whose NMS runs out of emitted candidates before it stops by itself (max_det
 * survivors or max_nms candidates consumed) although a cut was in force gets det_count = -2 (dropped candidates
 * could have mattered) and the caller re-runs without the cut (on the device: dsn_detect_nms_redo_levels); every other result is exactly
 * the uncut one.
 * Needs 16-byte aligned planes with ny*nx % 4 == 0 (DSN_E_UNSUPPORTED otherwise: run without a cut).
 *   target      candidates the NMS is expected to consume at most (e.g. 16*max_det), >= 1
 *   hist_ws     reserved, may be NULL (the histograms live in shared memory)
 *   cut_hi / cut_dropped   device (bs) uint32 / int32 */
int dsn_conf_cut_levels(const void *const *raw, int dtype, const int *ny, const int *nx, int nl, int bs, int na, int no,
                        float conf_thres, int multi_label, const uint64_t *class_mask, int target,
                        unsigned int *hist_ws, unsigned int *cut_hi, int *cut_dropped, void *stream);
int dsn_filter_levels_cut(const void *const *raw, int dtype, const int *ny, const int *nx, int nl, int bs, int na,
                          int no, float conf_thres, int multi_label, const uint64_t *class_mask,
                          uint64_t *keys, int64_t cap, int *counts, const unsigned int *cut_hi, void *stream);

/* Same filter on an already decoded prediction (bs, A, no) — the input of the reference's
 * non_max_suppression(prediction, ...) (general.py:659). */
int dsn_filter_z(const void *z, int dtype, int bs, int64_t A, int no, float conf_thres, int multi_label,
                 const uint64_t *class_mask, uint64_t *keys, int64_t cap, int *counts, void *stream);

/* Per-image NMS over the candidate keys: lazy descending-score selection (histogram select +
 * in-shared-memory sort of the next chunk), top-max_nms cut (general.py:728-729), class-offset
 * boxes (general.py:732-733), greedy IoU suppression with torchvision's CPU arithmetic
 * (call site general.py:734), stop at max_det (general.py:735-736), write rows
 * [x1,y1,x2,y2,conf,cls] (general.py:745).
 *   conf_thres the threshold the keys were produced with (dsn_filter_*): every key then satisfies
 *              conf > conf_thres, which bounds the key range the selection histogram covers
 *   iou_thres  compared as torchvision does: (double)iou_fp32 > iou_thres
 *   max_wh     class offset in pixels (4096, general.py:675); agnostic != 0 => offset 0
 *   det        (bs, max_det, 6) fp32, rows >= det_count[i] are left untouched
 *   det_count  (bs) int32; -1 when counts[i] > cap (keys were dropped: retry with a larger cap); -2 when a
 *              confidence cut was in force and the NMS ran out of emitted candidates (retry without the cut)
 *   cut_hi / cut_dropped (dsn_nms_levels only) device (bs) uint32 / int32 written by dsn_conf_cut_levels, or both NULL
 *              (cut_hi bounds the key range the selection looks at; cut_dropped arms the -2 report)
 * Box source is either the raw levels (dsn_nms_levels: decode on the fly, yolo.py:266-269 +
 * general.py:523-530) or a decoded prediction (dsn_nms_z). */
int dsn_nms_levels(const void *const *raw, int dtype, const int *ny, const int *nx, const float *stride,
                   const float *anchor_grid, int nl, int bs, int na, int no,
                   const uint64_t *keys, int64_t cap, const int *counts, float conf_thres,
                   double iou_thres, int agnostic, int max_nms, int max_det, float max_wh,
                   const unsigned int *cut_hi, const int *cut_dropped, float *det, int *det_count, void *stream);
/* Same, with the CTA size chosen by the caller: cta_threads = 512 (fastest per image), 256 (16 K registers per image: finds
 * room next to the short-lived CTAs of a concurrent streaming kernel) or 0 = dsn_nms_levels' own rule (256 when no
 * confidence cut is given, 512 otherwise).  Results are identical. */
int dsn_nms_levels_cta(const void *const *raw, int dtype, const int *ny, const int *nx, const float *stride,
                       const float *anchor_grid, int nl, int bs, int na, int no,
                       const uint64_t *keys, int64_t cap, const int *counts, float conf_thres,
                       double iou_thres, int agnostic, int max_nms, int max_det, float max_wh,
                       const unsigned int *cut_hi, const int *cut_dropped, float *det, int *det_count, int cta_threads,
                       void *stream);

/* A-priori labels on the raw-level path (general.py:690-696; scripts/val.py:228 passes them with save_hybrid): label j
 * of image b acts as prediction row A + j with obj = 1 and a one-hot class, i.e. one candidate of confidence 1.0.
 * dsn_append_label_keys adds those candidates to the keys of dsn_filter_levels (same stream, after it; cap must leave
 * room for max_labels more keys per image); dsn_nms_levels_labels is dsn_nms_levels with the label boxes as the box
 * source of rows >= A.
 *   labels       device (bs, max_labels, 5) fp32 rows [cls, x, y, w, h] in network-input pixels
 *   label_count  device (bs) int32 valid rows per image */
int dsn_append_label_keys(const float *labels, const int *label_count, int bs, int max_labels, int64_t A, int nc,
                          float conf_thres, const uint64_t *class_mask, uint64_t *keys, int64_t cap, int *counts,
                          void *stream);
int dsn_nms_levels_labels(const void *const *raw, int dtype, const int *ny, const int *nx, const float *stride,
                          const float *anchor_grid, int nl, int bs, int na, int no,
                          const uint64_t *keys, int64_t cap, const int *counts, float conf_thres,
                          double iou_thres, int agnostic, int max_nms, int max_det, float max_wh,
                          const float *labels, int max_labels, float *det, int *det_count, void *stream);

int dsn_nms_z(const void *z, int dtype, int bs, int64_t A, int no,
              const uint64_t *keys, int64_t cap, const int *counts, float conf_thres,
              double iou_thres, int agnostic, int max_nms, int max_det, float max_wh,
              float *det, int *det_count, void *stream);

/* Fused detection path: dsn_filter_levels + dsn_nms_levels in ONE launch, one CTA per image
 * (objectness scan -> shared-memory row list -> class pass -> keys -> lazy NMS).  Same results as the
 * two calls; `keys` (bs, cap) and `counts` (bs) are still written (workspace + candidate counts).
 * Replaces yolo.py:259-277 + general.py:659-750 for callers that hold the raw head outputs. */
int dsn_detect_nms_levels(const void *const *raw, int dtype, const int *ny, const int *nx, const float *stride,
                          const float *anchor_grid, int nl, int bs, int na, int no,
                          float conf_thres, int multi_label, const uint64_t *class_mask,
                          double iou_thres, int agnostic, int max_nms, int max_det, float max_wh,
                          uint64_t *keys, int64_t cap, int *counts, float *det, int *det_count, void *stream);

/* On-device redo for the confidence cut: the same kernel, but only images whose det_count[b] == -2 (the NMS of
 * dsn_nms_levels ran out of emitted candidates while a cut was in force) are processed - uncut filter + NMS by one
 * CTA per flagged image; every other CTA exits at once.  Enqueued right after dsn_nms_levels it makes the
 * cut -> filter -> NMS chain exact without a host round trip (no det_count == -2 ever reaches the caller), so the
 * chain can be captured in a CUDA graph.  keys / counts / det / det_count are the buffers of the preceding calls. */
int dsn_detect_nms_redo_levels(const void *const *raw, int dtype, const int *ny, const int *nx, const float *stride,
                               const float *anchor_grid, int nl, int bs, int na, int no,
                               float conf_thres, int multi_label, const uint64_t *class_mask,
                               double iou_thres, int agnostic, int max_nms, int max_det, float max_wh,
                               uint64_t *keys, int64_t cap, int *counts, float *det, int *det_count, void *stream);

/* ------------------------------------------------------------------------------------------
 * Segmentation tail.
 * ------------------------------------------------------------------------------------------ */

/* Fused  nn.Upsample(bilinear, align_corners=True) (core/models/yolo.py:183)
 *      + argmax over classes, lowest index on ties (core/utils/plots.py:228,
 *        core/utils/metrics.py:356,375)
 *      + optional mask write, nearest-resized to (mask_h, mask_w) (plots.py:229)
 *      + optional confusion-matrix accumulation against an int64 target
 *        (closed form of metrics.py:350-388).
 * logits   (bs, C, h, w) LOW-RES head output (input of the final Upsample), elements of type `dtype`
 * H, W     full-resolution size the reference would upsample to (8h, 8w)
 * target   (bs, H, W) int64 or NULL.  Values are wrapped to int32 as metrics.py:359,380 do,
 *          then clamped into [0, cm_rows-1] (negatives -> 0: neither labeled nor a class).
 * mask_f32 (bs, mask_h, mask_w) fp32 or NULL;  mask_u8 same shape uint8 or NULL
 * cm       (cm_slots, cm_rows, C) uint64 or NULL, ACCUMULATED into (caller zeroes once);
 *          image b adds into slot b % cm_slots; the confusion matrix is the sum over slots.
 *          cm[r][p] = #pixels with clamped target r and predicted class p.
 */
int dsn_seg_fused(const void *logits, int dtype, int bs, int C, int h, int w, int H, int W,
                  const int64_t *target, float *mask_f32, uint8_t *mask_u8, int mask_h, int mask_w,
                  unsigned long long *cm, int cm_slots, int cm_rows, void *stream);

/* Same outputs from FULL-RES logits (bs, C, H, W) — the tensor the reference's
 * segoutput_to_target / batch_pix_accuracy / batch_intersection_union receive. */
int dsn_seg_fullres(const void *logits, int dtype, int bs, int C, int H, int W,
                    const int64_t *target, float *mask_f32, uint8_t *mask_u8, int mask_h, int mask_w,
                    unsigned long long *cm, int cm_slots, int cm_rows, void *stream);

/* Producer of the seg tail's input (first slice; opt-in because it changes the boundary): the
 * seg head's last convolution nn.Conv2d(c_hid, n_segcls, 1) + bias (core/models/yolo.py:182) as a streaming kernel.
 *   features (bs, K, h, w) NCHW, elements of type `dtype`;  weight (C, K) fp32, bias (C) fp32 or NULL;  C <= 4
 *   logits   (bs, C, h, w) fp32 = the low-res logits dsn_seg_fused reads
 * fp32 accumulation in channel order, one FMA per term (cuDNN's summation order and, under --half, its fp16 rounding of the
 * result differ: parity with the reference's conv is a tolerance, not bit-exactness). */
int dsn_seg_head_conv1x1(const void *features, int dtype, const float *weight, const float *bias, int bs, int K, int C,
                         int h, int w, float *logits, void *stream);

/* Producer of the detection tail's input (opt-in because it changes the boundary): one
 * Detect output convolution nn.Conv2d(ch[i], na * no, 1) + bias (core/models/yolo.py:252, applied at :258).
 *   features (bs, K, ny, nx) NCHW, elements of type `dtype`;  weight (N, K) row-major, elements of type `dtype`;
 *   bias (N) fp32 or NULL;  out (bs, N, ny, nx) NCHW, elements of type `dtype` = raw level i of dsn_filter_levels / dsn_nms_levels
 * DSN_F16 / DSN_BF16: tcgen05 tensor cores, fp32 accumulation in TMEM, result rounded to `dtype` (what the reference's
 * cuDNN convolution does under --half); needs K % 64 == 0, (ny * nx) % 8 == 0, N <= 256 and N_pad * K * 2 + 50 176 bytes of
 * shared memory <= 200 KB (DSN_E_UNSUPPORTED otherwise: keep the framework's convolution for that level).
 * DSN_F32: exact fp32 FMA chain in channel order on the CUDA cores (TF32 would break the 1e-5 box tolerance).
 * Parity with the reference's convolution is a tolerance (summation order), not bit-exactness. */
int dsn_detect_head_conv1x1(const void *features, int dtype, const void *weight, const float *bias, int bs, int K, int N,
                            int ny, int nx, void *out, void *stream);

/* nn.Upsample(bilinear, align_corners=True) alone (core/models/yolo.py:183): (planes,h,w) ->
 * (planes,H,W), bit-identical to ATen's CPU kernel (fma(l0,a,l1*b) lerps, horizontal first). */
int dsn_upsample_bilinear_ac(const void *in, int dtype, int64_t planes, int h, int w, int H, int W,
                             float *out, void *stream);

/* ------------------------------------------------------------------------------------------
 * Next row of the scope table: the step every caller runs right after
 * non_max_suppression.
 * ------------------------------------------------------------------------------------------ */

/* scale_coords + clip_coords (core/utils/general.py:598-623), in place: box rows [x1,y1,x2,y2,...] are
 * mapped from the letterboxed network input back to the original image, (x - pad) / gain in fp32, and
 * clamped to [0, w0] x [0, h0]; do_round != 0 additionally rounds half-to-even (scripts/detect.py:218).
 *   boxes        device, image b's row i starts at boxes + b*image_stride + i*row_stride (floats)
 *   counts       device (bs) int32 valid rows per image, or NULL = rows_per_image for every image
 *   params       device (bs, 5) fp32: pad_x, pad_y, gain, w0, h0 — the reference computes them in Python
 *                (general.py:600-605) and applies them as fp32 scalars */
int dsn_scale_coords(float *boxes, int64_t row_stride, int64_t image_stride, const int *counts, int bs,
                     int rows_per_image, const float *params, int do_round, void *stream);

/* box_iou (core/utils/general.py:626-648): out (n, m) fp32 = IoU of every box1 row with every box2 row
 * (xyxy, rows `stride` floats apart), evaluated with the reference's fp32 tensor arithmetic. */
int dsn_box_iou(const float *box1, int64_t stride1, int n, const float *box2, int64_t stride2, int m,
                float *out, void *stream);

/* process_batch (scripts/val.py:101-122) for a whole batch, one CTA per image:
 *   det          device, image b's detection row i at det + b*det_image_stride + i*det_row_stride: [x1,y1,x2,y2,conf,cls]
 *   det_count    device (bs) int32 valid rows (NULL = max_rows everywhere)
 *   labels       device (L, 5) fp32 rows [cls, x1, y1, x2, y2] of all images back to back
 *   label_off    device (bs+1) int32: image b owns label rows [label_off[b], label_off[b+1])
 *   iouv         device (niou) fp32 IoU thresholds (val.py:188: linspace(0.5, 0.95, 10))
 *   correct      device (bs, max_rows, niou) uint8 (0/1); rows >= det_count[b] are left untouched
 * Same matching as the reference (per detection its best class-matching label with IoU >= iouv[0], then per label
 * the first such detection in detection order); equal IoUs of one detection resolve to the lowest label index. */
int dsn_process_batch(const float *det, int64_t det_image_stride, int64_t det_row_stride, const int *det_count,
                      int bs, int max_rows, const float *labels, const int *label_off, int max_labels_per_image,
                      const float *iouv, int niou, unsigned char *correct, void *stream);

/* All labels of a batch into native image space in one launch (scripts/val.py:227,243-255):
 *   targets   device (n, 6) fp32 rows [image, cls, x, y, w, h], xywh normalised to the network input
 *   width, height   network input size (val.py:227: targets[:, 2:] *= (width, height, width, height))
 *   params    device (bs, 5) fp32 per image: pad_x, pad_y, gain, w0, h0 (as dsn_scale_coords)
 *   labelsn   device (n, 5) fp32 rows [cls, x1, y1, x2, y2]: xywh2xyxy (general.py:523-530), then scale_coords +
 *             clip_coords (general.py:598-623) with the row's image parameters */
int dsn_labels_to_native(const float *targets, int64_t n, float width, float height, const float *params, int bs,
                         float *labelsn, void *stream);

/* ---- end-of-eval detection statistics ------------------------------------------------
 * ap_per_class (core/utils/metrics.py:31-92) over ALL detections of an evaluation, on the device:
 *   tp          device (n, niou) uint8/bool: detection i is a true positive at IoU threshold j (val.py:121,263)
 *   conf        device (n) fp32 confidences; pred_cls device (n) fp32 predicted classes
 *   classes     device (ncls) fp32: the unique TARGET classes, ascending (np.unique(target_cls), metrics.py:52)
 *   n_labels    device (ncls) int64: labels per class (metrics.py:59)
 *   x101, px    device float64: np.linspace(0, 1, 101) and np.linspace(0, 1, npx) (npx = 1000, metrics.py:56,114)
 *   workspace   device, 256-byte aligned, dsn_ap_workspace_bytes(n, niou, ncls) bytes
 *   ap          device (ncls, niou) float64 = metrics.py:79;  p, r device (ncls, npx) float64 = metrics.py:71,75
 * Bit-identical to the reference's numpy float64 results (np.interp / np.trapz restated operation by operation);
 * equal confidences keep their input order where the reference's quicksort (metrics.py:48) leaves it open.
 * Limits: n < 2^24, ncls < 255, niou <= 16.  The F1 curve / its argmax (metrics.py:85-91) are a (ncls, npx)
 * elementwise tail the host mirror evaluates from p and r. */
int64_t dsn_ap_workspace_bytes(int64_t n, int niou, int ncls);
int dsn_ap_per_class(const uint8_t *tp, const float *conf, const float *pred_cls, int64_t n, int niou,
                     const float *classes, const int64_t *n_labels, int ncls, const double *x101, const double *px,
                     int npx, void *workspace, int64_t workspace_bytes, double *ap, double *p, double *r, void *stream);

/* compute_ap (metrics.py:95-121, 'interp' method) of one recall / precision curve (device float64 arrays of n
 * points): precision envelope into env_ws (n doubles), 101-point interpolation, trapezoid sum -> ap[0]. */
int dsn_compute_ap(const double *recall, const double *precision, int64_t n, const double *x101, double *env_ws,
                   double *ap, void *stream);

/* ConfusionMatrix.process_batch (metrics.py:132-170) for a whole batch, one CTA per image; det / det_count / labels /
 * label_off as for dsn_process_batch.  matrix: device (nc+1, nc+1) uint64, rows = predicted class (nc = background),
 * columns = true class, ACCUMULATED.  Detections with conf <= `conf` are ignored, pairs need IoU > iou_thres (both
 * compared in fp32 as torch does); equal IoUs resolve to the lowest index (the reference's order there depends on
 * numpy's unstable argsort); classes outside [0, nc] are skipped (the reference would raise or wrap).
 * skip_empty != 0: images without detections contribute nothing, as in val.run, whose loop `continue`s before
 * process_batch for them (val.py:241-244); 0: their labels count as background FP, as a direct call would. */
int dsn_det_confusion_matrix(const float *det, int64_t det_image_stride, int64_t det_row_stride, const int *det_count,
                             int bs, int max_rows, const float *labels, const int *label_off,
                             int max_labels_per_image, int nc, float conf, float iou_thres, int skip_empty,
                             uint64_t *matrix, void *stream);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU detection gather over peer memory.  One process per GPU; the reference never shards
 * evaluation (scripts/val.py:170-172), so this is new: the gather root owns the gathered (global_bs, max_det, 6) /
 * (global_bs) buffers, every rank passes pointers into ITS slice of them as `det` / `det_count` of dsn_nms_levels, and
 * the NMS kernel's result rows go over NVLink as they are produced.  These five calls are host-side set-up (they are
 * NOT enqueue-only: they allocate / map memory on the current device):
 *   dsn_peer_alloc   cudaMalloc'ed, zero-filled buffer (CUDA IPC needs a cudaMalloc allocation of its own)
 *   dsn_peer_export  64-byte CUDA IPC handle of such a buffer, to be sent to the other processes
 *   dsn_peer_open    map another process's buffer into this one (peer access enabled lazily); dsn_peer_close unmaps */
int dsn_peer_alloc(int64_t bytes, void **ptr);
int dsn_peer_free(void *ptr);
int dsn_peer_export(const void *ptr, unsigned char *handle64);
int dsn_peer_open(const unsigned char *handle64, void **ptr);
int dsn_peer_close(void *ptr);
/* Enqueue-only.  dsn_peer_signal: after everything enqueued before it on `stream`, increment *local_step (device
 * word of this rank) and publish the new value to *remote_flag (a word of the root's buffer) with a system-scope
 * release store.  dsn_peer_wait: on the root, spin (acquire loads) until flags[i] - targets[i] >= 0 (both device
 * uint32 arrays) for every i < n, for at most timeout_ms; on a timeout *status (device int, caller-zeroed) becomes
 * 1 + index of a missing flag. */
int dsn_peer_signal(unsigned int *local_step, unsigned int *remote_flag, void *stream);
int dsn_peer_wait(const unsigned int *flags, const unsigned int *targets, int n, int64_t timeout_ms, int *status, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DESENET_B200_H_ */
