// End-of-eval detection statistics:
//   * ap_per_class / compute_ap        core/utils/metrics.py:31-121
//   * ConfusionMatrix.process_batch    core/utils/metrics.py:132-170
// ap_per_class runs once per evaluation over ALL detections (n ~ 300 x images): sort by confidence, per class
// cumulative TP counts for every IoU threshold, precision envelope, 101-point interpolated AP, and the 1000-point
// precision / recall curves.  Everything stays on the device: a bitonic sort of 64-bit keys
// (class slot | descending confidence | detection index), integer scans (exact in any order), a reverse max-scan
// (exact in any order), and the float64 interpolation / trapezoid arithmetic of np.interp / np.trapz restated
// operation by operation with explicit round-to-nearest intrinsics (no FMA contraction), so the results are
// bit-identical to the reference's numpy ones (oracle/port.py: interp_scalar, pairwise_sum_small, compute_ap).
#include <cstring>

#include "common.cuh"

namespace dsn {

constexpr int kSortTile = 2048;          // keys sorted / merged per CTA in shared memory
constexpr int kSortThreads = 1024;
constexpr int kNoClass = 255;            // class slot of detections whose class has no label
constexpr int kMaxIou = 16;

// ---- keys -------------------------------------------------------------------------------------------------------
// ascending key order = class slot ascending, then confidence DESCENDING (np.argsort(-conf), metrics.py:48), then
// detection index ascending (equal confidences keep their input order; the reference's quicksort leaves it open)
__global__ void __launch_bounds__(256) ap_keys_kernel(const float *__restrict__ conf, const float *__restrict__ pred_cls,
                                                      long long n, long long npad, const float *__restrict__ classes,
                                                      int ncls, unsigned long long *__restrict__ keys) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    if (i >= n) { keys[i] = ~0ull; return; }
    const float c = pred_cls[i];
    int lo = 0, hi = ncls;                                    // first class >= c
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (classes[mid] < c) lo = mid + 1; else hi = mid;
    }
    const unsigned long long slot = (lo < ncls && classes[lo] == c) ? (unsigned long long)lo : (unsigned long long)kNoClass;
    const unsigned int b = __float_as_uint(-conf[i]);
    const unsigned int u = (b & 0x80000000u) ? ~b : (b | 0x80000000u);          // ascending in -conf
    keys[i] = (slot << 56) | ((unsigned long long)u << 24) | (unsigned long long)i;
}

// ---- bitonic sort -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bitonic_smem_steps(unsigned long long *a, long long base, int k_first, int k_last, int j_first) {
    // runs the (k, j) steps with j < kSortTile on the tile a[0..kSortTile) whose first global index is `base`
    for (int k = k_first; k <= k_last; k <<= 1) {
        for (int j = (k == k_first ? j_first : k >> 1); j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < kSortTile / 2; t += kSortThreads) {
                const int ix = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int pj = ix | j;
                const unsigned long long x = a[ix], y = a[pj];
                const bool up = (((base + ix) / k) & 1) == 0;
                if ((x > y) == up) { a[ix] = y; a[pj] = x; }
            }
            __syncthreads();
        }
    }
}

// sorts every tile completely (k = 2 .. kSortTile)
__global__ void __launch_bounds__(kSortThreads) bitonic_tile_sort_kernel(unsigned long long *__restrict__ keys) {
    __shared__ unsigned long long a[kSortTile];
    const long long base = (long long)blockIdx.x * kSortTile;
    for (int t = threadIdx.x; t < kSortTile; t += kSortThreads) a[t] = keys[base + t];
    __syncthreads();
    bitonic_smem_steps(a, base, 2, kSortTile, 1);
    for (int t = threadIdx.x; t < kSortTile; t += kSortThreads) keys[base + t] = a[t];
}

// one (k, j) step with j >= kSortTile
__global__ void __launch_bounds__(256) bitonic_global_step_kernel(unsigned long long *__restrict__ keys, long long half, long long j,
                                                                   long long k) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= half) return;
    const long long ix = ((t & ~(j - 1)) << 1) | (t & (j - 1));
    const long long pj = ix | j;
    const unsigned long long x = keys[ix], y = keys[pj];
    const bool up = ((ix / k) & 1) == 0;
    if ((x > y) == up) { keys[ix] = y; keys[pj] = x; }
}

// the remaining steps j = kSortTile/2 .. 1 of merge level k (> kSortTile)
__global__ void __launch_bounds__(kSortThreads) bitonic_tile_merge_kernel(unsigned long long *__restrict__ keys, long long k) {
    __shared__ unsigned long long a[kSortTile];
    const long long base = (long long)blockIdx.x * kSortTile;
    for (int t = threadIdx.x; t < kSortTile; t += kSortThreads) a[t] = keys[base + t];
    __syncthreads();
    const bool up = ((base / k) & 1) == 0;                   // uniform for the whole tile since k > kSortTile
    for (int j = kSortTile / 2; j > 0; j >>= 1) {
        for (int t = threadIdx.x; t < kSortTile / 2; t += kSortThreads) {
            const int ix = ((t & ~(j - 1)) << 1) | (t & (j - 1));
            const int pj = ix | j;
            const unsigned long long x = a[ix], y = a[pj];
            if ((x > y) == up) { a[ix] = y; a[pj] = x; }
        }
        __syncthreads();
    }
    for (int t = threadIdx.x; t < kSortTile; t += kSortThreads) keys[base + t] = a[t];
}

// ---- per-class cumulative TP counts ---------------------------------------------------------------------------------
__device__ __forceinline__ long long lower_bound_key(const unsigned long long *keys, long long n, unsigned long long v) {
    long long lo = 0, hi = n;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (keys[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// One CTA per class: tpc[pos][j] = number of true positives at IoU threshold j among the class's detections up to
// and including sorted position pos (metrics.py:66: tp[i].cumsum(0)); conf_s[pos] = the detection's confidence.
__global__ void __launch_bounds__(256) ap_scan_kernel(const unsigned long long *__restrict__ keys, long long n,
                                                      const unsigned char *__restrict__ tp, const float *__restrict__ conf, int niou,
                                                      int *__restrict__ tpc, float *__restrict__ conf_s, long long *__restrict__ seg) {
    __shared__ int warp_tot[8][kMaxIou];
    __shared__ int carry[kMaxIou];
    const int c = blockIdx.x;
    const long long s = lower_bound_key(keys, n, (unsigned long long)c << 56);
    const long long e = lower_bound_key(keys, n, (unsigned long long)(c + 1) << 56);
    if (threadIdx.x == 0) { seg[2 * c] = s; seg[2 * c + 1] = e; }
    if (threadIdx.x < kMaxIou) carry[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long p0 = s; p0 < e; p0 += 256) {
        const long long pos = p0 + threadIdx.x;
        const bool in = pos < e;
        long long idx = 0;
        if (in) {
            idx = (long long)(keys[pos] & 0xFFFFFFull);
            conf_s[pos] = conf[idx];
        }
        int v[kMaxIou];
#pragma unroll
        for (int j = 0; j < kMaxIou; ++j) {
            v[j] = (in && j < niou && tp[idx * niou + j]) ? 1 : 0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, v[j], d);
                if (lane >= d) v[j] += up;
            }
            if (lane == 31) warp_tot[warp][j] = v[j];
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kMaxIou; ++j) {
            int add = carry[j];
            for (int w = 0; w < warp; ++w) add += warp_tot[w][j];
            v[j] += add;
            if (in && j < niou) tpc[pos * niou + j] = v[j];
        }
        __syncthreads();
        if (threadIdx.x == 255)
#pragma unroll
            for (int j = 0; j < kMaxIou; ++j) carry[j] = v[j];
        __syncthreads();
    }
}

// ---- curves ---------------------------------------------------------------------------------------------------------
// recall / precision at sorted position i of a class segment (metrics.py:70-75); every operation is IEEE binary64
// round-to-nearest exactly as numpy evaluates it: tpc / (n_l + 1e-16) and tpc / (tpc + fpc) with tpc + fpc = i + 1
struct CurveCounts {
    const int *tpc;          // segment start, threshold j: tpc[i * niou]
    int niou;
    double denom;            // n_l + 1e-16
    __device__ __forceinline__ double recall(long long i) const { return __ddiv_rn((double)tpc[i * niou], denom); }
    __device__ __forceinline__ double precision(long long i) const { return __ddiv_rn((double)tpc[i * niou], (double)(i + 1)); }
};
struct CurveArrays {         // compute_ap on explicit recall / precision arrays (metrics.py:95)
    const double *rec, *pre;
    __device__ __forceinline__ double recall(long long i) const { return rec[i]; }
    __device__ __forceinline__ double precision(long long i) const { return pre[i]; }
};

// precision envelope: env[i] = max(precision[i], precision[i+1], ..., 0) (metrics.py:110, reverse maximum.accumulate
// of [1, precision, 0]); max is exact, so the scan order does not matter.  Grid (curves), one CTA per curve.
template <class Curve>
__device__ void envelope_curve(const Curve &cv, long long len, double *env, long long env_stride) {
    __shared__ double warp_max[8];
    __shared__ double carry;
    if (threadIdx.x == 0) carry = 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long hi = len; hi > 0; hi -= 256) {                 // chunk [hi-256, hi), thread 0 = highest position
        const long long i = hi - 1 - threadIdx.x;
        double v = i >= 0 ? cv.precision(i) : 0.0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v = fmax(v, up);
        }
        if (lane == 31) warp_max[warp] = v;
        __syncthreads();
        double m = carry;
        for (int w = 0; w < warp; ++w) m = fmax(m, warp_max[w]);
        v = fmax(v, m);
        if (i >= 0) env[i * env_stride] = v;
        __syncthreads();
        if (threadIdx.x == 255) carry = v;
        __syncthreads();
    }
}

// np.interp for one x over a virtual sorted array xp(0..len) with values fp(0..len) (numpy arr_interp: j = last index
// with xp[j] <= x, slope form, NaN retries) - see oracle/port.py:interp_scalar
template <class XP, class FP>
__device__ __forceinline__ double interp_one(double x, long long len, const XP &xp, const FP &fp, double left, double right) {
    if (x != x) return x;
    if (x > xp(len - 1)) return right;
    if (x < xp(0)) return left;
    long long lo = 0, hi = len;                                  // first index with xp > x
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (xp(mid) <= x) lo = mid + 1; else hi = mid;
    }
    const long long j = lo - 1;
    if (j == len - 1) return fp(j);
    const double xj = xp(j), fj = fp(j);
    if (xj == x) return fj;
    const double xj1 = xp(j + 1), fj1 = fp(j + 1);
    const double slope = __ddiv_rn(__dsub_rn(fj1, fj), __dsub_rn(xj1, xj));
    double res = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, xj)), fj);
    if (res != res) {
        res = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, xj1)), fj1);
        if (res != res && fj == fj1) res = fj;
    }
    return res;
}

// compute_ap (metrics.py:95-121, 'interp'): 101-point interpolation of the envelope over [0, recall, 1] and
// np.trapz, whose sum is numpy's pairwise add.reduce of 100 terms (8 accumulators, then the 4-element remainder).
template <class Curve>
__device__ double integrate_curve(const Curve &cv, long long len, const double *env, long long env_stride,
                                  const double *__restrict__ x101) {
    __shared__ double y[101];
    __shared__ double result;
    if (threadIdx.x < 101) {
        auto xp = [&](long long i) -> double { return i == 0 ? 0.0 : (i == len + 1 ? 1.0 : cv.recall(i - 1)); };
        auto fp = [&](long long i) -> double {
            return i == 0 ? fmax(1.0, len > 0 ? env[0] : 0.0) : (i == len + 1 ? 0.0 : env[(i - 1) * env_stride]);
        };
        y[threadIdx.x] = interp_one(x101[threadIdx.x], len + 2, xp, fp, fp(0), 0.0);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double r[8];
        auto term = [&](int k) { return __ddiv_rn(__dmul_rn(__dsub_rn(x101[k + 1], x101[k]), __dadd_rn(y[k + 1], y[k])), 2.0); };
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = term(k);
        int i = 8;
        for (; i < 96; i += 8)
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], term(i + k));
        double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])), __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        for (; i < 100; ++i) res = __dadd_rn(res, term(i));
        result = res;
    }
    __syncthreads();
    return result;
}

// grid (ncls, niou): envelope of every (class, threshold) curve into env[pos][j]
__global__ void __launch_bounds__(256) ap_envelope_kernel(const int *__restrict__ tpc, int niou, const long long *__restrict__ seg,
                                                          const long long *__restrict__ n_labels, double *__restrict__ env) {
    const int c = blockIdx.x, j = blockIdx.y;
    const long long s = seg[2 * c], e = seg[2 * c + 1];
    if (e <= s || n_labels[c] <= 0) return;
    CurveCounts cv{tpc + s * niou + j, niou, __dadd_rn((double)n_labels[c], 1e-16)};
    envelope_curve(cv, e - s, env + s * niou + j, niou);
}

__global__ void __launch_bounds__(128) ap_integrate_kernel(const int *__restrict__ tpc, int niou, const long long *__restrict__ seg,
                                                           const long long *__restrict__ n_labels, const double *__restrict__ env,
                                                           const double *__restrict__ x101, double *__restrict__ ap) {
    const int c = blockIdx.x, j = blockIdx.y;
    const long long s = seg[2 * c], e = seg[2 * c + 1];
    if (e <= s || n_labels[c] <= 0) {                              // metrics.py:62-63: the row stays zero
        if (threadIdx.x == 0) ap[(size_t)c * niou + j] = 0.0;
        return;
    }
    CurveCounts cv{tpc + s * niou + j, niou, __dadd_rn((double)n_labels[c], 1e-16)};
    const double a = integrate_curve(cv, e - s, env + s * niou + j, niou, x101);
    if (threadIdx.x == 0) ap[(size_t)c * niou + j] = a;
}

// precision / recall at the npx confidence levels px (metrics.py:71,75: np.interp(-px, -conf[i], curve[:, 0]))
__global__ void __launch_bounds__(256) ap_pr_kernel(const int *__restrict__ tpc, int niou, const float *__restrict__ conf_s,
                                                    const long long *__restrict__ seg, const long long *__restrict__ n_labels,
                                                    const double *__restrict__ px, int npx, double *__restrict__ p,
                                                    double *__restrict__ r) {
    const int c = blockIdx.x;
    const long long s = seg[2 * c], e = seg[2 * c + 1];
    const long long len = e - s;
    const bool empty = len <= 0 || n_labels[c] <= 0;
    CurveCounts cv{tpc + s * niou, niou, empty ? 1.0 : __dadd_rn((double)n_labels[c], 1e-16)};
    const float *cs = conf_s + s;
    auto xp = [&](long long i) -> double { return -(double)cs[i]; };
    auto fr = [&](long long i) -> double { return cv.recall(i); };
    auto fpr = [&](long long i) -> double { return cv.precision(i); };
    for (int t = blockIdx.y * blockDim.x + threadIdx.x; t < npx; t += gridDim.y * blockDim.x) {
        double rv = 0.0, pv = 0.0;
        if (!empty) {
            const double x = -px[t];
            rv = interp_one(x, len, xp, fr, 0.0, fr(len - 1));
            pv = interp_one(x, len, xp, fpr, 1.0, fpr(len - 1));
        }
        r[(size_t)c * npx + t] = rv;
        p[(size_t)c * npx + t] = pv;
    }
}

// compute_ap on one explicit curve
__global__ void __launch_bounds__(256) compute_ap_kernel(const double *__restrict__ rec, const double *__restrict__ pre, long long len,
                                                         const double *__restrict__ x101, double *__restrict__ env,
                                                         double *__restrict__ ap) {
    CurveArrays cv{rec, pre};
    envelope_curve(cv, len, env, 1);
    __syncthreads();
    const double a = integrate_curve(cv, len, env, 1, x101);
    if (threadIdx.x == 0) ap[0] = a;
}

// ---- detection confusion matrix -------------------------------------------------------------------------------------
__device__ __forceinline__ float cm_min(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fminf(a, b); }
__device__ __forceinline__ float cm_max(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b); }
__device__ __forceinline__ float cm_iou(const float *a, const float *b) {                 // general.py:640-648
    const float area1 = __fmul_rn(__fsub_rn(a[2], a[0]), __fsub_rn(a[3], a[1]));
    const float area2 = __fmul_rn(__fsub_rn(b[2], b[0]), __fsub_rn(b[3], b[1]));
    float w = __fsub_rn(cm_min(a[2], b[2]), cm_max(a[0], b[0]));
    float h = __fsub_rn(cm_min(a[3], b[3]), cm_max(a[1], b[1]));
    w = w != w ? w : fmaxf(w, 0.f);
    h = h != h ? h : fmaxf(h, 0.f);
    const float inter = __fmul_rn(w, h);
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area1, area2), inter));
}

// One CTA per image (metrics.py:132-170).  Every detection above `conf` picks its best-IoU label among IoU >
// iou_thres (np.unique on the detection column of the IoU-sorted matches), every label keeps its best-IoU such
// detection (np.unique on the label column); equal IoUs resolve to the lowest index.  matrix is (nc+1, nc+1) int64,
// rows = predicted class (nc = background), columns = true class, accumulated with atomics.
__global__ void __launch_bounds__(256) det_confusion_kernel(const float *__restrict__ det, long long det_image_stride,
                                                            long long det_row_stride, const int *__restrict__ det_count,
                                                            int max_rows, const float *__restrict__ labels,
                                                            const int *__restrict__ label_off, int nc, float conf, float iou_thres,
                                                            int skip_empty, unsigned long long *__restrict__ matrix) {
    extern __shared__ unsigned long long win[];     // per label: (iou bits << 32) | ~detection of its best detection
    __shared__ int any_match;
    const int b = blockIdx.x;
    const int n = det_count ? min(det_count[b], max_rows) : max_rows;
    if (skip_empty && n <= 0) return;               // val.py:241-244: images without detections never reach process_batch
    const int l0 = label_off[b], m = label_off[b + 1] - l0;
    const float *lab = labels + (size_t)l0 * 5;
    const float *d0 = det + (size_t)b * det_image_stride;
    for (int i = threadIdx.x; i < m; i += blockDim.x) win[i] = 0ull;
    if (threadIdx.x == 0) any_match = 0;
    __syncthreads();
    constexpr int kPer = 8;
    int best_l[kPer];
    bool valid[kPer];
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        best_l[q] = -1;
        const int d = threadIdx.x + q * blockDim.x;
        valid[q] = false;
        if (d < n) {
            const float *dr = d0 + (size_t)d * det_row_stride;
            valid[q] = dr[4] > conf;                                               // metrics.py:143
            if (valid[q]) {
                const float box[4] = {dr[0], dr[1], dr[2], dr[3]};
                float best = 0.f;
                for (int l = 0; l < m; ++l) {
                    const float iou = cm_iou(lab + l * 5 + 1, box);
                    if (iou > iou_thres && (best_l[q] < 0 || iou > best)) { best_l[q] = l; best = iou; }   // metrics.py:148
                }
                if (best_l[q] >= 0) {
                    atomicMax(&win[best_l[q]], ((unsigned long long)__float_as_uint(best) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)d));
                    any_match = 1;
                }
            }
        }
    }
    __syncthreads();
    const int w = nc + 1;
    for (int l = threadIdx.x; l < m; l += blockDim.x) {
        const int gc = (int)lab[l * 5];
        if (gc < 0 || gc > nc) continue;
        int row = nc;                                                              // background FP, metrics.py:163
        if (win[l] != 0ull) {
            const unsigned d = 0xFFFFFFFFu - (unsigned)(win[l] & 0xFFFFFFFFull);
            const int dc = (int)d0[(size_t)d * det_row_stride + 5];
            if (dc < 0 || dc > nc) continue;
            row = dc;                                                              // metrics.py:161
        }
        atomicAdd(&matrix[(size_t)row * w + gc], 1ull);
    }
    if (any_match) {                                                               // metrics.py:165
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            const int d = threadIdx.x + q * blockDim.x;
            if (d < n && valid[q]) {
                const bool matched = best_l[q] >= 0 && (0xFFFFFFFFu - (unsigned)(win[best_l[q]] & 0xFFFFFFFFull)) == (unsigned)d;
                const int dc = (int)d0[(size_t)d * det_row_stride + 5];
                if (!matched && dc >= 0 && dc <= nc) atomicAdd(&matrix[(size_t)dc * w + nc], 1ull);   // metrics.py:168
            }
        }
    }
}

static inline long long pow2_at_least(long long n, long long lo) {
    long long p = lo;
    while (p < n) p <<= 1;
    return p;
}

struct ApWorkspace {
    unsigned long long *keys;
    long long *seg;
    float *conf_s;
    int *tpc;
    double *env;
    long long npad;
    long long bytes;
};

static inline ApWorkspace carve_ap_workspace(void *base, long long n, int niou, int ncls) {
    ApWorkspace w;
    w.npad = pow2_at_least(n > 0 ? n : 1, kSortTile);
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_keys = take(sizeof(unsigned long long) * (size_t)w.npad);
    const size_t o_seg = take(sizeof(long long) * 2 * (size_t)(ncls > 0 ? ncls : 1));
    const size_t o_conf = take(sizeof(float) * (size_t)(n > 0 ? n : 1));
    const size_t o_tpc = take(sizeof(int) * (size_t)(n > 0 ? n : 1) * niou);
    const size_t o_env = take(sizeof(double) * (size_t)(n > 0 ? n : 1) * niou);
    w.bytes = (long long)off;
    char *b = (char *)base;
    w.keys = (unsigned long long *)(b + o_keys);
    w.seg = (long long *)(b + o_seg);
    w.conf_s = (float *)(b + o_conf);
    w.tpc = (int *)(b + o_tpc);
    w.env = (double *)(b + o_env);
    return w;
}

}  // namespace dsn

using namespace dsn;

extern "C" __attribute__((visibility("default"))) int64_t dsn_ap_workspace_bytes(int64_t n, int niou, int ncls) {
    if (n < 0 || niou < 1 || niou > kMaxIou || ncls < 0) return -1;
    return carve_ap_workspace(nullptr, n, niou, ncls).bytes;
}

extern "C" __attribute__((visibility("default"))) int dsn_ap_per_class(const uint8_t *tp, const float *conf, const float *pred_cls,
                                                                       int64_t n, int niou, const float *classes,
                                                                       const int64_t *n_labels, int ncls, const double *x101,
                                                                       const double *px, int npx, void *workspace,
                                                                       int64_t workspace_bytes, double *ap, double *p, double *r,
                                                                       void *stream) {
    if (n < 0 || niou < 1 || ncls < 0 || npx < 0) return DSN_E_INVALID;
    if (niou > kMaxIou || ncls >= kNoClass || n >= (1ll << 24)) return DSN_E_UNSUPPORTED;
    if (ncls == 0) return DSN_OK;
    if (!classes || !n_labels || !x101 || !ap || !workspace || (npx > 0 && (!px || !p || !r))) return DSN_E_INVALID;
    if (n > 0 && (!tp || !conf || !pred_cls)) return DSN_E_INVALID;
    if (((uintptr_t)workspace & 255) != 0) return DSN_E_ALIGN;
    const ApWorkspace w = carve_ap_workspace(workspace, n, niou, ncls);
    if (workspace_bytes < w.bytes) return DSN_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    ap_keys_kernel<<<(unsigned)((w.npad + 255) / 256), 256, 0, st>>>(conf, pred_cls, n, w.npad, classes, ncls, w.keys);
    DSN_CHECK_LAUNCH();
    const unsigned tiles = (unsigned)(w.npad / kSortTile);
    bitonic_tile_sort_kernel<<<tiles, kSortThreads, 0, st>>>(w.keys);
    DSN_CHECK_LAUNCH();
    for (long long k = 2ll * kSortTile; k <= w.npad; k <<= 1) {
        for (long long j = k >> 1; j >= kSortTile; j >>= 1) {
            bitonic_global_step_kernel<<<(unsigned)((w.npad / 2 + 255) / 256), 256, 0, st>>>(w.keys, w.npad / 2, j, k);
            DSN_CHECK_LAUNCH();
        }
        bitonic_tile_merge_kernel<<<tiles, kSortThreads, 0, st>>>(w.keys, k);
        DSN_CHECK_LAUNCH();
    }
    ap_scan_kernel<<<ncls, 256, 0, st>>>(w.keys, n, tp, conf, niou, w.tpc, w.conf_s, w.seg);
    DSN_CHECK_LAUNCH();
    ap_envelope_kernel<<<dim3(ncls, niou), 256, 0, st>>>(w.tpc, niou, w.seg, (const long long *)n_labels, w.env);
    DSN_CHECK_LAUNCH();
    ap_integrate_kernel<<<dim3(ncls, niou), 128, 0, st>>>(w.tpc, niou, w.seg, (const long long *)n_labels, w.env, x101, ap);
    DSN_CHECK_LAUNCH();
    if (npx > 0) {
        ap_pr_kernel<<<dim3(ncls, (npx + 255) / 256), 256, 0, st>>>(w.tpc, niou, w.conf_s, w.seg, (const long long *)n_labels, px, npx, p, r);
        DSN_CHECK_LAUNCH();
    }
    return DSN_OK;
}

extern "C" __attribute__((visibility("default"))) int dsn_compute_ap(const double *recall, const double *precision, int64_t n,
                                                                     const double *x101, double *env_ws, double *ap, void *stream) {
    if (n < 0) return DSN_E_INVALID;
    if (!x101 || !ap || (n > 0 && (!recall || !precision || !env_ws))) return DSN_E_INVALID;
    compute_ap_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(recall, precision, n, x101, env_ws, ap);
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}

extern "C" __attribute__((visibility("default"))) int dsn_det_confusion_matrix(const float *det, int64_t det_image_stride,
                                                                               int64_t det_row_stride, const int *det_count, int bs,
                                                                               int max_rows, const float *labels, const int *label_off,
                                                                               int max_labels_per_image, int nc, float conf,
                                                                               float iou_thres, int skip_empty, uint64_t *matrix,
                                                                               void *stream) {
    if (bs < 0 || max_rows < 0 || nc < 1 || max_labels_per_image < 0 || det_row_stride < 6) return DSN_E_INVALID;
    if (bs == 0) return DSN_OK;
    if (!label_off || !matrix || (max_rows > 0 && !det) || (max_labels_per_image > 0 && !labels)) return DSN_E_INVALID;
    if (max_rows > 8 * 256 || max_labels_per_image > 6000) return DSN_E_UNSUPPORTED;
    const size_t smem = sizeof(unsigned long long) * (size_t)(max_labels_per_image > 0 ? max_labels_per_image : 1);
    det_confusion_kernel<<<bs, 256, smem, (cudaStream_t)stream>>>(det, det_image_stride, det_row_stride, det_count, max_rows, labels,
                                                                  label_off, nc, conf, iou_thres, skip_empty, (unsigned long long *)matrix);
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}
