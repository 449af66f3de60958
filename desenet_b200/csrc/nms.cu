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
#include <algorithm>
#include <cmath>
#include <cstring>

#include "common.cuh"
#include "geom.cuh"

namespace dsn {

constexpr int kNmsThreads = 512;
constexpr int kChunk = 2048;        // candidates sorted / resolved per round
constexpr int kBins = 2048;         // histogram bins per refinement level (11 bits)
constexpr int kBinBits = 11;
constexpr int kKeptSmem = 1024;     // survivors cached in shared memory (rest re-read from `det`)
constexpr int kSub = 64;            // sub-block resolved by one bitmask scan

struct NmsParams {
    const unsigned long long *keys;
    long long cap;
    const int *counts;
    int nc;
    float thr_f;        // largest fp32 <= iou_thres: (double)ovr > iou_thres  <=>  ovr > thr_f
    float off_mul;      // agnostic ? 0 : max_wh                                 general.py:732
    int max_nms, max_det;
    unsigned long long key_lo_hint;   // keys of conf <= 1.0 are >= this (hint, not required)
    unsigned long long key_hi;        // every key is <= this (conf > conf_thres)
    float *det;
    int *det_count;
};

struct BoxFromZ {
    const float *z;
    long long A;
    int no;
    __device__ __forceinline__ float4 operator()(int b, unsigned int row) const {
        const float *r = z + ((size_t)b * A + row) * no;
        return xywh2xyxy(__ldg(r), __ldg(r + 1), __ldg(r + 2), __ldg(r + 3));       // general.py:706
    }
};

struct BoxFromLevels {
    HeadGeom g;
    __device__ __forceinline__ float4 operator()(int b, unsigned int row) const {
        int l = 0;
        while (l + 1 < g.nl && (int)row >= g.row_off[l + 1]) ++l;
        const int plane = g.plane[l];
        const int r = (int)row - g.row_off[l];
        const int a = r / plane, cell = r - a * plane;
        const float *base = g.raw[l] + ((size_t)(b * g.na + a) * g.no) * plane + cell;
        const float x = decode_xy(__ldg(base), cell % g.nx[l], g.stride[l]);
        const float y = decode_xy(__ldg(base + plane), cell / g.nx[l], g.stride[l]);
        const float w = decode_wh(__ldg(base + 2 * (size_t)plane), g.anchor_w[l][a]);
        const float h = decode_wh(__ldg(base + 3 * (size_t)plane), g.anchor_h[l][a]);
        return xywh2xyxy(x, y, w, h);
    }
};

// torchvision CPU kernel arithmetic: std::max/min, fp32 products and sums rounded one by one.
__device__ __forceinline__ float tv_max(float a, float b) { return a < b ? b : a; }
__device__ __forceinline__ float tv_min(float a, float b) { return b < a ? b : a; }
__device__ __forceinline__ float box_area(const float4 &b) { return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y)); }
__device__ __forceinline__ bool suppresses(const float4 &a, float aa, const float4 &b, float ab, float thr_f) {
    const float xx1 = tv_max(a.x, b.x), yy1 = tv_max(a.y, b.y);
    const float xx2 = tv_min(a.z, b.z), yy2 = tv_min(a.w, b.w);
    const float w = tv_max(0.f, __fsub_rn(xx2, xx1)), h = tv_max(0.f, __fsub_rn(yy2, yy1));
    const float inter = __fmul_rn(w, h);
    const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa, ab), inter));
    return ovr > thr_f;      // NaN (0/0) never suppresses
}
__device__ __forceinline__ float4 offset_box(const float4 &b, float off) {
    return make_float4(__fadd_rn(b.x, off), __fadd_rn(b.y, off), __fadd_rn(b.z, off), __fadd_rn(b.w, off));   // general.py:733
}

struct NmsSmem {
    unsigned long long key[kChunk];
    float4 box[kChunk];               // un-offset xyxy (what is written out)
    float off[kChunk];                // cls * max_wh
    float4 kbox[kKeptSmem];           // survivors, offset boxes
    float karea[kKeptSmem];
    unsigned int hist[kBins];
    unsigned long long mask[kSub];
    unsigned int scan_tmp[kNmsThreads / 32];
    unsigned int dead[2];
    unsigned long long keptbits;
    unsigned int m;                   // gathered count
    unsigned int below;
    unsigned int bc_a, bc_b;          // block-reduce broadcasts
    int nkept;
};

__device__ __forceinline__ unsigned int block_sum(unsigned int v, unsigned int *tmp, unsigned int *bc) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = __reduce_add_sync(0xffffffffu, v);
    if (lane == 0) tmp[w] = v;
    __syncthreads();
    if (w == 0) {
        unsigned int t = lane < kNmsThreads / 32 ? tmp[lane] : 0u;
        t = __reduce_add_sync(0xffffffffu, t);
        if (lane == 0) *bc = t;
    }
    __syncthreads();
    return *bc;
}

// Select the next chunk: the keys in [lo, hi] (hi inclusive) that are the smallest <= kChunk of
// the remaining ones.  Returns the chunk's inclusive upper key in *chunk_hi and its size; size 0 =
// nothing left in [lo, hi].  `below` (first call only) counts keys < lo.
__device__ int select_chunk(NmsSmem &s, const unsigned long long *keys, int n, unsigned long long lo,
                            unsigned long long hi, unsigned long long *chunk_hi, unsigned int *below) {
    for (;;) {
        const unsigned long long width = hi - lo;
        const int bits = 64 - __clzll((long long)(width | 1ull));
        const int sh = bits > kBinBits ? bits - kBinBits : 0;
        for (int i = threadIdx.x; i < kBins; i += kNmsThreads) s.hist[i] = 0u;
        if (threadIdx.x == 0) s.below = 0u;
        __syncthreads();
        unsigned int nb = 0;
        for (int i = threadIdx.x; i < n; i += kNmsThreads) {
            const unsigned long long k = keys[i];
            if (k < lo) ++nb;
            else if (k <= hi) atomicAdd(&s.hist[(unsigned int)((k - lo) >> sh)], 1u);
        }
        if (nb) atomicAdd(&s.below, nb);
        __syncthreads();
        if (below) *below = s.below;
        // each thread owns kBins/kNmsThreads consecutive bins; exclusive block scan of the sums
        constexpr int per = kBins / kNmsThreads;
        unsigned int h[per], mine = 0;
#pragma unroll
        for (int j = 0; j < per; ++j) { h[j] = s.hist[threadIdx.x * per + j]; mine += h[j]; }
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        unsigned int incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned int up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        if (lane == 31) s.scan_tmp[w] = incl;
        __syncthreads();
        unsigned int wbase = 0, total = 0;
#pragma unroll
        for (int j = 0; j < kNmsThreads / 32; ++j) {
            const unsigned int t = s.scan_tmp[j];
            if (j < w) wbase += t;
            total += t;
        }
        __syncthreads();   // scan_tmp is reused by block_sum below
        if (total == 0) return 0;
        // number of bin boundaries j in 1..kBins whose exclusive prefix is <= kChunk, and that prefix
        unsigned int pre = wbase + incl - mine, nfit = 0, cnt_fit = 0;
#pragma unroll
        for (int j = 0; j < per; ++j) {
            pre += h[j];                       // exclusive prefix of boundary (bin index + 1)
            if (pre <= (unsigned)kChunk) { ++nfit; cnt_fit = pre; }
        }
        const unsigned int jmax = block_sum(nfit, s.scan_tmp, &s.bc_a);
        // cnt at jmax = max over threads of cnt_fit (prefix is monotone)
        unsigned int cmax = cnt_fit;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, d));
        if (threadIdx.x == 0) s.bc_b = 0u;
        __syncthreads();
        if (lane == 0 && cmax) atomicMax(&s.bc_b, cmax);
        __syncthreads();
        const unsigned int cnt = s.bc_b;
        __syncthreads();
        if (cnt > 0) {
            const unsigned long long span = jmax >= (unsigned)kBins ? ~0ull : (((unsigned long long)jmax << sh) - 1ull);
            *chunk_hi = span > width ? hi : lo + span;     // saturating: empty bins may lie beyond hi
            return (int)cnt;
        }
        // bins [0, jmax) are empty and bin jmax alone exceeds kChunk: refine inside it
        lo = lo + ((unsigned long long)jmax << sh);        // <= hi: the bin holds keys <= hi
        const unsigned long long bw = (1ull << sh) - 1ull;
        hi = bw > hi - lo ? hi : lo + bw;
        below = nullptr;
    }
}

__device__ void bitonic_sort(unsigned long long *a, int n_pow2) {
    for (int k = 2; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (n_pow2 >> 1); t += kNmsThreads) {
                const int ix = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int pj = ix | j;
                const unsigned long long x = a[ix], y = a[pj];
                const bool up = (ix & k) == 0;
                if ((x > y) == up) { a[ix] = y; a[pj] = x; }
            }
            __syncthreads();
        }
    }
}

template <class BoxSrc>
__global__ void __launch_bounds__(kNmsThreads) nms_kernel(const NmsParams p, const BoxSrc src) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NmsSmem &s = *reinterpret_cast<NmsSmem *>(smem_raw);
    const int b = blockIdx.x;
    const int total = p.counts[b];
    if (total > p.cap) {                     // keys were dropped: the caller must retry with a larger cap
        if (threadIdx.x == 0) p.det_count[b] = -1;
        return;
    }
    if (total <= 0 || p.max_det <= 0 || p.max_nms <= 0) {
        if (threadIdx.x == 0) p.det_count[b] = 0;
        return;
    }
    const int n = total;
    const unsigned long long *keys = p.keys + (size_t)b * p.cap;
    float *det = p.det + (size_t)b * p.max_det * 6;
    if (threadIdx.x == 0) s.nkept = 0;
    int consumed = 0;               // candidates taken in sorted order so far (max_nms cut)
    int seen = 0;                   // keys consumed including those beyond nothing (== consumed)
    unsigned long long cur = 0ull;
    bool first = true;
    __syncthreads();

    for (;;) {
        // ---- 1. next chunk of keys -------------------------------------------------------------
        unsigned long long chunk_hi = ~0ull;
        int m;
        if (first && n <= kChunk) {
            m = n;
        } else {
            if (first) {
                unsigned int below = 0;
                m = select_chunk(s, keys, n, p.key_lo_hint, p.key_hi, &chunk_hi, &below);
                if (below) {                 // confidences above 1.0 exist: start from key 0 instead
                    m = select_chunk(s, keys, n, 0ull, p.key_hi, &chunk_hi, nullptr);
                } else {
                    cur = p.key_lo_hint;
                }
            } else {
                m = select_chunk(s, keys, n, cur, p.key_hi, &chunk_hi, nullptr);
            }
        }
        first = false;
        if (m == 0) break;
        // ---- 2. gather + sort ------------------------------------------------------------------
        int mp = 32;
        while (mp < m) mp <<= 1;
        if (threadIdx.x == 0) s.m = 0u;
        for (int i = threadIdx.x; i < mp; i += kNmsThreads) s.key[i] = ~0ull;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += kNmsThreads) {
            const unsigned long long k = keys[i];
            if (k >= cur && k <= chunk_hi) s.key[atomicAdd(&s.m, 1u)] = k;
        }
        __syncthreads();
        bitonic_sort(s.key, mp);
        seen += m;
        if (m > p.max_nms - consumed) m = p.max_nms - consumed;          // general.py:728-729
        consumed += m;
        // ---- 3. boxes --------------------------------------------------------------------------
        for (int i = threadIdx.x; i < m; i += kNmsThreads) {
            const unsigned int ref = (unsigned int)s.key[i];
            const unsigned int row = ref / (unsigned)p.nc;
            s.box[i] = src(b, row);
            s.off[i] = __fmul_rn((float)(ref - row * (unsigned)p.nc), p.off_mul);      // general.py:732
        }
        __syncthreads();
        // ---- 4. resolve the chunk in 64-wide sub-blocks ------------------------------------------
        bool done = false;
        for (int s0 = 0; s0 < m && !done; s0 += kSub) {
            const int cnt = min(kSub, m - s0);
            const int nkept0 = s.nkept;
            if (threadIdx.x < kSub) s.mask[threadIdx.x] = 0ull;
            if (threadIdx.x < 2) s.dead[threadIdx.x] = 0u;
            __syncthreads();
            // (a) against the survivors so far: candidate = tid % 64, 8 threads stride the kept list
            {
                const int j = threadIdx.x & (kSub - 1), part = threadIdx.x >> 6;
                bool dead = false;
                if (j < cnt) {
                    const float4 bj = offset_box(s.box[s0 + j], s.off[s0 + j]);
                    const float aj = box_area(bj);
                    for (int k = part; k < nkept0 && !dead; k += kNmsThreads / kSub) {
                        float4 bk;
                        float ak;
                        if (k < kKeptSmem) {
                            bk = s.kbox[k];
                            ak = s.karea[k];
                        } else {
                            const float *r = det + (size_t)k * 6;
                            bk = offset_box(make_float4(r[0], r[1], r[2], r[3]), __fmul_rn(r[5], p.off_mul));
                            ak = box_area(bk);
                        }
                        dead = suppresses(bk, ak, bj, aj, p.thr_f);
                    }
                }
                const unsigned int bal = __ballot_sync(0xffffffffu, dead);
                if ((threadIdx.x & 31) == 0 && bal) atomicOr(&s.dead[(threadIdx.x >> 5) & 1], bal);
            }
            // (b) pairwise inside the sub-block: thread = (i, group of 8 j's)
            {
                const int i = threadIdx.x >> 3, jg = (threadIdx.x & 7) * 8;
                if (i < cnt && jg + 7 > i) {
                    const float4 bi = offset_box(s.box[s0 + i], s.off[s0 + i]);
                    const float ai = box_area(bi);
                    unsigned long long bits = 0ull;
#pragma unroll
                    for (int d = 0; d < 8; ++d) {
                        const int j = jg + d;
                        if (j > i && j < cnt) {
                            const float4 bj = offset_box(s.box[s0 + j], s.off[s0 + j]);
                            if (suppresses(bi, ai, bj, box_area(bj), p.thr_f)) bits |= 1ull << j;
                        }
                    }
                    if (bits) atomicOr(&s.mask[i], bits);
                }
            }
            __syncthreads();
            // (c) serial greedy scan over the survivors of (a)
            if (threadIdx.x == 0) {
                const unsigned long long valid = cnt == 64 ? ~0ull : ((1ull << cnt) - 1ull);
                unsigned long long alive = valid & ~((unsigned long long)s.dead[0] | ((unsigned long long)s.dead[1] << 32));
                unsigned long long kept = 0ull;
                int nk = nkept0;
                while (alive && nk < p.max_det) {
                    const int i = __ffsll((long long)alive) - 1;
                    kept |= 1ull << i;
                    ++nk;
                    alive &= ~(s.mask[i] | (1ull << i));
                }
                s.keptbits = kept;
                s.nkept = nk;
            }
            __syncthreads();
            // (d) append survivors                                               general.py:745
            if (threadIdx.x < cnt) {
                const unsigned long long kept = s.keptbits;
                if ((kept >> threadIdx.x) & 1ull) {
                    const int slot = nkept0 + __popcll(kept & ((1ull << threadIdx.x) - 1ull));
                    const int ci = s0 + threadIdx.x;
                    const float4 bx = s.box[ci];
                    const unsigned long long key = s.key[ci];
                    float *r = det + (size_t)slot * 6;
                    r[0] = bx.x; r[1] = bx.y; r[2] = bx.z; r[3] = bx.w;
                    r[4] = key_conf(key);
                    r[5] = (float)((unsigned int)key % (unsigned)p.nc);
                    if (slot < kKeptSmem) {
                        const float4 ob = offset_box(bx, s.off[ci]);
                        s.kbox[slot] = ob;
                        s.karea[slot] = box_area(ob);
                    }
                }
            }
            __syncthreads();
            done = s.nkept >= p.max_det;
        }
        if (done || consumed >= p.max_nms || seen >= n) break;
        if (chunk_hi >= p.key_hi) break;
        cur = chunk_hi + 1ull;
    }
    __syncthreads();
    if (threadIdx.x == 0) p.det_count[b] = s.nkept;
}

static int make_nms_params(NmsParams &p, const uint64_t *keys, int64_t cap, const int *counts, int bs, int nc,
                           float conf_thres, double iou_thres, int agnostic, int max_nms, int max_det, float max_wh, float *det,
                           int *det_count) {
    if (!keys || !counts || !det_count || cap < 1 || bs < 0 || nc < 1) return DSN_E_INVALID;
    if (max_det > 0 && !det) return DSN_E_INVALID;
    if (!(iou_thres >= 0.0 && iou_thres <= 1.0)) return DSN_E_INVALID;          // general.py:672
    p.keys = (const unsigned long long *)keys;
    p.cap = cap;
    p.counts = counts;
    p.nc = nc;
    float t = (float)iou_thres;                       // round to nearest, then step down if above
    if ((double)t > iou_thres) t = std::nextafterf(t, -INFINITY);
    p.thr_f = t;
    p.off_mul = agnostic ? 0.f : max_wh;
    p.max_nms = max_nms;
    p.max_det = max_det;
    p.key_lo_hint = (unsigned long long)(~0x3F800000u) << 32;    // conf == 1.0f
    uint32_t tb;
    const float ct = conf_thres > 0.f ? conf_thres : 0.f;
    memcpy(&tb, &ct, 4);
    p.key_hi = ((unsigned long long)(~tb) << 32) | 0xFFFFFFFFull;   // conf > conf_thres  =>  key <= this
    p.det = det;
    p.det_count = det_count;
    return DSN_OK;
}

template <class BoxSrc>
static int launch_nms(const NmsParams &p, const BoxSrc &src, int bs, cudaStream_t st) {
    if (bs == 0) return DSN_OK;
    const size_t smem = sizeof(NmsSmem);
    static bool configured = false;   // per template instantiation
    if (!configured) {
        DSN_CUDA(cudaFuncSetAttribute(nms_kernel<BoxSrc>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    nms_kernel<BoxSrc><<<bs, kNmsThreads, smem, st>>>(p, src);
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}

}  // namespace dsn

using namespace dsn;

extern "C" __attribute__((visibility("default"))) int dsn_nms_levels(const float *const *raw, const int *ny, const int *nx, const float *stride,
                              const float *anchor_grid, int nl, int bs, int na, int no, const uint64_t *keys,
                              int64_t cap, const int *counts, float conf_thres, double iou_thres, int agnostic,
                              int max_nms, int max_det, float max_wh, float *det, int *det_count, void *stream) {
    if (!stride || !anchor_grid) return DSN_E_INVALID;
    BoxFromLevels src;
    int rc = make_geom(src.g, raw, ny, nx, stride, anchor_grid, nl, bs, na, no);
    if (rc) return rc;
    NmsParams p;
    rc = make_nms_params(p, keys, cap, counts, bs, no - 5, conf_thres, iou_thres, agnostic, max_nms, max_det, max_wh, det, det_count);
    if (rc) return rc;
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
