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
    if (!(w > 0.f && h > 0.f)) return false;   // inter == 0 (or NaN): ovr is 0 or NaN, never > thr_f >= 0
    const float inter = __fmul_rn(w, h);
    const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa, ab), inter));
    return ovr > thr_f;      // NaN (0/0) never suppresses
}
__device__ __forceinline__ float4 offset_box(const float4 &b, float off) {
    return make_float4(__fadd_rn(b.x, off), __fadd_rn(b.y, off), __fadd_rn(b.z, off), __fadd_rn(b.w, off));   // general.py:733
}

constexpr int kRound = 256;          // candidates resolved per bitmask round
constexpr int kWords = kRound / 64;

struct NmsSmem {
    unsigned long long key[kChunk];
    float4 box[kChunk];               // un-offset xyxy (what is written out)
    float off[kChunk];                // cls * max_wh
    float4 kbox[kKeptSmem];           // survivors, offset boxes
    float karea[kKeptSmem];
    float4 rbox[kRound];              // offset boxes of the current round
    float rarea[kRound];
    union {
        unsigned int hist[kBins];                      // selection
        unsigned long long mask[kRound][kWords];       // resolution (never live at the same time)
    };
    unsigned int scan_tmp[kNmsThreads / 32];
    unsigned int dead[kRound / 32];
    unsigned int nz[kRound / 32];
    unsigned long long keptbits[kWords];
    unsigned int m;                   // gathered count
    unsigned int below;
    unsigned int red_sum, red_max, red_min;
    int nkept;
};

// Visit every key of the image once: 16 B loads, 4 independent loads in flight per thread (the
// selection passes are latency-bound on L2, not bandwidth-bound).
template <class F>
__device__ __forceinline__ void for_each_key(const unsigned long long *keys, int n, F f) {
    int head = 0;
    if (((uintptr_t)keys & 15) != 0) {          // align to 16 B
        head = 1;
        if (threadIdx.x == 0 && n > 0) f(keys[0]);
    }
    const ulonglong2 *k2 = reinterpret_cast<const ulonglong2 *>(keys + head);
    const int n2 = (n - head) >> 1;
    for (int i = threadIdx.x; i < n2; i += kNmsThreads * 4) {
        ulonglong2 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int idx = i + j * kNmsThreads;
            v[j] = idx < n2 ? k2[idx] : make_ulonglong2(~0ull, ~0ull);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (i + j * kNmsThreads < n2) { f(v[j].x); f(v[j].y); }
        }
    }
    if (((n - head) & 1) && threadIdx.x == 0) f(keys[n - 1]);
}

// Select the next chunk: the smallest keys in [lo, hi] (hi inclusive), about `target` of them and
// never more than kChunk.  Returns the chunk's inclusive upper key in *chunk_hi and its size; size
// 0 = nothing left in [lo, hi].  `below` (first call only) counts keys < lo.
__device__ int select_chunk(NmsSmem &s, const unsigned long long *keys, int n, unsigned long long lo,
                            unsigned long long hi, unsigned int target, unsigned long long *chunk_hi,
                            unsigned int *below) {
    for (;;) {
        const unsigned long long width = hi - lo;
        const int bits = 64 - __clzll((long long)(width | 1ull));
        const int sh = bits > kBinBits ? bits - kBinBits : 0;
        for (int i = threadIdx.x; i < kBins; i += kNmsThreads) s.hist[i] = 0u;
        if (threadIdx.x == 0) { s.below = 0u; s.red_sum = 0u; s.red_max = 0u; s.red_min = 0xFFFFFFFFu; }
        __syncthreads();
        unsigned int nb = 0;
        for_each_key(keys, n, [&](unsigned long long k) {
            if (k < lo) ++nb;
            else if (k <= hi) atomicAdd(&s.hist[(unsigned int)((k - lo) >> sh)], 1u);
        });
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
        if (total == 0) return 0;          // uniform: every thread sees the same total
        // boundaries j = 1..kBins (after bin j-1); prefix(j) = keys in bins < j.  nfit = number of
        // boundaries with prefix <= target (a prefix of 1..kBins since prefix is monotone),
        // cfit = the largest such prefix, over = the smallest prefix > target.
        unsigned int pre = wbase + incl - mine, nfit = 0, cfit = 0, over = 0xFFFFFFFFu;
#pragma unroll
        for (int j = 0; j < per; ++j) {
            pre += h[j];
            if (pre <= target) { ++nfit; cfit = pre; }
            else over = min(over, pre);
        }
        nfit = __reduce_add_sync(0xffffffffu, nfit);
        cfit = __reduce_max_sync(0xffffffffu, cfit);
        over = __reduce_min_sync(0xffffffffu, over);
        if (lane == 0) {
            if (nfit) atomicAdd(&s.red_sum, nfit);
            if (cfit) atomicMax(&s.red_max, cfit);
            atomicMin(&s.red_min, over);
        }
        __syncthreads();
        unsigned int jmax = s.red_sum, cnt = s.red_max;
        const unsigned int cnt_next = s.red_min;        // prefix(jmax + 1)
        __syncthreads();                                // red_* / hist are rewritten by the next pass
        if (cnt == 0 && cnt_next <= (unsigned)kChunk) { // first non-empty bin alone exceeds target but fits
            cnt = cnt_next;
            ++jmax;
        }
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

// m <= kNmsThreads unique keys: every thread counts the keys below its own (broadcast reads).
__device__ void rank_sort(unsigned long long *a, int m) {
    unsigned long long mine = 0ull;
    int rank = 0;
    if (threadIdx.x < m) {
        mine = a[threadIdx.x];
        for (int i = 0; i < m; ++i) rank += a[i] < mine ? 1 : 0;
    }
    __syncthreads();
    if (threadIdx.x < m) a[rank] = mine;
    __syncthreads();
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
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // candidates wanted per selection: survivors rarely need more than ~2x max_det candidates
    const unsigned int target = (unsigned)min(kChunk, max(kRound, 2 * p.max_det));
    if (threadIdx.x == 0) s.nkept = 0;
    int consumed = 0;               // candidates taken in sorted order so far (max_nms cut)
    int seen = 0;                   // keys gathered so far
    unsigned long long cur = 0ull;
    bool first = true;
    __syncthreads();

    for (;;) {
        // ---- 1. next chunk of keys -------------------------------------------------------------
        unsigned long long chunk_hi = ~0ull;
        int m;
        if (first && n <= kChunk) {
            m = n;
        } else if (first) {
            unsigned int below = 0;
            m = select_chunk(s, keys, n, p.key_lo_hint, p.key_hi, target, &chunk_hi, &below);
            if (below) m = select_chunk(s, keys, n, 0ull, p.key_hi, target, &chunk_hi, nullptr);   // conf > 1.0 exists
            else cur = p.key_lo_hint;
        } else {
            m = select_chunk(s, keys, n, cur, p.key_hi, target, &chunk_hi, nullptr);
        }
        first = false;
        if (m == 0) break;
        // ---- 2. gather + sort ------------------------------------------------------------------
        int mp = 32;
        while (mp < m) mp <<= 1;
        if (threadIdx.x == 0) s.m = 0u;
        if (m > kNmsThreads)
            for (int i = threadIdx.x; i < mp; i += kNmsThreads) s.key[i] = ~0ull;
        __syncthreads();
        for_each_key(keys, n, [&](unsigned long long k) {
            if (k >= cur && k <= chunk_hi) s.key[atomicAdd(&s.m, 1u)] = k;
        });
        __syncthreads();
        if (m <= kNmsThreads) rank_sort(s.key, m);
        else bitonic_sort(s.key, mp);
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
        // ---- 4. resolve the chunk in rounds of kRound candidates ---------------------------------
        bool done = false;
        for (int s0 = 0; s0 < m && !done; s0 += kRound) {
            const int cnt = min(kRound, m - s0);
            const int nkept0 = s.nkept;
            if (threadIdx.x < kRound) {
                float4 ob = make_float4(0.f, 0.f, 0.f, 0.f);
                if (threadIdx.x < cnt) ob = offset_box(s.box[s0 + threadIdx.x], s.off[s0 + threadIdx.x]);
                s.rbox[threadIdx.x] = ob;
                s.rarea[threadIdx.x] = box_area(ob);
            }
            if (threadIdx.x < kRound / 32) { s.dead[threadIdx.x] = 0u; s.nz[threadIdx.x] = 0u; }
            __syncthreads();
            // (a) against the survivors so far: candidate = tid % 256, two threads stride the kept list
            if (nkept0 > 0) {
                const int j = threadIdx.x & (kRound - 1), part = threadIdx.x / kRound;
                bool dead = false;
                if (j < cnt) {
                    const float4 bj = s.rbox[j];
                    const float aj = s.rarea[j];
                    for (int k = part; k < nkept0 && !dead; k += kNmsThreads / kRound) {
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
                if (lane == 0 && bal) atomicOr(&s.dead[warp & (kRound / 32 - 1)], bal);
            }
            // (b) pairwise inside the round: thread = (row i, half of the columns); each mask word
            //     has exactly one writer
            {
                const int i = threadIdx.x >> 1, half = threadIdx.x & 1;
                unsigned long long w0 = 0ull, w1 = 0ull;
                if (i < cnt) {
                    const float4 bi = s.rbox[i];
                    const float ai = s.rarea[i];
                    const int jb = half * 128;
                    const int jlo = max(i + 1, jb), jhi = min(cnt, jb + 128);
                    for (int j = jlo; j < jhi; ++j) {
                        if (suppresses(bi, ai, s.rbox[j], s.rarea[j], p.thr_f)) {
                            if (j - jb < 64) w0 |= 1ull << (j - jb);
                            else w1 |= 1ull << (j - jb - 64);
                        }
                    }
                }
                if (i < kRound) {
                    s.mask[i][half * 2] = w0;
                    s.mask[i][half * 2 + 1] = w1;
                }
                // rows with any suppression bit: lanes 2r, 2r+1 hold the two halves of row (16*warp + r)
                const unsigned int any = __ballot_sync(0xffffffffu, (w0 | w1) != 0ull);
                if (lane == 0 && any) {
                    unsigned int rows16 = 0;
#pragma unroll
                    for (int r = 0; r < 16; ++r) rows16 |= ((any >> (2 * r)) & 3u) ? (1u << r) : 0u;
                    atomicOr(&s.nz[warp >> 1], rows16 << (16 * (warp & 1)));
                }
            }
            __syncthreads();
            // (c) serial greedy scan, visiting only alive rows that suppress something
            if (threadIdx.x == 0) {
                unsigned long long alive[kWords], nzw[kWords];
#pragma unroll
                for (int w = 0; w < kWords; ++w) {
                    const int c = cnt - 64 * w;
                    const unsigned long long valid = c >= 64 ? ~0ull : (c > 0 ? ((1ull << c) - 1ull) : 0ull);
                    alive[w] = valid & ~((unsigned long long)s.dead[2 * w] | ((unsigned long long)s.dead[2 * w + 1] << 32));
                    nzw[w] = (unsigned long long)s.nz[2 * w] | ((unsigned long long)s.nz[2 * w + 1] << 32);
                }
#pragma unroll
                for (int w = 0; w < kWords; ++w) {
                    unsigned long long pend = alive[w] & nzw[w];
                    while (pend) {
                        const int i = __ffsll((long long)pend) - 1;
                        const unsigned long long *mrow = s.mask[64 * w + i];
#pragma unroll
                        for (int v = 0; v < kWords; ++v)
                            if (v >= w) alive[v] &= ~mrow[v];
                        const unsigned long long after = i == 63 ? 0ull : (~0ull << (i + 1));
                        pend = alive[w] & nzw[w] & after;
                    }
                }
                // keep the first (max_det - nkept0) alive rows                      general.py:735-736
                int room = p.max_det - nkept0;
#pragma unroll
                for (int w = 0; w < kWords; ++w) {
                    unsigned long long a = alive[w];
                    const int c = __popcll(a);
                    if (c > room) {
                        // drop the highest bits until `room` remain
                        unsigned long long keep = 0ull;
                        for (int q = 0; q < room; ++q) { const unsigned long long low = a & (~a + 1ull); keep |= low; a ^= low; }
                        a = keep;
                    }
                    room -= __popcll(a);
                    s.keptbits[w] = a;
                }
                s.nkept = p.max_det - room;
            }
            __syncthreads();
            // (d) append survivors                                               general.py:745
            if (threadIdx.x < cnt) {
                const int w = threadIdx.x >> 6, bit = threadIdx.x & 63;
                const unsigned long long kb = s.keptbits[w];
                if ((kb >> bit) & 1ull) {
                    int slot = nkept0 + __popcll(kb & ((1ull << bit) - 1ull));
                    for (int v = 0; v < w; ++v) slot += __popcll(s.keptbits[v]);
                    const int ci = s0 + threadIdx.x;
                    const float4 bx = s.box[ci];
                    const unsigned long long key = s.key[ci];
                    float *r = det + (size_t)slot * 6;
                    r[0] = bx.x; r[1] = bx.y; r[2] = bx.z; r[3] = bx.w;
                    r[4] = key_conf(key);
                    r[5] = (float)((unsigned int)key % (unsigned)p.nc);
                    if (slot < kKeptSmem) {
                        s.kbox[slot] = s.rbox[threadIdx.x];
                        s.karea[slot] = s.rarea[threadIdx.x];
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
