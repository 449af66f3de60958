// Per-image lazy NMS building blocks (selection, sorts, bitmask rounds) shared by the standalone NMS
// kernel (nms.cu) and the fused detection kernel (fused.cu).  See nms.cu for the algorithm.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>

#include "common.cuh"
#include "geom.cuh"

namespace dsn {

#ifndef DSN_NMS_THREADS
#define DSN_NMS_THREADS 512
#endif
constexpr int kNmsThreads = DSN_NMS_THREADS;
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
    const int *cut_dropped;           // optional (bs): the filter dropped candidates below a confidence cut
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
        const RowGeom rg = row_geom(g, (int)row);
        const int l = rg.l, plane = rg.plane, a = rg.a, cell = rg.cell;
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
    if (aa == aa && ab == ab) {
        // No coordinate is NaN (a NaN coordinate makes the area NaN), so std::max/min == fmaxf/fminf:
        // one FMNMX each.  Disjoint boxes (the common case) leave after 8 instructions.
        const float w = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), h = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
        if (!(w > 0.f && h > 0.f)) return false;       // inter == 0: ovr is 0 or NaN, never > thr_f >= 0
        const float inter = __fmul_rn(w, h);
        return __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa, ab), inter)) > thr_f;
    }
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
constexpr int kRankSortMax = 256;    // O(m^2) rank sort below, bitonic sort above
constexpr int kWords = kRound / 64;
constexpr int kChainClasses = 256;

struct NmsSmem {
    unsigned long long key[kChunk];
    float4 box[kChunk];               // un-offset xyxy (what is written out)
    float off[kChunk];                // cls * max_wh
    float4 kbox[kKeptSmem];           // survivors, offset boxes
    float karea[kKeptSmem];
    float4 rbox[kRound];              // offset boxes of the current round
    float rarea[kRound];
    float roff[kRound];               // class offsets of the current round / of the survivors
    float koff[kKeptSmem];
    float ext[4];                     // running extent of every box consumed so far: min x1, min y1, max x2, max y2
    float ext_w[kNmsThreads / 32][4]; // per-warp partials
    int ext_bad;                      // a consumed box had a non-finite coordinate
    union {
        unsigned int hist[kBins];                      // selection
        unsigned long long mask[kRound][kWords];       // resolution (never live at the same time)
    };
    unsigned int scan_tmp[kNmsThreads / 32];
    unsigned int dead[kRound / 32];
    unsigned int nz[kRound / 32];
    unsigned long long keptbits[kWords];
    // class chains (used while the class-aware culling holds and nc <= kChainClasses): survivors / round candidates of
    // one class are linked newest-to-oldest, so a candidate only meets the boxes that can overlap it
    int khead[kChainClasses];         // newest survivor slot of the class (-1: none); slots < kKeptSmem only
    short knext[kKeptSmem];           // next older survivor of the same class
    short rprev[kRound];              // previous (better-scored) candidate of the same class inside the round
    unsigned short rcls[kRound];      // class of the round's candidates
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
static __device__ int select_chunk(NmsSmem &s, const unsigned long long *keys, int n, unsigned long long lo,
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

static __device__ void bitonic_sort(unsigned long long *a, int n_pow2) {
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

// m <= kNmsThreads unique keys.  T = kNmsThreads / pow2ceil(m) (<= 32) adjacent lanes share one key:
// each counts the keys below it over a strided part of the list, a shuffle-xor tree adds the parts.
static __device__ void rank_sort(unsigned long long *a, int m) {
    int mp = 32;
    while (mp < m) mp <<= 1;
    int T = kNmsThreads / mp;
    T = T > 32 ? 32 : T;
    const int ki = threadIdx.x / T, sub = threadIdx.x - ki * T;
    unsigned long long mine = ~0ull;
    int rank = 0;
    if (ki < m) {
        mine = a[ki];
        for (int i = sub; i < m; i += T) rank += a[i] < mine ? 1 : 0;
    }
    for (int d = 1; d < T; d <<= 1) rank += __shfl_xor_sync(0xffffffffu, rank, d);
    __syncthreads();
    if (ki < m && sub == 0) a[rank] = mine;
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// NMS of one image by one CTA of kNmsThreads threads (see nms.cu for the algorithm).  `keys`/`n` =
// the image's candidate keys in global memory; results go to p.det / p.det_count.
// ---------------------------------------------------------------------------------------------
// `stashed`: the caller already placed the n <= kRankSortMax candidate keys, unsorted, in s.key — no gather from
// global memory is needed.
template <class BoxSrc>
__device__ void nms_image(NmsSmem &s, const NmsParams &p, const BoxSrc &src, int b, const unsigned long long *keys, int n,
                          bool stashed = false) {
    float *det = p.det + (size_t)b * p.max_det * 6;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // candidates wanted per selection: survivors rarely need more than ~2x max_det candidates
    const unsigned int target = (unsigned)min(kChunk, max(kRound, 2 * p.max_det));
    if (threadIdx.x == 0) {
        s.nkept = 0;
        s.ext[0] = s.ext[1] = INFINITY;
        s.ext[2] = s.ext[3] = -INFINITY;
        s.ext_bad = 0;
    }
    const bool chainable = p.nc <= kChainClasses;
    for (int i = threadIdx.x; i < kChainClasses; i += kNmsThreads) s.khead[i] = -1;
    int consumed = 0;               // candidates taken in sorted order so far (max_nms cut)
    int seen = 0;                   // keys gathered so far
    unsigned long long cur = 0ull;
    bool first = true;
    bool done = false;              // max_det survivors found
    __syncthreads();
    DSN_LAP_START();

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
        DSN_MARK(10);
        DSN_LAP(32);   // select
        // ---- 2. gather + sort ------------------------------------------------------------------
        int mp = 32;
        while (mp < m) mp <<= 1;
        if (!stashed) {                     // stashed implies first chunk, m == n <= kRankSortMax, keys in s.key
            if (threadIdx.x == 0) s.m = 0u;
            if (m > kRankSortMax)
                for (int i = threadIdx.x; i < mp; i += kNmsThreads) s.key[i] = ~0ull;
            __syncthreads();
            for_each_key(keys, n, [&](unsigned long long k) {
                if (k >= cur && k <= chunk_hi) s.key[atomicAdd(&s.m, 1u)] = k;
            });
            __syncthreads();
        }
        DSN_MARK(11);
        DSN_LAP(33);   // gather
        if (m <= kRankSortMax) rank_sort(s.key, m);
        else bitonic_sort(s.key, mp);
        DSN_MARK(12);
        DSN_LAP(34);   // sort
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
        DSN_MARK(13);
        DSN_LAP(35);   // boxes
        // ---- 3b. class-aware culling ------------------------------------------------------------------
        // Boxes of different classes are offset by >= max_wh (general.py:732-733).  If every box consumed so far
        // fits in an x- or y-range narrower than max_wh, two boxes of different classes cannot overlap after the
        // offset (fp32 rounding is monotone, so the order of the offset coordinates is preserved) and their IoU
        // test can be skipped: one compare instead of ~10 instructions for (nc-1)/nc of the pairs.
        {
            float e0 = INFINITY, e1 = INFINITY, e2 = -INFINITY, e3 = -INFINITY;
            bool bad = false;
            for (int i = threadIdx.x; i < m; i += kNmsThreads) {
                const float4 bx = s.box[i];
                bad = bad || !(fabsf(bx.x) < INFINITY && fabsf(bx.y) < INFINITY && fabsf(bx.z) < INFINITY && fabsf(bx.w) < INFINITY);
                e0 = fminf(e0, fminf(bx.x, bx.z)); e1 = fminf(e1, fminf(bx.y, bx.w));
                e2 = fmaxf(e2, fmaxf(bx.x, bx.z)); e3 = fmaxf(e3, fmaxf(bx.y, bx.w));
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                e0 = fminf(e0, __shfl_xor_sync(0xffffffffu, e0, d)); e1 = fminf(e1, __shfl_xor_sync(0xffffffffu, e1, d));
                e2 = fmaxf(e2, __shfl_xor_sync(0xffffffffu, e2, d)); e3 = fmaxf(e3, __shfl_xor_sync(0xffffffffu, e3, d));
            }
            if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(&s.ext_bad, 1);
            if (lane == 0) { s.ext_w[warp][0] = e0; s.ext_w[warp][1] = e1; s.ext_w[warp][2] = e2; s.ext_w[warp][3] = e3; }
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int w = 0; w < kNmsThreads / 32; ++w) {
                    s.ext[0] = fminf(s.ext[0], s.ext_w[w][0]); s.ext[1] = fminf(s.ext[1], s.ext_w[w][1]);
                    s.ext[2] = fmaxf(s.ext[2], s.ext_w[w][2]); s.ext[3] = fmaxf(s.ext[3], s.ext_w[w][3]);
                }
            }
            __syncthreads();
        }
        const double cull_lim = 0.999 * (double)p.off_mul;     // margin >> the fp32 rounding of cls * max_wh
        const bool cull = p.off_mul > 0.f && !s.ext_bad &&
                          (((double)s.ext[2] - (double)s.ext[0]) <= cull_lim || ((double)s.ext[3] - (double)s.ext[1]) <= cull_lim);
        DSN_LAP(36);   // extent
        // ---- 4. resolve the chunk in rounds of kRound candidates ---------------------------------
        for (int s0 = 0; s0 < m && !done; s0 += kRound) {
            const int cnt = min(kRound, m - s0);
            const int nkept0 = s.nkept;
            if (threadIdx.x < kRound) {
                float4 ob = make_float4(0.f, 0.f, 0.f, 0.f);
                if (threadIdx.x < cnt) ob = offset_box(s.box[s0 + threadIdx.x], s.off[s0 + threadIdx.x]);
                s.rbox[threadIdx.x] = ob;
                s.rarea[threadIdx.x] = box_area(ob);
                s.roff[threadIdx.x] = threadIdx.x < cnt ? s.off[s0 + threadIdx.x] : 0.f;
                s.rcls[threadIdx.x] = threadIdx.x < cnt ? (unsigned short)((unsigned int)s.key[s0 + threadIdx.x] % (unsigned)p.nc) : 0;
            }
            if (threadIdx.x < kRound / 32) { s.dead[threadIdx.x] = 0u; s.nz[threadIdx.x] = 0u; }
            static_assert((kRound * kWords) % kNmsThreads == 0, "mask words split evenly over the CTA");
#pragma unroll
            for (int i = 0; i < kRound * kWords / kNmsThreads; ++i) (&s.mask[0][0])[threadIdx.x + i * kNmsThreads] = 0ull;
            __syncthreads();
            DSN_LAP(37);   // round setup
            // class chains: previous candidate of the same class inside the round (built for every round so that the
            // survivor chains stay complete; used for the pair tests only while the class-aware culling holds)
            const bool use_chain = chainable && cull;
            if (chainable) {
                if (threadIdx.x < cnt) {
                    const unsigned short c = s.rcls[threadIdx.x];
                    int j = (int)threadIdx.x - 1;
                    while (j >= 0 && s.rcls[j] != c) --j;
                    s.rprev[threadIdx.x] = (short)j;
                }
                __syncthreads();
            }
            static_assert(kNmsThreads >= 2 * kRound, "the chain path runs (a) and (b) on disjoint halves of the CTA");
            if (use_chain) {
                if (threadIdx.x < kRound) {
                    // (a) against the survivors of the candidate's own class (newest first), then any survivors beyond
                    //     the shared-memory cache
                    const int j = threadIdx.x;
                    bool dead = false;
                    if (j < cnt && nkept0 > 0) {
                        const float4 bj = s.rbox[j];
                        const float aj = s.rarea[j], oj = s.roff[j];
                        // no early exit: the tests of successive hops are independent, only the hop itself (one LDS) is serial
#pragma unroll 4
                        for (int k = s.khead[s.rcls[j]]; k >= 0; k = s.knext[k]) dead = suppresses(s.kbox[k], s.karea[k], bj, aj, p.thr_f) || dead;
                        for (int k = kKeptSmem; k < nkept0 && !dead; ++k) {
                            const float *r = det + (size_t)k * 6;
                            const float ok = __fmul_rn(r[5], p.off_mul);
                            if (ok != oj) continue;
                            const float4 bk = offset_box(make_float4(r[0], r[1], r[2], r[3]), ok);
                            dead = suppresses(bk, box_area(bk), bj, aj, p.thr_f);
                        }
                    }
                    const unsigned int bal = __ballot_sync(0xffffffffu, dead);
                    if (lane == 0 && bal) atomicOr(&s.dead[warp], bal);
                } else if (threadIdx.x < 2 * kRound) {
                    // (b) inside the round: candidate i against the better-scored candidates of its class
                    const int i = threadIdx.x - kRound;
                    if (i < cnt) {
                        const float4 bi = s.rbox[i];
                        const float ai = s.rarea[i];
#pragma unroll 4
                        for (int j = s.rprev[i]; j >= 0; j = s.rprev[j]) {
                            if (suppresses(s.rbox[j], s.rarea[j], bi, ai, p.thr_f)) {
                                atomicOr(&s.mask[j][i >> 6], 1ull << (i & 63));
                                atomicOr(&s.nz[j >> 5], 1u << (j & 31));
                            }
                        }
                    }
                }
            } else {
            // (a) against the survivors so far: candidate = tid % kRound, the thread groups stride the kept list
            if (nkept0 > 0) {
                const int j = threadIdx.x & (kRound - 1), part = threadIdx.x / kRound;
                bool dead = false;
                if (j < cnt) {
                    const float4 bj = s.rbox[j];
                    const float aj = s.rarea[j], oj = s.roff[j];
                    for (int k = part; k < nkept0 && !dead; k += kNmsThreads / kRound) {
                        float4 bk;
                        float ak;
                        if (k < kKeptSmem) {
                            if (cull && s.koff[k] != oj) continue;          // other class: cannot overlap
                            bk = s.kbox[k];
                            ak = s.karea[k];
                        } else {
                            const float *r = det + (size_t)k * 6;
                            const float ok = __fmul_rn(r[5], p.off_mul);
                            if (cull && ok != oj) continue;
                            bk = offset_box(make_float4(r[0], r[1], r[2], r[3]), ok);
                            ak = box_area(bk);
                        }
                        dead = suppresses(bk, ak, bj, aj, p.thr_f);
                    }
                }
                const unsigned int bal = __ballot_sync(0xffffffffu, dead);
                if (lane == 0 && bal) atomicOr(&s.dead[warp & (kRound / 32 - 1)], bal);
            }
            DSN_LAP(38);   // (a) vs kept (no sync yet)
            // (b) pairwise inside the round.  Rows r and cnt-1-r together have cnt-1 later columns, so
            //     G threads per row pair get equal shares of the triangle; hits are rare -> atomics.
            {
                const int nrp = (cnt + 1) >> 1;
                int rpp = 16;
                while (rpp < nrp) rpp <<= 1;
                const int G = kNmsThreads / rpp;
                const int rp = threadIdx.x / G, sub = threadIdx.x - rp * G;
                if (rp < nrp) {
                    const int i1 = rp, i2 = cnt - 1 - rp;
                    const int len1 = cnt - 1 - i1, len2 = i2 != i1 ? cnt - 1 - i2 : 0;
                    const float4 b1 = s.rbox[i1], b2 = s.rbox[i2];
                    const float a1 = s.rarea[i1], a2 = s.rarea[i2], o1 = s.roff[i1], o2 = s.roff[i2];
                    for (int k = sub; k < len1 + len2; k += G) {
                        const bool fst = k < len1;
                        const int i = fst ? i1 : i2;
                        const int j = fst ? i1 + 1 + k : i2 + 1 + (k - len1);
                        if (cull && s.roff[j] != (fst ? o1 : o2)) continue;   // other class: cannot overlap
                        if (suppresses(fst ? b1 : b2, fst ? a1 : a2, s.rbox[j], s.rarea[j], p.thr_f)) {
                            atomicOr(&s.mask[i][j >> 6], 1ull << (j & 63));
                            atomicOr(&s.nz[i >> 5], 1u << (i & 31));
                        }
                    }
                }
            }
            }   // !use_chain
            __syncthreads();
            DSN_MARK(14);
            DSN_LAP(39);   // (b) pairs + sync
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
            DSN_MARK(15);
            DSN_LAP(40);   // (c) scan
            // (d) append survivors                                               general.py:745
            int my_slot = -1;
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
                        s.koff[slot] = s.roff[threadIdx.x];
                        if (chainable) {
                            // link behind the previous survivor of the class: the nearest kept one up the round's chain,
                            // else the class's newest survivor of earlier rounds
                            int j = s.rprev[threadIdx.x];
                            while (j >= 0 && !((s.keptbits[j >> 6] >> (j & 63)) & 1ull)) j = s.rprev[j];
                            int prev;
                            if (j >= 0) {
                                prev = nkept0 + __popcll(s.keptbits[j >> 6] & ((1ull << (j & 63)) - 1ull));
                                for (int v = 0; v < (j >> 6); ++v) prev += __popcll(s.keptbits[v]);
                            } else {
                                prev = s.khead[s.rcls[threadIdx.x]];
                            }
                            s.knext[slot] = (short)prev;
                            my_slot = slot;
                        }
                    }
                }
            }
            if (chainable) {
                __syncthreads();                                   // every knext is written before a head moves
                if (my_slot >= 0) atomicMax(&s.khead[s.rcls[threadIdx.x]], my_slot);
            }
            __syncthreads();
            DSN_LAP(41);   // (d) append
            done = s.nkept >= p.max_det;
        }
        if (done || consumed >= p.max_nms || seen >= n) break;
        if (chunk_hi >= p.key_hi) break;
        cur = chunk_hi + 1ull;
    }
    __syncthreads();
    DSN_MARK(16);
    // Confidence cut (dsn_conf_cut_levels): the emitted keys are exactly the best candidates of the image, so the
    // result is the reference's whenever the greedy pass stopped by itself (max_det survivors, or max_nms candidates
    // consumed).  Running out of emitted keys first means dropped candidates could still have mattered: -2.
    const bool starved = !done && consumed < p.max_nms && p.cut_dropped != nullptr && p.cut_dropped[b] != 0;
    if (threadIdx.x == 0) p.det_count[b] = starved ? -2 : s.nkept;
}

static inline int make_nms_params(NmsParams &p, const uint64_t *keys, int64_t cap, const int *counts, int bs, int nc,
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
    p.cut_dropped = nullptr;
    p.det = det;
    p.det_count = det_count;
    return DSN_OK;
}


}  // namespace dsn
