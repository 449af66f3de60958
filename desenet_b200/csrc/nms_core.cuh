// Per-image lazy NMS building blocks (selection, sorts, bitmask rounds) shared by the standalone NMS
// kernel (nms.cu) and the fused detection kernel (fused.cu).  See nms.cu for the algorithm.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <type_traits>

#include "common.cuh"
#include "geom.cuh"

#ifndef DSN_NMS_THREADS
#define DSN_NMS_THREADS 512
#endif
// Everything below depends on the CTA size, and the library holds TWO instantiations (nms.cu: 512 threads, nms_small.cu:
// 256 threads): the inline namespace keeps their symbols apart.
#ifndef DSN_NMS_NS
#define DSN_NMS_NS nms_t512
#endif

namespace dsn {
inline namespace DSN_NMS_NS {

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
    const unsigned int *cut_hi;       // optional (bs): that cut (confidence half of the largest emitted key) - bounds the selection range
    float *det;
    int *det_count;
};

// Box sources: fetch() issues the loads of one row (raw values + what finish() needs), finish() turns them into the
// xyxy box.  The NMS fetches several rows per thread before finishing any, so their DRAM round trips overlap.
struct RawBox {
    float x, y, w, h;
    int l, a, cell;        // l < 0: x, y, w, h are already pixels (decoded prediction or an a-priori label row)
};

template <typename T>
struct BoxFromZ {
    const T *z;
    long long A;
    int no;
    __device__ __forceinline__ RawBox fetch(int b, unsigned int row) const {
        const T *r = z + ((size_t)b * A + row) * no;
        return RawBox{Elem<T>::ld1c(r), Elem<T>::ld1c(r + 1), Elem<T>::ld1c(r + 2), Elem<T>::ld1c(r + 3), -1, 0, 0};
    }
    __device__ __forceinline__ float4 finish(const RawBox &q) const { return xywh2xyxy(q.x, q.y, q.w, q.h); }   // general.py:706
    __device__ __forceinline__ float4 operator()(int b, unsigned int row) const { return finish(fetch(b, row)); }
};

template <typename T>
struct BoxFromLevels {
    HeadGeom g;
    // optional a-priori label rows appended after the A anchor rows (general.py:690-696): (bs, extra_rows, 5) [cls, x, y, w, h]
    const float *extra = nullptr;
    int extra_rows = 0;
    __device__ __forceinline__ RawBox fetch(int b, unsigned int row) const {
        if (row >= (unsigned)g.A) {
            const float *r = extra + ((size_t)b * extra_rows + (row - (unsigned)g.A)) * 5;
            return RawBox{__ldg(r + 1), __ldg(r + 2), __ldg(r + 3), __ldg(r + 4), -1, 0, 0};       // general.py:692
        }
        const RowGeom rg = row_geom(g, (int)row);
        const T *base = static_cast<const T *>(g.raw[rg.l]) + ((size_t)(b * g.na + rg.a) * g.no) * rg.plane + rg.cell;
        return RawBox{Elem<T>::ld1c(base), Elem<T>::ld1c(base + rg.plane), Elem<T>::ld1c(base + 2 * (size_t)rg.plane),
                      Elem<T>::ld1c(base + 3 * (size_t)rg.plane), rg.l, rg.a, rg.cell};
    }
    __device__ __forceinline__ float4 finish(const RawBox &q) const {
        if (q.l < 0) return xywh2xyxy(q.x, q.y, q.w, q.h);                                           // general.py:706
        const int l = q.l;
        const float x = decode_xy(q.x, q.cell % g.nx[l], g.stride[l]);
        const float y = decode_xy(q.y, q.cell / g.nx[l], g.stride[l]);
        const float w = decode_wh(q.w, g.anchor_w[l][q.a]);
        const float h = decode_wh(q.h, g.anchor_h[l][q.a]);
        return xywh2xyxy(x, y, w, h);
    }
    __device__ __forceinline__ float4 operator()(int b, unsigned int row) const { return finish(fetch(b, row)); }
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
constexpr int kChainSplit = kNmsThreads / 256;   // = kNmsThreads / kRound: threads per candidate in the survivor test

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
    // Survivors of one class are split over kChainSplit interleaved chains (by slot parity), so that the threads testing
    // one candidate against the survivors each walk a fraction of them.
    int khead[kChainClasses][kChainSplit];   // newest survivor slot of (class, slot % kChainSplit) (-1: none); slots < kKeptSmem only
    short knext[kKeptSmem];           // next older survivor of the same class and parity
    short lastw[kRound / 32][kChainClasses];   // last candidate of each class inside each warp of the round (-1: none)
    short rprev[kRound];              // previous (better-scored) candidate of the same class inside the round
    short ralive[kRound];             // ... the previous one that survived the test against the survivors so far
    short rslot[kRound];              // survivor slot of the round's candidate (-1: not kept)
    unsigned short rcls[kRound];      // class of the round's candidates
    unsigned int m;                   // gathered count
    unsigned int below;
    unsigned int red_sum, red_max, red_min;
    int nkept;
};

// mask[row] |= bit col, as a native 32-bit shared-memory atomic (64-bit ones are compare-and-swap loops)
__device__ __forceinline__ void mask_set(NmsSmem &s, int row, int col) {
    atomicOr(reinterpret_cast<unsigned int *>(&s.mask[row][0]) + (col >> 5), 1u << (col & 31));
}

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
// On success the histogram is left as the FIRST FREE POSITION of every bin (exclusive prefix) and *bin_lo / *bin_sh /
// *crowded describe the bins, so that the gather can drop each key straight into its bin (counting sort): bin of key k =
// (k - *bin_lo) >> *bin_sh; *crowded != 0: some bin of the chunk holds more than kBinSortMax keys (ties in the confidence).
constexpr unsigned int kBinSortMax = 16;
static __device__ int select_chunk(NmsSmem &s, const unsigned long long *keys, int n, unsigned long long lo,
                            unsigned long long hi, unsigned int target, unsigned long long *chunk_hi,
                            unsigned int *below, unsigned long long *bin_lo, int *bin_sh, int *crowded) {
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
            unsigned int first = wbase + incl - mine, big = 0;
#pragma unroll
            for (int j = 0; j < per; ++j) {
                const unsigned int bin = threadIdx.x * per + j;
                s.hist[bin] = first;
                big |= (bin < jmax && h[j] > kBinSortMax) ? 1u : 0u;
                first += h[j];
            }
            *crowded = __syncthreads_or((int)big);         // (also: every bin start is written before the gather reads it)
            *bin_lo = lo;
            *bin_sh = sh;
            return (int)cnt;
        }
        // bins [0, jmax) are empty and bin jmax alone exceeds kChunk: refine inside it
        lo = lo + ((unsigned long long)jmax << sh);        // <= hi: the bin holds keys <= hi
        const unsigned long long bw = (1ull << sh) - 1ull;
        hi = bw > hi - lo ? hi : lo + bw;
        below = nullptr;
    }
}

// Bitonic sort of P = E * kNmsThreads keys in shared memory `a` (ascending).  Thread t keeps the E consecutive keys
// a[t*E .. t*E+E) in registers: partners at distance j < E are exchanged in registers, at E <= j < 32*E by warp shuffles,
// and only at j >= 32*E (a partner in another warp: 10 of the 55 / 66 steps at P = 1024 / 2048) through shared memory,
// alternating between `a` and `scratch` so that one barrier per such step suffices.
template <int E>
static __device__ void bitonic_sort_regs(unsigned long long *a, unsigned long long *scratch) {
    constexpr int P = E * kNmsThreads;
    const int t = threadIdx.x;
    unsigned long long v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = a[t * E + e];
    bool to_scratch = true;
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32 * E) {
                unsigned long long *buf = to_scratch ? scratch : a;
                to_scratch = !to_scratch;
#pragma unroll
                for (int e = 0; e < E; ++e) buf[t * E + e] = v[e];
                __syncthreads();
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int i = t * E + e;
                    const unsigned long long pv = buf[i ^ j];
                    const bool keep_min = ((i & j) == 0) == ((i & k) == 0);
                    v[e] = keep_min ? (pv < v[e] ? pv : v[e]) : (pv > v[e] ? pv : v[e]);
                }
            } else if (j >= E) {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int i = t * E + e;
                    const unsigned long long pv = __shfl_xor_sync(0xffffffffu, v[e], j / E);
                    const bool keep_min = ((i & j) == 0) == ((i & k) == 0);
                    v[e] = keep_min ? (pv < v[e] ? pv : v[e]) : (pv > v[e] ? pv : v[e]);
                }
            } else {
                // static register indices (a run-time e ^ j would put v[] in local memory).  Direction of element
                // i = t * E + e: ascending iff (i & k) == 0, i.e. (e & k) == 0 while k < E and ((t * E) & k) == 0 from then on
                const bool up = ((t * E) & k) == 0;
                auto cx = [&](unsigned long long &x, unsigned long long &y, bool ascending) {
                    if ((x > y) == ascending) { const unsigned long long tmp = x; x = y; y = tmp; }
                };
                auto step = [&](auto jc) {
                    constexpr int J = decltype(jc)::value;
                    if constexpr (J < E) {
#pragma unroll
                        for (int e = 0; e < E; ++e)
                            if ((e & J) == 0) cx(v[e], v[e | J], k < E ? (e & k) == 0 : up);
                    }
                };
                if (j == 1) step(std::integral_constant<int, 1>{});
                else if (j == 2) step(std::integral_constant<int, 2>{});
                else if (j == 4) step(std::integral_constant<int, 4>{});
                else step(std::integral_constant<int, 8>{});
            }
        }
    }
    __syncthreads();                     // the last exchange step may still be reading `a`
#pragma unroll
    for (int e = 0; e < E; ++e) a[t * E + e] = v[e];
    __syncthreads();
}
static __device__ void bitonic_sort(unsigned long long *a, unsigned long long *scratch, int n_pow2) {
    static_assert(kChunk <= 8 * kNmsThreads, "a thread sorts at most 8 keys in registers");
    if (n_pow2 <= kNmsThreads) bitonic_sort_regs<1>(a, scratch);          // callers pad to kNmsThreads keys at least
    else if (n_pow2 == 2 * kNmsThreads) bitonic_sort_regs<2>(a, scratch);
    else if (n_pow2 == 4 * kNmsThreads) bitonic_sort_regs<4>(a, scratch);
    else bitonic_sort_regs<8>(a, scratch);
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
    // candidates wanted per selection.  First chunk: survivors rarely need more than ~2x max_det candidates (the
    // real-time setting never gets past it); later chunks are as large as the sort allows
    unsigned int target = (unsigned)min(kChunk, max(kRound, 2 * p.max_det));
    if (threadIdx.x == 0) {
        s.nkept = 0;
        s.ext[0] = s.ext[1] = INFINITY;
        s.ext[2] = s.ext[3] = -INFINITY;
        s.ext_bad = 0;
    }
    const bool chainable = p.nc <= kChainClasses;
    for (int i = threadIdx.x; i < kChainClasses * kChainSplit; i += kNmsThreads) (&s.khead[0][0])[i] = -1;
    // every key is <= key_hi (conf > conf_thres); with a confidence cut in force, <= the cut: a tighter range spreads the
    // keys over more histogram bins (fewer shared-memory atomic conflicts, no refinement pass)
    unsigned long long key_hi = p.key_hi;
    if (p.cut_hi != nullptr) key_hi = min(key_hi, ((unsigned long long)p.cut_hi[b] << 32) | 0xFFFFFFFFull);
    int consumed = 0;               // candidates taken in sorted order so far (max_nms cut)
    int seen = 0;                   // keys gathered so far
    unsigned long long cur = 0ull;
    bool first = true;
    bool done = false;              // max_det survivors found
    __syncthreads();
    DSN_LAP_START();

    for (;;) {
        // ---- 1. next chunk of keys -------------------------------------------------------------
        unsigned long long chunk_hi = ~0ull, bin_lo = 0ull;
        int m, bin_sh = 0, crowded = 0;
        bool binned = true;               // the chunk came out of select_chunk: its histogram positions the keys
        if (first && n <= kChunk) {
            m = n;
            binned = false;
        } else if (first) {
            unsigned int below = 0;
            m = select_chunk(s, keys, n, p.key_lo_hint, key_hi, target, &chunk_hi, &below, &bin_lo, &bin_sh, &crowded);
            if (below) m = select_chunk(s, keys, n, 0ull, key_hi, target, &chunk_hi, nullptr, &bin_lo, &bin_sh, &crowded);   // conf > 1.0 exists
            else cur = p.key_lo_hint;
        } else {
            m = select_chunk(s, keys, n, cur, key_hi, target, &chunk_hi, nullptr, &bin_lo, &bin_sh, &crowded);
        }
        first = false;
        if (m == 0) break;
        DSN_MARK(10);
        DSN_LAP(32);   // select
        // ---- 2. gather + sort ------------------------------------------------------------------
        int mp = 32;
        while (mp < m) mp <<= 1;
        unsigned long long *scratch = reinterpret_cast<unsigned long long *>(s.box);      // s.box is free until step 3
        if (binned) {
            // counting sort: every key of the chunk goes to the next free position of its histogram bin (the bins are in
            // key order), then the few keys sharing a bin are ordered among themselves - O(m) instead of a bitonic network
            for_each_key(keys, n, [&](unsigned long long k) {
                if (k >= cur && k <= chunk_hi) s.key[atomicAdd(&s.hist[(unsigned int)((k - bin_lo) >> bin_sh)], 1u)] = k;
            });
            __syncthreads();
            DSN_MARK(11);
            DSN_LAP(33);   // gather
            if (!crowded) {
                for (int q = threadIdx.x; q < m; q += kNmsThreads) {
                    const unsigned long long k = s.key[q];
                    const unsigned int bin = (unsigned int)((k - bin_lo) >> bin_sh);
                    const unsigned int st = bin ? s.hist[bin - 1] : 0u, en = s.hist[bin];      // after the gather: hist = bin ends
                    unsigned int r = st;
                    for (unsigned int i = st; i < en; ++i) r += s.key[i] < k ? 1u : 0u;
                    scratch[en - st > 1u ? r : (unsigned)q] = k;
                }
                __syncthreads();
                for (int q = threadIdx.x; q < m; q += kNmsThreads) s.key[q] = scratch[q];
                __syncthreads();
            } else {
                mp = max(mp, kNmsThreads);
                for (int i = m + threadIdx.x; i < mp; i += kNmsThreads) s.key[i] = ~0ull;
                __syncthreads();
                bitonic_sort(s.key, scratch, mp);
            }
        } else {
            if (!stashed) {                     // stashed implies first chunk, m == n <= kRankSortMax, keys in s.key
                if (threadIdx.x == 0) s.m = 0u;
                if (m > kRankSortMax) {
                    mp = max(mp, kNmsThreads);          // the register sort works on kNmsThreads keys or more
                    for (int i = threadIdx.x; i < mp; i += kNmsThreads) s.key[i] = ~0ull;
                }
                __syncthreads();
                for_each_key(keys, n, [&](unsigned long long k) {
                    // one shared-memory atomic per warp and call instead of one per key
                    const unsigned int act = __activemask();
                    const bool take = k >= cur && k <= chunk_hi;
                    const unsigned int bal = __ballot_sync(act, take);
                    if (bal) {
                        const int leader = __ffs(bal) - 1;
                        unsigned int base = 0;
                        if (lane == leader) base = atomicAdd(&s.m, (unsigned)__popc(bal));
                        base = __shfl_sync(act, base, leader);
                        if (take) s.key[base + __popc(bal & ((1u << lane) - 1u))] = k;
                    }
                });
                __syncthreads();
            }
            DSN_MARK(11);
            DSN_LAP(33);   // gather
            if (m <= kRankSortMax) rank_sort(s.key, m);
            else bitonic_sort(s.key, scratch, mp);
        }
        DSN_MARK(12);
        DSN_LAP(34);   // sort
        DSN_COUNT(52, 1);              // chunks
        DSN_COUNT(53, m);              // keys gathered
        seen += m;
        if (m > p.max_nms - consumed) m = p.max_nms - consumed;          // general.py:728-729
        consumed += m;
        // ---- 3. boxes --------------------------------------------------------------------------
        for (int i0 = 0; i0 < m; i0 += 2 * kNmsThreads) {
            // two candidates per thread and pass: both rows' loads are in flight before either is decoded
            const int ia = i0 + threadIdx.x, ib = ia + kNmsThreads;
            unsigned int refa = 0, refb = 0, rowa = 0, rowb = 0;
            RawBox qa{}, qb{};
            if (ia < m) {
                refa = (unsigned int)s.key[ia];
                rowa = refa / (unsigned)p.nc;
                qa = src.fetch(b, rowa);
            }
            if (ib < m) {
                refb = (unsigned int)s.key[ib];
                rowb = refb / (unsigned)p.nc;
                qb = src.fetch(b, rowb);
            }
            if (ia < m) {
                s.box[ia] = src.finish(qa);
                s.off[ia] = __fmul_rn((float)(refa - rowa * (unsigned)p.nc), p.off_mul);      // general.py:732
            }
            if (ib < m) {
                s.box[ib] = src.finish(qb);
                s.off[ib] = __fmul_rn((float)(refb - rowb * (unsigned)p.nc), p.off_mul);
            }
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
                // previous candidate of the same class: inside the warp with one MATCH; across warps through a table of the
                // last candidate of every class in every warp (at most kRound / 32 - 1 look-ups, whatever the class count)
                for (int i = threadIdx.x; i < (kRound / 32) * p.nc; i += kNmsThreads) s.lastw[i / p.nc][i % p.nc] = -1;
                __syncthreads();
                const bool in_round = threadIdx.x < kRound;
                const bool valid = in_round && threadIdx.x < cnt;
                const unsigned int c = valid ? (unsigned)s.rcls[threadIdx.x] : 0xFFFFFFFFu;
                unsigned int peers = 0;
                if (in_round) {                                   // warps 0 .. kRound/32 - 1, all lanes
                    peers = __match_any_sync(0xffffffffu, c);
                    if (valid && (peers >> lane) == 1u) s.lastw[warp][c] = (short)threadIdx.x;      // no same-class lane above me
                }
                __syncthreads();
                if (valid) {
                    const unsigned int below = peers & ((1u << lane) - 1u);
                    int j = -1;
                    if (below) {
                        j = (int)(threadIdx.x & ~31u) + (31 - __clz(below));
                    } else {
                        for (int w = warp - 1; w >= 0 && j < 0; --w) j = s.lastw[w][c];
                    }
                    s.rprev[threadIdx.x] = (short)j;
                }
                __syncthreads();
            }
            static_assert(kNmsThreads == kChainSplit * kRound, "the chain path tests a candidate with kChainSplit threads");
            DSN_LAP(45);   // rprev chains
            DSN_COUNT(50, 1);          // rounds
            DSN_COUNT(51, cnt);        // candidates consumed
            if (use_chain) {
                // (a) against the survivors of the candidate's own class: thread (j, half) walks the chain of that parity,
                //     four hops at a time (the hops are the serial part, the tests of a group are independent), and leaves
                //     at the first hit - at the eval setting nine candidates in ten die here
                {
                    const int j = threadIdx.x & (kRound - 1), half = threadIdx.x / kRound;
                    bool dead = false;
                    if (j < cnt && nkept0 > 0) {
                        const float4 bj = s.rbox[j];
                        const float aj = s.rarea[j], oj = s.roff[j];
                        int k = s.khead[s.rcls[j]][half];
                        while (k >= 0) {
                            const int k1 = s.knext[k];
                            const int k2 = k1 >= 0 ? s.knext[k1] : -1;
                            const int k3 = k2 >= 0 ? s.knext[k2] : -1;
                            bool hit = suppresses(s.kbox[k], s.karea[k], bj, aj, p.thr_f);
                            if (k1 >= 0) hit = suppresses(s.kbox[k1], s.karea[k1], bj, aj, p.thr_f) || hit;
                            if (k2 >= 0) hit = suppresses(s.kbox[k2], s.karea[k2], bj, aj, p.thr_f) || hit;
                            if (k3 >= 0) hit = suppresses(s.kbox[k3], s.karea[k3], bj, aj, p.thr_f) || hit;
                            if (hit) { dead = true; break; }
                            k = k3 >= 0 ? s.knext[k3] : -1;
                        }
                        // survivors beyond the shared-memory cache (max_det > kKeptSmem): split by parity as well
                        for (int kk = kKeptSmem + half; kk < nkept0 && !dead; kk += kChainSplit) {
                            const float *r = det + (size_t)kk * 6;
                            const float ok = __fmul_rn(r[5], p.off_mul);
                            if (ok != oj) continue;
                            const float4 bk = offset_box(make_float4(r[0], r[1], r[2], r[3]), ok);
                            dead = suppresses(bk, box_area(bk), bj, aj, p.thr_f);
                        }
                    }
                    const unsigned int bal = __ballot_sync(0xffffffffu, dead);
                    if (lane == 0 && bal) atomicOr(&s.dead[warp & (kRound / 32 - 1)], bal);
                }
                __syncthreads();
                DSN_LAP(46);   // (a) chains vs kept + sync
                // (b) inside the round, among the candidates still alive: link each to the nearest better-scored alive
                //     candidate of its class, then test it against that chain
                if (threadIdx.x < cnt) {
                    int j = s.rprev[threadIdx.x];
                    while (j >= 0 && ((s.dead[j >> 5] >> (j & 31)) & 1u)) j = s.rprev[j];
                    s.ralive[threadIdx.x] = (short)j;
                }
                __syncthreads();
                DSN_LAP(47);   // alive links + sync
                {
                    // thread (i, half): the hop is a dependent shared-memory load, the test ~10x that - both threads of a
                    // candidate walk its whole chain and each tests every other hop
                    const int i = threadIdx.x & (kRound - 1), half = threadIdx.x / kRound;
                    if (i < cnt && !((s.dead[i >> 5] >> (i & 31)) & 1u)) {
                        const float4 bi = s.rbox[i];
                        const float ai = s.rarea[i];
                        int hop = 0;
#pragma unroll 2
                        for (int j = s.ralive[i]; j >= 0; j = s.ralive[j], ++hop) {
                            if ((hop & (kChainSplit - 1)) != half) continue;
                            if (suppresses(s.rbox[j], s.rarea[j], bi, ai, p.thr_f)) {
                                mask_set(s, j, i);
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
                            mask_set(s, i, j);
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
            if (threadIdx.x < kRound) {
                int slot = -1;
                if (threadIdx.x < cnt) {
                    const int w = threadIdx.x >> 6, bit = threadIdx.x & 63;
                    const unsigned long long kb = s.keptbits[w];
                    if ((kb >> bit) & 1ull) {
                        slot = nkept0 + __popcll(kb & ((1ull << bit) - 1ull));
                        for (int v = 0; v < w; ++v) slot += __popcll(s.keptbits[v]);
                    }
                }
                s.rslot[threadIdx.x] = (short)slot;
                if (slot >= 0) {
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
                        if (chainable) my_slot = slot;
                    }
                }
            }
            if (chainable) {
                __syncthreads();                                   // every rslot is written
                if (my_slot >= 0) {
                    // link behind the previous survivor of the class with the same slot parity: the nearest such kept
                    // candidate up the round's chain, else that chain's newest survivor of earlier rounds
                    int j = s.rprev[threadIdx.x];
                    while (j >= 0 && (s.rslot[j] < 0 || ((s.rslot[j] ^ my_slot) & (kChainSplit - 1)))) j = s.rprev[j];
                    s.knext[my_slot] = j >= 0 ? s.rslot[j] : (short)s.khead[s.rcls[threadIdx.x]][my_slot & (kChainSplit - 1)];
                }
                __syncthreads();                                   // every knext is written before a head moves
                if (my_slot >= 0) atomicMax(&s.khead[s.rcls[threadIdx.x]][my_slot & (kChainSplit - 1)], my_slot);
            }
            __syncthreads();
            DSN_LAP(41);   // (d) append
            done = s.nkept >= p.max_det;
        }
        if (done || consumed >= p.max_nms || seen >= n) break;
        if (chunk_hi >= key_hi) break;
        cur = chunk_hi + 1ull;
        // an image that needs a second chunk is in the many-candidates regime (a few percent survive, and the survival
        // rate keeps falling as survivors accumulate): pay for select + gather + sort as rarely as possible
        target = (unsigned)kChunk;
    }
    __syncthreads();
    DSN_MARK(16);
    DSN_COUNT_ALL(54, 55, consumed);      // candidates consumed: sum / max over the images
    DSN_COUNT_ALL(56, 57, seen);          // keys gathered (all chunks)
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
    p.cut_hi = nullptr;
    p.det = det;
    p.det_count = det_count;
    return DSN_OK;
}


}  // inline namespace DSN_NMS_NS
}  // namespace dsn
