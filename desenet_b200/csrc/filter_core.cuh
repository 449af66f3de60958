// Candidate-filter building blocks shared by the standalone filter kernels (detect.cu) and the fused
// detection kernel (fused.cu): core/utils/general.py:668-718 of the reference.
#pragma once
#include <cmath>

#include "common.cuh"
#include "geom.cuh"

namespace dsn {

// =============================================================================================
// Candidate filter
// =============================================================================================
struct FilterOut {
    unsigned long long *keys;
    long long cap;
    int *counts;
    float conf_thres;
    int multi_label;
    unsigned long long cmask[4];
    int use_cmask;
    // optional confidence cut (dsn_conf_cut_levels): candidates whose key's confidence half exceeds cut_hi[b] are not
    // emitted (they can only matter beyond the max_nms best)
    const unsigned int *cut_hi;
};

__device__ __forceinline__ bool class_allowed(const FilterOut &f, int cls) {
    if (!f.use_cmask) return true;
    const int w = (cls >> 6) & 3;            // static indices: a dynamic one would put the by-value parameter in local memory
    const unsigned long long m = w == 0 ? f.cmask[0] : (w == 1 ? f.cmask[1] : (w == 2 ? f.cmask[2] : f.cmask[3]));
    return (m >> (cls & 63)) & 1ull;
}

// Warp-aggregated slot allocation: every lane passes its emission count; returns the lane's first
// slot inside keys[b][*].  One global atomic per warp (ballot-style compaction of survivors).
__device__ __forceinline__ int warp_alloc(int *counter, int n_mine) {
    const int lane = threadIdx.x & 31;
    int incl = n_mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane == 31 && total > 0) base = atomicAdd(counter, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    return base + incl - n_mine;
}

constexpr int kClsChunk = 8;

// One row per lane.  `Cls` provides load(c) -> raw class value of this lane's row and prob(raw) ->
// class probability (sigmoid for raw head logits, identity for a decoded prediction).  The raw
// values of a chunk of kClsChunk classes are all loaded before any is used, so a row costs ONE
// memory round trip per chunk.  obj is the objectness probability; lanes whose row fails
// `obj > conf_thres` (or !valid) skip every class load.  Must be called by all 32 lanes.
// `counter` = the image's candidate counter: f.counts + b (global) or a CTA-local shared-memory word.
// Default emission sink: the key goes to the image's key buffer in global memory.
struct KeySinkGlobal {
    unsigned long long *kout;
    int cap;               // candidate capacity of the image's key buffer (fits 32 bits: A * nc < 2^32 and counts are int)
    __device__ __forceinline__ KeySinkGlobal(unsigned long long *k, long long c) : kout(k), cap(c > 0x7fffffffLL ? 0x7fffffff : (int)c) {}
    __device__ __forceinline__ void operator()(int slot, unsigned long long key, int /*cls*/) const {
        if (slot < cap) kout[slot] = key;
    }
};

template <class Cls, class Sink>
__device__ __forceinline__ void filter_emit(const FilterOut &f, unsigned int row, int nc, float obj, bool valid,
                                            const Cls &cls, int *counter, const Sink &sink,
                                            const unsigned int *cut_ptr = nullptr) {
    const bool pass = valid && obj > f.conf_thres;                                       // general.py:668,687
    if (!__any_sync(0xffffffffu, pass)) return;
    const bool cut_on = cut_ptr != nullptr;
    const unsigned int cut = cut_on ? *cut_ptr : 0xFFFFFFFFu;
    float best = 0.f;
    int bcls = 0;
    if (!f.multi_label && nc <= kClsChunk) {
        // Best-class mode, all classes in one chunk.  sigmoid and the product with obj > 0 are monotone,
        // so the largest class LOGIT bounds every conf of the row: if even that conf is clearly below
        // the threshold the row emits nothing and the other sigmoids are never evaluated.  Rows that
        // (nearly) reach the threshold take the exact path below: argmax over the rounded products,
        // first maximum wins (general.py:713).
        float raw[kClsChunk];
        bool exact = false;
        if (pass) {
            float mx = -INFINITY;
#pragma unroll
            for (int ci = 0; ci < kClsChunk; ++ci) {
                raw[ci] = ci < nc ? cls.load(ci) : -INFINITY;
                mx = fmaxf(mx, raw[ci]);                      // NaN logits are ignored here ...
                exact = exact || raw[ci] != raw[ci];          // ... and force the exact path
            }
            const float top = __fmul_rn(cls.prob(mx), obj);
            exact = exact || !(top <= f.conf_thres * 0.999f);
        }
        if (exact) {
#pragma unroll
            for (int ci = 0; ci < kClsChunk; ++ci) {
                if (ci < nc) {
                    const float conf = __fmul_rn(cls.prob(raw[ci]), obj);                     // general.py:703
                    if (ci == 0 || argmax_better(conf, best)) { best = conf; bcls = ci; }
                }
            }
        }
        bool emit = exact && best > f.conf_thres && class_allowed(f, bcls);                  // general.py:714,717
        if (cut_on) emit = emit && conf_key(best) <= cut;
        const int slot = warp_alloc(counter, emit ? 1 : 0);
        if (emit) sink(slot, ((unsigned long long)conf_key(best) << 32) | (unsigned long long)(row * (unsigned)nc + bcls), bcls);
        return;
    }
    for (int c0 = 0; c0 < nc; c0 += kClsChunk) {
        float raw[kClsChunk];
        if (pass) {
#pragma unroll
            for (int ci = 0; ci < kClsChunk; ++ci) raw[ci] = (c0 + ci < nc) ? cls.load(c0 + ci) : 0.f;
        }
        if (!f.multi_label) {
            // best class, first maximum wins                                   general.py:713
            if (pass) {
#pragma unroll
                for (int ci = 0; ci < kClsChunk; ++ci) {
                    const int c = c0 + ci;
                    if (c < nc) {
                        const float conf = __fmul_rn(cls.prob(raw[ci]), obj);                 // general.py:703
                        if (c == 0 || argmax_better(conf, best)) { best = conf; bcls = c; }
                    }
                }
            }
        } else {
            // every (row, class) with conf > thres                               general.py:710-711
            float cf[kClsChunk];
            unsigned int emit = 0;
            if (pass) {
#pragma unroll
                for (int ci = 0; ci < kClsChunk; ++ci) {
                    if (c0 + ci < nc) {
                        cf[ci] = __fmul_rn(cls.prob(raw[ci]), obj);
                        emit |= cf[ci] > f.conf_thres ? (1u << ci) : 0u;
                    }
                }
                if (f.use_cmask) {                            // rare: filter the chunk's classes in one go
#pragma unroll
                    for (int ci = 0; ci < kClsChunk; ++ci)
                        if (!class_allowed(f, c0 + ci)) emit &= ~(1u << ci);
                }
            }
            if (cut_on) {                                     // confidence cut: keep only keys that can reach the top max_nms
                unsigned int keep = 0;
#pragma unroll
                for (int ci = 0; ci < kClsChunk; ++ci)
                    if (((emit >> ci) & 1u) && conf_key(cf[ci]) <= cut) keep |= 1u << ci;
                emit = keep;
            }
            if (!__any_sync(0xffffffffu, emit != 0)) continue;
            int slot = warp_alloc(counter, __popc(emit));
            const unsigned int ref0 = row * (unsigned)nc + (unsigned)c0;
#pragma unroll
            for (int ci = 0; ci < kClsChunk; ++ci)
                if ((emit >> ci) & 1u) {
                    sink(slot, ((unsigned long long)conf_key(cf[ci]) << 32) | (unsigned long long)(ref0 + ci), c0 + ci);
                    ++slot;
                }
        }
    }
    if (!f.multi_label) {
        bool emit = pass && best > f.conf_thres && class_allowed(f, bcls);                   // general.py:714,717
        if (cut_on) emit = emit && conf_key(best) <= cut;
        const int slot = warp_alloc(counter, emit ? 1 : 0);
        if (emit) sink(slot, ((unsigned long long)conf_key(best) << 32) | (unsigned long long)(row * (unsigned)nc + bcls), bcls);
    }
}

template <typename T>
struct ClsFromPlanes {
    const T *p;            // class-0 plane of this (image, anchor), already offset to the cell
    size_t plane;
    __device__ __forceinline__ float load(int c) const { return Elem<T>::ld1(p + (size_t)c * plane); }
    __device__ __forceinline__ float prob(float raw) const { return sigmoidf_ref(raw); }
};

// conservative logit-space bound: sigmoid(x) > conf_thres is impossible for x below it (0.1 % relative
// + 1e-3 absolute margin, orders of magnitude above the fp32 error of the device sigmoid)
// ... and the matching upper bound: sigmoid(x) > conf_thres is certain for x at or above it.
static inline float obj_logit_upper_bound(float conf_thres) {
    if (!(conf_thres > 0.f)) return INFINITY;          // conf 0: decide by the exact sigmoid (underflow cases)
    const double c = (double)conf_thres * (1.0 + 1e-3);
    if (c >= 1.0) return INFINITY;
    return std::nextafterf((float)(std::log(c / (1.0 - c)) + 1e-3), INFINITY);
}

static inline float obj_logit_lower_bound(float conf_thres) {
    float lo = -INFINITY;
    if (conf_thres > 0.f) {
        const double c = (double)conf_thres * (1.0 - 1e-3);
        lo = c >= 1.0 ? INFINITY : (float)(std::log(c / (1.0 - c)) - 1e-3);
        if (std::isfinite(lo)) lo = std::nextafterf(lo, -INFINITY);
    }
    return lo;
}

static inline int make_filter_out(FilterOut &f, float conf_thres, int multi_label, int nc, const uint64_t *class_mask,
                                  uint64_t *keys, int64_t cap, int *counts) {
    if (!keys || !counts || cap < 1) return DSN_E_INVALID;
    if (!(conf_thres >= 0.f && conf_thres <= 1.f)) return DSN_E_INVALID;     // general.py:671
    if (class_mask && nc > 256) return DSN_E_UNSUPPORTED;
    f.keys = (unsigned long long *)keys;
    f.cap = cap;
    f.counts = counts;
    f.conf_thres = conf_thres;
    f.multi_label = (multi_label && nc > 1) ? 1 : 0;                          // general.py:679
    f.use_cmask = class_mask != nullptr;
    for (int i = 0; i < 4; ++i) f.cmask[i] = class_mask ? class_mask[i] : ~0ull;
    f.cut_hi = nullptr;
    return DSN_OK;
}

}  // namespace dsn
