// Segmentation tail: x8 bilinear upsample (align_corners=True) + per-pixel argmax + mask write +
// confusion-matrix accumulation, fused (core/models/yolo.py:183, core/utils/plots.py:222-229,
// core/utils/metrics.py:350-388 of the reference).  HBM-bound streaming kernels: the low-res
// logits (bs*C*h*w*4 B) stay in L1/L2, the int64 target is read once with 16 B no-allocate
// loads, the fp32 mask is written once with 16 B streaming stores.  The upsampled logits are
// never materialised, yet every upsampled value is rounded exactly as ATen's CPU kernel rounds
// it (horizontal lerps first, each lerp = fma(l0, a, l1*b)), so argmax ties break identically.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace dsn {

struct SegParams {
    const void *logits;    // elements of the kernel's template type T (fp32 / fp16 / bf16)
    int bs, C, h, w, H, W;
    const long long *target;
    float *mask_f32;
    uint8_t *mask_u8;
    int mask_h, mask_w;
    unsigned long long *cm;
    int cm_slots, cm_rows;
    float sy, sx;      // (h-1)/(H-1), (w-1)/(W-1) in fp32 (0 when the output extent is 1)
    float my, mx;      // nearest scales float(H)/float(mask_h), float(W)/float(mask_w)
};

constexpr int kMaxPackedC = 4;      // template instantiations of the fast path
// build-time tuning switches (no run-time environment knobs in the launchers)
#ifndef DSN_SEG_EVAL_CTAS_PER_SM
#define DSN_SEG_EVAL_CTAS_PER_SM 0   // eval-mode seg kernel: cap on the resident CTAs per SM its persistent grid is sized for (0: all that fit)
#endif
#ifndef DSN_SEG_MASK_CTAS_PER_SM
#define DSN_SEG_MASK_CTAS_PER_SM 0   // detect-mode mask kernel: resident CTAs per SM the grid is sized for (0: one CTA per block)
#endif
#ifndef DSN_SEG_MASK_RPB
#define DSN_SEG_MASK_RPB 32          // ... output rows per block (work item)
#endif
#ifndef DSN_SEG_MASK_PX
#define DSN_SEG_MASK_PX 4            // ... pixels per thread (4 or 8)
#endif
constexpr int kPrefetchRows = 8;    // eval kernel: target rows pulled into L2 this far ahead of the register loads

// ---- packed confusion-matrix counters ------------------------------------------------------
// Per thread: 8-bit fields, one per target row, packed in a 32-bit word (cm_rows <= 4, the nclass = 2
// case) or a 64-bit word (cm_rows <= 8).  For C == 2 two words are kept — all pixels and pixels predicted
// as class 1 (class 0 = difference, taken after the reduction); otherwise one word per class.  A thread adds
// at most 255 pixels per field between flushes.  Flush: widen to 16-bit fields, REDUX-add across the warp
// (32 lanes * 255 < 65536), one shared-memory atomic per (row, class) per warp; one global atomic per
// (row, class) per CTA at the end.
template <int C, typename W>
struct CmAcc {
    static constexpr int kWords = (C == 2) ? 2 : C;
    W w[kWords];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < kWords; ++i) w[i] = 0;
    }
    __device__ __forceinline__ void add(int pred, int row) {
        const W inc = (W)1 << (8 * row);
        if (C == 2) {
            w[0] += inc;
            w[1] += pred ? inc : (W)0;
        } else {
#pragma unroll
            for (int c = 0; c < C; ++c) w[c] += (pred == c) ? inc : (W)0;
        }
    }
    __device__ __forceinline__ void flush(unsigned int *cm_smem, int rows) {
        const int lane = threadIdx.x & 31;
        unsigned int cnt[kWords];
#pragma unroll
        for (int i = 0; i < kWords; ++i) {
            // rows 0,2 | 1,3 | 4,6 | 5,7 as 16-bit fields of four 32-bit words
            const unsigned long long v = (unsigned long long)w[i];
            const unsigned int e_lo = (unsigned int)v & 0x00FF00FFu, o_lo = ((unsigned int)v >> 8) & 0x00FF00FFu;
            const unsigned int s_e_lo = __reduce_add_sync(0xffffffffu, e_lo), s_o_lo = __reduce_add_sync(0xffffffffu, o_lo);
            unsigned int s_e_hi = 0, s_o_hi = 0;
            if (sizeof(W) == 8) {
                const unsigned int hi = (unsigned int)(v >> 32);
                s_e_hi = __reduce_add_sync(0xffffffffu, hi & 0x00FF00FFu);
                s_o_hi = __reduce_add_sync(0xffffffffu, (hi >> 8) & 0x00FF00FFu);
            }
            // lane r extracts row r
            const unsigned int word = (lane & 4) ? ((lane & 1) ? s_o_hi : s_e_hi) : ((lane & 1) ? s_o_lo : s_e_lo);
            cnt[i] = (word >> (16 * ((lane >> 1) & 1))) & 0xFFFFu;
            w[i] = 0;
        }
        if (lane < rows) {
            if (C == 2) {
                const unsigned int ones = cnt[1], zeros = cnt[0] - cnt[1];
                if (zeros) atomicAdd(&cm_smem[lane * C + 0], zeros);
                if (ones) atomicAdd(&cm_smem[lane * C + 1], ones);
            } else {
#pragma unroll
                for (int c = 0; c < kWords; ++c)
                    if (cnt[c]) atomicAdd(&cm_smem[lane * C + c], cnt[c]);
            }
        }
    }
};

// XOR of the HIGH words of four int64 targets.  Only the low words carry information (clamp_target_row), but with the
// high words dead ptxas narrows each 16-byte target load into two 4-byte loads of the used words: twice the requests,
// each pulling a whole 32-byte sector for 4 useful bytes.  Folding the high words into a sink that a never-true branch
// stores keeps the loads 128 bits wide: eval seg kernel 375 -> 288 us at bs 512 / 640 px (0.69 -> 0.90 of the measured
// HBM peak), 101 -> 80 us at bs 32 / 1280 px.
__device__ __forceinline__ unsigned int hi_words(const longlong2 &a, const longlong2 &b) {
    return (unsigned int)(a.x >> 32) ^ (unsigned int)(a.y >> 32) ^ (unsigned int)(b.x >> 32) ^ (unsigned int)(b.y >> 32);
}

__device__ __forceinline__ int clamp_target_row(long long t, int rows) {
    // metrics.py:359,380 wrap the int64 target to int32 (`astype('int32')`); negatives are neither
    // "labeled" (target > 0) nor inside any histogram range -> same statistics as row 0.
    int v = (int)(unsigned int)(unsigned long long)t;
    v = v < 0 ? 0 : v;
    return v > rows - 1 ? rows - 1 : v;
}

// ---- fast path: upsampling by >= 4 in x and >= 1 in y, W % 4 == 0, mask (if any) at full resolution,
//      C in 2..4 ------------------------------------------------------------------------------------
// One CTA = kTileRows output rows x full width of one image; one thread = a 4-pixel column group
// that walks down the rows.  The thread keeps the HORIZONTALLY interpolated values of the two
// source rows it is between (T = row y0, B = row y0+1) in registers; when y0 advances (every ~8
// output rows) T <- B and B is rebuilt from three raw source values per class that were prefetched
// one source row ahead.  Per output pixel that leaves one FMUL + one FFMA per class, the compare
// and the store — the kernel has to stay below ~10 issue slots per pixel to remain HBM-bound.
struct RowCoef {
    float l0, l1;
    int y0;
};

// Vertical lerp of 4 pixels of one class with Blackwell's packed fp32x2 arithmetic (FMUL2 / FFMA2, sm_100+): two
// correctly rounded fp32 results per instruction, bit-identical to fmaf(l0, t, l1*b) per pixel, half the issue slots.
__device__ __forceinline__ void vlerp4(const float2 lm, const float (&t)[4], const float (&b)[4], float (&o)[4]) {
    const float2 l0 = make_float2(lm.x, lm.x), l1 = make_float2(lm.y, lm.y);
    const float2 m01 = __fmul2_rn(l1, make_float2(b[0], b[1]));
    const float2 m23 = __fmul2_rn(l1, make_float2(b[2], b[3]));
    const float2 r01 = __ffma2_rn(l0, make_float2(t[0], t[1]), m01);
    const float2 r23 = __ffma2_rn(l0, make_float2(t[2], t[3]), m23);
    o[0] = r01.x; o[1] = r01.y; o[2] = r23.x; o[3] = r23.y;
}

template <int C>
struct SrcCols {
    float v[C][3];
};

template <int C, typename T>
__device__ __forceinline__ SrcCols<C> load_src_row(const T *img, int plane, int row_off, int c0, int c1, int c2) {
    SrcCols<C> r;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const T *pl = img + (size_t)c * plane + row_off;
        r.v[c][0] = Elem<T>::ld1c(pl + c0);
        r.v[c][1] = Elem<T>::ld1c(pl + c1);
        r.v[c][2] = Elem<T>::ld1c(pl + c2);
    }
    return r;
}

// MASK: bit 0 = write mask_f32, bit 1 = write mask_u8.
// Persistent form: the bs*H output rows of the batch are split into gridDim.x CONTIGUOUS runs (one
// per CTA, grid = SMs x resident CTAs), so the per-thread x set-up is paid once per CTA.  Inside a
// run the rows are walked SEGMENT by segment: a segment = the output rows that interpolate between
// the same two source rows (8-9 rows at scale 8), found through a small shared-memory table, so the
// inner loop is branch-free: two LDS for the row's (l0, l1), C x 4 FMUL+FFMA, the compares, one
// 16 B streaming store.  Between segments T/B swap roles and the stale one is rebuilt from raw
// values prefetched one source row ahead.
template <typename TIn, int C, bool HAS_TGT, int MASK, typename W>
__global__ void __maxnreg__(HAS_TGT ? 96 : 128) seg_fused_rows_kernel(const SegParams p, const int rows_per_cta) {
    extern __shared__ __align__(16) unsigned char seg_smem[];
    float2 *lam = reinterpret_cast<float2 *>(seg_smem);                    // [H]   (l0, l1) of output row Y
    int *y0_of = reinterpret_cast<int *>(lam + p.H);                       // [H]   source row y0 of output row Y
    int *seg_start = y0_of + p.H;                                          // [h+1] first output row whose y0 is the index
    __shared__ unsigned int cm_smem[DSN_MAX_CM_ROWS * kMaxPackedC];
    const int G = p.W >> 2;
    const int plane = p.h * p.w;
    const long long total_rows = (long long)p.bs * p.H;
    const long long r_begin = (long long)blockIdx.x * rows_per_cta;
    const long long r_end = min(total_rows, r_begin + rows_per_cta);
    if (HAS_TGT && threadIdx.x < DSN_MAX_CM_ROWS * kMaxPackedC) cm_smem[threadIdx.x] = 0u;
    for (int Y = threadIdx.x; Y < p.H; Y += blockDim.x) {
        const Lerp ly = lerp_coeff(p.sy, Y, p.h);
        lam[Y] = make_float2(ly.l0, ly.l1);
        y0_of[Y] = ly.i0;
    }
    for (int y = threadIdx.x; y <= p.h; y += blockDim.x) seg_start[y] = p.H;   // sentinel for unused / last+1
    __syncthreads();
    for (int Y = threadIdx.x; Y < p.H; Y += blockDim.x)
        if (Y == 0 || y0_of[Y] != y0_of[Y - 1]) seg_start[y0_of[Y]] = Y;       // y0 is monotone, steps of <= 1 (H >= h)
    __syncthreads();
    CmAcc<C, W> acc;
    acc.clear();

    for (int g0 = 0; g0 < G; g0 += blockDim.x) {
        if (g0 + (int)(threadIdx.x & ~31u) >= G) break;        // whole warp beyond the row: nothing to do
        const int g = g0 + threadIdx.x;
        const bool active = g < G;
        const int X = (active ? g : 0) << 2;
        float l0[4], l1[4];
        bool shx[4];
        int xa;
        {
            const Lerp lx0 = lerp_coeff(p.sx, X, p.w);
            xa = lx0.i0;
            l0[0] = lx0.l0; l1[0] = lx0.l1; shx[0] = false;
#pragma unroll
            for (int k = 1; k < 4; ++k) {
                const Lerp lx = lerp_coeff(p.sx, X + k, p.w);
                l0[k] = lx.l0; l1[k] = lx.l1; shx[k] = lx.i0 != xa;     // source column advanced by one
            }
        }
        const int c0 = xa, c1 = min(xa + 1, p.w - 1), c2 = min(xa + 2, p.w - 1);
        // horizontal lerp of one source row; returns false if a raw value is inf/NaN          (yolo.py:183)
        auto hrow = [&](const SrcCols<C> &s, float (&out)[C][4]) -> bool {
            bool fin = true;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                fin = fin && (fabsf(s.v[c][0]) < INFINITY) && (fabsf(s.v[c][1]) < INFINITY) && (fabsf(s.v[c][2]) < INFINITY);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float a = shx[k] ? s.v[c][1] : s.v[c][0];
                    const float bb = shx[k] ? s.v[c][2] : s.v[c][1];
                    out[c][k] = __fmaf_rn(l0[k], a, __fmul_rn(l1[k], bb));
                }
            }
            return fin;
        };
        float *m32 = (MASK & 1) ? p.mask_f32 + (size_t)r_begin * p.W + X : nullptr;     // (b*H + Y)*W + X == r*W + X
        uint8_t *m8 = (MASK & 2) ? p.mask_u8 + (size_t)r_begin * p.W + X : nullptr;
        const long long *tp = HAS_TGT ? p.target + (size_t)r_begin * p.W + X : nullptr;
        const long long *tgt_end = HAS_TGT ? p.target + (size_t)total_rows * p.W : nullptr;
        int since_flush = 0;
        bool finite = true;
        unsigned int hi_sink = 0;
        // predictions of one output row only (no mask, no target): used by the pure-eval group path
        auto row_pred = [&](const float2 lm, const float (&top)[C][4], const float (&bot)[C][4], int (&pred)[4]) {
            if (C == 2 && finite) {
                float v0[4], v1[4];
                vlerp4(lm, top[0], bot[0], v0);
                vlerp4(lm, top[C - 1], bot[C - 1], v1);
#pragma unroll
                for (int k = 0; k < 4; ++k) pred[k] = v1[k] > v0[k] ? 1 : 0;
            } else {
                float vals[C][4];
#pragma unroll
                for (int c = 0; c < C; ++c) vlerp4(lm, top[c], bot[c], vals[c]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float best = vals[0][k];
                    int bi = 0;
#pragma unroll
                    for (int c = 1; c < C; ++c)
                        if (argmax_better(vals[c][k], best)) { best = vals[c][k]; bi = c; }
                    pred[k] = bi;
                }
            }
        };
        // one output row from the row pair (top, bot)
        auto row_body = [&](const float2 lm, const float (&top)[C][4], const float (&bot)[C][4], const long long (&tv)[4]) {
            int pred[4];
            float predf[4];
            if (C == 2 && finite) {
                // no NaN can appear: plain compare, first maximum wins
                float v0[4], v1[4];
                vlerp4(lm, top[0], bot[0], v0);
                vlerp4(lm, top[C - 1], bot[C - 1], v1);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    pred[k] = v1[k] > v0[k] ? 1 : 0;
                    predf[k] = v1[k] > v0[k] ? 1.0f : 0.0f;
                }
            } else {
                float vals[C][4];
#pragma unroll
                for (int c = 0; c < C; ++c) vlerp4(lm, top[c], bot[c], vals[c]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float best = 0.f;
                    int bi = 0;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const float val = vals[c][k];
                        if (c == 0) {
                            best = val;
                        } else if (argmax_better(val, best)) {
                            best = val;
                            bi = c;
                        }
                    }
                    pred[k] = bi;
                    predf[k] = (float)bi;
                }
            }
            if (active) {
                if (MASK & 1) stg_stream_f4(m32, make_float4(predf[0], predf[1], predf[2], predf[3]));
                if (MASK & 2)
                    *reinterpret_cast<uint32_t *>(m8) =
                        (uint32_t)pred[0] | ((uint32_t)pred[1] << 8) | ((uint32_t)pred[2] << 16) | ((uint32_t)pred[3] << 24);
                if (HAS_TGT) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc.add(pred[k], clamp_target_row(tv[k], p.cm_rows));
                }
            }
            if (MASK & 1) m32 += p.W;
            if (MASK & 2) m8 += p.W;
        };
        // output rows [Y, ye) of one segment; the target loads of 4 rows are issued together (8 x 16 B in flight)
        auto segment = [&](int Y, const int ye, const float (&top)[C][4], const float (&bot)[C][4]) {
            for (; Y + 4 <= ye; Y += 4) {
                long long tv[4][4];
                if (HAS_TGT) {
                    // Registers bound the bytes a thread can have in flight (8 x 16 B here), and a dependent DRAM round
                    // trip is ~2.5 us on this part, so the rows two groups ahead are pulled into L2 without registers:
                    // one prefetch per 128 B line (4 lanes x 32 B), i.e. by every fourth lane.
                    if (active && (threadIdx.x & 3) == 0) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const long long *pa = tp + (size_t)(q + kPrefetchRows) * p.W;
                            if (pa < tgt_end) asm volatile("prefetch.global.L2 [%0];" ::"l"(pa));
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        tv[q][0] = tv[q][1] = tv[q][2] = tv[q][3] = 0;
                        if (active) {
                            const longlong2 ta = ldg_stream_ll2(tp + (size_t)q * p.W);
                            const longlong2 tb = ldg_stream_ll2(tp + (size_t)q * p.W + 2);
                            tv[q][0] = ta.x; tv[q][1] = ta.y; tv[q][2] = tb.x; tv[q][3] = tb.y;
                            hi_sink ^= hi_words(ta, tb);          // keeps the loads 128 bits wide, see hi_words
                        }
                    }
                    tp += (size_t)4 * p.W;
                }
                if constexpr (HAS_TGT && MASK == 0 && C <= 4) {
                    // pure eval: the predictions of all four rows (independent of the target) are computed FIRST, packed
                    // two bits per pixel, so that ~300 instructions separate the target loads from their first use and
                    // the DRAM / L2 round trip overlaps this thread's own arithmetic; only then are the targets consumed
                    unsigned int packed = 0;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        int pr[4];
                        row_pred(lam[Y + q], top, bot, pr);
#pragma unroll
                        for (int k = 0; k < 4; ++k) packed |= (unsigned)pr[k] << (2 * (q * 4 + k));
                    }
                    if (active) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                acc.add((int)((packed >> (2 * (q * 4 + k))) & 3u), clamp_target_row(tv[q][k], p.cm_rows));
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) row_body(lam[Y + q], top, bot, tv[q]);
                }
                if (HAS_TGT) {
                    since_flush += 4;
                    if (since_flush >= 60) {              // 4 px per row: an 8-bit field holds at most 255
                        acc.flush(cm_smem, p.cm_rows);
                        since_flush = 0;
                    }
                }
            }
            for (; Y < ye; ++Y) {
                long long tv[4] = {0, 0, 0, 0};
                if (HAS_TGT) {
                    if (active) {
                        const longlong2 ta = ldg_stream_ll2(tp);
                        const longlong2 tb = ldg_stream_ll2(tp + 2);
                        tv[0] = ta.x; tv[1] = ta.y; tv[2] = tb.x; tv[3] = tb.y;
                        hi_sink ^= hi_words(ta, tb);
                    }
                    tp += p.W;
                    ++since_flush;
                }
                row_body(lam[Y], top, bot, tv);
            }
        };

        long long r = r_begin;
        while (r < r_end) {
            const int b = (int)(r / p.H);
            int Y = (int)(r - (long long)b * p.H);
            const int Yend = (int)min((long long)p.H, Y + (r_end - r));        // rows of image b inside this run
            r += Yend - Y;
            const TIn *img = static_cast<const TIn *>(p.logits) + (size_t)b * C * plane;
            int y0 = y0_of[Y];
            float T[C][4], B[C][4];
            {
                const SrcCols<C> s0 = load_src_row<C>(img, plane, y0 * p.w, c0, c1, c2);
                const SrcCols<C> s1 = load_src_row<C>(img, plane, min(y0 + 1, p.h - 1) * p.w, c0, c1, c2);
                const bool f0 = hrow(s0, T);
                const bool f1 = hrow(s1, B);
                finite = finite && f0 && f1;
            }
            for (;;) {
                // T = source row y0, B = y0+1
                SrcCols<C> nxt = load_src_row<C>(img, plane, min(y0 + 2, p.h - 1) * p.w, c0, c1, c2);   // prefetch
                int ye = min(seg_start[y0 + 1], Yend);
                segment(Y, ye, T, B);
                Y = ye;
                if (Y >= Yend) break;
                ++y0;
                finite = hrow(nxt, T) && finite;          // T <- source row y0+1 (new bottom); B is the new top
                nxt = load_src_row<C>(img, plane, min(y0 + 2, p.h - 1) * p.w, c0, c1, c2);
                ye = min(seg_start[y0 + 1], Yend);
                segment(Y, ye, B, T);
                Y = ye;
                if (Y >= Yend) break;
                ++y0;
                finite = hrow(nxt, B) && finite;          // back to T = top, B = bottom
            }
        }
        if (HAS_TGT) acc.flush(cm_smem, p.cm_rows);
        if (HAS_TGT && hi_sink == 0x9E3779B9u && p.bs < 0) p.cm[0] = hi_sink;      // never true: see hi_sink above
    }
    if (HAS_TGT) {
        __syncthreads();
        const int nb = p.cm_rows * C;
        if (threadIdx.x < nb) {
            const unsigned int cnt = cm_smem[threadIdx.x];
            if (cnt) atomicAdd(p.cm + (size_t)(blockIdx.x % p.cm_slots) * nb + threadIdx.x, (unsigned long long)cnt);
        }
    }
}

// ---- detect mode (mask only): the store stream ---------------------------------------------------------------------
// The fp32 mask is 98 % of the bytes of a detect-mode step (4 B per pixel out, 1/16 B per pixel in), so this kernel is a
// pure write stream with ~25 instructions of arithmetic per 16 bytes.  A thread owns a PX-pixel column group and keeps the
// horizontally interpolated source rows T / B in registers, as above; what differs is how the work is cut:
//   * a work item is a BLOCK of `rows_per_block` consecutive output rows of one image (full width), and CTA c takes the
//     blocks c, c + gridDim.x, ... — at any moment the resident CTAs write ONE compact, advancing window of the mask
//     instead of one far-apart run per CTA.  A pure store stream with the same cut reaches 6.2 TB/s on B200, against
//     5.2 TB/s for 296 separate runs and for TMA bulk stores out of shared memory (profiles/experiments/stbw.cu);
//   * no shared-memory tables and no barrier: row coefficients are computed on the fly (6 instructions per row), so a
//     CTA's first store is one L2 round trip (its two source rows) after launch.
template <int PX>
struct XCoef {
    float l0[PX], l1[PX];
    unsigned shx;          // bit k: pixel k interpolates between source columns (c1, c2) instead of (c0, c1)
    int c0, c1, c2;
};
template <int PX>
__device__ __forceinline__ XCoef<PX> make_xcoef(const SegParams &p, int X) {
    XCoef<PX> xc;
    const Lerp lx0 = lerp_coeff(p.sx, X, p.w);
    xc.l0[0] = lx0.l0; xc.l1[0] = lx0.l1; xc.shx = 0u;
#pragma unroll
    for (int k = 1; k < PX; ++k) {
        const Lerp lx = lerp_coeff(p.sx, X + k, p.w);
        xc.l0[k] = lx.l0; xc.l1[k] = lx.l1;
        xc.shx |= (lx.i0 != lx0.i0 ? 1u : 0u) << k;                  // source column advanced by one (never by two: see the dispatch)
    }
    xc.c0 = lx0.i0; xc.c1 = min(lx0.i0 + 1, p.w - 1); xc.c2 = min(lx0.i0 + 2, p.w - 1);
    return xc;
}
// horizontal lerp of one source row (yolo.py:183, horizontal first); returns false if a raw value is inf / NaN
template <int C, int PX>
__device__ __forceinline__ bool hrow_px(const XCoef<PX> &xc, const SrcCols<C> &sc, float (&out)[C][PX]) {
    bool fin = true;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        fin = fin && (fabsf(sc.v[c][0]) < INFINITY) && (fabsf(sc.v[c][1]) < INFINITY) && (fabsf(sc.v[c][2]) < INFINITY);
#pragma unroll
        for (int k = 0; k < PX; ++k) {
            const bool sh = (xc.shx >> k) & 1u;
            const float a = sh ? sc.v[c][1] : sc.v[c][0];
            const float bb = sh ? sc.v[c][2] : sc.v[c][1];
            out[c][k] = __fmaf_rn(xc.l0[k], a, __fmul_rn(xc.l1[k], bb));
        }
    }
    return fin;
}

template <typename TIn, int C, int MASK, int PX>
__global__ void __launch_bounds__(512) seg_mask_rows_kernel(const SegParams p, const int rows_per_block, const int blocks_per_image,
                                                            const long long nblocks) {
    static_assert(PX == 4 || PX == 8, "a thread owns 4 or 8 pixels of a row");
    const int G = p.W / PX;
    const int plane = p.h * p.w;
    for (int g0 = 0; g0 < G; g0 += blockDim.x) {
        if (g0 + (int)(threadIdx.x & ~31u) >= G) break;        // whole warp beyond the row: nothing to do
        const int g = g0 + threadIdx.x;
        const bool active = g < G;
        const int X = (active ? g : 0) * PX;
        const XCoef<PX> xc = make_xcoef<PX>(p, X);
        bool finite = true;
        for (long long blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
            const int b = (int)(blk / blocks_per_image);
            int Y = (int)(blk - (long long)b * blocks_per_image) * rows_per_block;
            const int Yend = min(p.H, Y + rows_per_block);
            const TIn *img = static_cast<const TIn *>(p.logits) + (size_t)b * C * plane;
            int y0 = lerp_coeff(p.sy, Y, p.h).i0;
            float T[C][PX], B[C][PX];
            SrcCols<C> nxt;
            {
                const SrcCols<C> s0 = load_src_row<C>(img, plane, y0 * p.w, xc.c0, xc.c1, xc.c2);
                const SrcCols<C> s1 = load_src_row<C>(img, plane, min(y0 + 1, p.h - 1) * p.w, xc.c0, xc.c1, xc.c2);
                nxt = load_src_row<C>(img, plane, min(y0 + 2, p.h - 1) * p.w, xc.c0, xc.c1, xc.c2);     // one source row ahead
                const bool f0 = hrow_px<C, PX>(xc, s0, T);
                const bool f1 = hrow_px<C, PX>(xc, s1, B);
                finite = finite && f0 && f1;
            }
            const size_t off = ((size_t)b * p.H + Y) * p.W + X;
            float *m32 = (MASK & 1) ? p.mask_f32 + off : nullptr;
            uint8_t *m8 = (MASK & 2) ? p.mask_u8 + off : nullptr;
#pragma unroll 2
            for (; Y < Yend; ++Y) {
                const Lerp ly = lerp_coeff(p.sy, Y, p.h);
                if (ly.i0 != y0) {                                   // CTA-uniform: the source row pair moves down by one
                    ++y0;
#pragma unroll
                    for (int c = 0; c < C; ++c)
#pragma unroll
                        for (int k = 0; k < PX; ++k) T[c][k] = B[c][k];
                    const bool f2 = hrow_px<C, PX>(xc, nxt, B);
                    finite = finite && f2;
                    nxt = load_src_row<C>(img, plane, min(y0 + 2, p.h - 1) * p.w, xc.c0, xc.c1, xc.c2);
                }
                const float2 lm = make_float2(ly.l0, ly.l1);
                int pred[PX];
                float predf[PX];
                if (C == 2 && finite) {
                    // no NaN can appear: plain compare, first maximum wins
#pragma unroll
                    for (int q = 0; q < PX; q += 4) {
                        float v0[4], v1[4];
                        vlerp4(lm, *reinterpret_cast<const float(*)[4]>(&T[0][q]), *reinterpret_cast<const float(*)[4]>(&B[0][q]), v0);
                        vlerp4(lm, *reinterpret_cast<const float(*)[4]>(&T[C - 1][q]), *reinterpret_cast<const float(*)[4]>(&B[C - 1][q]), v1);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            pred[q + k] = v1[k] > v0[k] ? 1 : 0;
                            predf[q + k] = v1[k] > v0[k] ? 1.0f : 0.0f;
                        }
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < PX; q += 4) {
                        float vals[C][4];
#pragma unroll
                        for (int c = 0; c < C; ++c)
                            vlerp4(lm, *reinterpret_cast<const float(*)[4]>(&T[c][q]), *reinterpret_cast<const float(*)[4]>(&B[c][q]), vals[c]);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            float best = vals[0][k];
                            int bi = 0;
#pragma unroll
                            for (int c = 1; c < C; ++c)
                                if (argmax_better(vals[c][k], best)) { best = vals[c][k]; bi = c; }
                            pred[q + k] = bi;
                            predf[q + k] = (float)bi;
                        }
                    }
                }
                if (active) {
#pragma unroll
                    for (int q = 0; q < PX; q += 4) {
                        if (MASK & 1) stg_stream_f4(m32 + q, make_float4(predf[q], predf[q + 1], predf[q + 2], predf[q + 3]));
                        if (MASK & 2)
                            *reinterpret_cast<uint32_t *>(m8 + q) = (uint32_t)pred[q] | ((uint32_t)pred[q + 1] << 8) |
                                                                    ((uint32_t)pred[q + 2] << 16) | ((uint32_t)pred[q + 3] << 24);
                    }
                }
                if (MASK & 1) m32 += p.W;
                if (MASK & 2) m8 += p.W;
            }
        }
    }
}

// ---- general path: any C, any scale, any mask size, any cm_rows -------------------------------
template <typename T>
__device__ __forceinline__ int pred_bilinear(const T *img, int C, int h, int w, float sy, float sx, int Y, int X) {
    const Lerp ly = lerp_coeff(sy, Y, h);
    const Lerp lx = lerp_coeff(sx, X, w);
    const int y1 = ly.i0 + (ly.i0 < h - 1), x1 = lx.i0 + (lx.i0 < w - 1);
    float best = 0.f;
    int bi = 0;
    for (int c = 0; c < C; ++c) {
        const T *pl = img + (size_t)c * h * w;
        const float t = __fmaf_rn(lx.l0, Elem<T>::ld1c(pl + ly.i0 * w + lx.i0), __fmul_rn(lx.l1, Elem<T>::ld1c(pl + ly.i0 * w + x1)));
        const float bt = __fmaf_rn(lx.l0, Elem<T>::ld1c(pl + y1 * w + lx.i0), __fmul_rn(lx.l1, Elem<T>::ld1c(pl + y1 * w + x1)));
        const float val = __fmaf_rn(ly.l0, t, __fmul_rn(ly.l1, bt));
        if (c == 0) {
            best = val;
        } else if (argmax_better(val, best)) {
            best = val;
            bi = c;
        }
    }
    return bi;
}

template <typename T>
__device__ __forceinline__ int pred_fullres(const T *img, int C, size_t plane, size_t off) {
    float best = Elem<T>::ld1c(img + off);
    int bi = 0;
    for (int c = 1; c < C; ++c) {
        const float val = Elem<T>::ld1c(img + (size_t)c * plane + off);
        if (argmax_better(val, best)) {
            best = val;
            bi = c;
        }
    }
    return bi;
}

// ATen legacy 'nearest' source index (plots.py:229): min(floor(dst * float(in/out)), in-1)
__device__ __forceinline__ int nearest_src(int dst, float scale, int n_in) {
    const int s = (int)floorf(__fmul_rn((float)dst, scale));
    return s < n_in - 1 ? s : n_in - 1;
}

// warp-aggregated shared-memory histogram update for an arbitrary number of bins
__device__ __forceinline__ void cm_add_match(unsigned int *cm_smem, int bin, bool valid) {
    const unsigned int act = __ballot_sync(0xffffffffu, valid);
    if (!valid) return;
    const unsigned int peers = __match_any_sync(act, bin);
    if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&cm_smem[bin], (unsigned int)__popc(peers));
}

template <typename T, bool FULLRES>
__global__ void __launch_bounds__(256) seg_general_kernel(const SegParams p) {
    extern __shared__ unsigned int cm_dyn[];
    const int b = blockIdx.y;
    const int nb = p.cm_rows * p.C;
    const size_t plane_lo = (size_t)p.h * p.w;
    const size_t plane_hi = (size_t)p.H * p.W;
    const T *img = static_cast<const T *>(p.logits) + (size_t)b * p.C * (FULLRES ? plane_hi : plane_lo);
    if (p.target) {
        for (int i = threadIdx.x; i < nb; i += blockDim.x) cm_dyn[i] = 0u;
        __syncthreads();
        const size_t npx = plane_hi;
        const size_t per = ((npx + gridDim.x - 1) / gridDim.x + 31) / 32 * 32;
        const size_t lo = per * blockIdx.x, hi = min(lo + per, npx);
        for (size_t i0 = lo; i0 < hi; i0 += blockDim.x) {
            const size_t i = i0 + threadIdx.x;
            const bool valid = i < hi;
            int bin = 0;
            if (valid) {
                const int Y = (int)(i / p.W), X = (int)(i % p.W);
                const int pr = FULLRES ? pred_fullres(img, p.C, plane_hi, i) : pred_bilinear(img, p.C, p.h, p.w, p.sy, p.sx, Y, X);
                bin = clamp_target_row(p.target[(size_t)b * npx + i], p.cm_rows) * p.C + pr;
            }
            cm_add_match(cm_dyn, bin, valid);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < nb; i += blockDim.x)
            if (cm_dyn[i]) atomicAdd(p.cm + (size_t)(b % p.cm_slots) * nb + i, (unsigned long long)cm_dyn[i]);
    }
    if (p.mask_f32 || p.mask_u8) {
        const size_t npx = (size_t)p.mask_h * p.mask_w;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (size_t)gridDim.x * blockDim.x) {
            const int Y = nearest_src((int)(i / p.mask_w), p.my, p.H);
            const int X = nearest_src((int)(i % p.mask_w), p.mx, p.W);
            const int pr = FULLRES ? pred_fullres(img, p.C, plane_hi, (size_t)Y * p.W + X)
                                   : pred_bilinear(img, p.C, p.h, p.w, p.sy, p.sx, Y, X);
            if (p.mask_f32) p.mask_f32[(size_t)b * npx + i] = (float)pr;
            if (p.mask_u8) p.mask_u8[(size_t)b * npx + i] = (uint8_t)pr;
        }
    }
}

// ---- full-resolution logits, vectorised: W*H % 4 == 0, mask (if any) at full resolution --------
// One thread = 4 consecutive pixels x kQuadsPerThread strided quads; C planes read with 16 B
// streaming loads (each logit is read exactly once).
constexpr int kQuadsPerThread = 8;
template <typename T, bool HAS_TGT, bool HAS_MASK>
__global__ void __launch_bounds__(256) seg_fullres_vec_kernel(const SegParams p) {
    __shared__ unsigned int cm_smem[DSN_MAX_CM_ROWS * 32];
    const int b = blockIdx.y;
    const size_t plane = (size_t)p.H * p.W;
    const size_t nquads = plane >> 2;
    const T *img = static_cast<const T *>(p.logits) + (size_t)b * p.C * plane;
    const int nb = p.cm_rows * p.C;
    if (HAS_TGT) {
        for (int i = threadIdx.x; i < nb; i += blockDim.x) cm_smem[i] = 0u;
        __syncthreads();
    }
    const size_t q0 = (size_t)blockIdx.x * blockDim.x * kQuadsPerThread;
    unsigned int hi_sink = 0;
#pragma unroll 2
    for (int j = 0; j < kQuadsPerThread; ++j) {
        const size_t q = q0 + (size_t)j * blockDim.x + threadIdx.x;
        const bool valid = q < nquads;
        int pred[4] = {0, 0, 0, 0};
        if (valid) {
            float4 best = Elem<T>::ld4(img + 4 * q);
            for (int c = 1; c < p.C; ++c) {
                const float4 v = Elem<T>::ld4(img + (size_t)c * plane + 4 * q);
                if (argmax_better(v.x, best.x)) { best.x = v.x; pred[0] = c; }
                if (argmax_better(v.y, best.y)) { best.y = v.y; pred[1] = c; }
                if (argmax_better(v.z, best.z)) { best.z = v.z; pred[2] = c; }
                if (argmax_better(v.w, best.w)) { best.w = v.w; pred[3] = c; }
            }
            const size_t pix = (size_t)b * plane + 4 * q;
            if (HAS_MASK) {
                if (p.mask_f32)
                    stg_stream_f4(p.mask_f32 + pix, make_float4((float)pred[0], (float)pred[1], (float)pred[2], (float)pred[3]));
                if (p.mask_u8)
                    *reinterpret_cast<uint32_t *>(p.mask_u8 + pix) =
                        (uint32_t)pred[0] | ((uint32_t)pred[1] << 8) | ((uint32_t)pred[2] << 16) | ((uint32_t)pred[3] << 24);
            }
        }
        if (HAS_TGT) {
            long long tv[4] = {0, 0, 0, 0};
            if (valid) {
                const size_t pix = (size_t)b * plane + 4 * q;
                const longlong2 ta = ldg_stream_ll2(p.target + pix);
                const longlong2 tb = ldg_stream_ll2(p.target + pix + 2);
                tv[0] = ta.x; tv[1] = ta.y; tv[2] = tb.x; tv[3] = tb.y;
                hi_sink ^= hi_words(ta, tb);              // keeps the loads 128 bits wide, see hi_words
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                cm_add_match(cm_smem, clamp_target_row(tv[k], p.cm_rows) * p.C + pred[k], valid);
        }
    }
    if (HAS_TGT) {
        if (hi_sink == 0x9E3779B9u && p.bs < 0) p.cm[0] = hi_sink;      // never true: see hi_words
        __syncthreads();
        for (int i = threadIdx.x; i < nb; i += blockDim.x)
            if (cm_smem[i]) atomicAdd(p.cm + (size_t)(b % p.cm_slots) * nb + i, (unsigned long long)cm_smem[i]);
    }
}

// ---- plain upsample (strict drop-in for nn.Upsample, yolo.py:183) ------------------------------
template <typename T>
__global__ void __launch_bounds__(256) upsample_bilinear_ac_kernel(const T *__restrict__ in, long long planes, int h, int w,
                                                                   int H, int W, float sy, float sx, float *__restrict__ out) {
    const size_t total = (size_t)planes * H * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int X = (int)(i % W);
        const size_t t = i / W;
        const int Y = (int)(t % H);
        const T *pl = in + (t / H) * (size_t)h * w;
        const Lerp ly = lerp_coeff(sy, Y, h);
        const Lerp lx = lerp_coeff(sx, X, w);
        const int y1 = ly.i0 + (ly.i0 < h - 1), x1 = lx.i0 + (lx.i0 < w - 1);
        const float tp = __fmaf_rn(lx.l0, Elem<T>::ld1c(pl + ly.i0 * w + lx.i0), __fmul_rn(lx.l1, Elem<T>::ld1c(pl + ly.i0 * w + x1)));
        const float bt = __fmaf_rn(lx.l0, Elem<T>::ld1c(pl + y1 * w + lx.i0), __fmul_rn(lx.l1, Elem<T>::ld1c(pl + y1 * w + x1)));
        out[i] = __fmaf_rn(ly.l0, tp, __fmul_rn(ly.l1, bt));
    }
}

static inline float ac_scale(int n_in, int n_out) { return n_out > 1 ? (float)(n_in - 1) / (float)(n_out - 1) : 0.0f; }

static int check_seg_args(const void *logits, int dtype, int bs, int C, int H, int W, const int64_t *target, float *mask_f32,
                          uint8_t *mask_u8, int mask_h, int mask_w, unsigned long long *cm, int cm_slots, int cm_rows) {
    if (!logits || bs < 0 || C < 1 || H < 1 || W < 1 || dtype_size(dtype) == 0) return DSN_E_INVALID;
    if (((uintptr_t)logits & (uintptr_t)(dtype_size(dtype) - 1)) != 0) return DSN_E_ALIGN;
    if ((target != nullptr) != (cm != nullptr)) return DSN_E_INVALID;
    if (target && (cm_slots < 1 || cm_rows < 1)) return DSN_E_INVALID;
    if ((mask_f32 || mask_u8) && (mask_h < 1 || mask_w < 1)) return DSN_E_INVALID;
    if (bs > 65535) return DSN_E_UNSUPPORTED;
    if ((long long)cm_rows * C > 12288) return DSN_E_UNSUPPORTED;   // 48 KB of shared-memory bins
    return DSN_OK;
}

template <typename T, bool FULLRES>
static int launch_general(const SegParams &p, cudaStream_t st) {
    const size_t px = std::max((size_t)p.H * p.W, (size_t)p.mask_h * p.mask_w);
    int gx = (int)std::min<size_t>((px + 256 * 16 - 1) / (256 * 16), 4096);
    gx = std::max(gx, 1);
    const size_t smem = p.target ? (size_t)p.cm_rows * p.C * sizeof(unsigned int) : 0;
    seg_general_kernel<T, FULLRES><<<dim3(gx, p.bs), 256, smem, st>>>(p);
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}

static int sm_count() { return sm_count_current_device(); }

// Persistent grid for the row kernels: SMs x resident CTAs, at least 8 output rows per CTA.
static inline void rows_grid(const SegParams &p, int ctas_per_sm, long long &grid, int &rows_per_cta) {
    const long long total_rows = (long long)p.bs * p.H;
    grid = std::min<long long>((long long)sm_count() * ctas_per_sm, (total_rows + 7) / 8);
    rows_per_cta = (int)((total_rows + grid - 1) / grid);
    grid = (total_rows + rows_per_cta - 1) / rows_per_cta;
}

// eval mode (target -> confusion matrix, optionally the mask as well)
template <typename T, int C, int M, typename W>
static int launch_eval_rows_w(const SegParams &p, cudaStream_t st) {
    const int G = p.W / 4;
    const int threads = std::min(512, (G + 31) / 32 * 32);
    const size_t smem = (sizeof(float2) + sizeof(int)) * (size_t)p.H + sizeof(int) * (size_t)(p.h + 1);
    static DeviceOnce once;                   // per template instantiation
    if (int rc = raise_dynamic_smem(once, seg_fused_rows_kernel<T, C, true, M, W>, 200 * 1024)) return rc;
    int occ = 0;
    DSN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, seg_fused_rows_kernel<T, C, true, M, W>, threads, smem));
    if (occ < 1) occ = 1;
    if (DSN_SEG_EVAL_CTAS_PER_SM > 0 && occ > DSN_SEG_EVAL_CTAS_PER_SM) occ = DSN_SEG_EVAL_CTAS_PER_SM;
    long long grid;
    int rows_per_cta;
    rows_grid(p, occ, grid, rows_per_cta);
    seg_fused_rows_kernel<T, C, true, M, W><<<(unsigned)grid, threads, smem, st>>>(p, rows_per_cta);
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}

// detect mode (mask only): blocks of kMaskRowsPerBlock rows dealt round-robin to a grid of kMaskCtasPerSm CTAs per SM
// (0: one CTA per block)
constexpr int kMaskCtasPerSm = DSN_SEG_MASK_CTAS_PER_SM;
constexpr int kMaskRowsPerBlock = DSN_SEG_MASK_RPB;
template <typename T, int C, int M, int PX>
static int launch_mask_rows(const SegParams &p, cudaStream_t st) {
    const int G = p.W / PX;
    const int threads = std::min(512, (G + 31) / 32 * 32);
    const int rpb = std::min(p.H, kMaskRowsPerBlock);
    const int bpi = (p.H + rpb - 1) / rpb;
    const long long nblocks = (long long)p.bs * bpi;
    long long grid = kMaskCtasPerSm > 0 ? std::min<long long>(nblocks, (long long)sm_count() * kMaskCtasPerSm) : nblocks;
    grid = std::min<long long>(grid, 0x7fffffffLL);
    seg_mask_rows_kernel<T, C, M, PX><<<(unsigned)grid, threads, 0, st>>>(p, rpb, bpi, nblocks);
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}
template <typename T, int C, int M>
static int launch_mask(const SegParams &p, cudaStream_t st) {
#if DSN_SEG_MASK_PX == 8
    // 8 pixels per thread when they cannot span more than two source intervals (7 * sx < 1) and the rows allow 32-byte groups
    const bool px8 = p.W % 8 == 0 && (double)(p.W - 1) > 7.0 * (double)(p.w - 1) + 1e-3 &&
                     ((M & 1) == 0 || (((uintptr_t)p.mask_f32 & 15) == 0)) && ((M & 2) == 0 || (((uintptr_t)p.mask_u8 & 3) == 0));
    if (px8) return launch_mask_rows<T, C, M, 8>(p, st);
#endif
    return launch_mask_rows<T, C, M, 4>(p, st);
}

template <typename T, int C>
static int launch_tile(const SegParams &p, cudaStream_t st) {
    const bool tgt = p.target != nullptr;
    const int mk = (p.mask_f32 ? 1 : 0) | (p.mask_u8 ? 2 : 0);
    if (tgt) {
        const bool small = p.cm_rows <= 4;      // 8-bit counters of 4 rows fit one 32-bit word
        switch (mk) {
            case 0: return small ? launch_eval_rows_w<T, C, 0, unsigned int>(p, st) : launch_eval_rows_w<T, C, 0, unsigned long long>(p, st);
            case 1: return small ? launch_eval_rows_w<T, C, 1, unsigned int>(p, st) : launch_eval_rows_w<T, C, 1, unsigned long long>(p, st);
            case 2: return small ? launch_eval_rows_w<T, C, 2, unsigned int>(p, st) : launch_eval_rows_w<T, C, 2, unsigned long long>(p, st);
            default: return small ? launch_eval_rows_w<T, C, 3, unsigned int>(p, st) : launch_eval_rows_w<T, C, 3, unsigned long long>(p, st);
        }
    }
    if (mk == 1) return launch_mask<T, C, 1>(p, st);
    if (mk == 2) return launch_mask<T, C, 2>(p, st);
    return launch_mask<T, C, 3>(p, st);
}

template <typename T>
static int seg_fused_dispatch(const SegParams &p, bool fast, cudaStream_t st) {
    if (!fast) return launch_general<T, false>(p, st);
    switch (p.C) {
        case 2: return launch_tile<T, 2>(p, st);
        case 3: return launch_tile<T, 3>(p, st);
        default: return launch_tile<T, 4>(p, st);
    }
}

template <typename T>
static int seg_fullres_dispatch(const SegParams &p, bool fast, bool has_mask, cudaStream_t st) {
    if (!fast) return launch_general<T, true>(p, st);
    const size_t nquads = (size_t)p.H * p.W / 4;
    dim3 grid((unsigned)((nquads + 256 * kQuadsPerThread - 1) / (256 * kQuadsPerThread)), p.bs);
    const bool tgt = p.target != nullptr;
    if (tgt && has_mask) seg_fullres_vec_kernel<T, true, true><<<grid, 256, 0, st>>>(p);
    else if (tgt) seg_fullres_vec_kernel<T, true, false><<<grid, 256, 0, st>>>(p);
    else seg_fullres_vec_kernel<T, false, true><<<grid, 256, 0, st>>>(p);
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}

}  // namespace dsn

using namespace dsn;

extern "C" __attribute__((visibility("default"))) int dsn_seg_fused(const void *logits, int dtype, int bs, int C, int h, int w, int H, int W,
                             const int64_t *target, float *mask_f32, uint8_t *mask_u8, int mask_h, int mask_w,
                             unsigned long long *cm, int cm_slots, int cm_rows, void *stream) {
    int rc = check_seg_args(logits, dtype, bs, C, H, W, target, mask_f32, mask_u8, mask_h, mask_w, cm, cm_slots, cm_rows);
    if (rc) return rc;
    if (h < 1 || w < 1) return DSN_E_INVALID;
    const bool has_mask = mask_f32 || mask_u8;
    if (bs == 0 || (!target && !has_mask)) return DSN_OK;
    SegParams p{};
    p.logits = logits; p.bs = bs; p.C = C; p.h = h; p.w = w; p.H = H; p.W = W;
    p.target = (const long long *)target; p.mask_f32 = mask_f32; p.mask_u8 = mask_u8;
    p.mask_h = has_mask ? mask_h : 0; p.mask_w = has_mask ? mask_w : 0;
    p.cm = cm; p.cm_slots = cm_slots; p.cm_rows = cm_rows;
    p.sy = ac_scale(h, H); p.sx = ac_scale(w, W);
    p.my = has_mask ? (float)H / (float)mask_h : 1.f;
    p.mx = has_mask ? (float)W / (float)mask_w : 1.f;
    cudaStream_t st = (cudaStream_t)stream;
    const bool fast = C >= 2 && C <= kMaxPackedC && (W % 4 == 0) && H >= h && H <= 4096 && (W - 1) >= 4 * (w - 1) && h * w < (1 << 24) &&
                      (!has_mask || (mask_h == H && mask_w == W)) && (!target || cm_rows <= DSN_MAX_CM_ROWS) &&
                      (((uintptr_t)mask_f32 & 15) == 0) && (((uintptr_t)mask_u8 & 3) == 0) && (((uintptr_t)target & 15) == 0);
    DSN_DISPATCH_DTYPE(dtype, T, { return seg_fused_dispatch<T>(p, fast, st); })
    return DSN_OK;
}

extern "C" __attribute__((visibility("default"))) int dsn_seg_fullres(const void *logits, int dtype, int bs, int C, int H, int W, const int64_t *target,
                               float *mask_f32, uint8_t *mask_u8, int mask_h, int mask_w, unsigned long long *cm, int cm_slots,
                               int cm_rows, void *stream) {
    int rc = check_seg_args(logits, dtype, bs, C, H, W, target, mask_f32, mask_u8, mask_h, mask_w, cm, cm_slots, cm_rows);
    if (rc) return rc;
    const bool has_mask = mask_f32 || mask_u8;
    if (bs == 0 || (!target && !has_mask)) return DSN_OK;
    SegParams p{};
    p.logits = logits; p.bs = bs; p.C = C; p.h = H; p.w = W; p.H = H; p.W = W;
    p.target = (const long long *)target; p.mask_f32 = mask_f32; p.mask_u8 = mask_u8;
    p.mask_h = has_mask ? mask_h : 0; p.mask_w = has_mask ? mask_w : 0;
    p.cm = cm; p.cm_slots = cm_slots; p.cm_rows = cm_rows;
    p.sy = 1.f; p.sx = 1.f;
    p.my = has_mask ? (float)H / (float)mask_h : 1.f;
    p.mx = has_mask ? (float)W / (float)mask_w : 1.f;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t plane = (size_t)H * W;
    const uintptr_t vmask = (uintptr_t)(4 * dtype_size(dtype) - 1);        // 4-element loads
    const bool fast = (plane % 4 == 0) && (!has_mask || (mask_h == H && mask_w == W)) &&
                      (!target || cm_rows * C <= DSN_MAX_CM_ROWS * 32) && (((uintptr_t)logits & vmask) == 0) &&
                      (((uintptr_t)mask_f32 & 15) == 0) && (((uintptr_t)mask_u8 & 3) == 0) && (((uintptr_t)target & 15) == 0);
    DSN_DISPATCH_DTYPE(dtype, T, { return seg_fullres_dispatch<T>(p, fast, has_mask, st); })
    return DSN_OK;
}

extern "C" __attribute__((visibility("default"))) int dsn_upsample_bilinear_ac(const void *in, int dtype, int64_t planes, int h, int w, int H, int W,
                                        float *out, void *stream) {
    if (!in || !out || planes < 0 || h < 1 || w < 1 || H < 1 || W < 1 || dtype_size(dtype) == 0) return DSN_E_INVALID;
    if (planes == 0) return DSN_OK;
    const size_t total = (size_t)planes * H * W;
    const int grid = (int)std::min<size_t>((total + 256 * 8 - 1) / (256 * 8), 148 * 32);
    DSN_DISPATCH_DTYPE(dtype, T, {
        upsample_bilinear_ac_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(static_cast<const T *>(in), planes, h, w, H, W, ac_scale(h, H),
                                                                             ac_scale(w, W), out);
    })
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}
