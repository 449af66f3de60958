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
    const float *logits;
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

template <int C>
__device__ __forceinline__ SrcCols<C> load_src_row(const float *img, int plane, int row_off, int c0, int c1, int c2) {
    SrcCols<C> r;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float *pl = img + (size_t)c * plane + row_off;
        r.v[c][0] = __ldg(pl + c0);
        r.v[c][1] = __ldg(pl + c1);
        r.v[c][2] = __ldg(pl + c2);
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
// RING (eval mode without a mask): the int64 target does not pass through registers on its way in.  Every thread
// streams its own 32 B of each row into a private shared-memory ring with cp.async (LDGSTS, 16 B, L2 only), kRingAhead
// rows ahead of the row it is working on, and reads it back with two LDS.128 when the row's predictions are done.
// Registers bound the bytes a thread can keep in flight with plain loads, and those were consumed right after they
// were issued (nothing in flight during the arithmetic); with the ring the next rows' bytes are in flight all the time.
constexpr int kRingSlots = 8;      // rows per thread in the ring (power of two)
constexpr int kRingAhead = 4;      // rows in flight beyond the group being consumed
constexpr int kRingPrefetch = 12;  // rows ahead of the issue pointer that are prefetched into L2
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int C, bool HAS_TGT, int MASK, typename W, bool RING = false>
__global__ void __maxnreg__(HAS_TGT ? 96 : 128) seg_fused_rows_kernel(const SegParams p, const int rows_per_cta) {
    extern __shared__ __align__(16) unsigned char seg_smem[];
    float2 *lam = reinterpret_cast<float2 *>(seg_smem);                    // [H]   (l0, l1) of output row Y
    int *y0_of = reinterpret_cast<int *>(lam + p.H);                       // [H]   source row y0 of output row Y
    int *seg_start = y0_of + p.H;                                          // [h+1] first output row whose y0 is the index
    // [RING] target ring: slot s, half q of thread t at ((s * 2 + q) * blockDim.x + t) * 16 bytes (conflict-free LDS.128)
    unsigned char *ring = seg_smem + (((sizeof(float2) + sizeof(int)) * (size_t)p.H + sizeof(int) * (size_t)(p.h + 1) + 15) & ~(size_t)15);
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
        // [RING] stream state: next row to issue, rows of the run still to issue, issue / consume counters
        const long long *tp_iss = tp;
        long long iss_left = r_end - r_begin;
        unsigned int iss = 0, con = 0;
        auto ring_issue = [&]() {
            // the row kRingPrefetch rows further down is pulled into L2 (one prefetch per 128 B line = every fourth lane),
            // so that the cp.async of a row finds it there
            if (iss_left > kRingPrefetch && active && (threadIdx.x & 3) == 0)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(tp_iss + (size_t)kRingPrefetch * p.W));
            if (iss_left > 0 && active) {
                unsigned char *d = ring + ((size_t)((iss & (kRingSlots - 1)) * 2) * blockDim.x + threadIdx.x) * 16;
                cp_async16(d, tp_iss);
                cp_async16(d + (size_t)blockDim.x * 16, tp_iss + 2);
            }
            cp_async_commit();
            ++iss;
            tp_iss += p.W;
            --iss_left;
        };
        auto ring_read = [&](unsigned int slot, long long (&tv)[4]) {
            const unsigned char *d = ring + ((size_t)((slot & (kRingSlots - 1)) * 2) * blockDim.x + threadIdx.x) * 16;
            const longlong2 ta = *reinterpret_cast<const longlong2 *>(d);
            const longlong2 tb = *reinterpret_cast<const longlong2 *>(d + (size_t)blockDim.x * 16);
            tv[0] = ta.x; tv[1] = ta.y; tv[2] = tb.x; tv[3] = tb.y;
        };
        if constexpr (RING) {
#pragma unroll
            for (int q = 0; q < kRingAhead; ++q) ring_issue();
        }

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
                if constexpr (RING) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) ring_issue();
                    cp_async_wait<kRingAhead>();
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        long long t4[4] = {0, 0, 0, 0};
                        if (active) ring_read(con + q, t4);
                        row_body(lam[Y + q], top, bot, t4);
                    }
                    con += 4;
                    since_flush += 4;
                    if (since_flush >= 60) {
                        acc.flush(cm_smem, p.cm_rows);
                        since_flush = 0;
                    }
                    continue;
                }
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
                if constexpr (RING) {
                    ring_issue();
                    cp_async_wait<kRingAhead>();
                    if (active) ring_read(con, tv);
                    ++con;
                    ++since_flush;
                } else if (HAS_TGT) {
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
            const float *img = p.logits + (size_t)b * C * plane;
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
        if constexpr (RING) cp_async_wait<0>();
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

// ---- eval mode with a TMA-fed target ring -------------------------------------------------------------
// The int64 target is 8 B per pixel — 98 % of the bytes of an eval step — and a CTA's run of output rows is
// one contiguous range of it.  Instead of per-thread loads (registers bound the bytes in flight and the kernel
// sat at ~3.8 TB/s on long-scoreboard stalls), a dedicated PRODUCER warp streams the run through a ring of
// shared-memory stages with TMA bulk copies (cp.async.bulk + mbarrier complete_tx); the CONSUMER warps wait on
// a stage's "full" barrier, read their pixels' targets with LDS.128 and release the stage through its "empty"
// barrier.  The ring keeps ~56 KB per CTA in flight with zero register cost; everything else (segment walk,
// packed confusion-matrix counters) is as in seg_fused_rows_kernel.
__device__ __forceinline__ uint32_t seg_smem_u32(const void *ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void seg_mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(seg_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void seg_mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(seg_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void seg_mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(seg_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void seg_mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SEG_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SEG_DONE;\n"
        "bra SEG_WAIT;\n"
        "SEG_DONE:\n"
        "}\n" ::"r"(seg_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void seg_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(seg_smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(seg_smem_u32(bar))
                 : "memory");
}

constexpr int kMaxStages = 16;

template <int C, int MASK, typename W>
__global__ void __launch_bounds__(544) seg_eval_tma_kernel(const SegParams p, const int rows_per_cta, const int stage_rows,
                                                           const int nstage) {
    extern __shared__ __align__(128) unsigned char seg_smem[];
    // layout: ring | lam[H] | y0_of[H]
    const size_t stage_bytes = (size_t)stage_rows * p.W * sizeof(long long);
    long long *ring = reinterpret_cast<long long *>(seg_smem);
    float2 *lam = reinterpret_cast<float2 *>(seg_smem + stage_bytes * nstage);
    int *y0_of = reinterpret_cast<int *>(lam + p.H);
    __shared__ unsigned int cm_smem[DSN_MAX_CM_ROWS * kMaxPackedC];
    __shared__ __align__(8) unsigned long long full_bar[kMaxStages], empty_bar[kMaxStages];
    const int G = p.W >> 2;
    const int plane = p.h * p.w;
    const int n_cons = blockDim.x - 32;                       // consumer threads (multiple of 32, >= G)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool producer = warp == (n_cons >> 5);
    const long long total_rows = (long long)p.bs * p.H;
    const long long r_begin = (long long)blockIdx.x * rows_per_cta;
    const long long r_end = min(total_rows, r_begin + rows_per_cta);
    const int run_rows = (int)(r_end - r_begin);
    if (threadIdx.x < DSN_MAX_CM_ROWS * kMaxPackedC) cm_smem[threadIdx.x] = 0u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < nstage; ++i) {
            seg_mbar_init(&full_bar[i], 1);
            seg_mbar_init(&empty_bar[i], n_cons >> 5);        // one arrival per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int Y = threadIdx.x; Y < p.H; Y += blockDim.x) {
        const Lerp ly = lerp_coeff(p.sy, Y, p.h);
        lam[Y] = make_float2(ly.l0, ly.l1);
        y0_of[Y] = ly.i0;
    }
    __syncthreads();
    const int n_stages_total = (run_rows + stage_rows - 1) / stage_rows;

    if (producer) {
        if (lane == 0) {
            const long long *src = p.target + (size_t)r_begin * p.W;
            int st = 0;
            uint32_t ph = 1;                                  // parity of the PREVIOUS use of the slot
            for (int k = 0; k < n_stages_total; ++k) {
                if (k >= nstage) seg_mbar_wait(&empty_bar[st], ph);
                const int rows_here = min(stage_rows, run_rows - k * stage_rows);
                const uint32_t bytes = (uint32_t)((size_t)rows_here * p.W * sizeof(long long));
                seg_mbar_expect_tx(&full_bar[st], bytes);
                seg_bulk_g2s(reinterpret_cast<unsigned char *>(ring) + stage_bytes * st,
                             src + (size_t)k * stage_rows * p.W, bytes, &full_bar[st]);
                if (++st == nstage) { st = 0; ph ^= 1u; }
            }
        }
    } else {
        const int g = threadIdx.x;
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
                l0[k] = lx.l0; l1[k] = lx.l1; shx[k] = lx.i0 != xa;
            }
        }
        const int c0 = xa, c1 = min(xa + 1, p.w - 1), c2 = min(xa + 2, p.w - 1);
        auto hrow = [&](const SrcCols<C> &sc, float (&out)[C][4]) -> bool {
            bool fin = true;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                fin = fin && (fabsf(sc.v[c][0]) < INFINITY) && (fabsf(sc.v[c][1]) < INFINITY) && (fabsf(sc.v[c][2]) < INFINITY);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float a = shx[k] ? sc.v[c][1] : sc.v[c][0];
                    const float bb = shx[k] ? sc.v[c][2] : sc.v[c][1];
                    out[c][k] = __fmaf_rn(l0[k], a, __fmul_rn(l1[k], bb));
                }
            }
            return fin;
        };
        CmAcc<C, W> acc;
        acc.clear();
        float *m32 = (MASK & 1) ? p.mask_f32 + (size_t)r_begin * p.W + X : nullptr;
        uint8_t *m8 = (MASK & 2) ? p.mask_u8 + (size_t)r_begin * p.W + X : nullptr;
        bool finite = true;
        auto row_body = [&](const float2 lm, const float (&top)[C][4], const float (&bot)[C][4], const long long *trow) {
            int pred[4];
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
                }
            }
            if (active) {
                const longlong2 ta = *reinterpret_cast<const longlong2 *>(trow + X);
                const longlong2 tb = *reinterpret_cast<const longlong2 *>(trow + X + 2);
                const long long tv[4] = {ta.x, ta.y, tb.x, tb.y};
#pragma unroll
                for (int k = 0; k < 4; ++k) acc.add(pred[k], clamp_target_row(tv[k], p.cm_rows));
                if (MASK & 1) stg_stream_f4(m32, make_float4((float)pred[0], (float)pred[1], (float)pred[2], (float)pred[3]));
                if (MASK & 2)
                    *reinterpret_cast<uint32_t *>(m8) =
                        (uint32_t)pred[0] | ((uint32_t)pred[1] << 8) | ((uint32_t)pred[2] << 16) | ((uint32_t)pred[3] << 24);
            }
            if (MASK & 1) m32 += p.W;
            if (MASK & 2) m8 += p.W;
        };

        int b = (int)(r_begin / p.H);
        int Y = (int)(r_begin - (long long)b * p.H);
        const float *img = p.logits + (size_t)b * C * plane;
        float T[C][4], B[C][4];
        SrcCols<C> nxt;
        int ycur = 0, since_flush = 0;
        bool fresh = true, flip = false;
        int st = 0;                                           // ring slot of the current stage
        uint32_t phase = 0;                                   // parity of that slot's current fill
        for (int rel0 = 0; rel0 < run_rows; rel0 += stage_rows) {
            const int rows_here = min(stage_rows, run_rows - rel0);
            seg_mbar_wait(&full_bar[st], phase);
            const long long *trow = reinterpret_cast<const long long *>(reinterpret_cast<const unsigned char *>(ring) + stage_bytes * st);
            for (int rs = 0; rs < rows_here; ++rs, trow += p.W) {
                if (fresh) {                                  // first row of the run or of an image
                    fresh = false;
                    flip = false;
                    ycur = y0_of[Y];
                    const SrcCols<C> s0 = load_src_row<C>(img, plane, ycur * p.w, c0, c1, c2);
                    const SrcCols<C> s1 = load_src_row<C>(img, plane, min(ycur + 1, p.h - 1) * p.w, c0, c1, c2);
                    nxt = load_src_row<C>(img, plane, min(ycur + 2, p.h - 1) * p.w, c0, c1, c2);
                    const bool f0 = hrow(s0, T);
                    const bool f1 = hrow(s1, B);
                    finite = finite && f0 && f1;
                } else if (y0_of[Y] != ycur) {                // CTA-uniform: source row pair moved down by one
                    ycur = y0_of[Y];
                    const bool f2 = flip ? hrow(nxt, B) : hrow(nxt, T);
                    finite = finite && f2;
                    flip = !flip;
                    nxt = load_src_row<C>(img, plane, min(ycur + 2, p.h - 1) * p.w, c0, c1, c2);
                }
                if (!flip) row_body(lam[Y], T, B, trow); else row_body(lam[Y], B, T, trow);
                if (++Y == p.H) {
                    Y = 0;
                    img += (size_t)C * plane;
                    fresh = true;
                }
            }
            __syncwarp();                                     // done with this stage: hand it back to the producer
            if (lane == 0) seg_mbar_arrive(&empty_bar[st]);
            if (++st == nstage) { st = 0; phase ^= 1u; }
            since_flush += rows_here;
            if (since_flush >= 60) {                          // 4 px per row: an 8-bit field holds at most 255
                acc.flush(cm_smem, p.cm_rows);
                since_flush = 0;
            }
        }
        acc.flush(cm_smem, p.cm_rows);
    }
    __syncthreads();
    const int nb = p.cm_rows * C;
    if (threadIdx.x < nb) {
        const unsigned int cnt = cm_smem[threadIdx.x];
        if (cnt) atomicAdd(p.cm + (size_t)(blockIdx.x % p.cm_slots) * nb + threadIdx.x, (unsigned long long)cnt);
    }
}

// ---- general path: any C, any scale, any mask size, any cm_rows -------------------------------
__device__ __forceinline__ int pred_bilinear(const float *img, int C, int h, int w, float sy, float sx, int Y, int X) {
    const Lerp ly = lerp_coeff(sy, Y, h);
    const Lerp lx = lerp_coeff(sx, X, w);
    const int y1 = ly.i0 + (ly.i0 < h - 1), x1 = lx.i0 + (lx.i0 < w - 1);
    float best = 0.f;
    int bi = 0;
    for (int c = 0; c < C; ++c) {
        const float *pl = img + (size_t)c * h * w;
        const float t = __fmaf_rn(lx.l0, __ldg(pl + ly.i0 * w + lx.i0), __fmul_rn(lx.l1, __ldg(pl + ly.i0 * w + x1)));
        const float bt = __fmaf_rn(lx.l0, __ldg(pl + y1 * w + lx.i0), __fmul_rn(lx.l1, __ldg(pl + y1 * w + x1)));
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

__device__ __forceinline__ int pred_fullres(const float *img, int C, size_t plane, size_t off) {
    float best = __ldg(img + off);
    int bi = 0;
    for (int c = 1; c < C; ++c) {
        const float val = __ldg(img + (size_t)c * plane + off);
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

template <bool FULLRES>
__global__ void __launch_bounds__(256) seg_general_kernel(const SegParams p) {
    extern __shared__ unsigned int cm_dyn[];
    const int b = blockIdx.y;
    const int nb = p.cm_rows * p.C;
    const size_t plane_lo = (size_t)p.h * p.w;
    const size_t plane_hi = (size_t)p.H * p.W;
    const float *img = p.logits + (size_t)b * p.C * (FULLRES ? plane_hi : plane_lo);
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
template <bool HAS_TGT, bool HAS_MASK>
__global__ void __launch_bounds__(256) seg_fullres_vec_kernel(const SegParams p) {
    __shared__ unsigned int cm_smem[DSN_MAX_CM_ROWS * 32];
    const int b = blockIdx.y;
    const size_t plane = (size_t)p.H * p.W;
    const size_t nquads = plane >> 2;
    const float *img = p.logits + (size_t)b * p.C * plane;
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
            float4 best = ldg_stream_f4(img + 4 * q);
            for (int c = 1; c < p.C; ++c) {
                const float4 v = ldg_stream_f4(img + (size_t)c * plane + 4 * q);
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
__global__ void __launch_bounds__(256) upsample_bilinear_ac_kernel(const float *__restrict__ in, long long planes, int h, int w,
                                                                   int H, int W, float sy, float sx, float *__restrict__ out) {
    const size_t total = (size_t)planes * H * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int X = (int)(i % W);
        const size_t t = i / W;
        const int Y = (int)(t % H);
        const float *pl = in + (t / H) * (size_t)h * w;
        const Lerp ly = lerp_coeff(sy, Y, h);
        const Lerp lx = lerp_coeff(sx, X, w);
        const int y1 = ly.i0 + (ly.i0 < h - 1), x1 = lx.i0 + (lx.i0 < w - 1);
        const float tp = __fmaf_rn(lx.l0, __ldg(pl + ly.i0 * w + lx.i0), __fmul_rn(lx.l1, __ldg(pl + ly.i0 * w + x1)));
        const float bt = __fmaf_rn(lx.l0, __ldg(pl + y1 * w + lx.i0), __fmul_rn(lx.l1, __ldg(pl + y1 * w + x1)));
        out[i] = __fmaf_rn(ly.l0, tp, __fmul_rn(ly.l1, bt));
    }
}

static inline float ac_scale(int n_in, int n_out) { return n_out > 1 ? (float)(n_in - 1) / (float)(n_out - 1) : 0.0f; }

static int check_seg_args(const float *logits, int bs, int C, int H, int W, const int64_t *target, float *mask_f32,
                          uint8_t *mask_u8, int mask_h, int mask_w, unsigned long long *cm, int cm_slots, int cm_rows) {
    if (!logits || bs < 0 || C < 1 || H < 1 || W < 1) return DSN_E_INVALID;
    if ((target != nullptr) != (cm != nullptr)) return DSN_E_INVALID;
    if (target && (cm_slots < 1 || cm_rows < 1)) return DSN_E_INVALID;
    if ((mask_f32 || mask_u8) && (mask_h < 1 || mask_w < 1)) return DSN_E_INVALID;
    if (bs > 65535) return DSN_E_UNSUPPORTED;
    if ((long long)cm_rows * C > 12288) return DSN_E_UNSUPPORTED;   // 48 KB of shared-memory bins
    return DSN_OK;
}

template <bool FULLRES>
static int launch_general(const SegParams &p, cudaStream_t st) {
    const size_t px = std::max((size_t)p.H * p.W, (size_t)p.mask_h * p.mask_w);
    int gx = (int)std::min<size_t>((px + 256 * 16 - 1) / (256 * 16), 4096);
    gx = std::max(gx, 1);
    const size_t smem = p.target ? (size_t)p.cm_rows * p.C * sizeof(unsigned int) : 0;
    seg_general_kernel<FULLRES><<<dim3(gx, p.bs), 256, smem, st>>>(p);
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}

static inline int G_of(const SegParams &p) { return p.W / 4; }

static int sm_count() { return sm_count_current_device(); }

template <int C, bool T, int M, typename W, bool RING = false>
static int launch_rows_w(const SegParams &p, cudaStream_t st) {
    const int G = p.W / 4;
    const int threads = std::min(512, (G + 31) / 32 * 32);
    size_t smem = (sizeof(float2) + sizeof(int)) * (size_t)p.H + sizeof(int) * (size_t)(p.h + 1);
    if (RING) smem = ((smem + 15) & ~(size_t)15) + (size_t)kRingSlots * 32 * threads;
    static DeviceOnce once;                   // per template instantiation
    if (int rc = raise_dynamic_smem(once, seg_fused_rows_kernel<C, T, M, W, RING>, 200 * 1024)) return rc;
    int occ = 0;
    DSN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, seg_fused_rows_kernel<C, T, M, W, RING>, threads, smem));
    if (occ < 1) occ = 1;
    static int occ_cap = -1;                 // experiment knob: DSN_SEG_OCC=<resident CTAs per SM the persistent grid is sized for>
    if (occ_cap < 0) {
        const char *e = getenv("DSN_SEG_OCC");
        occ_cap = e ? atoi(e) : 0;
    }
    // detect mode is a pure store stream: two 160-thread CTAs per SM already saturate it (21.8 vs 22.1 us), and the
    // register file they leave free lets an NMS CTA of a batch in flight share the SM (step 29.85 -> 29.3 us at cfg 2)
    if (!T && occ_cap == 0 && occ > 2) occ = 2;
    if (occ_cap > 0 && occ > occ_cap) occ = occ_cap;
    const long long total_rows = (long long)p.bs * p.H;
    long long grid = std::min<long long>((long long)sm_count() * occ, (total_rows + 7) / 8);
    int rows_per_cta = (int)((total_rows + grid - 1) / grid);
    static int forced = -1;                  // experiment knob: DSN_SEG_ROWS=<rows per CTA>
    if (forced < 0) {
        const char *e = getenv("DSN_SEG_ROWS");
        forced = e ? atoi(e) : 0;
    }
    if (forced > 0) rows_per_cta = forced;
    grid = (total_rows + rows_per_cta - 1) / rows_per_cta;
    seg_fused_rows_kernel<C, T, M, W, RING><<<(unsigned)grid, threads, smem, st>>>(p, rows_per_cta);
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}

template <int C, int M, typename W>
static int launch_eval_tma_w(const SegParams &p, cudaStream_t st) {
    const int G = p.W / 4;
    const int cons = (G + 31) / 32 * 32;
    const int threads = cons + 32;
    const int stage_rows = 2;
    const size_t stage_bytes = (size_t)stage_rows * p.W * sizeof(long long);
    int nstage = (int)std::min<size_t>(kMaxStages, std::max<size_t>(3, (56 * 1024) / stage_bytes));
    const size_t smem = stage_bytes * nstage + (sizeof(float2) + sizeof(int)) * (size_t)p.H;
    static DeviceOnce once;                   // per template instantiation
    if (int rc = raise_dynamic_smem(once, seg_eval_tma_kernel<C, M, W>, 200 * 1024)) return rc;
    int occ = 0;
    DSN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, seg_eval_tma_kernel<C, M, W>, threads, smem));
    if (occ < 1) return DSN_E_UNSUPPORTED;
    const long long total_rows = (long long)p.bs * p.H;
    long long grid = std::min<long long>((long long)sm_count() * occ, (total_rows + 7) / 8);
    const int rows_per_cta = (int)((total_rows + grid - 1) / grid);
    grid = (total_rows + rows_per_cta - 1) / rows_per_cta;
    seg_eval_tma_kernel<C, M, W><<<(unsigned)grid, threads, smem, st>>>(p, rows_per_cta, stage_rows, nstage);
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}

template <int C, bool T, int M>
static int launch_rows_one(const SegParams &p, cudaStream_t st) {
    if constexpr (T && M == 0) {
        // pure eval, opt-in (DSN_SEG_RING=1): target streamed through a per-thread cp.async ring.  Measured on B200
        // (cfg 3, bs 512): the kernel ALONE is slower than the plain-load variant (463 vs 363 us: 4 rows ahead do not
        // cover a DRAM round trip and 8 slots x 32 B x threads of shared memory cap the ring depth), while the whole
        // step with 3 batches in flight is ~6 % faster (the ring CTAs share the SMs better with the detection chain);
        // at 1280 px it loses both ways.  The default stays the variant with the higher stand-alone bandwidth.
        static int ring = -1;
        if (ring < 0) { const char *e = getenv("DSN_SEG_RING"); ring = (e && atoi(e)) ? 1 : 0; }
        if (ring && (size_t)kRingSlots * 32 * std::min(512, (p.W / 4 + 31) / 32 * 32) <= 160 * 1024) {
            if (p.cm_rows <= 4) return launch_rows_w<C, T, M, unsigned int, true>(p, st);
            return launch_rows_w<C, T, M, unsigned long long, true>(p, st);
        }
    }
    if (!T || p.cm_rows <= 4) return launch_rows_w<C, T, M, unsigned int>(p, st);
    return launch_rows_w<C, T, M, unsigned long long>(p, st);
}
template <int C, int M>
static int launch_eval_tma(const SegParams &p, cudaStream_t st) {
    if (p.cm_rows <= 4) return launch_eval_tma_w<C, M, unsigned int>(p, st);
    return launch_eval_tma_w<C, M, unsigned long long>(p, st);
}

template <int C>
static int launch_tile(const SegParams &p, cudaStream_t st) {
    const bool tgt = p.target != nullptr;
    const int mk = (p.mask_f32 ? 1 : 0) | (p.mask_u8 ? 2 : 0);
    // eval mode: the TMA-fed target ring (seg_eval_tma_kernel) is parity-green but measured no faster than the
    // streaming-load variant on B200 (55.6 vs 55.3 us at bs 64, 640x640), so it is opt-in: DSN_SEG_TMA=1
    static int use_tma = -1;
    if (use_tma < 0) { const char *e = getenv("DSN_SEG_TMA"); use_tma = (e && atoi(e)) ? 1 : 0; }
    if (tgt && use_tma && G_of(p) <= 512 && (size_t)2 * p.W * 8 * 3 <= 160 * 1024) {
        if (mk == 0) return launch_eval_tma<C, 0>(p, st);
        if (mk == 1) return launch_eval_tma<C, 1>(p, st);
        if (mk == 2) return launch_eval_tma<C, 2>(p, st);
        return launch_eval_tma<C, 3>(p, st);
    }
    if (tgt) {
        if (mk == 0) return launch_rows_one<C, true, 0>(p, st);
        if (mk == 1) return launch_rows_one<C, true, 1>(p, st);
        if (mk == 2) return launch_rows_one<C, true, 2>(p, st);
        return launch_rows_one<C, true, 3>(p, st);
    }
    if (mk == 1) return launch_rows_one<C, false, 1>(p, st);
    if (mk == 2) return launch_rows_one<C, false, 2>(p, st);
    return launch_rows_one<C, false, 3>(p, st);
}

}  // namespace dsn

using namespace dsn;

extern "C" __attribute__((visibility("default"))) int dsn_seg_fused(const float *logits, int bs, int C, int h, int w, int H, int W, const int64_t *target,
                             float *mask_f32, uint8_t *mask_u8, int mask_h, int mask_w, unsigned long long *cm,
                             int cm_slots, int cm_rows, void *stream) {
    int rc = check_seg_args(logits, bs, C, H, W, target, mask_f32, mask_u8, mask_h, mask_w, cm, cm_slots, cm_rows);
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
    if (!fast) return launch_general<false>(p, st);
    switch (C) {
        case 2: return launch_tile<2>(p, st);
        case 3: return launch_tile<3>(p, st);
        default: return launch_tile<4>(p, st);
    }
}

extern "C" __attribute__((visibility("default"))) int dsn_seg_fullres(const float *logits, int bs, int C, int H, int W, const int64_t *target, float *mask_f32,
                               uint8_t *mask_u8, int mask_h, int mask_w, unsigned long long *cm, int cm_slots,
                               int cm_rows, void *stream) {
    int rc = check_seg_args(logits, bs, C, H, W, target, mask_f32, mask_u8, mask_h, mask_w, cm, cm_slots, cm_rows);
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
    const bool fast = (plane % 4 == 0) && (!has_mask || (mask_h == H && mask_w == W)) &&
                      (!target || cm_rows * C <= DSN_MAX_CM_ROWS * 32) && (((uintptr_t)logits & 15) == 0) &&
                      (((uintptr_t)mask_f32 & 15) == 0) && (((uintptr_t)mask_u8 & 3) == 0) && (((uintptr_t)target & 15) == 0);
    if (!fast) return launch_general<true>(p, st);
    const size_t nquads = plane / 4;
    dim3 grid((unsigned)((nquads + 256 * kQuadsPerThread - 1) / (256 * kQuadsPerThread)), bs);
    const bool tgt = target != nullptr;
    if (tgt && has_mask) seg_fullres_vec_kernel<true, true><<<grid, 256, 0, st>>>(p);
    else if (tgt) seg_fullres_vec_kernel<true, false><<<grid, 256, 0, st>>>(p);
    else seg_fullres_vec_kernel<false, true><<<grid, 256, 0, st>>>(p);
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}

extern "C" __attribute__((visibility("default"))) int dsn_upsample_bilinear_ac(const float *in, int64_t planes, int h, int w, int H, int W, float *out,
                                        void *stream) {
    if (!in || !out || planes < 0 || h < 1 || w < 1 || H < 1 || W < 1) return DSN_E_INVALID;
    if (planes == 0) return DSN_OK;
    const size_t total = (size_t)planes * H * W;
    const int grid = (int)std::min<size_t>((total + 256 * 8 - 1) / (256 * 8), 148 * 32);
    upsample_bilinear_ac_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, planes, h, w, H, W, ac_scale(h, H), ac_scale(w, W), out);
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}
