// Fused detection kernel: decode + confidence filter + per-image NMS in ONE launch, one CTA per
// image (core/models/yolo.py:259-277 + core/utils/general.py:659-750 of the reference).
//
// The two-kernel path (dsn_filter_levels -> dsn_nms_levels) pays a kernel boundary, a global
// candidate counter and several dependent HBM round trips (~2 us each on B200) per CTA on each
// side.  At the real-time setting (conf 0.25: ~2.6 % of the anchors pass the objectness test, ~130
// candidates per image) that chain is pure latency, so here one CTA owns an image end to end and
// the number of dependent round trips is cut to three:
//   1. objectness stage: the image's objectness planes are contiguous runs in the NCHW head
//      outputs, so ONE thread issues a handful of TMA bulk copies (cp.async.bulk global->shared,
//      completion on an mbarrier) that bring up to 100 KB of objectness logits into shared memory
//      in a single round trip with no register staging; the CTA scans them from shared memory
//      against a conservative LOGIT bound (no sigmoid) and appends the surviving rows to a shared
//      row list (warp-aggregated shared atomics);
//   2. class pass over the list (one row per lane): objectness, class AND box logits of a row are
//      loaded together (one round trip), the exact `sigmoid(obj) > conf` test is applied, conf =
//      obj*cls, best-class / multi-label selection; keys go to the image's key buffer with a
//      CTA-local counter (no global atomics) and, while there are few, keys + decoded boxes are also
//      stashed in shared memory;
//   3. the lazy NMS of nms_core.cuh: straight from the shared-memory stash when the image has
//      <= 256 candidates (no key gather, no box fetch), otherwise from the key buffer.
// The row list is flushed through step 2 whenever it could overflow, so dense inputs (conf 0.001:
// nearly every anchor passes) stream through the same code.
#include "filter_core.cuh"
#include "nms_core.cuh"

namespace dsn {

constexpr int kStageFloats = 14336;      // objectness logits staged per round (56 KB)
constexpr int kListCap = kStageFloats;   // rows in the shared list: a whole stage can pass, so the list never overflows
constexpr int kMaxPieces = 24;           // bulk copies per round

struct FusedSmem {
    NmsSmem nms;
    alignas(16) float stage[kStageFloats];       // kStageFloats ELEMENTS of the input type (fp16 / bf16 use half of it)
    unsigned int list[kListCap];
    alignas(8) unsigned long long mbar;
    int piece_off[kMaxPieces + 1];       // first stage float of the piece
    int piece_row[kMaxPieces];           // image row of the piece's first float
    int npieces;
    // resume point of the plane walk (thread 0 only)
    int next_l, next_a, next_c;
    unsigned int list_n;
    int ncand;
};

static_assert(sizeof(FusedSmem) <= 227 * 1024, "FusedSmem must fit the 227 KB opt-in shared memory of an sm_100a CTA");

// ---- mbarrier + TMA bulk copy (cp.async.bulk: 1-D, no tensor map needed) ----------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "DSN_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DSN_DONE;\n"
        "bra DSN_WAIT;\n"
        "DSN_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// key -> global key buffer; while the image has few candidates the key is also stashed in shared memory, and the
// four box-logit sectors of an emitting row are started towards L2 so that the NMS's box decode hits there
template <typename T>
struct KeySinkStash {
    unsigned long long *kout;
    long long cap;
    NmsSmem *nms;
    const T *box0;         // box-logit 0 of this lane's row
    size_t plane;
    __device__ __forceinline__ void operator()(int slot, unsigned long long key, int /*cls*/) const {
        if (slot < cap) kout[slot] = key;
        if (slot < kRankSortMax) nms->key[slot] = key;
#pragma unroll
        for (int o = 0; o < 4; ++o) asm volatile("prefetch.global.L2 [%0];" ::"l"(box0 + (size_t)o * plane));
    }
};

// raw class logits were loaded by the caller together with the objectness logit (one round trip)
template <typename T>
struct ClsPreloaded {
    const float *raw;      // kClsChunk values of the first chunk, already in registers
    const T *p;            // class-0 logit of this row (further chunks are loaded on demand)
    size_t plane;
    __device__ __forceinline__ float load(int c) const { return c < kClsChunk ? raw[c] : Elem<T>::ld1(p + (size_t)c * plane); }
    __device__ __forceinline__ float prob(float v) const { return sigmoidf_ref(v); }
};

// class pass over the rows listed in shared memory                         general.py:687-718
template <typename T>
__device__ __forceinline__ void fused_flush_list(FusedSmem &s, const HeadGeom &g, const FilterOut &f, const NmsParams &p, int b) {
    const int n = (int)s.list_n;
    const int nc = g.no - 5;
    unsigned long long *kout = f.keys + (size_t)b * f.cap;
    for (int i0 = 0; i0 < n; i0 += kNmsThreads) {
        const int i = i0 + threadIdx.x;
        const bool valid = i < n;
        const unsigned int row = valid ? s.list[i] : 0u;
        const RowGeom rg = row_geom(g, (int)row);
        const int plane = rg.plane;
        const T *base = static_cast<const T *>(g.raw[rg.l]) + ((size_t)(b * g.na + rg.a) * g.no) * plane + rg.cell;
        // the list only holds rows whose objectness logit cleared the conservative bound, so nearly all of them pass
        // the exact test: load the objectness logit AND the first chunk of class logits in the same round trip
        float ol = -INFINITY, raw[kClsChunk];
#pragma unroll
        for (int ci = 0; ci < kClsChunk; ++ci) raw[ci] = 0.f;
        if (valid) {
            ol = Elem<T>::ld1(base + (size_t)4 * plane);
#pragma unroll
            for (int ci = 0; ci < kClsChunk; ++ci)
                if (ci < nc) raw[ci] = Elem<T>::ld1(base + (size_t)(5 + ci) * plane);
        }
        const float obj = valid ? sigmoidf_ref(ol) : 0.f;                         // yolo.py:266
        KeySinkStash<T> sink{kout, f.cap, &s.nms, base, (size_t)plane};
        ClsPreloaded<T> cls{raw, base + (size_t)5 * plane, (size_t)plane};
        filter_emit(f, row, nc, obj, valid, cls, &s.ncand, sink);
    }
}

// append the rows of stage floats [lo, hi) whose logit clears the bound to the shared list
template <typename T>
__device__ __forceinline__ void fused_scan_range(FusedSmem &s, int lo, int hi, float obj_logit_lo) {
    const T *stage = reinterpret_cast<const T *>(s.stage);
    const int lane = threadIdx.x & 31;
    for (int o0 = lo; o0 < hi; o0 += kNmsThreads * 4) {
        const int o = o0 + threadIdx.x * 4;
        unsigned int pass = 0;
        if (o < hi) {
            const float4 v = Elem<T>::lds4(stage + o);
            pass = (v.x >= obj_logit_lo ? 1u : 0u) | (v.y >= obj_logit_lo ? 2u : 0u) | (v.z >= obj_logit_lo ? 4u : 0u) |
                   (v.w >= obj_logit_lo ? 8u : 0u);
        }
        if (!__any_sync(0xffffffffu, pass != 0)) continue;
        const int mine = __popc(pass);
        int incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        const int tot = __shfl_sync(0xffffffffu, incl, 31);
        unsigned int base = 0;
        if (lane == 31) base = atomicAdd(&s.list_n, (unsigned)tot);
        base = __shfl_sync(0xffffffffu, base, 31);
        unsigned int slot = base + incl - mine;
        if (pass) {
            int pc = 0;
            while (pc + 1 < s.npieces && o >= s.piece_off[pc + 1]) ++pc;
            const unsigned int row = (unsigned)(s.piece_row[pc] + (o - s.piece_off[pc]));
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if ((pass >> k) & 1u) s.list[slot++] = row + k;
        }
    }
}

// only_flagged: the on-device redo of the confidence cut (dsn_detect_nms_redo_levels) - only images whose NMS reported
// det_count == -2 are processed, every other CTA leaves at once.
template <typename T>
__global__ void __launch_bounds__(kNmsThreads) detect_fused_kernel(const HeadGeom g, const FilterOut f, const NmsParams p,
                                                                   const float obj_logit_lo, const int only_flagged) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FusedSmem &s = *reinterpret_cast<FusedSmem *>(smem_raw);
    const int b = blockIdx.x;
    if (only_flagged && p.det_count[b] != -2) return;
    DSN_MARK(0);
    if (threadIdx.x == 0) {
        s.list_n = 0u;
        s.ncand = 0;
        s.next_l = 0; s.next_a = 0; s.next_c = 0;
        mbar_init(&s.mbar, 1);
    }
    __syncthreads();
    uint32_t parity = 0;
    DSN_LAP_START();
    for (int round = 0;; ++round) {
        // ---- 1a. one thread walks the (level, anchor) objectness planes and issues the bulk copies ----
        if (threadIdx.x == 0) {
            int l = s.next_l, a = s.next_a, c = s.next_c, used = 0, np = 0;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // stage was read through the generic proxy
            uint32_t bytes = 0;
            // plan first (the expect_tx must be posted before or with the copies)
            int pl_[kMaxPieces], pa_[kMaxPieces], pc_[kMaxPieces], pn_[kMaxPieces];
            while (l < g.nl && np < kMaxPieces && used < kStageFloats) {
                const int plane = g.plane[l];
                const int take = min(plane - c, kStageFloats - used);
                pl_[np] = l; pa_[np] = a; pc_[np] = c; pn_[np] = take;
                s.piece_off[np] = used;
                s.piece_row[np] = g.row_off[l] + a * plane + c;
                used += take;
                bytes += (uint32_t)take * (uint32_t)sizeof(T);
                ++np;
                c += take;
                if (c == plane) { c = 0; if (++a == g.na) { a = 0; ++l; } }
            }
            s.piece_off[np] = used;
            s.npieces = np;
            s.next_l = l; s.next_a = a; s.next_c = c;
            if (np > 0) {
                mbar_expect_tx(&s.mbar, bytes);
                for (int i = 0; i < np; ++i) {
                    const int plane = g.plane[pl_[i]];
                    const T *src = static_cast<const T *>(g.raw[pl_[i]]) + ((size_t)(b * g.na + pa_[i]) * g.no + 4) * plane + pc_[i];
                    bulk_g2s(reinterpret_cast<T *>(s.stage) + s.piece_off[i], src, (uint32_t)pn_[i] * (uint32_t)sizeof(T), &s.mbar);
                }
            }
        }
        __syncthreads();
        const int np = s.npieces;
        if (np == 0) break;                                   // every plane has been scanned
        const int used = s.piece_off[np];
        mbar_wait(&s.mbar, parity);
        parity ^= 1u;
        DSN_LAP(42);   // plan + TMA stage wait
        DSN_MARK(1 + (round < 3 ? round : 3));
        // ---- 1b. scan from shared memory (the list can hold a whole stage, so one pass suffices) --------------------
        fused_scan_range<T>(s, 0, used, obj_logit_lo);
        __syncthreads();
        DSN_LAP(43);   // scan
        // ---- 2. class pass over the rows of this stage ---------------------------------------------------------------
        fused_flush_list<T>(s, g, f, p, b);
        __syncthreads();
        DSN_LAP(44);   // class pass
        if (threadIdx.x == 0) s.list_n = 0u;
        // (the loop head's __syncthreads orders this reset and the stage reuse before the next round)
    }
    DSN_MARK(8);
    const int total = s.ncand;
    if (threadIdx.x == 0) f.counts[b] = total;
    if (total > p.cap) {
        if (threadIdx.x == 0) p.det_count[b] = -1;
        return;
    }
    if (total <= 0 || p.max_det <= 0 || p.max_nms <= 0) {
        if (threadIdx.x == 0) p.det_count[b] = 0;
        return;
    }
    // ---- 3. NMS ------------------------------------------------------------------------------------------
    BoxFromLevels<T> src;
    src.g = g;
    nms_image(s.nms, p, src, b, p.keys + (size_t)b * p.cap, total, /*stashed_keys=*/total <= kRankSortMax);
}

}  // namespace dsn

using namespace dsn;

static int detect_nms_levels_impl(
    const void *const *raw, int dtype, const int *ny, const int *nx, const float *stride, const float *anchor_grid, int nl, int bs,
    int na, int no, float conf_thres, int multi_label, const uint64_t *class_mask, double iou_thres, int agnostic,
    int max_nms, int max_det, float max_wh, uint64_t *keys, int64_t cap, int *counts, float *det, int *det_count,
    void *stream, int only_flagged) {
    if (!stride || !anchor_grid) return DSN_E_INVALID;
    HeadGeom g;
    int rc = make_geom(g, raw, dtype, ny, nx, stride, anchor_grid, nl, bs, na, no);
    if (rc) return rc;
    FilterOut f;
    rc = make_filter_out(f, conf_thres, multi_label, no - 5, class_mask, keys, cap, counts);
    if (rc) return rc;
    NmsParams p;
    rc = make_nms_params(p, keys, cap, counts, bs, no - 5, conf_thres, iou_thres, agnostic, max_nms, max_det, max_wh, det,
                         det_count);
    if (rc) return rc;
    if (bs == 0) return DSN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // the TMA bulk copies need 16 B aligned, 16 B granular planes; otherwise use the two-kernel path
    const bool vec = geom_vec4(g) && geom_bulk(g);
    if (!vec) {
        if (only_flagged) return DSN_OK;       // no confidence cut exists for such layouts: nothing can be flagged
        rc = dsn_filter_levels(raw, dtype, ny, nx, nl, bs, na, no, conf_thres, multi_label, class_mask, keys, cap, counts, stream);
        if (rc) return rc;
        return dsn_nms_levels(raw, dtype, ny, nx, stride, anchor_grid, nl, bs, na, no, keys, cap, counts, conf_thres, iou_thres,
                              agnostic, max_nms, max_det, max_wh, nullptr, nullptr, det, det_count, stream);
    }
    const float lo = obj_logit_lower_bound(conf_thres);
    const size_t smem = sizeof(FusedSmem);
    DSN_DISPATCH_DTYPE(dtype, T, {
        static DeviceOnce once;               // per element type
        rc = raise_dynamic_smem(once, detect_fused_kernel<T>, (int)smem);
        if (rc) return rc;
        detect_fused_kernel<T><<<bs, kNmsThreads, smem, st>>>(g, f, p, lo, only_flagged);
    })
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}

extern "C" __attribute__((visibility("default"))) int dsn_detect_nms_levels(
    const void *const *raw, int dtype, const int *ny, const int *nx, const float *stride, const float *anchor_grid, int nl, int bs,
    int na, int no, float conf_thres, int multi_label, const uint64_t *class_mask, double iou_thres, int agnostic,
    int max_nms, int max_det, float max_wh, uint64_t *keys, int64_t cap, int *counts, float *det, int *det_count,
    void *stream) {
    return detect_nms_levels_impl(raw, dtype, ny, nx, stride, anchor_grid, nl, bs, na, no, conf_thres, multi_label, class_mask, iou_thres,
                                  agnostic, max_nms, max_det, max_wh, keys, cap, counts, det, det_count, stream, 0);
}

extern "C" __attribute__((visibility("default"))) int dsn_detect_nms_redo_levels(
    const void *const *raw, int dtype, const int *ny, const int *nx, const float *stride, const float *anchor_grid, int nl, int bs,
    int na, int no, float conf_thres, int multi_label, const uint64_t *class_mask, double iou_thres, int agnostic,
    int max_nms, int max_det, float max_wh, uint64_t *keys, int64_t cap, int *counts, float *det, int *det_count,
    void *stream) {
    return detect_nms_levels_impl(raw, dtype, ny, nx, stride, anchor_grid, nl, bs, na, no, conf_thres, multi_label, class_mask, iou_thres,
                                  agnostic, max_nms, max_det, max_wh, keys, cap, counts, det, det_count, stream, 1);
}
