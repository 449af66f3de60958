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

constexpr int kStageFloats = 25600;      // objectness logits staged per round (100 KB)
constexpr int kListCap = 6144;           // rows in the shared list
constexpr int kMaxPieces = 24;           // bulk copies per round

struct FusedSmem {
    NmsSmem nms;
    alignas(16) float stage[kStageFloats];
    unsigned int list[kListCap];
    alignas(8) unsigned long long mbar;
    int piece_off[kMaxPieces + 1];       // first stage float of the piece
    int piece_row[kMaxPieces];           // image row of the piece's first float
    int npieces;
    // resume point of the plane walk (thread 0 only)
    int next_l, next_a, next_c;
    unsigned int list_n;
    unsigned int pass_total;
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

struct ClsFromRowPlanes {
    const float *p;        // class-0 logit of this row
    size_t plane;
    __device__ __forceinline__ float load(int c) const { return ldg_stream_f1(p + (size_t)c * plane); }
    __device__ __forceinline__ float prob(float raw) const { return sigmoidf_ref(raw); }
};

// key -> global key buffer; while the image has few candidates also key + decoded box -> shared stash
// (the box is decoded only for rows that actually emit a candidate)
struct KeySinkStash {
    unsigned long long *kout;
    long long cap;
    NmsSmem *nms;
    float bl[4];           // raw box logits of this lane's row
    int gx, gy;
    float stride, aw, ah, off_mul;
    __device__ __forceinline__ void operator()(int slot, unsigned long long key, int cls) const {
        if (slot < cap) kout[slot] = key;
        if (slot < kRankSortMax) {
            nms->key[slot] = key;
            nms->box[slot] = xywh2xyxy(decode_xy(bl[0], gx, stride), decode_xy(bl[1], gy, stride), decode_wh(bl[2], aw),
                                       decode_wh(bl[3], ah));                    // yolo.py:268-269, general.py:706
            nms->off[slot] = __fmul_rn((float)cls, off_mul);                     // general.py:732
        }
    }
};

// class pass over the rows listed in shared memory                         general.py:687-718
__device__ __forceinline__ void fused_flush_list(FusedSmem &s, const HeadGeom &g, const FilterOut &f, const NmsParams &p, int b) {
    const int n = (int)s.list_n;
    unsigned long long *kout = f.keys + (size_t)b * f.cap;
    for (int i0 = 0; i0 < n; i0 += kNmsThreads) {
        const int i = i0 + threadIdx.x;
        const bool valid = i < n;
        const unsigned int row = valid ? s.list[i] : 0u;
        const RowGeom rg = row_geom(g, (int)row);
        const int plane = rg.plane;
        const float *base = g.raw[rg.l] + ((size_t)(b * g.na + rg.a) * g.no) * plane + rg.cell;
        const int nx = g.nx[rg.l];
        const int gy = rg.cell / nx;
        KeySinkStash sink{kout, f.cap, &s.nms, {0.f, 0.f, 0.f, 0.f}, rg.cell - gy * nx, gy, g.stride[rg.l],
                          g.anchor_w[rg.l][rg.a], g.anchor_h[rg.l][rg.a], p.off_mul};
        float ol = -INFINITY;
        if (valid) {
            // objectness + box logits ride in the same round trip as the class logits filter_emit loads
            ol = ldg_stream_f1(base + (size_t)4 * plane);
#pragma unroll
            for (int o = 0; o < 4; ++o) sink.bl[o] = ldg_stream_f1(base + (size_t)o * plane);
        }
        const float obj = valid ? sigmoidf_ref(ol) : 0.f;                         // yolo.py:266
        ClsFromRowPlanes cls{base + (size_t)5 * plane, (size_t)plane};
        filter_emit(f, row, g.no - 5, obj, valid, cls, &s.ncand, sink);
    }
}

// append the rows of stage floats [lo, hi) whose logit clears the bound to the shared list
__device__ __forceinline__ void fused_scan_range(FusedSmem &s, int lo, int hi, float obj_logit_lo, bool count_only) {
    const int lane = threadIdx.x & 31;
    unsigned int cnt = 0;
    for (int o0 = lo; o0 < hi; o0 += kNmsThreads * 4) {
        const int o = o0 + threadIdx.x * 4;
        unsigned int pass = 0;
        if (o < hi) {
            const float4 v = *reinterpret_cast<const float4 *>(&s.stage[o]);
            pass = (v.x >= obj_logit_lo ? 1u : 0u) | (v.y >= obj_logit_lo ? 2u : 0u) | (v.z >= obj_logit_lo ? 4u : 0u) |
                   (v.w >= obj_logit_lo ? 8u : 0u);
        }
        if (count_only) {
            cnt += __popc(pass);
            continue;
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
    if (count_only) {
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0 && cnt) atomicAdd(&s.pass_total, cnt);
    }
}

__global__ void __launch_bounds__(kNmsThreads) detect_fused_kernel(const HeadGeom g, const FilterOut f, const NmsParams p,
                                                                   const float obj_logit_lo) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FusedSmem &s = *reinterpret_cast<FusedSmem *>(smem_raw);
    const int b = blockIdx.x;
    DSN_MARK(0);
    if (threadIdx.x == 0) {
        s.list_n = 0u;
        s.ncand = 0;
        s.next_l = 0; s.next_a = 0; s.next_c = 0;
        mbar_init(&s.mbar, 1);
    }
    __syncthreads();
    uint32_t parity = 0;
    for (int round = 0;; ++round) {
        // ---- 1a. one thread walks the (level, anchor) objectness planes and issues the bulk copies ----
        if (threadIdx.x == 0) {
            s.pass_total = 0u;
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
                bytes += (uint32_t)take * 4u;
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
                    const float *src = g.raw[pl_[i]] + ((size_t)(b * g.na + pa_[i]) * g.no + 4) * plane + pc_[i];
                    bulk_g2s(&s.stage[s.piece_off[i]], src, (uint32_t)pn_[i] * 4u, &s.mbar);
                }
            }
        }
        __syncthreads();
        const int np = s.npieces;
        if (np == 0) break;                                   // every plane has been scanned
        const int used = s.piece_off[np];
        mbar_wait(&s.mbar, parity);
        parity ^= 1u;
        DSN_MARK(1 + (round < 3 ? round : 3));
        // ---- 1b. scan from shared memory ------------------------------------------------------------------
        fused_scan_range(s, 0, used, obj_logit_lo, true);     // how many rows would be listed?
        __syncthreads();
        if (s.list_n + s.pass_total <= (unsigned)kListCap) {
            fused_scan_range(s, 0, used, obj_logit_lo, false);
            __syncthreads();
        } else {
            // dense input: scan in slices that cannot overflow the list, flushing in between
            constexpr int kSlice = kNmsThreads * 4;
            for (int lo = 0; lo < used; lo += kSlice) {
                if (s.list_n + kSlice > (unsigned)kListCap) {
                    __syncthreads();
                    fused_flush_list(s, g, f, p, b);
                    __syncthreads();
                    if (threadIdx.x == 0) s.list_n = 0u;
                    __syncthreads();
                }
                fused_scan_range(s, lo, min(used, lo + kSlice), obj_logit_lo, false);
                __syncthreads();
            }
        }
    }
    DSN_MARK(5);
    // ---- 2. class pass over what is left in the list ------------------------------------------------------
    fused_flush_list(s, g, f, p, b);
    __syncthreads();
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
    BoxFromLevels src{g};
    nms_image(s.nms, p, src, b, p.keys + (size_t)b * p.cap, total, /*stashed=*/total <= kRankSortMax);
}

}  // namespace dsn

using namespace dsn;

extern "C" __attribute__((visibility("default"))) int dsn_detect_nms_levels(
    const float *const *raw, const int *ny, const int *nx, const float *stride, const float *anchor_grid, int nl, int bs,
    int na, int no, float conf_thres, int multi_label, const uint64_t *class_mask, double iou_thres, int agnostic,
    int max_nms, int max_det, float max_wh, uint64_t *keys, int64_t cap, int *counts, float *det, int *det_count,
    void *stream) {
    if (!stride || !anchor_grid) return DSN_E_INVALID;
    HeadGeom g;
    int rc = make_geom(g, raw, ny, nx, stride, anchor_grid, nl, bs, na, no);
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
    bool vec = true;
    for (int l = 0; l < nl; ++l) vec = vec && (g.plane[l] % 4 == 0) && (((uintptr_t)raw[l] & 15) == 0);
    if (!vec) {
        rc = dsn_filter_levels(raw, ny, nx, nl, bs, na, no, conf_thres, multi_label, class_mask, keys, cap, counts, stream);
        if (rc) return rc;
        return dsn_nms_levels(raw, ny, nx, stride, anchor_grid, nl, bs, na, no, keys, cap, counts, conf_thres, iou_thres,
                              agnostic, max_nms, max_det, max_wh, det, det_count, stream);
    }
    const float lo = obj_logit_lower_bound(conf_thres);
    const size_t smem = sizeof(FusedSmem);
    static DeviceOnce once;
    rc = raise_dynamic_smem(once, detect_fused_kernel, (int)smem);
    if (rc) return rc;
    detect_fused_kernel<<<bs, kNmsThreads, smem, st>>>(g, f, p, lo);
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}
