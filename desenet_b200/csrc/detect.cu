// Detection head tail: Detect.forward's post-conv decode (core/models/yolo.py:259-277) and the
// candidate filter of non_max_suppression (core/utils/general.py:668-718), as
//   * dsn_detect_decode  — strict drop-in: materialises z (bs,A,no) and the permuted raw logits
//   * dsn_filter_levels  — fused: raw NCHW head outputs -> 64-bit candidate keys, nothing else
//   * dsn_filter_z       — the same filter on an already decoded prediction
// All three are HBM-bound streaming kernels: plane-major 16 B loads (threadIdx.x runs along the
// cell index, the contiguous dimension of every channel plane), no shared-memory reuse needed
// except the decode transposition.
#include <algorithm>
#include <cmath>

#include "common.cuh"
#include "filter_core.cuh"
#include "geom.cuh"

namespace dsn {

// =============================================================================================
// Strict decode: z + xperm                                                 yolo.py:259-277
// One CTA = one (image, level, anchor) x a run of `cells` cells; the no x cells input tile is read
// plane-major (coalesced along cells), decoded, transposed through shared memory and written as
// ONE contiguous run of cells*no floats of z (rows of one anchor plane are adjacent in z) and of
// xperm.
// =============================================================================================
__global__ void __launch_bounds__(256) detect_decode_kernel(const HeadGeom g, float *__restrict__ z, XpermPtrs xp, int cells) {
    extern __shared__ float sm[];
    const int no = g.no;
    float *sz = sm;                       // [cells][no] decoded
    float *sx = sm + (size_t)cells * no;  // [cells][no] raw
    // blockIdx.x enumerates (level, anchor, tile) for one image
    int t = blockIdx.x, l = 0;
    while (l + 1 < g.nl && t >= g.tile_off[l + 1]) ++l;
    t -= g.tile_off[l];
    const int plane = g.plane[l];
    const int tiles_per_plane = (plane + cells - 1) / cells;
    const int a = t / tiles_per_plane;
    const int c0 = (t % tiles_per_plane) * cells;
    const int n = min(cells, plane - c0);
    const int b = blockIdx.y;
    const float *src = g.raw[l] + ((size_t)(b * g.na + a) * no) * plane + c0;
    const float stride = g.stride[l], aw = g.anchor_w[l][a], ah = g.anchor_h[l][a];
    const int nx = g.nx[l];
    for (int i = threadIdx.x; i < no * n; i += blockDim.x) {
        const int o = i / n, c = i - o * n;
        const float v = ldg_stream_f1(src + (size_t)o * plane + c);
        float r;
        const int cell = c0 + c;
        if (o == 0) r = decode_xy(v, cell % nx, stride);                     // yolo.py:268
        else if (o == 1) r = decode_xy(v, cell / nx, stride);
        else if (o < 4) r = decode_wh(v, o == 2 ? aw : ah);                  // yolo.py:269
        else r = sigmoidf_ref(v);                                            // yolo.py:266
        sz[c * no + o] = r;
        sx[c * no + o] = v;
    }
    __syncthreads();
    const size_t row0 = (size_t)g.row_off[l] + (size_t)a * plane + c0;
    if (z) {
        float *dst = z + ((size_t)b * g.A + row0) * no;                        // yolo.py:275,277
        for (int i = threadIdx.x; i < no * n; i += blockDim.x) dst[i] = sz[i];
    }
    if (xp.p[l]) {
        float *dst = xp.p[l] + (((size_t)(b * g.na + a)) * plane + c0) * no;   // yolo.py:260
        for (int i = threadIdx.x; i < no * n; i += blockDim.x) dst[i] = sx[i];
    }
}

// ---- from the raw levels -----------------------------------------------------------------------
// One CTA = a run of kFilterCells cells of ONE (image, level, anchor) plane, so all plane geometry
// is CTA-uniform.  Phase 1: every thread issues all of its objectness loads (16 B, coalesced,
// kFilterUnroll in flight), evaluates the sigmoid only for logits above a conservative logit-space
// bound, and appends the cells that pass `sigmoid(obj) > conf_thres` to a shared-memory list
// (warp-ballot/scan compaction, cell order preserved inside a warp).  Phase 2: the list is
// processed one cell per lane — class logits loaded in chunks of 8 (one round trip), conf = obj*cls,
// best-class or multi-label selection, warp-aggregated emission of the 64-bit keys.  At conf 0.25
// (~2.6 % of the anchors pass) the kernel is two dependent HBM round trips per CTA and touches only
// the objectness plane plus the sectors of surviving cells; at conf 0.001 the list holds nearly
// every cell and phase 2 streams the class planes with coalesced 4 B loads.
constexpr int kFilterUnroll = 4;
constexpr int kFilterThreads = 256;

template <int V>
__global__ void __launch_bounds__(kFilterThreads) filter_levels_kernel(const HeadGeom g, const FilterOut f, const float obj_logit_lo) {
    __shared__ unsigned int list_cell[kFilterThreads * kFilterUnroll * V];
    __shared__ float list_obj[kFilterThreads * kFilterUnroll * V];
    __shared__ unsigned int list_n;
    const int b = blockIdx.y;
    int l = 0;
    while (l + 1 < g.nl && (int)blockIdx.x >= g.cta_off[l + 1]) ++l;
    const int cp = blockIdx.x - g.cta_off[l];
    const int a = cp / g.ctas_per_plane[l];
    const int tile = cp - a * g.ctas_per_plane[l];
    const int plane = g.plane[l];
    const int units = plane / V;
    const float *pbase = g.raw[l] + ((size_t)(b * g.na + a) * g.no) * plane;
    const unsigned int prow0 = (unsigned)(g.row_off[l] + a * plane);
    const int u0 = tile * (kFilterThreads * kFilterUnroll) + threadIdx.x;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) list_n = 0u;
    float objraw[kFilterUnroll][V];
#pragma unroll
    for (int j = 0; j < kFilterUnroll; ++j) {
        const int u = u0 + j * kFilterThreads;
#pragma unroll
        for (int k = 0; k < V; ++k) objraw[j][k] = -INFINITY;
        if (u < units) {
            if constexpr (V == 4) {
                const float4 t = ldg_stream_f4(pbase + (size_t)4 * plane + u * 4);
                objraw[j][0] = t.x; objraw[j][1] = t.y; objraw[j][2] = t.z; objraw[j][3] = t.w;
            } else {
                objraw[j][0] = ldg_stream_f1(pbase + (size_t)4 * plane + u);
            }
        }
    }
    __syncthreads();
    // ---- phase 1: objectness test + ordered compaction into the shared list ----------------------
#pragma unroll
    for (int j = 0; j < kFilterUnroll; ++j) {
        float obj[V];
        unsigned int pass = 0;
#pragma unroll
        for (int k = 0; k < V; ++k) {
            obj[k] = objraw[j][k] >= obj_logit_lo ? sigmoidf_ref(objraw[j][k]) : 0.f;    // yolo.py:266
            pass |= obj[k] > f.conf_thres ? (1u << k) : 0u;                              // general.py:668
        }
        if (!__any_sync(0xffffffffu, pass != 0)) continue;
        const int mine = __popc(pass);
        int incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        unsigned int base = 0;
        if (lane == 31) base = atomicAdd(&list_n, (unsigned)total);
        base = __shfl_sync(0xffffffffu, base, 31);
        unsigned int slot = base + incl - mine;
        const unsigned int cell = (unsigned)(u0 + j * kFilterThreads) * V;
#pragma unroll
        for (int k = 0; k < V; ++k)
            if ((pass >> k) & 1u) {
                list_cell[slot] = cell + k;
                list_obj[slot] = obj[k];
                ++slot;
            }
    }
    __syncthreads();
    // ---- phase 2: one surviving cell per lane -------------------------------------------------------
    const int n = (int)list_n;
    for (int i0 = 0; i0 < n; i0 += kFilterThreads) {
        const int i = i0 + threadIdx.x;
        const bool valid = i < n;
        const unsigned int cell = valid ? list_cell[i] : 0u;
        const float obj = valid ? list_obj[i] : 0.f;
        ClsFromPlanes cls{pbase + (size_t)5 * plane + cell, (size_t)plane};
        filter_emit(f, prow0 + cell, g.no - 5, obj, valid, cls, f.counts + b, KeySinkGlobal(f.keys + (size_t)b * f.cap, f.cap));
    }
}

// ---- from a decoded prediction (bs, A, no): lane = one row ---------------------------------------
struct ClsFromRow {
    const float *r;
    __device__ __forceinline__ float load(int c) const { return __ldg(r + 5 + c); }
    __device__ __forceinline__ float prob(float raw) const { return raw; }
};

__global__ void __launch_bounds__(256) filter_z_kernel(const float *__restrict__ z, long long A, int no, const FilterOut f) {
    const int b = blockIdx.y;
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = row < A;
    const float *r = z + ((size_t)b * A + (valid ? row : 0)) * no;
    const float obj = valid ? __ldg(r + 4) : 0.f;
    filter_emit(f, (unsigned)row, no - 5, obj, valid, ClsFromRow{r}, f.counts + b, KeySinkGlobal(f.keys + (size_t)b * f.cap, f.cap));
}

}  // namespace dsn

using namespace dsn;

extern "C" __attribute__((visibility("default"))) int dsn_detect_decode(const float *const *raw, const int *ny, const int *nx, const float *stride,
                                 const float *anchor_grid, int nl, int bs, int na, int no, float *z,
                                 float *const *xperm, void *stream) {
    HeadGeom g;
    int rc = make_geom(g, raw, ny, nx, stride, anchor_grid, nl, bs, na, no);
    if (rc) return rc;
    if (bs == 0 || (!z && !xperm)) return DSN_OK;
    const int cells = no <= 32 ? 128 : 32;
    XpermPtrs xp{};
    for (int l = 0; l < nl; ++l) xp.p[l] = xperm ? xperm[l] : nullptr;
    int tiles = 0;
    for (int l = 0; l < nl; ++l) {
        g.tile_off[l] = tiles;
        tiles += na * ((g.plane[l] + cells - 1) / cells);
    }
    g.tile_off[nl] = tiles;
    const size_t smem = (size_t)2 * cells * no * sizeof(float);
    if (smem > 200 * 1024) return DSN_E_UNSUPPORTED;
    if (smem > 48 * 1024) {                   // large `no` (e.g. 85): opt in to more dynamic shared memory, once per device
        static DeviceOnce once;
        if (int rc2 = raise_dynamic_smem(once, detect_decode_kernel, 200 * 1024)) return rc2;
    }
    detect_decode_kernel<<<dim3(tiles, bs), 256, smem, (cudaStream_t)stream>>>(g, z, xp, cells);
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}

extern "C" __attribute__((visibility("default"))) int dsn_filter_levels(const float *const *raw, const int *ny, const int *nx, int nl, int bs, int na, int no,
                                 float conf_thres, int multi_label, const uint64_t *class_mask, uint64_t *keys,
                                 int64_t cap, int *counts, void *stream) {
    HeadGeom g;
    int rc = make_geom(g, raw, ny, nx, nullptr, nullptr, nl, bs, na, no);
    if (rc) return rc;
    FilterOut f;
    rc = make_filter_out(f, conf_thres, multi_label, no - 5, class_mask, keys, cap, counts);
    if (rc) return rc;
    if (bs == 0) return DSN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DSN_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)bs, st));
    bool vec = true;
    for (int l = 0; l < nl; ++l) vec = vec && (g.plane[l] % 4 == 0) && (((uintptr_t)raw[l] & 15) == 0);
    const int V = vec ? 4 : 1;
    int ctas = 0;
    for (int l = 0; l < nl; ++l) {
        g.cta_off[l] = ctas;
        g.ctas_per_plane[l] = (g.plane[l] / V + kFilterThreads * kFilterUnroll - 1) / (kFilterThreads * kFilterUnroll);
        ctas += na * g.ctas_per_plane[l];
    }
    g.cta_off[nl] = ctas;
    const float lo = obj_logit_lower_bound(conf_thres);
    dim3 grid(ctas, bs);
    if (vec) filter_levels_kernel<4><<<grid, kFilterThreads, 0, st>>>(g, f, lo);
    else filter_levels_kernel<1><<<grid, kFilterThreads, 0, st>>>(g, f, lo);
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}

extern "C" __attribute__((visibility("default"))) int dsn_filter_z(const float *z, int bs, int64_t A, int no, float conf_thres, int multi_label,
                            const uint64_t *class_mask, uint64_t *keys, int64_t cap, int *counts, void *stream) {
    if (!z || bs < 0 || A < 1 || no < 6) return DSN_E_INVALID;
    if (bs > 65535 || (unsigned long long)A * (unsigned long long)(no - 5) >= (1ull << 32)) return DSN_E_UNSUPPORTED;
    FilterOut f;
    int rc = make_filter_out(f, conf_thres, multi_label, no - 5, class_mask, keys, cap, counts);
    if (rc) return rc;
    if (bs == 0) return DSN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DSN_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)bs, st));
    dim3 grid((unsigned)((A + 255) / 256), bs);
    filter_z_kernel<<<grid, 256, 0, st>>>(z, A, no, f);
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}
