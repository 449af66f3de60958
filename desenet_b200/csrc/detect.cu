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
#include <cstring>

#include <cooperative_groups.h>
#include <cuda.h>            // CUtensorMap (types only: the encoder is reached through cudaGetDriverEntryPoint)

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
template <typename T>
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
    const T *src = static_cast<const T *>(g.raw[l]) + ((size_t)(b * g.na + a) * no) * plane + c0;
    const float stride = g.stride[l], aw = g.anchor_w[l][a], ah = g.anchor_h[l][a];
    const int nx = g.nx[l];
    for (int i = threadIdx.x; i < no * n; i += blockDim.x) {
        const int o = i / n, c = i - o * n;
        const float v = Elem<T>::ld1(src + (size_t)o * plane + c);
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

template <typename T, int V>
__global__ void __launch_bounds__(kFilterThreads) filter_levels_kernel(const HeadGeom g, const FilterOut f, const float obj_logit_lo) {
    __shared__ unsigned int list_cell[kFilterThreads * kFilterUnroll * V];
    __shared__ float list_obj[kFilterThreads * kFilterUnroll * V];
    __shared__ unsigned int list_n;
    const int b = blockIdx.z, a = blockIdx.y;
    const TileGeom tg = tile_geom(g, (int)blockIdx.x);
    const int tile = tg.tile;
    const int plane = tg.plane;
    const int units = plane / V;
    const T *pbase = static_cast<const T *>(tg.raw) + ((size_t)(b * g.na + a) * g.no) * plane;
    const unsigned int prow0 = (unsigned)(tg.row_off + a * plane);
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
                const float4 t = Elem<T>::ld4(pbase + (size_t)4 * plane + u * 4);
                objraw[j][0] = t.x; objraw[j][1] = t.y; objraw[j][2] = t.z; objraw[j][3] = t.w;
            } else {
                objraw[j][0] = Elem<T>::ld1(pbase + (size_t)4 * plane + u);
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
        ClsFromPlanes<T> cls{pbase + (size_t)5 * plane + cell, (size_t)plane};
        filter_emit(f, prow0 + cell, g.no - 5, obj, valid, cls, f.counts + b, KeySinkGlobal(f.keys + (size_t)b * f.cap, f.cap),
                    f.cut_hi ? f.cut_hi + b : nullptr);
    }
}

// ---- dense multi-label filter --------------------------------------------------------------------------------------
// Eval setting (conf 0.001, multi_label): nearly every cell passes the objectness test and every class plane has to
// be streamed, so the kernel above is bound by instruction issue (one sigmoid per (cell, class)), not by HBM.  Here
// the test `sigmoid(cls) * obj > T` is first decided in LOGIT space against a per-cell bound
//     L = logit(T' / obj) - margin,   T' = (1 - 1e-3) * max(conf_thres, confidence of the cut)
// (one log per cell instead of one sigmoid per class): a class logit below L cannot give a candidate, whatever the
// last-ulp behaviour of the device sigmoid (margins are ~1000x its error).  The few (cell, class) pairs above L are
// compacted per warp through shared memory (16-bit indices into the warp's parked chunk of raw logits) and only
// those get the exact arithmetic (sigmoid, product, the strict compares of
// general.py:710-711 and the cut on the key) - one pair per lane, keys stored coalesced.
// One warp = 128 consecutive cells (4 per lane, 16-byte loads), classes in chunks of 8; one CTA = 1024 cells of
// one (image, level, anchor) plane.  Warps never synchronise with each other.
constexpr int kDenseThreads = 256;
constexpr int kDenseWarps = kDenseThreads / 32;
constexpr int kDenseCells = kDenseThreads * 4;
constexpr int kDenseChunk = 8;
constexpr int kDenseRound = 256;
struct DenseSmem {
    float tile[kDenseWarps][kDenseChunk][128];          // raw class logits of the warp's cells, one chunk
    float obj[kDenseWarps][128];                        // raw objectness logit per cell
    unsigned short idx[kDenseWarps][kDenseChunk * 128]; // compacted (class-in-chunk << 7 | cell) of the pairs above L
    unsigned long long kbuf[kDenseWarps][kDenseRound];  // keys of one round, written out together
};

template <typename T>
__global__ void __launch_bounds__(kDenseThreads, 3) filter_dense_kernel(const HeadGeom g, const FilterOut f, const float obj_logit_lo,
                                                                                   const float obj_logit_hi) {
    extern __shared__ __align__(16) unsigned char dense_smem_raw[];
    DenseSmem &sm = *reinterpret_cast<DenseSmem *>(dense_smem_raw);
    const int b = blockIdx.z, a = blockIdx.y;
    const TileGeom tg = tile_geom(g, (int)blockIdx.x);
    const int tile = tg.tile;
    const int plane = tg.plane;
    const int nc = g.no - 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cell0 = tile * kDenseCells + threadIdx.x * 4;       // plane % 4 == 0: the 4 cells are all in or all out
    const bool in = cell0 < plane;
    const T *pbase = static_cast<const T *>(tg.raw) + ((size_t)(b * g.na + a) * g.no) * plane + (in ? cell0 : 0);
    const unsigned int row0 = (unsigned)(tg.row_off + a * plane + tile * kDenseCells + warp * 128);   // warp's first row
    const unsigned int cut = f.cut_hi ? f.cut_hi[b] : 0xFFFFFFFFu;
    const float cut_conf = __uint_as_float(~cut);                 // keys <= cut  <=>  conf >= cut_conf  (0 when no cut)
    const float t_lo = fmaxf(f.conf_thres, cut_conf) * 0.999f;
    // ---- first chunk of class planes: issued together with the objectness load (one round trip, not two) ----------
    float4 v[kDenseChunk];
#pragma unroll
    for (int ci = 0; ci < kDenseChunk; ++ci) {
        v[ci] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        if (in && ci < nc) v[ci] = Elem<T>::ld4(pbase + (size_t)(5 + ci) * plane);
    }
    // ---- objectness: the per-cell logit bound L = -ln(obj / T' - 1) - margin, obj = 1 / (1 + exp(-o)) ------------------
    // Approximate exp / division / log are enough here: the margins (0.1 % on T, 4e-3 on L) are >1000x their error
    // wherever a candidate can exist (obj >= T, so obj / T' - 1 >= 1e-3 and the subtraction loses < 1e-3 relative).
    // The strict test obj > conf_thres (general.py:668) is decided in logit space outside the narrow band
    // [obj_logit_lo, obj_logit_hi) and by the exact sigmoid inside it.
    float L[4];
    bool any = false;
    {
        float4 o4 = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        if (in) o4 = Elem<T>::ld4(pbase + (size_t)4 * plane);
        const float ol[4] = {o4.x, o4.y, o4.z, o4.w};
        const float inv_t = __fdividef(1.0f, t_lo);                                      // +inf when T' == 0
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            L[k] = INFINITY;
            bool pass = ol[k] >= obj_logit_hi;
            if (!pass && ol[k] >= obj_logit_lo) pass = sigmoidf_ref(ol[k]) > f.conf_thres;   // yolo.py:266, general.py:668
            if (pass) {
                const float q = __fdividef(inv_t, 1.0f + __expf(-ol[k])) - 1.0f;
                if (q > 0.f) L[k] = -__logf(q) - 4e-3f;                                 // q == +inf -> -inf: everything passes
            }
            any = any || L[k] < INFINITY;
        }
        *reinterpret_cast<float4 *>(&sm.obj[warp][lane * 4]) = o4;                       // raw logits; sigmoid per survivor
    }
    if (!__any_sync(0xffffffffu, any)) return;
    unsigned long long *kout = f.keys + (size_t)b * f.cap;
    const unsigned int cap = f.cap > 0x7fffffffLL ? 0x7fffffffu : (unsigned)f.cap;
    int *counter = f.counts + b;
    // class planes in chunks of kDenseChunk; the NEXT chunk's loads are issued before the current chunk's survivors are
    // worked on, so the HBM round trip overlaps the compaction and the exact pass
    int nk = 0;                                         // keys waiting in the warp's kbuf (warp-uniform)
    auto flush = [&]() {                                // ONE atomic on the image's counter, coalesced 8-byte stores
        unsigned int base = 0;
        if (lane == 0) base = (unsigned)atomicAdd(counter, nk);
        base = __shfl_sync(0xffffffffu, base, 0);
        __syncwarp();
        for (int i = lane; i < nk; i += 32)
            if (base + i < cap) kout[base + i] = sm.kbuf[warp][i];
        __syncwarp();
        nk = 0;
    };
    for (int c0 = 0; c0 < nc; c0 += kDenseChunk) {
        // ---- logit-space test of the chunk in registers -------------------------------------------------------------
        unsigned int mask = 0;
#pragma unroll
        for (int ci = 0; ci < kDenseChunk; ++ci) {
            mask |= (v[ci].x > L[0] ? 1u : 0u) << (ci * 4 + 0);
            mask |= (v[ci].y > L[1] ? 1u : 0u) << (ci * 4 + 1);
            mask |= (v[ci].z > L[2] ? 1u : 0u) << (ci * 4 + 2);
            mask |= (v[ci].w > L[3] ? 1u : 0u) << (ci * 4 + 3);
        }
        const bool work = __any_sync(0xffffffffu, mask != 0);
        if (work) {                                     // park the raw logits for the exact pass
            __syncwarp();                               // previous chunk's readers are done with tile / idx
#pragma unroll
            for (int ci = 0; ci < kDenseChunk; ++ci) *reinterpret_cast<float4 *>(&sm.tile[warp][ci][lane * 4]) = v[ci];
        }
        if (c0 + kDenseChunk < nc) {
#pragma unroll
            for (int ci = 0; ci < kDenseChunk; ++ci) {
                v[ci] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
                if (any && c0 + kDenseChunk + ci < nc) v[ci] = Elem<T>::ld4(pbase + (size_t)(5 + c0 + kDenseChunk + ci) * plane);
            }
        }
        if (!work) continue;
        const int mine = __popc(mask);
        int incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        int slot = incl - mine;
        for (unsigned int m = mask; m; m &= m - 1) {
            const int bit = __ffs(m) - 1;
            sm.idx[warp][slot++] = (unsigned short)(((bit >> 2) << 7) | (lane * 4 + (bit & 3)));
        }
        __syncwarp();
        // ---- exact arithmetic for the surviving pairs, one per lane; keys are staged in the warp's kbuf and leave
        // together when it fills up (or at the end) ----------------------------------------------------------------
        for (int e0 = 0; e0 < total; e0 += 32) {
            if (nk > kDenseRound - 32) flush();
            const int e = e0 + lane;
            bool emit = false;
            unsigned long long key = 0ull;
            if (e < total) {
                const unsigned int id = sm.idx[warp][e];
                const int ci = id >> 7, cell = id & 127;
                const float conf = __fmul_rn(sigmoidf_ref(sm.tile[warp][ci][cell]), sigmoidf_ref(sm.obj[warp][cell]));   // general.py:703
                const unsigned int ck = conf_key(conf);
                emit = conf > f.conf_thres && ck <= cut && class_allowed(f, c0 + ci);                     // general.py:710,717
                key = ((unsigned long long)ck << 32) | (unsigned long long)((row0 + cell) * (unsigned)nc + (unsigned)(c0 + ci));
            }
            const unsigned int bal = __ballot_sync(0xffffffffu, emit);
            if (emit) sm.kbuf[warp][nk + __popc(bal & ((1u << lane) - 1u))] = key;
            nk += __popc(bal);
        }
    }
    if (nk > 0) flush();
}

// ---- dense multi-label filter, nc <= 8: TMA-fed persistent warps -------------------------------------------------------
// With few classes (the reference's own head: nc = 6) the kernel above is neither HBM- nor issue-bound but LATENCY-bound:
// a warp loads its tile (one DRAM round trip), works on it with nothing in flight, and leaves; the parked copy of the tile
// costs 68 KB of shared memory per CTA and caps the occupancy at 24 warps per SM (cfg 3: 0.46 of the HBM peak).  Here every
// warp is a small persistent pipeline of its own: it walks a contiguous range of tiles (128 cells x (1 + nc) planes) and
// keeps kTmaStages tiles in flight with TMA - ONE 3-D tensor-map copy per tile (cp.async.bulk.tensor.3d, box = 128 cells x
// (1 + nc) channels x 1 image; per-plane bulk copies when no tensor map can be encoded), completion on one mbarrier per
// stage - no registers hold data in flight, the tile arrives where the later passes index it, and nothing is parked.
// What is left is instruction issue, so the tile body is written for instruction count:
//   * a lane owns 4 cells; per cell the logit-space bound L of the kernel above, +inf for cells that cannot pass at all;
//   * per (cell, class) ONE subtraction L - logit whose sign bit is shifted into the lane's 24-bit mask (funnel shift):
//     two instructions per pair, no predicates;
//   * the pairs above the bound (4-5 % at the eval setting with the confidence cut) are appended, behind one warp scan
//     (one per lane and round for a dense tile), to a small PENDING list that survives across tiles, and the exact
//     arithmetic (sigmoid * sigmoid, the strict compares of general.py:710-711, the cut on the key) runs only on FULL
//     batches of 32 pending pairs;
//   * keys are staged per warp and leave with one atomic per batch of up to 128.
// Warps never synchronise with each other.
#ifndef DSN_TMA_WARPS
#define DSN_TMA_WARPS 8
#endif
#ifndef DSN_TMA_STAGES
#define DSN_TMA_STAGES 2
#endif
#ifndef DSN_FILTER_TENSORMAP
#define DSN_FILTER_TENSORMAP 1      // one 3-D tensor-map copy per tile instead of 1 + nc bulk copies
#endif
constexpr int kTmaWarps = DSN_TMA_WARPS;     // per CTA
constexpr int kTmaStages = DSN_TMA_STAGES;   // tiles in flight per warp (more stages = fewer resident warps: measured worse)
constexpr int kTmaKbuf = 128;                // keys staged per warp
constexpr int kTmaTileCells = 128;
constexpr int kTmaPend = 80;                 // pending (cell, class) pairs per warp: < 32 between tiles

struct TmaLevel {
    const void *raw;
    int plane, row_off, tiles;           // tiles per (image, anchor) plane
};

__device__ __forceinline__ uint32_t flt_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct TileCursor {
    int b, l, a, t;
};

// One 3-D tensor map per level over the raw head output (cells, channels, images): a tile = ONE
// cp.async.bulk.tensor.3d with box (128 cells, 1 + nc channels, 1 image) starting at (cell0, a * no + 4, b) - the same
// bytes in the same shared-memory order as the per-plane bulk copies, for one issued instruction instead of 1 + nc
// (each a 10-instruction uniform-register sequence).  Cells beyond the plane are zero-filled and counted.
struct FilterTensorMaps {
    CUtensorMap m[DSN_MAX_LEVELS];
};

// per-warp shared memory: stages [kTmaStages][planes][128] elements | kbuf | pending (class logit, obj logit, ref) | mbarriers
__host__ __device__ inline size_t tma_warp_bytes(int nc, int esize) {
    const size_t stage_bytes = (size_t)(nc + 1) * kTmaTileCells * esize;
    return ((kTmaStages * stage_bytes + kTmaKbuf * 8 + kTmaPend * 12 + kTmaStages * 8) + 127) & ~(size_t)127;
}

template <typename T, bool TENSOR>
__global__ void __launch_bounds__(kTmaWarps * 32) filter_dense_tma_kernel(const HeadGeom g, const FilterOut f, const int tiles_per_image,
                                                                         const long long total_tiles, const int tiles_per_warp,
                                                                         const __grid_constant__ FilterTensorMaps maps) {
    extern __shared__ __align__(128) unsigned char tma_smem[];
    __shared__ TmaLevel lv[DSN_MAX_LEVELS];
    const int nc = g.no - 5;
    const int planes = nc + 1;                                   // objectness + classes: channels 4 .. 4 + nc of an anchor
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t stage_bytes = (size_t)planes * kTmaTileCells * sizeof(T);
    unsigned char *wbase = tma_smem + tma_warp_bytes(nc, (int)sizeof(T)) * warp;
    T *stages = reinterpret_cast<T *>(wbase);
    unsigned long long *kbuf = reinterpret_cast<unsigned long long *>(wbase + kTmaStages * stage_bytes);
    float *pend_v = reinterpret_cast<float *>(kbuf + kTmaKbuf);                  // class logit of a pending pair
    float *pend_o = pend_v + kTmaPend;                                           // its objectness logit
    unsigned int *pend_r = reinterpret_cast<unsigned int *>(pend_o + kTmaPend);  // row * nc + class (low half of the key)
    unsigned long long *full = reinterpret_cast<unsigned long long *>(pend_r + kTmaPend);
#pragma unroll
    for (int i = 0; i < DSN_MAX_LEVELS; ++i)
        if ((int)threadIdx.x == i && i < g.nl)
            lv[i] = TmaLevel{g.raw[i], g.plane[i], g.row_off[i], (g.plane[i] + kTmaTileCells - 1) / kTmaTileCells};
    if (lane == 0) {
        for (int i = 0; i < kTmaStages; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(flt_smem_u32(&full[i])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const long long w_global = (long long)blockIdx.x * kTmaWarps + warp;
    const long long first = w_global * tiles_per_warp;
    if (first >= total_tiles) return;
    const int my_tiles = (int)min((long long)tiles_per_warp, total_tiles - first);

    auto advance = [&](TileCursor &c) {
        if (++c.t == lv[c.l].tiles) {
            c.t = 0;
            if (++c.a == g.na) {
                c.a = 0;
                if (++c.l == g.nl) { c.l = 0; ++c.b; }
            }
        }
    };
    TileCursor cons;
    {
        cons.b = (int)(first / tiles_per_image);
        int r = (int)(first - (long long)cons.b * tiles_per_image);
        cons.l = 0;
        while (r >= g.na * lv[cons.l].tiles) { r -= g.na * lv[cons.l].tiles; ++cons.l; }
        cons.a = r / lv[cons.l].tiles;
        cons.t = r - cons.a * lv[cons.l].tiles;
    }
    TileCursor prod = cons;
    // lane 0 posts the byte count and issues one bulk copy per plane (512-byte rows; row 0: objectness)
    auto issue = [&](const TileCursor &c, int st) {
        if (TENSOR) {
            if (lane == 0) {
                const uint32_t bar = flt_smem_u32(&full[st]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)stage_bytes) : "memory");
                asm volatile(
                    "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                        flt_smem_u32(stages + (size_t)st * planes * kTmaTileCells)),
                    "l"(reinterpret_cast<uint64_t>(&maps.m[c.l])), "r"(c.t * kTmaTileCells), "r"(c.a * g.no + 4), "r"(c.b), "r"(bar)
                    : "memory");
            }
            return;
        }
        if (lane == 0) {
            const TmaLevel L = lv[c.l];
            const int cell0 = c.t * kTmaTileCells;
            const unsigned bytes = (unsigned)min(kTmaTileCells, L.plane - cell0) * (unsigned)sizeof(T);
            const uint32_t bar = flt_smem_u32(&full[st]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes * (unsigned)planes) : "memory");
            const T *src = static_cast<const T *>(L.raw) + ((size_t)(c.b * g.na + c.a) * g.no + 4) * L.plane + cell0;
            uint32_t dst = flt_smem_u32(stages + (size_t)st * planes * kTmaTileCells);
            for (int pl = 0; pl < planes; ++pl) {
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                             "r"(bytes), "r"(bar)
                             : "memory");
                src += L.plane;
                dst += kTmaTileCells * (unsigned)sizeof(T);
            }
        }
    };
    const int prefill = min(kTmaStages, my_tiles);
    for (int i = 0; i < prefill; ++i) {
        issue(prod, i);
        advance(prod);
    }
    int issued = prefill;

    const unsigned int cap = f.cap > 0x7fffffffLL ? 0x7fffffffu : (unsigned)f.cap;
    const unsigned int lt = (1u << lane) - 1u;
    int nk = 0;                                         // keys waiting in the warp's kbuf (warp-uniform)
    int np = 0;                                         // pending pairs (warp-uniform)
    int kb = cons.b;                                    // image the staged keys and the pending pairs belong to
    unsigned long long *kout = f.keys + (size_t)kb * f.cap;
    auto flush = [&]() {                                // ONE atomic on the image's counter, coalesced 8-byte stores
        unsigned int base = 0;
        if (lane == 0) base = (unsigned)atomicAdd(f.counts + kb, nk);
        base = __shfl_sync(0xffffffffu, base, 0);
        __syncwarp();
        unsigned long long *dst = kout + base;
        if (base + (unsigned)nk <= cap) {
            for (int i = lane; i < nk; i += 32) dst[i] = kbuf[i];
        } else {
            for (int i = lane; i < nk; i += 32)
                if (base + i < cap) dst[i] = kbuf[i];
        }
        __syncwarp();
        nk = 0;
    };
    unsigned int cut;
    float inv_t;
    auto image_consts = [&]() {
        cut = f.cut_hi ? f.cut_hi[kb] : 0xFFFFFFFFu;
        const float t_lo = fmaxf(f.conf_thres, __uint_as_float(~cut)) * 0.999f;      // see filter_dense_kernel
        inv_t = __fdividef(1.0f, t_lo);                                              // +inf when T' == 0
    };
    image_consts();
    // exact arithmetic for the newest `n` (<= 32) pending pairs, one per lane                          general.py:703,710,717
    auto exact_batch = [&](int n) {
        if (nk > kTmaKbuf - 32) flush();
        const int e = np - n + lane;
        bool emit = false;
        unsigned long long key = 0ull;
        if (lane < n) {
            const unsigned int ref = pend_r[e];
            const float conf = __fmul_rn(sigmoidf_ref(pend_v[e]), sigmoidf_ref(pend_o[e]));
            const unsigned int ck = conf_key(conf);
            emit = conf > f.conf_thres && ck <= cut;
            if (f.use_cmask) emit = emit && class_allowed(f, (int)(ref % (unsigned)nc));
            key = ((unsigned long long)ck << 32) | (unsigned long long)ref;
        }
        const unsigned int bal = __ballot_sync(0xffffffffu, emit);
        if (emit) kbuf[nk + __popc(bal & lt)] = key;
        nk += __popc(bal);
        np -= n;
        __syncwarp();                                    // the pending slots are rewritten by the next insertion round
    };

    for (int it = 0; it < my_tiles; ++it) {
        const int st = it % kTmaStages;
        const unsigned parity = (unsigned)(it / kTmaStages) & 1u;
        if (cons.b != kb) {                              // next image: its keys go to another buffer, its cut differs
            if (np > 0) exact_batch(np);
            if (nk > 0) flush();
            kb = cons.b;
            kout = f.keys + (size_t)kb * f.cap;
            image_consts();
        }
        {
            const uint32_t bar = flt_smem_u32(&full[st]);
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "FLT_WAIT:\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                "@p bra FLT_DONE;\n"
                "bra FLT_WAIT;\n"
                "FLT_DONE:\n"
                "}\n" ::"r"(bar),
                "r"(parity)
                : "memory");
        }
        const TmaLevel L = lv[cons.l];
        const int cell0 = cons.t * kTmaTileCells;
        const T *tile = stages + (size_t)st * planes * kTmaTileCells;              // [planes][128]: row 0 objectness, 1 + c class c
        const bool in = lane * 4 < min(kTmaTileCells, L.plane - cell0);           // plane % 4 == 0: the 4 cells are all in or all out
        // ---- per-cell logit bound: sigmoid(cls) * sigmoid(obj) > T' needs cls > L = -ln(1 / (T' * sigmoid(obj)) - 1) ---------
        const float4 o4 = Elem<T>::lds4(tile + lane * 4);
        const float ol[4] = {o4.x, o4.y, o4.z, o4.w};
        float Lb[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float q = __fdividef(inv_t, 1.0f + __expf(-ol[k])) - 1.0f;      // NaN objectness -> NaN -> +inf below
            Lb[k] = q > 0.f ? -__logf(q) - 4e-3f : INFINITY;                      // q == +inf (T' == 0) -> -inf: everything passes
        }
        // ---- logit-space test: the sign of L - logit, shifted into the mask (pair p = class * 4 + cell -> bit 4 * nc - 1 - p) -
        unsigned int mask = 0;
#pragma unroll
        for (int ci = 0; ci < kDenseChunk; ++ci) {
            if (ci < nc) {
                const float4 v = Elem<T>::lds4(tile + (size_t)(1 + ci) * kTmaTileCells + lane * 4);
                mask = __funnelshift_l(__float_as_uint(Lb[0] - v.x), mask, 1);
                mask = __funnelshift_l(__float_as_uint(Lb[1] - v.y), mask, 1);
                mask = __funnelshift_l(__float_as_uint(Lb[2] - v.z), mask, 1);
                mask = __funnelshift_l(__float_as_uint(Lb[3] - v.w), mask, 1);
            }
        }
        if (!in) mask = 0;                               // cells beyond the plane: the stage holds stale bytes there
        // ---- pairs above the bound -> pending list, one per lane and round; full batches of 32 go through the exact pass ---
        const unsigned int row0 = (unsigned)(L.row_off + cons.a * L.plane + cell0) + (unsigned)lane * 4u;
        const int top = 4 * nc - 1;
        // pending-list entry of the pair at bit `pos` of this lane's mask
        auto put = [&](int slot, int pos) {
            const int pr = top - pos, ci = pr >> 2, k = pr & 3;
            pend_v[slot] = Elem<T>::cvt(tile[(1 + ci) * kTmaTileCells + lane * 4 + k]);
            pend_o[slot] = k == 0 ? ol[0] : (k == 1 ? ol[1] : (k == 2 ? ol[2] : ol[3]));
            pend_r[slot] = (row0 + (unsigned)k) * (unsigned)nc + (unsigned)ci;
        };
        if (__any_sync(0xffffffffu, mask != 0)) {
            const int mine = __popc(mask);
            int incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += up;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            if (np + total <= kTmaPend) {
                // the usual case (a few dozen pairs): every lane appends all of its pairs at once behind one warp scan
                int slot = np + incl - mine;
                for (unsigned int m = mask; m;) {
                    const int pos = 31 - __clz(m);           // highest set bit = lowest pair index
                    m &= ~(1u << pos);
                    put(slot++, pos);
                }
                np += total;
                __syncwarp();
                while (np >= 32) exact_batch(32);
            } else {
                // dense tile: one pair per lane and round, full batches drained as they form (np stays below 64)
                while (__any_sync(0xffffffffu, mask != 0)) {
                    const bool have = mask != 0;
                    const unsigned int bal = __ballot_sync(0xffffffffu, have);
                    if (have) {
                        const int pos = 31 - __clz(mask);
                        mask &= ~(1u << pos);
                        put(np + __popc(bal & lt), pos);
                    }
                    np += __popc(bal);
                    __syncwarp();
                    if (np >= 32) exact_batch(32);
                }
            }
        }
        __syncwarp();                                    // every lane is done with the stage
        if (issued < my_tiles) {                         // refill it with the tile kTmaStages ahead
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(prod, st);
            advance(prod);
            ++issued;
        }
        advance(cons);
    }
    if (np > 0) exact_batch(np);
    if (nk > 0) flush();
}

// ---- confidence cut for huge candidate sets (eval setting, multi-label) -------------------------------------------
// With conf 0.001 and multi_label the filter would emit up to A*nc candidates per image (1.28 M at nc = 80), while
// the greedy NMS stops after max_det survivors and never looks past the max_nms = 30000 best (general.py:728-736).
// A SAMPLE of the image (every kCutStride-th unit of 4 cells, all classes) is histogrammed by confidence; from it a
// cut is chosen above which about 2 x `target` candidates are expected, and the filter then emits only candidates at
// or above the cut - exactly the best ones, so the NMS result is unchanged as long as the NMS stops by itself.  If it
// runs out of emitted candidates first (cut_dropped[b] says a cut was in force) it reports det_count = -2 and the
// caller re-runs that batch without the cut.
// One CLUSTER of up to kCutCluster CTAs per image (sized at launch from the sampled work): every CTA histograms its share of the sample in its own shared memory,
// CTA 0 then sums the histograms over distributed shared memory and picks the bin.  A task = one sampled unit x one
// chunk of 8 classes, so an image with many classes spreads over all threads and each waits for two round trips.
constexpr int kCutStride = 64;
constexpr int kCutBins = 2048;
// the cut aims at kCutNum / kCutDen x `target` candidates: the estimate comes from the ~1.5 * target / kCutStride sampled
// candidates above the cut (192 at the default target: 7 % standard error), so 1.5 x leaves four to five standard errors
// before an image could starve - which would only cost the on-device redo, never the result
constexpr int kCutNum = 3, kCutDen = 2;
constexpr int kCutThreads = 256;
constexpr int kCutCluster = 4;

__device__ __forceinline__ unsigned int cut_bin(unsigned int key_hi, unsigned int lo_hi, int sh) {
    return key_hi <= lo_hi ? 0u : min((key_hi - lo_hi) >> sh, (unsigned)kCutBins - 1u);
}

template <typename T>
__global__ void __launch_bounds__(kCutThreads)
conf_cut_kernel(const HeadGeom g, const FilterOut f, const float obj_logit_lo, unsigned int lo_hi, int sh, unsigned int want,
                unsigned int *__restrict__ cut_hi, int *__restrict__ cut_dropped) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ unsigned int sh_hist[kCutBins];
    __shared__ unsigned int warp_sum[kCutThreads / 32];
    __shared__ unsigned int chosen;
    const int csize = (int)cluster.num_blocks();     // 1 .. kCutCluster CTAs per image (launch attribute)
    const int b = blockIdx.x / csize;
    const int rank = (int)cluster.block_rank();
    for (int i = threadIdx.x; i < kCutBins; i += kCutThreads) sh_hist[i] = 0u;
    if (threadIdx.x == 0) chosen = 0xFFFFFFFFu;
    __syncthreads();
    const int total_units = g.unit_off[g.nl];
    const int nc = g.no - 5;
    const int sampled = (total_units + kCutStride - 1) / kCutStride;
    const int nsplit = f.multi_label ? (nc + 7) / 8 : 1;          // best-class mode needs all classes of a cell in one thread
    const int span = f.multi_label ? 8 : nc;
    for (int t = rank * kCutThreads + threadIdx.x; t < sampled * nsplit; t += csize * kCutThreads) {
        const int su = t / nsplit;
        const int cbeg = (t - su * nsplit) * 8;
        const RowGeom rg = row_geom(g, su * kCutStride * 4);          // sampled unit = 4 consecutive rows
        const T *base = static_cast<const T *>(g.raw[rg.l]) + ((size_t)(b * g.na + rg.a) * g.no) * rg.plane + rg.cell;
        const float4 o4 = Elem<T>::ld4(base + (size_t)4 * rg.plane);
        const float ol[4] = {o4.x, o4.y, o4.z, o4.w};
        float obj[4];
        unsigned int pass = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            obj[k] = ol[k] >= obj_logit_lo ? sigmoidf_ref(ol[k]) : 0.f;
            pass |= obj[k] > f.conf_thres ? (1u << k) : 0u;
        }
        if (!pass) continue;
        float best[4] = {0.f, 0.f, 0.f, 0.f};
        int bcls[4] = {0, 0, 0, 0};
        for (int c0 = cbeg; c0 < min(nc, cbeg + span); c0 += 8) {      // 8 class planes per round trip
            float4 c4[8];
#pragma unroll
            for (int ci = 0; ci < 8; ++ci)
                c4[ci] = c0 + ci < nc ? Elem<T>::ld4(base + (size_t)(5 + c0 + ci) * rg.plane) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int ci = 0; ci < 8; ++ci) {
                const int c = c0 + ci;
                if (c >= nc) break;
                const float cl[4] = {c4[ci].x, c4[ci].y, c4[ci].z, c4[ci].w};
                const bool ok = class_allowed(f, c);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (!((pass >> k) & 1u)) continue;
                    const float conf = __fmul_rn(sigmoidf_ref(cl[k]), obj[k]);
                    if (f.multi_label) {
                        if (ok && conf > f.conf_thres) atomicAdd(&sh_hist[cut_bin(conf_key(conf), lo_hi, sh)], 1u);
                    } else if (c == 0 || argmax_better(conf, best[k])) {
                        best[k] = conf;
                        bcls[k] = c;
                    }
                }
            }
        }
        if (!f.multi_label) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (((pass >> k) & 1u) && best[k] > f.conf_thres && class_allowed(f, bcls[k]))
                    atomicAdd(&sh_hist[cut_bin(conf_key(best[k]), lo_hi, sh)], 1u);
        }
    }
    cluster.sync();                                  // every CTA's histogram is complete and visible cluster-wide
    if (rank == 0) {
        // prefix sum over the bins (best confidences first), summed over the cluster, scaled by the sampling stride
        constexpr int per = kCutBins / kCutThreads;
        static_assert(per * kCutThreads == kCutBins, "bins must split evenly over the CTA");
        unsigned int h[per], mine = 0;
#pragma unroll
        for (int j = 0; j < per; ++j) h[j] = sh_hist[threadIdx.x * per + j];
        for (int r = 1; r < csize; ++r) {
            const unsigned int *remote = cluster.map_shared_rank(sh_hist, r);      // DSMEM read
#pragma unroll
            for (int j = 0; j < per; ++j) h[j] += remote[threadIdx.x * per + j];
        }
#pragma unroll
        for (int j = 0; j < per; ++j) { h[j] *= kCutStride; mine += h[j]; }
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        unsigned int incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned int up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        unsigned int pre = incl - mine;
        for (int w = 0; w < warp; ++w) pre += warp_sum[w];
#pragma unroll
        for (int j = 0; j < per; ++j) {
            pre += h[j];
            if (pre >= want) { atomicMin(&chosen, (unsigned)(threadIdx.x * per + j)); break; }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            // keys in bins <= chosen pass; boundary in the confidence half of the key: lo_hi + ((chosen+1) << sh) - 1
            unsigned int c = 0xFFFFFFFFu;
            if (chosen != 0xFFFFFFFFu) {
                const unsigned long long edge = (unsigned long long)lo_hi + (((unsigned long long)chosen + 1ull) << sh) - 1ull;
                c = edge > 0xFFFFFFFFull ? 0xFFFFFFFFu : (unsigned int)edge;
            }
            cut_hi[b] = c;
            cut_dropped[b] = c != 0xFFFFFFFFu;        // a cut is in force: the filter may drop candidates of this image
        }
    }
    cluster.sync();                                  // keep the other CTAs' shared memory alive until CTA 0 has read it
}

// ---- from a decoded prediction (bs, A, no): lane = one row ---------------------------------------
template <typename T>
struct ClsFromRow {
    const T *r;
    __device__ __forceinline__ float load(int c) const { return Elem<T>::ld1c(r + 5 + c); }
    __device__ __forceinline__ float prob(float raw) const { return raw; }
};

template <typename T>
__global__ void __launch_bounds__(256) filter_z_kernel(const T *__restrict__ z, long long A, int no, const FilterOut f) {
    const int b = blockIdx.y;
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = row < A;
    const T *r = z + ((size_t)b * A + (valid ? row : 0)) * no;
    const float obj = valid ? Elem<T>::ld1c(r + 4) : 0.f;
    filter_emit(f, (unsigned)row, no - 5, obj, valid, ClsFromRow<T>{r}, f.counts + b, KeySinkGlobal(f.keys + (size_t)b * f.cap, f.cap));
}

}  // namespace dsn

using namespace dsn;

extern "C" __attribute__((visibility("default"))) int dsn_detect_decode(const void *const *raw, int dtype, const int *ny, const int *nx,
                                 const float *stride, const float *anchor_grid, int nl, int bs, int na, int no, float *z,
                                 float *const *xperm, void *stream) {
    HeadGeom g;
    int rc = make_geom(g, raw, dtype, ny, nx, stride, anchor_grid, nl, bs, na, no);
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
    DSN_DISPATCH_DTYPE(dtype, T, {
        if (smem > 48 * 1024) {               // large `no` (e.g. 85): opt in to more dynamic shared memory, once per device
            static DeviceOnce once;           // per element type
            if (int rc2 = raise_dynamic_smem(once, detect_decode_kernel<T>, 200 * 1024)) return rc2;
        }
        detect_decode_kernel<T><<<dim3(tiles, bs), 256, smem, (cudaStream_t)stream>>>(g, z, xp, cells);
    })
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}

extern "C" __attribute__((visibility("default"))) int dsn_filter_levels(const void *const *raw, int dtype, const int *ny, const int *nx, int nl, int bs,
                                 int na, int no, float conf_thres, int multi_label, const uint64_t *class_mask, uint64_t *keys,
                                 int64_t cap, int *counts, void *stream) {
    return dsn_filter_levels_cut(raw, dtype, ny, nx, nl, bs, na, no, conf_thres, multi_label, class_mask, keys, cap, counts, nullptr,
                                 stream);
}

// confidence half of a key for a confidence value / the histogram geometry shared by sampler, cut and filter
static inline unsigned int host_conf_key(float conf) {
    uint32_t b;
    memcpy(&b, &conf, 4);
    return ~b;
}
static inline void cut_geometry(float conf_thres, unsigned int &lo_hi, int &sh) {
    lo_hi = host_conf_key(1.0f);                                  // confidences above 1.0 fall into bin 0
    const unsigned int hi_hi = host_conf_key(conf_thres > 0.f ? conf_thres : 0.f);
    const unsigned int width = hi_hi > lo_hi ? hi_hi - lo_hi : 1u;
    sh = 0;
    while ((width >> sh) >= (unsigned)kCutBins) ++sh;
}

extern "C" __attribute__((visibility("default"))) int dsn_conf_cut_levels(const void *const *raw, int dtype, const int *ny, const int *nx, int nl,
                                                                          int bs, int na, int no, float conf_thres, int multi_label,
                                                                          const uint64_t *class_mask, int target,
                                                                          unsigned int *hist_ws, unsigned int *cut_hi,
                                                                          int *cut_dropped, void *stream) {
    HeadGeom g;
    int rc = make_geom(g, raw, dtype, ny, nx, nullptr, nullptr, nl, bs, na, no);
    if (rc) return rc;
    (void)hist_ws;                              // reserved (the histogram lives in shared memory)
    if (!cut_hi || !cut_dropped || target < 1 || target > (1 << 28)) return DSN_E_INVALID;
    FilterOut f;
    uint64_t dummy_keys;
    int dummy_counts;
    rc = make_filter_out(f, conf_thres, multi_label, no - 5, class_mask, &dummy_keys, 1, &dummy_counts);
    if (rc) return rc;
    if (bs == 0) return DSN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // the cut -> filter -> NMS -> redo chain needs planes the 4-element loads and the TMA bulk copies of the redo kernel can walk
    if (!geom_vec4(g) || !geom_bulk(g)) return DSN_E_UNSUPPORTED;   // caller: run without a cut
    int units = 0;
    for (int l = 0; l < nl; ++l) {
        g.unit_off[l] = units;
        units += na * g.plane[l] / 4;
    }
    g.unit_off[nl] = units;
    unsigned int lo_hi;
    int sh;
    cut_geometry(conf_thres, lo_hi, sh);
    // cluster size from the amount of sampled work: one CTA per image when a single pass of its threads covers it
    const int sampled = (units + kCutStride - 1) / kCutStride;
    const int tasks = sampled * (f.multi_label ? (no - 5 + 7) / 8 : 1);
    int csize = 1;
    while (csize < kCutCluster && tasks > csize * kCutThreads) csize <<= 1;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)(bs * csize));
    lc.blockDim = dim3(kCutThreads);
    lc.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    lc.attrs = attr;
    lc.numAttrs = 1;
    DSN_DISPATCH_DTYPE(dtype, T, {
        DSN_CUDA(cudaLaunchKernelEx(&lc, conf_cut_kernel<T>, g, f, obj_logit_lower_bound(conf_thres), lo_hi, sh,
                                    (unsigned)((long long)kCutNum * target / kCutDen), cut_hi, cut_dropped));
    })
    return DSN_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (libcuda is not linked)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        (void)cudaGetLastError();
    }
    return fn;
}
// maps of all levels; false when the encoder is missing or refuses a level (the caller then uses the per-plane bulk copies)
static bool make_filter_tensor_maps(const HeadGeom &g, int nl, int bs, int na, int no, FilterTensorMaps &maps) {
#if DSN_FILTER_TENSORMAP
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return false;
    const int es = dtype_size(g.dtype);
    const CUtensorMapDataType dt = g.dtype == DSN_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                                       : (g.dtype == DSN_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
    for (int l = 0; l < nl; ++l) {
        const cuuint64_t dims[3] = {(cuuint64_t)g.plane[l], (cuuint64_t)na * no, (cuuint64_t)bs};
        const cuuint64_t strides[2] = {(cuuint64_t)g.plane[l] * es, (cuuint64_t)g.plane[l] * es * na * no};
        const cuuint32_t box[3] = {(cuuint32_t)kTmaTileCells, (cuuint32_t)(no - 4), 1u};
        const cuuint32_t estr[3] = {1u, 1u, 1u};
        if (enc(&maps.m[l], dt, 3, const_cast<void *>(g.raw[l]), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    return true;
#else
    (void)g; (void)nl; (void)bs; (void)na; (void)no; (void)maps;
    return false;
#endif
}

template <typename T>
static int launch_filter(HeadGeom &g, const FilterOut &f, float conf_thres, int nl, int bs, int na, int no, cudaStream_t st) {
    const bool vec = geom_vec4(g);
    const int V = vec ? 4 : 1;
    int ctas = 0;
    for (int l = 0; l < nl; ++l) {
        g.cta_off[l] = ctas;
        g.ctas_per_plane[l] = (g.plane[l] / V + kFilterThreads * kFilterUnroll - 1) / (kFilterThreads * kFilterUnroll);
        ctas += g.ctas_per_plane[l];
    }
    g.cta_off[nl] = ctas;
    const float lo = obj_logit_lower_bound(conf_thres);
    if (vec && f.multi_label && no - 5 <= kDenseChunk && geom_bulk(g)) {      // few classes: TMA-fed persistent warps
        const int nc = no - 5;
        int tiles_per_image = 0;
        for (int l = 0; l < nl; ++l) tiles_per_image += na * ((g.plane[l] + kTmaTileCells - 1) / kTmaTileCells);
        const long long total_tiles = (long long)bs * tiles_per_image;
        const size_t smem = tma_warp_bytes(nc, (int)sizeof(T)) * kTmaWarps;
        FilterTensorMaps maps;                  // a by-value kernel argument (captured with the launch in a CUDA graph)
        const bool tensor = make_filter_tensor_maps(g, nl, bs, na, no, maps);
        static DeviceOnce once_t, once_b;       // per element type and form
        if (int rc2 = tensor ? raise_dynamic_smem(once_t, filter_dense_tma_kernel<T, true>, 200 * 1024)
                             : raise_dynamic_smem(once_b, filter_dense_tma_kernel<T, false>, 200 * 1024))
            return rc2;
        // resident CTAs per SM from the resources themselves (the occupancy API answers for the carveout in force when it is
        // called, which is not yet this kernel's before its first launch): 227 KB of shared memory, 1 KB reserved per CTA
        const int occ = std::max(1, std::min((int)((227 * 1024) / (smem + 1024)), 2048 / (kTmaWarps * 32)));
        // a warp takes a contiguous range of at least kTmaStages + 1 tiles; the grid is what the SMs hold at once
        const long long max_warps = (long long)sm_count_current_device() * occ * kTmaWarps;
        long long tpw = std::max<long long>(kTmaStages + 1, (total_tiles + max_warps - 1) / max_warps);
        const long long warps = (total_tiles + tpw - 1) / tpw;
        const long long grid = (warps + kTmaWarps - 1) / kTmaWarps;
        if (tensor)
            filter_dense_tma_kernel<T, true><<<(unsigned)grid, kTmaWarps * 32, smem, st>>>(g, f, tiles_per_image, total_tiles, (int)tpw, maps);
        else
            filter_dense_tma_kernel<T, false><<<(unsigned)grid, kTmaWarps * 32, smem, st>>>(g, f, tiles_per_image, total_tiles, (int)tpw, maps);
        DSN_CHECK_LAUNCH();
        return DSN_OK;
    }
    if (vec && f.multi_label) {                // dense multi-label path: every class plane is streamed
        ctas = 0;
        for (int l = 0; l < nl; ++l) {
            g.cta_off[l] = ctas;
            g.ctas_per_plane[l] = (g.plane[l] + kDenseCells - 1) / kDenseCells;
            ctas += g.ctas_per_plane[l];
        }
        g.cta_off[nl] = ctas;
        static DeviceOnce once;               // per element type
        if (int rc2 = raise_dynamic_smem(once, filter_dense_kernel<T>, (int)sizeof(DenseSmem))) return rc2;
        filter_dense_kernel<T><<<dim3(ctas, na, bs), kDenseThreads, sizeof(DenseSmem), st>>>(g, f, lo, obj_logit_upper_bound(conf_thres));
        DSN_CHECK_LAUNCH();
        return DSN_OK;
    }
    dim3 grid(ctas, na, bs);
    if (vec) filter_levels_kernel<T, 4><<<grid, kFilterThreads, 0, st>>>(g, f, lo);
    else filter_levels_kernel<T, 1><<<grid, kFilterThreads, 0, st>>>(g, f, lo);
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}

extern "C" __attribute__((visibility("default"))) int dsn_filter_levels_cut(const void *const *raw, int dtype, const int *ny, const int *nx, int nl, int bs,
                                 int na, int no, float conf_thres, int multi_label, const uint64_t *class_mask, uint64_t *keys,
                                 int64_t cap, int *counts, const unsigned int *cut_hi, void *stream) {
    HeadGeom g;
    int rc = make_geom(g, raw, dtype, ny, nx, nullptr, nullptr, nl, bs, na, no);
    if (rc) return rc;
    FilterOut f;
    rc = make_filter_out(f, conf_thres, multi_label, no - 5, class_mask, keys, cap, counts);
    if (rc) return rc;
    f.cut_hi = cut_hi;
    if (bs == 0) return DSN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DSN_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)bs, st));
    DSN_DISPATCH_DTYPE(dtype, T, { return launch_filter<T>(g, f, conf_thres, nl, bs, na, no, st); })
    return DSN_OK;
}

extern "C" __attribute__((visibility("default"))) int dsn_filter_z(const void *z, int dtype, int bs, int64_t A, int no, float conf_thres, int multi_label,
                            const uint64_t *class_mask, uint64_t *keys, int64_t cap, int *counts, void *stream) {
    if (!z || bs < 0 || A < 1 || no < 6 || dtype_size(dtype) == 0) return DSN_E_INVALID;
    if (((uintptr_t)z & (uintptr_t)(dtype_size(dtype) - 1)) != 0) return DSN_E_ALIGN;
    if (bs > 65535 || (unsigned long long)A * (unsigned long long)(no - 5) >= (1ull << 32)) return DSN_E_UNSUPPORTED;
    FilterOut f;
    int rc = make_filter_out(f, conf_thres, multi_label, no - 5, class_mask, keys, cap, counts);
    if (rc) return rc;
    if (bs == 0) return DSN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DSN_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)bs, st));
    dim3 grid((unsigned)((A + 255) / 256), bs);
    DSN_DISPATCH_DTYPE(dtype, T, { filter_z_kernel<T><<<grid, 256, 0, st>>>(static_cast<const T *>(z), A, no, f); })
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}
