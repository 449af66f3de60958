// Detection-head geometry passed by value to the kernels (fits the 4 KB kernel-parameter space).
#pragma once
#include "common.cuh"

namespace dsn {

struct HeadGeom {
    const void *raw[DSN_MAX_LEVELS];    // (bs, na*no, ny, nx) per level, elements of type `dtype` (DSN_F32 / DSN_F16 / DSN_BF16)
    int dtype;
    int ny[DSN_MAX_LEVELS], nx[DSN_MAX_LEVELS], plane[DSN_MAX_LEVELS];
    int row_off[DSN_MAX_LEVELS + 1];    // first row of the level inside z (per image); [nl] = A
    int unit_off[DSN_MAX_LEVELS + 1];   // fused kernel: first work unit (V cells) of the level
    int cta_off[DSN_MAX_LEVELS + 1];    // filter kernels: first blockIdx.x of the level (grid = tiles x anchors x images)
    int ctas_per_plane[DSN_MAX_LEVELS]; // filter kernels: CTAs covering one (level, anchor) plane
    int tile_off[DSN_MAX_LEVELS + 1];   // decode kernel: first tile of the level
    float stride[DSN_MAX_LEVELS];
    float anchor_w[DSN_MAX_LEVELS][DSN_MAX_ANCHORS], anchor_h[DSN_MAX_LEVELS][DSN_MAX_ANCHORS];
    int nl, na, no, bs, A;
};

struct XpermPtrs {
    float *p[DSN_MAX_LEVELS];
};

// planes that 4-element vector loads can walk: plane % 4 == 0 and level pointers aligned to 4 elements
static inline bool geom_vec4(const HeadGeom &g) {
    const uintptr_t m = (uintptr_t)(4 * dtype_size(g.dtype) - 1);
    bool ok = true;
    for (int l = 0; l < g.nl; ++l) ok = ok && (g.plane[l] % 4 == 0) && (((uintptr_t)g.raw[l] & m) == 0);
    return ok;
}
// ... and that 1-D TMA bulk copies can walk in tiles of 128 cells: 16-byte aligned pointers, rows a multiple of 16 bytes
static inline bool geom_bulk(const HeadGeom &g) {
    const int per16 = 16 / dtype_size(g.dtype);
    bool ok = true;
    for (int l = 0; l < g.nl; ++l) ok = ok && (g.plane[l] % per16 == 0) && (((uintptr_t)g.raw[l] & 15) == 0);
    return ok;
}

// stride / anchor_grid may be NULL for kernels that never decode boxes (the key-only filter).
static inline int make_geom(HeadGeom &g, const void *const *raw, int dtype, const int *ny, const int *nx, const float *stride,
                            const float *anchor_grid, int nl, int bs, int na, int no) {
    if (!raw || !ny || !nx || nl < 1 || bs < 0 || na < 1 || no < 6) return DSN_E_INVALID;
    const int esize = dtype_size(dtype);
    if (esize == 0) return DSN_E_INVALID;
    if (nl > DSN_MAX_LEVELS || na > DSN_MAX_ANCHORS || bs > 65535) return DSN_E_UNSUPPORTED;
    g = HeadGeom{};
    g.nl = nl; g.na = na; g.no = no; g.bs = bs; g.dtype = dtype;
    long long rows = 0;
    for (int l = 0; l < nl; ++l) {
        if (!raw[l] || ny[l] < 1 || nx[l] < 1) return DSN_E_INVALID;
        if (((uintptr_t)raw[l] & (uintptr_t)(esize - 1)) != 0) return DSN_E_ALIGN;
        g.raw[l] = raw[l];
        g.ny[l] = ny[l]; g.nx[l] = nx[l];
        const long long plane = (long long)ny[l] * nx[l];
        if (plane > (1 << 28)) return DSN_E_UNSUPPORTED;
        g.plane[l] = (int)plane;
        g.row_off[l] = (int)rows;
        rows += (long long)na * plane;
        g.stride[l] = stride ? stride[l] : 0.f;
        for (int a = 0; a < na; ++a) {
            g.anchor_w[l][a] = anchor_grid ? anchor_grid[((size_t)l * na + a) * 2 + 0] : 0.f;
            g.anchor_h[l][a] = anchor_grid ? anchor_grid[((size_t)l * na + a) * 2 + 1] : 0.f;
        }
    }
    // candidate keys carry row*nc + cls in 32 bits
    if ((unsigned long long)rows * (unsigned long long)(no - 5) >= (1ull << 32)) return DSN_E_UNSUPPORTED;
    g.row_off[nl] = (int)rows;
    g.A = (int)rows;
    return DSN_OK;
}

// Level of a filter CTA from blockIdx.x, with STATIC indices into the by-value kernel parameter (dynamic indexing would
// make the compiler copy the arrays to local memory) and no integer division: the anchor is blockIdx.y.
struct TileGeom {
    const void *raw;
    int plane, row_off, tile;
};
__device__ __forceinline__ TileGeom tile_geom(const HeadGeom &g, int x) {
    TileGeom t{g.raw[0], g.plane[0], g.row_off[0], x};
#pragma unroll
    for (int i = 1; i < DSN_MAX_LEVELS; ++i)
        if (i < g.nl && x >= g.cta_off[i]) {
            t.raw = g.raw[i];
            t.plane = g.plane[i];
            t.row_off = g.row_off[i];
            t.tile = x - g.cta_off[i];
        }
    return t;
}

// (level, anchor, cell) of a row index inside one image, without integer division (nl, na are small).
struct RowGeom {
    int l, a, cell, plane;
};
__device__ __forceinline__ RowGeom row_geom(const HeadGeom &g, int row) {
    RowGeom r;
    r.l = 0;
    while (r.l + 1 < g.nl && row >= g.row_off[r.l + 1]) ++r.l;
    r.plane = g.plane[r.l];
    int c = row - g.row_off[r.l];
    r.a = 0;
    while (c >= r.plane) { c -= r.plane; ++r.a; }
    r.cell = c;
    return r;
}

// Box decode of one anchor row from the raw level planes (core/models/yolo.py:266-269), every op
// rounded separately so dsn_detect_decode and the NMS box fetch agree bit-for-bit.
__device__ __forceinline__ float decode_xy(float logit, int g, float stride) {
    const float two = __fmul_rn(sigmoidf_ref(logit), 2.0f);
    return __fmul_rn(__fadd_rn(__fsub_rn(two, 0.5f), (float)g), stride);
}
__device__ __forceinline__ float decode_wh(float logit, float anchor) {
    const float two = __fmul_rn(sigmoidf_ref(logit), 2.0f);
    return __fmul_rn(__fmul_rn(two, two), anchor);
}
// xywh2xyxy (core/utils/general.py:523-530)
__device__ __forceinline__ float4 xywh2xyxy(float x, float y, float w, float h) {
    const float hw = __fmul_rn(w, 0.5f), hh = __fmul_rn(h, 0.5f);
    return make_float4(__fsub_rn(x, hw), __fsub_rn(y, hh), __fadd_rn(x, hw), __fadd_rn(y, hh));
}

}  // namespace dsn
