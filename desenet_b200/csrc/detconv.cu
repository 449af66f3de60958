// Producer of the detection tail's input: the Detect head's output convolutions,
// nn.Conv2d(ch[i], na * no, kernel_size=1) with bias (core/models/yolo.py:252, applied at :258), one per level.
//     out[b, n, p] = bias[n] + sum_k W[n, k] * x[b, k, p]          x: (bs, K, ny, nx) NCHW, out: (bs, N, ny, nx) NCHW
// This one IS a contraction (N = 33 outputs per pixel for the reference's nc = 6, up to 255 at nc = 80; K = 128 / 256 /
// 512 input channels), so under --half (scripts/val.py:144,178-179), where the reference's cuDNN convolution runs on
// tensor cores with fp32 accumulation and rounds the result to fp16, it runs on the 5th-generation tensor cores:
//
//   detect_head_conv_umma_kernel<T = __half | __nv_bfloat16>
//     GEMM  D[M = 128 pixels, N_pad] += A[128 pixels, 16 k] * B[N_pad, 16 k]   tcgen05.mma.cta_group::1.kind::f16, fp32 in TMEM
//     * A = the activations.  NCHW means the pixels of one channel are contiguous: an MN-major operand.  A K-chunk of 64
//       channels x 128 pixels is brought in by per-thread 16-byte cp.async (8 pixels of one channel each, coalesced along
//       the pixel axis) straight into the canonical 128-byte-swizzle MN-major layout (the swizzle is applied to the
//       cp.async destination), kUmmaStages chunks in flight per CTA; partial tiles are zero-filled by cp.async's src-size.
//     * B = the weights, (N, K) row-major = K-major, laid out once per CTA in the canonical 128-byte-swizzle K-major
//       layout (rows >= N are zero); N_pad = N rounded up to 16 (48 for nc = 6).
//     * warp-specialised: 8 warps copy and later drain, one elected lane of a ninth warp issues the four K = 16 MMAs of
//       a chunk as soon as the chunk's mbarrier has counted all 256 cp.async completions; tcgen05.commit onto a second
//       mbarrier hands the stage back to the loaders, and a commit after the last chunk of a tile tells the epilogue that
//       the accumulator (one of two TMEM buffers) is complete.  No CTA-wide barrier inside the loop.
//     * epilogue: tcgen05.ld 32x32b (thread = pixel, 16 channels per load), + bias in fp32, round to T, one store per
//       (channel, pixel) - a warp writes 32 consecutive pixels of one channel plane.
//   The kernel is HBM-bound by two orders of magnitude (2 * N_pad flops per loaded element): the tensor cores are used
//   because N * K multiply-adds per pixel would otherwise cost ~N/2 CUDA-core instructions per loaded element.
//
//   detect_head_conv_f32_kernel: fp32 activations (the reference's default precision).  Exact fp32 arithmetic, one FMA
//   per term in channel order (what a plain fp32 convolution computes up to summation order); TF32 tensor cores would
//   break the 1e-5 box tolerance of the north star.  One thread = 2 pixels x 36 outputs, weights transposed in shared memory.
#include <algorithm>

#include "common.cuh"

namespace dsn {

// ---------------------------------------------------------------------------------------------------------------------
// fp32: CUDA cores, exact FMA chain
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kConvG = 36;      // outputs per thread: one group covers the 33 channels of nc = 6 (input read once)

// One thread = V pixels x kConvG outputs.  Weights sit transposed in shared memory, sw[k][kConvG], so a k step costs
// kConvG / 4 LDS.128 (broadcast) for V * kConvG FMAs: FMA-bound, not LDS-bound.
template <int V>
__global__ void __launch_bounds__(256) detect_head_conv_f32_kernel(const float *__restrict__ x, const float *__restrict__ weight,
                                                                   const float *__restrict__ bias, int K, int N, long long plane,
                                                                   float *__restrict__ out) {
    extern __shared__ __align__(16) float sw[];        // [K][kConvG] of this output group (columns >= N: zero)
    const int n0 = blockIdx.z * kConvG;
    for (int i = threadIdx.x; i < kConvG * K; i += blockDim.x) {
        const int c = i / K, k = i - c * K;            // coalesced read of weight row n0 + c
        sw[k * kConvG + c] = n0 + c < N ? weight[(size_t)(n0 + c) * K + k] : 0.f;
    }
    __syncthreads();
    const int b = blockIdx.y;
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;          // unit of V pixels
    if (u * V >= plane) return;
    const float *px = x + (size_t)b * K * plane + u * V;
    float acc[kConvG][V];
#pragma unroll
    for (int c = 0; c < kConvG; ++c)
#pragma unroll
        for (int j = 0; j < V; ++j) acc[c][j] = (bias && n0 + c < N) ? bias[n0 + c] : 0.f;
#pragma unroll 8
    for (int k = 0; k < K; ++k) {
        float v[V];
        if constexpr (V == 2) {
            const float2 t = __ldg(reinterpret_cast<const float2 *>(px + (size_t)k * plane));
            v[0] = t.x; v[1] = t.y;
        } else {
            v[0] = ldg_stream_f1(px + (size_t)k * plane);
        }
        const float4 *wk = reinterpret_cast<const float4 *>(sw + k * kConvG);
#pragma unroll
        for (int c4 = 0; c4 < kConvG / 4; ++c4) {
            const float4 w4 = wk[c4];
#pragma unroll
            for (int j = 0; j < V; ++j) {
                acc[c4 * 4 + 0][j] = __fmaf_rn(w4.x, v[j], acc[c4 * 4 + 0][j]);
                acc[c4 * 4 + 1][j] = __fmaf_rn(w4.y, v[j], acc[c4 * 4 + 1][j]);
                acc[c4 * 4 + 2][j] = __fmaf_rn(w4.z, v[j], acc[c4 * 4 + 2][j]);
                acc[c4 * 4 + 3][j] = __fmaf_rn(w4.w, v[j], acc[c4 * 4 + 3][j]);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < kConvG; ++c) {
        if (n0 + c < N) {
            float *po = out + ((size_t)b * N + n0 + c) * plane + u * V;
            if constexpr (V == 2) *reinterpret_cast<float2 *>(po) = make_float2(acc[c][0], acc[c][1]);
            else po[0] = acc[c][0];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// fp16 / bf16: tcgen05 + TMEM
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kUmmaM = 128;          // pixels per tile = TMEM lanes
constexpr int kUmmaKChunk = 64;      // channels per pipeline stage (four K = 16 MMAs)
constexpr int kUmmaStages = 3;
constexpr int kUmmaThreads = 256;
// Shared-memory operand layouts: the canonical 128-byte-swizzle forms (a first version with the no-swizzle "interleave"
// forms was correct but every MMA took ~0.4 us of tensor-pipe time - the operand fetch, not HBM, bounded the kernel).
//   A stage (MN-major, SW128): 2 pixel blocks x 8 k-groups x [8 k-rows x 128 bytes]; a row = 64 consecutive pixels of one
//     channel, its eight 16-byte pieces XOR-swizzled with the row index (address bits [4,7) ^= bits [7,10)).
//     SBO (next k-group) = 1024 bytes, LBO (next 64-pixel block) = 8192 bytes.  A quarter-warp's cp.async fill one row.
//   B (K-major, SW128): per 64-channel chunk a block of n_pad rows x 128 bytes, same swizzle; SBO (next 8 rows) = 1024.
constexpr int kABlockBytes = (kUmmaKChunk / 8) * 1024;    // one 64-pixel block of a stage: 8 192
constexpr int kAStageBytes = (kUmmaM / 64) * kABlockBytes; // 16 384

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, SWIZZLE_128B (layout type 2 in bits [61,64)): start address, leading and stride byte
// offsets in 16-byte units; bits [46,48) = 1 is the sm_100 descriptor version
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) |
           (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "CONV_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra CONV_DONE;\n"
        "bra CONV_WAIT;\n"
        "CONV_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

template <typename T>
struct UmmaFmt;
template <>
struct UmmaFmt<__half> {
    static constexpr uint32_t code = 0;                 // F16
    static __device__ __forceinline__ __half from_f32(float v) { return __float2half_rn(v); }
};
template <>
struct UmmaFmt<__nv_bfloat16> {
    static constexpr uint32_t code = 1;                 // BF16
    static __device__ __forceinline__ __nv_bfloat16 from_f32(float v) { return __float2bfloat16_rn(v); }
};

struct ConvParams {
    const void *x;
    const void *weight;      // (N, K) row-major, element type T
    const float *bias;       // (N) fp32 or null
    void *out;
    int bs, K, N, n_pad, tmem_cols;
    int plane;               // ny * nx, multiple of 8
    int tiles_per_image;
    long long total_tiles;
};

// Warp roles: warps 0-7 (256 threads) load the activations and run the epilogues; warp 8 (one elected lane) issues the
// MMAs.  Barriers: full[s] (256 cp.async completions, noinc), empty[s] (tcgen05.commit: the MMAs that read stage s are done),
// acc_full[a] / acc_empty[a] for the two accumulator buffers in TMEM (tile t uses buffer t & 1), so the MMAs of the next
// tile run while a tile's epilogue drains.  Nobody meets at a CTA barrier inside the loop.
template <typename T>
__global__ void __launch_bounds__(kUmmaThreads + 32) detect_head_conv_umma_kernel(const ConvParams p) {
    extern __shared__ __align__(1024) unsigned char conv_smem[];
    __shared__ __align__(8) unsigned long long bars[2 * kUmmaStages + 4];      // full[S], empty[S], acc_full[2], acc_empty[2]
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = p.K, nchunks = K / kUmmaKChunk;
    unsigned char *a_stages = conv_smem + ((1024u - (smem_u32(conv_smem) & 1023u)) & 1023u);   // swizzle atoms are 1024-byte aligned
    unsigned char *b_tile = a_stages + kUmmaStages * kAStageBytes;              // [K / 64][n_pad rows x 128 bytes], swizzled
    const uint32_t b_block = (uint32_t)p.n_pad * 128u;                           // one 64-channel chunk of the weights
    const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[kUmmaStages]);
    const uint32_t bar_acc_full = smem_u32(&bars[2 * kUmmaStages]), bar_acc_empty = smem_u32(&bars[2 * kUmmaStages + 2]);

    // ---- one-time set-up: barriers, TMEM, weights --------------------------------------------------------------------
    if (tid == 0) {
        for (int i = 0; i < kUmmaStages; ++i) {
            mbar_init(bar_full + 8 * i, kUmmaThreads);
            mbar_init(bar_empty + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_acc_full + 8 * i, 1);
            mbar_init(bar_acc_empty + 8 * i, kUmmaThreads);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"((uint32_t)p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {
        const T *w = static_cast<const T *>(p.weight);
        const int units = p.n_pad * (K / 8);                 // one unit = 8 consecutive k of one output row = 16 bytes
        for (int u = tid; u < units; u += kUmmaThreads + 32) {
            const int n = u / (K / 8), kg = u - n * (K / 8);
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (n < p.N) v = *reinterpret_cast<const uint4 *>(w + (size_t)n * K + kg * 8);
            *reinterpret_cast<uint4 *>(b_tile + (size_t)(kg >> 3) * b_block + (size_t)n * 128 + (((kg & 7) ^ (n & 7)) << 4)) = v;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic-proxy writes of B -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_slot;

    // ---- this CTA's tiles: blockIdx.x, + gridDim.x, ... ; its chunk stream = tiles x nchunks ---------------------------
    const long long my_tiles = p.total_tiles > blockIdx.x ? (p.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long total_chunks = my_tiles * nchunks;

    if (warp == kUmmaThreads / 32) {
        // ================================================ MMA issuer ===================================================
        if (lane == 0) {
            // instruction descriptor: D = F32, A = B = T, A MN-major, B K-major, N = n_pad, M = 128
            const uint32_t idesc = (1u << 4) | (UmmaFmt<T>::code << 7) | (UmmaFmt<T>::code << 10) | (1u << 15) | (0u << 16) |
                                   ((uint32_t)(p.n_pad >> 3) << 17) | ((uint32_t)(kUmmaM >> 4) << 24);
            long long t = 0;
            int c = 0;
            for (long long g = 0; g < total_chunks; ++g) {
                const int s = (int)(g % kUmmaStages);
                const uint32_t acc = (uint32_t)(t & 1);
                if (c == 0 && t >= 2) mbar_wait(bar_acc_empty + 8 * acc, (uint32_t)((t / 2 - 1) & 1));     // epilogue of tile t - 2 has read it
                mbar_wait(bar_full + 8 * s, (uint32_t)((g / kUmmaStages) & 1));                            // the chunk has landed
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = smem_u32(a_stages + (size_t)s * kAStageBytes);
                const uint32_t b_addr = smem_u32(b_tile) + (uint32_t)c * b_block;
#pragma unroll
                for (int kk = 0; kk < kUmmaKChunk / 16; ++kk) {
                    const uint64_t adesc = umma_desc(a_addr + kk * 2048, kABlockBytes, 1024);     // 16 channels = two k-groups further
                    const uint64_t bdesc = umma_desc(b_addr + kk * 32, 16, 1024);                 // 16 channels = 32 bytes into the row
                    umma_f16(tmem_d + acc * (uint32_t)p.n_pad, adesc, bdesc, idesc, (c > 0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(bar_empty + 8 * s);                                  // stage free once these MMAs have read it
                if (c == nchunks - 1) {
                    umma_commit(bar_acc_full + 8 * acc);                         // accumulator of the tile complete
                    c = 0;
                    ++t;
                } else {
                    ++c;
                }
            }
        }
    } else {
        // ========================================= loaders + epilogue (256 threads) ======================================
        // a thread copies 16-byte pieces: piece q of a chunk = (k = q / 16, pixel group mg = q % 16); 1024 pieces per chunk
        auto epilogue = [&](long long t) {
            const long long tile = blockIdx.x + t * gridDim.x;
            const int b = (int)(tile / p.tiles_per_image);
            const int p0 = (int)(tile - (long long)b * p.tiles_per_image) * kUmmaM;
            const uint32_t acc = (uint32_t)(t & 1);
            mbar_wait(bar_acc_full + 8 * acc, (uint32_t)((t / 2) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int lane_grp = warp & 3;                                        // a warp reads TMEM lanes 32 * (warp % 4) ...
            const int pix = p0 + lane_grp * 32 + lane;
            T *obase = static_cast<T *>(p.out) + (size_t)b * p.N * p.plane + pix;
            // warps 0-3 take the even 16-column groups, warps 4-7 the odd ones
            for (int cg = warp >> 2; cg * 16 < p.n_pad; cg += 2) {
                uint32_t r[16];
                tmem_ld16(tmem_d + ((uint32_t)(lane_grp * 32) << 16) + acc * (uint32_t)p.n_pad + (uint32_t)(cg * 16), r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (pix < p.plane) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int n = cg * 16 + j;
                        if (n < p.N) obase[(size_t)n * p.plane] = UmmaFmt<T>::from_f32(__fadd_rn(__uint_as_float(r[j]), p.bias ? p.bias[n] : 0.f));
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_acc_empty + 8 * acc) : "memory");
        };
        long long t_load = 0;           // tile / chunk-in-tile of the chunk being loaded
        int c_load = 0;
        long long t_epi = 0;            // next tile whose epilogue is due
        for (long long g = 0; g < total_chunks; ++g) {
            const int s = (int)(g % kUmmaStages);
            if (g >= kUmmaStages) mbar_wait(bar_empty + 8 * s, (uint32_t)((g / kUmmaStages - 1) & 1));   // chunk g - S has been consumed
            {
                const long long tile = blockIdx.x + t_load * gridDim.x;
                const int b = (int)(tile / p.tiles_per_image);
                const int p0 = (int)(tile - (long long)b * p.tiles_per_image) * kUmmaM;
                const T *src0 = static_cast<const T *>(p.x) + ((size_t)b * K + (size_t)c_load * kUmmaKChunk) * p.plane + p0;
                const uint32_t dst0 = smem_u32(a_stages + (size_t)s * kAStageBytes);
#pragma unroll
                for (int i = 0; i < (kUmmaKChunk * (kUmmaM / 8)) / kUmmaThreads; ++i) {
                    const int q = tid + i * kUmmaThreads;
                    const int k = q >> 4, mg = q & 15;
                    const bool inside = p0 + mg * 8 < p.plane;                    // plane % 8 == 0: a piece is all in or all out
                    const T *src = inside ? src0 + (size_t)k * p.plane + mg * 8 : src0;
                    const uint32_t dst = dst0 + (uint32_t)((mg >> 3) * kABlockBytes + (k >> 3) * 1024 + (k & 7) * 128 + (((mg & 7) ^ (k & 7)) << 4));
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(inside ? 16 : 0) : "memory");
                }
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar_full + 8 * s) : "memory");
            }
            if (++c_load == nchunks) { c_load = 0; ++t_load; }
            // chunk g - S was the last of tile t_epi: that tile's MMAs are complete (its stage was just handed back), drain it
            if (g >= kUmmaStages && (g - kUmmaStages + 1) == (t_epi + 1) * nchunks) epilogue(t_epi++);
        }
        for (; t_epi < my_tiles; ++t_epi) epilogue(t_epi);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

template <typename T>
static int launch_conv_umma(const ConvParams &p, cudaStream_t st) {
    const size_t smem = (size_t)kUmmaStages * kAStageBytes + (size_t)p.n_pad * p.K * sizeof(T) + 1024;
    if (smem > 200 * 1024) return DSN_E_UNSUPPORTED;
    static DeviceOnce once;           // per element type
    if (int rc = once.run([&]() {
            int e = (int)cudaFuncSetAttribute(detect_head_conv_umma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e == 0) e = (int)cudaFuncSetAttribute(detect_head_conv_umma_kernel<T>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                      (int)cudaSharedmemCarveoutMaxShared);
            return e;
        }))
        return rc;
    // resident CTAs per SM from the resources themselves (the occupancy API answers for the carveout in force at the time
    // of the call, which is not yet this kernel's on its first launch): 227 KB of shared memory with 1 KB reserved per CTA,
    // 512 TMEM columns, 64 K registers at 56 per thread
    int occ = (int)std::min<size_t>((227 * 1024) / (smem + 1024), (size_t)(512 / p.tmem_cols));
    occ = std::max(1, std::min(occ, 4));
    const long long grid = std::min<long long>(p.total_tiles, (long long)sm_count_current_device() * occ);
    detect_head_conv_umma_kernel<T><<<(unsigned)grid, kUmmaThreads + 32, smem, st>>>(p);
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}

}  // namespace dsn

using namespace dsn;

extern "C" __attribute__((visibility("default"))) int dsn_detect_head_conv1x1(const void *features, int dtype, const void *weight, const float *bias,
                                                                              int bs, int K, int N, int ny, int nx, void *out, void *stream) {
    if (!features || !weight || !out || bs < 0 || K < 1 || N < 1 || ny < 1 || nx < 1 || dtype_size(dtype) == 0) return DSN_E_INVALID;
    if (bs > 65535 || N > 256) return DSN_E_UNSUPPORTED;
    if (bs == 0) return DSN_OK;
    const long long plane = (long long)ny * nx;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DSN_F32) {
        if ((((uintptr_t)features | (uintptr_t)weight | (uintptr_t)out) & 3) != 0) return DSN_E_ALIGN;
        const size_t smem = (size_t)kConvG * K * sizeof(float);
        if (smem > 160 * 1024) return DSN_E_UNSUPPORTED;
        const bool vec = plane % 2 == 0 && (((uintptr_t)features | (uintptr_t)out) & 7) == 0;
        const long long units = vec ? plane / 2 : plane;
        dim3 grid((unsigned)((units + 255) / 256), bs, (N + kConvG - 1) / kConvG);
        static DeviceOnce once2, once1;
        if (vec) {
            if (int rc = raise_dynamic_smem(once2, detect_head_conv_f32_kernel<2>, 160 * 1024)) return rc;
            detect_head_conv_f32_kernel<2><<<grid, 256, smem, st>>>(static_cast<const float *>(features), static_cast<const float *>(weight), bias, K,
                                                                    N, plane, static_cast<float *>(out));
        } else {
            if (int rc = raise_dynamic_smem(once1, detect_head_conv_f32_kernel<1>, 160 * 1024)) return rc;
            detect_head_conv_f32_kernel<1><<<grid, 256, smem, st>>>(static_cast<const float *>(features), static_cast<const float *>(weight), bias, K,
                                                                    N, plane, static_cast<float *>(out));
        }
        DSN_CHECK_LAUNCH();
        return DSN_OK;
    }
    // tensor-core path: 16-byte pieces along the pixel axis and along K
    if (K % kUmmaKChunk != 0 || plane % 8 != 0 || plane > 0x7fffffffLL) return DSN_E_UNSUPPORTED;
    if ((((uintptr_t)features | (uintptr_t)weight) & 15) != 0 || ((uintptr_t)out & 1) != 0) return DSN_E_ALIGN;
    ConvParams p;
    p.x = features;
    p.weight = weight;
    p.bias = bias;
    p.out = out;
    p.bs = bs;
    p.K = K;
    p.N = N;
    p.n_pad = (N + 15) / 16 * 16;
    p.tmem_cols = 32;
    while (p.tmem_cols < 2 * p.n_pad) p.tmem_cols <<= 1;                  // two accumulator buffers
    p.plane = (int)plane;
    p.tiles_per_image = (int)((plane + kUmmaM - 1) / kUmmaM);
    p.total_tiles = (long long)bs * p.tiles_per_image;
    if (dtype == DSN_F16) return launch_conv_umma<__half>(p, st);
    return launch_conv_umma<__nv_bfloat16>(p, st);
}
