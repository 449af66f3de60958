/* TEST INFRASTRUCTURE ONLY — plain-C restatement of the integer / fp32 pieces of DeSeNet's
 * post-network path that the CUDA kernels must match bit-for-bit.  Never linked into the
 * product library.  Build: see oracle/c/Makefile (-ffp-contract=off: no FMA contraction, every
 * product and sum below is rounded separately, like the reference's scalar fp32 code paths).
 *
 * Reference file:line (splwany/DeSeNet) each routine follows is given at the routine.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- torchvision.ops.nms CPU kernel (third-party; call site core/utils/general.py:734) ----
 * order[] must be the stable descending-score permutation (ties: lower index first).
 * fp32 area / intersection / union, ovr = inter / (a_i + a_j - inter), suppress iff
 * (double)ovr > thr (strict; NaN never suppresses).  Returns number kept; keep[] in score order. */
typedef struct { float s; int64_t i; } sc_t;
static int cmp_desc_stable(const void *a, const void *b) {
    const sc_t *x = (const sc_t *)a, *y = (const sc_t *)b;
    if (x->s > y->s) return -1;
    if (x->s < y->s) return 1;
    return (x->i > y->i) - (x->i < y->i);
}

int64_t orc_nms_greedy(const float *boxes, const float *scores, int64_t n, double thr, int64_t *keep) {
    if (n <= 0) return 0;
    sc_t *ord = (sc_t *)malloc(sizeof(sc_t) * (size_t)n);
    float *area = (float *)malloc(sizeof(float) * (size_t)n);
    uint8_t *dead = (uint8_t *)calloc((size_t)n, 1);
    for (int64_t i = 0; i < n; ++i) {
        ord[i].s = scores[i]; ord[i].i = i;
        const float *b = boxes + 4 * i;
        area[i] = (b[2] - b[0]) * (b[3] - b[1]);
    }
    qsort(ord, (size_t)n, sizeof(sc_t), cmp_desc_stable);
    int64_t nk = 0;
    for (int64_t a = 0; a < n; ++a) {
        int64_t i = ord[a].i;
        if (dead[i]) continue;
        keep[nk++] = i;
        const float *bi = boxes + 4 * i;
        float ai = area[i];
        for (int64_t c = a + 1; c < n; ++c) {
            int64_t j = ord[c].i;
            if (dead[j]) continue;
            const float *bj = boxes + 4 * j;
            float xx1 = bi[0] > bj[0] ? bi[0] : bj[0];
            float yy1 = bi[1] > bj[1] ? bi[1] : bj[1];
            float xx2 = bi[2] < bj[2] ? bi[2] : bj[2];
            float yy2 = bi[3] < bj[3] ? bi[3] : bj[3];
            float w = xx2 - xx1; if (!(w > 0.0f)) w = 0.0f;
            float h = yy2 - yy1; if (!(h > 0.0f)) h = 0.0f;
            float inter = w * h;
            float ovr = inter / (ai + area[j] - inter);
            if ((double)ovr > thr) dead[j] = 1;
        }
    }
    free(ord); free(area); free(dead);
    return nk;
}

/* ---- nn.Upsample(bilinear, align_corners=True)   core/models/yolo.py:183 ----
 * ATen index rule in fp32: scale=(in-1)/(out-1); src=scale*dst; i0=(int)src; i1=i0+(i0<in-1);
 * l1=src-i0; l0=1-l1.  Each lerp `w0*a + w1*b` is evaluated as fma(w0, a, w1*b) — the
 * contraction ATen's CPU kernel is compiled with (probed: 100% bit-identical to
 * F.interpolate on torch 2.11 CPU; the un-contracted form matches only 62%).  Horizontal
 * lerps first, then the vertical one: out = fma(l0y, T, l1y*B), T = fma(l0x, v00, l1x*v01),
 * B = fma(l0x, v10, l1x*v11). */
static void coeffs(int n_in, int n_out, int *i0, int *i1, float *l0, float *l1) {
    float scale = n_out > 1 ? (float)(n_in - 1) / (float)(n_out - 1) : 0.0f;
    for (int d = 0; d < n_out; ++d) {
        float src = scale * (float)d;
        int a = (int)src; if (a > n_in - 1) a = n_in - 1;
        i0[d] = a; i1[d] = a + (a < n_in - 1 ? 1 : 0);
        l1[d] = src - (float)a; l0[d] = 1.0f - l1[d];
    }
}

void orc_upsample_bilinear_ac(const float *in, int64_t planes, int h, int w, int H, int W, float *out) {
    int *y0 = malloc(sizeof(int) * H), *y1 = malloc(sizeof(int) * H);
    int *x0 = malloc(sizeof(int) * W), *x1 = malloc(sizeof(int) * W);
    float *ly0 = malloc(sizeof(float) * H), *ly1 = malloc(sizeof(float) * H);
    float *lx0 = malloc(sizeof(float) * W), *lx1 = malloc(sizeof(float) * W);
    coeffs(h, H, y0, y1, ly0, ly1);
    coeffs(w, W, x0, x1, lx0, lx1);
    for (int64_t p = 0; p < planes; ++p) {
        const float *src = in + p * (int64_t)h * w;
        float *dst = out + p * (int64_t)H * W;
        for (int Y = 0; Y < H; ++Y) {
            const float *r0 = src + (int64_t)y0[Y] * w, *r1 = src + (int64_t)y1[Y] * w;
            for (int X = 0; X < W; ++X) {
                float t = fmaf(lx0[X], r0[x0[X]], lx1[X] * r0[x1[X]]);
                float b = fmaf(lx0[X], r1[x0[X]], lx1[X] * r1[x1[X]]);
                dst[(int64_t)Y * W + X] = fmaf(ly0[Y], t, ly1[Y] * b);
            }
        }
    }
    free(y0); free(y1); free(x0); free(x1); free(ly0); free(ly1); free(lx0); free(lx1);
}

/* ---- argmax over channels, first max wins, NaN maximal (torch.max / argmax semantics;
 *      core/utils/metrics.py:356,375, core/utils/plots.py:228) ---- */
void orc_argmax_first(const float *logits, int64_t bs, int C, int64_t hw, int32_t *out) {
    for (int64_t b = 0; b < bs; ++b)
        for (int64_t p = 0; p < hw; ++p) {
            const float *v = logits + b * C * hw + p;
            float best = v[0]; int bi = 0;
            for (int c = 1; c < C; ++c) {
                float cur = v[(int64_t)c * hw];
                if (cur > best || (isnan(cur) && !isnan(best))) { best = cur; bi = c; }
            }
            out[b * hw + p] = bi;
        }
}

/* ---- target-clamped confusion matrix: closed form behind batch_pix_accuracy /
 *      batch_intersection_union (core/utils/metrics.py:350-388).  Targets are wrapped to int32
 *      first, as metrics.py:359,380 (`astype('int32')`).  rows: see oracle/port.py:cm_rows (clamp into [0, rows-1], negatives -> 0). ---- */
void orc_confusion_matrix(const int32_t *pred, const int64_t *target, int64_t n, int C, int rows, int64_t *cm) {
    for (int64_t i = 0; i < n; ++i) {
        int32_t t = (int32_t)(uint32_t)(uint64_t)target[i];
        int r = t < 0 ? 0 : (t > rows - 1 ? rows - 1 : t);
        cm[(int64_t)r * C + pred[i]] += 1;
    }
}

/* ---- 1x1 convolution with bias as ONE fused multiply-add per term, in input-channel order, fp32:
 *      the heads' last convolutions, nn.Conv2d(c_hid, n_segcls, 1) (core/models/yolo.py:182) and Detect.m[i] =
 *      nn.Conv2d(ch[i], na * no, 1) (core/models/yolo.py:252, applied at :258).  The reference's own convolution (MKL-DNN /
 *      cuDNN) has an unspecified summation order, so this restatement pins the ORDER the CUDA producers use
 *      (acc = bias; acc = fma(w[n][k], x[k], acc) for k = 0 .. K-1), which the GPU tests then match bit for bit; its
 *      distance to the exact (float64) convolution is checked separately.
 *      x (bs, K, plane), w (N, K), bias (N) or NULL, out (bs, N, plane). ---- */
void orc_conv1x1_fma(const float *x, const float *w, const float *bias, int64_t bs, int K, int N, int64_t plane, float *out) {
    for (int64_t b = 0; b < bs; ++b)
        for (int n = 0; n < N; ++n)
            for (int64_t p = 0; p < plane; ++p) {
                float acc = bias ? bias[n] : 0.0f;
                for (int k = 0; k < K; ++k) acc = fmaf(w[(int64_t)n * K + k], x[(b * K + k) * plane + p], acc);
                out[(b * N + n) * plane + p] = acc;
            }
}
