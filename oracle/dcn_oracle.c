/*
 * oracle/dcn_oracle.c — TEST INFRASTRUCTURE ONLY (the checker, never the product).
 *
 * CPU restatement of the reference's modulated deformable convolution (DCNv2),
 * forward and backward, in plain C.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.
 *
 * Semantics followed (paths relative to /root/reference/basicsr/models/ops/dcn/src/):
 *   bilinear sample ............ deform_conv_cuda_kernel.cu:467-497
 *   gather (im2col) ............ deform_conv_cuda_kernel.cu:570-633
 *   out = W·col + bias ......... deform_conv_cuda.cpp:518-568
 *   input-grad weights ......... deform_conv_cuda_kernel.cu:499-524
 *   coordinate-grad weights .... deform_conv_cuda_kernel.cu:526-568
 *   grad_input scatter ......... deform_conv_cuda_kernel.cu:635-693
 *   grad_offset / grad_mask .... deform_conv_cuda_kernel.cu:695-767
 *   backward GEMM order ........ deform_conv_cuda.cpp:617-681
 *
 * Parity pinning: the reference ships no golden vectors for this path;
 * this restatement is pinned by tests/test_oracle.py against torchvision 0.26
 * deform_conv2d (forward, grad_input, grad_mask, grad_weight, grad_bias everywhere;
 * grad_offset away from the exact -1 coordinate, where the reference returns 0 and
 * torchvision does not) and against tests/golden/dcn_ref_cuda_*.npz, produced by the
 * UNMODIFIED reference CUDA extension on a B200 (tools/first_light.py, section refext).
 *
 * The arithmetic is fp32 at the interface with fp64 accumulation inside the
 * contractions, so the oracle is at least as accurate as the reference's SGEMM.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int N, C, H, W;          /* input  [N,C,H,W]                          */
    int Cout, kh, kw;        /* weight [Cout, C/groups, kh, kw]            */
    int sh, sw, ph, pw, dh, dw; /* stride / pad / dilation per axis (v2 passes equal pairs) */
    int groups, dg;          /* weight groups, deformable groups           */
    int Ho, Wo;              /* derived                                    */
} dcn_shape;

static void derive(dcn_shape *s)
{
    s->Ho = (s->H + 2 * s->ph - (s->dh * (s->kh - 1) + 1)) / s->sh + 1;
    s->Wo = (s->W + 2 * s->pw - (s->dw * (s->kw - 1) + 1)) / s->sw + 1;
}

/* kernel.cu:467-497 — corners outside [0,H-1]x[0,W-1] contribute zero */
static float bilinear(const float *im, int H, int W, float h, float w)
{
    int h_low = (int)floorf(h), w_low = (int)floorf(w);
    int h_high = h_low + 1, w_high = w_low + 1;
    float lh = h - h_low, lw = w - w_low, hh = 1 - lh, hw = 1 - lw;
    float v1 = 0, v2 = 0, v3 = 0, v4 = 0;
    if (h_low >= 0 && w_low >= 0) v1 = im[h_low * W + w_low];
    if (h_low >= 0 && w_high <= W - 1) v2 = im[h_low * W + w_high];
    if (h_high <= H - 1 && w_low >= 0) v3 = im[h_high * W + w_low];
    if (h_high <= H - 1 && w_high <= W - 1) v4 = im[h_high * W + w_high];
    return hh * hw * v1 + hh * lw * v2 + lh * hw * v3 + lh * lw * v4;
}

/* kernel.cu:499-524 */
static float grad_weight_of_cell(float ah, float aw, int h, int w, int H, int W)
{
    if (ah <= -1 || ah >= H || aw <= -1 || aw >= W) return 0;
    int hl = (int)floorf(ah), wl = (int)floorf(aw), hh = hl + 1, wh = wl + 1;
    float wt = 0;
    if (h == hl && w == wl) wt = (h + 1 - ah) * (w + 1 - aw);
    if (h == hl && w == wh) wt = (h + 1 - ah) * (aw + 1 - w);
    if (h == hh && w == wl) wt = (ah + 1 - h) * (w + 1 - aw);
    if (h == hh && w == wh) wt = (ah + 1 - h) * (aw + 1 - w);
    return wt;
}

/* kernel.cu:526-568 — dir 0: d/dh, dir 1: d/dw */
static float coord_weight(float ah, float aw, int H, int W, const float *im, int dir)
{
    if (ah <= -1 || ah >= H || aw <= -1 || aw >= W) return 0;
    int hl = (int)floorf(ah), wl = (int)floorf(aw), hh = hl + 1, wh = wl + 1;
    float wt = 0;
    if (dir == 0) {
        if (hl >= 0 && wl >= 0) wt += -1 * (wl + 1 - aw) * im[hl * W + wl];
        if (hl >= 0 && wh <= W - 1) wt += -1 * (aw - wl) * im[hl * W + wh];
        if (hh <= H - 1 && wl >= 0) wt += (wl + 1 - aw) * im[hh * W + wl];
        if (hh <= H - 1 && wh <= W - 1) wt += (aw - wl) * im[hh * W + wh];
    } else {
        if (hl >= 0 && wl >= 0) wt += -1 * (hl + 1 - ah) * im[hl * W + wl];
        if (hl >= 0 && wh <= W - 1) wt += (hl + 1 - ah) * im[hl * W + wh];
        if (hh <= H - 1 && wl >= 0) wt += -1 * (ah - hl) * im[hh * W + wl];
        if (hh <= H - 1 && wh <= W - 1) wt += (ah - hl) * im[hh * W + wh];
    }
    return wt;
}

/* columns[(c*K + k), (ho*Wo + wo)] for one sample — kernel.cu:570-633 */
static void im2col_sample(const dcn_shape *s, const float *x, const float *off,
                          const float *msk, float *col)
{
    const int K = s->kh * s->kw, HW = s->Ho * s->Wo, cpg = s->C / s->dg;
#pragma omp parallel for schedule(static)
    for (int c = 0; c < s->C; ++c) {
        const int g = c / cpg;
        const float *im = x + (size_t)c * s->H * s->W;
        const float *og = off + (size_t)g * 2 * K * HW;
        const float *mg = msk ? msk + (size_t)g * K * HW : NULL;   /* NULL: DCNv1 (no modulation) */
        for (int ho = 0; ho < s->Ho; ++ho)
            for (int wo = 0; wo < s->Wo; ++wo) {
                const int p = ho * s->Wo + wo;
                for (int i = 0; i < s->kh; ++i)
                    for (int j = 0; j < s->kw; ++j) {
                        const int k = i * s->kw + j;
                        const float dh = og[(size_t)(2 * k) * HW + p];
                        const float dw = og[(size_t)(2 * k + 1) * HW + p];
                        const float m = mg ? mg[(size_t)k * HW + p] : 1.0f;
                        const float h_im = ho * s->sh - s->ph + i * s->dh + dh;
                        const float w_im = wo * s->sw - s->pw + j * s->dw + dw;
                        float v = 0;
                        if (h_im > -1 && w_im > -1 && h_im < s->H && w_im < s->W)
                            v = bilinear(im, s->H, s->W, h_im, w_im);
                        col[((size_t)c * K + k) * HW + p] = v * m;
                    }
            }
    }
}

/* returns 0 on success, negative on invalid arguments */
/* General form: per-axis stride/pad/dilation, mask may be NULL (== DCNv1,
 * deform_conv_cuda_kernel.cu:190-243 — same sampling rule without the modulation). */
int dcn_oracle_forward_ex(const float *x, const float *offset, const float *mask,
                          const float *weight, const float *bias /* may be NULL */,
                          float *out, int N, int C, int H, int W, int Cout, int kh,
                          int kw, int sh, int sw, int ph, int pw, int dh, int dw, int groups, int dg)
{
    dcn_shape s = {N, C, H, W, Cout, kh, kw, sh, sw, ph, pw, dh, dw, groups, dg, 0, 0};
    if (N < 0 || C <= 0 || Cout <= 0 || groups <= 0 || dg <= 0 || C % groups ||
        Cout % groups || C % dg || sh <= 0 || sw <= 0)
        return -1;
    derive(&s);
    if (s.Ho <= 0 || s.Wo <= 0) return -2;
    const int K = kh * kw, HW = s.Ho * s.Wo;
    const int cg = C / groups, og = Cout / groups; /* per weight group */
    float *col = (float *)malloc((size_t)C * K * HW * sizeof(float));
    if (!col) return -3;
    for (int n = 0; n < N; ++n) {
        im2col_sample(&s, x + (size_t)n * C * H * W, offset + (size_t)n * dg * 2 * K * HW,
                      mask ? mask + (size_t)n * dg * K * HW : NULL, col);
        /* deform_conv_cuda.cpp:550-568: out[n][g] = W[g].flatten(1) @ col[g] + bias */
#pragma omp parallel for schedule(static)
        for (int co = 0; co < Cout; ++co) {
            const int g = co / og;
            const float *wrow = weight + (size_t)co * cg * K;
            float *o = out + ((size_t)n * Cout + co) * HW;
            double *acc = (double *)calloc(HW, sizeof(double));
            for (int r = 0; r < cg * K; ++r) {
                const double wv = wrow[r];
                const float *crow = col + ((size_t)g * cg * K + r) * HW;
                for (int p = 0; p < HW; ++p) acc[p] += wv * crow[p];
            }
            const double b = bias ? bias[co] : 0.0;
            for (int p = 0; p < HW; ++p) o[p] = (float)(acc[p] + b);
            free(acc);
        }
    }
    free(col);
    return 0;
}

int dcn_oracle_forward(const float *x, const float *offset, const float *mask,
                       const float *weight, const float *bias, float *out, int N, int C, int H,
                       int W, int Cout, int kh, int kw, int stride, int pad, int dil, int groups, int dg)
{
    return dcn_oracle_forward_ex(x, offset, mask, weight, bias, out, N, C, H, W, Cout, kh, kw, stride,
                                 stride, pad, pad, dil, dil, groups, dg);
}

/* grad_weight / grad_bias are ACCUMULATED into (caller zero-fills), like
 * deform_conv_cuda.cpp:659-671; the other three grads are overwritten.
 * mask == NULL: DCNv1 backward (deform_conv_cuda_kernel.cu:279-436, deform_conv_cuda.cpp:239-488);
 * grad_mask is then ignored.  grad_weight contributions are multiplied by `scale`
 * (deform_conv_backward_parameters, deform_conv_cuda.cpp:471-478). */
int dcn_oracle_backward_ex(const float *x, const float *offset, const float *mask,
                           const float *weight, const float *grad_out, float *grad_x,
                           float *grad_offset, float *grad_mask, float *grad_weight,
                           float *grad_bias /* may be NULL */, float scale, int N, int C, int H, int W,
                           int Cout, int kh, int kw, int sh, int sw, int ph, int pw, int dh_, int dw_,
                           int groups, int dg)
{
    dcn_shape s = {N, C, H, W, Cout, kh, kw, sh, sw, ph, pw, dh_, dw_, groups, dg, 0, 0};
    const int stride_h = sh, stride_w = sw, pad_h = ph, pad_w = pw, dil_h = dh_, dil_w = dw_;
    if (N < 0 || C <= 0 || Cout <= 0 || groups <= 0 || dg <= 0 || C % groups ||
        Cout % groups || C % dg || sh <= 0 || sw <= 0)
        return -1;
    derive(&s);
    if (s.Ho <= 0 || s.Wo <= 0) return -2;
    const int K = kh * kw, HW = s.Ho * s.Wo, cpg = C / dg;
    const int cg = C / groups, og = Cout / groups;
    float *col = (float *)malloc((size_t)C * K * HW * sizeof(float));
    float *gcol = (float *)malloc((size_t)C * K * HW * sizeof(float));
    if (!col || !gcol) { free(col); free(gcol); return -3; }
    memset(grad_x, 0, (size_t)N * C * H * W * sizeof(float));

    for (int n = 0; n < N; ++n) {
        const float *xn = x + (size_t)n * C * H * W;
        const float *on = offset + (size_t)n * dg * 2 * K * HW;
        const float *mn = mask ? mask + (size_t)n * dg * K * HW : NULL;
        const float *gon = grad_out + (size_t)n * Cout * HW;
        float *gxn = grad_x + (size_t)n * C * H * W;
        float *goffn = grad_offset + (size_t)n * dg * 2 * K * HW;
        float *gmn = (mask && grad_mask) ? grad_mask + (size_t)n * dg * K * HW : NULL;

        /* gcol = W^T · grad_out   (deform_conv_cuda.cpp:623-626) */
#pragma omp parallel for schedule(static)
        for (int r = 0; r < C * K; ++r) {
            const int g = (r / K) / cg;         /* weight group of input channel */
            const int rr = r - g * cg * K;      /* row inside the group          */
            double *acc = (double *)calloc(HW, sizeof(double));
            for (int o = 0; o < og; ++o) {
                const int co = g * og + o;
                const double wv = weight[(size_t)co * cg * K + rr];
                const float *grow = gon + (size_t)co * HW;
                for (int p = 0; p < HW; ++p) acc[p] += wv * grow[p];
            }
            float *dst = gcol + (size_t)r * HW;
            for (int p = 0; p < HW; ++p) dst[p] = (float)acc[p];
            free(acc);
        }

        /* grad_offset, grad_mask  (kernel.cu:695-767) */
#pragma omp parallel for schedule(static)
        for (int oc = 0; oc < dg * 2 * K; ++oc) {
            const int g = oc / (2 * K), within = oc % (2 * K), k = within / 2, dir = within % 2;
            const int i = k / kw, j = k % kw;
            for (int ho = 0; ho < s.Ho; ++ho)
                for (int wo = 0; wo < s.Wo; ++wo) {
                    const int p = ho * s.Wo + wo;
                    const float dh = on[((size_t)g * 2 * K + 2 * k) * HW + p];
                    const float dw = on[((size_t)g * 2 * K + 2 * k + 1) * HW + p];
                    const float m = mn ? mn[((size_t)g * K + k) * HW + p] : 1.0f;
                    float ih = ho * stride_h - pad_h + i * dil_h + dh;
                    float iw = wo * stride_w - pad_w + j * dil_w + dw;
                    const int outside = (ih <= -1 || iw <= -1 || ih >= H || iw >= W);
                    if (outside) ih = iw = -2; /* kernel.cu:747-750 sentinel */
                    double val = 0, mval = 0;
                    for (int cc = 0; cc < cpg; ++cc) {
                        const int c = g * cpg + cc;
                        const float *im = xn + (size_t)c * H * W;
                        const float gc = gcol[((size_t)c * K + k) * HW + p];
                        if (!outside) mval += (double)gc * bilinear(im, H, W, ih, iw);
                        val += (double)coord_weight(ih, iw, H, W, im, dir) * gc * m;
                    }
                    goffn[(size_t)oc * HW + p] = (float)val;
                    if (dir == 0 && gmn) gmn[((size_t)g * K + k) * HW + p] = (float)mval;
                }
        }

        /* grad_input scatter (kernel.cu:635-693); serial per channel => deterministic */
#pragma omp parallel for schedule(static)
        for (int c = 0; c < C; ++c) {
            const int g = c / cpg;
            float *gim = gxn + (size_t)c * H * W;
            for (int k = 0; k < K; ++k) {
                const int i = k / kw, j = k % kw;
                for (int ho = 0; ho < s.Ho; ++ho)
                    for (int wo = 0; wo < s.Wo; ++wo) {
                        const int p = ho * s.Wo + wo;
                        const float dh = on[((size_t)g * 2 * K + 2 * k) * HW + p];
                        const float dw = on[((size_t)g * 2 * K + 2 * k + 1) * HW + p];
                        const float m = mn ? mn[((size_t)g * K + k) * HW + p] : 1.0f;
                        const float ih = ho * stride_h - pad_h + i * dil_h + dh;
                        const float iw = wo * stride_w - pad_w + j * dil_w + dw;
                        const float top = gcol[((size_t)c * K + k) * HW + p] * m;
                        const int ch = (int)ih, cw = (int)iw; /* truncation, :675-676 */
                        for (int dy = -2; dy <= 2; ++dy)
                            for (int dx = -2; dx <= 2; ++dx) {
                                const int yy = ch + dy, xx = cw + dx;
                                if (yy >= 0 && yy < H && xx >= 0 && xx < W &&
                                    fabsf(ih - yy) < 1 && fabsf(iw - xx) < 1)
                                    gim[yy * W + xx] +=
                                        grad_weight_of_cell(ih, iw, yy, xx, H, W) * top;
                            }
                    }
            }
        }

        /* recompute columns, grad_weight += grad_out · col^T, grad_bias += grad_out · 1
         * (deform_conv_cuda.cpp:647-671) */
        im2col_sample(&s, xn, on, mn, col);
#pragma omp parallel for schedule(static)
        for (int co = 0; co < Cout; ++co) {
            const int g = co / og;
            const float *grow = gon + (size_t)co * HW;
            for (int r = 0; r < cg * K; ++r) {
                const float *crow = col + ((size_t)g * cg * K + r) * HW;
                double acc = 0;
                for (int p = 0; p < HW; ++p) acc += (double)grow[p] * crow[p];
                grad_weight[(size_t)co * cg * K + r] += (float)(acc * scale);
            }
            if (grad_bias) {
                double acc = 0;
                for (int p = 0; p < HW; ++p) acc += grow[p];
                grad_bias[co] += (float)acc;
            }
        }
    }
    free(col);
    free(gcol);
    return 0;
}

int dcn_oracle_backward(const float *x, const float *offset, const float *mask,
                        const float *weight, const float *grad_out, float *grad_x,
                        float *grad_offset, float *grad_mask, float *grad_weight, float *grad_bias,
                        int N, int C, int H, int W, int Cout, int kh, int kw, int stride, int pad,
                        int dil, int groups, int dg)
{
    return dcn_oracle_backward_ex(x, offset, mask, weight, grad_out, grad_x, grad_offset, grad_mask,
                                  grad_weight, grad_bias, 1.0f, N, C, H, W, Cout, kh, kw, stride, stride,
                                  pad, pad, dil, dil, groups, dg);
}
