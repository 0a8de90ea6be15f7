/*
 * oracle/oracle_train.c -- CPU ORACLE, training side.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * See orc_train.inc for provenance.  Parity status: unpinned against the JVM (oracle.c header).
 */
#define _GNU_SOURCE
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "orc_math.h"
#include "oracle.h"

#define REAL float
#define SUF(x) x##_f32
#define FMA fmaf
#define EXP orc_expf
#define SQRT sqrtf
#include "orc_din.inc"
#include "orc_train.inc"
#undef REAL
#undef SUF
#undef FMA
#undef EXP
#undef SQRT

#define REAL double
#define SUF(x) x##_f64
#define FMA fma
#define EXP orc_exp
#define SQRT sqrt
#include "orc_din.inc"
#include "orc_train.inc"
#undef REAL
#undef SUF
#undef FMA
#undef EXP
#undef SQRT

/* params: compact vector (read), grad: same length, zeroed here (zeroGradParameters) */
int orc_din_gradients_f32(int64_t rows, int E, int T, const float *params, int64_t n, const int32_t *node,
                          const int32_t *seq, const int32_t *mask_flat, int64_t n_mask, const float *labels,
                          float *grad, float *loss)
{
    orc_din_f32 m;
    const float *emb = params, *watt = emb + rows * E, *w1 = watt + (int64_t)E * E;
    const float *b1 = w1 + (int64_t)2 * E * E, *w2 = b1 + E, *b2 = w2 + E;
    if (orc_din_init_f32(&m, rows, E, T, emb, watt, w1, b1, w2, b2)) return -9;
    memset(grad, 0, sizeof(float) * (size_t)(rows * E + 3 * (int64_t)E * E + 2 * E + 1));
    int rc = orc_din_grad_f32(&m, n, node, seq, mask_flat, n_mask, labels, grad, loss);
    orc_din_free_f32(&m);
    return rc;
}

int orc_din_gradients_f64(int64_t rows, int E, int T, const double *params, int64_t n, const int32_t *node,
                          const int32_t *seq, const int32_t *mask_flat, int64_t n_mask, const double *labels,
                          double *grad, double *loss)
{
    orc_din_f64 m;
    const double *emb = params, *watt = emb + rows * E, *w1 = watt + (int64_t)E * E;
    const double *b1 = w1 + (int64_t)2 * E * E, *w2 = b1 + E, *b2 = w2 + E;
    if (orc_din_init_f64(&m, rows, E, T, emb, watt, w1, b1, w2, b2)) return -9;
    memset(grad, 0, sizeof(double) * (size_t)(rows * E + 3 * (int64_t)E * E + 2 * E + 1));
    int rc = orc_din_grad_f64(&m, n, node, seq, mask_flat, n_mask, labels, grad, loss);
    orc_din_free_f64(&m);
    return rc;
}

void orc_adam_f32(float *w, const float *g, float *s, float *r, int64_t n, double lr, int t)
{
    orc_adam_impl_f32(w, g, s, r, n, lr, t);
}
void orc_adam_f64(double *w, const double *g, double *s, double *r, int64_t n, double lr, int t)
{
    orc_adam_impl_f64(w, g, s, r, n, lr, t);
}

/* ---- DeepFM in the training loop (tdm/src/main/scala/com/mass/tdm/model/DeepFM.scala:11-44 behind LocalOptimizer.trainBatch) -----
 * forward as orc_deepfm_row_f32 (oracle.c); BCECriterionWithLogits (mean); backward of Add / Linear(T+1, 1) / ReLU /
 * Linear((T+1)E, T+1) / Concat / Reshape / FM (scalann/.../nn/FM.scala:46-72: gradInput_j = (buffer - F_j) * gradOutput) /
 * EmbeddingShare scatter-add (padding skipped).  Gradient layout = the compact vector [emb | W1 | b1 | W2 | b2], zeroed here.
 * Rows in order, accumulated in float like a single-threaded run (tolerance-based parity as for the DIN step). */
int orc_deepfm_gradients_f32(int64_t rows, int E, int T, const float *params, int64_t n, const int32_t *node, const int32_t *seq,
                             const float *labels, float *grad, float *loss_out)
{
    const int F = T + 1, IN = F * E;
    const float *emb = params, *w1 = params + rows * E, *b1 = w1 + (int64_t)F * IN, *w2 = b1 + F, *b2 = w2 + F;
    memset(grad, 0, sizeof(float) * (size_t)(rows * E + (int64_t)F * IN + 2 * F + 1));
    float *g_emb = grad, *g_w1 = grad + rows * E, *g_b1 = g_w1 + (int64_t)F * IN, *g_w2 = g_b1 + F, *g_b2 = g_w2 + F;
    float *X = (float *)malloc(sizeof(float) * (size_t)IN), *buf = (float *)malloc(sizeof(float) * (size_t)E);
    float *z = (float *)malloc(sizeof(float) * (size_t)F), *dz = (float *)malloc(sizeof(float) * (size_t)F);
    float *dX = (float *)malloc(sizeof(float) * (size_t)IN);
    const float inv_n = (float)(1.0 / (double)n);
    float loss = 0.0f;
    int rc = 0;
    for (int64_t r = 0; r < n && !rc; r++) {
        for (int s = 0; s < F; s++) {
            const int32_t c = s == 0 ? node[r] : seq[r * T + s - 1];
            if (c < -1 || (int64_t)c >= rows) { rc = -1; break; }
            for (int k = 0; k < E; k++) X[s * E + k] = c < 0 ? 0.0f : emb[(size_t)c * E + k];
        }
        if (rc) break;
        for (int k = 0; k < E; k++) buf[k] = 0.0f;
        for (int s = 0; s < F; s++) for (int k = 0; k < E; k++) buf[k] = buf[k] + X[s * E + k];
        float sum_square = 0.0f, square_sum = 0.0f;
        for (int k = 0; k < E; k++) sum_square = fmaf(buf[k], buf[k], sum_square);
        for (int k = 0; k < IN; k++) square_sum = fmaf(X[k], X[k], square_sum);
        const float fm = (sum_square - square_sum) / 2.0f;
        float dnn = 0.0f;
        for (int o = 0; o < F; o++) {
            float acc = 0.0f;
            for (int k = 0; k < IN; k++) acc = fmaf(X[k], w1[(size_t)o * IN + k], acc);
            z[o] = acc + b1[o];
            dnn = fmaf(z[o] > 0.0f ? z[o] : 0.0f, w2[o], dnn);
        }
        const float y = fm + (dnn + b2[0]);
        const float t = labels[r], ay = y < 0 ? -y : y;
        loss = loss + ((y > 0 ? y : 0.0f) - y * t + (float)log(1.0 + (double)orc_expf(-ay)));
        const float dy = (1.0f / (1.0f + orc_expf(-y)) - t) * inv_n;
        for (int o = 0; o < F; o++) {
            g_w2[o] = fmaf(dy, z[o] > 0.0f ? z[o] : 0.0f, g_w2[o]);
            dz[o] = z[o] <= 0.0f ? 0.0f : dy * w2[o];
            g_b1[o] = g_b1[o] + dz[o];
            for (int k = 0; k < IN; k++) g_w1[(size_t)o * IN + k] = fmaf(dz[o], X[k], g_w1[(size_t)o * IN + k]);
        }
        g_b2[0] = g_b2[0] + dy;
        for (int k = 0; k < IN; k++) {
            float acc = 0.0f;
            for (int o = 0; o < F; o++) acc = fmaf(dz[o], w1[(size_t)o * IN + k], acc);
            dX[k] = acc + (buf[k % E] - X[k]) * dy;                 /* DNN branch + FM branch (both read the same features) */
        }
        for (int s = 0; s < F; s++) {
            const int32_t c = s == 0 ? node[r] : seq[r * T + s - 1];
            if (c < 0) continue;
            for (int k = 0; k < E; k++) g_emb[(size_t)c * E + k] += dX[s * E + k];
        }
    }
    *loss_out = loss * inv_n;
    free(X); free(buf); free(z); free(dz); free(dX);
    return rc;
}
