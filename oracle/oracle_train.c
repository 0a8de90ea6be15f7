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
