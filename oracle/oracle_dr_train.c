/*
 * oracle/oracle_dr_train.c -- CPU ORACLE, Deep Retrieval training step.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Restates one mini-batch iteration of deep-retrieval/src/main/scala/com/mass/dr/optim/LocalOptimizer.scala:58-194:
 *   layer model   MiniBatch.transformLayerData (dataset/MiniBatch.scala:19-50) -> LayerModel graph forward (model/LayerModel.scala:22-39:
 *                 EmbeddingShare -> Reshape -> Linear per layer) -> CrossEntropyLayer (loss/CrossEntropyLayer.scala:13-24 =
 *                 scalann CrossEntropyCriterion.scala:15-27 = LogSoftMax.scala:36-67 + ClassNLLCriterion.scala:17-104, sizeAverage)
 *                 -> backward -> syncGradients over `parallelism` thread chunks (LocalOptimizer.scala:139-187)
 *   rerank model  MiniBatch.transformRerankData (:52-61) -> RerankModel graph (model/RerankModel.scala:20-36: Embedding ->
 *                 Reshape -> Linear) -> SampledSoftmaxLoss (scalann/.../nn/SampledSoftmaxLoss.scala:49-153, batchMode = false)
 *                 -> its own Adam over the softmax weights / biases (nn/mixin/ParameterOptimizer.scala:28-88) -> model backward
 * Arithmetic spec as in oracle.c: every GEMM / dot element is ONE chain over ascending k (acc = fma(a, b, acc) from 0), everything
 * else a single IEEE operation in the order the Scala code issues it; MKL's own order is unobservable (oracle.c header: parity
 * unpinned against the JVM).  Pinned by the reference's known-answer tests CrossEntropyTest.scala:26-43 and the property of
 * SampledSoftmaxLossTest.scala:42-52 (tests/test_oracle_known_answers.py) and by float64 finite differences.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "orc_math.h"
#include "oracle.h"

/* CrossEntropyCriterion on [R, C] logits with integer targets; weight = 1 / R (sizeAverage).  Returns the loss and, if
 * grad != NULL, writes gradInput (LogSoftMax.updateGradInput on ClassNLLCriterion.updateGradInput). */
double orc_cross_entropy_f64(int64_t R, int C, const double *logits, const int32_t *target, double *grad)
{
    double out_loss = 0.0;
    double *buf = (double *)malloc(sizeof(double) * (size_t)C);
    for (int64_t r = 0; r < R; r++) {
        const double *in = logits + r * C;
        double mx = in[0];
        for (int j = 1; j < C; j++) mx = in[j] > mx ? in[j] : mx;                 /* in.max() */
        double sum = 0.0;
        for (int j = 0; j < C; j++) { buf[j] = orc_exp(-mx + in[j]); sum = fma(buf[j], 1.0, sum); }   /* fill(-max).add(in).exp(); dot(ones) */
        const double log_sum = mx + log(sum);
        const int t = target[r];
        out_loss = out_loss - (in[t] + -log_sum);                                  /* out.add(-logSum); output -= out[target] */
        if (grad) {
            double *g = grad + r * C;
            const double go_t = -1.0 / (double)R;                                  /* ClassNLL: -1 then / batchSize */
            const double out_sum = go_t;                                           /* gradOut.dot(ones): one non-zero */
            for (int j = 0; j < C; j++) {
                const double e = orc_exp(in[j] + -log_sum);                        /* buffer.exp(out) */
                g[j] = fma(-out_sum, e, j == t ? go_t : 0.0);                      /* gradOut.add(-outSum, buffer) */
            }
        }
    }
    free(buf);
    return out_loss / (double)R;
}

/* rows of one chunk: sample s in [s0, s1), path p < P; idx = seq ++ (path[i] + numItem + i K), i < D - 1 */
static void layer_chunk(int num_item, int K, int D, int T, int E, const double *emb, const double *const *w, const double *const *b,
                        int s0, int s1, const int32_t *seq, const int32_t *target, const int32_t *item_paths, int P,
                        double *g_emb, double *const *g_w, double *const *g_b, double *loss)
{
    const int64_t R = (int64_t)(s1 - s0) * P;
    const int W = T + D - 1, IN = W * E;
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)(R * W));
    int32_t *tg = (int32_t *)malloc(sizeof(int32_t) * (size_t)(R * D));
    double *X = (double *)calloc((size_t)(R * IN), sizeof(double));
    double *GX = (double *)calloc((size_t)(R * IN), sizeof(double));
    double *lg = (double *)malloc(sizeof(double) * (size_t)(R * K));
    double *dl = (double *)malloc(sizeof(double) * (size_t)(R * K));
    for (int s = s0; s < s1; s++)
        for (int p = 0; p < P; p++) {
            const int64_t r = (int64_t)(s - s0) * P + p;
            const int32_t *path = item_paths + ((int64_t)target[s] * P + p) * D;
            for (int j = 0; j < T; j++) idx[r * W + j] = seq[(int64_t)s * T + j];
            for (int i = 0; i < D - 1; i++) idx[r * W + T + i] = path[i] + num_item + i * K;
            for (int d = 0; d < D; d++) tg[r * D + d] = path[d];
            for (int j = 0; j < W; j++) {
                const int32_t c = idx[r * W + j];
                if (c >= 0) memcpy(X + r * IN + (int64_t)j * E, emb + (int64_t)c * E, sizeof(double) * (size_t)E);   /* paddingIdx -> zero row */
            }
        }
    int32_t *tcol = (int32_t *)malloc(sizeof(int32_t) * (size_t)R);
    for (int d = 0; d < D; d++) {
        const int in = (T + d) * E;
        for (int64_t r = 0; r < R; r++) {
            for (int k = 0; k < K; k++) {
                double acc = 0.0;
                const double *wr = w[d] + (int64_t)k * in, *x = X + r * IN;
                for (int i = 0; i < in; i++) acc = fma(x[i], wr[i], acc);
                lg[r * K + k] = acc + b[d][k];                                     /* Linear: addmm then add bias */
            }
            tcol[r] = tg[r * D + d];
        }
        loss[d] = orc_cross_entropy_f64(R, K, lg, tcol, dl);
        for (int k = 0; k < K; k++) {                                              /* gradWeight = gradOutput^T . input, gradBias = column sums */
            for (int i = 0; i < in; i++) {
                double acc = 0.0;
                for (int64_t r = 0; r < R; r++) acc = fma(dl[r * K + k], X[r * IN + i], acc);
                g_w[d][(int64_t)k * in + i] += acc;
            }
            double acc = 0.0;
            for (int64_t r = 0; r < R; r++) acc = fma(dl[r * K + k], 1.0, acc);
            g_b[d][k] += acc;
        }
        for (int64_t r = 0; r < R; r++)                                            /* gradInput = gradOutput . W */
            for (int i = 0; i < in; i++) {
                double acc = 0.0;
                for (int k = 0; k < K; k++) acc = fma(dl[r * K + k], w[d][(int64_t)k * in + i], acc);
                GX[r * IN + i] += acc;
            }
    }
    for (int64_t r = 0; r < R; r++)                                                /* EmbeddingShare backward: scatter-add, padding skipped */
        for (int j = 0; j < W; j++) {
            const int32_t c = idx[r * W + j];
            if (c < 0) continue;
            for (int e = 0; e < E; e++) g_emb[(int64_t)c * E + e] += GX[r * IN + (int64_t)j * E + e];
        }
    free(idx); free(tg); free(X); free(GX); free(lg); free(dl); free(tcol);
}

/* trainLayerBatch + syncGradients: gradients of the layer model for one mini-batch, zeroed here; loss[D] = mean over chunks */
int orc_dr_layer_grad(int num_item, int K, int D, int T, int E, const double *emb, const double *const *w, const double *const *b,
                      int n, const int32_t *seq, const int32_t *target, const int32_t *item_paths, int P, int parallelism,
                      double *g_emb, double *const *g_w, double *const *g_b, double *loss)
{
    const int64_t emb_n = ((int64_t)num_item + (int64_t)K * (D - 1)) * E;
    for (int64_t i = 0; i < (int64_t)n * T; i++)
        if (seq[i] < -1 || seq[i] >= num_item + K * (D - 1)) return -2;
    for (int i = 0; i < n; i++)
        if (target[i] < 0 || target[i] >= num_item) return -2;
    const int task = n / parallelism, extra = n % parallelism;
    const int par = task == 0 ? extra : parallelism;                               /* LocalOptimizer.scala:146-148 */
    memset(g_emb, 0, sizeof(double) * (size_t)emb_n);
    for (int d = 0; d < D; d++) {
        memset(g_w[d], 0, sizeof(double) * (size_t)K * (size_t)((T + d) * E));
        memset(g_b[d], 0, sizeof(double) * (size_t)K);
        loss[d] = 0.0;
    }
    double *closs = (double *)malloc(sizeof(double) * (size_t)D);
    for (int c = 0; c < par; c++) {                                                /* gradient buffers summed in chunk order (:170-181) */
        const int off = c * task + (c < extra ? c : extra), len = task + (c < extra ? 1 : 0);
        layer_chunk(num_item, K, D, T, E, emb, w, b, off, off + len, seq, target, item_paths, P, g_emb, g_w, g_b, closs);
        for (int d = 0; d < D; d++) loss[d] += closs[d];
    }
    free(closs);
    if (par > 1) {                                                                 /* totalLayerGradients.div(syncNum) */
        for (int64_t i = 0; i < emb_n; i++) g_emb[i] /= (double)par;
        for (int d = 0; d < D; d++) {
            const int64_t nw = (int64_t)K * (T + d) * E;
            for (int64_t i = 0; i < nw; i++) g_w[d][i] /= (double)par;
            for (int k = 0; k < K; k++) g_b[d][k] /= (double)par;
        }
    }
    for (int d = 0; d < D; d++) loss[d] /= (double)par;                            /* lossSum.map(_ / parallelism) */
    return 0;
}

/* SampledSoftmaxLoss.updateOutput + backward on user vectors u[n][E]: loss, gradInput gu[n][E], and the parameter gradients
 * ACCUMULATED into g_sm_w / g_sm_b (ParameterOptimizer.computeParameterGrad, :67-88: axpy into gradWeights, which nothing in the
 * reference ever zeroes).  sampled[n][S + 1]: positive first. */
double orc_sampled_softmax_f64(int n, int E, int S, const double *u, const double *sm_w, const double *sm_b, const int32_t *sampled,
                               double *gu, double *g_sm_w, double *g_sm_b)
{
    const int C = S + 1;
    double *lg = (double *)malloc(sizeof(double) * (size_t)n * C), *dl = (double *)malloc(sizeof(double) * (size_t)n * C);
    int32_t *zero = (int32_t *)calloc((size_t)n, sizeof(int32_t));               /* labelPosition: the positive is in the first place */
    for (int i = 0; i < n; i++)
        for (int j = 0; j < C; j++) {
            const int32_t it = sampled[(int64_t)i * C + j];
            double acc = 0.0;
            for (int e = 0; e < E; e++) acc = fma(sm_w[(int64_t)it * E + e], u[(int64_t)i * E + e], acc);   /* out.addmv(w, vec) */
            lg[(int64_t)i * C + j] = acc + sm_b[it];                               /* out.add(b) */
        }
    const double loss = orc_cross_entropy_f64(n, C, lg, zero, dl);
    for (int i = 0; i < n; i++)                                                    /* linearBackward: grad = w^T . logitGrad */
        for (int e = 0; e < E; e++) {
            double acc = 0.0;
            for (int j = 0; j < C; j++) acc = fma(sm_w[(int64_t)sampled[(int64_t)i * C + j] * E + e], dl[(int64_t)i * C + j], acc);
            gu[(int64_t)i * E + e] = acc;
        }
    for (int i = 0; i < n; i++)                                                    /* computeParameterGradInput + computeParameterGrad */
        for (int j = 0; j < C; j++) {
            const int32_t it = sampled[(int64_t)i * C + j];
            const double g = dl[(int64_t)i * C + j];
            for (int e = 0; e < E; e++) g_sm_w[(int64_t)it * E + e] += fma(g, u[(int64_t)i * E + e], 0.0);
            g_sm_b[it] = g + g_sm_b[it];
        }
    free(lg); free(dl); free(zero);
    return loss;
}

/* trainRerank (LocalOptimizer.scala:122-137): model gradients zeroed here, softmax-parameter gradients accumulated (see above).
 * rr_w is [E][T E] row-major.  The softmax parameters are NOT updated here (orc_adam_eps_f64 does that, in the reference's order:
 * after gradInput has been computed from the old weights). */
int orc_dr_rerank_grad(int num_item, int T, int E, const double *rr_emb, const double *rr_w, const double *rr_b, const double *sm_w,
                       const double *sm_b, int n, const int32_t *seq, const int32_t *sampled, int S, double *g_rr_emb, double *g_rr_w,
                       double *g_rr_b, double *g_sm_w, double *g_sm_b, double *loss)
{
    const int IN = T * E, C = S + 1;
    for (int64_t i = 0; i < (int64_t)n * T; i++)
        if (seq[i] < -1 || seq[i] >= num_item) return -2;
    for (int64_t i = 0; i < (int64_t)n * C; i++)
        if (sampled[i] < 0 || sampled[i] >= num_item) return -2;
    double *X = (double *)calloc((size_t)n * IN, sizeof(double)), *u = (double *)malloc(sizeof(double) * (size_t)n * E);
    double *gu = (double *)malloc(sizeof(double) * (size_t)n * E);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < T; j++) {
            const int32_t c = seq[(int64_t)i * T + j];
            if (c >= 0) memcpy(X + (int64_t)i * IN + (int64_t)j * E, rr_emb + (int64_t)c * E, sizeof(double) * (size_t)E);
        }
    for (int i = 0; i < n; i++)
        for (int o = 0; o < E; o++) {
            double acc = 0.0;
            for (int k = 0; k < IN; k++) acc = fma(X[(int64_t)i * IN + k], rr_w[(int64_t)o * IN + k], acc);
            u[(int64_t)i * E + o] = acc + rr_b[o];
        }
    *loss = orc_sampled_softmax_f64(n, E, S, u, sm_w, sm_b, sampled, gu, g_sm_w, g_sm_b);
    memset(g_rr_emb, 0, sizeof(double) * (size_t)num_item * E);
    for (int o = 0; o < E; o++) {
        for (int k = 0; k < IN; k++) {
            double acc = 0.0;
            for (int i = 0; i < n; i++) acc = fma(gu[(int64_t)i * E + o], X[(int64_t)i * IN + k], acc);
            g_rr_w[(int64_t)o * IN + k] = acc;
        }
        double acc = 0.0;
        for (int i = 0; i < n; i++) acc = fma(gu[(int64_t)i * E + o], 1.0, acc);
        g_rr_b[o] = acc;
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < T; j++) {
            const int32_t c = seq[(int64_t)i * T + j];
            if (c < 0) continue;
            for (int e = 0; e < E; e++) {
                double acc = 0.0;
                for (int o = 0; o < E; o++) acc = fma(gu[(int64_t)i * E + o], rr_w[(int64_t)o * IN + (int64_t)j * E + e], acc);
                g_rr_emb[(int64_t)c * E + e] += acc;
            }
        }
    free(X); free(u); free(gu);
    return 0;
}

/* Adam with an explicit epsilon and without touching the gradient: Adam.optimize (optim/Adam.scala:19-73, eps 1e-8) and
 * ParameterOptimizer.optimize (nn/mixin/ParameterOptimizer.scala:38-65, eps 1e-7) are the same sequence of tensor operations. */
void orc_adam_eps_f64(double *w, const double *g, double *s, double *r, int64_t n, double lr, double eps, int t)
{
    const double beta1 = 0.9, beta2 = 0.999;
    const double step = lr * sqrt(1 - pow(beta2, t)) / (1 - pow(beta1, t));
    for (int64_t i = 0; i < n; i++) {
        s[i] = s[i] * beta1 + (1 - beta1) * g[i];
        r[i] = r[i] * beta2 + (1 - beta2) * (g[i] * g[i]);
        const double denom = sqrt(r[i]) + eps;
        w[i] = w[i] + -step * (s[i] / denom);
    }
}
