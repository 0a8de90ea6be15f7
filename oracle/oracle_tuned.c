/*
 * oracle/oracle_tuned.c -- TEST / MEASUREMENT INFRASTRUCTURE, never linked into the product.
 *
 * The SECOND CPU form SURVEY.md 8(d) asks for next to the faithful port (oracle.c): what a tuned CPU implementation of
 * the same retrieval (tdm/src/main/scala/com/mass/tdm/model/Recommender.scala:40-107 over DIN.scala:14-43) would do when
 * it is free to re-associate the arithmetic -- so its logits are NOT bit-equal to the reference's, only close (the test
 * states the tolerance), and it is reported as a baseline ("port-tuned"), never used as a checker.
 *
 *   - node-side precomputation at load: y[node] = W1[:, :E] . emb[node] + b1, stored next to the row (one 2E record per node);
 *   - weight collapse: M = W1[:, E:] . Watt  (Attention.scala:32,46 followed by Linear(2E,E)), so per USER
 *     G_j = M . K_j for the T history rows, once;
 *   - per candidate row only: s = K x (T dot products), softmax over T, h = relu(y + sum_j p_j G_j), logit = w2 . h + b2
 *     -- 2TE + E fused multiply-adds instead of 2TE + 3E^2 + E;
 *   - polynomial expf, everything vectorised across E (AVX2 / AVX-512 clones picked at load time by the ifunc resolver);
 *   - beam cut by quickselect on (score, position) keys instead of a full sort;
 *   - users spread over threads exactly like Evaluator.scala:28-66.
 */
#define _GNU_SOURCE
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

#ifndef CLONES
#define CLONES __attribute__((target_clones("avx512f", "avx2", "default")))
#endif

struct orc_tuned {
    const orc_tree *t;
    int max_level; int64_t n_codes;
    const uint8_t *exists, *is_leaf; const int32_t *node_id;
    int E, T; int64_t rows;
    float scale, b2;
    float *rec;      /* rows x 2E : [emb row | W1x . emb row + b1] */
    float *Mt;       /* E x E, k-major: Mt[k][o] = sum_a W1[o][E+a] Watt[a][k] */
    float *w2;
};

typedef struct { orc_tuned *m; const float *emb, *w1, *b1; int64_t lo, hi; } prep_job;

CLONES static void prep_rows(orc_tuned *m, const float *emb, const float *w1t, const float *b1, int64_t lo, int64_t hi)
{
    const int E = m->E;
    for (int64_t r = lo; r < hi; r++) {
        const float *x = emb + r * E;
        float *o = m->rec + r * 2 * E;
        memcpy(o, x, sizeof(float) * E);
        float *y = o + E;
        for (int i = 0; i < E; i++) y[i] = b1[i];
        for (int k = 0; k < E; k++) {
            const float xk = x[k];
            const float *w = w1t + (size_t)k * E;
            for (int i = 0; i < E; i++) y[i] += xk * w[i];
        }
    }
}

typedef struct { orc_tuned *m; const float *emb, *w1t, *b1; int64_t lo, hi; } prep_arg;
static void *prep_worker(void *p)
{
    prep_arg *a = (prep_arg *)p;
    prep_rows(a->m, a->emb, a->w1t, a->b1, a->lo, a->hi);
    return NULL;
}

orc_tuned *orc_tuned_create(const orc_tree *t, const orc_tdm_model *model, int n_threads)
{
    int64_t rows; int E, T; const float *emb, *watt, *w1, *b1, *w2, *b2;
    if (orc_tdm_model_view(model, &rows, &E, &T, &emb, &watt, &w1, &b1, &w2, &b2)) return NULL;   /* DIN only */
    orc_tuned *m = (orc_tuned *)calloc(1, sizeof(*m));
    m->t = t;
    orc_tree_view(t, &m->max_level, &m->n_codes, &m->exists, &m->is_leaf, &m->node_id);
    m->E = E; m->T = T; m->rows = rows;
    m->scale = (float)(1.0 / sqrt((double)E));
    m->b2 = b2[0];
    m->rec = (float *)aligned_alloc(64, sizeof(float) * (size_t)rows * 2 * E);
    m->Mt = (float *)aligned_alloc(64, sizeof(float) * E * E);
    m->w2 = (float *)aligned_alloc(64, sizeof(float) * E);
    float *w1t = (float *)aligned_alloc(64, sizeof(float) * E * E);
    if (!m->rec || !m->Mt || !m->w2 || !w1t) { free(w1t); orc_tuned_destroy(m); return NULL; }
    memcpy(m->w2, w2, sizeof(float) * E);
    for (int o = 0; o < E; o++)
        for (int k = 0; k < E; k++) w1t[(size_t)k * E + o] = w1[(size_t)o * 2 * E + k];
    for (int o = 0; o < E; o++)
        for (int k = 0; k < E; k++) {
            double acc = 0.0;
            for (int a = 0; a < E; a++) acc += (double)w1[(size_t)o * 2 * E + E + a] * (double)watt[(size_t)a * E + k];
            m->Mt[(size_t)k * E + o] = (float)acc;
        }
    if (n_threads < 1) n_threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    prep_arg *args = (prep_arg *)malloc(sizeof(prep_arg) * n_threads);
    for (int i = 0; i < n_threads; i++) {
        args[i] = (prep_arg){ m, emb, w1t, b1, rows * i / n_threads, rows * (i + 1) / n_threads };
        pthread_create(&th[i], NULL, prep_worker, &args[i]);
    }
    for (int i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
    free(th); free(args); free(w1t);
    return m;
}

void orc_tuned_destroy(orc_tuned *m)
{
    if (!m) return;
    free(m->rec); free(m->Mt); free(m->w2); free(m);
}

/* One AVX-512 register of floats as a GCC vector type: the avx2 clone lowers it to two ymm, the default clone to four xmm. */
typedef float v16 __attribute__((vector_size(64)));
typedef float v16u __attribute__((vector_size(64), aligned(4)));
typedef int32_t v16i __attribute__((vector_size(64)));
#define TP 16   /* history positions padded to one vector; T <= TP */

/* exp(x) for x <= 0 (softmax arguments), ~2 ulp: 2^n * p(r), |r| <= ln2/2, Cephes' degree-5 polynomial. */
static inline __attribute__((always_inline)) v16 tuned_exp16(v16 x)
{
    const v16 lo = { -87.0f, -87.0f, -87.0f, -87.0f, -87.0f, -87.0f, -87.0f, -87.0f, -87.0f, -87.0f, -87.0f, -87.0f, -87.0f, -87.0f, -87.0f, -87.0f };
    x = (v16)(((v16i)x & (x > lo)) | ((v16i)lo & ~(x > lo)));
    const v16i ni = __builtin_convertvector(x * 1.44269504088896341f - 0.5f, v16i);      /* x <= 0: truncation rounds to nearest */
    const v16 n = __builtin_convertvector(ni, v16);
    const v16 r = (x - n * 0.693359375f) - n * -2.12194440e-4f;
    v16 p = r * 1.9875691500e-4f + 1.3981999507e-3f;
    p = p * r + 8.3334519073e-3f;
    p = p * r + 4.1665795894e-2f;
    p = p * r + 1.6666665459e-1f;
    p = p * r + 5.0000001201e-1f;
    p = p * r * r + r + 1.0f;
    return p * (v16)((ni + 127) << 23);
}

static inline __attribute__((always_inline)) v16 rot16(v16 v, int by)
{
    const v16i i8 = { 8, 9, 10, 11, 12, 13, 14, 15, 0, 1, 2, 3, 4, 5, 6, 7 }, i4 = { 4, 5, 6, 7, 0, 1, 2, 3, 12, 13, 14, 15, 8, 9, 10, 11 };
    const v16i i2 = { 2, 3, 0, 1, 6, 7, 4, 5, 10, 11, 8, 9, 14, 15, 12, 13 }, i1 = { 1, 0, 3, 2, 5, 4, 7, 6, 9, 8, 11, 10, 13, 12, 15, 14 };
    return __builtin_shuffle(v, by == 8 ? i8 : by == 4 ? i4 : by == 2 ? i2 : i1);
}
static inline __attribute__((always_inline)) v16 hsum16(v16 v)      /* every lane = the sum */
{
    v += rot16(v, 8); v += rot16(v, 4); v += rot16(v, 2); v += rot16(v, 1);
    return v;
}
static inline __attribute__((always_inline)) v16 hmax16(v16 v)
{
    v16 w;
    w = rot16(v, 8); v = (v16)(((v16i)v & (v > w)) | ((v16i)w & ~(v > w)));
    w = rot16(v, 4); v = (v16)(((v16i)v & (v > w)) | ((v16i)w & ~(v > w)));
    w = rot16(v, 2); v = (v16)(((v16i)v & (v > w)) | ((v16i)w & ~(v > w)));
    w = rot16(v, 1); v = (v16)(((v16i)v & (v > w)) | ((v16i)w & ~(v > w)));
    return v;
}

/* scores of n candidate codes for one user.  Kt: E x TP (history rows transposed, padded lanes zero), G: T x E (= M . K_j),
 * neg: 0 or -FLT_MAX per lane (masked position / padding lane). */
static inline __attribute__((always_inline)) void score_rows_e(const orc_tuned *m, const int E, int n, const int32_t *codes,
                                                               const float *restrict Kt, const float *restrict G,
                                                               const float *restrict neg, float *restrict out)
{
    const int T = m->T, EV = E / 16;
    const float scale = m->scale, b2 = m->b2;
    const v16 *restrict kt = (const v16 *)Kt;
    const v16 *restrict gv = (const v16 *)G;
    const v16 *restrict w2 = (const v16 *)m->w2;
    const v16 negv = *(const v16u *)neg;
    v16 live;
    for (int j = 0; j < TP; j++) live[j] = j < T ? 1.0f : 0.0f;
    for (int i = 0; i < n; i++) {
        const float *restrict x = m->rec + (size_t)codes[i] * 2 * E;
        const v16u *restrict y = (const v16u *)(x + E);
        if (i + 4 < n) {
            const char *nx = (const char *)(m->rec + (size_t)codes[i + 4] * 2 * E);
            for (int b = 0; b < 2 * E * 4; b += 64) __builtin_prefetch(nx + b);
        }
        /* s = K x with the T positions across the lanes: no horizontal reductions */
        v16 s0 = { 0 }, s1 = { 0 }, s2 = { 0 }, s3 = { 0 };
        for (int k = 0; k < E; k += 4) {
            s0 += x[k] * kt[k]; s1 += x[k + 1] * kt[k + 1]; s2 += x[k + 2] * kt[k + 2]; s3 += x[k + 3] * kt[k + 3];
        }
        v16 s = ((s0 + s1) + (s2 + s3)) * scale + negv;
        v16 e = tuned_exp16(s - hmax16(s)) * live;
        e = e / hsum16(e);
        v16 acc = { 0 };
        for (int v = 0; v < EV; v++) {
            v16 h = y[v];
            for (int j = 0; j < T; j++) h += e[j] * gv[(size_t)j * EV + v];
            h = (v16)((v16i)h & (h > 0.0f));
            acc += h * w2[v];
        }
        out[i] = hsum16(acc)[0] + b2;
    }
}

/* embed_size as a compile-time constant for the sizes the reference configures (16) and the bench uses (64) */
CLONES static void score_rows(const orc_tuned *m, int n, const int32_t *codes, const float *restrict Kt,
                              const float *restrict G, const float *restrict neg, float *restrict out)
{
    switch (m->E) {
    case 64: score_rows_e(m, 64, n, codes, Kt, G, neg, out); break;
    case 32: score_rows_e(m, 32, n, codes, Kt, G, neg, out); break;
    case 16: score_rows_e(m, 16, n, codes, Kt, G, neg, out); break;
    default: score_rows_e(m, m->E, n, codes, Kt, G, neg, out); break;            /* any multiple of 16 */
    }
}

CLONES static void user_prologue(const orc_tuned *m, const int32_t *hist, float *restrict K, float *restrict G, float *restrict Kt)
{
    const int E = m->E, T = m->T;
    memset(Kt, 0, sizeof(float) * E * TP);
    for (int j = 0; j < T; j++) {
        float *kj = K + (size_t)j * E, *gj = G + (size_t)j * E;
        if (hist[j] < 0) { memset(kj, 0, sizeof(float) * E); memset(gj, 0, sizeof(float) * E); continue; }
        memcpy(kj, m->rec + (size_t)hist[j] * 2 * E, sizeof(float) * E);
        for (int o = 0; o < E; o++) gj[o] = 0.0f;
        for (int k = 0; k < E; k++) {
            const float xk = kj[k];
            const float *restrict w = m->Mt + (size_t)k * E;
            for (int o = 0; o < E; o++) gj[o] += xk * w[o];
        }
        for (int k = 0; k < E; k++) Kt[(size_t)k * TP + j] = kj[k];
    }
}

static inline uint64_t tuned_key(float f, int pos)
{   /* larger key = better: score in Float.compare order, then EARLIER position (the stable sort of Recommender.scala:75-87) */
    union { float f; uint32_t u; } c; c.f = f;
    uint32_t u = c.u;
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((uint64_t)u << 32) | (uint32_t)(0x7fffffff - pos);
}

/* partition so that the `keep` largest keys come first (Hoare quickselect, median of three) */
static void select_top(uint64_t *a, int n, int keep)
{
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        uint64_t x = a[lo], y = a[(lo + hi) >> 1], z = a[hi];
        uint64_t piv = x > y ? (y > z ? y : (x > z ? z : x)) : (x > z ? x : (y > z ? z : y));
        int i = lo, j = hi;
        while (i <= j) {
            while (a[i] > piv) i++;
            while (a[j] < piv) j--;
            if (i <= j) { uint64_t tmp = a[i]; a[i] = a[j]; a[j] = tmp; i++; j--; }
        }
        if (keep - 1 <= j) hi = j;
        else if (keep - 1 >= i) lo = i;
        else break;
    }
}

typedef struct {
    const orc_tuned *m; const int32_t *seq; int B, beam, topk, use_mask, lo, hi, rc;
    int32_t *out_items; float *out_logits; int32_t *out_counts;
} tuned_job;

static int tuned_user(const orc_tuned *m, const int32_t *seq_ids, int beam, int topk, int use_mask,
                      int32_t *out_items, float *out_logits, float *K, float *G, float *Kt, int32_t *cand, float *pred,
                      int32_t *next, uint64_t *keys, int32_t *lcode, float *lpred, int lcap)
{
    const int T = m->T;
    int32_t hist[64]; uint8_t masked[64];
    orc_tdm_id_to_code(m->t, T, seq_ids, hist, masked);
    for (int j = 0; j < T; j++) if (hist[j] != -1 && (hist[j] < 0 || hist[j] >= m->rows)) return -2;
    if (!use_mask) memset(masked, 0, T);
    user_prologue(m, hist, K, G, Kt);
    float neg[TP];
    for (int j = 0; j < TP; j++) neg[j] = (j >= T || masked[j]) ? -FLT_MAX : 0.0f;
    int level = (int)floor(log((double)beam) / log(2.0));          /* getLevelStart, Recommender.scala:210-216 */
    int64_t start = ((int64_t)1 << level) - 1, end = 2 * start + 1;
    int ncand = 0, nleaf = 0;
    if (level <= m->max_level)
        for (int64_t c = start; c < end && c < m->n_codes; c++)
            if (m->exists[c]) { cand[ncand] = (int32_t)c; pred[ncand] = 0.0f; ncand++; }
    for (int it = level; it <= m->max_level && ncand; it++) {
        int nnl = 0;
        for (int i = 0; i < ncand; i++) {
            if (m->is_leaf[cand[i]]) { if (nleaf >= lcap) return -3; lcode[nleaf] = cand[i]; lpred[nleaf] = pred[i]; nleaf++; }
            else { keys[nnl] = tuned_key(pred[i], nnl); next[nnl] = cand[i]; nnl++; }
        }
        int nb = nnl;
        if (nnl > beam) {
            select_top(keys, nnl, beam);
            nb = beam;
            for (int i = 0; i < nb; i++) cand[i] = next[0x7fffffff - (int)(keys[i] & 0x7fffffff)];
        } else {
            for (int i = 0; i < nb; i++) cand[i] = next[i];
        }
        int nc = 0;
        for (int i = 0; i < nb; i++)
            for (int s = 1; s <= 2; s++) {
                int64_t c = 2 * (int64_t)cand[i] + s;
                if (c < m->n_codes && m->exists[c]) next[nc++] = (int32_t)c;
            }
        score_rows(m, nc, next, Kt, G, neg, pred);
        memcpy(cand, next, sizeof(int32_t) * nc);
        ncand = nc;
    }
    /* top-k of the leaves */
    for (int i = 0; i < nleaf; i++) keys[i] = tuned_key(lpred[i], i);
    int k = nleaf < topk ? nleaf : topk;
    if (nleaf > k) select_top(keys, nleaf, k);
    for (int i = 1; i < k; i++) {                           /* k is small: insertion sort, descending */
        uint64_t v = keys[i]; int j = i - 1;
        while (j >= 0 && keys[j] < v) { keys[j + 1] = keys[j]; j--; }
        keys[j + 1] = v;
    }
    for (int i = 0; i < k; i++) {
        int p = 0x7fffffff - (int)(keys[i] & 0x7fffffff);
        out_items[i] = m->node_id[lcode[p]];
        out_logits[i] = lpred[p];
    }
    return k;
}

static void *tuned_worker(void *p)
{
    tuned_job *j = (tuned_job *)p;
    const orc_tuned *m = j->m;
    const int E = m->E, T = m->T, cap = 2 * j->beam + 2, lcap = 2 * j->beam * (m->max_level + 2) + 8;
    float *K = (float *)aligned_alloc(64, sizeof(float) * T * E), *G = (float *)aligned_alloc(64, sizeof(float) * T * E);
    float *Kt = (float *)aligned_alloc(64, sizeof(float) * E * TP);
    int32_t *cand = (int32_t *)malloc(sizeof(int32_t) * cap), *next = (int32_t *)malloc(sizeof(int32_t) * cap);
    float *pred = (float *)malloc(sizeof(float) * cap);
    uint64_t *keys = (uint64_t *)malloc(sizeof(uint64_t) * (lcap > cap ? lcap : cap));
    int32_t *lcode = (int32_t *)malloc(sizeof(int32_t) * lcap);
    float *lpred = (float *)malloc(sizeof(float) * lcap);
    for (int u = j->lo; u < j->hi; u++) {
        int32_t *oi = j->out_items + (size_t)u * j->topk; float *ol = j->out_logits + (size_t)u * j->topk;
        for (int i = 0; i < j->topk; i++) { oi[i] = -1; ol[i] = 0.0f; }
        int n = tuned_user(m, j->seq + (size_t)u * T, j->beam, j->topk, j->use_mask, oi, ol, K, G, Kt, cand, pred, next, keys,
                           lcode, lpred, lcap);
        if (n < 0) { j->rc = n; n = 0; }
        j->out_counts[u] = n;
    }
    free(K); free(G); free(Kt); free(cand); free(next); free(pred); free(keys); free(lcode); free(lpred);
    return NULL;
}

int orc_tuned_retrieve_batch(const orc_tuned *m, int B, const int32_t *seq_ids, int beam, int topk, int use_mask,
                             int n_threads, int32_t *out_items, float *out_logits, int32_t *out_counts)
{
    if (m->E > 256 || (m->E & 15) || m->T > TP || beam < 1) return -1;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > B) n_threads = B > 0 ? B : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    tuned_job *jobs = (tuned_job *)malloc(sizeof(tuned_job) * n_threads);
    int task = B / n_threads, extra = B % n_threads, rc = 0;
    for (int i = 0; i < n_threads; i++) {
        jobs[i] = (tuned_job){ m, seq_ids, B, beam, topk, use_mask, 0, 0, 0, out_items, out_logits, out_counts };
        jobs[i].lo = i * task + (i < extra ? i : extra);
        jobs[i].hi = jobs[i].lo + task + (i < extra ? 1 : 0);
        pthread_create(&th[i], NULL, tuned_worker, &jobs[i]);
    }
    for (int i = 0; i < n_threads; i++) { pthread_join(th[i], NULL); if (jobs[i].rc) rc = jobs[i].rc; }
    free(th); free(jobs);
    return rc;
}
