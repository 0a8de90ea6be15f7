/*
 * oracle/oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (dismember_b200/) never
 * links, imports or executes anything under oracle/.
 *
 * What it is: a plain-C restatement of massquantity/dismember's retrieval hot
 * path, function by function (citations inline, relative to /root/reference).
 *
 * PARITY STATUS: "parity unpinned" against the real JVM reference.  The
 * reference is Scala + Intel MKL over JNI; no JVM/sbt exists in the build
 * image and the MKL JNI jar is not vendored, so the reference cannot be run or
 * compiled here.  Upstream holds NO golden vector for beam
 * search, the DIN scorer or sgemm; its only known-answer tests are
 * scalann/src/test/scala/SoftMaxTest.scala:8-27 and CrossEntropyTest.scala,
 * which tests/test_oracle_known_answers.py replays against this file, and its
 * bundled fixtures (data/jtm, data/otm, data/dr) supply the real trees and
 * trained weights the golden vectors in tests/golden/ are computed on.
 * Where MKL's order of floating-point operations is unknowable the oracle
 * fixes one (orc_math.h) and the CUDA product reproduces it bit for bit.
 */
#define _GNU_SOURCE
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "orc_math.h"
#include "oracle.h"

/* ------------------------------------------------------------------ DIN -- */
#define REAL float
#define SUF(x) x##_f32
#define FMA fmaf
#define EXP orc_expf
#include "orc_din.inc"
#undef REAL
#undef SUF
#undef FMA
#undef EXP

#define REAL double
#define SUF(x) x##_f64
#define FMA fma
#define EXP orc_exp
#include "orc_din.inc"
#undef REAL
#undef SUF
#undef FMA
#undef EXP

/* ----------------------------------------------------------------- tree -- */
/* DistTree.loadItems / loadData  (tdm/src/main/scala/com/mass/tdm/tree/DistTree.scala:25-87):
 *   codeNodeMap : code -> Node{id, probality, is_leaf}
 *   idCodeMap   : leaf item id -> code          (from the Part_* lists)
 *   nonLeafOffset = max(leaf id) + 1 ; maxCode = max(leaf code) ; maxLevel = meta */
struct orc_tree {
    int max_level;
    int64_t n_codes;            /* 2^(max_level+1) - 1 */
    uint8_t *exists;            /* codeNodeMap.contains(code) */
    uint8_t *is_leaf;
    int32_t *node_id;           /* codeNodeMap(code).id */
    int32_t non_leaf_offset;
    int32_t max_code;
    int32_t *id_code;           /* [0, non_leaf_offset) -> code or -1 */
};

orc_tree *orc_tree_create(int max_level, int64_t n_nodes, const int32_t *codes,
                          const int32_t *node_ids, const uint8_t *is_leaf,
                          int64_t n_items, const int32_t *leaf_ids, const int32_t *leaf_codes)
{
    orc_tree *t = (orc_tree *)calloc(1, sizeof(*t));
    t->max_level = max_level;
    t->n_codes = ((int64_t)1 << (max_level + 1)) - 1;
    t->exists = (uint8_t *)calloc((size_t)t->n_codes, 1);
    t->is_leaf = (uint8_t *)calloc((size_t)t->n_codes, 1);
    t->node_id = (int32_t *)calloc((size_t)t->n_codes, sizeof(int32_t));
    for (int64_t i = 0; i < n_nodes; i++) {
        int64_t c = codes[i];
        if (c < 0 || c >= t->n_codes) { orc_tree_destroy(t); return NULL; }
        t->exists[c] = 1; t->is_leaf[c] = is_leaf[i] ? 1 : 0; t->node_id[c] = node_ids[i];
    }
    int32_t mx_id = -1, mx_code = -1;
    for (int64_t i = 0; i < n_items; i++) {
        if (leaf_ids[i] > mx_id) mx_id = leaf_ids[i];
        if (leaf_codes[i] > mx_code) mx_code = leaf_codes[i];
    }
    t->non_leaf_offset = mx_id + 1;
    t->max_code = mx_code;
    t->id_code = (int32_t *)malloc(sizeof(int32_t) * (size_t)(t->non_leaf_offset > 0 ? t->non_leaf_offset : 1));
    for (int32_t i = 0; i < t->non_leaf_offset; i++) t->id_code[i] = -1;
    for (int64_t i = 0; i < n_items; i++)
        if (leaf_ids[i] >= 0) t->id_code[leaf_ids[i]] = leaf_codes[i];   /* later entries win, like ++= */
    return t;
}

void orc_tree_destroy(orc_tree *t)
{
    if (!t) return;
    free(t->exists); free(t->is_leaf); free(t->node_id); free(t->id_code); free(t);
}

/* TDMTree.idToCode  (tdm/src/main/scala/com/mass/tdm/tree/TDMTree.scala:35-56).
 * NB a faithful quirk: an id that is neither padding nor a known leaf goes
 * through `id - nonLeafOffset`; only results > maxCode are masked, so -1 yields
 * an UNMASKED zero row and other negatives an invalid index. */
void orc_tdm_id_to_code(const orc_tree *t, int T, const int32_t *ids, int32_t *codes, uint8_t *masked)
{
    for (int i = 0; i < T; i++) {
        int32_t id = ids[i];
        masked[i] = 0;
        if (id == 0) { masked[i] = 1; codes[i] = -1; }
        else if (id < t->non_leaf_offset && id >= 0 && t->id_code[id] >= 0) codes[i] = t->id_code[id];
        else {
            int64_t tmp = (int64_t)id - t->non_leaf_offset;
            if (tmp > t->max_code) { masked[i] = 1; codes[i] = -1; }
            else codes[i] = (int32_t)tmp;
        }
    }
}

/* Recommender.getLevelStart  (tdm/.../model/Recommender.scala:210-216):
 * level = floor(log(n)/log(2)) in Double; start code = 2^level - 1. */
static int orc_lower_log2(int n)
{
    return (int)floor(log((double)n) / log(2.0));
}

/* stable descending sort of indices by key (java.util.Arrays.sort on objects
 * = TimSort = stable; any stable sort gives the same permutation). */
typedef struct { uint64_t key; int32_t pos; } orc_kp;

/* bottom-up merge sort, ties keep the left (earlier) element first */
static void orc_sort_desc(orc_kp *a, int n)
{
    if (n < 2) return;
    orc_kp *buf = (orc_kp *)malloc(sizeof(orc_kp) * (size_t)n);
    orc_kp *src = a, *dst = buf;
    for (int w = 1; w < n; w *= 2) {
        for (int lo = 0; lo < n; lo += 2 * w) {
            int mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            int i = lo, j = mid, k = lo;
            while (i < mid && j < hi) dst[k++] = (src[j].key > src[i].key) ? src[j++] : src[i++];
            while (i < mid) dst[k++] = src[i++];
            while (j < hi) dst[k++] = src[j++];
        }
        orc_kp *s = src; src = dst; dst = s;
    }
    if (src != a) memcpy(a, src, sizeof(orc_kp) * (size_t)n);
    free(buf);
}

/* ------------------------------------------------------------ TDM / JTM -- */
/* DeepFM scorer, the other `model.deep_model` of the TDM/JTM tasks
 * (tdm/src/main/scala/com/mass/tdm/model/DeepFM.scala:11-44): features F = [item row ; T history rows]
 * ((T+1) x E, padding rows zero, NO mask);
 *   FM   (scalann/.../nn/FM.scala:14-44): buffer = F_0 + F_1 + ... (vAdd in row order, from zero),
 *         (dot(buffer, buffer) - dot(Fflat, Fflat)) / 2
 *   DNN : Linear((T+1)E, T+1) -> ReLU -> Linear(T+1, 1) on Fflat  (Linear.scala:19-56: bias after the product)
 *   Add  (nn/Add.scala): fm + dnn.
 * Compact parameter vector (Graph.scala:37-48, topological order): [emb rows*E | W1 (T+1) x (T+1)E | b1 T+1 | W2 T+1 | b2 1]. */
typedef struct { const float *w1, *b1, *w2, *b2; } orc_dfm_f32;
struct orc_tdm_model { orc_din_f32 din; int deepfm; orc_dfm_f32 dfm; };

static float orc_deepfm_row_f32(const orc_tdm_model *m, const float *q, const float *K, float *scratch)
{
    const int E = m->din.E, T = m->din.T, F = T + 1;
    float *buf = scratch;                                           /* E */
    for (int k = 0; k < E; k++) buf[k] = 0.0f;
    for (int k = 0; k < E; k++) buf[k] = buf[k] + q[k];             /* vAdd, feature 0 = the item */
    for (int j = 0; j < T; j++)
        for (int k = 0; k < E; k++) buf[k] = buf[k] + K[(size_t)j * E + k];
    float sum_square = 0.0f, square_sum = 0.0f;
    for (int k = 0; k < E; k++) sum_square = fmaf(buf[k], buf[k], sum_square);
    for (int k = 0; k < E; k++) square_sum = fmaf(q[k], q[k], square_sum);
    for (int k = 0; k < T * E; k++) square_sum = fmaf(K[k], K[k], square_sum);
    const float fm = (sum_square - square_sum) / 2.0f;
    float dnn = 0.0f;
    for (int o = 0; o < F; o++) {
        const float *w = m->dfm.w1 + (size_t)o * F * E;
        float acc = 0.0f;
        for (int k = 0; k < E; k++) acc = fmaf(q[k], w[k], acc);
        for (int k = 0; k < T * E; k++) acc = fmaf(K[k], w[E + k], acc);
        float h = acc + m->dfm.b1[o];
        h = h > 0.0f ? h : (h != h ? h : 0.0f);                     /* ReLU.scala:30-44: math.max(x, 0), NaN propagates */
        dnn = fmaf(h, m->dfm.w2[o], dnn);
    }
    dnn = dnn + m->dfm.b2[0];
    return fm + dnn;
}

orc_tdm_model *orc_tdm_model_create(int64_t rows, int E, int T, const float *params)
{
    orc_tdm_model *m = (orc_tdm_model *)calloc(1, sizeof(*m));
    const float *emb = params, *watt = emb + rows * E, *w1 = watt + (int64_t)E * E;
    const float *b1 = w1 + (int64_t)2 * E * E, *w2 = b1 + E, *b2 = w2 + E;
    if (orc_din_init_f32(&m->din, rows, E, T, emb, watt, w1, b1, w2, b2)) { free(m); return NULL; }
    return m;
}
orc_tdm_model *orc_tdm_deepfm_create(int64_t rows, int E, int T, const float *params)
{
    orc_tdm_model *m = (orc_tdm_model *)calloc(1, sizeof(*m));
    const int F = T + 1;
    m->deepfm = 1;
    m->din.rows = rows; m->din.E = E; m->din.T = T; m->din.emb = params;      /* gather + index checks reuse the DIN struct */
    m->dfm.w1 = params + rows * E;
    m->dfm.b1 = m->dfm.w1 + (int64_t)F * F * E;
    m->dfm.w2 = m->dfm.b1 + F;
    m->dfm.b2 = m->dfm.w2 + F;
    return m;
}
/* read-only views for oracle_tuned.c (the second CPU form of SURVEY 8(d)) */
void orc_tree_view(const orc_tree *t, int *max_level, int64_t *n_codes, const uint8_t **exists, const uint8_t **is_leaf,
                   const int32_t **node_id)
{
    *max_level = t->max_level; *n_codes = t->n_codes; *exists = t->exists; *is_leaf = t->is_leaf; *node_id = t->node_id;
}
int orc_tdm_model_view(const orc_tdm_model *m, int64_t *rows, int *E, int *T, const float **emb, const float **watt,
                       const float **w1, const float **b1, const float **w2, const float **b2)
{
    if (m->deepfm) return -1;
    *rows = m->din.rows; *E = m->din.E; *T = m->din.T; *emb = m->din.emb; *watt = m->din.watt; *w1 = m->din.w1;
    *b1 = m->din.b1; *w2 = m->din.w2; *b2 = m->din.b2;
    return 0;
}
void orc_tdm_model_destroy(orc_tdm_model *m) { if (m) { if (!m->deepfm) orc_din_free_f32(&m->din); free(m); } }

int orc_din_forward_f32_api(const orc_tdm_model *m, int64_t n, const int32_t *node, const int32_t *seq,
                            const int32_t *mask_flat, int64_t n_mask, float *out)
{
    if (m->deepfm) {                                                 /* DeepFM takes no mask input (DeepFM.scala:14-15) */
        const int E = m->din.E, T = m->din.T;
        float *K = (float *)malloc(sizeof(float) * T * E), *q = (float *)malloc(sizeof(float) * E);
        float *scratch = (float *)malloc(sizeof(float) * E);
        int rc = 0;
        for (int64_t r = 0; r < n; r++) {
            if (orc_gather_history_f32(&m->din, seq + r * T, K)) { rc = -1; break; }
            int32_t c = node[r];
            if (c == -1) { for (int k = 0; k < E; k++) q[k] = 0.0f; }
            else if (c >= 0 && (int64_t)c < m->din.rows) memcpy(q, m->din.emb + (size_t)c * E, sizeof(float) * E);
            else { rc = -1; break; }
            out[r] = orc_deepfm_row_f32(m, q, K, scratch);
        }
        free(K); free(q); free(scratch);
        return rc;
    }
    return orc_din_forward_f32(&m->din, n, node, seq, mask_flat, n_mask, out);
}

/* Recommender._recommend  (tdm/.../model/Recommender.scala:40-107).
 * out_items/out_logits need room for every scored leaf: 2*beam*(levels+1) is a
 * safe bound (the writers always put leaves at max_level, so 2*beam in
 * practice).  Returns the number of (id, logit) pairs, or <0 on index error. */
int orc_tdm_recommend_raw(const orc_tree *t, const orc_tdm_model *m, const int32_t *seq_ids, int beam,
                          int use_mask, const int32_t *consumed, int n_consumed,
                          int32_t *out_items, float *out_logits, int cap)
{
    const orc_din_f32 *d = &m->din;
    const int E = d->E, T = d->T;
    int32_t *hist = (int32_t *)malloc(sizeof(int32_t) * T);
    uint8_t *masked = (uint8_t *)malloc(T);
    float *K = (float *)malloc(sizeof(float) * T * E);
    float *scratch = (float *)malloc(sizeof(float) * (3 * E + T));
    int ncap = 2 * beam + 2;
    int32_t *cand = (int32_t *)malloc(sizeof(int32_t) * ncap);     /* candidate codes  */
    float *pred = (float *)malloc(sizeof(float) * ncap);
    int32_t *nl = (int32_t *)malloc(sizeof(int32_t) * ncap);       /* non-leaf subset  */
    float *nlp = (float *)malloc(sizeof(float) * ncap);
    orc_kp *kp = (orc_kp *)malloc(sizeof(orc_kp) * ncap);
    /* leaf list: the reference PREPENDS each level's leaves (`leafNodes ++: acc`) */
    int lcap = cap, nleaf = 0;
    int32_t *lcode = (int32_t *)malloc(sizeof(int32_t) * lcap);
    float *lpred = (float *)malloc(sizeof(float) * lcap);
    int rc = 0;

    orc_tdm_id_to_code(t, T, seq_ids, hist, masked);               /* duplicateSequence :116-158 */
    if (!use_mask) memset(masked, 0, T);
    if (orc_gather_history_f32(d, hist, K)) { rc = -2; goto done; }

    int level = orc_lower_log2(beam);                              /* getLevelStart :210-216 */
    int64_t start = ((int64_t)1 << level) - 1, end = start * 2 + 1;
    int ncand = 0;
    if (level <= t->max_level)
        for (int64_t c = start; c < end && c < t->n_codes; c++)
            if (t->exists[c]) {   /* at most 2^level <= beam of them */
                cand[ncand] = (int32_t)c; pred[ncand] = 0.0f; ncand++;
            }

    for (int it = level; it <= t->max_level; it++) {               /* foldLeft :58 */
        if (ncand == 0) continue;
        /* partition :62-64 */
        int nnl = 0, nl_leaf = 0;
        for (int i = 0; i < ncand; i++) if (t->is_leaf[cand[i]]) nl_leaf++;
        if (nl_leaf) {                                             /* prepend, keeping their order */
            if (nleaf + nl_leaf > lcap) { rc = -3; goto done; }
            memmove(lcode + nl_leaf, lcode, sizeof(int32_t) * nleaf);
            memmove(lpred + nl_leaf, lpred, sizeof(float) * nleaf);
            int w = 0;
            for (int i = 0; i < ncand; i++) if (t->is_leaf[cand[i]]) { lcode[w] = cand[i]; lpred[w] = pred[i]; w++; }
            nleaf += nl_leaf;
        }
        for (int i = 0; i < ncand; i++) if (!t->is_leaf[cand[i]]) { nl[nnl] = cand[i]; nlp[nnl] = pred[i]; nnl++; }
        if (nnl == 0) { ncand = 0; continue; }
        /* beam cut :75-87 : stable sort by y.pred.compareTo(x.pred), take(beam) */
        int nb = nnl;
        if (nnl > beam) {
            for (int i = 0; i < nnl; i++) { kp[i].key = orc_key_f32(nlp[i]); kp[i].pos = i; }
            orc_sort_desc(kp, nnl);
            nb = beam;
            for (int i = 0; i < nb; i++) cand[i] = nl[kp[i].pos];
        } else {
            for (int i = 0; i < nb; i++) cand[i] = nl[i];
        }
        /* children :88-92 ; scoring :93-99 */
        int nc = 0;
        for (int i = 0; i < nb; i++)
            for (int s = 1; s <= 2; s++) {
                int64_t c = 2 * (int64_t)cand[i] + s;
                if (c < t->n_codes && t->exists[c]) nl[nc++] = (int32_t)c;
            }
        for (int i = 0; i < nc; i++) {
            const float *q = d->emb + (size_t)nl[i] * E;
            cand[i] = nl[i];
            pred[i] = m->deepfm ? orc_deepfm_row_f32(m, q, K, scratch) : orc_din_row_f32(d, q, K, masked, scratch);
        }
        ncand = nc;
    }
    /* :103-106 */
    int n = 0;
    for (int i = 0; i < nleaf; i++) {
        int32_t id = t->node_id[lcode[i]];
        int skip = 0;
        for (int j = 0; j < n_consumed; j++) if (consumed[j] == id) { skip = 1; break; }
        if (skip) continue;
        if (n >= cap) { rc = -3; goto done; }
        out_items[n] = id; out_logits[n] = lpred[i]; n++;
    }
    rc = n;
done:
    free(hist); free(masked); free(K); free(scratch); free(cand); free(pred); free(nl); free(nlp);
    free(kp); free(lcode); free(lpred);
    return rc;
}

/* Recommender.recommendItems :18-38 (consumed != NULL: eval variant, beam widened
 * to max((|consumed|+topk)/2, beam)) and TDM.recommend (TDM.scala:17-22).
 * out_prob (nullable) = TDM.sigmoid in Double (TDM.scala:56-58). */
int orc_tdm_recommend(const orc_tree *t, const orc_tdm_model *m, const int32_t *seq_ids, int beam, int topk,
                      int use_mask, const int32_t *consumed, int n_consumed, int widen_beam,
                      int32_t *out_items, float *out_logits, double *out_prob)
{
    int b = beam;
    if (widen_beam) { int w = (n_consumed + topk) / 2; if (w > b) b = w; }
    int cap = 2 * b * (t->max_level + 2) + 8;
    int32_t *items = (int32_t *)malloc(sizeof(int32_t) * cap);
    float *logits = (float *)malloc(sizeof(float) * cap);
    int n = orc_tdm_recommend_raw(t, m, seq_ids, b, use_mask, consumed, n_consumed, items, logits, cap);
    if (n < 0) { free(items); free(logits); return n; }
    orc_kp *kp = (orc_kp *)malloc(sizeof(orc_kp) * (size_t)(n + 1));
    for (int i = 0; i < n; i++) { kp[i].key = orc_key_f32(logits[i]); kp[i].pos = i; }
    orc_sort_desc(kp, n);                       /* sortBy(_._2)(Ordering[Float].reverse) -- stable */
    int k = n < topk ? n : topk;
    for (int i = 0; i < k; i++) {
        out_items[i] = items[kp[i].pos];
        if (out_logits) out_logits[i] = logits[kp[i].pos];
        if (out_prob) out_prob[i] = 1.0 / (1.0 + exp(-(double)logits[kp[i].pos]));
    }
    free(items); free(logits); free(kp);
    return k;
}

/* ------------------------------------------------------------------ OTM -- */
/* OTM's DeepFM (otm/src/main/scala/com/mass/otm/model/DeepFM.scala:12-48) is the TDM graph instantiated for Double:
 * same features, FM (scalann/.../nn/FM.scala:14-44), Linear((T+1)E, T+1) -> ReLU -> Linear(T+1, 1), Add; no mask input. */
typedef struct { const double *w1, *b1, *w2, *b2; } orc_dfm_f64;
struct orc_otm_model { orc_din_f64 din; int deepfm; orc_dfm_f64 dfm; };

static double orc_deepfm_row_f64(const orc_otm_model *m, const double *q, const double *K, double *scratch)
{
    const int E = m->din.E, T = m->din.T, F = T + 1;
    double *buf = scratch;                                          /* E */
    for (int k = 0; k < E; k++) buf[k] = 0.0;
    for (int k = 0; k < E; k++) buf[k] = buf[k] + q[k];             /* vAdd, feature 0 = the item */
    for (int j = 0; j < T; j++)
        for (int k = 0; k < E; k++) buf[k] = buf[k] + K[(size_t)j * E + k];
    double sum_square = 0.0, square_sum = 0.0;
    for (int k = 0; k < E; k++) sum_square = fma(buf[k], buf[k], sum_square);
    for (int k = 0; k < E; k++) square_sum = fma(q[k], q[k], square_sum);
    for (int k = 0; k < T * E; k++) square_sum = fma(K[k], K[k], square_sum);
    const double fm = (sum_square - square_sum) / 2.0;
    double dnn = 0.0;
    for (int o = 0; o < F; o++) {
        const double *w = m->dfm.w1 + (size_t)o * F * E;
        double acc = 0.0;
        for (int k = 0; k < E; k++) acc = fma(q[k], w[k], acc);
        for (int k = 0; k < T * E; k++) acc = fma(K[k], w[E + k], acc);
        double h = acc + m->dfm.b1[o];
        h = h > 0.0 ? h : (h != h ? h : 0.0);                       /* ReLU.scala:30-44 */
        dnn = fma(h, m->dfm.w2[o], dnn);
    }
    dnn = dnn + m->dfm.b2[0];
    return fm + dnn;
}

/* params = [emb rows*E | W1 (T+1) x (T+1)E | b1 T+1 | W2 T+1 | b2 1] (Graph.scala:37-48, topological order) */
orc_otm_model *orc_otm_deepfm_create(int64_t rows, int E, int T, const double *params)
{
    orc_otm_model *m = (orc_otm_model *)calloc(1, sizeof(*m));
    const int F = T + 1;
    m->deepfm = 1;
    m->din.rows = rows; m->din.E = E; m->din.T = T; m->din.emb = params;      /* gather + index checks reuse the DIN struct */
    m->dfm.w1 = params + rows * E;
    m->dfm.b1 = m->dfm.w1 + (int64_t)F * F * E;
    m->dfm.w2 = m->dfm.b1 + F;
    m->dfm.b2 = m->dfm.w2 + F;
    return m;
}

orc_otm_model *orc_otm_model_create(int64_t rows, int E, int T, const double *params)
{
    orc_otm_model *m = (orc_otm_model *)calloc(1, sizeof(*m));
    const double *emb = params, *watt = emb + rows * E, *w1 = watt + (int64_t)E * E;
    const double *b1 = w1 + (int64_t)2 * E * E, *w2 = b1 + E, *b2 = w2 + E;
    if (orc_din_init_f64(&m->din, rows, E, T, emb, watt, w1, b1, w2, b2)) { free(m); return NULL; }
    return m;
}
void orc_otm_model_destroy(orc_otm_model *m) { if (m) { if (!m->deepfm) orc_din_free_f64(&m->din); free(m); } }

int orc_din_forward_f64_api(const orc_otm_model *m, int64_t n, const int32_t *node, const int32_t *seq,
                            const int32_t *mask_flat, int64_t n_mask, double *out)
{
    if (m->deepfm) {                                                 /* no mask input (DeepFM.scala:17-18) */
        const int E = m->din.E, T = m->din.T;
        double *K = (double *)malloc(sizeof(double) * T * E), *q = (double *)malloc(sizeof(double) * E);
        double *scratch = (double *)malloc(sizeof(double) * E);
        int rc = 0;
        for (int64_t r = 0; r < n; r++) {
            if (orc_gather_history_f64(&m->din, seq + r * T, K)) { rc = -1; break; }
            int32_t c = node[r];
            if (c == -1) { for (int k = 0; k < E; k++) q[k] = 0.0; }
            else if (c >= 0 && (int64_t)c < m->din.rows) memcpy(q, m->din.emb + (size_t)c * E, sizeof(double) * E);
            else { rc = -1; break; }
            out[r] = orc_deepfm_row_f64(m, q, K, scratch);
        }
        free(K); free(q); free(scratch);
        return rc;
    }
    return orc_din_forward_f64(&m->din, n, node, seq, mask_flat, n_mask, out);
}

/* CandidateSearcher.beamSearch / batchBeamSearch per user
 * (otm/src/main/scala/com/mass/otm/model/CandidateSearcher.scala:15-80,106-121;
 * OTMTree.initializeBeam otm/.../tree/OTMTree.scala:16-23).  seq = leaf node
 * ids (-1 = padding, masked when use_mask).  Writes the candidates of the LAST
 * scored level (2*min(beam, ...) ids + scores); returns their count. */
/* ---- OTM pseudo targets -------------------------------------------------------------------------------------------------
 * OTMTree.optimalPseudoTargets (otm/src/main/scala/com/mass/otm/tree/OTMTree.scala:27-46): the leaf level holds the target
 * items with score 1.0; every level above is computeTargets of the level below (:104-129):
 *   computeChildrenScores (:131-165): for every node n of the user's list, its sibling s (n even -> n-1, else n+1), the model's
 *     logits of n and of s with the user's history (mask = positions of the padding id), negLabel = score of s if s is in the
 *     user's list else 0.  QUIRK kept: without a mask input the reference scores the NEGATIVE tensor for both (:157-161).
 *   label(n) = score(n) if pred(n) >= pred(s) else negLabel (:117-118); the parent's target is the sum of the labels of its
 *     children that are in the list (groupMapReduce(_ + _), :120-121), clipped to [0, 1] (:123, otm/package.scala clipValue).
 * The lists are Maps turned into Lists, so their order is a HashMap artefact; every consumer looks nodes up by id
 * (MiniBatch.batchTransform, otm/.../dataset/MiniBatch.scala:27-34), so the oracle keeps each user's list sorted by id. */
int orc_otm_pseudo_targets(const orc_otm_model *m, int B, int T, const int32_t *seqs, const int64_t *target_off,
                           const int32_t *targets, int leaf_level, int start_level, int use_mask, int M,
                           int32_t *out_ids, double *out_vals, int32_t *out_cnt)
{
    const int n_lvl = leaf_level - start_level;
    if (n_lvl <= 0) return 0;
    int32_t *pos = (int32_t *)malloc(sizeof(int32_t) * (size_t)B * M), *neg = (int32_t *)malloc(sizeof(int32_t) * (size_t)B * M);
    int32_t *rseq = (int32_t *)malloc(sizeof(int32_t) * (size_t)B * M * T), *mask = (int32_t *)malloc(sizeof(int32_t) * (size_t)B * M * T);
    double *pp = (double *)malloc(sizeof(double) * (size_t)B * M), *pn = (double *)malloc(sizeof(double) * (size_t)B * M);
    int rc = 0;
    for (int li = 0; li < n_lvl; li++)
        for (size_t i = 0; i < (size_t)B * M; i++) { out_ids[(size_t)li * B * M + i] = -1; out_vals[(size_t)li * B * M + i] = 0.0; }
    /* leaf level: Node(target, 1.0); duplicates collapse (all scores are 1.0: a duplicate can only push a sum further past the clip) */
    {
        int32_t *ids = out_ids + (size_t)(n_lvl - 1) * B * M;
        double *vals = out_vals + (size_t)(n_lvl - 1) * B * M;
        for (int u = 0; u < B; u++) {
            int c = 0;
            for (int64_t q = target_off[u]; q < target_off[u + 1]; q++) {
                const int32_t t = targets[q];
                int k = 0;
                while (k < c && ids[(size_t)u * M + k] < t) k++;
                if (k < c && ids[(size_t)u * M + k] == t) continue;
                if (c >= M) { rc = -1; goto done; }
                for (int j = c; j > k; j--) ids[(size_t)u * M + j] = ids[(size_t)u * M + j - 1];
                ids[(size_t)u * M + k] = t;
                c++;
            }
            for (int k = 0; k < c; k++) vals[(size_t)u * M + k] = 1.0;
            out_cnt[(size_t)(n_lvl - 1) * B + u] = c;
        }
    }
    for (int li = n_lvl - 1; li > 0; li--) {                      /* children at list li -> parents at list li - 1 */
        const int32_t *cid = out_ids + (size_t)li * B * M;
        const double *cval = out_vals + (size_t)li * B * M;
        const int32_t *ccnt = out_cnt + (size_t)li * B;
        int64_t n = 0, nm = 0;
        for (int u = 0; u < B; u++)
            for (int k = 0; k < ccnt[u]; k++) {
                const int32_t id = cid[(size_t)u * M + k];
                pos[n] = id;
                neg[n] = (id % 2 == 0) ? id - 1 : id + 1;
                for (int j = 0; j < T; j++) {
                    rseq[n * T + j] = seqs[(size_t)u * T + j];
                    if (use_mask && seqs[(size_t)u * T + j] == -1) mask[nm++] = (int32_t)(n * T + j);
                }
                n++;
            }
        if (n > 0) {
            rc = orc_din_forward_f64_api(m, n, use_mask ? pos : neg, rseq, use_mask ? mask : NULL, use_mask ? nm : 0, pp);
            if (rc) goto done;
            rc = orc_din_forward_f64_api(m, n, neg, rseq, use_mask ? mask : NULL, use_mask ? nm : 0, pn);
            if (rc) goto done;
        }
        int32_t *pid = out_ids + (size_t)(li - 1) * B * M;
        double *pval = out_vals + (size_t)(li - 1) * B * M;
        int64_t row = 0;
        for (int u = 0; u < B; u++) {
            int c = 0;
            for (int k = 0; k < ccnt[u]; k++, row++) {
                const int32_t id = cid[(size_t)u * M + k], sib = neg[row];
                double neg_label = 0.0;
                for (int q = 0; q < ccnt[u]; q++) if (cid[(size_t)u * M + q] == sib) { neg_label = cval[(size_t)u * M + q]; break; }
                const double label = pp[row] >= pn[row] ? cval[(size_t)u * M + k] : neg_label;
                const int32_t par = (id - 1) >> 1;
                int j = 0;
                while (j < c && pid[(size_t)u * M + j] != par) j++;
                if (j == c) { pid[(size_t)u * M + c] = par; pval[(size_t)u * M + c] = 0.0; c++; }   /* children are id-sorted: parents arrive ascending */
                pval[(size_t)u * M + j] = pval[(size_t)u * M + j] + label;
            }
            for (int j = 0; j < c; j++) {
                const double v = pval[(size_t)u * M + j];
                pval[(size_t)u * M + j] = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
            }
            out_cnt[(size_t)(li - 1) * B + u] = c;
        }
    }
done:
    free(pos); free(neg); free(rseq); free(mask); free(pp); free(pn);
    return rc;
}

int orc_otm_beam_search(const orc_otm_model *m, const int32_t *seq, int leaf_level, int beam, int use_mask,
                        int32_t *out_ids, double *out_scores)
{
    const orc_din_f64 *d = &m->din;
    const int E = d->E, T = d->T;
    uint8_t *masked = (uint8_t *)calloc(T, 1);
    double *K = (double *)malloc(sizeof(double) * T * E);
    double *scratch = (double *)malloc(sizeof(double) * (3 * E + T));
    int start_level = orc_lower_log2(beam);                        /* otm/package.scala:15 */
    int64_t start = ((int64_t)1 << start_level) - 1, end = start * 2 + 1;
    int64_t n0 = end - start;
    int cap = (int)(2 * (n0 > beam ? n0 : beam)) + 2;
    int32_t *ids = (int32_t *)malloc(sizeof(int32_t) * cap), *nxt = (int32_t *)malloc(sizeof(int32_t) * cap);
    double *sc = (double *)malloc(sizeof(double) * cap);
    orc_kp *kp = (orc_kp *)malloc(sizeof(orc_kp) * cap);
    int rc = 0, n = 0;
    if (use_mask) for (int j = 0; j < T; j++) masked[j] = (seq[j] == -1);
    if (orc_gather_history_f64(d, seq, K)) { rc = -2; goto done; }
    for (int64_t c = start; c < end; c++) { ids[n] = (int32_t)c; sc[n] = 0.0; n++; }
    for (int level = start_level; level < leaf_level; level++) {
        int nb = n;
        if (level != start_level) {                                /* sortBy(_.score)(reverse).take(beam) */
            for (int i = 0; i < n; i++) { kp[i].key = orc_key_f64(sc[i]); kp[i].pos = i; }
            orc_sort_desc(kp, n);
            nb = n < beam ? n : beam;
            for (int i = 0; i < nb; i++) nxt[i] = ids[kp[i].pos];
            memcpy(ids, nxt, sizeof(int32_t) * nb);
        }
        for (int i = 0; i < nb; i++) { nxt[2 * i] = ids[i] * 2 + 1; nxt[2 * i + 1] = ids[i] * 2 + 2; }
        n = 2 * nb;
        for (int i = 0; i < n; i++) {
            if (nxt[i] < 0 || nxt[i] >= d->rows) { rc = -2; goto done; }
            ids[i] = nxt[i];
            sc[i] = m->deepfm ? orc_deepfm_row_f64(m, d->emb + (size_t)nxt[i] * E, K, scratch)
                              : orc_din_row_f64(d, d->emb + (size_t)nxt[i] * E, K, masked, scratch);
        }
    }
    memcpy(out_ids, ids, sizeof(int32_t) * n);
    memcpy(out_scores, sc, sizeof(double) * n);
    rc = n;
done:
    free(masked); free(K); free(scratch); free(ids); free(nxt); free(sc); free(kp);
    return rc;
}

/* OTM.recommend (otm/.../model/OTM.scala:14-23): keep candidates whose id maps
 * back to an item, stable sort desc, topk.  leaf_item[id - leaf_start] = item
 * id or -1 (idItemMapping), leaf_start = 2^leaf_level - 1. */
int orc_otm_recommend(const orc_otm_model *m, const int32_t *seq_leaf_ids, int leaf_level, int beam, int topk,
                      int use_mask, const int32_t *leaf_item, int32_t *out_items, double *out_scores,
                      double *out_prob)
{
    int start_level = orc_lower_log2(beam);
    int64_t n0 = (int64_t)1 << start_level;
    int cap = (int)(2 * (n0 > beam ? n0 : beam)) + 2;
    int32_t *ids = (int32_t *)malloc(sizeof(int32_t) * cap);
    double *sc = (double *)malloc(sizeof(double) * cap);
    orc_kp *kp = (orc_kp *)malloc(sizeof(orc_kp) * cap);
    int n = orc_otm_beam_search(m, seq_leaf_ids, leaf_level, beam, use_mask, ids, sc);
    if (n < 0) { free(ids); free(sc); free(kp); return n; }
    int64_t leaf_start = ((int64_t)1 << leaf_level) - 1, n_leaf = (int64_t)1 << leaf_level;
    int nk = 0;
    for (int i = 0; i < n; i++) {
        int64_t off = (int64_t)ids[i] - leaf_start;
        if (off >= 0 && off < n_leaf && leaf_item[off] >= 0) { kp[nk].key = orc_key_f64(sc[i]); kp[nk].pos = i; nk++; }
    }
    orc_sort_desc(kp, nk);
    int k = nk < topk ? nk : topk;
    for (int i = 0; i < k; i++) {
        int p = kp[i].pos;
        out_items[i] = leaf_item[ids[p] - leaf_start];
        if (out_scores) out_scores[i] = sc[p];
        if (out_prob) out_prob[i] = 1.0 / (1.0 + exp(-sc[p]));
    }
    free(ids); free(sc); free(kp);
    return k;
}

/* -------------------------------------------------------- Deep Retrieval -- */
/* LayerModel (deep-retrieval/src/main/scala/com/mass/dr/model/LayerModel.scala:22-84),
 * RerankModel (RerankModel.scala:20-94).  All Double. */
struct orc_dr_model {
    int num_item, K, D, T, E;
    const double *layer_emb;        /* (num_item + K*(D-1)) x E */
    const double **layer_w;         /* D : K x ((T+d)E)  [out][in] */
    const double **layer_b;         /* D : K */
    const double *rr_emb;           /* num_item x E */
    const double *rr_w;             /* E x (T*E) */
    const double *rr_b;             /* E */
    const double *sm_w;             /* num_item x E */
    const double *sm_b;             /* num_item */
};

orc_dr_model *orc_dr_model_create(int num_item, int K, int D, int T, int E, const double *layer_emb,
                                  const double *const *layer_w, const double *const *layer_b,
                                  const double *rr_emb, const double *rr_w, const double *rr_b,
                                  const double *sm_w, const double *sm_b)
{
    orc_dr_model *m = (orc_dr_model *)calloc(1, sizeof(*m));
    m->num_item = num_item; m->K = K; m->D = D; m->T = T; m->E = E;
    m->layer_emb = layer_emb;
    m->layer_w = (const double **)malloc(sizeof(double *) * D);
    m->layer_b = (const double **)malloc(sizeof(double *) * D);
    for (int d = 0; d < D; d++) { m->layer_w[d] = layer_w[d]; m->layer_b[d] = layer_b[d]; }
    m->rr_emb = rr_emb; m->rr_w = rr_w; m->rr_b = rr_b; m->sm_w = sm_w; m->sm_b = sm_b;
    return m;
}
void orc_dr_model_destroy(orc_dr_model *m) { if (m) { free(m->layer_w); free(m->layer_b); free(m); } }

/* LayerModel.inference :68-84 : x = concat(emb rows of inputSeq), out = W x (addmv) then + bias */
static int orc_dr_layer_logits(const orc_dr_model *m, const int32_t *input, int len, int rank, double *x,
                               double *out)
{
    const int E = m->E;
    const int64_t rows = (int64_t)m->num_item + (int64_t)m->K * (m->D - 1);
    for (int i = 0; i < len; i++) {
        if (input[i] == -1) for (int k = 0; k < E; k++) x[(size_t)i * E + k] = 0.0;
        else if (input[i] >= 0 && input[i] < rows) memcpy(x + (size_t)i * E, m->layer_emb + (size_t)input[i] * E, sizeof(double) * E);
        else return -2;
    }
    const int in = len * E;
    const double *W = m->layer_w[rank], *b = m->layer_b[rank];
    for (int o = 0; o < m->K; o++) {
        double acc = 0.0;
        const double *w = W + (size_t)o * in;
        for (int k = 0; k < in; k++) acc = fma(w[k], x[k], acc);
        out[o] = acc + b[o];
    }
    return 0;
}

/* CandidateSearcher.beamSearch (dr/model/CandidateSearcher.scala:22-60) with
 * softmax of dr/package.scala:23-28 (max, exp(x-max), left-to-right sum, divide).
 * out_paths: beam x D node indices, out_prob: beam.  Returns #paths. */
int orc_dr_beam_search(const orc_dr_model *m, const int32_t *seq, int beam, int32_t *out_paths, double *out_prob)
{
    const int K = m->K, D = m->D, T = m->T, E = m->E;
    int cap = beam > 1 ? beam : 1;
    int32_t *path = (int32_t *)calloc((size_t)cap * D, sizeof(int32_t));   /* live paths */
    int32_t *npath = (int32_t *)calloc((size_t)cap * D, sizeof(int32_t));
    double *prob = (double *)malloc(sizeof(double) * cap), *nprob = (double *)malloc(sizeof(double) * cap);
    double *cprob = (double *)malloc(sizeof(double) * (size_t)cap * K);
    orc_kp *kp = (orc_kp *)malloc(sizeof(orc_kp) * (size_t)cap * K);
    int32_t *input = (int32_t *)malloc(sizeof(int32_t) * (T + D));
    double *x = (double *)malloc(sizeof(double) * (size_t)(T + D) * E);
    double *logit = (double *)malloc(sizeof(double) * K);
    int live = 1, rc = 0;
    prob[0] = 1.0;
    for (int i = 0; i < D && rc == 0; i++) {
        for (int p = 0; p < live; p++) {
            memcpy(input, seq, sizeof(int32_t) * T);
            for (int j = 0; j < i; j++) input[T + j] = path[(size_t)p * D + j] + m->num_item + j * K;
            if (orc_dr_layer_logits(m, input, T + i, i, x, logit)) { rc = -2; break; }
            double mx = logit[0];
            for (int c = 1; c < K; c++) if (logit[c] > mx) mx = logit[c];       /* logits.max */
            double sum = 0.0;
            for (int c = 0; c < K; c++) { logit[c] = orc_exp(logit[c] - mx); sum = sum + logit[c]; }
            for (int c = 0; c < K; c++) {
                size_t idx = (size_t)p * K + c;
                cprob[idx] = prob[p] * (logit[c] / sum);
                kp[idx].key = orc_key_f64(cprob[idx]); kp[idx].pos = (int32_t)idx;
            }
        }
        if (rc) break;
        int n = live * K;
        orc_sort_desc(kp, n);                                      /* sortBy(_.probability)(reverse) */
        int nb = n < beam ? n : beam;
        for (int q = 0; q < nb; q++) {
            int p = kp[q].pos / K, c = kp[q].pos % K;
            memcpy(npath + (size_t)q * D, path + (size_t)p * D, sizeof(int32_t) * D);
            npath[(size_t)q * D + i] = c;
            nprob[q] = cprob[kp[q].pos];
        }
        memcpy(path, npath, sizeof(int32_t) * (size_t)nb * D);
        memcpy(prob, nprob, sizeof(double) * nb);
        live = nb;
    }
    if (rc == 0) {
        memcpy(out_paths, path, sizeof(int32_t) * (size_t)live * D);
        memcpy(out_prob, prob, sizeof(double) * live);
        rc = live;
    }
    free(path); free(npath); free(prob); free(nprob); free(cprob); free(kp); free(input); free(x); free(logit);
    return rc;
}

/* RerankModel.inference :43-68 : u = W_r x + b_r ; score_i = smW[item_i] . u (addmv) + smB[item_i] */
int orc_dr_rerank(const orc_dr_model *m, const int32_t *seq, int n_cand, const int32_t *cand, double *out)
{
    const int T = m->T, E = m->E, in = T * E;
    double *x = (double *)malloc(sizeof(double) * in), *u = (double *)malloc(sizeof(double) * E);
    int rc = 0;
    for (int i = 0; i < T; i++) {
        if (seq[i] == -1) for (int k = 0; k < E; k++) x[(size_t)i * E + k] = 0.0;
        else if (seq[i] >= 0 && seq[i] < m->num_item) memcpy(x + (size_t)i * E, m->rr_emb + (size_t)seq[i] * E, sizeof(double) * E);
        else { rc = -2; goto done; }
    }
    for (int o = 0; o < E; o++) {
        double acc = 0.0;
        const double *w = m->rr_w + (size_t)o * in;
        for (int k = 0; k < in; k++) acc = fma(w[k], x[k], acc);
        u[o] = acc + m->rr_b[o];
    }
    for (int i = 0; i < n_cand; i++) {
        if (cand[i] < 0 || cand[i] >= m->num_item) { rc = -2; goto done; }
        const double *w = m->sm_w + (size_t)cand[i] * E;
        double acc = 0.0;
        for (int k = 0; k < E; k++) acc = fma(w[k], u[k], acc);
        out[i] = acc + m->sm_b[cand[i]];
    }
done:
    free(x); free(u);
    return rc;
}

/* DeepRetrieval.recommend (dr/model/DeepRetrieval.scala:26-46) given the
 * path -> items CSR (MappingOp.pathItemMapping; path key = sum_d c_d K^(D-1-d)).
 * Candidate order = beam order, then the path's item list order
 * (CandidateSearcher.searchCandidate :8-20); duplicates are kept. */
int orc_dr_recommend(const orc_dr_model *m, const int32_t *seq, int beam, int topk, const int64_t *path_off,
                     const int32_t *path_items, int32_t *out_ids, double *out_scores, double *out_prob)
{
    const int D = m->D, K = m->K;
    int cap = beam > 1 ? beam : 1;
    int32_t *paths = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap * D);
    double *pp = (double *)malloc(sizeof(double) * cap);
    int np = orc_dr_beam_search(m, seq, beam, paths, pp);
    if (np < 0) { free(paths); free(pp); return np; }
    int64_t ncand = 0;
    for (int p = 0; p < np; p++) {
        int64_t key = 0;
        for (int d = 0; d < D; d++) key = key * K + paths[(size_t)p * D + d];
        ncand += path_off[key + 1] - path_off[key];
    }
    int32_t *cand = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ncand + 1));
    double *sc = (double *)malloc(sizeof(double) * (size_t)(ncand + 1));
    orc_kp *kp = (orc_kp *)malloc(sizeof(orc_kp) * (size_t)(ncand + 1));
    int64_t w = 0;
    for (int p = 0; p < np; p++) {
        int64_t key = 0;
        for (int d = 0; d < D; d++) key = key * K + paths[(size_t)p * D + d];
        for (int64_t i = path_off[key]; i < path_off[key + 1]; i++) cand[w++] = path_items[i];
    }
    int rc = orc_dr_rerank(m, seq, (int)ncand, cand, sc);
    int k = 0;
    if (rc == 0) {
        for (int64_t i = 0; i < ncand; i++) { kp[i].key = orc_key_f64(sc[i]); kp[i].pos = (int32_t)i; }
        orc_sort_desc(kp, (int)ncand);
        k = ncand < topk ? (int)ncand : topk;
        for (int i = 0; i < k; i++) {
            out_ids[i] = cand[kp[i].pos];
            if (out_scores) out_scores[i] = sc[kp[i].pos];
            if (out_prob) out_prob[i] = 1.0 / (1.0 + exp(-sc[kp[i].pos]));
        }
    } else k = rc;
    free(paths); free(pp); free(cand); free(sc); free(kp);
    return k;
}

/* ------------------------------------------------- batched, multi-thread -- */
/* Evaluator.evaluate splits a mini-batch of users evenly over
 * Engine.coreNumber() threads and loops recommendItems sequentially in each
 * (tdm/src/main/scala/com/mass/tdm/evaluation/Evaluator.scala:28-66).  Same
 * partition rule here: taskSize = B / n, the first B % n threads take one more. */
typedef struct {
    const orc_tree *t; const orc_tdm_model *m; const orc_otm_model *om; const int32_t *seq;
    int T, beam, topk, use_mask, lo, hi, leaf_level, rc;
    const int32_t *leaf_item;
    const int64_t *cons_off; const int32_t *cons; int widen;
    int32_t *out_items; float *out_logits; double *out_scores; int32_t *out_counts;
} orc_job;

static void *orc_tdm_worker(void *p)
{
    orc_job *j = (orc_job *)p;
    for (int u = j->lo; u < j->hi; u++) {
        const int32_t *cons = NULL; int nc = 0;
        if (j->cons_off) { cons = j->cons + j->cons_off[u]; nc = (int)(j->cons_off[u + 1] - j->cons_off[u]); }
        for (int i = 0; i < j->topk; i++) { j->out_items[(size_t)u * j->topk + i] = -1; j->out_logits[(size_t)u * j->topk + i] = 0.0f; }
        int n = orc_tdm_recommend(j->t, j->m, j->seq + (size_t)u * j->T, j->beam, j->topk, j->use_mask, cons, nc,
                                  j->widen, j->out_items + (size_t)u * j->topk, j->out_logits + (size_t)u * j->topk, NULL);
        if (n < 0) { j->rc = n; n = 0; }
        j->out_counts[u] = n;
    }
    return NULL;
}

static void *orc_otm_worker(void *p)
{
    orc_job *j = (orc_job *)p;
    for (int u = j->lo; u < j->hi; u++) {
        for (int i = 0; i < j->topk; i++) { j->out_items[(size_t)u * j->topk + i] = -1; j->out_scores[(size_t)u * j->topk + i] = 0.0; }
        int n = orc_otm_recommend(j->om, j->seq + (size_t)u * j->T, j->leaf_level, j->beam, j->topk, j->use_mask,
                                  j->leaf_item, j->out_items + (size_t)u * j->topk, j->out_scores + (size_t)u * j->topk, NULL);
        if (n < 0) { j->rc = n; n = 0; }
        j->out_counts[u] = n;
    }
    return NULL;
}

static int orc_run(orc_job *proto, int B, int n_threads, void *(*fn)(void *))
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > B) n_threads = B > 0 ? B : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    orc_job *jobs = (orc_job *)malloc(sizeof(orc_job) * n_threads);
    int task = B / n_threads, extra = B % n_threads, rc = 0;
    for (int i = 0; i < n_threads; i++) {
        jobs[i] = *proto;
        jobs[i].lo = i * task + (i < extra ? i : extra);
        jobs[i].hi = jobs[i].lo + task + (i < extra ? 1 : 0);
        jobs[i].rc = 0;
        pthread_create(&th[i], NULL, fn, &jobs[i]);
    }
    for (int i = 0; i < n_threads; i++) { pthread_join(th[i], NULL); if (jobs[i].rc) rc = jobs[i].rc; }
    free(th); free(jobs);
    return rc;
}

int orc_tdm_retrieve_batch(const orc_tree *t, const orc_tdm_model *m, int B, const int32_t *seq_ids, int beam,
                           int topk, int use_mask, const int64_t *cons_off, const int32_t *cons, int widen_beam,
                           int n_threads, int32_t *out_items, float *out_logits, int32_t *out_counts)
{
    orc_job j; memset(&j, 0, sizeof(j));
    j.t = t; j.m = m; j.seq = seq_ids; j.T = m->din.T; j.beam = beam; j.topk = topk; j.use_mask = use_mask;
    j.cons_off = cons_off; j.cons = cons; j.widen = widen_beam;
    j.out_items = out_items; j.out_logits = out_logits; j.out_counts = out_counts;
    return orc_run(&j, B, n_threads, orc_tdm_worker);
}

int orc_otm_retrieve_batch(const orc_otm_model *m, int B, const int32_t *seq_leaf_ids, int leaf_level, int beam,
                           int topk, int use_mask, const int32_t *leaf_item, int n_threads, int32_t *out_items,
                           double *out_scores, int32_t *out_counts)
{
    orc_job j; memset(&j, 0, sizeof(j));
    j.om = m; j.seq = seq_leaf_ids; j.T = m->din.T; j.beam = beam; j.topk = topk; j.use_mask = use_mask;
    j.leaf_level = leaf_level; j.leaf_item = leaf_item;
    j.out_items = out_items; j.out_scores = out_scores; j.out_counts = out_counts;
    return orc_run(&j, B, n_threads, orc_otm_worker);
}

/* known-answer hooks for scalann/src/test/scala/SoftMaxTest.scala */
void orc_softmax_f32(int n, int dim, const float *in, float *out)
{
    for (int i = 0; i < n; i++) {
        const float *x = in + (size_t)i * dim; float *y = out + (size_t)i * dim;
        float mx = x[0];
        for (int d = 1; d < dim; d++) mx = x[d] > mx ? x[d] : mx;
        for (int d = 0; d < dim; d++) y[d] = orc_expf(x[d] - mx);
        float sum = 0.0f;
        for (int d = 0; d < dim; d++) sum = fmaf(y[d], 1.0f, sum);
        float inv = 1.0f / sum;
        for (int d = 0; d < dim; d++) y[d] = y[d] * inv;
    }
}
/* SoftMax.updateGradInput (SoftMax.scala:46-65): g = (go - dot(go, y)) * y */
void orc_softmax_grad_f32(int n, int dim, const float *y, const float *go, float *gi)
{
    for (int i = 0; i < n; i++) {
        const float *yy = y + (size_t)i * dim, *g = go + (size_t)i * dim; float *o = gi + (size_t)i * dim;
        float sum = 0.0f;
        for (int d = 0; d < dim; d++) sum = fmaf(g[d], yy[d], sum);
        for (int d = 0; d < dim; d++) o[d] = (g[d] - sum) * yy[d];
    }
}
float orc_expf_api(float x) { return orc_expf(x); }
double orc_exp_api(double x) { return orc_exp(x); }
