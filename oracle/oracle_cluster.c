/*
 * oracle/oracle_cluster.c -- CPU ORACLE, k-means tree rebuild.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Restates tdm/src/main/scala/com/mass/tdm/cluster/RecursiveCluster.scala:34-214 (clusterType = "kmeans"):
 *   run / train / miniBatch (:34-62, :144-172)   recursive balanced bisection, code 2p+1 / 2p+2 per half
 *   cluster (:176-192)                           2-means on the segment, distances to centroids.head
 *   balanceTree (:194-198)                       argPartition at len / 2 (tdm/.../utils/Utils.scala:130-199: quickselect with a
 *                                                median-of-three pivot and a three-way partition, indices carried along)
 *   squaredDistance (:200-211)
 * The 2-means itself is a THIRD-PARTY dependency absent from /root/reference: smile-core 2.6.0 (project/Dependencies.scala:17-19),
 * smile.clustering.KMeans.fit(data, 2) = k-means++ seeding, centroids = means of the seed partition, Lloyd iterations while the
 * distortion falls by more than tol = 1e-4, at most 100, and PartitionClustering.run(clusterIterNum, ...) = the run with the least
 * distortion.  smile draws from its own MathEx generator, so the reference's trees are not reproducible run to run; this
 * restatement follows the published algorithm with a counter-based generator (splitmix64 keyed by seed, node code, run) and
 * fixes the summation order (chunks of 1024 points, sequential inside a chunk, chunks in order) so that the CUDA path can match it
 * bit for bit.  Parity vs the JVM: unpinned (oracle.c header); pinned here: argPartition / balanceTree against a line-by-line
 * Python port and their defining properties (tests/test_oracle_known_answers.py).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

#define KM_CHUNK 1024

static uint64_t sm64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static uint64_t km_key(uint64_t seed, int64_t pcode, int run, int what)
{
    return sm64(sm64(seed ^ sm64((uint64_t)pcode)) + (uint64_t)run * 2 + (uint64_t)what);
}

/* Utils.argPartition (Utils.scala:130-199) */
static void ap_swap(double *e, int32_t *ix, int a, int b)
{
    const double t = e[a]; e[a] = e[b]; e[b] = t;
    const int32_t u = ix[a]; ix[a] = ix[b]; ix[b] = u;
}
static int ap_med(const double *e, int p1, int p2, int p3)
{
    if (e[p1] < e[p2]) return e[p2] < e[p3] ? p2 : (e[p1] < e[p3] ? p3 : p1);
    return e[p2] > e[p3] ? p2 : (e[p1] > e[p3] ? p3 : p1);
}
int orc_arg_partition(double *e, int n, int position, int32_t *ix)
{
    int left = 0, right = n - 1;
    if (position < left || position > right) return -1;
    while (left < right) {
        const int pvt = ap_med(e, left, right, (int)(((int64_t)left + right) / 2));
        const double pv = e[pvt];
        ap_swap(e, ix, pvt, left);
        int i = left, lt = left, gt = right;
        while (i <= gt) {
            if (e[i] < pv) { ap_swap(e, ix, lt, i); lt++; i++; }
            else if (e[i] > pv) { ap_swap(e, ix, gt, i); gt--; }
            else if (e[i] == pv) i++;
            else return -3;                                        /* "Nan element detected" */
        }
        if (lt <= position && position <= gt) left = right;
        else if (position < lt) right = lt - 1;
        else left = gt + 1;
    }
    return 0;
}

static double sqdist(const double *x, const double *y, int E)
{
    double sum = 0.0;
    for (int i = 0; i < E; i++) { const double d = x[i] - y[i]; sum += d * d; }
    return sum;
}

/* one assignment pass: labels by the nearer centroid (ties -> 0), chunked sums; returns the within-cluster sum of squares */
static double km_pass(const double *emb, int E, const int32_t *idx, int len, const double *c, double *S, double *N)
{
    double wcss = 0.0;
    memset(S, 0, sizeof(double) * 2 * (size_t)E);
    N[0] = N[1] = 0.0;
    double *cs = (double *)malloc(sizeof(double) * 2 * (size_t)E);
    for (int k0 = 0; k0 < len; k0 += KM_CHUNK) {
        const int k1 = k0 + KM_CHUNK < len ? k0 + KM_CHUNK : len;
        double wc = 0.0, cn[2] = {0.0, 0.0};
        memset(cs, 0, sizeof(double) * 2 * (size_t)E);
        for (int i = k0; i < k1; i++) {
            const double *x = emb + (int64_t)idx[i] * E;
            const double d0 = sqdist(x, c, E), d1 = sqdist(x, c + E, E);
            const int lab = d1 < d0 ? 1 : 0;
            wc += lab ? d1 : d0;
            cn[lab] += 1.0;
            for (int e = 0; e < E; e++) cs[lab * E + e] += x[e];
        }
        wcss += wc;
        N[0] += cn[0]; N[1] += cn[1];
        for (int e = 0; e < 2 * E; e++) S[e] += cs[e];
    }
    free(cs);
    return wcss;
}

/* smile KMeans.fit(data, 2, 100, 1e-4), one run: -> distortion, c[2][E] */
static double km_run(const double *emb, int E, const int32_t *idx, int len, uint64_t seed, int64_t pcode, int run, double *c)
{
    double *S = (double *)malloc(sizeof(double) * 2 * (size_t)E), N[2];
    double *d2 = (double *)malloc(sizeof(double) * (size_t)len);
    const int i0 = (int)(km_key(seed, pcode, run, 0) % (uint64_t)len);
    memcpy(c, emb + (int64_t)idx[i0] * E, sizeof(double) * (size_t)E);
    const int nch = (len + KM_CHUNK - 1) / KM_CHUNK;
    double *ct = (double *)malloc(sizeof(double) * (size_t)nch), total = 0.0;
    for (int k = 0; k < nch; k++) {
        double t = 0.0;
        for (int i = k * KM_CHUNK; i < len && i < (k + 1) * KM_CHUNK; i++) { d2[i] = sqdist(emb + (int64_t)idx[i] * E, c, E); t += d2[i]; }
        ct[k] = t;
        total += t;
    }
    int i1 = (i0 + 1) % len;
    if (total > 0.0) {                                             /* k-means++: the second seed with probability proportional to D^2 */
        const double r = (double)(km_key(seed, pcode, run, 1) >> 11) * 0x1.0p-53 * total;
        double pre = 0.0;
        int k = 0;
        while (k < nch - 1 && !(r < pre + ct[k])) { pre += ct[k]; k++; }
        const int k1 = (k + 1) * KM_CHUNK < len ? (k + 1) * KM_CHUNK : len;
        double acc = pre;
        i1 = k1 - 1;
        for (int i = k * KM_CHUNK; i < k1; i++) { acc += d2[i]; if (r < acc) { i1 = i; break; } }
    }
    memcpy(c + E, emb + (int64_t)idx[i1] * E, sizeof(double) * (size_t)E);
    double distortion = 0.0, diff = 1.7976931348623157e308;
    for (int iter = 0; iter <= 100 && diff > 1e-4; iter++) {       /* iter 0 = the seed partition and its means */
        const double w = km_pass(emb, E, idx, len, c, S, N);
        for (int cl = 0; cl < 2; cl++)
            if (N[cl] > 0.0)
                for (int e = 0; e < E; e++) c[cl * E + e] = S[cl * E + e] / N[cl];
        if (iter > 0) diff = distortion - w;
        distortion = w;
    }
    free(S); free(d2); free(ct);
    return distortion;
}

/* RecursiveCluster.cluster + balanceTree on perm[start, start + len): reorders the range into (left | right), returns mid */
static int km_cluster(const double *emb, int E, int32_t *idx, int len, int iters, uint64_t seed, int64_t pcode)
{
    double *c = (double *)malloc(sizeof(double) * 2 * (size_t)E), *best = (double *)malloc(sizeof(double) * (size_t)E);
    double best_d = 0.0;
    for (int run = 0; run < iters; run++) {
        const double d = km_run(emb, E, idx, len, seed, pcode, run, c);
        if (run == 0 || d < best_d) { best_d = d; memcpy(best, c, sizeof(double) * (size_t)E); }
    }
    double *dist = (double *)malloc(sizeof(double) * (size_t)len);
    int32_t *ix = (int32_t *)malloc(sizeof(int32_t) * (size_t)len), *out = (int32_t *)malloc(sizeof(int32_t) * (size_t)len);
    for (int i = 0; i < len; i++) { dist[i] = sqdist(emb + (int64_t)idx[i] * E, best, E); ix[i] = i; }
    orc_arg_partition(dist, len, len / 2, ix);
    for (int i = 0; i < len; i++) out[i] = idx[ix[i]];
    memcpy(idx, out, sizeof(int32_t) * (size_t)len);
    free(c); free(best); free(dist); free(ix); free(out);
    return len / 2;
}

/* RecursiveCluster.run without the file: codes[n] for points 0..n-1 (train + miniBatch give the same codes as this level order:
 * a segment's result depends only on its node code and its index order) */
int orc_kmeans_tree(int n, int E, const double *emb, int iters, uint64_t seed, int32_t *codes)
{
    if (n < 2 || E < 1 || iters < 1) return -1;
    int32_t *perm = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    for (int i = 0; i < n; i++) perm[i] = i;
    typedef struct { int64_t pcode; int start, len; } seg_t;
    seg_t *cur = (seg_t *)malloc(sizeof(seg_t) * (size_t)n), *nxt = (seg_t *)malloc(sizeof(seg_t) * (size_t)n);
    int nc = 1, nn = 0;
    cur[0].pcode = 0; cur[0].start = 0; cur[0].len = n;
    while (nc) {
        nn = 0;
        for (int s = 0; s < nc; s++) {
            const seg_t g = cur[s];
            const int64_t lc = 2 * g.pcode + 1, rc = 2 * g.pcode + 2;
            if (g.len == 2) { codes[perm[g.start]] = (int32_t)lc; codes[perm[g.start + 1]] = (int32_t)rc; continue; }
            const int mid = km_cluster(emb, E, perm + g.start, g.len, iters, seed, g.pcode);
            if (mid == 1) codes[perm[g.start]] = (int32_t)lc;
            else { nxt[nn].pcode = lc; nxt[nn].start = g.start; nxt[nn].len = mid; nn++; }
            if (g.len - mid == 1) codes[perm[g.start + mid]] = (int32_t)rc;
            else { nxt[nn].pcode = rc; nxt[nn].start = g.start + mid; nxt[nn].len = g.len - mid; nn++; }
        }
        seg_t *t = cur; cur = nxt; nxt = t;
        nc = nn;
    }
    free(perm); free(cur); free(nxt);
    return 0;
}
