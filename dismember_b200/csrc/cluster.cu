// cluster.cu -- k-means tree rebuild: tdm/src/main/scala/com/mass/tdm/cluster/RecursiveCluster.scala:34-214, "kmeans".
//
// The reference bisects the catalogue recursively: 2-means on the segment's item embeddings (smile-core 2.6.0 KMeans.fit(data, 2)
// inside PartitionClustering.run(clusterIterNum, ...): third-party, its published algorithm is restated here -- k-means++ seeding,
// means of the seed partition, Lloyd passes while the distortion falls by more than 1e-4, at most 100, best of the runs), squared
// distances to centroids.head (:200-211), balanced split at the median by Utils.argPartition (utils/Utils.scala:130-199), node
// codes 2p+1 / 2p+2 per half (:52-62, :144-172).  Here the recursion runs LEVEL BY LEVEL: every segment of a level is clustered
// in the same launches --
//   segments of more than 1024 points   one CTA per (segment, run, 1024-point chunk): km_seed_kernel / km_pick_kernel, then
//                                       km_pass_kernel + km_update_kernel per Lloyd pass (chunk partials reduced in chunk order),
//                                       km_best_kernel, km_dist_kernel
//   segments of 3..1024 points          one CTA per segment runs every run to convergence and writes the distances (km_small_kernel)
// and the order-defining quickselect (argPartition: sequential by construction, like JTM's reBalance) runs on the host between
// levels.  Sums are sequential inside a chunk and chunks are added in order, distances are mul-then-add chains over ascending k:
// the arithmetic of oracle/oracle_cluster.c, so the codes are identical to the oracle's for the same seed.  smile's own random
// generator is not reproducible, so neither side can match a JVM run tree for tree.
#include <algorithm>
#include <cstring>
#include <vector>

#include "device_utils.cuh"
#include "rows_kernels.cuh"

using namespace dmg;

namespace {

constexpr int kKmChunk = 1024, kKmTP = 64, kKmThreads = 128, kKmMaxIter = 100;
constexpr double kKmTol = 1e-4;

__host__ __device__ inline uint64_t km_sm64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ inline uint64_t km_key(uint64_t seed, int64_t pcode, int run, int what)
{
    return km_sm64(km_sm64(seed ^ km_sm64((uint64_t)pcode)) + (uint64_t)run * 2 + (uint64_t)what);
}

struct KmSmem {                         // carve of the dynamic shared memory (doubles first)
    double *tile, *c, *acc, *best, *d2, *mn, *wn;
    int *lab;
    __device__ KmSmem(unsigned char *raw, int E)
    {
        tile = reinterpret_cast<double *>(raw);
        c = tile + kKmTP * (E + 1);
        acc = c + 2 * E;
        best = acc + 2 * E;
        d2 = best + E;
        mn = d2 + kKmChunk;
        wn = mn + kKmTP;
        lab = reinterpret_cast<int *>(wn + 4);
    }
    static size_t bytes(int E) { return (size_t)(kKmTP * (E + 1) + 5 * E + kKmChunk + kKmTP + 4) * 8 + kKmTP * 4; }
};

__device__ __forceinline__ void km_load_tile(const double *__restrict__ emb, int E, const int32_t *__restrict__ pseg, int t0, int np, double *tile)
{
    for (int q = threadIdx.x; q < np * E; q += blockDim.x) {
        const int p = q / E, e = q - p * E;
        tile[p * (E + 1) + e] = __ldg(emb + (int64_t)pseg[t0 + p] * E + e);
    }
}
__device__ __forceinline__ double km_sqdist(const double *x, const double *y, int E)      // RecursiveCluster.squaredDistance
{
    double sum = 0.0;
    for (int i = 0; i < E; i++) { const double d = __dsub_rn(x[i], y[i]); sum = __dadd_rn(sum, __dmul_rn(d, d)); }
    return sum;
}

// distances of the points [k0, k1) of a segment to s.c[0]: out[i - k0] (shared or global), returns nothing
__device__ void km_chunk_dist(const double *emb, int E, const int32_t *pseg, int k0, int k1, const double *cen, KmSmem &s, double *out)
{
    for (int t0 = k0; t0 < k1; t0 += kKmTP) {
        const int np = min(kKmTP, k1 - t0);
        km_load_tile(emb, E, pseg, t0, np, s.tile);
        __syncthreads();
        if ((int)threadIdx.x < np) out[t0 - k0 + threadIdx.x] = km_sqdist(s.tile + threadIdx.x * (E + 1), cen, E);
        __syncthreads();
    }
}

// one assignment pass over the points [k0, k1): s.acc[2E] += sums per label, s.wn = {wcss, n0, n1} (all zeroed by the caller)
__device__ void km_chunk_pass(const double *emb, int E, const int32_t *pseg, int k0, int k1, KmSmem &s)
{
    for (int t0 = k0; t0 < k1; t0 += kKmTP) {
        const int np = min(kKmTP, k1 - t0);
        km_load_tile(emb, E, pseg, t0, np, s.tile);
        __syncthreads();
        if ((int)threadIdx.x < np) {
            const double *x = s.tile + threadIdx.x * (E + 1);
            const double d0 = km_sqdist(x, s.c, E), d1 = km_sqdist(x, s.c + E, E);
            const int lab = d1 < d0 ? 1 : 0;
            s.lab[threadIdx.x] = lab;
            s.mn[threadIdx.x] = lab ? d1 : d0;
        }
        __syncthreads();
        for (int q = threadIdx.x; q < 2 * E; q += blockDim.x) {
            const int cl = q / E, e = q - cl * E;
            double a = s.acc[q];
            for (int p = 0; p < np; p++)
                if (s.lab[p] == cl) a = __dadd_rn(a, s.tile[p * (E + 1) + e]);
            s.acc[q] = a;
        }
        if (threadIdx.x == 0) {
            double w = s.wn[0], n0 = s.wn[1], n1 = s.wn[2];
            for (int p = 0; p < np; p++) { w = __dadd_rn(w, s.mn[p]); if (s.lab[p]) n1 += 1.0; else n0 += 1.0; }
            s.wn[0] = w; s.wn[1] = n0; s.wn[2] = n1;
        }
        __syncthreads();
    }
}

// the second seed: r = u total, first point whose running sum of D^2 exceeds r (chunk prefix `pre`, points [k0, k1) with D^2 in d2)
__device__ __forceinline__ int km_scan_pick(const double *d2, int k0, int k1, double pre, double r)
{
    double acc = pre;
    for (int i = k0; i < k1; i++) { acc = __dadd_rn(acc, d2[i - k0]); if (r < acc) return i; }
    return k1 - 1;
}

struct KmSeg { int64_t pcode; int32_t start, len; };

// ---- segments of 3..1024 points: everything in one CTA ----------------------------------------------------------------------
__global__ void __launch_bounds__(kKmThreads) km_small_kernel(const double *__restrict__ emb, int E, const int32_t *__restrict__ perm,
                                                              const KmSeg *__restrict__ segs, int iters, uint64_t seed, double *__restrict__ dist)
{
    extern __shared__ __align__(16) unsigned char km_raw[];
    KmSmem s(km_raw, E);
    __shared__ int sI1;
    __shared__ double sState[3];        // distortion, diff, best distortion
    __shared__ int sGo;
    const KmSeg g = segs[blockIdx.x];
    const int32_t *pseg = perm + g.start;
    const int len = g.len, tid = threadIdx.x;
    for (int run = 0; run < iters; run++) {
        const int i0 = (int)(km_key(seed, g.pcode, run, 0) % (uint64_t)len);
        for (int e = tid; e < E; e += blockDim.x) s.c[e] = __ldg(emb + (int64_t)pseg[i0] * E + e);
        __syncthreads();
        km_chunk_dist(emb, E, pseg, 0, len, s.c, s, s.d2);
        if (tid == 0) {
            double t = 0.0;
            for (int i = 0; i < len; i++) t = __dadd_rn(t, s.d2[i]);
            int i1 = (i0 + 1) % len;
            if (t > 0.0) i1 = km_scan_pick(s.d2, 0, len, 0.0, __dmul_rn(__dmul_rn((double)(km_key(seed, g.pcode, run, 1) >> 11), 0x1.0p-53), t));
            sI1 = i1;
            sState[1] = 1.7976931348623157e308;
        }
        __syncthreads();
        for (int e = tid; e < E; e += blockDim.x) s.c[E + e] = __ldg(emb + (int64_t)pseg[sI1] * E + e);
        __syncthreads();
        for (int iter = 0; iter <= kKmMaxIter; iter++) {
            for (int q = tid; q < 2 * E; q += blockDim.x) s.acc[q] = 0.0;
            if (tid == 0) { s.wn[0] = 0.0; s.wn[1] = 0.0; s.wn[2] = 0.0; }
            __syncthreads();
            km_chunk_pass(emb, E, pseg, 0, len, s);
            for (int q = tid; q < 2 * E; q += blockDim.x) {
                const double n = s.wn[1 + q / E];
                if (n > 0.0) s.c[q] = __ddiv_rn(s.acc[q], n);
            }
            if (tid == 0) {
                const double w = s.wn[0];
                if (iter > 0) sState[1] = __dsub_rn(sState[0], w);
                sState[0] = w;
                sGo = (iter + 1 <= kKmMaxIter && sState[1] > kKmTol) ? 1 : 0;
            }
            __syncthreads();
            if (!sGo) break;
        }
        if (run == 0 || sState[0] < sState[2]) {                    // PartitionClustering.run: the least distortion, first run wins a tie
            for (int e = tid; e < E; e += blockDim.x) s.best[e] = s.c[e];
            __syncthreads();
            if (tid == 0) sState[2] = sState[0];
        }
        __syncthreads();
    }
    km_chunk_dist(emb, E, pseg, 0, len, s.best, s, dist + g.start);
}

// ---- segments of more than 1024 points: one CTA per (problem = segment x run, chunk) ------------------------------------------
struct KmBig {
    const double *emb; int E; const int32_t *perm;
    const KmSeg *prob;                  // per problem (segment x run): the segment
    const int32_t *prob_run, *prob_chunk0, *prob_nch;
    const int32_t *chunk_prob;          // per chunk: its problem; chunk index within the problem = k - prob_chunk0
    double *cen;                        // [problem][2][E]
    double *d2;                         // [problem-local point] D^2 to the first seed (offset d2_off[problem])
    const int64_t *d2_off;
    double *ps, *pw;                    // per chunk: sums [2][E], {wcss, n0, n1, chunk total of D^2}
    double *state;                      // per problem: distortion, diff
    int32_t *iter, *done, *pending;
    uint64_t seed;
};

__global__ void __launch_bounds__(kKmThreads) km_seed_kernel(const KmBig b)
{
    extern __shared__ __align__(16) unsigned char km_raw[];
    KmSmem s(km_raw, b.E);
    const int k = blockIdx.x, q = b.chunk_prob[k], E = b.E, tid = threadIdx.x;
    const KmSeg g = b.prob[q];
    const int32_t *pseg = b.perm + g.start;
    const int kc = k - b.prob_chunk0[q], k0 = kc * kKmChunk, k1 = min(k0 + kKmChunk, g.len);
    const int i0 = (int)(km_key(b.seed, g.pcode, b.prob_run[q], 0) % (uint64_t)g.len);
    for (int e = tid; e < E; e += blockDim.x) s.c[e] = __ldg(b.emb + (int64_t)pseg[i0] * E + e);
    __syncthreads();
    km_chunk_dist(b.emb, E, pseg, k0, k1, s.c, s, s.d2);
    double *out = b.d2 + b.d2_off[q] + k0;
    for (int i = tid; i < k1 - k0; i += blockDim.x) out[i] = s.d2[i];
    if (tid == 0) {
        double t = 0.0;
        for (int i = 0; i < k1 - k0; i++) t = __dadd_rn(t, s.d2[i]);
        b.pw[(size_t)k * 4 + 3] = t;
    }
}

__global__ void km_pick_kernel(const KmBig b, int n_prob)
{
    const int q = blockIdx.x, E = b.E;
    if (q >= n_prob) return;
    __shared__ int sI[2];
    const KmSeg g = b.prob[q];
    const int32_t *pseg = b.perm + g.start;
    if (threadIdx.x == 0) {
        const int run = b.prob_run[q], nch = b.prob_nch[q], c0 = b.prob_chunk0[q];
        const int i0 = (int)(km_key(b.seed, g.pcode, run, 0) % (uint64_t)g.len);
        double total = 0.0;
        for (int k = 0; k < nch; k++) total = __dadd_rn(total, b.pw[(size_t)(c0 + k) * 4 + 3]);
        int i1 = (i0 + 1) % g.len;
        if (total > 0.0) {
            const double r = __dmul_rn(__dmul_rn((double)(km_key(b.seed, g.pcode, run, 1) >> 11), 0x1.0p-53), total);
            double pre = 0.0;
            int k = 0;
            while (k < nch - 1 && !(r < __dadd_rn(pre, b.pw[(size_t)(c0 + k) * 4 + 3]))) { pre = __dadd_rn(pre, b.pw[(size_t)(c0 + k) * 4 + 3]); k++; }
            const int k0 = k * kKmChunk, k1 = min(k0 + kKmChunk, g.len);
            i1 = km_scan_pick(b.d2 + b.d2_off[q] + k0, k0, k1, pre, r);
        }
        sI[0] = i0; sI[1] = i1;
        b.state[(size_t)q * 2] = 0.0; b.state[(size_t)q * 2 + 1] = 1.7976931348623157e308;
        b.iter[q] = 0; b.done[q] = 0;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 2 * E; e += blockDim.x) b.cen[(size_t)q * 2 * E + e] = __ldg(b.emb + (int64_t)pseg[sI[e / E]] * E + (e % E));
}

__global__ void __launch_bounds__(kKmThreads) km_pass_kernel(const KmBig b)
{
    extern __shared__ __align__(16) unsigned char km_raw[];
    KmSmem s(km_raw, b.E);
    const int k = blockIdx.x, q = b.chunk_prob[k], E = b.E, tid = threadIdx.x;
    if (b.done[q]) return;
    const KmSeg g = b.prob[q];
    const int kc = k - b.prob_chunk0[q], k0 = kc * kKmChunk, k1 = min(k0 + kKmChunk, g.len);
    for (int e = tid; e < 2 * E; e += blockDim.x) { s.c[e] = b.cen[(size_t)q * 2 * E + e]; s.acc[e] = 0.0; }
    if (tid == 0) { s.wn[0] = 0.0; s.wn[1] = 0.0; s.wn[2] = 0.0; }
    __syncthreads();
    km_chunk_pass(b.emb, E, b.perm + g.start, k0, k1, s);
    for (int e = tid; e < 2 * E; e += blockDim.x) b.ps[(size_t)k * 2 * E + e] = s.acc[e];
    if (tid < 3) b.pw[(size_t)k * 4 + tid] = s.wn[tid];
}

__global__ void km_update_kernel(const KmBig b, int n_prob)
{
    const int q = blockIdx.x, E = b.E;
    if (q >= n_prob || b.done[q]) return;
    const int nch = b.prob_nch[q], c0 = b.prob_chunk0[q];
    for (int e = threadIdx.x; e < 2 * E; e += blockDim.x) {
        double S = 0.0, N = 0.0;
        for (int k = 0; k < nch; k++) {
            S = __dadd_rn(S, b.ps[(size_t)(c0 + k) * 2 * E + e]);
            N = __dadd_rn(N, b.pw[(size_t)(c0 + k) * 4 + 1 + e / E]);
        }
        if (N > 0.0) b.cen[(size_t)q * 2 * E + e] = __ddiv_rn(S, N);
    }
    if (threadIdx.x == 0) {
        double w = 0.0;
        for (int k = 0; k < nch; k++) w = __dadd_rn(w, b.pw[(size_t)(c0 + k) * 4]);
        const int it = b.iter[q];
        double diff = b.state[(size_t)q * 2 + 1];
        if (it > 0) diff = __dsub_rn(b.state[(size_t)q * 2], w);
        b.state[(size_t)q * 2] = w; b.state[(size_t)q * 2 + 1] = diff;
        b.iter[q] = it + 1;
        if (it + 1 <= kKmMaxIter && diff > kKmTol) atomicAdd(b.pending, 1);
        else b.done[q] = 1;
    }
}

// per big segment: the run with the least distortion (first wins a tie) -> its first centroid
__global__ void km_best_kernel(const KmBig b, int n_seg, int iters, double *__restrict__ best)
{
    const int sgi = blockIdx.x, E = b.E;
    if (sgi >= n_seg) return;
    int br = 0;
    double bd = b.state[(size_t)(sgi * iters) * 2];
    for (int r = 1; r < iters; r++) {
        const double d = b.state[(size_t)(sgi * iters + r) * 2];
        if (d < bd) { bd = d; br = r; }
    }
    for (int e = threadIdx.x; e < E; e += blockDim.x) best[(size_t)sgi * E + e] = b.cen[(size_t)(sgi * iters + br) * 2 * E + e];
}

// distances of the big segments' points to their best centroid: the chunks of run 0 cover every point once
__global__ void __launch_bounds__(kKmThreads) km_dist_kernel(const KmBig b, int iters, const double *__restrict__ best, double *__restrict__ dist)
{
    extern __shared__ __align__(16) unsigned char km_raw[];
    KmSmem s(km_raw, b.E);
    const int k = blockIdx.x, q = b.chunk_prob[k], E = b.E;
    if (b.prob_run[q] != 0) return;
    const KmSeg g = b.prob[q];
    const int kc = k - b.prob_chunk0[q], k0 = kc * kKmChunk, k1 = min(k0 + kKmChunk, g.len);
    for (int e = threadIdx.x; e < E; e += blockDim.x) s.best[e] = best[(size_t)(q / iters) * E + e];
    __syncthreads();
    km_chunk_dist(b.emb, E, b.perm + g.start, k0, k1, s.best, s, s.d2);
    for (int i = threadIdx.x; i < k1 - k0; i += blockDim.x) dist[g.start + k0 + i] = s.d2[i];
}

// Utils.argPartition (Utils.scala:130-199): quickselect, median-of-three pivot, three-way partition, indices carried along
inline void ap_swap(double *e, int32_t *ix, int a, int b) { std::swap(e[a], e[b]); std::swap(ix[a], ix[b]); }
inline int ap_med(const double *e, int p1, int p2, int p3)
{
    if (e[p1] < e[p2]) return e[p2] < e[p3] ? p2 : (e[p1] < e[p3] ? p3 : p1);
    return e[p2] > e[p3] ? p2 : (e[p1] > e[p3] ? p3 : p1);
}
bool arg_partition(double *e, int n, int position, int32_t *ix)
{
    int left = 0, right = n - 1;
    while (left < right) {
        const int pvt = ap_med(e, left, right, (int)(((int64_t)left + right) / 2));
        const double pv = e[pvt];
        ap_swap(e, ix, pvt, left);
        int i = left, lt = left, gt = right;
        while (i <= gt) {
            if (e[i] < pv) { ap_swap(e, ix, lt, i); lt++; i++; }
            else if (e[i] > pv) { ap_swap(e, ix, gt, i); gt--; }
            else if (e[i] == pv) i++;
            else return false;                                     // "Nan element detected"
        }
        if (lt <= position && position <= gt) left = right;
        else if (position < lt) right = lt - 1;
        else left = gt + 1;
    }
    return true;
}

template <typename T> int32_t to_dev(dmg_handle_t h, T **dst, const std::vector<T> &v)
{
    DMG_CUDA(h, cudaMalloc(dst, std::max<size_t>(v.size(), 1) * sizeof(T)));
    if (!v.empty()) DMG_CUDA(h, cudaMemcpyAsync(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    return DMG_OK;
}

}  // namespace

/* RecursiveCluster.run (kmeans) without the file: codes[n] of the points 0..n-1.  TreeBuilder.build writes them out
 * (dismember_b200/formats/tree_file.py: build_tree). */
DMG_API int32_t dmg_kmeans_tree(dmg_handle_t h, int32_t n, int32_t E, const double *emb, int32_t iters, uint64_t seed, int32_t *out_codes)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    if (n < 2 || E < 1 || iters < 1 || !emb || !out_codes) return fail(h, DMG_ERR_INVALID_ARG, "dmg_kmeans_tree: bad arguments (n >= 2, clusterIterNum >= 1)");
    const size_t smem = KmSmem::bytes(E);
    if (smem > h->smem_optin) return fail(h, DMG_ERR_UNSUPPORTED, "embedding size %d needs %zu B of shared memory", E, smem);
    DMG_CUDA(h, cudaSetDevice(h->device));
    DMG_CUDA(h, cudaFuncSetAttribute(km_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DMG_CUDA(h, cudaFuncSetAttribute(km_seed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DMG_CUDA(h, cudaFuncSetAttribute(km_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DMG_CUDA(h, cudaFuncSetAttribute(km_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    double *d_emb = nullptr, *d_dist = nullptr;
    int32_t *d_perm = nullptr;
    std::vector<void *> level_allocs;
    auto cleanup = [&]() { for (void *p : level_allocs) cudaFree(p); level_allocs.clear(); };
    struct Guard {
        double *&a, *&b; int32_t *&c; std::vector<void *> &lv;
        ~Guard() { cudaFree(a); cudaFree(b); cudaFree(c); for (void *p : lv) cudaFree(p); }
    } guard{d_emb, d_dist, d_perm, level_allocs};
    DMG_CUDA(h, cudaMalloc(&d_emb, (size_t)n * E * 8));
    DMG_CUDA(h, cudaMalloc(&d_dist, (size_t)n * 8));
    DMG_CUDA(h, cudaMalloc(&d_perm, (size_t)n * 4));
    DMG_CUDA(h, cudaMemcpyAsync(d_emb, emb, (size_t)n * E * 8, cudaMemcpyHostToDevice, h->stream));
    std::vector<int32_t> perm(n), ix, tmp;
    for (int i = 0; i < n; i++) perm[i] = i;
    std::vector<double> dist(n);
    std::vector<KmSeg> cur{{0, 0, n}}, nxt;
    while (!cur.empty()) {
        std::vector<KmSeg> small, big;
        for (const KmSeg &g : cur) {
            if (g.len == 2) { out_codes[perm[g.start]] = (int32_t)(2 * g.pcode + 1); out_codes[perm[g.start + 1]] = (int32_t)(2 * g.pcode + 2); }
            else (g.len > kKmChunk ? big : small).push_back(g);
        }
        if (small.empty() && big.empty()) break;
        DMG_CUDA(h, cudaMemcpyAsync(d_perm, perm.data(), (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
        if (!small.empty()) {
            KmSeg *d_seg = nullptr;
            DMG_TRY(to_dev(h, &d_seg, small)); level_allocs.push_back(d_seg);
            km_small_kernel<<<(unsigned)small.size(), kKmThreads, smem, h->stream>>>(d_emb, E, d_perm, d_seg, iters, seed, d_dist);
            h->launches += 1;
        }
        if (!big.empty()) {
            const int n_seg = (int)big.size(), n_prob = n_seg * iters;
            std::vector<KmSeg> prob(n_prob);
            std::vector<int32_t> prob_run(n_prob), prob_chunk0(n_prob), prob_nch(n_prob), chunk_prob;
            std::vector<int64_t> d2_off(n_prob);
            int64_t d2_total = 0;
            for (int sgi = 0; sgi < n_seg; sgi++)
                for (int r = 0; r < iters; r++) {
                    const int q = sgi * iters + r, nch = (big[sgi].len + kKmChunk - 1) / kKmChunk;
                    prob[q] = big[sgi]; prob_run[q] = r; prob_chunk0[q] = (int32_t)chunk_prob.size(); prob_nch[q] = nch;
                    d2_off[q] = d2_total; d2_total += big[sgi].len;
                    chunk_prob.insert(chunk_prob.end(), nch, q);
                }
            const int n_chunk = (int)chunk_prob.size();
            KmBig b;
            memset(&b, 0, sizeof(b));
            b.emb = d_emb; b.E = E; b.perm = d_perm; b.seed = seed;
            KmSeg *d_prob = nullptr; int32_t *d_run = nullptr, *d_c0 = nullptr, *d_nch = nullptr, *d_cp = nullptr; int64_t *d_off = nullptr;
            DMG_TRY(to_dev(h, &d_prob, prob)); level_allocs.push_back(d_prob);
            DMG_TRY(to_dev(h, &d_run, prob_run)); level_allocs.push_back(d_run);
            DMG_TRY(to_dev(h, &d_c0, prob_chunk0)); level_allocs.push_back(d_c0);
            DMG_TRY(to_dev(h, &d_nch, prob_nch)); level_allocs.push_back(d_nch);
            DMG_TRY(to_dev(h, &d_cp, chunk_prob)); level_allocs.push_back(d_cp);
            DMG_TRY(to_dev(h, &d_off, d2_off)); level_allocs.push_back(d_off);
            b.prob = d_prob; b.prob_run = d_run; b.prob_chunk0 = d_c0; b.prob_nch = d_nch; b.chunk_prob = d_cp; b.d2_off = d_off;
            double *d_best = nullptr;
            void *blk = nullptr;
            const size_t need = Carver::need({(size_t)n_prob * 2 * E * 8, (size_t)d2_total * 8, (size_t)n_chunk * 2 * E * 8, (size_t)n_chunk * 4 * 8,
                                              (size_t)n_prob * 2 * 8, (size_t)n_prob * 4, (size_t)n_prob * 4, 256, (size_t)n_seg * E * 8});
            DMG_CUDA(h, cudaMalloc(&blk, need)); level_allocs.push_back(blk);
            Carver cw(blk);
            b.cen = cw.take<double>((size_t)n_prob * 2 * E); b.d2 = cw.take<double>(d2_total); b.ps = cw.take<double>((size_t)n_chunk * 2 * E);
            b.pw = cw.take<double>((size_t)n_chunk * 4); b.state = cw.take<double>((size_t)n_prob * 2); b.iter = cw.take<int32_t>(n_prob);
            b.done = cw.take<int32_t>(n_prob); b.pending = cw.take<int32_t>(64); d_best = cw.take<double>((size_t)n_seg * E);
            km_seed_kernel<<<n_chunk, kKmThreads, smem, h->stream>>>(b);
            km_pick_kernel<<<n_prob, 128, 0, h->stream>>>(b, n_prob);
            h->launches += 2;
            for (int it = 0; it <= kKmMaxIter; it++) {
                DMG_CUDA(h, cudaMemsetAsync(b.pending, 0, 4, h->stream));
                km_pass_kernel<<<n_chunk, kKmThreads, smem, h->stream>>>(b);
                km_update_kernel<<<n_prob, 128, 0, h->stream>>>(b, n_prob);
                h->launches += 2;
                int32_t pending = 0;
                DMG_CUDA(h, cudaMemcpyAsync(&pending, b.pending, 4, cudaMemcpyDeviceToHost, h->stream));
                DMG_CUDA(h, cudaStreamSynchronize(h->stream));
                if (!pending) break;
            }
            km_best_kernel<<<n_seg, 64, 0, h->stream>>>(b, n_seg, iters, d_best);
            km_dist_kernel<<<n_chunk, kKmThreads, smem, h->stream>>>(b, iters, d_best, d_dist);
            h->launches += 2;
        }
        DMG_CUDA(h, cudaGetLastError());
        DMG_CUDA(h, cudaMemcpyAsync(dist.data(), d_dist, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
        DMG_CUDA(h, cudaStreamSynchronize(h->stream));
        cleanup();
        nxt.clear();
        for (const std::vector<KmSeg> *lst : {&small, &big})
            for (const KmSeg &g : *lst) {                             // balanceTree (:194-198)
                const int mid = g.len / 2;
                ix.resize(g.len); tmp.resize(g.len);
                for (int i = 0; i < g.len; i++) ix[i] = i;
                if (!arg_partition(dist.data() + g.start, g.len, mid, ix.data())) return fail(h, DMG_ERR_INVALID_ARG, "Nan element detected");
                for (int i = 0; i < g.len; i++) tmp[i] = perm[g.start + ix[i]];
                std::copy(tmp.begin(), tmp.end(), perm.begin() + g.start);
                const int64_t lc = 2 * g.pcode + 1, rc = 2 * g.pcode + 2;
                if (lc > 0x7fffffff / 2) return fail(h, DMG_ERR_UNSUPPORTED, "tree deeper than 30 levels");
                if (mid == 1) out_codes[perm[g.start]] = (int32_t)lc; else nxt.push_back({lc, g.start, mid});
                if (g.len - mid == 1) out_codes[perm[g.start + mid]] = (int32_t)rc; else nxt.push_back({rc, g.start + mid, g.len - mid});
            }
        cur.swap(nxt);
    }
    return DMG_OK;
}
