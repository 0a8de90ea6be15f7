// dr.cu -- Deep Retrieval (K11, K12): layer-wise beam search over K^D paths and the rerank step.
//
// Replaces  CandidateSearcher.beamSearch   deep-retrieval/src/main/scala/com/mass/dr/model/CandidateSearcher.scala:22-60
//           LayerModel.inference           deep-retrieval/.../model/LayerModel.scala:68-84
//           softmax                        deep-retrieval/.../dr/package.scala:23-28
//           RerankModel.inference          deep-retrieval/.../model/RerankModel.scala:43-68
//           DeepRetrieval.recommend        deep-retrieval/.../model/DeepRetrieval.scala:26-46
// All arithmetic is Double, in the strict order of dmg_math.cuh (sequential-k fma chains, the
// spec'd exp, left-to-right softmax sum, divide).  One persistent CTA per user:
//   layer i logits = W_i . [emb(seq) | emb(chosen nodes)] + b_i.  The chain over the first T*E
//   inputs is the same for every live path, so it is computed once per user and layer (U_i) and
//   each path only continues it over its i*E node inputs -- identical bits, (T+i)/i times less work.
//   The beam*K candidate probabilities never get sorted as a whole: a streaming threshold buffer in
//   shared memory keeps every candidate that can still reach the top `beam` under the reference's
//   (probability desc, candidate index asc) order and is trimmed by a bitonic sort when it fills.
#include <algorithm>

#include "device_utils.cuh"
#include "rows_kernels.cuh"
#include "shard_common.cuh"

using namespace dmg;

namespace {

constexpr int kDrMaxD = 8;
constexpr int kDrPC = 16;          // paths per GEMM chunk
constexpr int kDrCap = 2048;       // streaming buffer capacity (power of two)
constexpr int kDrChunk = 1024;     // candidates examined between two capacity checks

struct DrBeamParams {
    int num_item, K, D, T, E, B, beam;
    const double *layer_emb;
    const double *wT[kDrMaxD];     // [in][K]
    const double *b[kDrMaxD];
    const int32_t *seq;            // B x T, -1 = padding
    const double *hist;            // nullable: B x T x E history rows already gathered (sharded item tables)
    const double *node_emb;        // rows of the path nodes: layer_emb + num_item * E when the table is whole
    double *scratch;               // grid x (beam*K)
    int32_t *out_paths;            // B x beam x D
    double *out_probs;             // B x beam
    int32_t *out_counts;
};

using KO = KeyOf<double>;

__device__ __forceinline__ bool key_better(const Key128 &a, const Key128 &b) { return key_less(b, a); }

// sort the buffer, keep the best `beam`, refresh the threshold
__device__ void trim_buffer(Key128 *buf, int *s_count, Key128 *s_tau, int beam)
{
    const int count = *s_count;
    for (int i = threadIdx.x; i < kDrCap; i += blockDim.x)
        if (i >= count) buf[i] = KO::lowest();
    __syncthreads();
    bitonic_sort_desc(buf, kDrCap);
    if (threadIdx.x == 0) {
        const int keep = count < beam ? count : beam;
        *s_count = keep;
        *s_tau = keep >= beam ? buf[beam - 1] : KO::lowest();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kThreads) dr_beam_kernel(const DrBeamParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = p.K, D = p.D, T = p.T, E = p.E, beam = p.beam;
    const int nodeE = (D - 1) * E;
    double *sX0 = reinterpret_cast<double *>(smem_raw);            // T*E
    double *sU = sX0 + T * E;                                      // K
    double *sXn = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(sU + K) + 15) & ~(uintptr_t)15);   // nodeE x kDrPC, read as double2
    double *sSum = sXn + kDrPC * (nodeE > 0 ? nodeE : 1);          // beam
    double *sProb0 = sSum + beam;                                  // beam
    double *sProb1 = sProb0 + beam;                                // beam
    Key128 *sBuf = reinterpret_cast<Key128 *>((reinterpret_cast<uintptr_t>(sProb1 + beam) + 15) & ~(uintptr_t)15);
    Key128 *sTau = sBuf + kDrCap;                                  // 1
    int32_t *sPath0 = reinterpret_cast<int32_t *>(sTau + 1);       // beam x D
    int32_t *sPath1 = sPath0 + beam * D;
    int *sCount = sPath1 + beam * D;
    const int tid = threadIdx.x;
    double *Lmat = p.scratch + (size_t)blockIdx.x * (size_t)beam * K;

    for (int user = blockIdx.x; user < p.B; user += gridDim.x) {
        for (int i = tid; i < T * E; i += kThreads) {
            const int32_t c = p.seq[(size_t)user * T + i / E];
            sX0[i] = p.hist ? p.hist[(size_t)user * T * E + i] : (c < 0 ? 0.0 : p.layer_emb[(size_t)c * E + i % E]);
        }
        double *prob = sProb0, *nprob = sProb1;
        int32_t *path = sPath0, *npath = sPath1;
        int live = 1;
        if (tid == 0) prob[0] = 1.0;
        __syncthreads();

        for (int layer = 0; layer < D; layer++) {
            const double *wT = p.wT[layer], *bias = p.b[layer];
            // (1) user part of the chain, once per user and layer
            for (int o = tid; o < K; o += kThreads) {
                double acc = 0.0;
#pragma unroll 8
                for (int k = 0; k < T * E; k++) acc = fma_(__ldg(wT + (size_t)k * K + o), sX0[k], acc);      // 8 weight loads in flight per thread
                sU[o] = acc;
            }
            __syncthreads();
            // (2) every live path continues the chain over its node inputs
            const int nk = layer * E;
            for (int pb = 0; pb < live; pb += kDrPC) {
                const int np = live - pb < kDrPC ? live - pb : kDrPC;
                for (int i = tid; i < np * nk; i += kThreads) {
                    const int pp = i / nk, k = i % nk, j = k / E;
                    const int32_t row = path[(pb + pp) * D + j] + j * K;                // CandidateSearcher.scala:54 (numItem + j K + c)
                    sXn[k * kDrPC + pp] = p.node_emb[(size_t)row * E + k % E];          // [k][path]: the paths of one k are 8 16-byte loads
                }
                __syncthreads();
                for (int o = tid; o < K; o += kThreads) {
                    double acc[kDrPC];
                    const double u = sU[o];
#pragma unroll
                    for (int pp = 0; pp < kDrPC; pp++) acc[pp] = u;
#pragma unroll 8
                    for (int k = 0; k < nk; k++) {
                        const double w = __ldg(wT + (size_t)(T * E + k) * K + o);
#pragma unroll
                        for (int pp = 0; pp < kDrPC; pp += 2) {
                            const double2 x2 = *reinterpret_cast<const double2 *>(sXn + k * kDrPC + pp);
                            acc[pp] = fma_(w, x2.x, acc[pp]);
                            acc[pp + 1] = fma_(w, x2.y, acc[pp + 1]);
                        }
                    }
                    const double bo = __ldg(bias + o);
#pragma unroll
                    for (int pp = 0; pp < kDrPC; pp++)
                        if (pp < np) Lmat[(size_t)o * beam + (pb + pp)] = add_(acc[pp], bo);       // node-major: Lmat[c][path]
                }
                __syncthreads();
            }
            // (3) softmax per path: max, exp(x - max), left-to-right sum (dr/package.scala:23-28)
            // (thread = path; with the node-major layout the path threads of a warp read consecutive addresses at every step)
            for (int pp = tid; pp < live; pp += kThreads) {
                double *col = Lmat + pp;
                double mx = col[0];
                int c = 1;
                for (; c + 4 <= K; c += 4) {                                   // four loads in flight per step (the max is order-free)
                    const double v0 = col[(size_t)c * beam], v1 = col[(size_t)(c + 1) * beam], v2 = col[(size_t)(c + 2) * beam], v3 = col[(size_t)(c + 3) * beam];
                    mx = v0 > mx ? v0 : mx; mx = v1 > mx ? v1 : mx; mx = v2 > mx ? v2 : mx; mx = v3 > mx ? v3 : mx;
                }
                for (; c < K; c++) { const double v = col[(size_t)c * beam]; mx = v > mx ? v : mx; }
                double sum = 0.0;
                c = 0;
                for (; c + 4 <= K; c += 4) {                                   // loads and exps of four steps together; the sum stays left to right
                    const double v0 = col[(size_t)c * beam], v1 = col[(size_t)(c + 1) * beam], v2 = col[(size_t)(c + 2) * beam], v3 = col[(size_t)(c + 3) * beam];
                    const double e0 = exp_(sub_(v0, mx)), e1 = exp_(sub_(v1, mx)), e2 = exp_(sub_(v2, mx)), e3 = exp_(sub_(v3, mx));
                    col[(size_t)c * beam] = e0; col[(size_t)(c + 1) * beam] = e1; col[(size_t)(c + 2) * beam] = e2; col[(size_t)(c + 3) * beam] = e3;
                    sum = add_(add_(add_(add_(sum, e0), e1), e2), e3);
                }
                for (; c < K; c++) { const double e = exp_(sub_(col[(size_t)c * beam], mx)); col[(size_t)c * beam] = e; sum = add_(sum, e); }
                sSum[pp] = sum;
            }
            if (tid == 0) { *sCount = 0; *sTau = KO::lowest(); }
            __syncthreads();
            // (4) candidates (path, node): probability = parent * (e / sum); keep the best `beam`
            const int total = live * K;
            for (int base = 0; base < total; base += kDrChunk) {
                const Key128 tau = *sTau;
                const int lim = total < base + kDrChunk ? total : base + kDrChunk;
                for (int e0 = base + tid; e0 < lim; e0 += 4 * kThreads) {       // four candidates per thread and step: their loads go out together
                    double lv[4];
                    int cc[4], pq[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int e = e0 + q * kThreads;
                        cc[q] = e < lim ? e / live : 0; pq[q] = e < lim ? e - cc[q] * live : 0;   // visit order is free: the key carries the candidate index
                        lv[q] = Lmat[(size_t)cc[q] * beam + pq[q]];
                    }
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        if (e0 + q * kThreads >= lim) continue;
                        const double cp = mul_(prob[pq[q]], __ddiv_rn(lv[q], sSum[pq[q]]));
                        const Key128 key = KO::make(cp, pq[q] * K + cc[q]);
                        if (key_better(key, tau)) sBuf[atomicAdd(sCount, 1)] = key;
                    }
                }
                __syncthreads();
                const int cnt = *sCount;
                __syncthreads();
                if (cnt > kDrCap - kDrChunk) trim_buffer(sBuf, sCount, sTau, beam);
            }
            trim_buffer(sBuf, sCount, sTau, beam);
            const int nb = *sCount;
            // (5) extend the surviving paths (sorted by probability desc, index asc)
            for (int q = tid; q < nb; q += kThreads) {
                const int pos = KO::pos(sBuf[q]);
                const int pp = pos / K, c = pos % K;
                for (int j = 0; j < layer; j++) npath[q * D + j] = path[pp * D + j];
                npath[q * D + layer] = c;
                nprob[q] = mul_(prob[pp], __ddiv_rn(Lmat[(size_t)c * beam + pp], sSum[pp]));
            }
            __syncthreads();
            { double *t = prob; prob = nprob; nprob = t; }
            { int32_t *t = path; path = npath; npath = t; }
            live = nb;
        }
        for (int i = tid; i < beam * D; i += kThreads)
            p.out_paths[(size_t)user * beam * D + i] = i < live * D ? path[i] : -1;
        for (int i = tid; i < beam; i += kThreads) p.out_probs[(size_t)user * beam + i] = i < live ? prob[i] : 0.0;
        if (tid == 0) p.out_counts[user] = live;
        __syncthreads();
    }
}

struct DrRerankParams {
    int num_item, K, D, T, E, B, beam, topk;
    const double *rr_emb, *rr_wT, *rr_b, *sm_w, *sm_b;
    const int32_t *seq;
    const int32_t *paths;          // B x beam x D
    const int32_t *path_counts;    // B
    const int64_t *path_off;
    const int32_t *path_items;
    const double *pre_scores;      // nullable: candidate scores computed by the owners of the items (sharded tables) ...
    const int32_t *cand_off;       // ... candidate idx of user u at pre_scores[cand_off[u] + idx]
    int32_t *out_items;            // B x topk
    double *out_scores;
    int32_t *out_counts;
};

__global__ void __launch_bounds__(kThreads) dr_rerank_kernel(const DrRerankParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = p.T, E = p.E, D = p.D, K = p.K, beam = p.beam;
    double *sX = reinterpret_cast<double *>(smem_raw);             // T*E
    double *sUv = sX + T * E;                                      // E
    Key128 *sBuf = reinterpret_cast<Key128 *>((reinterpret_cast<uintptr_t>(sUv + E) + 15) & ~(uintptr_t)15);
    Key128 *sTau = sBuf + kDrCap;
    int64_t *sStart = reinterpret_cast<int64_t *>(sTau + 1);       // beam + 1 (candidate index where a path starts)
    int64_t *sBase = sStart + beam + 1;                            // beam (offset of the path's items)
    int *sCount = reinterpret_cast<int *>(sBase + beam);
    const int tid = threadIdx.x;

    for (int user = blockIdx.x; user < p.B; user += gridDim.x) {
        for (int i = tid; i < T * E && !p.pre_scores; i += kThreads) {
            const int32_t c = p.seq[(size_t)user * T + i / E];
            sX[i] = c < 0 ? 0.0 : p.rr_emb[(size_t)c * E + i % E];
        }
        __syncthreads();
        // user vector u = W_r x + b_r (RerankModel.inferenceUserVector :54-68)
        for (int o = tid; o < E && !p.pre_scores; o += kThreads) {
            double acc = 0.0;
            for (int k = 0; k < T * E; k++) acc = fma_(__ldg(p.rr_wT + (size_t)k * E + o), sX[k], acc);
            sUv[o] = add_(acc, __ldg(p.rr_b + o));
        }
        // candidate list = items of each surviving path, in beam order (searchCandidate :8-20)
        const int np = p.path_counts[user];
        if (tid == 0) {
            int64_t run = 0;
            for (int q = 0; q < np; q++) {
                int64_t key = 0;
                for (int d = 0; d < D; d++) key = key * K + p.paths[((size_t)user * beam + q) * D + d];
                sStart[q] = run;
                sBase[q] = p.path_off[key];
                run += p.path_off[key + 1] - p.path_off[key];
            }
            sStart[np] = run;
            *sCount = 0;
            *sTau = KO::lowest();
        }
        __syncthreads();
        const int64_t total = sStart[np];
        const int keep = p.topk < kDrCap ? p.topk : kDrCap;
        for (int64_t base = 0; base < total; base += kDrChunk) {
            const Key128 tau = *sTau;
            for (int64_t idx = base + tid; idx < total && idx < base + kDrChunk; idx += kThreads) {
                int lo = 0, hi = np;                                // last q with sStart[q] <= idx
                while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (sStart[mid] <= idx) lo = mid; else hi = mid; }
                double sc;
                if (p.pre_scores) sc = p.pre_scores[(size_t)p.cand_off[user] + idx];
                else {
                    const int32_t item = p.path_items[sBase[lo] + (idx - sStart[lo])];
                    const double *w = p.sm_w + (size_t)item * E;
                    double acc = 0.0;
                    for (int k = 0; k < E; k++) acc = fma_(__ldg(w + k), sUv[k], acc);
                    sc = add_(acc, __ldg(p.sm_b + item));
                }
                const Key128 key = KO::make(sc, (int)idx);
                if (key_better(key, tau)) sBuf[atomicAdd(sCount, 1)] = key;
            }
            __syncthreads();
            const int cnt = *sCount;
            __syncthreads();
            if (cnt > kDrCap - kDrChunk) trim_buffer(sBuf, sCount, sTau, keep);
        }
        trim_buffer(sBuf, sCount, sTau, keep);
        const int n = *sCount;
        for (int i = tid; i < p.topk; i += kThreads) {
            int32_t item = -1;
            double sc = 0.0;
            if (i < n) {
                const int64_t idx = KO::pos(sBuf[i]);
                int lo = 0, hi = np;
                while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (sStart[mid] <= idx) lo = mid; else hi = mid; }
                item = p.path_items[sBase[lo] + (idx - sStart[lo])];
                if (p.pre_scores) sc = p.pre_scores[(size_t)p.cand_off[user] + idx];
                else {
                    const double *w = p.sm_w + (size_t)item * E;
                    double acc = 0.0;
                    for (int k = 0; k < E; k++) acc = fma_(__ldg(w + k), sUv[k], acc);
                    sc = add_(acc, __ldg(p.sm_b + item));
                }
            }
            p.out_items[(size_t)user * p.topk + i] = item;
            p.out_scores[(size_t)user * p.topk + i] = sc;
        }
        if (tid == 0) p.out_counts[user] = n;
        __syncthreads();
    }
}

int32_t up(dmg_handle_t h, double **dst, const double *src, size_t n)
{
    DMG_CUDA(h, cudaMalloc(dst, n * sizeof(double)));
    DMG_CUDA(h, cudaMemcpyAsync(*dst, src, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    return DMG_OK;
}

}  // namespace

void dmg_free_dr(DrDev &d)
{
    cudaFree(d.d_layer_emb);
    for (double *p : d.d_layer_w) cudaFree(p);
    for (double *p : d.d_layer_b) cudaFree(p);
    for (double *p : d.d_layer_wT) cudaFree(p);
    cudaFree(d.d_rr_emb); cudaFree(d.d_rr_w); cudaFree(d.d_rr_b); cudaFree(d.d_sm_w); cudaFree(d.d_sm_b);
    cudaFree(d.d_path_off); cudaFree(d.d_path_items);
    for (double *p : d.tr_g) cudaFree(p);
    for (double *p : d.tr_s) cudaFree(p);
    for (double *p : d.tr_r) cudaFree(p);
    cudaFree(d.d_item_paths);
    d = DrDev();
}

DMG_API int32_t dmg_dr_load(dmg_handle_t h, int32_t num_item, int32_t K, int32_t D, int32_t T, int32_t E,
                            const double *layer_emb, const double *const *layer_w, const double *const *layer_b,
                            const double *rr_emb, const double *rr_w, const double *rr_b, const double *sm_w,
                            const double *sm_b)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    DMG_TRY(model_is_shared(h, "dmg_dr_load"));
    if (num_item <= 0 || K <= 0 || T <= 0 || E <= 0 || !layer_emb || !layer_w || !layer_b || !rr_emb || !rr_w || !rr_b ||
        !sm_w || !sm_b)
        return fail(h, DMG_ERR_INVALID_ARG, "dmg_dr_load: bad arguments");
    if (D < 2) return fail(h, DMG_ERR_INVALID_ARG, "number of layers must be at least 2");   // LayerModel.scala:24
    if (D > kDrMaxD) return fail(h, DMG_ERR_UNSUPPORTED, "at most %d layers", kDrMaxD);
    DMG_CUDA(h, cudaSetDevice(h->device));
    dmg_free_dr(h->dr);
    DrDev &d = h->dr;
    d.num_item = num_item; d.K = K; d.D = D; d.T = T; d.E = E;
    const size_t emb_rows = (size_t)num_item + (size_t)K * (D - 1);
    DMG_TRY(up(h, &d.d_layer_emb, layer_emb, emb_rows * E));
    for (int i = 0; i < D; i++) {
        const int in = (T + i) * E;
        double *w = nullptr, *b = nullptr, *wT = nullptr;
        DMG_TRY(up(h, &w, layer_w[i], (size_t)K * in));
        DMG_TRY(up(h, &b, layer_b[i], (size_t)K));
        DMG_CUDA(h, cudaMalloc(&wT, (size_t)K * in * sizeof(double)));
        transpose_kernel<double><<<(K * in + 255) / 256, 256, 0, h->stream>>>(w, wT, K, in);
        h->launches += 1;
        d.d_layer_w.push_back(w); d.d_layer_b.push_back(b); d.d_layer_wT.push_back(wT);
    }
    DMG_TRY(up(h, &d.d_rr_emb, rr_emb, (size_t)num_item * E));
    double *rr_tmp = nullptr;
    DMG_TRY(up(h, &rr_tmp, rr_w, (size_t)E * T * E));
    DMG_CUDA(h, cudaMalloc(&d.d_rr_w, (size_t)E * T * E * sizeof(double)));                 // kept transposed [T*E][E]
    transpose_kernel<double><<<(E * T * E + 255) / 256, 256, 0, h->stream>>>(rr_tmp, d.d_rr_w, E, T * E);
    h->launches += 1;
    DMG_CUDA(h, cudaGetLastError());
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(rr_tmp);
    DMG_TRY(up(h, &d.d_rr_b, rr_b, (size_t)E));
    DMG_TRY(up(h, &d.d_sm_w, sm_w, (size_t)num_item * E));
    DMG_TRY(up(h, &d.d_sm_b, sm_b, (size_t)num_item));
    d.loaded = true;
    return DMG_OK;
}

DMG_API int32_t dmg_dr_load_paths(dmg_handle_t h, const int64_t *path_off, const int32_t *path_items)
{
    if (!h || !path_off) return DMG_ERR_INVALID_ARG;
    DMG_TRY(model_is_shared(h, "dmg_dr_load_paths"));
    DrDev &d = h->dr;
    if (!d.loaded) return fail(h, DMG_ERR_STATE, "dmg_dr_load first");
    double nk = 1;
    for (int i = 0; i < d.D; i++) nk *= d.K;
    if (nk > 2.0e9) return fail(h, DMG_ERR_UNSUPPORTED, "K^D = %.3g path keys: dense CSR too large", nk);
    const int64_t n_keys = (int64_t)nk;
    const int64_t n_items = path_off[n_keys];
    if (n_items < 0 || (n_items > 0 && !path_items)) return fail(h, DMG_ERR_INVALID_ARG, "bad path CSR");
    DMG_CUDA(h, cudaSetDevice(h->device));
    cudaFree(d.d_path_off); cudaFree(d.d_path_items);
    d.d_path_off = nullptr; d.d_path_items = nullptr;
    DMG_CUDA(h, cudaMalloc(&d.d_path_off, (size_t)(n_keys + 1) * sizeof(int64_t)));
    DMG_CUDA(h, cudaMalloc(&d.d_path_items, (size_t)std::max<int64_t>(n_items, 1) * sizeof(int32_t)));
    DMG_CUDA(h, cudaMemcpyAsync(d.d_path_off, path_off, (size_t)(n_keys + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    if (n_items)
        DMG_CUDA(h, cudaMemcpyAsync(d.d_path_items, path_items, (size_t)n_items * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    d.paths_loaded = true;
    return DMG_OK;
}

// shared by beam_search / retrieve: runs the beam kernel, leaves paths/probs/counts on the device
static int32_t dr_beam_enqueue(dmg_handle_t h, int32_t B, const int32_t *seq_host, int32_t beam, bool for_rerank,
                               int32_t **d_paths, double **d_probs, int32_t **d_counts, int32_t **d_seq_out,
                               const double *hist_tiles = nullptr)
{
    DrDev &d = h->dr;
    if (!d.loaded) return fail(h, DMG_ERR_STATE, "dmg_dr_load first");
    if (d.sharded && !hist_tiles) return fail(h, DMG_ERR_STATE, "the Deep Retrieval item tables are sharded: use dmg_shard_dr_retrieve");
    if (B <= 0 || beam <= 0 || !seq_host) return fail(h, DMG_ERR_INVALID_ARG, "bad arguments");
    if ((double)beam * d.K > 2.0e9) return fail(h, DMG_ERR_UNSUPPORTED, "beam*K too large");
    if (beam > kDrCap - kDrChunk) return fail(h, DMG_ERR_UNSUPPORTED, "beam must be <= %d", kDrCap - kDrChunk);
    DMG_CUDA(h, cudaSetDevice(h->device));
    const int T = d.T, E = d.E, K = d.K, D = d.D;
    const size_t b_seq = (size_t)B * T * 4;
    DMG_TRY(ensure_host(h, h->s_in, b_seq));
    DMG_TRY(ensure_dev(h, h->s_in, b_seq));
    memcpy(h->s_in.h, seq_host, b_seq);
    DMG_CUDA(h, cudaMemcpyAsync(h->s_in.d, h->s_in.h, b_seq, cudaMemcpyHostToDevice, h->stream));
    int32_t *d_seq = (int32_t *)h->s_in.d;
    const int64_t emb_rows = (int64_t)d.num_item + (int64_t)K * (D - 1);
    // LayerModel.inference indexes the shared table [0, numItem + K(D-1)); RerankModel only [0, numItem)
    check_index_kernel<<<(unsigned)(((int64_t)B * T + 255) / 256), 256, 0, h->stream>>>(
        d_seq, (int64_t)B * T, for_rerank ? (int64_t)d.num_item : emb_rows, h->d_flags);
    h->launches += 1;
    const int nodeE = std::max((D - 1) * E, 1);
    size_t smem = ((size_t)T * E + K + 1 + (size_t)kDrPC * nodeE + 3 * (size_t)beam + 2) * 8 + (size_t)kDrCap * 16 +
                  (size_t)2 * beam * D * 4 + 64 * 4 + 32;
    if (smem > h->smem_optin)
        return fail(h, DMG_ERR_UNSUPPORTED, "Deep Retrieval shape needs %zu B of shared memory (limit %zu)", smem, h->smem_optin);
    DMG_CUDA(h, cudaFuncSetAttribute(dr_beam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int resident = 1;                                            // persistent CTAs: as many per SM as registers and shared memory allow
    DMG_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, dr_beam_kernel, kThreads, smem));
    const int grid = std::min(B, h->sm_count * std::max(resident, 1));
    const size_t scratch = (size_t)grid * beam * K * sizeof(double);
    const size_t work = Carver::need({scratch, (size_t)B * beam * D * 4, (size_t)B * beam * 8, (size_t)B * 4});
    DMG_TRY(ensure_dev(h, h->s_work, work));
    Carver cw(h->s_work.d);
    double *d_scr = cw.take<double>((size_t)grid * beam * K);
    *d_paths = cw.take<int32_t>((size_t)B * beam * D);
    *d_probs = cw.take<double>((size_t)B * beam);
    *d_counts = cw.take<int32_t>(B);
    *d_seq_out = d_seq;
    DrBeamParams p;
    memset(&p, 0, sizeof(p));
    p.num_item = d.num_item; p.K = K; p.D = D; p.T = T; p.E = E; p.B = B; p.beam = beam;
    p.layer_emb = d.d_layer_emb;
    p.hist = hist_tiles;
    p.node_emb = d.d_layer_emb + (size_t)(hist_tiles ? d.local_items : d.num_item) * E;
    for (int i = 0; i < D; i++) { p.wT[i] = d.d_layer_wT[i]; p.b[i] = d.d_layer_b[i]; }
    p.seq = d_seq; p.scratch = d_scr; p.out_paths = *d_paths; p.out_probs = *d_probs; p.out_counts = *d_counts;
    dr_beam_kernel<<<grid, kThreads, smem, h->stream>>>(p);
    h->launches += 1;
    DMG_CUDA(h, cudaGetLastError());
    return DMG_OK;
}

static int32_t dr_check_flag(dmg_handle_t h)
{
    int32_t flag = 0;
    DMG_CUDA(h, cudaMemcpyAsync(&flag, h->d_flags, 4, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    if (flag) {
        DMG_CUDA(h, cudaMemsetAsync(h->d_flags, 0, 4, h->stream));
        return fail(h, DMG_ERR_INDEX, "Deep Retrieval: history index outside [-1, num_item)");
    }
    return DMG_OK;
}

DMG_API int32_t dmg_dr_beam_search(dmg_handle_t h, int32_t B, const int32_t *seq, int32_t beam, int32_t *out_paths,
                                   double *out_probs, int32_t *out_counts)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    if (!out_paths || !out_probs || !out_counts) return fail(h, DMG_ERR_INVALID_ARG, "null output");
    int32_t *d_paths, *d_counts, *d_seq;
    double *d_probs;
    DMG_TRY(dr_beam_enqueue(h, B, seq, beam, false, &d_paths, &d_probs, &d_counts, &d_seq));
    const int D = h->dr.D;
    DMG_TRY(dr_check_flag(h));
    DMG_CUDA(h, cudaMemcpyAsync(out_paths, d_paths, (size_t)B * beam * D * 4, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaMemcpyAsync(out_probs, d_probs, (size_t)B * beam * 8, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaMemcpyAsync(out_counts, d_counts, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    return DMG_OK;
}

DMG_API int32_t dmg_dr_retrieve(dmg_handle_t h, int32_t B, const int32_t *seq, int32_t beam, int32_t topk,
                                int32_t *out_items, double *out_scores, int32_t *out_counts)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    if (!out_items || !out_scores || !out_counts || topk <= 0) return fail(h, DMG_ERR_INVALID_ARG, "bad arguments");
    if (!h->dr.paths_loaded) return fail(h, DMG_ERR_STATE, "dmg_dr_load_paths first");
    if (topk > kDrCap - kDrChunk) return fail(h, DMG_ERR_UNSUPPORTED, "topk must be <= %d", kDrCap - kDrChunk);
    int32_t *d_paths, *d_counts, *d_seq;
    double *d_probs;
    DMG_TRY(dr_beam_enqueue(h, B, seq, beam, true, &d_paths, &d_probs, &d_counts, &d_seq));
    DrDev &d = h->dr;
    const size_t out_bytes = Carver::need({(size_t)B * topk * 4, (size_t)B * topk * 8, (size_t)B * 4});
    DMG_TRY(ensure_dev(h, h->s_out, out_bytes));
    Carver od(h->s_out.d);
    int32_t *d_items = od.take<int32_t>((size_t)B * topk);
    double *d_sc = od.take<double>((size_t)B * topk);
    int32_t *d_cnt = od.take<int32_t>(B);
    DrRerankParams p;
    memset(&p, 0, sizeof(p));
    p.num_item = d.num_item; p.K = d.K; p.D = d.D; p.T = d.T; p.E = d.E; p.B = B; p.beam = beam; p.topk = topk;
    p.rr_emb = d.d_rr_emb; p.rr_wT = d.d_rr_w; p.rr_b = d.d_rr_b; p.sm_w = d.d_sm_w; p.sm_b = d.d_sm_b;
    p.seq = d_seq; p.paths = d_paths; p.path_counts = d_counts; p.path_off = d.d_path_off; p.path_items = d.d_path_items;
    p.out_items = d_items; p.out_scores = d_sc; p.out_counts = d_cnt;
    size_t smem = ((size_t)d.T * d.E + d.E + 2) * 8 + (size_t)kDrCap * 16 + ((size_t)2 * beam + 2) * 8 + 64 * 4 + 32;
    if (smem > h->smem_optin) return fail(h, DMG_ERR_UNSUPPORTED, "rerank needs %zu B of shared memory", smem);
    DMG_CUDA(h, cudaFuncSetAttribute(dr_rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dr_rerank_kernel<<<std::min(B, h->sm_count * 2), kThreads, smem, h->stream>>>(p);
    h->launches += 1;
    DMG_CUDA(h, cudaGetLastError());
    DMG_TRY(dr_check_flag(h));
    DMG_CUDA(h, cudaMemcpyAsync(out_items, d_items, (size_t)B * topk * 4, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaMemcpyAsync(out_scores, d_sc, (size_t)B * topk * 8, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaMemcpyAsync(out_counts, d_cnt, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    return DMG_OK;
}

// ---- Deep Retrieval with the item-indexed tables sharded over the GPUs of one box (BASELINE config 5) --------
// layer-embedding item rows, rerank embedding, softmax weights and biases are split by contiguous item-id range (rank r owns
// items [r * chunk, (r+1) * chunk)); the (D-1) K path-node rows and every Linear are replicated.  Per batch:
//   history rows of both embedding tables: each rank fills the rows it owns into zero [G*B, T, E] buffers, integer-sum
//   all-reduce (ncclUint64: x + 0 keeps every bit) -> every rank holds its users' inputs;
//   beam search over the K^D paths and the rerank user vector run locally (dr_beam_kernel with `hist`);
//   rerank candidates (candidate index, item) go to the owner of the item (ncclSend/ncclRecv), which answers
//   softmaxW[item].u + bias[item] (user vectors all-gathered once, E doubles per user); scores come back (8 B) and the
//   requester picks its topk by the reference's (score desc, candidate index asc) order.
// Same arithmetic and order as the unsharded kernels, so results are bit-identical to dmg_dr_retrieve.
namespace {

__global__ void dr_fill_tiles_kernel(const double *__restrict__ table, int64_t item_base, int64_t local_items, const int32_t *__restrict__ seq_all,
                                     int64_t n_slots, int E, unsigned long long *__restrict__ tiles)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots * E; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t sl = i / E;
        const int k = (int)(i % E);
        const int64_t c = seq_all[sl];
        unsigned long long v = 0ull;
        if (c >= item_base && c < item_base + local_items) v = (unsigned long long)__double_as_longlong(table[(size_t)(c - item_base) * E + k]);
        tiles[i] = v;
    }
}

// u = W_r x + b_r (RerankModel.inferenceUserVector :54-68) from the gathered rerank-embedding tile
__global__ void __launch_bounds__(kThreads) dr_uservec_kernel(int B, int T, int E, const double *__restrict__ hist_rr, const double *__restrict__ rr_wT,
                                                              const double *__restrict__ rr_b, double *__restrict__ uvec)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sX = reinterpret_cast<double *>(smem_raw);
    for (int user = blockIdx.x; user < B; user += gridDim.x) {
        __syncthreads();
        for (int i = threadIdx.x; i < T * E; i += kThreads) sX[i] = hist_rr[(size_t)user * T * E + i];
        __syncthreads();
        for (int o = threadIdx.x; o < E; o += kThreads) {
            double acc = 0.0;
            for (int k = 0; k < T * E; k++) acc = fma_(__ldg(rr_wT + (size_t)k * E + o), sX[k], acc);
            uvec[(size_t)user * E + o] = add_(acc, __ldg(rr_b + o));
        }
    }
}

__device__ __forceinline__ int64_t dr_path_key(const int32_t *path, int D, int K)
{
    int64_t key = 0;
    for (int d = 0; d < D; d++) key = key * K + path[d];
    return key;
}
// candidates of a user = items of each surviving path, in beam order (searchCandidate :8-20): count them ...
__global__ void dr_cand_count_kernel(int B, int beam, int D, int K, const int32_t *__restrict__ paths, const int32_t *__restrict__ path_counts,
                                     const int64_t *__restrict__ path_off, int32_t *__restrict__ totals)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= B) return;
    int64_t run = 0;
    for (int q = 0; q < path_counts[u]; q++) {
        const int64_t key = dr_path_key(paths + ((size_t)u * beam + q) * D, D, K);
        run += path_off[key + 1] - path_off[key];
    }
    totals[u] = (int32_t)run;
}
// ... and list them: cand_item[cand_off[u] + idx]
__global__ void __launch_bounds__(kThreads) dr_cand_fill_kernel(int B, int beam, int D, int K, const int32_t *__restrict__ paths,
                                                                const int32_t *__restrict__ path_counts, const int64_t *__restrict__ path_off,
                                                                const int32_t *__restrict__ path_items, const int32_t *__restrict__ cand_off,
                                                                int32_t *__restrict__ cand_item)
{
    for (int u = blockIdx.x; u < B; u += gridDim.x) {
        int64_t run = cand_off[u];
        for (int q = 0; q < path_counts[u]; q++) {               // every thread walks the (short) path list, the items are split
            const int64_t key = dr_path_key(paths + ((size_t)u * beam + q) * D, D, K);
            const int64_t b0 = path_off[key], n = path_off[key + 1] - b0;
            for (int64_t i = threadIdx.x; i < n; i += kThreads) cand_item[run + i] = path_items[b0 + i];
            run += n;
        }
    }
}
// requests (candidate index, item) appended to the region of the item's owner, warp-aggregated
__global__ void dr_bucket_kernel(int n, const int32_t *__restrict__ cand_item, int64_t chunk, int2 *__restrict__ region, int64_t region_stride,
                                 int32_t *__restrict__ n_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = i < n;
    const int item = ok ? cand_item[i] : 0;
    const int o = ok ? (int)(item / chunk) : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, o);
    const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
    int base = 0;
    if (ok && lane == leader) base = atomicAdd(&n_out[o], __popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (ok) region[(size_t)o * region_stride + base + __popc(peers & ((1u << lane) - 1u))] = make_int2(i, item);
}
// owner: softmaxW[item].u + bias[item] (RerankModel.inference :43-52) for the requests of requester `p`
__global__ void dr_score_cands_kernel(int n, const int2 *__restrict__ req, const int32_t *__restrict__ cand_off_p /* B + 1 of the requester */,
                                      int B, int E, const double *__restrict__ uvec_p /* B x E of the requester */, int64_t item_base,
                                      const double *__restrict__ sm_w, const double *__restrict__ sm_b, double *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int gci = req[i].x;
    int lo = 0, hi = B;                                          // user of the candidate: last u with cand_off[u] <= gci
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (cand_off_p[mid] <= gci) lo = mid; else hi = mid; }
    const int64_t li = (int64_t)req[i].y - item_base;
    const double *w = sm_w + (size_t)li * E, *u = uvec_p + (size_t)lo * E;
    double acc = 0.0;
    for (int k = 0; k < E; k++) acc = fma_(__ldg(w + k), u[k], acc);
    out[i] = add_(acc, __ldg(sm_b + li));
}
__global__ void dr_scatter_kernel(int n, const int2 *__restrict__ req, const double *__restrict__ reply, double *__restrict__ score)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) score[req[i].x] = reply[i];
}

}  // namespace

// Item-indexed tables are passed WHOLE (layer_emb: num_item + K (D-1) rows, rr_emb / sm_w: num_item rows, sm_b: num_item);
// only this rank's item range is uploaded.  Call after dmg_shard_init, on every rank.
DMG_API int32_t dmg_shard_dr_load(dmg_handle_t h, int32_t num_item, int32_t K, int32_t D, int32_t T, int32_t E,
                                  const double *layer_emb, const double *const *layer_w, const double *const *layer_b,
                                  const double *rr_emb, const double *rr_w, const double *rr_b, const double *sm_w, const double *sm_b)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    DMG_TRY(model_is_shared(h, "dmg_shard_dr_load"));
    ShardState *s = h->shard;
    if (!s) return fail(h, DMG_ERR_STATE, "call dmg_shard_init first");
    if (num_item <= 0 || K <= 0 || T <= 0 || E <= 0 || !layer_emb || !layer_w || !layer_b || !rr_emb || !rr_w || !rr_b || !sm_w || !sm_b)
        return fail(h, DMG_ERR_INVALID_ARG, "dmg_shard_dr_load: bad arguments");
    if (D < 2 || D > kDrMaxD) return fail(h, DMG_ERR_INVALID_ARG, "number of layers must be in [2, %d]", kDrMaxD);
    DMG_CUDA(h, cudaSetDevice(h->device));
    dmg_free_dr(h->dr);
    DrDev &d = h->dr;
    d.num_item = num_item; d.K = K; d.D = D; d.T = T; d.E = E;
    d.item_chunk = ((int64_t)num_item + s->world - 1) / s->world;
    d.item_base = std::min<int64_t>((int64_t)s->rank * d.item_chunk, num_item);
    d.local_items = std::min<int64_t>(d.item_chunk, num_item - d.item_base);
    d.sharded = true;
    const size_t node_rows = (size_t)K * (D - 1);
    // local layer table = [owned item rows | path-node rows]
    DMG_CUDA(h, cudaMalloc(&d.d_layer_emb, ((size_t)d.local_items + node_rows) * E * sizeof(double)));
    DMG_CUDA(h, cudaMemcpyAsync(d.d_layer_emb, layer_emb + (size_t)d.item_base * E, (size_t)d.local_items * E * 8, cudaMemcpyHostToDevice, h->stream));
    DMG_CUDA(h, cudaMemcpyAsync(d.d_layer_emb + (size_t)d.local_items * E, layer_emb + (size_t)num_item * E, node_rows * E * 8,
                                cudaMemcpyHostToDevice, h->stream));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int i = 0; i < D; i++) {
        const int in = (T + i) * E;
        double *w = nullptr, *b = nullptr, *wT = nullptr;
        DMG_TRY(up(h, &w, layer_w[i], (size_t)K * in));
        DMG_TRY(up(h, &b, layer_b[i], (size_t)K));
        DMG_CUDA(h, cudaMalloc(&wT, (size_t)K * in * sizeof(double)));
        transpose_kernel<double><<<(K * in + 255) / 256, 256, 0, h->stream>>>(w, wT, K, in);
        h->launches += 1;
        d.d_layer_w.push_back(w); d.d_layer_b.push_back(b); d.d_layer_wT.push_back(wT);
    }
    const size_t li = (size_t)std::max<int64_t>(d.local_items, 1);
    DMG_TRY(up(h, &d.d_rr_emb, rr_emb + (size_t)d.item_base * E, li * E));
    double *rr_tmp = nullptr;
    DMG_TRY(up(h, &rr_tmp, rr_w, (size_t)E * T * E));
    DMG_CUDA(h, cudaMalloc(&d.d_rr_w, (size_t)E * T * E * sizeof(double)));                 // kept transposed [T*E][E]
    transpose_kernel<double><<<(E * T * E + 255) / 256, 256, 0, h->stream>>>(rr_tmp, d.d_rr_w, E, T * E);
    h->launches += 1;
    DMG_CUDA(h, cudaGetLastError());
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(rr_tmp);
    DMG_TRY(up(h, &d.d_rr_b, rr_b, (size_t)E));
    DMG_TRY(up(h, &d.d_sm_w, sm_w + (size_t)d.item_base * E, li * E));
    DMG_TRY(up(h, &d.d_sm_b, sm_b + (size_t)d.item_base, li));
    d.loaded = true;
    return DMG_OK;
}

// DeepRetrieval.recommend for this rank's B users; collective (same B, beam, topk on every rank).
DMG_API int32_t dmg_shard_dr_retrieve(dmg_handle_t h, int32_t B, const int32_t *seq, int32_t beam, int32_t topk,
                                      int32_t *out_items, double *out_scores, int32_t *out_counts)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    ShardState *s = h->shard;
    DrDev &d = h->dr;
    if (!s || !d.loaded || !d.sharded) return fail(h, DMG_ERR_STATE, "dmg_shard_init and dmg_shard_dr_load first");
    if (!d.paths_loaded) return fail(h, DMG_ERR_STATE, "dmg_dr_load_paths first");
    if (B <= 0 || beam <= 0 || topk <= 0 || !seq || !out_items || !out_scores || !out_counts) return fail(h, DMG_ERR_INVALID_ARG, "bad arguments");
    if (topk > kDrCap - kDrChunk) return fail(h, DMG_ERR_UNSUPPORTED, "topk must be <= %d", kDrCap - kDrChunk);
    DMG_CUDA(h, cudaSetDevice(h->device));
    const int G = s->world, T = d.T, E = d.E, D = d.D, K = d.K;
    const int64_t BU = (int64_t)G * B;
    cudaStream_t st = h->stream;
    for (int64_t i = 0; i < (int64_t)B * T; i++)
        if (seq[i] < -1 || seq[i] >= d.num_item) return fail(h, DMG_ERR_INDEX, "Deep Retrieval: history index outside [-1, num_item)");
    // ---- histories of every rank, both embedding tiles ---------------------------------------------------------------
    Scratch &sb = s->buf;
    const size_t fixed = Carver::need({(size_t)BU * T * 4, (size_t)BU * T * E * 8, (size_t)BU * T * E * 8, (size_t)BU * E * 8,
                                       (size_t)B * 4, (size_t)G * (B + 1) * 4, (size_t)G * 4, (size_t)G * G * 4});
    DMG_TRY(ensure_dev(h, sb, fixed));
    DMG_TRY(ensure_host(h, sb, (size_t)(B + 1) * 8 + (size_t)G * G * 4 + 1024));
    Carver cd(sb.d);
    int32_t *d_seq_all = cd.take<int32_t>((size_t)BU * T);
    double *d_tile_l = cd.take<double>((size_t)BU * T * E), *d_tile_r = cd.take<double>((size_t)BU * T * E);
    double *d_uvec = cd.take<double>((size_t)BU * E);
    int32_t *d_totals = cd.take<int32_t>((size_t)B);
    int32_t *d_off_all = cd.take<int32_t>((size_t)G * (B + 1));
    int32_t *d_nreq = cd.take<int32_t>((size_t)G), *d_matrix = cd.take<int32_t>((size_t)G * G);
    int32_t *h_off = (int32_t *)sb.h;                               // B + 1
    int32_t *h_matrix = h_off + ((B + 1 + 63) & ~63);
    int32_t *d_seq_mine = d_seq_all + (size_t)s->rank * B * T;
    DMG_CUDA(h, cudaMemcpyAsync(d_seq_mine, seq, (size_t)B * T * 4, cudaMemcpyHostToDevice, st));
    DMG_CUDA(h, cudaStreamSynchronize(st));                        // `seq` is the caller's
    if (G > 1) DMG_NCCL(h, g_nccl.AllGather(d_seq_mine, d_seq_all, (size_t)B * T, ncclInt32, s->comm, st));
    dr_fill_tiles_kernel<<<h->sm_count * 4, 256, 0, st>>>(d.d_layer_emb, d.item_base, d.local_items, d_seq_all, BU * T, E, (unsigned long long *)d_tile_l);
    dr_fill_tiles_kernel<<<h->sm_count * 4, 256, 0, st>>>(d.d_rr_emb, d.item_base, d.local_items, d_seq_all, BU * T, E, (unsigned long long *)d_tile_r);
    h->launches += 2;
    if (G > 1) {
        DMG_NCCL(h, g_nccl.AllReduce(d_tile_l, d_tile_l, (size_t)BU * T * E, ncclUint64, ncclSum, s->comm, st));
        DMG_NCCL(h, g_nccl.AllReduce(d_tile_r, d_tile_r, (size_t)BU * T * E, ncclUint64, ncclSum, s->comm, st));
    }
    const double *hist_l = d_tile_l + (size_t)s->rank * B * T * E, *hist_r = d_tile_r + (size_t)s->rank * B * T * E;
    // ---- beam search (local), user vectors (local, then replicated) ------------------------------------------------------
    int32_t *d_paths, *d_counts, *d_seq_unused;
    double *d_probs;
    DMG_TRY(dr_beam_enqueue(h, B, seq, beam, true, &d_paths, &d_probs, &d_counts, &d_seq_unused, hist_l));
    double *d_u_mine = d_uvec + (size_t)s->rank * B * E;
    DMG_CUDA(h, cudaFuncSetAttribute(dr_uservec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)T * E * 8)));
    dr_uservec_kernel<<<std::min(B, h->sm_count * 4), kThreads, (size_t)T * E * 8, st>>>(B, T, E, hist_r, d.d_rr_w, d.d_rr_b, d_u_mine);
    h->launches += 1;
    if (G > 1) DMG_NCCL(h, g_nccl.AllGather(d_u_mine, d_uvec, (size_t)B * E, ncclFloat64, s->comm, st));
    // ---- candidates ----------------------------------------------------------------------------------------------------------
    dr_cand_count_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, beam, D, K, d_paths, d_counts, d.d_path_off, d_totals);
    h->launches += 1;
    DMG_CUDA(h, cudaMemcpyAsync(h_off + 1, d_totals, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    DMG_CUDA(h, cudaStreamSynchronize(st));
    h_off[0] = 0;
    int64_t total64 = 0;
    for (int u = 0; u < B; u++) { total64 += h_off[u + 1]; if (total64 > 0x7fffffff) return fail(h, DMG_ERR_UNSUPPORTED, "too many rerank candidates in one batch"); h_off[u + 1] = (int32_t)total64; }
    const int total = (int)total64;
    int32_t *d_off_mine = d_off_all + (size_t)s->rank * (B + 1);
    DMG_CUDA(h, cudaMemcpyAsync(d_off_mine, h_off, (size_t)(B + 1) * 4, cudaMemcpyHostToDevice, st));
    if (G > 1) DMG_NCCL(h, g_nccl.AllGather(d_off_mine, d_off_all, (size_t)B + 1, ncclInt32, s->comm, st));
    // region capacity = the largest candidate count of any rank (sizes are exchanged below; buffers use a common bound)
    int32_t cap_all = total;
    if (G > 1) {
        DMG_CUDA(h, cudaMemcpyAsync(d_nreq, &cap_all, 4, cudaMemcpyHostToDevice, st));
        DMG_NCCL(h, g_nccl.AllReduce(d_nreq, d_nreq, 1, ncclInt32, ncclMax, s->comm, st));
        DMG_CUDA(h, cudaMemcpyAsync(&cap_all, d_nreq, 4, cudaMemcpyDeviceToHost, st));
        DMG_CUDA(h, cudaStreamSynchronize(st));
    }
    const int64_t stride = std::max<int64_t>(cap_all, 1);
    Scratch &sw = h->s_out;                                          // second scratch block: candidate-sized buffers
    const size_t cbytes = Carver::need({(size_t)stride * 4, (size_t)stride * 8, (size_t)G * stride * 8, (size_t)G * stride * 8,
                                        (size_t)G * stride * 8, (size_t)G * stride * 8, (size_t)B * topk * 4, (size_t)B * topk * 8, (size_t)B * 4});
    DMG_TRY(ensure_dev(h, sw, cbytes));
    Carver cc(sw.d);
    int32_t *d_cand_item = cc.take<int32_t>((size_t)stride);
    double *d_cand_score = cc.take<double>((size_t)stride);
    int2 *d_req = cc.take<int2>((size_t)G * stride), *d_rreq = cc.take<int2>((size_t)G * stride);
    double *d_rsc = cc.take<double>((size_t)G * stride), *d_reply = cc.take<double>((size_t)G * stride);
    int32_t *d_items = cc.take<int32_t>((size_t)B * topk);
    double *d_sc = cc.take<double>((size_t)B * topk);
    int32_t *d_cnt = cc.take<int32_t>((size_t)B);
    dr_cand_fill_kernel<<<std::min(B, h->sm_count * 4), kThreads, 0, st>>>(B, beam, D, K, d_paths, d_counts, d.d_path_off, d.d_path_items, d_off_mine, d_cand_item);
    DMG_CUDA(h, cudaMemsetAsync(d_nreq, 0, (size_t)G * 4, st));
    if (total) dr_bucket_kernel<<<(total + 255) / 256, 256, 0, st>>>(total, d_cand_item, d.item_chunk, d_req, stride, d_nreq);
    h->launches += 2;
    if (G > 1) DMG_NCCL(h, g_nccl.AllGather(d_nreq, d_matrix, (size_t)G, ncclInt32, s->comm, st));
    else DMG_CUDA(h, cudaMemcpyAsync(d_matrix, d_nreq, 4, cudaMemcpyDeviceToDevice, st));
    DMG_CUDA(h, cudaMemcpyAsync(h_matrix, d_matrix, (size_t)G * G * 4, cudaMemcpyDeviceToHost, st));
    DMG_CUDA(h, cudaStreamSynchronize(st));
    if (G > 1) {
        DMG_NCCL(h, g_nccl.GroupStart());
        for (int p = 0; p < G; p++) {
            if (p == s->rank) continue;
            const int ns = h_matrix[s->rank * G + p], nr = h_matrix[p * G + s->rank];
            if (ns) DMG_NCCL(h, g_nccl.Send(d_req + (size_t)p * stride, (size_t)ns * 2, ncclInt32, p, s->comm, st));
            if (nr) DMG_NCCL(h, g_nccl.Recv(d_rreq + (size_t)p * stride, (size_t)nr * 2, ncclInt32, p, s->comm, st));
        }
        DMG_NCCL(h, g_nccl.GroupEnd());
    }
    for (int p = 0; p < G; p++) {
        const int nr = h_matrix[p * G + s->rank];
        if (!nr) continue;
        const int2 *rq = p == s->rank ? d_req + (size_t)p * stride : d_rreq + (size_t)p * stride;
        double *ro = p == s->rank ? d_reply + (size_t)p * stride : d_rsc + (size_t)p * stride;
        dr_score_cands_kernel<<<(nr + 127) / 128, 128, 0, st>>>(nr, rq, d_off_all + (size_t)p * (B + 1), B, E, d_uvec + (size_t)p * B * E, d.item_base,
                                                                 d.d_sm_w, d.d_sm_b, ro);
        h->launches += 1;
        if (p != s->rank) s->exchanged_rows += nr;
    }
    if (G > 1) {
        DMG_NCCL(h, g_nccl.GroupStart());
        for (int p = 0; p < G; p++) {
            if (p == s->rank) continue;
            const int ns = h_matrix[s->rank * G + p], nr = h_matrix[p * G + s->rank];
            if (nr) DMG_NCCL(h, g_nccl.Send(d_rsc + (size_t)p * stride, (size_t)nr, ncclFloat64, p, s->comm, st));
            if (ns) DMG_NCCL(h, g_nccl.Recv(d_reply + (size_t)p * stride, (size_t)ns, ncclFloat64, p, s->comm, st));
        }
        DMG_NCCL(h, g_nccl.GroupEnd());
    }
    for (int p = 0; p < G; p++) {
        const int ns = h_matrix[s->rank * G + p];
        if (!ns) continue;
        dr_scatter_kernel<<<(ns + 255) / 256, 256, 0, st>>>(ns, d_req + (size_t)p * stride, d_reply + (size_t)p * stride, d_cand_score);
        h->launches += 1;
    }
    // ---- topk per user by (score desc, candidate index asc) -----------------------------------------------------------------
    DrRerankParams p;
    memset(&p, 0, sizeof(p));
    p.num_item = d.num_item; p.K = K; p.D = D; p.T = T; p.E = E; p.B = B; p.beam = beam; p.topk = topk;
    p.rr_emb = d.d_rr_emb; p.rr_wT = d.d_rr_w; p.rr_b = d.d_rr_b; p.sm_w = d.d_sm_w; p.sm_b = d.d_sm_b;
    p.seq = d_seq_mine; p.paths = d_paths; p.path_counts = d_counts; p.path_off = d.d_path_off; p.path_items = d.d_path_items;
    p.pre_scores = d_cand_score; p.cand_off = d_off_mine;
    p.out_items = d_items; p.out_scores = d_sc; p.out_counts = d_cnt;
    const size_t smem = ((size_t)T * E + E + 2) * 8 + (size_t)kDrCap * 16 + ((size_t)2 * beam + 2) * 8 + 64 * 4 + 32;
    if (smem > h->smem_optin) return fail(h, DMG_ERR_UNSUPPORTED, "rerank needs %zu B of shared memory", smem);
    DMG_CUDA(h, cudaFuncSetAttribute(dr_rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dr_rerank_kernel<<<std::min(B, h->sm_count * 2), kThreads, smem, st>>>(p);
    h->launches += 1;
    DMG_CUDA(h, cudaGetLastError());
    DMG_CUDA(h, cudaMemcpyAsync(out_items, d_items, (size_t)B * topk * 4, cudaMemcpyDeviceToHost, st));
    DMG_CUDA(h, cudaMemcpyAsync(out_scores, d_sc, (size_t)B * topk * 8, cudaMemcpyDeviceToHost, st));
    DMG_CUDA(h, cudaMemcpyAsync(out_counts, d_cnt, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    DMG_CUDA(h, cudaStreamSynchronize(st));
    return DMG_OK;
}

// ---- synthetic Deep Retrieval model generated on the device (benchmarks / BASELINE config 5) ---------------------------------------
// Tensor.randn at model construction (LayerModel.scala:22-39, RerankModel.scala:20-36: EmbeddingShare / Embedding / Linear init,
// softmaxWeights randn(0, 0.05)) as counter-based values of the GLOBAL element index, so a rank of a sharded engine fills exactly
// the slice an unsharded engine would hold; and a synthetic item -> path assignment (J hashed paths per item) turned into
// MappingOp.pathItemMapping's shape -- ONE item per path (MappingOp.scala:23-28), here the largest item id -- as a CSR over the
// K^D path keys, built on the device: a 100 M item catalogue never exists on the host.
namespace {

inline uint64_t splitmix_host(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__global__ void dr_randn_kernel(double *__restrict__ dst, int64_t n, int64_t goff, uint64_t seed, double std)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t hsh = splitmix64(seed ^ splitmix64((uint64_t)(goff + i)));
        const uint32_t a = (uint32_t)hsh, b = (uint32_t)(hsh >> 32);
        const float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = ((float)(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
        dst[i] = (double)(sqrtf(-2.0f * __logf(u1)) * __cosf(6.28318530718f * u2)) * std;
    }
}
__global__ void dr_fill_i32_kernel(int32_t *__restrict__ dst, int64_t n, int32_t v)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = v;
}
// path key of (item, j): splitmix64(seed ^ splitmix64(item J + j)) mod K^D; the path keeps its largest item
__global__ void dr_syn_winner_kernel(int64_t num_item, int J, int64_t n_keys, uint64_t seed, int32_t *__restrict__ winner)
{
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < num_item * J; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t key = (int64_t)(splitmix64(seed ^ splitmix64((uint64_t)q)) % (uint64_t)n_keys);
        atomicMax(winner + key, (int32_t)(q / J));
    }
}
constexpr int kScanPer = 8, kScanBlock = 1024 * kScanPer;
__global__ void __launch_bounds__(1024) dr_syn_count_kernel(int64_t n_keys, const int32_t *__restrict__ winner, int64_t *__restrict__ blk)
{
    __shared__ int sW[32];
    const int64_t base = (int64_t)blockIdx.x * kScanBlock + (int64_t)threadIdx.x * kScanPer;
    int c = 0;
    for (int q = 0; q < kScanPer; q++) c += (base + q < n_keys && winner[base + q] >= 0) ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) sW[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 32; w++) t += sW[w];
        blk[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(1024) dr_syn_scan_blocks_kernel(int64_t nb, int64_t *__restrict__ blk, int64_t *__restrict__ total)
{
    __shared__ int64_t sT[1024];
    const int64_t per = (nb + 1023) / 1024, b0 = (int64_t)threadIdx.x * per, b1 = min(b0 + per, nb);
    int64_t t = 0;
    for (int64_t b = b0; b < b1; b++) t += blk[b];
    sT[threadIdx.x] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t run = 0;
        for (int i = 0; i < 1024; i++) { const int64_t v = sT[i]; sT[i] = run; run += v; }
        *total = run;
    }
    __syncthreads();
    int64_t run = sT[threadIdx.x];
    for (int64_t b = b0; b < b1; b++) { const int64_t v = blk[b]; blk[b] = run; run += v; }
}
__global__ void __launch_bounds__(1024) dr_syn_csr_kernel(int64_t n_keys, const int32_t *__restrict__ winner, const int64_t *__restrict__ blk,
                                                          const int64_t *__restrict__ total, int64_t *__restrict__ path_off, int32_t *__restrict__ path_items)
{
    __shared__ int sW[32];
    const int64_t base = (int64_t)blockIdx.x * kScanBlock + (int64_t)threadIdx.x * kScanPer;
    int c = 0;
    for (int q = 0; q < kScanPer; q++) c += (base + q < n_keys && winner[base + q] >= 0) ? 1 : 0;
    int incl = c;
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if ((int)(threadIdx.x & 31) >= o) incl += v; }
    if ((threadIdx.x & 31) == 31) sW[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x == 0) { int run = 0; for (int w = 0; w < 32; w++) { const int v = sW[w]; sW[w] = run; run += v; } }
    __syncthreads();
    int64_t off = blk[blockIdx.x] + sW[threadIdx.x >> 5] + (incl - c);
    for (int q = 0; q < kScanPer; q++)
        if (base + q < n_keys) {
            path_off[base + q] = off;
            const int32_t w = winner[base + q];
            if (w >= 0) path_items[off++] = w;
        }
    if (blockIdx.x == 0 && threadIdx.x == 0) path_off[n_keys] = *total;
}

}  // namespace

DMG_API int32_t dmg_dr_init_synthetic(dmg_handle_t h, int32_t num_item, int32_t K, int32_t D, int32_t T, int32_t E, int32_t J, uint64_t seed)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    DMG_TRY(model_is_shared(h, "dmg_dr_init_synthetic"));
    if (num_item <= 0 || K <= 0 || T <= 0 || E <= 0 || J <= 0) return fail(h, DMG_ERR_INVALID_ARG, "dmg_dr_init_synthetic: bad arguments");
    if (D < 2 || D > kDrMaxD) return fail(h, DMG_ERR_INVALID_ARG, "number of layers must be in [2, %d]", kDrMaxD);
    double nk = 1;
    for (int i = 0; i < D; i++) nk *= K;
    if (nk > 2.0e9) return fail(h, DMG_ERR_UNSUPPORTED, "K^D = %.3g path keys: dense CSR too large", nk);
    const int64_t n_keys = (int64_t)nk;
    DMG_CUDA(h, cudaSetDevice(h->device));
    dmg_free_dr(h->dr);
    DrDev &d = h->dr;
    ShardState *s = h->shard;
    const int world = s ? s->world : 1, rank = s ? s->rank : 0;
    d.num_item = num_item; d.K = K; d.D = D; d.T = T; d.E = E;
    d.sharded = s != nullptr;                                     // after dmg_shard_init: this rank's item range, served by dmg_shard_dr_retrieve
    d.item_chunk = ((int64_t)num_item + world - 1) / world;
    d.item_base = std::min<int64_t>((int64_t)rank * d.item_chunk, num_item);
    d.local_items = std::min<int64_t>(d.item_chunk, num_item - d.item_base);
    if (!d.sharded) { d.item_chunk = 0; d.item_base = 0; d.local_items = num_item; }
    const int64_t node_rows = (int64_t)K * (D - 1), li = std::max<int64_t>(d.local_items, 1);
    const int grid = h->sm_count * 16;
    const double std_ = 0.05;
    auto fill = [&](double *dst, int64_t n, int64_t goff, uint64_t sd, double sdv) {
        if (n > 0) dr_randn_kernel<<<grid, 256, 0, h->stream>>>(dst, n, goff, seed ^ splitmix_host(sd), sdv);
        h->launches += 1;
    };
    // local layer table = [owned item rows | path-node rows] (whole table when not sharded: the node rows follow the items anyway)
    DMG_CUDA(h, cudaMalloc(&d.d_layer_emb, (size_t)(d.local_items + node_rows) * E * 8));
    fill(d.d_layer_emb, d.local_items * E, d.item_base * E, 1, std_);
    fill(d.d_layer_emb + d.local_items * E, node_rows * E, (int64_t)num_item * E, 1, std_);
    for (int i = 0; i < D; i++) {
        const int in = (T + i) * E;
        double *w = nullptr, *b = nullptr, *wT = nullptr;
        DMG_CUDA(h, cudaMalloc(&w, (size_t)K * in * 8));
        DMG_CUDA(h, cudaMalloc(&b, (size_t)K * 8));
        DMG_CUDA(h, cudaMalloc(&wT, (size_t)K * in * 8));
        fill(w, (int64_t)K * in, 0, 10 + i, std_);
        DMG_CUDA(h, cudaMemsetAsync(b, 0, (size_t)K * 8, h->stream));
        transpose_kernel<double><<<(K * in + 255) / 256, 256, 0, h->stream>>>(w, wT, K, in);
        h->launches += 1;
        d.d_layer_w.push_back(w); d.d_layer_b.push_back(b); d.d_layer_wT.push_back(wT);
    }
    DMG_CUDA(h, cudaMalloc(&d.d_rr_emb, (size_t)li * E * 8));
    fill(d.d_rr_emb, d.local_items * E, d.item_base * E, 2, std_);
    DMG_CUDA(h, cudaMalloc(&d.d_rr_w, (size_t)E * T * E * 8));                              // kept [T E][E]: the values are defined in this layout
    fill(d.d_rr_w, (int64_t)E * T * E, 0, 3, std_);
    DMG_CUDA(h, cudaMalloc(&d.d_rr_b, (size_t)E * 8));
    DMG_CUDA(h, cudaMemsetAsync(d.d_rr_b, 0, (size_t)E * 8, h->stream));
    DMG_CUDA(h, cudaMalloc(&d.d_sm_w, (size_t)li * E * 8));
    fill(d.d_sm_w, d.local_items * E, d.item_base * E, 4, std_);
    DMG_CUDA(h, cudaMalloc(&d.d_sm_b, (size_t)li * 8));
    fill(d.d_sm_b, d.local_items, d.item_base, 5, 0.01);
    // path CSR (replicated)
    int32_t *winner = nullptr;
    int64_t *blk = nullptr;
    const int64_t nb = (n_keys + kScanBlock - 1) / kScanBlock;
    DMG_CUDA(h, cudaMalloc(&winner, (size_t)n_keys * 4));
    DMG_CUDA(h, cudaMalloc(&blk, (size_t)(nb + 1) * 8));
    dr_fill_i32_kernel<<<grid, 256, 0, h->stream>>>(winner, n_keys, -1);
    dr_syn_winner_kernel<<<grid, 256, 0, h->stream>>>(num_item, J, n_keys, seed ^ splitmix_host(6), winner);
    dr_syn_count_kernel<<<(unsigned)nb, 1024, 0, h->stream>>>(n_keys, winner, blk);
    dr_syn_scan_blocks_kernel<<<1, 1024, 0, h->stream>>>(nb, blk, blk + nb);
    int64_t total = 0;
    DMG_CUDA(h, cudaMemcpyAsync(&total, blk + nb, 8, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    DMG_CUDA(h, cudaMalloc(&d.d_path_off, (size_t)(n_keys + 1) * 8));
    DMG_CUDA(h, cudaMalloc(&d.d_path_items, (size_t)std::max<int64_t>(total, 1) * 4));
    dr_syn_csr_kernel<<<(unsigned)nb, 1024, 0, h->stream>>>(n_keys, winner, blk, blk + nb, d.d_path_off, d.d_path_items);
    h->launches += 5;
    DMG_CUDA(h, cudaGetLastError());
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(winner); cudaFree(blk);
    d.loaded = true;
    d.paths_loaded = true;
    return DMG_OK;
}
