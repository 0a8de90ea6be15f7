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
    double *sXn = sU + K;                                          // kDrPC x nodeE
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
            sX0[i] = c < 0 ? 0.0 : p.layer_emb[(size_t)c * E + i % E];
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
                for (int k = 0; k < T * E; k++) acc = fma_(__ldg(wT + (size_t)k * K + o), sX0[k], acc);
                sU[o] = acc;
            }
            __syncthreads();
            // (2) every live path continues the chain over its node inputs
            const int nk = layer * E;
            for (int pb = 0; pb < live; pb += kDrPC) {
                const int np = live - pb < kDrPC ? live - pb : kDrPC;
                for (int i = tid; i < np * nk; i += kThreads) {
                    const int pp = i / nk, k = i % nk, j = k / E;
                    const int32_t row = path[(pb + pp) * D + j] + p.num_item + j * K;   // CandidateSearcher.scala:54
                    sXn[pp * nodeE + k] = p.layer_emb[(size_t)row * E + k % E];
                }
                __syncthreads();
                for (int o = tid; o < K; o += kThreads) {
                    double acc[kDrPC];
                    const double u = sU[o];
#pragma unroll
                    for (int pp = 0; pp < kDrPC; pp++) acc[pp] = u;
                    for (int k = 0; k < nk; k++) {
                        const double w = __ldg(wT + (size_t)(T * E + k) * K + o);
#pragma unroll
                        for (int pp = 0; pp < kDrPC; pp++) acc[pp] = fma_(w, sXn[pp * nodeE + k], acc[pp]);
                    }
                    const double bo = __ldg(bias + o);
#pragma unroll
                    for (int pp = 0; pp < kDrPC; pp++)
                        if (pp < np) Lmat[(size_t)(pb + pp) * K + o] = add_(acc[pp], bo);
                }
                __syncthreads();
            }
            // (3) softmax per path: max, exp(x - max), left-to-right sum (dr/package.scala:23-28)
            for (int pp = tid; pp < live; pp += kThreads) {
                double *row = Lmat + (size_t)pp * K;
                double mx = row[0];
                for (int c = 1; c < K; c++) { const double v = row[c]; mx = v > mx ? v : mx; }
                double sum = 0.0;
                for (int c = 0; c < K; c++) { const double e = exp_(sub_(row[c], mx)); row[c] = e; sum = add_(sum, e); }
                sSum[pp] = sum;
            }
            if (tid == 0) { *sCount = 0; *sTau = KO::lowest(); }
            __syncthreads();
            // (4) candidates (path, node): probability = parent * (e / sum); keep the best `beam`
            const int total = live * K;
            for (int base = 0; base < total; base += kDrChunk) {
                const Key128 tau = *sTau;
                for (int idx = base + tid; idx < total && idx < base + kDrChunk; idx += kThreads) {
                    const int pp = idx / K;
                    const double cp = mul_(prob[pp], __ddiv_rn(Lmat[idx], sSum[pp]));
                    const Key128 key = KO::make(cp, idx);
                    if (key_better(key, tau)) sBuf[atomicAdd(sCount, 1)] = key;
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
                nprob[q] = mul_(prob[pp], __ddiv_rn(Lmat[pos], sSum[pp]));
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
        for (int i = tid; i < T * E; i += kThreads) {
            const int32_t c = p.seq[(size_t)user * T + i / E];
            sX[i] = c < 0 ? 0.0 : p.rr_emb[(size_t)c * E + i % E];
        }
        __syncthreads();
        // user vector u = W_r x + b_r (RerankModel.inferenceUserVector :54-68)
        for (int o = tid; o < E; o += kThreads) {
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
                const int32_t item = p.path_items[sBase[lo] + (idx - sStart[lo])];
                const double *w = p.sm_w + (size_t)item * E;
                double acc = 0.0;
                for (int k = 0; k < E; k++) acc = fma_(__ldg(w + k), sUv[k], acc);
                const double sc = add_(acc, __ldg(p.sm_b + item));
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
                const double *w = p.sm_w + (size_t)item * E;
                double acc = 0.0;
                for (int k = 0; k < E; k++) acc = fma_(__ldg(w + k), sUv[k], acc);
                sc = add_(acc, __ldg(p.sm_b + item));
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
    d = DrDev();
}

DMG_API int32_t dmg_dr_load(dmg_handle_t h, int32_t num_item, int32_t K, int32_t D, int32_t T, int32_t E,
                            const double *layer_emb, const double *const *layer_w, const double *const *layer_b,
                            const double *rr_emb, const double *rr_w, const double *rr_b, const double *sm_w,
                            const double *sm_b)
{
    if (!h) return DMG_ERR_INVALID_ARG;
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
                               int32_t **d_paths, double **d_probs, int32_t **d_counts, int32_t **d_seq_out)
{
    DrDev &d = h->dr;
    if (!d.loaded) return fail(h, DMG_ERR_STATE, "dmg_dr_load first");
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
    const int grid = std::min(B, h->sm_count);
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
    for (int i = 0; i < D; i++) { p.wT[i] = d.d_layer_wT[i]; p.b[i] = d.d_layer_b[i]; }
    p.seq = d_seq; p.scratch = d_scr; p.out_paths = *d_paths; p.out_probs = *d_probs; p.out_counts = *d_counts;
    const int nodeE = std::max((D - 1) * E, 1);
    size_t smem = ((size_t)T * E + K + (size_t)kDrPC * nodeE + 3 * (size_t)beam + 2) * 8 + (size_t)kDrCap * 16 +
                  (size_t)2 * beam * D * 4 + 64 * 4 + 32;
    if (smem > h->smem_optin)
        return fail(h, DMG_ERR_UNSUPPORTED, "Deep Retrieval shape needs %zu B of shared memory (limit %zu)", smem, h->smem_optin);
    DMG_CUDA(h, cudaFuncSetAttribute(dr_beam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
