// dr_train.cu -- Deep Retrieval training step: one mini-batch iteration of
//   deep-retrieval/src/main/scala/com/mass/dr/optim/LocalOptimizer.scala:58-194
//     layer model   MiniBatch.transformLayerData (dataset/MiniBatch.scala:19-50) -> EmbeddingShare -> Reshape -> Linear per layer
//                   (model/LayerModel.scala:22-39) -> CrossEntropyLayer (loss/CrossEntropyLayer.scala:13-24: LogSoftMax.scala:36-67 +
//                   ClassNLLCriterion.scala:17-104, sizeAverage) -> backward -> syncGradients over the thread chunks (:139-187)
//                   -> Adam (scalann/.../optim/Adam.scala:19-73)
//     rerank model  transformRerankData (:52-61) -> Embedding -> Reshape -> Linear (model/RerankModel.scala:20-36) ->
//                   SampledSoftmaxLoss (scalann/.../nn/SampledSoftmaxLoss.scala:49-153, batchMode = false) with its own Adam over the
//                   softmax weights / biases (nn/mixin/ParameterOptimizer.scala:28-88) -> model backward -> Adam
// Everything is Double.  The three GEMM shapes of a Linear (forward, gradWeight, gradInput) run on one tiled kernel whose every
// output element is ONE fma chain over ascending k from 0 -- the arithmetic spec of the oracle (oracle/oracle_dr_train.c), so
// logits and Linear weight gradients carry the oracle's bits; `log` (CUDA vs glibc), the chunked bias / loss sums and the atomic scatter-adds into the embedding
// / softmax-parameter gradients differ in the last bits, so training parity is tolerance-based (1e-12 relative, tests/test_gpu_dr_train.py).
#include <algorithm>
#include <cmath>
#include <vector>

#include "device_utils.cuh"
#include "rows_kernels.cuh"

using namespace dmg;

namespace {

constexpr int kGK = 8;                   // k step; C tile TILE x TILE per 256-thread CTA (TILE / 16 squared per thread), two shared-memory stages

// C(i, j) <- epilogue(sum_k A(i, k) B(k, j)),  A(i, k) = A[i sa_i + k sa_k],  B(k, j) = B[k sb_k + j sb_j]
// mode 0: C = acc + bias[j] (Linear.updateOutput: addmm then add bias)   1: C = acc   2: C = C + acc
// Thread (ty, tx) of the 16 x 16 grid owns rows {32 m + 2 ty, + 1 : m < 4} and the same pattern of columns: its a / b fragments are
// four 16-byte shared-memory loads each per k, the rows of a warp broadcast and the 16 lanes of a column load read 256 contiguous
// bytes.  Every accumulator is ONE fma chain over ascending k (the oracle's arithmetic).
template <int TILE>
__global__ void __launch_bounds__(256) dr_gemm_kernel(int M, int N, int Kd, const double *__restrict__ A, int64_t sa_i, int64_t sa_k,
                                                       const double *__restrict__ B, int64_t sb_k, int64_t sb_j, double *__restrict__ C,
                                                       int64_t ldc, const double *__restrict__ bias, int mode)
{
    constexpr int R = TILE / 16, NQ = TILE * kGK / 256, LG = TILE == 128 ? 7 : 6;      // outputs per thread and dimension, loads per thread and tile
    __shared__ __align__(16) double sA[2][kGK][TILE], sB[2][kGK][TILE];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int i0 = blockIdx.y * TILE, j0 = blockIdx.x * TILE;
    double acc[R][R];
#pragma unroll
    for (int a = 0; a < R; a++)
#pragma unroll
        for (int b = 0; b < R; b++) acc[a][b] = 0.0;
    double ra[NQ], rb[NQ];
    auto gload = [&](int k0) {                                      // TILE x 8 elements of each tile, consecutive threads along the unit stride
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            const int e = tid + 256 * q;
            int ai, ak, bk, bj;
            if (sa_k == 1) { ak = e & 7; ai = e >> 3; } else { ai = e & (TILE - 1); ak = e >> LG; }
            if (sb_k == 1) { bk = e & 7; bj = e >> 3; } else { bj = e & (TILE - 1); bk = e >> LG; }
            ra[q] = (i0 + ai < M && k0 + ak < Kd) ? __ldg(A + (int64_t)(i0 + ai) * sa_i + (int64_t)(k0 + ak) * sa_k) : 0.0;
            rb[q] = (j0 + bj < N && k0 + bk < Kd) ? __ldg(B + (int64_t)(k0 + bk) * sb_k + (int64_t)(j0 + bj) * sb_j) : 0.0;
        }
    };
    auto sstore = [&](int st) {
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            const int e = tid + 256 * q;
            int ai, ak, bk, bj;
            if (sa_k == 1) { ak = e & 7; ai = e >> 3; } else { ai = e & (TILE - 1); ak = e >> LG; }
            if (sb_k == 1) { bk = e & 7; bj = e >> 3; } else { bj = e & (TILE - 1); bk = e >> LG; }
            sA[st][ak][ai] = ra[q];
            sB[st][bk][bj] = rb[q];
        }
    };
    gload(0);
    sstore(0);
    __syncthreads();
    int st = 0;
    for (int k0 = 0; k0 < Kd; k0 += kGK, st ^= 1) {
        const bool more = k0 + kGK < Kd;
        if (more) gload(k0 + kGK);                                  // next tile in flight while this one is multiplied
        const int kn = Kd - k0 < kGK ? Kd - k0 : kGK;
#pragma unroll
        for (int k = 0; k < kGK; k++) {
            if (k < kn) {
                double a[R], b[R];
#pragma unroll
                for (int m = 0; m < R / 2; m++) {
                    const double2 av = *reinterpret_cast<const double2 *>(&sA[st][k][32 * m + ty * 2]), bv = *reinterpret_cast<const double2 *>(&sB[st][k][32 * m + tx * 2]);
                    a[2 * m] = av.x; a[2 * m + 1] = av.y; b[2 * m] = bv.x; b[2 * m + 1] = bv.y;
                }
#pragma unroll
                for (int x = 0; x < R; x++)
#pragma unroll
                    for (int y = 0; y < R; y++) acc[x][y] = __fma_rn(a[x], b[y], acc[x][y]);
            }
        }
        if (more) sstore(st ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int x = 0; x < R; x++)
#pragma unroll
        for (int y = 0; y < R; y++) {
            const int i = i0 + 32 * (x >> 1) + ty * 2 + (x & 1), j = j0 + 32 * (y >> 1) + tx * 2 + (y & 1);
            if (i < M && j < N) {
                double *c = C + (int64_t)i * ldc + j;
                *c = mode == 0 ? __dadd_rn(acc[x][y], bias[j]) : (mode == 1 ? acc[x][y] : __dadd_rn(*c, acc[x][y]));
            }
        }
}

// out[j] (+)= sum_r M[r][j] (a chain over ascending r): gradBias of a Linear
// blockIdx.y = one of kColChunks row ranges (a sequential chain each, written to part[chunk][j]); dr_colsum_finish_kernel adds the
// chunks in order.  Deterministic; the order differs from the oracle's single chain over all rows in the last bits.
constexpr int kColChunks = 32;
__global__ void dr_colsum_kernel(int64_t R, int N, const double *__restrict__ m, double *__restrict__ part)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const int64_t per = (R + kColChunks - 1) / kColChunks, r0 = blockIdx.y * per, r1 = r0 + per < R ? r0 + per : R;
    double acc = 0.0;
    for (int64_t r = r0; r < r1; r++) acc = __dadd_rn(acc, m[r * N + j]);
    part[(size_t)blockIdx.y * N + j] = acc;
}
__global__ void dr_colsum_finish_kernel(int N, const double *__restrict__ part, double *__restrict__ out, int accumulate)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    double acc = 0.0;
    for (int c = 0; c < kColChunks; c++) acc = __dadd_rn(acc, part[(size_t)c * N + j]);
    out[j] = accumulate ? __dadd_rn(out[j], acc) : acc;
}

// rows of MiniBatch.transformLayerData: row (sample s, path p): idx = seq ++ (path[i] + numItem + i K), tgt[d] = path[d];
// X = the embedding rows (padding -> zeros).  One CTA per row.  P == 0: rerank rows (idx = seq only, no paths).
__global__ void dr_rows_kernel(int T, int W, int E, int D, int P, int num_item, int K, const int32_t *__restrict__ seq,
                               const int32_t *__restrict__ target, const int32_t *__restrict__ item_paths, const double *__restrict__ emb,
                               int32_t *__restrict__ idx, int32_t *__restrict__ tgt, double *__restrict__ X)
{
    const int64_t r = blockIdx.x;
    const int64_t s = P ? r / P : r;
    const int p = P ? (int)(r % P) : 0;
    const int32_t *path = P ? item_paths + ((int64_t)target[s] * P + p) * D : nullptr;
    for (int j = threadIdx.x; j < W; j += blockDim.x) idx[r * W + j] = j < T ? seq[s * T + j] : path[j - T] + num_item + (j - T) * K;
    if (P)
        for (int d = threadIdx.x; d < D; d += blockDim.x) tgt[(int64_t)d * gridDim.x + r] = path[d];      // [D][R]: one column per layer
    for (int i = threadIdx.x; i < W * E; i += blockDim.x) {
        const int j = i / E;
        const int32_t c = j < T ? seq[s * T + j] : path[j - T] + num_item + (j - T) * K;
        X[r * (int64_t)W * E + i] = c >= 0 ? __ldg(emb + (int64_t)c * E + (i - j * E)) : 0.0;
    }
}

// EmbeddingShare / Embedding backward: g_emb[idx] += GX rows (padding skipped); layer d only reaches positions < T + d, which is
// how GX was accumulated.
__global__ void dr_scatter_kernel(int W, int E, const int32_t *__restrict__ idx, const double *__restrict__ GX, double *__restrict__ g_emb)
{
    const int64_t r = blockIdx.x;
    for (int i = threadIdx.x; i < W * E; i += blockDim.x) {
        const int j = i / E;
        const int32_t c = idx[r * W + j];
        if (c >= 0) atomicAdd(g_emb + (int64_t)c * E + (i - j * E), GX[r * (int64_t)W * E + i]);
    }
}

// CrossEntropyCriterion on one row per CTA: LogSoftMax.updateOutputOne (max, exp(in - max), dot with ones as ONE chain, logSum =
// max + log(sum)), ClassNLLCriterion (loss_r = -(in[t] - logSum)), then gradInput = gradOut + (1 / R) exp(out) in place of the logits.
// `R` is the row count of the row's thread chunk (sizeAverage inside each thread's criterion).
__global__ void __launch_bounds__(256) dr_ce_kernel(int C, double *__restrict__ logits, const int32_t *__restrict__ target, double inv_R,
                                                     double *__restrict__ row_loss)
{
    extern __shared__ double sBuf[];
    __shared__ double sRed[8];
    __shared__ double sLogSum;
    const int64_t r = blockIdx.x;
    double *in = logits + r * C;
    const int tid = threadIdx.x;
    double mx = -INFINITY;
    for (int j = tid; j < C; j += blockDim.x) mx = fmax(mx, in[j]);
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) sRed[tid >> 5] = mx;
    __syncthreads();
    mx = sRed[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) mx = fmax(mx, sRed[w]);
    for (int j = tid; j < C; j += blockDim.x) sBuf[j] = exp_(__dadd_rn(-mx, in[j]));
    __syncthreads();
    if (tid == 0) {
        double sum = 0.0;
        for (int j = 0; j < C; j++) sum = __fma_rn(sBuf[j], 1.0, sum);
        const double ls = __dadd_rn(mx, log(sum));
        sLogSum = ls;
        row_loss[r] = __dadd_rn(in[target[r]], -ls);                  // out[target]; the caller subtracts it from the running loss
    }
    __syncthreads();
    const double ls = sLogSum, go_t = -inv_R;
    const int t = target[r];
    for (int j = tid; j < C; j += blockDim.x) in[j] = __fma_rn(inv_R, exp_(__dadd_rn(in[j], -ls)), j == t ? go_t : 0.0);
}

// output = 0; output -= out[target] row after row; output /= R   (ClassNLLCriterion.updateOutput :44-64)
__global__ void dr_loss_kernel(int64_t R, const double *__restrict__ row_out, double *__restrict__ loss, int accumulate)
{
    if (blockIdx.x || threadIdx.x >= 32) return;                    // one warp: 32 interleaved chains, then a fixed shuffle tree
    double o = 0.0;
    for (int64_t r = threadIdx.x; r < R; r += 32) o = __dadd_rn(o, -row_out[r]);
    for (int s_ = 16; s_ > 0; s_ >>= 1) o = __dadd_rn(o, __shfl_xor_sync(0xffffffffu, o, s_));
    if (threadIdx.x) return;
    o = __ddiv_rn(o, (double)R);
    *loss = accumulate ? __dadd_rn(*loss, o) : o;
}

__global__ void dr_scale_kernel(double *__restrict__ x, int64_t n, double div)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = __ddiv_rn(x[i], div);
}

// SampledSoftmaxLoss on one sample per CTA: logits over the S + 1 sampled items (addmv + bias), CrossEntropy against position 0,
// gradInput = w^T . logitGrad from the weights BEFORE their update, parameter gradients accumulated (ParameterOptimizer :67-88).
__global__ void __launch_bounds__(128) dr_sampled_softmax_kernel(int E, int C, const double *__restrict__ u, const double *__restrict__ sm_w,
                                                                  const double *__restrict__ sm_b, const int32_t *__restrict__ sampled,
                                                                  double inv_n, double *__restrict__ gu, double *__restrict__ g_sm_w,
                                                                  double *__restrict__ g_sm_b, double *__restrict__ row_out)
{
    extern __shared__ double sm[];
    double *sU = sm, *sLg = sU + E, *sEx = sLg + C;
    __shared__ double sLogSum;
    const int i = blockIdx.x, tid = threadIdx.x;
    const int32_t *it = sampled + (int64_t)i * C;
    for (int e = tid; e < E; e += blockDim.x) sU[e] = u[(int64_t)i * E + e];
    __syncthreads();
    for (int j = tid; j < C; j += blockDim.x) {
        const double *w = sm_w + (int64_t)it[j] * E;
        double acc = 0.0;
        for (int e = 0; e < E; e++) acc = __fma_rn(w[e], sU[e], acc);
        sLg[j] = __dadd_rn(acc, sm_b[it[j]]);
    }
    __syncthreads();
    if (tid == 0) {
        double mx = sLg[0];
        for (int j = 1; j < C; j++) mx = sLg[j] > mx ? sLg[j] : mx;
        double sum = 0.0;
        for (int j = 0; j < C; j++) sum = __fma_rn(exp_(__dadd_rn(-mx, sLg[j])), 1.0, sum);
        sLogSum = __dadd_rn(mx, log(sum));
        row_out[i] = __dadd_rn(sLg[0], -sLogSum);
    }
    __syncthreads();
    const double ls = sLogSum;
    for (int j = tid; j < C; j += blockDim.x) sEx[j] = __fma_rn(inv_n, exp_(__dadd_rn(sLg[j], -ls)), j == 0 ? -inv_n : 0.0);   // logitGrad
    __syncthreads();
    for (int e = tid; e < E; e += blockDim.x) {
        double acc = 0.0;
        for (int j = 0; j < C; j++) acc = __fma_rn(sm_w[(int64_t)it[j] * E + e], sEx[j], acc);
        gu[(int64_t)i * E + e] = acc;
    }
    for (int q = tid; q < C * E; q += blockDim.x) {
        const int j = q / E, e = q - j * E;
        atomicAdd(g_sm_w + (int64_t)it[j] * E + e, __dmul_rn(sEx[j], sU[e]));
    }
    for (int j = tid; j < C; j += blockDim.x) atomicAdd(g_sm_b + it[j], sEx[j]);
}

// SampledSoftmaxLoss.uniformSampler (:156-178): the positive first, then numSampled distinct uniform negatives != positive in
// ascending order (BitSet.foreach).  The reference draws from ThreadLocalRandom; here a counter-based generator (splitmix64).
__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// One warp per sample, the sorted list in shared memory: membership and insert position by a parallel scan, the tail shifted 32 at a time.
__global__ void __launch_bounds__(128) dr_sample_kernel(int n, int S, int num_item, const int32_t *__restrict__ target, uint64_t seed,
                                                        int32_t *__restrict__ out)
{
    extern __shared__ int32_t sList[];                              // [4][S]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * 4 + warp;
    if (i >= n) return;
    int32_t *row = sList + warp * S;
    const int32_t pos = target[i];
    int have = 0;
    uint64_t ctr = 0;
    while (have < S) {
        const int32_t s = (int32_t)(splitmix64(seed ^ splitmix64(((uint64_t)i << 32) | ctr++)) % (uint64_t)num_item);   // the same draw on every lane
        if (s == pos) continue;
        int less = 0, dup = 0;
        for (int q = lane; q < have; q += 32) { const int32_t v = row[q]; less += v < s ? 1 : 0; dup |= v == s ? 1 : 0; }
        less = __reduce_add_sync(0xffffffffu, less);
        if (__any_sync(0xffffffffu, dup)) continue;
        for (int hi = have; hi > less; hi -= 32) {                  // shift row[less, have) one slot to the right, from the end
            const int q = hi - 1 - lane;
            const int32_t v = q >= less ? row[q] : 0;
            __syncwarp();
            if (q >= less) row[q + 1] = v;
            __syncwarp();
        }
        if (lane == 0) row[less] = s;
        __syncwarp();
        have++;
    }
    int32_t *dst = out + (int64_t)i * (S + 1);
    if (lane == 0) dst[0] = pos;
    for (int q = lane; q < S; q += 32) dst[1 + q] = row[q];
}

// Adam.optimize / ParameterOptimizer.optimize: the same tensor operations (s = s b1 + (1 - b1) g; r = r b2 + (1 - b2) g g;
// denom = sqrt(r) + eps; w += -step s / denom).  ZERO: fuse the next iteration's zeroGradParameters.
template <bool ZERO>
__global__ void __launch_bounds__(256) dr_adam_kernel(double *__restrict__ w, double *__restrict__ g, double *__restrict__ s, double *__restrict__ r,
                                                       int64_t n, double b1, double omb1, double b2, double omb2, double eps, double nstep)
{
    auto one = [&](double &wi, double &gi, double &si_, double &ri_) {
        const double si = __dadd_rn(__dmul_rn(si_, b1), __dmul_rn(omb1, gi));
        const double ri = __dadd_rn(__dmul_rn(ri_, b2), __dmul_rn(omb2, __dmul_rn(gi, gi)));
        const double denom = __dadd_rn(__dsqrt_rn(ri), eps);
        wi = __dadd_rn(wi, __dmul_rn(nstep, __ddiv_rn(si, denom)));
        si_ = si; ri_ = ri;
        if (ZERO) gi = 0.0;
    };
    const int64_t nv = n / 2, stride = (int64_t)gridDim.x * blockDim.x;      // 16-byte accesses (every tensor is cudaMalloc-aligned)
    double2 *wv = reinterpret_cast<double2 *>(w), *gv = reinterpret_cast<double2 *>(g), *sv = reinterpret_cast<double2 *>(s), *rv = reinterpret_cast<double2 *>(r);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        double2 a = wv[i], b = gv[i], c = sv[i], d = rv[i];
        one(a.x, b.x, c.x, d.x);
        one(a.y, b.y, c.y, d.y);
        wv[i] = a; sv[i] = c; rv[i] = d;
        if (ZERO) gv[i] = b;
    }
    for (int64_t i = nv * 2 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) one(w[i], g[i], s[i], r[i]);
}

struct Tensors {                        // DrDev's tensors in training order with their element counts
    std::vector<double *> w;
    std::vector<int64_t> n;
};
Tensors tensors_of(DrDev &d)
{
    Tensors t;
    const int64_t E = d.E;
    t.w.push_back(d.d_layer_emb); t.n.push_back(((int64_t)d.num_item + (int64_t)d.K * (d.D - 1)) * E);
    for (int i = 0; i < d.D; i++) {
        t.w.push_back(d.d_layer_w[i]); t.n.push_back((int64_t)d.K * (d.T + i) * E);
        t.w.push_back(d.d_layer_b[i]); t.n.push_back(d.K);
    }
    t.w.push_back(d.d_rr_emb); t.n.push_back((int64_t)d.num_item * E);
    t.w.push_back(d.d_rr_w); t.n.push_back((int64_t)E * d.T * E);
    t.w.push_back(d.d_rr_b); t.n.push_back(E);
    t.w.push_back(d.d_sm_w); t.n.push_back((int64_t)d.num_item * E);
    t.w.push_back(d.d_sm_b); t.n.push_back(d.num_item);
    return t;
}

int32_t ensure_train_state(dmg_handle_t h)
{
    DrDev &d = h->dr;
    if (!d.tr_g.empty()) return DMG_OK;
    const Tensors t = tensors_of(d);
    for (size_t i = 0; i < t.w.size(); i++) {
        double *g = nullptr, *s = nullptr, *r = nullptr;
        const size_t bytes = (size_t)t.n[i] * sizeof(double);
        DMG_CUDA(h, cudaMalloc(&g, bytes)); d.tr_g.push_back(g);
        DMG_CUDA(h, cudaMalloc(&s, bytes)); d.tr_s.push_back(s);
        DMG_CUDA(h, cudaMalloc(&r, bytes)); d.tr_r.push_back(r);
        DMG_CUDA(h, cudaMemsetAsync(g, 0, bytes, h->stream));
        DMG_CUDA(h, cudaMemsetAsync(s, 0, bytes, h->stream));
        DMG_CUDA(h, cudaMemsetAsync(r, 0, bytes, h->stream));
    }
    return DMG_OK;
}

void gemm(dmg_handle_t h, int M, int N, int Kd, const double *A, int64_t sa_i, int64_t sa_k, const double *B, int64_t sb_k, int64_t sb_j,
          double *C, int64_t ldc, const double *bias, int mode)
{
    if (M <= 0 || N <= 0) return;
    // 128 x 128 tiles (8 x 8 per thread: 1 shared-memory load per 8 fma) when they fill the machine twice over, else 64 x 64
    const int64_t big = (int64_t)((N + 127) / 128) * ((M + 127) / 128);
    if (big >= 2 * (int64_t)h->sm_count) {
        dim3 grid((unsigned)((N + 127) / 128), (unsigned)((M + 127) / 128));
        dr_gemm_kernel<128><<<grid, 256, 0, h->stream>>>(M, N, Kd, A, sa_i, sa_k, B, sb_k, sb_j, C, ldc, bias, mode);
    } else {
        dim3 grid((unsigned)((N + 63) / 64), (unsigned)((M + 63) / 64));
        dr_gemm_kernel<64><<<grid, 256, 0, h->stream>>>(M, N, Kd, A, sa_i, sa_k, B, sb_k, sb_j, C, ldc, bias, mode);
    }
    h->launches += 1;
}

template <bool ZERO> void adam(dmg_handle_t h, double *w, double *g, double *s, double *r, int64_t n, double lr, double eps, int t)
{
    const double beta1 = 0.9, beta2 = 0.999;
    const double step = lr * std::sqrt(1 - std::pow(beta2, t)) / (1 - std::pow(beta1, t));
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 16);
    dr_adam_kernel<ZERO><<<grid, 256, 0, h->stream>>>(w, g, s, r, n, beta1, 1 - beta1, beta2, 1 - beta2, eps, -step);
    h->launches += 1;
}

}  // namespace

/* itemPathMapping (LocalDataSet: item index -> its numPathPerItem paths), kept on the device for the training steps */
DMG_API int32_t dmg_dr_load_item_paths(dmg_handle_t h, int32_t P, const int32_t *item_paths)
{
    if (!h || !item_paths || P <= 0) return DMG_ERR_INVALID_ARG;
    DMG_TRY(model_is_shared(h, "dmg_dr_load_item_paths"));
    DrDev &d = h->dr;
    if (!d.loaded || d.sharded) return fail(h, DMG_ERR_STATE, "dmg_dr_load first (whole tables)");
    const int64_t n = (int64_t)d.num_item * P * d.D;
    for (int64_t i = 0; i < n; i++)
        if (item_paths[i] < 0 || item_paths[i] >= d.K) return fail(h, DMG_ERR_INDEX, "path node %d outside [0, %d)", item_paths[i], d.K);
    DMG_CUDA(h, cudaSetDevice(h->device));
    cudaFree(d.d_item_paths);
    d.d_item_paths = nullptr;
    DMG_CUDA(h, cudaMalloc(&d.d_item_paths, (size_t)n * 4));
    DMG_CUDA(h, cudaMemcpyAsync(d.d_item_paths, item_paths, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    d.P = P;
    return DMG_OK;
}

DMG_API int32_t dmg_dr_train_step(dmg_handle_t h, int32_t n, const int32_t *seq, const int32_t *target, const int32_t *sampled,
                                  int32_t num_sampled, uint64_t seed, double lr, int32_t step_t, int32_t rerank_step_t,
                                  int32_t parallelism, int32_t apply, double *out_layer_loss, double *out_rerank_loss)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    DMG_TRY(model_is_shared(h, "dmg_dr_train_step"));
    DrDev &d = h->dr;
    if (!d.loaded || d.sharded || !d.d_item_paths) return fail(h, DMG_ERR_STATE, "dmg_dr_load (whole tables) and dmg_dr_load_item_paths first");
    if (n <= 0 || !seq || !target || step_t < 1 || rerank_step_t < 0 || parallelism < 1 || !out_layer_loss)
        return fail(h, DMG_ERR_INVALID_ARG, "dmg_dr_train_step: bad arguments");
    const int T = d.T, E = d.E, K = d.K, D = d.D, P = d.P, S = num_sampled, Cs = S + 1;
    if (rerank_step_t && (S < 1 || S >= d.num_item))
        return fail(h, DMG_ERR_INVALID_ARG, "numSampled %d < numClasses %d, try using Softmax directly.", S, d.num_item);   // SampledSoftmaxLoss.scala:31-34
    for (int i = 0; i < n; i++)
        if (target[i] < 0 || target[i] >= d.num_item) return fail(h, DMG_ERR_INDEX, "target item %d outside [0, %d)", target[i], d.num_item);
    if (rerank_step_t && sampled)
        for (int64_t i = 0; i < (int64_t)n * Cs; i++)
            if (sampled[i] < 0 || sampled[i] >= d.num_item) return fail(h, DMG_ERR_INDEX, "sampled item %d outside [0, %d)", sampled[i], d.num_item);
    DMG_CUDA(h, cudaSetDevice(h->device));
    DMG_TRY(ensure_train_state(h));
    const Tensors tw = tensors_of(d);
    const int iRR = 1 + 2 * D, iSM = iRR + 3;
    if (d.tr_dirty) {                                                   // zeroGradParameters (otherwise fused into the Adam pass)
        for (int i = 0; i < iSM; i++) DMG_CUDA(h, cudaMemsetAsync(d.tr_g[i], 0, (size_t)tw.n[i] * 8, h->stream));
        d.tr_dirty = false;
    }
    const int W = T + D - 1, IN = W * E, INr = T * E;
    const int64_t R = (int64_t)n * P;
    const size_t need = Carver::need({(size_t)n * T * 4, (size_t)n * 4, (size_t)n * Cs * 4, (size_t)R * W * 4, (size_t)R * D * 4, (size_t)R * IN * 8,
                                      (size_t)R * IN * 8, (size_t)R * K * 8, (size_t)R * 8, (size_t)(D + 1) * 8, (size_t)n * E * 8, (size_t)n * E * 8, (size_t)kColChunks * std::max(K, E) * 8});
    DMG_TRY(ensure_dev(h, h->s_work, need));
    Carver cw(h->s_work.d);
    int32_t *d_seq = cw.take<int32_t>((size_t)n * T), *d_tgt_item = cw.take<int32_t>(n), *d_sampled = cw.take<int32_t>((size_t)n * Cs);
    int32_t *d_idx = cw.take<int32_t>((size_t)R * W), *d_tgt = cw.take<int32_t>((size_t)R * D);
    double *d_X = cw.take<double>((size_t)R * IN), *d_GX = cw.take<double>((size_t)R * IN), *d_lg = cw.take<double>((size_t)R * K);
    double *d_row = cw.take<double>(R), *d_loss = cw.take<double>(D + 1), *d_u = cw.take<double>((size_t)n * E), *d_gu = cw.take<double>((size_t)n * E);
    double *d_part = cw.take<double>((size_t)kColChunks * std::max(K, E));
    DMG_CUDA(h, cudaMemcpyAsync(d_seq, seq, (size_t)n * T * 4, cudaMemcpyHostToDevice, h->stream));
    DMG_CUDA(h, cudaMemcpyAsync(d_tgt_item, target, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
    // the rerank Embedding has numItem rows: the stricter of the two range checks applies when the rerank model trains
    check_index_kernel<<<(unsigned)(((int64_t)n * T + 255) / 256), 256, 0, h->stream>>>(
        d_seq, (int64_t)n * T, rerank_step_t ? (int64_t)d.num_item : (int64_t)d.num_item + (int64_t)K * (D - 1), h->d_flags);
    h->launches += 1;

    // ---- layer model: trainLayerBatch + syncGradients ----
    const int task = n / parallelism, extra = n % parallelism, par = task == 0 ? extra : parallelism;       // LocalOptimizer.scala:146-148
    dr_rows_kernel<<<(unsigned)R, 128, 0, h->stream>>>(T, W, E, D, P, d.num_item, K, d_seq, d_tgt_item, d.d_item_paths, d.d_layer_emb, d_idx, d_tgt, d_X);
    DMG_CUDA(h, cudaMemsetAsync(d_GX, 0, (size_t)R * IN * 8, h->stream));
    h->launches += 1;
    if ((size_t)K * 8 > 200 * 1024) return fail(h, DMG_ERR_UNSUPPORTED, "K = %d: a logits row does not fit shared memory", K);
    DMG_CUDA(h, cudaFuncSetAttribute(dr_ce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)K * 8)));
    for (int c = 0; c < par; c++) {
        const int off = c * task + std::min(c, extra), len = task + (c < extra ? 1 : 0);
        const int64_t r0 = (int64_t)off * P, Rc = (int64_t)len * P;
        for (int l = 0; l < D; l++) {
            const int in = (T + l) * E;
            double *lg = d_lg + r0 * K;
            const double *Xc = d_X + r0 * IN;
            gemm(h, (int)Rc, K, in, Xc, IN, 1, d.d_layer_w[l], 1, in, lg, K, d.d_layer_b[l], 0);                       // logits = X . W^T + b
            dr_ce_kernel<<<(unsigned)Rc, 256, (size_t)K * 8, h->stream>>>(K, lg, d_tgt + (int64_t)l * R + r0, 1.0 / (double)Rc, d_row + r0);
            dr_loss_kernel<<<1, 32, 0, h->stream>>>(Rc, d_row + r0, d_loss + l, c > 0);
            gemm(h, K, in, (int)Rc, lg, 1, K, Xc, IN, 1, d.tr_g[1 + 2 * l], in, nullptr, 2);                           // gradWeight += gradOut^T . X
            dr_colsum_kernel<<<dim3((K + 127) / 128, kColChunks), 128, 0, h->stream>>>(Rc, K, lg, d_part);            // gradBias
            dr_colsum_finish_kernel<<<(K + 127) / 128, 128, 0, h->stream>>>(K, d_part, d.tr_g[2 + 2 * l], 1);
            gemm(h, (int)Rc, in, K, lg, K, 1, d.d_layer_w[l], in, 1, d_GX + r0 * IN, IN, nullptr, 2);                   // gradInput += gradOut . W
            h->launches += 3;
        }
    }
    dr_scatter_kernel<<<(unsigned)R, 128, 0, h->stream>>>(W, E, d_idx, d_GX, d.tr_g[0]);
    h->launches += 1;
    if (par > 1) {
        for (int i = 0; i < iRR; i++) dr_scale_kernel<<<(int)std::min<int64_t>((tw.n[i] + 255) / 256, 4096), 256, 0, h->stream>>>(d.tr_g[i], tw.n[i], (double)par);
        dr_scale_kernel<<<1, 32, 0, h->stream>>>(d_loss, D, (double)par);
        h->launches += iRR + 1;
    }
    if (apply) {
        for (int i = 0; i < iRR; i++) adam<true>(h, tw.w[i], d.tr_g[i], d.tr_s[i], d.tr_r[i], tw.n[i], lr, 1e-8, step_t);
        for (int l = 0; l < D; l++) {                                   // the beam search reads the in-major copies
            const int in = (T + l) * E;
            transpose_kernel<double><<<(K * in + 255) / 256, 256, 0, h->stream>>>(d.d_layer_w[l], d.d_layer_wT[l], K, in);
            h->launches += 1;
        }
    }

    // ---- rerank model: trainRerank + reRankOptimizer ----
    if (rerank_step_t) {
        if (sampled) DMG_CUDA(h, cudaMemcpyAsync(d_sampled, sampled, (size_t)n * Cs * 4, cudaMemcpyHostToDevice, h->stream));
        else dr_sample_kernel<<<(n + 3) / 4, 128, (size_t)4 * S * 4, h->stream>>>(n, S, d.num_item, d_tgt_item, seed, d_sampled);
        double *Xr = d_X, *GXr = d_GX;                                  // [n][T E]
        dr_rows_kernel<<<(unsigned)n, 128, 0, h->stream>>>(T, T, E, D, 0, d.num_item, K, d_seq, d_tgt_item, nullptr, d.d_rr_emb, d_idx, nullptr, Xr);
        gemm(h, n, E, INr, Xr, INr, 1, d.d_rr_w, E, 1, d_u, E, d.d_rr_b, 0);                                             // u = X . W_r^T + b_r (W_r kept [T E][E])
        const size_t smem = (size_t)(E + 2 * Cs) * 8;
        if (smem > 200 * 1024) return fail(h, DMG_ERR_UNSUPPORTED, "numSampled = %d does not fit shared memory", S);
        DMG_CUDA(h, cudaFuncSetAttribute(dr_sampled_softmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dr_sampled_softmax_kernel<<<n, 128, smem, h->stream>>>(E, Cs, d_u, d.d_sm_w, d.d_sm_b, d_sampled, 1.0 / (double)n, d_gu,
                                                               d.tr_g[iSM], d.tr_g[iSM + 1], d_row);
        dr_loss_kernel<<<1, 32, 0, h->stream>>>(n, d_row, d_loss + D, 0);
        h->launches += 4 - (sampled ? 1 : 0);
        if (apply) {                                                    // inside reRankCriterion.backward: updateParameters (gradients are never zeroed)
            adam<false>(h, d.d_sm_w, d.tr_g[iSM], d.tr_s[iSM], d.tr_r[iSM], tw.n[iSM], lr, 1e-7, rerank_step_t);
            adam<false>(h, d.d_sm_b, d.tr_g[iSM + 1], d.tr_s[iSM + 1], d.tr_r[iSM + 1], tw.n[iSM + 1], lr, 1e-7, rerank_step_t);
        }
        gemm(h, INr, E, n, Xr, 1, INr, d_gu, E, 1, d.tr_g[iRR + 1], E, nullptr, 1);                                       // gradWeight^T [T E][E] = X^T . gu
        dr_colsum_kernel<<<dim3((E + 127) / 128, kColChunks), 128, 0, h->stream>>>(n, E, d_gu, d_part);
        dr_colsum_finish_kernel<<<(E + 127) / 128, 128, 0, h->stream>>>(E, d_part, d.tr_g[iRR + 2], 0);
        gemm(h, n, INr, E, d_gu, E, 1, d.d_rr_w, 1, E, GXr, INr, nullptr, 1);                                             // gradInput = gu . W_r
        dr_scatter_kernel<<<(unsigned)n, 128, 0, h->stream>>>(T, E, d_idx, GXr, d.tr_g[iRR]);
        h->launches += 2;
        if (apply)
            for (int i = iRR; i < iRR + 3; i++) adam<true>(h, tw.w[i], d.tr_g[i], d.tr_s[i], d.tr_r[i], tw.n[i], lr, 1e-8, rerank_step_t);
    }
    if (!apply) d.tr_dirty = true;
    DMG_CUDA(h, cudaGetLastError());
    std::vector<double> loss(D + 1, 0.0);
    DMG_CUDA(h, cudaMemcpyAsync(loss.data(), d_loss, (size_t)(D + 1) * 8, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    if (h->h_flags[0]) {
        h->h_flags[0] = 0;
        return fail(h, DMG_ERR_INDEX, "dmg_dr_train_step: embeddingLookup failed, history id out of range");
    }
    for (int l = 0; l < D; l++) out_layer_loss[l] = loss[l];
    if (out_rerank_loss) *out_rerank_loss = rerank_step_t ? loss[D] : std::nan("");
    return DMG_OK;
}

/* Parameters (which = 0) or the gradients of the last dmg_dr_train_step(apply = 0) (which = 1) back to the host, same tensor
 * layout as dmg_dr_load (rr_w as [E][T E]); any pointer may be NULL. */
DMG_API int32_t dmg_dr_download(dmg_handle_t h, int32_t which, double *layer_emb, double *const *layer_w, double *const *layer_b,
                                double *rr_emb, double *rr_w, double *rr_b, double *sm_w, double *sm_b)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    DrDev &d = h->dr;
    if (!d.loaded || d.sharded) return fail(h, DMG_ERR_STATE, "dmg_dr_load first (whole tables)");
    if (which == 1 && d.tr_g.empty()) return fail(h, DMG_ERR_STATE, "no training step has run");
    if (which != 0 && which != 1) return fail(h, DMG_ERR_INVALID_ARG, "which must be 0 (parameters) or 1 (gradients)");
    DMG_CUDA(h, cudaSetDevice(h->device));
    const Tensors tw = tensors_of(d);
    auto src = [&](int i) -> const double * { return which ? d.tr_g[i] : tw.w[i]; };
    auto down = [&](double *dst, int i) -> int32_t {
        if (dst) DMG_CUDA(h, cudaMemcpyAsync(dst, src(i), (size_t)tw.n[i] * 8, cudaMemcpyDeviceToHost, h->stream));
        return DMG_OK;
    };
    const int D = d.D, iRR = 1 + 2 * D;
    DMG_TRY(down(layer_emb, 0));
    for (int l = 0; l < D; l++) {
        if (layer_w) DMG_TRY(down(layer_w[l], 1 + 2 * l));
        if (layer_b) DMG_TRY(down(layer_b[l], 2 + 2 * l));
    }
    DMG_TRY(down(rr_emb, iRR));
    double *tmp = nullptr;
    if (rr_w) {                                                         // kept [T E][E] on the device
        const int INr = d.T * d.E;
        DMG_CUDA(h, cudaMalloc(&tmp, (size_t)tw.n[iRR + 1] * 8));
        transpose_kernel<double><<<(int)((tw.n[iRR + 1] + 255) / 256), 256, 0, h->stream>>>(src(iRR + 1), tmp, INr, d.E);
        DMG_CUDA(h, cudaMemcpyAsync(rr_w, tmp, (size_t)tw.n[iRR + 1] * 8, cudaMemcpyDeviceToHost, h->stream));
    }
    DMG_TRY(down(rr_b, iRR + 2));
    DMG_TRY(down(sm_w, iRR + 3));
    DMG_TRY(down(sm_b, iRR + 4));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(tmp);
    return DMG_OK;
}
