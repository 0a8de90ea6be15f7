// train.cu -- training path (K4-K7) and JTM item->node weights (K9).
//
//   K5 din_train_kernel      forward + BCECriterionWithLogits + backward of the DIN graph on an
//                            expanded batch      (StaticGraph.scala:23-116, BCECriterionWithLogits.scala:28-91,
//                            Linear.scala:58-114, ReLU.scala:46-88, Concat.scala:45-80, MatMul.scala:46-70,
//                            SoftMax.scala:46-65, Mask.scala:35-56)
//   K6 embedding scatter-add atomicAdd into the dense node-table gradient (LookupTable.scala:56-88)
//   K7 adam_dense_kernel     dense Adam over the whole flat vector, zeroGradParameters fused
//                            (Adam.scala:19-73, AbstractModule.scala:43-50)
//   K4 tdm_sample_kernel     ancestor positives + per-level uniform negatives (NegativeSampler.scala:76-158,
//                            MiniBatch.scala:49-88)
//   K9 jtm kernels           TreeLearning.aggregateWeights (jtm/.../optim/TreeLearning.scala:152-174)
// Floating-point accumulation order differs from a single JVM thread (atomics), so training parity
// is tolerance-based (1e-5 relative, tests/test_gpu_train.py); Adam itself is elementwise and exact.
#include <algorithm>
#include <cmath>

#include "beam_kernels.cuh"
#include "rows_kernels.cuh"
#include "shard_common.cuh"

using namespace dmg;

int32_t dmg_refresh_transposes(dmg_handle_t h);     // capi.cu
int32_t dmg_tdm_ids_to_codes(dmg_handle_t h, const int32_t *d_ids, int64_t n, int use_mask, int32_t *d_codes, uint8_t *d_mask);  // capi.cu

namespace {

constexpr int kTrThreads = 128;
constexpr int kTrRB = 8;

__device__ __forceinline__ float log1pexp_(float e) { return logf(1.0f + e); }
__device__ __forceinline__ double log1pexp_(double e) { return log(1.0 + e); }
__device__ __forceinline__ float sqrt_(float x) { return __fsqrt_rn(x); }
__device__ __forceinline__ double sqrt_(double x) { return __dsqrt_rn(x); }
__device__ __forceinline__ float div_(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_(double a, double b) { return __ddiv_rn(a, b); }

template <typename real> struct TrainParams {
    const real *emb, *watt, *w1, *b1, *w2, *b2;     // row-major [out][in]
    const real *wattT, *w1T;                        // [in][out]
    real scale;
    int E, T;
    int64_t n;
    double n_norm;                                  // the mean of BCECriterionWithLogits divides by this (n, or the global batch when rows are split over ranks)
    const int32_t *node, *seq;
    const uint8_t *mask;
    const real *labels;
    real *g_emb, *g_watt, *g_w1, *g_b1, *g_w2, *g_b2;
    double *loss_acc;                               // sum of per-row losses
};

template <typename real>
__global__ void __launch_bounds__(kTrThreads) din_train_kernel(const TrainParams<real> p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int E = p.E, T = p.T, PL = T + 1;
    real *sQ = reinterpret_cast<real *>(smem_raw);          // RB x E
    real *sK = sQ + kTrRB * E;                               // RB x T x E
    real *sP = sK + kTrRB * T * E;                           // RB x PL   probabilities
    real *sA = sP + kTrRB * PL;                              // RB x E
    real *sAtt = sA + kTrRB * E;                             // RB x E
    real *sZ = sAtt + kTrRB * E;                             // RB x E    pre-activation
    real *sDz = sZ + kTrRB * E;                              // RB x E
    real *sDx = sDz + kTrRB * E;                             // RB x 2E
    real *sDa = sDx + kTrRB * 2 * E;                         // RB x E
    real *sDs = sDa + kTrRB * E;                             // RB x PL   dp then ds
    real *sDy = sDs + kTrRB * PL;                            // RB
    real *gWatt = sDy + kTrRB;                               // E*E
    real *gW1 = gWatt + E * E;                               // 2*E*E
    real *gB1 = gW1 + 2 * E * E;                             // E
    real *gW2 = gB1 + E;                                     // E
    real *gB2 = gW2 + E;                                     // 1 (+1 loss)
    const int tid = threadIdx.x;
    for (int i = tid; i < 3 * E * E + 2 * E + 2; i += kTrThreads) gWatt[i] = (real)0;
    double loss_local = 0.0;
    const real inv_n = (real)(1.0 / p.n_norm);
    __syncthreads();

    for (int64_t g0 = (int64_t)blockIdx.x * kTrRB; g0 < p.n; g0 += (int64_t)gridDim.x * kTrRB) {
        const int nr = (int)((p.n - g0) < kTrRB ? (p.n - g0) : kTrRB);
        for (int idx = tid; idx < nr * (T + 1) * E; idx += kTrThreads) {
            const int k = idx % E, slot = (idx / E) % (T + 1), r = idx / (E * (T + 1));
            const int32_t c = slot == 0 ? p.node[g0 + r] : p.seq[(g0 + r) * T + slot - 1];
            const real v = c < 0 ? (real)0 : p.emb[(size_t)c * E + k];
            if (slot == 0) sQ[r * E + k] = v; else sK[(r * T + slot - 1) * E + k] = v;
        }
        __syncthreads();
        // ---- forward (same chains as rows_kernels.cuh) ----
        for (int idx = tid; idx < nr * T; idx += kTrThreads) {
            const int r = idx / T, j = idx % T;
            const real *q = sQ + r * E, *kj = sK + (r * T + j) * E;
            real acc = (real)0;
            for (int k = 0; k < E; k++) acc = fma_(q[k], kj[k], acc);
            real s = mul_(acc, p.scale);
            if (p.mask[(g0 + r) * T + j]) s = mask_value<real>::get();
            sP[r * PL + j] = s;
        }
        __syncthreads();
        if (tid < nr) {
            real *pr = sP + tid * PL;
            real mx = pr[0];
            for (int j = 1; j < T; j++) { real v = pr[j]; mx = v > mx ? v : mx; }
            real sum = (real)0;
            for (int j = 0; j < T; j++) { real e = exp_(sub_(pr[j], mx)); pr[j] = e; sum = add_(sum, e); }
            const real inv = inv_(sum);
            for (int j = 0; j < T; j++) pr[j] = mul_(pr[j], inv);
        }
        __syncthreads();
        for (int idx = tid; idx < nr * E; idx += kTrThreads) {
            const int r = idx / E, k = idx % E;
            real acc = (real)0;
            for (int j = 0; j < T; j++) acc = fma_(sP[r * PL + j], sK[(r * T + j) * E + k], acc);
            sA[idx] = acc;
        }
        __syncthreads();
        for (int idx = tid; idx < nr * E; idx += kTrThreads) {
            const int r = idx / E, o = idx % E;
            real acc = (real)0;
            for (int k = 0; k < E; k++) acc = fma_(sA[r * E + k], __ldg(p.wattT + (size_t)k * E + o), acc);
            sAtt[idx] = acc;
        }
        __syncthreads();
        for (int idx = tid; idx < nr * E; idx += kTrThreads) {
            const int r = idx / E, o = idx % E;
            real acc = (real)0;
            for (int k = 0; k < E; k++) acc = fma_(sQ[r * E + k], __ldg(p.w1T + (size_t)k * E + o), acc);
            for (int k = 0; k < E; k++) acc = fma_(sAtt[r * E + k], __ldg(p.w1T + (size_t)(E + k) * E + o), acc);
            sZ[idx] = add_(acc, __ldg(p.b1 + o));
        }
        __syncthreads();
        if (tid < nr) {
            real y = (real)0;
            for (int o = 0; o < E; o++) y = fma_(relu_(sZ[tid * E + o]), __ldg(p.w2 + o), y);
            y = add_(y, __ldg(p.b2));
            const real t = p.labels[g0 + tid];
            const real ay = y < (real)0 ? -y : y;
            loss_local += (double)((y > (real)0 ? y : (real)0) - y * t + log1pexp_(exp_(-ay)));
            sDy[tid] = mul_(sub_(div_((real)1, add_((real)1, exp_(-y))), t), inv_n);
        }
        __syncthreads();
        // ---- backward ----
        for (int idx = tid; idx < nr * E; idx += kTrThreads) {
            const int r = idx / E, o = idx % E;
            sDz[idx] = sZ[idx] <= (real)0 ? (real)0 : mul_(sDy[r], __ldg(p.w2 + o));
        }
        __syncthreads();
        for (int o = tid; o < E; o += kTrThreads) {
            real a2 = (real)0, a1 = (real)0;
            for (int r = 0; r < nr; r++) { a2 = fma_(sDy[r], relu_(sZ[r * E + o]), a2); a1 = add_(a1, sDz[r * E + o]); }
            gW2[o] = add_(gW2[o], a2);
            gB1[o] = add_(gB1[o], a1);
        }
        if (tid == 0) { real a = (real)0; for (int r = 0; r < nr; r++) a = add_(a, sDy[r]); gB2[0] = add_(gB2[0], a); }
        for (int idx = tid; idx < 2 * E * E; idx += kTrThreads) {      // gW1[o][k] += dz[o] * x[k]
            const int o = idx / (2 * E), k = idx % (2 * E);
            real acc = (real)0;
            for (int r = 0; r < nr; r++) acc = fma_(sDz[r * E + o], k < E ? sQ[r * E + k] : sAtt[r * E + k - E], acc);
            gW1[idx] = add_(gW1[idx], acc);
        }
        for (int idx = tid; idx < nr * 2 * E; idx += kTrThreads) {     // dx = W1^T dz
            const int r = idx / (2 * E), k = idx % (2 * E);
            real acc = (real)0;
            for (int o = 0; o < E; o++) acc = fma_(sDz[r * E + o], __ldg(p.w1 + (size_t)o * 2 * E + k), acc);
            sDx[idx] = acc;
        }
        __syncthreads();
        for (int idx = tid; idx < E * E; idx += kTrThreads) {          // gWatt[o][k] += datt[o] * a[k]
            const int o = idx / E, k = idx % E;
            real acc = (real)0;
            for (int r = 0; r < nr; r++) acc = fma_(sDx[r * 2 * E + E + o], sA[r * E + k], acc);
            gWatt[idx] = add_(gWatt[idx], acc);
        }
        for (int idx = tid; idx < nr * E; idx += kTrThreads) {         // da = Watt^T datt
            const int r = idx / E, k = idx % E;
            real acc = (real)0;
            for (int o = 0; o < E; o++) acc = fma_(sDx[r * 2 * E + E + o], __ldg(p.watt + (size_t)o * E + k), acc);
            sDa[idx] = acc;
        }
        __syncthreads();
        for (int idx = tid; idx < nr * T; idx += kTrThreads) {         // dp_j = da . K_j
            const int r = idx / T, j = idx % T;
            real acc = (real)0;
            for (int k = 0; k < E; k++) acc = fma_(sDa[r * E + k], sK[(r * T + j) * E + k], acc);
            sDs[r * PL + j] = acc;
        }
        __syncthreads();
        if (tid < nr) {                                                // softmax + mask backward
            real *ds = sDs + tid * PL;
            const real *pr = sP + tid * PL;
            real dot = (real)0;
            for (int j = 0; j < T; j++) dot = fma_(ds[j], pr[j], dot);
            for (int j = 0; j < T; j++) {
                real v = mul_(mul_(sub_(ds[j], dot), pr[j]), p.scale);
                if (p.mask[(g0 + tid) * T + j]) v = (real)0;
                ds[j] = v;
            }
        }
        __syncthreads();
        for (int idx = tid; idx < nr * (T + 1) * E; idx += kTrThreads) {   // K6: scatter-add into the table
            const int k = idx % E, slot = (idx / E) % (T + 1), r = idx / (E * (T + 1));
            if (slot == 0) {
                const int32_t c = p.node[g0 + r];
                if (c >= 0) {
                    real acc = sDx[r * 2 * E + k];
                    for (int j = 0; j < T; j++) acc = fma_(sDs[r * PL + j], sK[(r * T + j) * E + k], acc);
                    atomicAdd(p.g_emb + (size_t)c * E + k, acc);
                }
            } else {
                const int j = slot - 1;
                const int32_t c = p.seq[(g0 + r) * T + j];
                if (c >= 0) {
                    real v = mul_(sP[r * PL + j], sDa[r * E + k]);
                    v = fma_(sDs[r * PL + j], sQ[r * E + k], v);
                    atomicAdd(p.g_emb + (size_t)c * E + k, v);
                }
            }
        }
        __syncthreads();
    }
    for (int i = tid; i < E * E; i += kTrThreads) atomicAdd(p.g_watt + i, gWatt[i]);
    for (int i = tid; i < 2 * E * E; i += kTrThreads) atomicAdd(p.g_w1 + i, gW1[i]);
    for (int i = tid; i < E; i += kTrThreads) { atomicAdd(p.g_b1 + i, gB1[i]); atomicAdd(p.g_w2 + i, gW2[i]); }
    if (tid == 0) atomicAdd(p.g_b2, gB2[0]);
    // loss: warp reduce then one atomic per warp
    for (int o = 16; o > 0; o >>= 1) loss_local += __shfl_xor_sync(0xffffffffu, loss_local, o);
    if ((tid & 31) == 0 && loss_local != 0.0) atomicAdd(p.loss_acc, loss_local);
}

// ---- DeepFM in the training loop (tdm/.../model/DeepFM.scala:11-44; FM.scala:46-72 backward) -----------------------------------
// 8 rows per CTA iteration: features F = [item row ; T history rows] in shared memory, forward as deepfm_rows_forward_kernel (chains
// over Fflat in order), BCE-with-logits (mean over n_norm), backward of Add / Linear(T+1, 1) / ReLU / Linear((T+1)E, T+1) / FM,
// dense-weight gradients reduced per CTA in shared memory then atomicAdd, embedding gradients scatter-added (padding skipped).
constexpr int kDfmRB = 8;
__global__ void __launch_bounds__(kTrThreads) deepfm_train_kernel(const float *__restrict__ emb, const float *__restrict__ dense, int E, int T, int64_t n,
                                                                   double n_norm, const int32_t *__restrict__ node, const int32_t *__restrict__ seq,
                                                                   const float *__restrict__ labels, float *__restrict__ g_emb,
                                                                   float *__restrict__ g_dense, double *__restrict__ loss_acc)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int F = T + 1, IN = F * E, tid = threadIdx.x;
    float *sF = reinterpret_cast<float *>(smem_raw);             // RB x IN
    float *sBuf = sF + kDfmRB * IN;                              // RB x E   sum of the feature rows (FM buffer)
    float *sZ = sBuf + kDfmRB * E;                               // RB x F   pre-activations
    float *sDz = sZ + kDfmRB * F;                                // RB x F
    float *sDy = sDz + kDfmRB * F;                               // RB
    float *gW1 = sDy + kDfmRB;                                   // F x IN | b1 F | w2 F | b2 1: this CTA's share of the dense gradient
    const int n_dense = F * IN + 2 * F + 1;
    const float *w1 = dense, *b1 = w1 + (size_t)F * IN, *w2 = b1 + F, *b2 = w2 + F;
    float *gB1 = gW1 + F * IN, *gW2 = gB1 + F, *gB2 = gW2 + F;
    for (int i = tid; i < n_dense; i += kTrThreads) gW1[i] = 0.0f;
    const float inv_n = (float)(1.0 / n_norm);
    double loss_local = 0.0;
    __syncthreads();
    for (int64_t g0 = (int64_t)blockIdx.x * kDfmRB; g0 < n; g0 += (int64_t)gridDim.x * kDfmRB) {
        const int nr = (int)((n - g0) < kDfmRB ? (n - g0) : kDfmRB);
        for (int idx = tid; idx < nr * IN; idx += kTrThreads) {
            const int r = idx / IN, q = idx - r * IN, slot = q / E;
            const int32_t c = slot == 0 ? node[g0 + r] : seq[(g0 + r) * T + slot - 1];
            sF[idx] = c < 0 ? 0.0f : emb[(size_t)c * E + (q - slot * E)];
        }
        __syncthreads();
        for (int idx = tid; idx < nr * E; idx += kTrThreads) {    // FM buffer: vAdd in row order from zero
            const int r = idx / E, k = idx - r * E;
            float b = 0.0f;
            for (int s_ = 0; s_ < F; s_++) b = add_(b, sF[r * IN + s_ * E + k]);
            sBuf[idx] = b;
        }
        for (int idx = tid; idx < nr * F; idx += kTrThreads) {    // Linear((T+1)E, T+1): one chain per (row, output)
            const int r = idx % nr, o = idx / nr;
            const float *x = sF + r * IN, *w = w1 + (size_t)o * IN;
            float acc = 0.0f;
            for (int k = 0; k < IN; k++) acc = fma_(x[k], __ldg(w + k), acc);
            sZ[r * F + o] = add_(acc, b1[o]);
        }
        __syncthreads();
        if (tid < nr) {
            const int r = tid;
            const float *x = sF + r * IN, *bf = sBuf + r * E;
            float sum_square = 0.0f, square_sum = 0.0f;
            for (int k = 0; k < E; k++) sum_square = fma_(bf[k], bf[k], sum_square);
            for (int k = 0; k < IN; k++) square_sum = fma_(x[k], x[k], square_sum);
            const float fm = div_(sub_(sum_square, square_sum), 2.0f);
            float dnn = 0.0f;
            for (int o = 0; o < F; o++) { const float z = sZ[r * F + o]; dnn = fma_(z > 0.0f ? z : 0.0f, w2[o], dnn); }
            const float y = add_(fm, add_(dnn, b2[0]));
            const float t = labels[g0 + r], ay = y < 0.0f ? -y : y;
            loss_local += (double)add_(sub_(y > 0.0f ? y : 0.0f, mul_(y, t)), log1pexp_(exp_(-ay)));
            sDy[r] = mul_(sub_(div_(1.0f, add_(1.0f, exp_(-y))), t), inv_n);
        }
        __syncthreads();
        for (int idx = tid; idx < nr * F; idx += kTrThreads) {
            const int r = idx / F, o = idx - r * F;
            sDz[idx] = sZ[idx] <= 0.0f ? 0.0f : mul_(sDy[r], w2[o]);
        }
        __syncthreads();
        if (tid < F) {                                            // Linear(T+1, 1) weight / bias, first Linear's bias
            float a = gW2[tid], b = gB1[tid];
            for (int r = 0; r < nr; r++) { const float z = sZ[r * F + tid]; a = fma_(sDy[r], z > 0.0f ? z : 0.0f, a); b = add_(b, sDz[r * F + tid]); }
            gW2[tid] = a; gB1[tid] = b;
        } else if (tid == F) {
            float a = gB2[0];
            for (int r = 0; r < nr; r++) a = add_(a, sDy[r]);
            gB2[0] = a;
        }
        for (int idx = tid; idx < F * IN; idx += kTrThreads) {    // gradWeight of the first Linear: dz^T . Fflat
            const int o = idx / IN, k = idx - o * IN;
            float a = gW1[idx];
            for (int r = 0; r < nr; r++) a = fma_(sDz[r * F + o], sF[r * IN + k], a);
            gW1[idx] = a;
        }
        for (int idx = tid; idx < nr * IN; idx += kTrThreads) {   // gradInput: DNN branch + FM branch, scattered into the table gradient
            const int r = idx / IN, q = idx - r * IN, slot = q / E, k = q - slot * E;
            const int32_t c = slot == 0 ? node[g0 + r] : seq[(g0 + r) * T + slot - 1];
            if (c < 0) continue;
            float acc = 0.0f;
            for (int o = 0; o < F; o++) acc = fma_(sDz[r * F + o], __ldg(w1 + (size_t)o * IN + q), acc);
            acc = add_(acc, mul_(sub_(sBuf[r * E + k], sF[idx]), sDy[r]));
            atomicAdd(g_emb + (size_t)c * E + k, acc);
        }
        __syncthreads();
    }
    for (int i = tid; i < n_dense; i += kTrThreads) { const float v = gW1[i]; if (v != 0.0f) atomicAdd(g_dense + i, v); }
    for (int o = 16; o > 0; o >>= 1) loss_local += __shfl_xor_sync(0xffffffffu, loss_local, o);
    if ((tid & 31) == 0 && loss_local != 0.0) atomicAdd(loss_acc, loss_local);
}

// K7: dense Adam, gradient zeroed in the same pass.  Pure streaming (4 reads + 4 writes per parameter, no reuse):
// 16-byte vector loads, four vectors per thread in flight, so that ~64 KB per SM is outstanding -- what HBM3e needs
// at ~800 ns latency; the scalar form moved 2.9 TB/s, this one is bound by the copy bandwidth.
template <typename real>
__device__ __forceinline__ void adam_one(real &w, real &g, real &s, real &r, real b1, real omb1, real b2, real omb2, real eps, real nstep)
{
    const real si = add_(mul_(s, b1), mul_(omb1, g));
    const real ri = add_(mul_(r, b2), mul_(omb2, mul_(g, g)));
    const real denom = add_(sqrt_(ri), eps);
    w = add_(w, mul_(nstep, div_(si, denom)));
    s = si; r = ri; g = (real)0;
}
template <typename real>
__global__ void __launch_bounds__(256) adam_dense_kernel(real *__restrict__ w, real *__restrict__ g, real *__restrict__ s, real *__restrict__ r,
                                                         int64_t n, real b1, real omb1, real b2, real omb2, real eps, real nstep)
{
    constexpr int V = 16 / sizeof(real);
    struct alignas(16) Vec { real v[V]; };
    const int64_t nv = n / V;
    Vec *wv = reinterpret_cast<Vec *>(w), *gv = reinterpret_cast<Vec *>(g), *sv = reinterpret_cast<Vec *>(s), *rv = reinterpret_cast<Vec *>(r);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        Vec a = wv[i], b = gv[i], c = sv[i], d = rv[i];
#pragma unroll
        for (int q = 0; q < V; q++) adam_one(a.v[q], b.v[q], c.v[q], d.v[q], b1, omb1, b2, omb2, eps, nstep);
        wv[i] = a; gv[i] = b; sv[i] = c; rv[i] = d;
    }
    for (int64_t i = nv * V + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        adam_one(w[i], g[i], s[i], r[i], b1, omb1, b2, omb2, eps, nstep);
}

// K4: one thread per (target, level): positive = ancestor at that level, then neg distinct uniform
// codes of the level that exist and differ from the positive, emitted in ascending order.
__global__ void tdm_sample_kernel(int n_targets, const int32_t *__restrict__ target_items, const int32_t *__restrict__ id_code,
                                  int32_t non_leaf_offset, const uint32_t *__restrict__ exists, int max_level,
                                  const int32_t *__restrict__ layer_neg, const int32_t *__restrict__ level_off,
                                  int start_level, int layer_sum, uint64_t seed, int32_t *__restrict__ out_node,
                                  float *__restrict__ out_label, int32_t *__restrict__ err_flag,
                                  const double *__restrict__ cdf, int tolerance)
{
    const int n_lv = max_level + 1 - start_level;
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n_targets * n_lv) return;
    const int t = gid / n_lv, level = start_level + gid % n_lv;
    const int32_t item = target_items[t];
    int32_t leaf = (item > 0 && item < non_leaf_offset) ? id_code[item] : -1;
    if (leaf < 0) { atomicExch(err_flag, 1); return; }
    int64_t pos = leaf;
    for (int l = max_level; l > level; l--) pos = (pos - 1) >> 1;       // TDMTree.pathNodes upTrace
    const int neg = layer_neg[level];
    int32_t *dst = out_node + (size_t)t * layer_sum + level_off[level];
    float *lab = out_label + (size_t)t * layer_sum + level_off[level];
    dst[0] = (int32_t)pos;
    lab[0] = 1.0f;
    const int64_t lstart = ((int64_t)1 << level) - 1, lsize = (int64_t)1 << level;
    int got = 0;
    uint64_t ctr = 0;
    const uint64_t key = splitmix64(seed ^ splitmix64(((uint64_t)t << 8) | (uint64_t)level));
    const int max_try = 64 * (neg + 4);
    if (cdf) {
        // sampleFromCategoricalDistribution (NegativeSampler.scala:116-144): at most neg + tolerance draws from the level's
        // EnumeratedIntegerDistribution over Node.probality, keeping new codes != the positive; what is still missing is drawn
        // uniformly over the level WITHOUT excluding the positive (the reference's fallback loop only asks codeNodeMap.contains)
        const double total = cdf[lstart + lsize - 1];
        int tries = 0;
        while (total > 0.0 && got < neg && tries < neg + tolerance) {
            tries++;
            const double u = (double)(splitmix64(key + ctr++) >> 11) * (1.0 / 9007199254740992.0) * total;
            int64_t lo = 0, hi = lsize - 1;                              // first code of the level with cdf > u
            while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (cdf[lstart + mid] > u) hi = mid; else lo = mid + 1; }
            const int64_t c = lstart + lo;
            if (c == pos) continue;
            bool dup = false;
            for (int i = 0; i < got; i++) dup |= (dst[1 + i] == (int32_t)c);
            if (dup) continue;
            int i = got++;
            while (i > 0 && dst[i] > (int32_t)c) { dst[1 + i] = dst[i]; i--; }
            dst[1 + i] = (int32_t)c;
        }
        int guard = 0;
        while (got < neg && guard++ < max_try) {
            const int64_t c = lstart + (int64_t)(splitmix64(key + ctr++) % (uint64_t)lsize);
            if (!code_exists(exists, c)) continue;
            bool dup = false;
            for (int i = 0; i < got; i++) dup |= (dst[1 + i] == (int32_t)c);
            if (dup) continue;
            int i = got++;
            while (i > 0 && dst[i] > (int32_t)c) { dst[1 + i] = dst[i]; i--; }
            dst[1 + i] = (int32_t)c;
        }
        for (int i = got; i < neg; i++) dst[1 + i] = -1;                // fewer nodes than negatives: padding rows
        for (int i = 0; i < neg; i++) lab[1 + i] = 0.0f;
        return;
    }
    while (got < neg && (int)ctr < max_try) {
        const int64_t c = lstart + (int64_t)(splitmix64(key + ctr++) % (uint64_t)lsize);
        if (c == pos || !code_exists(exists, c)) continue;
        bool dup = false;
        for (int i = 0; i < got; i++) dup |= (dst[1 + i] == (int32_t)c);
        if (dup) continue;
        int i = got++;                                                  // insertion keeps ascending order (BitSet.toList)
        while (i > 0 && dst[i] > (int32_t)c) { dst[1 + i] = dst[i]; i--; }
        dst[1 + i] = (int32_t)c;
    }
    if (got < neg) {                                                    // sparse level: take what exists, in order
        for (int64_t c = lstart; c < lstart + lsize && got < neg; c++) {
            if (c == pos || !code_exists(exists, c)) continue;
            bool dup = false;
            for (int i = 0; i < got; i++) dup |= (dst[1 + i] == (int32_t)c);
            if (dup) continue;
            int i = got++;
            while (i > 0 && dst[i] > (int32_t)c) { dst[1 + i] = dst[i]; i--; }
            dst[1 + i] = (int32_t)c;
        }
        for (int i = got; i < neg; i++) dst[1 + i] = -1;                // fewer nodes than negatives: padding rows
    }
    for (int i = 0; i < neg; i++) lab[1 + i] = 0.0f;
}

// history of target t repeated layer_sum times (MiniBatch.transformWithMask)
__global__ void repeat_seq_kernel(const int32_t *__restrict__ codes, const uint8_t *__restrict__ mask, int n_targets, int T,
                                  int layer_sum, int32_t *__restrict__ out_seq, uint8_t *__restrict__ out_mask)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = (int64_t)n_targets * layer_sum * T;
    if (i >= n) return;
    const int j = (int)(i % T);
    const int64_t t = i / ((int64_t)layer_sum * T);
    out_seq[i] = codes[t * T + j];
    if (out_mask) out_mask[i] = mask[t * T + j];
}

// K9 second half: per (item, node) the in-order fp32 sum of the item's sample logits (Tensor.sum),
// then per (item, child) the in-order sum along child -> parent (TreeLearning.scala:163-172).
__global__ void jtm_reduce_kernel(int n_items, const int64_t *__restrict__ sample_off, const float *__restrict__ logits,
                                  int n_nodes /* nodes per item = 2^(gap+1)-2 */, int gap, float *__restrict__ node_sum,
                                  float *__restrict__ out_weights)
{
    // phase A: one thread per (item, node)
    const int64_t total = (int64_t)n_items * n_nodes;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int item = (int)(g / n_nodes), nd = (int)(g % n_nodes);
        const int64_t s0 = sample_off[item], s1 = sample_off[item + 1];
        // logits layout: for item i, block of (s1-s0)*n_nodes values, node-major [nd][sample]
        const float *src = logits + s0 * n_nodes + (int64_t)nd * (s1 - s0);
        float acc = 0.0f;
        for (int64_t k = 0; k < s1 - s0; k++) acc = __fadd_rn(acc, src[k]);
        node_sum[g] = acc;
    }
    (void)gap; (void)out_weights;
}

__global__ void jtm_path_kernel(int n_items, const int64_t *__restrict__ sample_off, int gap, int n_nodes,
                                const float *__restrict__ node_sum, float *__restrict__ out_weights)
{
    const int n_child = 1 << gap;
    const int64_t total = (int64_t)n_items * n_child;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int item = (int)(g / n_child), ch = (int)(g % n_child);
        if (sample_off[item + 1] == sample_off[item]) { out_weights[g] = -1e6f; continue; }   // TreeLearning.scala:160
        // local heap index inside the parent's subtree: root = 0, children 2i+1, 2i+2; node slot = idx-1
        int idx = (1 << gap) - 1 + ch;
        float w = 0.0f;
        while (idx > 0) { w = __fadd_rn(w, node_sum[(int64_t)item * n_nodes + idx - 1]); idx = (idx - 1) >> 1; }
        out_weights[g] = w;
    }
}

// rows for K9: (item-sample, subtree node) -> node code + history codes at the node's level
__global__ void jtm_rows_kernel(int n_items, const int64_t *__restrict__ sample_off, const int32_t *__restrict__ sample_seq,
                                const int32_t *__restrict__ parent_code, int old_level, int gap, int n_nodes, int T,
                                const int32_t *__restrict__ id_code, int32_t non_leaf_offset, int32_t max_code, int max_level,
                                int hierarchical, int min_level, int use_mask, const int64_t *__restrict__ item_of_row,
                                int64_t n_rows, int32_t *__restrict__ out_node, int32_t *__restrict__ out_seq,
                                uint8_t *__restrict__ out_mask)
{
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n_rows; row += (int64_t)gridDim.x * blockDim.x) {
        const int item = (int)item_of_row[row];
        const int64_t s0 = sample_off[item], ns = sample_off[item + 1] - s0;
        const int64_t local = row - s0 * n_nodes;
        const int nd = (int)(local / ns);
        const int64_t smp = s0 + local % ns;
        // subtree slot nd -> heap index nd+1 -> depth below the parent and offset within that depth
        const int hidx = nd + 1;
        const int depth = 31 - __clz(hidx + 1);
        const int64_t first = ((int64_t)parent_code[item] + 1) * ((int64_t)1 << depth) - 1;   // leftmost descendant at that depth
        const int32_t code = (int32_t)(first + (hidx + 1 - (1 << depth)));
        const int level = old_level + depth;
        out_node[row] = code;
        for (int j = 0; j < T; j++) {                                   // JTMTree.idToCodeWithMask :86-113
            const int32_t id = sample_seq[smp * T + j];
            int32_t c;
            uint8_t m = 0;
            if (id == 0) { c = -1; m = 1; }
            else if (id > 0 && id < non_leaf_offset && id_code[id] >= 0) {
                c = id_code[id];
                if (hierarchical && level >= min_level) {               // getAncestorAtLevel :36-43
                    const int64_t lim = ((int64_t)1 << (level + 1)) - 1;
                    int64_t cc = c;
                    while (cc >= lim) cc = (cc - 1) >> 1;
                    c = (int32_t)cc;
                }
            } else {
                int64_t tmp = (int64_t)id - non_leaf_offset;
                c = tmp > max_code ? -1 : (int32_t)tmp;                 // NB: not added to the mask (JTMTree.scala:104-107)
            }
            out_seq[row * T + j] = c;
            out_mask[row * T + j] = use_mask ? m : 0;
        }
        (void)gap; (void)max_level;
    }
}

template <typename real> int32_t ensure_train_state(dmg_handle_t h)
{
    DinDev &d = h->din;
    if (d.d_grad) return DMG_OK;
    const size_t bytes = (size_t)d.n_params * sizeof(real);
    DMG_CUDA(h, cudaMalloc(&d.d_grad, bytes));
    DMG_CUDA(h, cudaMalloc(&d.d_m, bytes));
    DMG_CUDA(h, cudaMalloc(&d.d_v, bytes));
    DMG_CUDA(h, cudaMemsetAsync(d.d_grad, 0, bytes, h->stream));
    DMG_CUDA(h, cudaMemsetAsync(d.d_m, 0, bytes, h->stream));
    DMG_CUDA(h, cudaMemsetAsync(d.d_v, 0, bytes, h->stream));
    return DMG_OK;
}

// uploads (node, seq, mask, labels), validates, runs K5/K6 into d_grad.  Loss (mean) -> *loss_out.
// forward + BCE + backward on device-resident rows: gradients accumulate into d.d_grad, the summed loss into d_loss
template <typename real>
int32_t grad_enqueue(dmg_handle_t h, int64_t n, const int32_t *dn, const int32_t *ds, const uint8_t *d_mask, const real *dl, double *d_loss,
                     const real *emb_override = nullptr, real *g_emb_override = nullptr, double n_norm = 0.0)
{
    DinDev &d = h->din;
    const int E = d.E, T = d.T;
    if (d.kind == 1) {                                            // DeepFM scorer: no mask input (DeepFM.scala:14-15)
        if (sizeof(real) != 4 || emb_override) return fail(h, DMG_ERR_UNSUPPORTED, "DeepFM training is built for the Float model on an unsharded table");
        const int F = T + 1, IN = F * E;
        const size_t smem = ((size_t)kDfmRB * (IN + E + 2 * F + 1) + (size_t)F * IN + 2 * F + 1) * 4;
        if (smem > h->smem_optin) return fail(h, DMG_ERR_UNSUPPORTED, "embed_size %d / seq_len %d too large for the DeepFM training kernel", E, T);
        DMG_CUDA(h, cudaFuncSetAttribute(deepfm_train_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int grid = (int)std::min<int64_t>((n + kDfmRB - 1) / kDfmRB, (int64_t)h->sm_count * 2);
        float *g = (float *)d.d_grad;
        deepfm_train_kernel<<<grid, kTrThreads, smem, h->stream>>>(d.emb<float>(), d.tail<float>(), E, T, n, n_norm > 0.0 ? n_norm : (double)n, dn, ds,
                                                                   (const float *)dl, g, g + d.rows * E, d_loss);
        h->launches += 1;
        DMG_CUDA(h, cudaGetLastError());
        return DMG_OK;
    }
    TrainParams<real> p;
    p.emb = d.emb<real>(); p.watt = d.watt<real>(); p.w1 = d.w1<real>(); p.b1 = d.b1<real>(); p.w2 = d.w2<real>(); p.b2 = d.b2<real>();
    p.wattT = (const real *)d.d_wattT; p.w1T = (const real *)d.d_w1T;
    p.scale = (real)(1.0 / std::sqrt((double)E));
    p.E = E; p.T = T; p.n = n; p.node = dn; p.seq = ds; p.mask = d_mask; p.labels = dl;
    p.n_norm = n_norm > 0.0 ? n_norm : (double)n;
    if (emb_override) p.emb = emb_override;          // rows staged per occurrence (sharded table): indices are occurrence ids
    real *g = (real *)d.d_grad;
    p.g_emb = g_emb_override ? g_emb_override : g; p.g_watt = g + d.rows * E; p.g_w1 = p.g_watt + (int64_t)E * E; p.g_b1 = p.g_w1 + (int64_t)2 * E * E;
    p.g_w2 = p.g_b1 + E; p.g_b2 = p.g_w2 + E;
    p.loss_acc = d_loss;
    const size_t smem = ((size_t)kTrRB * ((size_t)8 * E + (size_t)T * E + 2 * (T + 1) + 1) + 3 * (size_t)E * E + 2 * E + 2) * sizeof(real);
    if (smem > h->smem_optin) return fail(h, DMG_ERR_UNSUPPORTED, "embed_size %d too large for the training kernel", E);
    auto kern = din_train_kernel<real>;
    DMG_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<int64_t>((n + kTrRB - 1) / kTrRB, (int64_t)h->sm_count * 4);
    kern<<<grid, kTrThreads, smem, h->stream>>>(p);
    h->launches += 1;
    DMG_CUDA(h, cudaGetLastError());
    return DMG_OK;
}

template <typename real>
int32_t grad_pass(dmg_handle_t h, int64_t n, const int32_t *node, const int32_t *seq, const int32_t *mask_flat, int64_t n_mask,
                  const real *labels, double *loss_out)
{
    DinDev &d = h->din;
    const int E = d.E, T = d.T;
    DMG_TRY(ensure_train_state<real>(h));
    const size_t b_node = (size_t)n * 4, b_seq = (size_t)n * T * 4, b_mask = (size_t)n_mask * 4, b_lab = (size_t)n * sizeof(real);
    const size_t in_bytes = Carver::need({b_node, b_seq, b_mask, b_lab});
    DMG_TRY(ensure_host(h, h->s_in, in_bytes));
    DMG_TRY(ensure_dev(h, h->s_in, in_bytes));
    Carver ch(h->s_in.h), cd(h->s_in.d);
    int32_t *hn = ch.take<int32_t>((size_t)n), *dn = cd.take<int32_t>((size_t)n);
    int32_t *hs = ch.take<int32_t>((size_t)n * T), *ds = cd.take<int32_t>((size_t)n * T);
    int32_t *hm = ch.take<int32_t>((size_t)n_mask), *dm = cd.take<int32_t>((size_t)n_mask);
    real *hl = ch.take<real>((size_t)n), *dl = cd.take<real>((size_t)n);
    memcpy(hn, node, b_node); memcpy(hs, seq, b_seq); memcpy(hl, labels, b_lab);
    if (n_mask) memcpy(hm, mask_flat, b_mask);
    DMG_CUDA(h, cudaMemcpyAsync(h->s_in.d, h->s_in.h, ch.off, cudaMemcpyHostToDevice, h->stream));
    DMG_TRY(ensure_dev(h, h->s_work, Carver::need({(size_t)n * T, 64})));
    Carver cw(h->s_work.d);
    uint8_t *d_mask = cw.take<uint8_t>((size_t)n * T);
    double *d_loss = cw.take<double>(1);
    DMG_CUDA(h, cudaMemsetAsync(d_mask, 0, (size_t)n * T, h->stream));
    DMG_CUDA(h, cudaMemsetAsync(d_loss, 0, sizeof(double), h->stream));
    if (n_mask) {
        mask_scatter_kernel<<<(unsigned)((n_mask + 255) / 256), 256, 0, h->stream>>>(dm, n_mask, n * T, d_mask, h->d_flags);
        h->launches += 1;
    }
    check_index_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(dn, n, d.rows, h->d_flags);
    check_index_kernel<<<(unsigned)((n * T + 255) / 256), 256, 0, h->stream>>>(ds, n * T, d.rows, h->d_flags);
    h->launches += 2;
    int32_t flag = 0;
    DMG_CUDA(h, cudaMemcpyAsync(&flag, h->d_flags, 4, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    if (flag) {
        DMG_CUDA(h, cudaMemsetAsync(h->d_flags, 0, 4, h->stream));
        return fail(h, DMG_ERR_INDEX, "training batch: embeddingLookup failed, index outside [0, %lld) or bad mask position", (long long)d.rows);
    }
    DMG_TRY(grad_enqueue<real>(h, n, dn, ds, d_mask, dl, d_loss));
    double loss_sum = 0.0;
    DMG_CUDA(h, cudaMemcpyAsync(&loss_sum, d_loss, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    *loss_out = loss_sum / (double)n;
    return DMG_OK;
}

template <typename real> int32_t adam_pass(dmg_handle_t h, double lr, int step_t)
{
    DinDev &d = h->din;
    const double beta1 = 0.9, beta2 = 0.999, eps = 1e-8;                      // Adam.scala:10-14
    const double step = lr * std::sqrt(1 - std::pow(beta2, step_t)) / (1 - std::pow(beta1, step_t));
    adam_dense_kernel<real><<<h->sm_count * 16, 256, 0, h->stream>>>(
        (real *)d.d_params, (real *)d.d_grad, (real *)d.d_m, (real *)d.d_v, d.n_params, (real)beta1, (real)(1 - beta1),
        (real)beta2, (real)(1 - beta2), (real)eps, (real)(-step));
    h->launches += 1;
    DMG_CUDA(h, cudaGetLastError());
    if (d.kind == 1) { h->fast_dirty = true; return DMG_OK; }   // DeepFM keeps no transposed copies
    return dmg_refresh_transposes(h);
}

int32_t train_precheck(dmg_handle_t h, int64_t rows, const void *node, const void *seq, const void *labels, const void *out)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    DMG_TRY(model_is_shared(h, "training"));
    if (!h->din.loaded) return fail(h, DMG_ERR_STATE, "DIN weights must be loaded first");
    // a Float DeepFM model carries the `sharded` mark as "level-synchronous path only"; its table is whole unless a communicator says otherwise
    const bool whole_deepfm = h->din.kind == 1 && h->din.dtype == DMG_F32 && (!h->shard || h->shard->world == 1);
    if (h->din.sharded && !whole_deepfm) return fail(h, DMG_ERR_STATE, "the node table is sharded (dmg_shard_init): use the dmg_shard_* entry points");
    if (rows <= 0 || !node || !seq || !labels || !out) return fail(h, DMG_ERR_INVALID_ARG, "bad arguments");
    DMG_CUDA(h, cudaSetDevice(h->device));
    return DMG_OK;
}

}  // namespace

DMG_API int32_t dmg_din_gradients(dmg_handle_t h, int64_t rows, const int32_t *node, const int32_t *seq, const int32_t *mask_flat,
                                  int64_t n_mask, const void *labels, void *out_loss, void *out_grad, int64_t n_grad)
{
    DMG_TRY(train_precheck(h, rows, node, seq, labels, out_loss));
    DinDev &d = h->din;
    if (!out_grad || n_grad != d.n_params) return fail(h, DMG_ERR_INVALID_ARG, "out_grad must hold %lld values", (long long)d.n_params);
    double loss = 0.0;
    if (d.dtype == DMG_F32) {
        DMG_TRY(ensure_train_state<float>(h));
        DMG_CUDA(h, cudaMemsetAsync(d.d_grad, 0, (size_t)d.n_params * 4, h->stream));
        DMG_TRY(grad_pass<float>(h, rows, node, seq, mask_flat, n_mask, (const float *)labels, &loss));
        *(float *)out_loss = (float)loss;
    } else {
        DMG_TRY(ensure_train_state<double>(h));
        DMG_CUDA(h, cudaMemsetAsync(d.d_grad, 0, (size_t)d.n_params * 8, h->stream));
        DMG_TRY(grad_pass<double>(h, rows, node, seq, mask_flat, n_mask, (const double *)labels, &loss));
        *(double *)out_loss = loss;
    }
    DMG_CUDA(h, cudaMemcpyAsync(out_grad, d.d_grad, (size_t)d.n_params * d.esz, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaMemsetAsync(d.d_grad, 0, (size_t)d.n_params * d.esz, h->stream));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    return DMG_OK;
}

DMG_API int32_t dmg_train_step(dmg_handle_t h, int64_t rows, const int32_t *node, const int32_t *seq, const int32_t *mask_flat,
                               int64_t n_mask, const void *labels, double lr, int32_t step_t, void *out_loss)
{
    DMG_TRY(train_precheck(h, rows, node, seq, labels, out_loss));
    if (step_t < 1) return fail(h, DMG_ERR_INVALID_ARG, "step_t is the 1-based Adam timestep");
    double loss = 0.0;
    if (h->din.dtype == DMG_F32) {
        DMG_TRY(grad_pass<float>(h, rows, node, seq, mask_flat, n_mask, (const float *)labels, &loss));
        DMG_TRY(adam_pass<float>(h, lr, step_t));
        *(float *)out_loss = (float)loss;
    } else {
        DMG_TRY(grad_pass<double>(h, rows, node, seq, mask_flat, n_mask, (const double *)labels, &loss));
        DMG_TRY(adam_pass<double>(h, lr, step_t));
        *(double *)out_loss = loss;
    }
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    return DMG_OK;
}

// ---- device-buffer variant: nothing is copied, nothing is synchronised ----------------------------------------------------------
namespace {
// embeddingLookup's range check without a host round trip: a bad index raises the handle's flag (reported by dmg_synchronize) and is
// replaced by the padding index so that the kernels behind it stay inside the table
__global__ void sanitize_index_kernel(const int32_t *__restrict__ src, int32_t *__restrict__ dst, int64_t n, int64_t rows, int32_t *__restrict__ flag)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t c = src[i];
    if (c < -1 || (int64_t)c >= rows) { atomicExch(flag, 1); c = -1; }
    dst[i] = c;
}
template <typename real> __global__ void loss_mean_kernel(const double *__restrict__ sum, int64_t n, real *__restrict__ out) { *out = (real)(*sum / (double)n); }
}  // namespace

int32_t dmg_sanitize_indices(dmg_handle_t h, const int32_t *src, int32_t *dst, int64_t n, int64_t rows)
{
    if (n > 0) sanitize_index_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(src, dst, n, rows, h->d_flags);
    h->launches += 1;
    DMG_CUDA(h, cudaGetLastError());
    return DMG_OK;
}

DMG_API int32_t dmg_train_step_dev(dmg_handle_t h, int64_t rows, const int32_t *d_node, const int32_t *d_seq, const uint8_t *d_mask,
                                   const void *d_labels, double lr, int32_t step_t, void *d_out_loss)
{
    DMG_TRY(train_precheck(h, rows, d_node, d_seq, d_labels, d_out_loss));
    if (step_t < 1) return fail(h, DMG_ERR_INVALID_ARG, "step_t is the 1-based Adam timestep");
    DinDev &d = h->din;
    const int T = d.T;
    const bool f32 = d.dtype == DMG_F32;
    if (f32) DMG_TRY(ensure_train_state<float>(h)); else DMG_TRY(ensure_train_state<double>(h));
    DMG_TRY(ensure_dev(h, h->s_work, Carver::need({(size_t)rows * 4, (size_t)rows * T * 4, (size_t)rows * T, 64})));
    Carver cw(h->s_work.d);
    int32_t *dn = cw.take<int32_t>((size_t)rows), *ds = cw.take<int32_t>((size_t)rows * T);
    uint8_t *zmask = cw.take<uint8_t>((size_t)rows * T);
    double *d_loss = cw.take<double>(1);
    DMG_TRY(dmg_sanitize_indices(h, d_node, dn, rows, d.rows));
    DMG_TRY(dmg_sanitize_indices(h, d_seq, ds, rows * T, d.rows));
    if (!d_mask) DMG_CUDA(h, cudaMemsetAsync(zmask, 0, (size_t)rows * T, h->stream));
    DMG_CUDA(h, cudaMemsetAsync(d_loss, 0, sizeof(double), h->stream));
    if (f32) {
        DMG_TRY(grad_enqueue<float>(h, rows, dn, ds, d_mask ? d_mask : zmask, (const float *)d_labels, d_loss));
        DMG_TRY(adam_pass<float>(h, lr, step_t));
        loss_mean_kernel<float><<<1, 1, 0, h->stream>>>(d_loss, rows, (float *)d_out_loss);
    } else {
        DMG_TRY(grad_enqueue<double>(h, rows, dn, ds, d_mask ? d_mask : zmask, (const double *)d_labels, d_loss));
        DMG_TRY(adam_pass<double>(h, lr, step_t));
        loss_mean_kernel<double><<<1, 1, 0, h->stream>>>(d_loss, rows, (double *)d_out_loss);
    }
    h->launches += 1;
    DMG_CUDA(h, cudaGetLastError());
    return DMG_OK;
}

DMG_API int32_t dmg_tdm_sample_expand(dmg_handle_t h, int32_t n_targets, const int32_t *target_items, const int32_t *item_seq,
                                      const int32_t *layer_neg, int32_t start_level, int32_t with_prob, int32_t tolerance,
                                      uint64_t seed, int32_t *out_node, int32_t *out_seq, float *out_label, int32_t *out_rows)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    const TreeDev &t = h->tree;
    if (!t.loaded || t.complete || !t.d_id_code) return fail(h, DMG_ERR_STATE, "needs a tree loaded with dmg_load_tree_tdm");
    if (!h->din.loaded) return fail(h, DMG_ERR_STATE, "DIN weights (seq_len) must be loaded first");
    if (n_targets <= 0 || !target_items || !item_seq || !layer_neg || !out_node || !out_seq || !out_label || !out_rows)
        return fail(h, DMG_ERR_INVALID_ARG, "bad arguments");
    if (start_level < 1 || start_level > t.max_level)
        return fail(h, DMG_ERR_INVALID_ARG, "start sample level should be at least 1, got %d", start_level);   // NegativeSampler.scala:23
    if (with_prob && !t.d_cdf) return fail(h, DMG_ERR_STATE, "withProb sampling needs node probabilities (dmg_load_tree_tdm prob)");
    if (tolerance < 0) return fail(h, DMG_ERR_INVALID_ARG, "tolerance must be >= 0");
    const int L = t.max_level, T = h->din.T;
    std::vector<int32_t> level_off(L + 2, 0);
    int layer_sum = 0;
    for (int l = start_level; l <= L; l++) {
        if (layer_neg[l] < 0 || (double)layer_neg[l] >= std::pow(2.0, l))
            return fail(h, DMG_ERR_INVALID_ARG, "Num of negative samples must not exceed max numbers in current layer");  // :49-53
        level_off[l] = layer_sum;
        layer_sum += 1 + layer_neg[l];
    }
    DMG_CUDA(h, cudaSetDevice(h->device));
    const size_t rows = (size_t)n_targets * layer_sum;
    const size_t in_bytes = Carver::need({(size_t)n_targets * 4, (size_t)n_targets * T * 4, (size_t)(L + 1) * 4, (size_t)(L + 2) * 4});
    DMG_TRY(ensure_host(h, h->s_in, in_bytes));
    DMG_TRY(ensure_dev(h, h->s_in, in_bytes));
    Carver ch(h->s_in.h), cd(h->s_in.d);
    int32_t *ht = ch.take<int32_t>(n_targets), *dt = cd.take<int32_t>(n_targets);
    int32_t *hs = ch.take<int32_t>((size_t)n_targets * T), *dsq = cd.take<int32_t>((size_t)n_targets * T);
    int32_t *hneg = ch.take<int32_t>(L + 1), *dneg = cd.take<int32_t>(L + 1);
    int32_t *hoff = ch.take<int32_t>(L + 2), *doff = cd.take<int32_t>(L + 2);
    memcpy(ht, target_items, (size_t)n_targets * 4); memcpy(hs, item_seq, (size_t)n_targets * T * 4);
    memcpy(hneg, layer_neg, (size_t)(L + 1) * 4); memcpy(hoff, level_off.data(), (size_t)(L + 2) * 4);
    DMG_CUDA(h, cudaMemcpyAsync(h->s_in.d, h->s_in.h, ch.off, cudaMemcpyHostToDevice, h->stream));
    const size_t out_bytes = Carver::need({rows * 4, rows * T * 4, rows * 4});
    DMG_TRY(ensure_dev(h, h->s_out, out_bytes));
    Carver od(h->s_out.d);
    int32_t *d_node = od.take<int32_t>(rows), *d_seq = od.take<int32_t>(rows * T);
    float *d_lab = od.take<float>(rows);
    DMG_TRY(ensure_dev(h, h->s_work, Carver::need({(size_t)n_targets * T * 4, (size_t)n_targets * T})));
    Carver cw(h->s_work.d);
    int32_t *d_codes = cw.take<int32_t>((size_t)n_targets * T);
    uint8_t *d_mask = cw.take<uint8_t>((size_t)n_targets * T);
    const int n_lv = L + 1 - start_level;
    tdm_sample_kernel<<<(n_targets * n_lv + 127) / 128, 128, 0, h->stream>>>(n_targets, dt, t.d_id_code, t.non_leaf_offset, t.d_exists, L,
                                                                           dneg, doff, start_level, layer_sum, seed, d_node, d_lab, h->d_flags,
                                                                           with_prob ? t.d_cdf : nullptr, tolerance);
    const int64_t nseq = (int64_t)n_targets * T;
    // history ids -> codes (TDMTree.idToCode) then repeated layer_sum times
    DMG_TRY(dmg_tdm_ids_to_codes(h, dsq, nseq, 1, d_codes, d_mask));
    repeat_seq_kernel<<<(unsigned)((rows * T + 255) / 256), 256, 0, h->stream>>>(d_codes, d_mask, n_targets, T, layer_sum, d_seq, nullptr);
    h->launches += 2;
    DMG_CUDA(h, cudaGetLastError());
    int32_t flag = 0;
    DMG_CUDA(h, cudaMemcpyAsync(&flag, h->d_flags, 4, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaMemcpyAsync(out_node, d_node, rows * 4, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaMemcpyAsync(out_seq, d_seq, rows * T * 4, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaMemcpyAsync(out_label, d_lab, rows * 4, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    if (flag) {
        DMG_CUDA(h, cudaMemsetAsync(h->d_flags, 0, 4, h->stream));
        return fail(h, DMG_ERR_INDEX, "dmg_tdm_sample_expand: a target is not a leaf item of the tree, or a history id is invalid");
    }
    *out_rows = (int32_t)rows;
    return DMG_OK;
}

// ---- K8: OTM pseudo targets, bottom up, on the device ---------------------------------------------------------------------------
// OTMTree.optimalPseudoTargets / computeTargets / computeChildrenScores (otm/.../tree/OTMTree.scala:27-46, 104-172).  Per level:
// otm_pt_expand_kernel lists (node, sibling, history) rows for every node of every user's list, the fp64 row scorer (model.forward)
// scores the nodes and the siblings, otm_pt_combine_kernel turns them into the parents' clipped targets.  Lists are kept sorted by
// node id in [B][M] slots (-1 padding): the reference's Lists come out of Maps and are only ever looked up by id.
namespace {
__global__ void otm_pt_expand_kernel(int B, int M, int T, const int32_t *__restrict__ ids, const int32_t *__restrict__ cnt,
                                     const int32_t *__restrict__ seqs, int use_mask, int32_t *__restrict__ pos, int32_t *__restrict__ neg,
                                     int32_t *__restrict__ rseq, uint8_t *__restrict__ rmask)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * M) return;
    const int u = i / M, k = i % M;
    const bool live = k < cnt[u];
    const int32_t id = live ? ids[i] : -1;
    pos[i] = id;
    neg[i] = live ? ((id % 2 == 0) ? id - 1 : id + 1) : -1;          // OTMTree.scala:143
    for (int j = 0; j < T; j++) {
        const int32_t c = live ? seqs[(size_t)u * T + j] : -1;
        rseq[(size_t)i * T + j] = c;
        rmask[(size_t)i * T + j] = (use_mask && c == -1) ? 1 : 0;
    }
}
__global__ void otm_pt_combine_kernel(int B, int M, const int32_t *__restrict__ cid, const double *__restrict__ cval,
                                      const int32_t *__restrict__ ccnt, const double *__restrict__ ppos, const double *__restrict__ pneg,
                                      int32_t *__restrict__ pid, double *__restrict__ pval, int32_t *__restrict__ pcnt)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= B) return;
    const int n = ccnt[u];
    int c = 0;
    for (int k = 0; k < M; k++) { pid[(size_t)u * M + k] = -1; pval[(size_t)u * M + k] = 0.0; }
    for (int k = 0; k < n; k++) {
        const int32_t id = cid[(size_t)u * M + k], sib = (id % 2 == 0) ? id - 1 : id + 1;
        double neg_label = 0.0;                                   // nodes.find(_.id == nn): the sibling's score if it is listed (:146-150)
        for (int q = 0; q < n; q++) if (cid[(size_t)u * M + q] == sib) { neg_label = cval[(size_t)u * M + q]; break; }
        const double label = ppos[(size_t)u * M + k] >= pneg[(size_t)u * M + k] ? cval[(size_t)u * M + k] : neg_label;   // :117-118
        const int32_t par = (id - 1) >> 1;
        int j = 0;
        while (j < c && pid[(size_t)u * M + j] != par) j++;
        if (j == c) { pid[(size_t)u * M + c] = par; c++; }            // children are id-sorted, so parents arrive in ascending order
        pval[(size_t)u * M + j] = __dadd_rn(pval[(size_t)u * M + j], label);                                             // groupMapReduce(_ + _)
    }
    for (int j = 0; j < c; j++) {
        const double v = pval[(size_t)u * M + j];
        pval[(size_t)u * M + j] = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);                                                   // clipValue(_, 0, 1)
    }
    pcnt[u] = c;
}
}  // namespace

DMG_API int32_t dmg_otm_pseudo_targets(dmg_handle_t h, int32_t B, const int32_t *leaf_seq, const int64_t *target_off,
                                       const int32_t *target_leaves, int32_t start_level, int32_t use_mask, int32_t M,
                                       int32_t *out_ids, double *out_vals, int32_t *out_counts)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    DinDev &d = h->din;
    const TreeDev &t = h->tree;
    if (!t.loaded || !t.complete || !d.loaded) return fail(h, DMG_ERR_STATE, "needs a complete tree (dmg_load_tree_complete) and DIN weights");
    if (d.dtype != DMG_F64 || d.kind != 0) return fail(h, DMG_ERR_STATE, "OTM scorer is DeepModel[Double] DIN: load DMG_F64 DIN weights");
    if (d.sharded) return fail(h, DMG_ERR_STATE, "the node table is sharded: not supported here");
    const int L = t.max_level, n_lvl = L - start_level;
    if (B <= 0 || M <= 0 || !leaf_seq || !target_off || !target_leaves || !out_ids || !out_vals || !out_counts || start_level < 0 || n_lvl <= 0)
        return fail(h, DMG_ERR_INVALID_ARG, "bad arguments");
    DMG_CUDA(h, cudaSetDevice(h->device));
    const int T = d.T, E = d.E;
    const size_t BM = (size_t)B * M;
    // leaf lists on the host: Node(target, 1.0), sorted by id, duplicates collapsed
    std::vector<int32_t> ids0(BM, -1), cnt0(B, 0);
    std::vector<double> val0(BM, 0.0);
    const int64_t leaf_start = ((int64_t)1 << L) - 1, leaf_end = ((int64_t)2 << L) - 1;
    for (int u = 0; u < B; u++) {
        std::vector<int32_t> v(target_leaves + target_off[u], target_leaves + target_off[u + 1]);
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
        if ((int)v.size() > M) return fail(h, DMG_ERR_INVALID_ARG, "user %d has %d distinct targets, M = %d", u, (int)v.size(), M);
        for (size_t k = 0; k < v.size(); k++) {
            if (v[k] < leaf_start || v[k] >= leaf_end) return fail(h, DMG_ERR_INDEX, "target %d of user %d is not a leaf node id", v[k], u);
            ids0[(size_t)u * M + k] = v[k]; val0[(size_t)u * M + k] = 1.0;
        }
        cnt0[u] = (int32_t)v.size();
    }
    const size_t need = Carver::need({(size_t)B * T * 4, (size_t)n_lvl * BM * 4, (size_t)n_lvl * BM * 8, (size_t)n_lvl * B * 4, BM * 4, BM * 4,
                                      BM * T * 4, BM * T, BM * 8, BM * 8});
    DMG_TRY(ensure_dev(h, h->s_work, need));
    Carver cw(h->s_work.d);
    int32_t *d_seq = cw.take<int32_t>((size_t)B * T);
    int32_t *d_ids = cw.take<int32_t>((size_t)n_lvl * BM);
    double *d_vals = cw.take<double>((size_t)n_lvl * BM);
    int32_t *d_cnt = cw.take<int32_t>((size_t)n_lvl * B);
    int32_t *d_pos = cw.take<int32_t>(BM), *d_neg = cw.take<int32_t>(BM), *d_rseq = cw.take<int32_t>(BM * T);
    uint8_t *d_rmask = cw.take<uint8_t>(BM * T);
    double *d_pp = cw.take<double>(BM), *d_pn = cw.take<double>(BM);
    DMG_CUDA(h, cudaMemcpyAsync(d_seq, leaf_seq, (size_t)B * T * 4, cudaMemcpyHostToDevice, h->stream));
    DMG_CUDA(h, cudaMemcpyAsync(d_ids + (size_t)(n_lvl - 1) * BM, ids0.data(), BM * 4, cudaMemcpyHostToDevice, h->stream));
    DMG_CUDA(h, cudaMemcpyAsync(d_vals + (size_t)(n_lvl - 1) * BM, val0.data(), BM * 8, cudaMemcpyHostToDevice, h->stream));
    DMG_CUDA(h, cudaMemcpyAsync(d_cnt + (size_t)(n_lvl - 1) * B, cnt0.data(), (size_t)B * 4, cudaMemcpyHostToDevice, h->stream));
    check_index_kernel<<<(unsigned)(((size_t)B * T + 255) / 256), 256, 0, h->stream>>>(d_seq, (int64_t)B * T, d.rows, h->d_flags);
    h->launches += 1;
    const double scale = 1.0 / std::sqrt((double)E);
    for (int li = n_lvl - 1; li > 0; li--) {
        const int32_t *cid = d_ids + (size_t)li * BM, *ccnt = d_cnt + (size_t)li * B;
        const double *cval = d_vals + (size_t)li * BM;
        otm_pt_expand_kernel<<<(unsigned)((BM + 255) / 256), 256, 0, h->stream>>>(B, M, T, cid, ccnt, d_seq, use_mask, d_pos, d_neg, d_rseq, d_rmask);
        for (int pass = 0; pass < 2; pass++) {
            // QUIRK kept: without a mask input the reference scores the NEGATIVE tensor for both predictions (OTMTree.scala:157-161)
            const int32_t *nodes = (pass == 0 && use_mask) ? d_pos : d_neg;
            double *out = pass == 0 ? d_pp : d_pn;
            cudaError_t terr = cudaSuccess;
            if (!rows_forward_tiled<double>(E, d.emb<double>(), (const double *)d.d_wattT, (const double *)d.d_w1T, d.b1<double>(), d.w2<double>(),
                                            d.b2<double>(), scale, T, (int64_t)BM, nodes, d_rseq, d_rmask, out, h->sm_count, h->smem_per_sm,
                                            h->smem_optin, h->stream, &terr)) {
                const size_t smem = (size_t)kRowsRB * ((size_t)4 * E + (size_t)T * E + T + 1) * 8;
                auto kern = din_rows_forward_kernel<double>;
                DMG_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                kern<<<(int)std::min<int64_t>(((int64_t)BM + kRowsRB - 1) / kRowsRB, (int64_t)h->sm_count * 8), kRowsThreads, smem, h->stream>>>(
                    d.emb<double>(), (const double *)d.d_wattT, (const double *)d.d_w1T, d.b1<double>(), d.w2<double>(), d.b2<double>(), scale, E, T,
                    (int64_t)BM, nodes, d_rseq, d_rmask, out);
            }
            DMG_CUDA(h, terr);
        }
        otm_pt_combine_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(B, M, cid, cval, ccnt, d_pp, d_pn, d_ids + (size_t)(li - 1) * BM,
                                                                     d_vals + (size_t)(li - 1) * BM, d_cnt + (size_t)(li - 1) * B);
        h->launches += 4;
    }
    DMG_CUDA(h, cudaGetLastError());
    int32_t flag = 0;
    DMG_CUDA(h, cudaMemcpyAsync(&flag, h->d_flags, 4, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaMemcpyAsync(out_ids, d_ids, (size_t)n_lvl * BM * 4, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaMemcpyAsync(out_vals, d_vals, (size_t)n_lvl * BM * 8, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaMemcpyAsync(out_counts, d_cnt, (size_t)n_lvl * B * 4, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    if (flag) {
        DMG_CUDA(h, cudaMemsetAsync(h->d_flags, 0, 4, h->stream));
        return fail(h, DMG_ERR_INDEX, "dmg_otm_pseudo_targets: embeddingLookup failed, history id outside [-1, %lld)", (long long)d.rows);
    }
    return DMG_OK;
}

DMG_API int32_t dmg_jtm_item_weights(dmg_handle_t h, int32_t n_items, const int64_t *sample_off, const int32_t *sample_seq,
                                     const int32_t *parent_code, int32_t old_level, int32_t level, int32_t hierarchical,
                                     int32_t min_level, int32_t use_mask, float *out_weights)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    const TreeDev &t = h->tree;
    DinDev &d = h->din;
    if (!t.loaded || t.complete || !t.d_id_code) return fail(h, DMG_ERR_STATE, "needs a tree loaded with dmg_load_tree_tdm");
    if (!d.loaded || d.dtype != DMG_F32) return fail(h, DMG_ERR_STATE, "JTM scorer is Module[Float]: load DMG_F32 weights");
    if (h->din.sharded) return fail(h, DMG_ERR_STATE, "the node table is sharded (dmg_shard_init): use the dmg_shard_* entry points");
    const int gap = level - old_level;
    if (n_items <= 0 || !sample_off || !parent_code || !out_weights || gap < 1 || gap > 8 || old_level < 0 || level > t.max_level)
        return fail(h, DMG_ERR_INVALID_ARG, "bad arguments (1 <= level - old_level <= 8, level <= max_level)");
    const int64_t n_samples = sample_off[n_items];
    if (n_samples > 0 && !sample_seq) return fail(h, DMG_ERR_INVALID_ARG, "sample_seq is null");
    DMG_CUDA(h, cudaSetDevice(h->device));
    const int T = d.T, E = d.E;
    const int n_nodes = (1 << (gap + 1)) - 2, n_child = 1 << gap;
    const int64_t n_rows = n_samples * n_nodes;
    // host: row -> item map
    std::vector<int64_t> item_of_row((size_t)std::max<int64_t>(n_rows, 1));
    for (int i = 0; i < n_items; i++) {
        const int64_t b = sample_off[i] * n_nodes, e = sample_off[i + 1] * n_nodes;
        for (int64_t r = b; r < e; r++) item_of_row[(size_t)r] = i;
    }
    const size_t in_bytes = Carver::need({(size_t)(n_items + 1) * 8, (size_t)n_samples * T * 4, (size_t)n_items * 4, (size_t)n_rows * 8});
    DMG_TRY(ensure_host(h, h->s_in, in_bytes));
    DMG_TRY(ensure_dev(h, h->s_in, in_bytes));
    Carver ch(h->s_in.h), cd(h->s_in.d);
    int64_t *ho = ch.take<int64_t>(n_items + 1), *dof = cd.take<int64_t>(n_items + 1);
    int32_t *hs = ch.take<int32_t>((size_t)n_samples * T), *dsq = cd.take<int32_t>((size_t)n_samples * T);
    int32_t *hp = ch.take<int32_t>(n_items), *dp = cd.take<int32_t>(n_items);
    int64_t *hr = ch.take<int64_t>((size_t)n_rows), *dr = cd.take<int64_t>((size_t)n_rows);
    memcpy(ho, sample_off, (size_t)(n_items + 1) * 8);
    if (n_samples) memcpy(hs, sample_seq, (size_t)n_samples * T * 4);
    memcpy(hp, parent_code, (size_t)n_items * 4);
    if (n_rows) memcpy(hr, item_of_row.data(), (size_t)n_rows * 8);
    DMG_CUDA(h, cudaMemcpyAsync(h->s_in.d, h->s_in.h, ch.off, cudaMemcpyHostToDevice, h->stream));
    const size_t work = Carver::need({(size_t)n_rows * 4, (size_t)n_rows * T * 4, (size_t)n_rows * T, (size_t)n_rows * 4,
                                      (size_t)n_items * n_nodes * 4, (size_t)n_items * n_child * 4});
    DMG_TRY(ensure_dev(h, h->s_work, work));
    Carver cw(h->s_work.d);
    int32_t *d_node = cw.take<int32_t>((size_t)n_rows), *d_seq = cw.take<int32_t>((size_t)n_rows * T);
    uint8_t *d_mask = cw.take<uint8_t>((size_t)n_rows * T);
    float *d_logit = cw.take<float>((size_t)n_rows), *d_nsum = cw.take<float>((size_t)n_items * n_nodes);
    float *d_w = cw.take<float>((size_t)n_items * n_child);
    const int grid = h->sm_count * 8;
    if (n_rows) {
        jtm_rows_kernel<<<grid, 256, 0, h->stream>>>(n_items, dof, dsq, dp, old_level, gap, n_nodes, T, t.d_id_code, t.non_leaf_offset,
                                                    t.max_code, t.max_level, hierarchical, min_level, use_mask, dr, n_rows, d_node, d_seq, d_mask);
        check_index_kernel<<<(unsigned)((n_rows * T + 255) / 256), 256, 0, h->stream>>>(d_seq, n_rows * T, d.rows, h->d_flags);
        cudaError_t terr = cudaSuccess;
        if (!rows_forward_tiled<float>(E, d.emb<float>(), (const float *)d.d_wattT, (const float *)d.d_w1T, d.b1<float>(), d.w2<float>(),
                                       d.b2<float>(), (float)(1.0 / std::sqrt((double)E)), T, n_rows, d_node, d_seq, d_mask, d_logit,
                                       h->sm_count, h->smem_per_sm, h->smem_optin, h->stream, &terr)) {
            const size_t smem = (size_t)kRowsRB * ((size_t)4 * E + (size_t)T * E + T + 1) * 4;
            auto kern = din_rows_forward_kernel<float>;
            DMG_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const int g2 = (int)std::min<int64_t>((n_rows + kRowsRB - 1) / kRowsRB, (int64_t)h->sm_count * 8);
            kern<<<g2, kRowsThreads, smem, h->stream>>>(d.emb<float>(), (const float *)d.d_wattT, (const float *)d.d_w1T, d.b1<float>(),
                                                        d.w2<float>(), d.b2<float>(), (float)(1.0 / std::sqrt((double)E)), E, T, n_rows,
                                                        d_node, d_seq, d_mask, d_logit);
        }
        DMG_CUDA(h, terr);
        h->launches += 3;
    }
    jtm_reduce_kernel<<<grid, 256, 0, h->stream>>>(n_items, dof, d_logit, n_nodes, gap, d_nsum, d_w);
    jtm_path_kernel<<<grid, 256, 0, h->stream>>>(n_items, dof, gap, n_nodes, d_nsum, d_w);
    h->launches += 2;
    DMG_CUDA(h, cudaGetLastError());
    int32_t flag = 0;
    DMG_CUDA(h, cudaMemcpyAsync(&flag, h->d_flags, 4, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaMemcpyAsync(out_weights, d_w, (size_t)n_items * n_child * 4, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    if (flag) {
        DMG_CUDA(h, cudaMemsetAsync(h->d_flags, 0, 4, h->stream));
        return fail(h, DMG_ERR_INDEX, "dmg_jtm_item_weights: a history id maps outside the node table");
    }
    return DMG_OK;
}

// ---- JTM: per-level item -> child assignment (K10), host code like the reference's ---------------------------------
// TreeLearning.getChildrenProjection / sortNodeWeights / reBalance (jtm/.../optim/TreeLearning.scala:48-97,137-150,
// 217-265) for ONE level step over all parents.  Sequential greedy per parent by construction (the heaviest child is
// frozen first, its overflow moves to the next-best unfrozen child), so it runs on the host; the scorer work that feeds
// it is dmg_jtm_item_weights.  Items of one parent are taken in array order (the reference's Map order is a JVM
// artefact the caller fixes by the order of its arrays).
//   parent_code[i]  node of level old_level the item sits under
//   old_child[i]    JTMTree.getAncestorAtLevel(item, level) in the CURRENT tree (TreeLearning.scala:224-225: items already
//                   under a child are kept there first)
//   weights         n_items x n_child from dmg_jtm_item_weights (children left to right)
//   out_node[i]     new node of level `level`; the parent code when every candidate child was full (the reference keeps
//                   the old projection for such an item)
namespace {
inline uint32_t jtm_order_key(float w)                                  // Ordering[Float].reverse under Float.compare
{
    uint32_t u;
    memcpy(&u, &w, 4);
    if (w != w) u = 0x7fc00000u;
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
}  // namespace

DMG_API int32_t dmg_jtm_assign_level(dmg_handle_t h, int32_t n_items, const int32_t *parent_code, const int32_t *old_child,
                                     int32_t n_child, const float *weights, int32_t max_assign, int32_t *out_node)
{
    // host-only: h may be NULL (no device is touched), errors are then reported by the status alone
    if (n_items < 0 || n_child <= 0 || max_assign <= 0 || (n_items > 0 && (!parent_code || !old_child || !weights || !out_node)))
        return fail(h, DMG_ERR_INVALID_ARG, "dmg_jtm_assign_level: bad arguments");
    struct Entry { int32_t item; float w; int32_t next; };
    // items grouped by parent, array order kept
    std::vector<int32_t> order(n_items);
    for (int32_t i = 0; i < n_items; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return parent_code[a] < parent_code[b]; });
    std::vector<int32_t> cand((size_t)n_child);
    std::vector<std::vector<int32_t>> cand_of;                      // per item of the group: children by weight desc (stable)
    for (int32_t g0 = 0; g0 < n_items;) {
        int32_t g1 = g0;
        while (g1 < n_items && parent_code[order[g1]] == parent_code[order[g0]]) g1++;
        const int32_t par = parent_code[order[g0]], m = g1 - g0;
        const int64_t first = ((int64_t)par + 1) * n_child - 1;       // getChildrenAtLevel: left to right
        cand_of.assign((size_t)m, std::vector<int32_t>());
        std::vector<std::vector<Entry>> res((size_t)n_child);
        std::vector<char> processed((size_t)n_child, 0);
        for (int32_t k = 0; k < m; k++) {
            const int32_t it = order[g0 + k];
            const float *w = weights + (size_t)it * n_child;
            for (int32_t c = 0; c < n_child; c++) cand[c] = c;
            std::stable_sort(cand.begin(), cand.end(), [&](int32_t a, int32_t b) { return jtm_order_key(w[a]) > jtm_order_key(w[b]); });
            cand_of[k] = cand;
            res[cand[0]].push_back({k, w[cand[0]], 1});
            out_node[it] = par;
        }
        for (;;) {
            int32_t best = -1, best_cnt = -1;
            for (int32_t c = 0; c < n_child; c++) {                   // getMaxNode: first maximum wins
                const int32_t cnt = (!processed[c] && !res[c].empty()) ? (int32_t)res[c].size() : -1;
                if (cnt > best_cnt) { best_cnt = cnt; best = c; }
            }
            if (best_cnt <= max_assign) break;
            processed[best] = 1;
            std::vector<Entry> &lst = res[best];
            // sortBy(i => (oldItemNodeMap(i.id) != node, i.weight)) under (Boolean asc, Float desc), stable
            std::stable_sort(lst.begin(), lst.end(), [&](const Entry &a, const Entry &b) { return jtm_order_key(a.w) > jtm_order_key(b.w); });
            std::stable_sort(lst.begin(), lst.end(), [&](const Entry &a, const Entry &b) {
                const bool ma = old_child[order[g0 + a.item]] != first + best, mb = old_child[order[g0 + b.item]] != first + best;
                return !ma && mb;
            });
            std::vector<Entry> rest(lst.begin() + max_assign, lst.end());
            lst.resize((size_t)max_assign);
            for (const Entry &e : rest) {
                const float *w = weights + (size_t)order[g0 + e.item] * n_child;
                for (int32_t k = e.next; k < n_child; k++) {
                    const int32_t c = cand_of[e.item][k];
                    if (!processed[c]) { res[c].push_back({e.item, w[c], k + 1}); break; }
                }
            }
        }
        for (int32_t c = 0; c < n_child; c++)
            for (const Entry &e : res[c]) out_node[order[g0 + e.item]] = (int32_t)(first + c);
        g0 = g1;
    }
    return DMG_OK;
}

// ---- evaluation metrics ----------------------------------------------------------------------------
// Metrics.computeMetrics (tdm/.../evaluation/Metrics.scala:5-25) for a batch of users: precision = hits / k, recall =
// hits / |labels|, NDCG = dcg / idcg with gains log(2) / log(rank + 2), k = number of items actually recommended;
// (0, 0, 0) without a hit.  One thread per user; out[u] = (precision, recall, ndcg) in double, to be summed by the caller
// in user order like EvalResult.+= (Evaluator.scala:62-66).
namespace {
__global__ void eval_metrics_kernel(int B, int stride, const int32_t *__restrict__ rec, const int32_t *__restrict__ rec_count,
                                    const int64_t *__restrict__ label_off, const int32_t *__restrict__ labels, double *__restrict__ out)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= B) return;
    const int k = rec_count[u];
    const int64_t l0 = label_off[u], l1 = label_off[u + 1];
    int common = 0, j = 0;
    double dcg = 0.0, idcg = 0.0;
    const double ln2 = log(2.0);
    for (int i = 0; i < k; i++) {
        const int32_t it = rec[(size_t)u * stride + i];
        bool hit = false;
        for (int64_t q = l0; q < l1 && !hit; q++) hit = labels[q] == it;
        if (hit) {
            common++;
            dcg += ln2 / log((double)(i + 2));
            idcg += ln2 / log((double)(j + 2));
            j++;
        }
    }
    // labels.toSet: the recall denominator is labels.length (duplicates counted), Metrics.scala:20
    out[(size_t)u * 3 + 0] = common ? (double)common / (double)k : 0.0;
    out[(size_t)u * 3 + 1] = common ? (double)common / (double)(l1 - l0) : 0.0;
    out[(size_t)u * 3 + 2] = common ? dcg / idcg : 0.0;
}
}  // namespace

DMG_API int32_t dmg_eval_metrics(dmg_handle_t h, int32_t B, int32_t topk, const int32_t *rec_items, const int32_t *rec_counts,
                                 const int64_t *label_off, const int32_t *labels, double *out_metrics)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    if (B <= 0 || topk <= 0 || !rec_items || !rec_counts || !label_off || !out_metrics || (label_off[B] > 0 && !labels))
        return fail(h, DMG_ERR_INVALID_ARG, "dmg_eval_metrics: bad arguments");
    for (int32_t u = 0; u < B; u++)
        if (rec_counts[u] < 0 || rec_counts[u] > topk || label_off[u + 1] < label_off[u])
            return fail(h, DMG_ERR_INVALID_ARG, "dmg_eval_metrics: bad count / offsets for user %d", u);
    DMG_CUDA(h, cudaSetDevice(h->device));
    const size_t nl = (size_t)label_off[B];
    DMG_TRY(ensure_dev(h, h->s_in, Carver::need({(size_t)B * topk * 4, (size_t)B * 4, ((size_t)B + 1) * 8, nl * 4, (size_t)B * 24})));
    Carver cd(h->s_in.d);
    int32_t *d_rec = cd.take<int32_t>((size_t)B * topk), *d_cnt = cd.take<int32_t>((size_t)B);
    int64_t *d_off = cd.take<int64_t>((size_t)B + 1);
    int32_t *d_lab = cd.take<int32_t>(nl);
    double *d_out = cd.take<double>((size_t)B * 3);
    DMG_CUDA(h, cudaMemcpyAsync(d_rec, rec_items, (size_t)B * topk * 4, cudaMemcpyHostToDevice, h->stream));
    DMG_CUDA(h, cudaMemcpyAsync(d_cnt, rec_counts, (size_t)B * 4, cudaMemcpyHostToDevice, h->stream));
    DMG_CUDA(h, cudaMemcpyAsync(d_off, label_off, ((size_t)B + 1) * 8, cudaMemcpyHostToDevice, h->stream));
    if (nl) DMG_CUDA(h, cudaMemcpyAsync(d_lab, labels, nl * 4, cudaMemcpyHostToDevice, h->stream));
    eval_metrics_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(B, topk, d_rec, d_cnt, d_off, d_lab, d_out);
    h->launches += 1;
    DMG_CUDA(h, cudaGetLastError());
    DMG_CUDA(h, cudaMemcpyAsync(out_metrics, d_out, (size_t)B * 24, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    return DMG_OK;
}


// ---- data-parallel training step over the replicas of one box ------------------------------------------------
// LocalOptimizer.trainBatch / syncGradients (tdm/.../optim/LocalOptimizer.scala:139-187) with GPUs in the place of threads:
// every rank holds the whole model (same weights), runs forward / BCE / backward on ITS rows of the mini-batch, the gradient
// replicas are averaged -- one ncclAllReduce(ncclAvg) of the flat gradient over NVLink instead of the per-thread reduce --
// and every rank applies the same dense Adam step, so the replicas stay bit-identical.  Call after dmg_shard_init (which only
// provides the communicator here: the table is loaded whole with dmg_load_din_weights / dmg_init_din_weights).
// out_loss: mean loss of this rank's rows.
DMG_API int32_t dmg_dp_train_step(dmg_handle_t h, int64_t rows, const int32_t *node, const int32_t *seq, const int32_t *mask_flat,
                                  int64_t n_mask, const void *labels, double lr, int32_t step_t, void *out_loss)
{
    DMG_TRY(train_precheck(h, rows, node, seq, labels, out_loss));
    ShardState *s = h->shard;
    if (!s) return fail(h, DMG_ERR_STATE, "call dmg_shard_init first (it provides the NCCL communicator)");
    if (step_t < 1) return fail(h, DMG_ERR_INVALID_ARG, "step_t is the 1-based Adam timestep");
    DinDev &d = h->din;
    double loss = 0.0;
    if (d.dtype == DMG_F32) DMG_TRY(grad_pass<float>(h, rows, node, seq, mask_flat, n_mask, (const float *)labels, &loss));
    else DMG_TRY(grad_pass<double>(h, rows, node, seq, mask_flat, n_mask, (const double *)labels, &loss));
    if (s->world > 1)
        DMG_NCCL(h, g_nccl.AllReduce(d.d_grad, d.d_grad, (size_t)d.n_params, d.dtype == DMG_F32 ? ncclFloat32 : ncclFloat64, ncclAvg, s->comm,
                                     h->stream));
    if (d.dtype == DMG_F32) { DMG_TRY(adam_pass<float>(h, lr, step_t)); *(float *)out_loss = (float)loss; }
    else { DMG_TRY(adam_pass<double>(h, lr, step_t)); *(double *)out_loss = loss; }
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    return DMG_OK;
}

// ---- training on the SHARDED node table -----------------------------------------------------------------------------
// LocalOptimizer.trainBatch / syncGradients (tdm/.../optim/LocalOptimizer.scala:139-187) when no rank holds the whole table: every
// rank brings its rows of the mini-batch with GLOBAL node codes.  Every embedding occurrence (row, slot) is fetched from the row's
// owner (4 B out, E floats back), forward / BCE / backward run locally on the staged rows (the same din_train_kernel, indices =
// occurrence ids, mean over the GLOBAL batch), the per-occurrence gradients travel back and are APPLIED BY THE OWNER (atomic
// scatter-add into its shard of the dense gradient); only the 3 E^2 + 2 E + 1 dense scorer weights (12 417 at E = 64) and the
// 2^bits - 1 replicated top rows are all-reduced.  Every rank then runs the dense Adam over its own shard.
namespace {
__global__ void shtr_owner_kernel(int64_t n_occ, int T, const int32_t *__restrict__ node, const int32_t *__restrict__ seq, ShardGeo g,
                                  int64_t global_rows, int32_t *__restrict__ occ_code, int32_t *__restrict__ counts, int32_t *__restrict__ flag)
{
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_occ) return;
    const int64_t r = o / (T + 1);
    const int slot = (int)(o - r * (T + 1));
    int32_t c = slot == 0 ? node[r] : seq[r * T + slot - 1];
    if (c < -1 || (int64_t)c >= global_rows || (slot == 0 && c < 0)) { atomicExch(flag, 1); c = -1; }
    occ_code[o] = c;
    if (c >= 0) {
        const int ow = shard_owner(g, c);
        if (ow != g.rank) atomicAdd(counts + ow, 1);
    }
}
// remote occurrences appended to the region of their owner: req_code / req_occ at send_off[owner] + cursor
__global__ void shtr_bucket_kernel(int64_t n_occ, const int32_t *__restrict__ occ_code, ShardGeo g, const int32_t *__restrict__ send_off,
                                   int32_t *__restrict__ cursor, int32_t *__restrict__ req_code, int32_t *__restrict__ req_occ)
{
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_occ) return;
    const int32_t c = occ_code[o];
    if (c < 0) return;
    const int ow = shard_owner(g, c);
    if (ow == g.rank) return;
    const int pos = send_off[ow] + atomicAdd(cursor + ow, 1);
    req_code[pos] = c; req_occ[pos] = (int32_t)o;
}
// owner side: rows of the requested codes
__global__ void shtr_gather_kernel(int64_t n, int E, const int32_t *__restrict__ codes, ShardGeo g, const float *__restrict__ emb, float *__restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * E; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = i / E;
        out[i] = emb[shard_local_row(g, codes[q]) * E + (i - q * E)];
    }
}
// staged rows per occurrence: own / replicated rows straight from the local table, remote ones from the replies; remapped indices
__global__ void shtr_stage_local_kernel(int64_t n_occ, int E, const int32_t *__restrict__ occ_code, ShardGeo g, const float *__restrict__ emb,
                                        float *__restrict__ stage)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_occ * E; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t o = i / E;
        const int32_t c = occ_code[o];
        if (c >= 0 && shard_owner(g, c) == g.rank) stage[i] = emb[shard_local_row(g, c) * E + (i - o * E)];
    }
}
__global__ void shtr_stage_remote_kernel(int64_t n, int E, const int32_t *__restrict__ req_occ, const float *__restrict__ rows, float *__restrict__ stage)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * E; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = i / E;
        stage[(int64_t)req_occ[q] * E + (i - q * E)] = rows[i];
    }
}
__global__ void shtr_remap_kernel(int64_t rows, int T, const int32_t *__restrict__ occ_code, int32_t *__restrict__ node2, int32_t *__restrict__ seq2)
{
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= rows * (T + 1)) return;
    const int64_t r = o / (T + 1);
    const int slot = (int)(o - r * (T + 1));
    const int32_t v = occ_code[o] >= 0 ? (int32_t)o : -1;
    if (slot == 0) node2[r] = v; else seq2[r * T + slot - 1] = v;
}
// gradients back: own rows added in place, remote ones packed in request order
__global__ void shtr_grad_local_kernel(int64_t n_occ, int E, const int32_t *__restrict__ occ_code, ShardGeo g, const float *__restrict__ sgrad,
                                       float *__restrict__ g_emb)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_occ * E; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t o = i / E;
        const int32_t c = occ_code[o];
        if (c >= 0 && shard_owner(g, c) == g.rank) { const float v = sgrad[i]; if (v != 0.0f) atomicAdd(g_emb + shard_local_row(g, c) * E + (i - o * E), v); }
    }
}
__global__ void shtr_grad_pack_kernel(int64_t n, int E, const int32_t *__restrict__ req_occ, const float *__restrict__ sgrad, float *__restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * E; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = i / E;
        out[i] = sgrad[(int64_t)req_occ[q] * E + (i - q * E)];
    }
}
__global__ void shtr_grad_apply_kernel(int64_t n, int E, const int32_t *__restrict__ codes, ShardGeo g, const float *__restrict__ in, float *__restrict__ g_emb)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * E; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = i / E;
        const float v = in[i];
        if (v != 0.0f) atomicAdd(g_emb + shard_local_row(g, codes[q]) * E + (i - q * E), v);
    }
}
}  // namespace

DMG_API int32_t dmg_shard_train_step(dmg_handle_t h, int64_t rows, const int32_t *node, const int32_t *seq, const int32_t *mask_flat,
                                     int64_t n_mask, const float *labels, double lr, int32_t step_t, float *out_loss)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    DMG_TRY(model_is_shared(h, "training"));
    ShardState *s = h->shard;
    DinDev &d = h->din;
    if (!s || !d.loaded || d.kind != 0 || d.dtype != DMG_F32) return fail(h, DMG_ERR_STATE, "dmg_shard_init and a sharded Float DIN table first");
    if (rows <= 0 || !node || !seq || !labels || !out_loss || n_mask < 0 || (n_mask > 0 && !mask_flat) || step_t < 1)
        return fail(h, DMG_ERR_INVALID_ARG, "dmg_shard_train_step: bad arguments");
    DMG_CUDA(h, cudaSetDevice(h->device));
    const int G = s->world, E = d.E, T = d.T, me = s->rank;
    const ShardGeo g = s->geo();
    const int64_t n_occ = rows * (T + 1);
    if (n_occ > 0x7fffffff) return fail(h, DMG_ERR_UNSUPPORTED, "batch too large");
    cudaStream_t st = h->stream;
    DMG_TRY(ensure_train_state<float>(h));
    // inputs
    const size_t in_bytes = Carver::need({(size_t)rows * 4, (size_t)rows * T * 4, (size_t)n_mask * 4, (size_t)rows * 4});
    DMG_TRY(ensure_host(h, h->s_in, in_bytes + (size_t)(G + 1) * (G + 1) * 4 + 64));
    DMG_TRY(ensure_dev(h, h->s_in, in_bytes));
    Carver ch(h->s_in.h), cd(h->s_in.d);
    int32_t *hn = ch.take<int32_t>((size_t)rows), *dn = cd.take<int32_t>((size_t)rows);
    int32_t *hs = ch.take<int32_t>((size_t)rows * T), *ds = cd.take<int32_t>((size_t)rows * T);
    int32_t *hm = ch.take<int32_t>((size_t)n_mask), *dm = cd.take<int32_t>((size_t)n_mask);
    float *hl = ch.take<float>((size_t)rows), *dl = cd.take<float>((size_t)rows);
    int32_t *h_matrix = ch.take<int32_t>((size_t)(G + 1) * G);
    memcpy(hn, node, (size_t)rows * 4); memcpy(hs, seq, (size_t)rows * T * 4); memcpy(hl, labels, (size_t)rows * 4);
    if (n_mask) memcpy(hm, mask_flat, (size_t)n_mask * 4);
    DMG_CUDA(h, cudaMemcpyAsync(h->s_in.d, h->s_in.h, in_bytes, cudaMemcpyHostToDevice, st));
    // fixed-size work
    const size_t fixed = Carver::need({(size_t)rows * T, 64, (size_t)n_occ * 4, (size_t)(G + 1) * 4, (size_t)(G + 1) * G * 4, (size_t)G * 4, (size_t)G * 4,
                                       (size_t)n_occ * 4, (size_t)n_occ * 4, (size_t)rows * 4, (size_t)rows * T * 4,
                                       (size_t)n_occ * E * 4, (size_t)n_occ * E * 4, (size_t)n_occ * E * 4});
    DMG_TRY(ensure_dev(h, h->s_work, fixed));
    Carver cw(h->s_work.d);
    uint8_t *d_mask = cw.take<uint8_t>((size_t)rows * T);
    double *d_loss = cw.take<double>(2);
    int32_t *occ_code = cw.take<int32_t>(n_occ), *d_counts = cw.take<int32_t>(G + 1), *d_matrix = cw.take<int32_t>((size_t)(G + 1) * G);
    int32_t *d_send_off = cw.take<int32_t>(G), *d_cursor = cw.take<int32_t>(G);
    int32_t *req_code = cw.take<int32_t>(n_occ), *req_occ = cw.take<int32_t>(n_occ), *node2 = cw.take<int32_t>(rows), *seq2 = cw.take<int32_t>((size_t)rows * T);
    float *stage = cw.take<float>((size_t)n_occ * E), *sgrad = cw.take<float>((size_t)n_occ * E), *xfer = cw.take<float>((size_t)n_occ * E);
    DMG_CUDA(h, cudaMemsetAsync(d_mask, 0, (size_t)rows * T, st));
    DMG_CUDA(h, cudaMemsetAsync(d_loss, 0, 16, st));
    DMG_CUDA(h, cudaMemsetAsync(d_counts, 0, (size_t)(G + 1) * 4, st));
    DMG_CUDA(h, cudaMemsetAsync(d_cursor, 0, (size_t)G * 4, st));
    DMG_CUDA(h, cudaMemsetAsync(sgrad, 0, (size_t)n_occ * E * 4, st));
    if (n_mask) { mask_scatter_kernel<<<(unsigned)((n_mask + 255) / 256), 256, 0, st>>>(dm, n_mask, rows * T, d_mask, h->d_flags); h->launches += 1; }
    const unsigned gocc = (unsigned)((n_occ + 255) / 256);
    shtr_owner_kernel<<<gocc, 256, 0, st>>>(n_occ, T, dn, ds, g, s->global_rows, occ_code, d_counts, h->d_flags);
    const int32_t my_rows = (int32_t)rows;
    DMG_CUDA(h, cudaMemcpyAsync(d_counts + G, &my_rows, 4, cudaMemcpyHostToDevice, st));
    h->launches += 1;
    // count matrix [rank][G + 1]: requests rank -> owner, and the rank's row count
    if (G > 1) DMG_NCCL(h, g_nccl.AllGather(d_counts, d_matrix, (size_t)(G + 1), ncclInt32, s->comm, st));
    else DMG_CUDA(h, cudaMemcpyAsync(d_matrix, d_counts, (size_t)(G + 1) * 4, cudaMemcpyDeviceToDevice, st));
    DMG_CUDA(h, cudaMemcpyAsync(h_matrix, d_matrix, (size_t)(G + 1) * G * 4, cudaMemcpyDeviceToHost, st));
    DMG_CUDA(h, cudaStreamSynchronize(st));
    if (*(volatile int32_t *)h->h_flags) {
        h->h_flags[0] = 0;
        return fail(h, DMG_ERR_INDEX, "dmg_shard_train_step: node code outside [0, %lld), history code outside [-1, %lld) or bad mask position",
                    (long long)s->global_rows, (long long)s->global_rows);
    }
    auto cnt = [&](int from, int to) { return h_matrix[from * (G + 1) + to]; };
    int64_t n_total = 0, n_out = 0, n_in = 0;
    std::vector<int32_t> send_off(G, 0), recv_off(G, 0);
    for (int p = 0; p < G; p++) {
        n_total += h_matrix[p * (G + 1) + G];
        send_off[p] = (int32_t)n_out; n_out += cnt(me, p);
        recv_off[p] = (int32_t)n_in; n_in += cnt(p, me);
    }
    Scratch &sb = s->buf;
    DMG_TRY(ensure_dev(h, sb, Carver::need({(size_t)std::max<int64_t>(n_in, 1) * 4, (size_t)std::max<int64_t>(n_in, 1) * E * 4})));
    Carver cb(sb.d);
    int32_t *in_code = cb.take<int32_t>(std::max<int64_t>(n_in, 1));
    float *in_rows = cb.take<float>((size_t)std::max<int64_t>(n_in, 1) * E);
    DMG_CUDA(h, cudaMemcpyAsync(d_send_off, send_off.data(), (size_t)G * 4, cudaMemcpyHostToDevice, st));
    shtr_bucket_kernel<<<gocc, 256, 0, st>>>(n_occ, occ_code, g, d_send_off, d_cursor, req_code, req_occ);
    h->launches += 1;
    const int grid = h->sm_count * 8;
    auto exchange = [&](const void *out, void *in, size_t elem, ncclDataType_t dt, size_t per, bool reverse) -> int32_t {
        // forward: my requests -> owners; reverse: owners' answers -> requesters (same counts, roles swapped)
        if (G == 1) return DMG_OK;
        DMG_NCCL(h, g_nccl.GroupStart());
        for (int p = 0; p < G; p++) {
            if (p == me) continue;
            const int64_t ns = reverse ? cnt(p, me) : cnt(me, p), nr = reverse ? cnt(me, p) : cnt(p, me);
            const int64_t so = reverse ? recv_off[p] : send_off[p], ro = reverse ? send_off[p] : recv_off[p];
            if (ns) DMG_NCCL(h, g_nccl.Send((const char *)out + (size_t)so * per * elem, (size_t)ns * per, dt, p, s->comm, st));
            if (nr) DMG_NCCL(h, g_nccl.Recv((char *)in + (size_t)ro * per * elem, (size_t)nr * per, dt, p, s->comm, st));
        }
        DMG_NCCL(h, g_nccl.GroupEnd());
        return DMG_OK;
    };
    DMG_TRY(exchange(req_code, in_code, 4, ncclInt32, 1, false));                       // codes to the owners
    if (n_in) shtr_gather_kernel<<<grid, 256, 0, st>>>(n_in, E, in_code, g, d.emb<float>(), in_rows);
    DMG_TRY(exchange(in_rows, xfer, 4, ncclFloat32, (size_t)E, true));                  // rows back
    shtr_stage_local_kernel<<<grid, 256, 0, st>>>(n_occ, E, occ_code, g, d.emb<float>(), stage);
    if (n_out) shtr_stage_remote_kernel<<<grid, 256, 0, st>>>(n_out, E, req_occ, xfer, stage);
    shtr_remap_kernel<<<gocc, 256, 0, st>>>(rows, T, occ_code, node2, seq2);
    h->launches += 4;
    DMG_TRY(grad_enqueue<float>(h, rows, node2, seq2, d_mask, dl, d_loss, stage, sgrad, (double)n_total));
    float *gflat = (float *)d.d_grad;
    shtr_grad_local_kernel<<<grid, 256, 0, st>>>(n_occ, E, occ_code, g, sgrad, gflat);
    if (n_out) shtr_grad_pack_kernel<<<grid, 256, 0, st>>>(n_out, E, req_occ, sgrad, xfer);
    DMG_TRY(exchange(xfer, in_rows, 4, ncclFloat32, (size_t)E, false));                 // gradients to the owners
    if (n_in) shtr_grad_apply_kernel<<<grid, 256, 0, st>>>(n_in, E, in_code, g, in_rows, gflat);
    h->launches += 3;
    s->exchanged_rows += n_in;
    if (G > 1) {
        if (g.repl_rows > 0) DMG_NCCL(h, g_nccl.AllReduce(gflat, gflat, (size_t)g.repl_rows * E, ncclFloat32, ncclSum, s->comm, st));
        const size_t dense = (size_t)(d.n_params - d.rows * E);
        DMG_NCCL(h, g_nccl.AllReduce(gflat + d.rows * E, gflat + d.rows * E, dense, ncclFloat32, ncclSum, s->comm, st));
        DMG_NCCL(h, g_nccl.AllReduce(d_loss, d_loss, 1, ncclFloat64, ncclSum, s->comm, st));
    }
    DMG_TRY(adam_pass<float>(h, lr, step_t));
    double loss_sum = 0.0;
    DMG_CUDA(h, cudaMemcpyAsync(&loss_sum, d_loss, 8, cudaMemcpyDeviceToHost, st));
    DMG_CUDA(h, cudaStreamSynchronize(st));
    *out_loss = (float)(loss_sum / (double)n_total);
    return DMG_OK;
}
