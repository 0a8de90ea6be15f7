// rows_kernels.cuh -- DIN forward on n independent (node, history) rows: model.forward itself.
//
// The seam the reference calls once per tree level (Recommender.scala:94,
// otm CandidateSearcher.scala:41,77, OTMTree.scala:168,198, jtm TreeLearning.scala:168).
// Unlike the beam kernel every row carries its own T history indices, so the history tile
// cannot be shared; each CTA stages RB rows (1 + T embedding rows each) in shared memory and
// walks the graph phase by phase with one sequential-k fma chain per output element
// (same arithmetic spec as dmg_math.cuh / the beam kernel => identical bits).
#pragma once
#include "dmg_common.cuh"
#include "dmg_math.cuh"

namespace dmg {

constexpr int kRowsThreads = 128;
constexpr int kRowsRB = 8;          // rows per CTA iteration

template <typename real>
__global__ void __launch_bounds__(kRowsThreads) din_rows_forward_kernel(
    const real *__restrict__ emb, const real *__restrict__ wattT, const real *__restrict__ w1T,
    const real *__restrict__ b1, const real *__restrict__ w2, const real *__restrict__ b2, real scale, int E, int T,
    int64_t n, const int32_t *__restrict__ node, const int32_t *__restrict__ seq, const uint8_t *__restrict__ mask,
    real *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    real *sQ = reinterpret_cast<real *>(smem_raw);        // RB x E
    real *sK = sQ + kRowsRB * E;                           // RB x T x E
    real *sP = sK + kRowsRB * T * E;                       // RB x (T+1)
    real *sA = sP + kRowsRB * (T + 1);                     // RB x E
    real *sAtt = sA + kRowsRB * E;                         // RB x E
    real *sH = sAtt + kRowsRB * E;                         // RB x E
    const int tid = threadIdx.x;
    const int PL = T + 1;

    for (int64_t g0 = (int64_t)blockIdx.x * kRowsRB; g0 < n; g0 += (int64_t)gridDim.x * kRowsRB) {
        const int nr = (int)((n - g0) < kRowsRB ? (n - g0) : kRowsRB);
        // gather: EmbeddingShare on (item, seq), paddingIdx -> zeros
        for (int idx = tid; idx < nr * (T + 1) * E; idx += kRowsThreads) {
            const int k = idx % E, slot = (idx / E) % (T + 1), r = idx / (E * (T + 1));
            const int32_t c = slot == 0 ? node[g0 + r] : seq[(g0 + r) * T + slot - 1];
            const real v = c < 0 ? (real)0 : emb[(size_t)c * E + k];
            if (slot == 0) sQ[r * E + k] = v; else sK[(r * T + slot - 1) * E + k] = v;
        }
        __syncthreads();
        for (int idx = tid; idx < nr * T; idx += kRowsThreads) {
            const int r = idx / T, j = idx % T;
            const real *q = sQ + r * E, *kj = sK + (r * T + j) * E;
            real acc = (real)0;
            for (int k = 0; k < E; k++) acc = fma_(q[k], kj[k], acc);
            real s = mul_(acc, scale);
            if (mask[(g0 + r) * T + j]) s = mask_value<real>::get();
            sP[r * PL + j] = s;
        }
        __syncthreads();
        if (tid < nr) {
            real *pr = sP + tid * PL;
            real mx = pr[0];
            for (int j = 1; j < T; j++) { real v = pr[j]; mx = (v > mx || v != v) ? v : mx; }
            real sum = (real)0;
            for (int j = 0; j < T; j++) { real e = exp_(sub_(pr[j], mx)); pr[j] = e; sum = add_(sum, e); }
            const real inv = inv_(sum);
            for (int j = 0; j < T; j++) pr[j] = mul_(pr[j], inv);
        }
        __syncthreads();
        for (int idx = tid; idx < nr * E; idx += kRowsThreads) {
            const int r = idx / E, k = idx % E;
            real acc = (real)0;
            for (int j = 0; j < T; j++) acc = fma_(sP[r * PL + j], sK[(r * T + j) * E + k], acc);
            sA[idx] = acc;
        }
        __syncthreads();
        for (int idx = tid; idx < nr * E; idx += kRowsThreads) {
            const int r = idx / E, o = idx % E;
            real acc = (real)0;
            for (int k = 0; k < E; k++) acc = fma_(sA[r * E + k], __ldg(wattT + (size_t)k * E + o), acc);
            sAtt[idx] = acc;
        }
        __syncthreads();
        for (int idx = tid; idx < nr * E; idx += kRowsThreads) {
            const int r = idx / E, o = idx % E;
            real acc = (real)0;
            for (int k = 0; k < E; k++) acc = fma_(sQ[r * E + k], __ldg(w1T + (size_t)k * E + o), acc);
            for (int k = 0; k < E; k++) acc = fma_(sAtt[r * E + k], __ldg(w1T + (size_t)(E + k) * E + o), acc);
            sH[idx] = relu_(add_(acc, __ldg(b1 + o)));
        }
        __syncthreads();
        if (tid < nr) {
            real l = (real)0;
            for (int o = 0; o < E; o++) l = fma_(sH[tid * E + o], __ldg(w2 + o), l);
            out[g0 + tid] = add_(l, __ldg(b2));
        }
        __syncthreads();
    }
}

// mask tensor (flat positions row*T+j, Recommender.scala:141-147) -> dense n x T bytes
static __global__ void mask_scatter_kernel(const int32_t *__restrict__ flat, int64_t n_mask, int64_t limit,
                                    uint8_t *__restrict__ mask, int32_t *__restrict__ err_flag)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_mask) return;
    int64_t p = flat[i];
    if (p < 0 || p >= limit) { atomicExch(err_flag, 1); return; }
    mask[p] = 1;
}

// index validation for model.forward inputs: [-1, rows)
static __global__ void check_index_kernel(const int32_t *__restrict__ idx, int64_t n, int64_t rows, int32_t *__restrict__ err_flag)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t c = idx[i];
    if (c < -1 || (int64_t)c >= rows) atomicExch(err_flag, 1);
}

// [out][in] -> [in][out]
template <typename real>
__global__ void transpose_kernel(const real *__restrict__ src, real *__restrict__ dst, int n_out, int n_in)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out * n_in) return;
    int o = i / n_in, k = i % n_in;
    dst[(size_t)k * n_out + o] = src[i];
}

// randn(0, std) like Tensor.randn at model construction (EmbeddingShare.scala:21, Linear.scala:12);
// counter-based (splitmix64 -> Box-Muller) so any slice can be generated independently.
__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
template <typename real>
__global__ void randn_fill_kernel(real *__restrict__ dst, int64_t n, uint64_t seed, double std)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        uint64_t h = splitmix64(seed ^ splitmix64((uint64_t)i));
        uint32_t a = (uint32_t)h, b = (uint32_t)(h >> 32);
        float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);
        float u2 = ((float)(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
        float z = sqrtf(-2.0f * __logf(u1)) * __cosf(6.28318530718f * u2);
        dst[i] = (real)((double)z * std);
    }
}

}  // namespace dmg
