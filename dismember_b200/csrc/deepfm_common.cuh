// deepfm_common.cuh -- the strict (oracle-order) arithmetic of one DeepFM row, shared by the level-synchronous scorer (shard.cu) and the
// certified fast path (beam_wave_dfm.cuh).  tdm/src/main/scala/com/mass/tdm/model/DeepFM.scala:11-44, scalann/.../nn/FM.scala:14-44.
#pragma once
#include "dmg_math.cuh"

namespace dmg {

__device__ __forceinline__ float deepfm_finish(const float *x, const float *sK, const float *hrow, float square_sum,
                                               const float *sW2, float b2, int E, int T)
{
    float sum_square = 0.0f;
    for (int k = 0; k < E; k++) {
        float b = add_(0.0f, x[k]);
        for (int j = 0; j < T; j++) b = add_(b, sK[j * E + k]);
        sum_square = fma_(b, b, sum_square);
    }
    const float fm = __fdiv_rn(sub_(sum_square, square_sum), 2.0f);
    float dnn = 0.0f;
    for (int o = 0; o <= T; o++) dnn = fma_(hrow[o], sW2[o], dnn);
    return add_(fm, add_(dnn, b2));
}
// chain c of a row: c <= T hidden unit c (returns relu(acc + b1[c])), c == T + 1 the square sum
__device__ __forceinline__ float deepfm_chain(const float *x, const float *sK, const float *sW1, const float *sB1, int c, int E, int T)
{
    float acc = 0.0f;
    if (c <= T) {
        const float *w = sW1 + (size_t)c * (T + 1) * E;
        for (int k = 0; k < E; k++) acc = fma_(x[k], w[k], acc);
        for (int k = 0; k < T * E; k++) acc = fma_(sK[k], w[E + k], acc);
        return relu_(add_(acc, sB1[c]));
    }
    for (int k = 0; k < E; k++) acc = fma_(x[k], x[k], acc);
    for (int k = 0; k < T * E; k++) acc = fma_(sK[k], sK[k], acc);
    return acc;
}

// NC chains of one row advanced together (independent accumulators give the FMA pipe its ILP; every chain is still its
// own sequential-k chain, so the bits do not change).  Chains c0 .. c0+NC-1; chain T+1 is the square sum.
template <int NC>
__device__ __forceinline__ void deepfm_chains(const float *x, const float *sK, const float *sW1, const float *sB1, int c0, int E, int T,
                                              float *hout)
{
    const int F = T + 1;
    float acc[NC];
    const float *w[NC];
#pragma unroll
    for (int i = 0; i < NC; i++) { acc[i] = 0.0f; w[i] = sW1 + (size_t)(c0 + i < F ? c0 + i : 0) * F * E; }
    // 16-byte shared-memory loads: one of x (or of the history) and one per chain of its weights feed 4 fma steps each
#pragma unroll 2
    for (int k = 0; k < E; k += 4) {
        const float4 xv = *reinterpret_cast<const float4 *>(x + k);
#pragma unroll
        for (int i = 0; i < NC; i++) {
            const float4 wv = c0 + i < F ? *reinterpret_cast<const float4 *>(w[i] + k) : xv;
            acc[i] = fma_(xv.x, wv.x, acc[i]);
            acc[i] = fma_(xv.y, wv.y, acc[i]);
            acc[i] = fma_(xv.z, wv.z, acc[i]);
            acc[i] = fma_(xv.w, wv.w, acc[i]);
        }
    }
#pragma unroll 2
    for (int k = 0; k < T * E; k += 4) {
        const float4 kv = *reinterpret_cast<const float4 *>(sK + k);
#pragma unroll
        for (int i = 0; i < NC; i++) {
            const float4 wv = c0 + i < F ? *reinterpret_cast<const float4 *>(w[i] + E + k) : kv;
            acc[i] = fma_(kv.x, wv.x, acc[i]);
            acc[i] = fma_(kv.y, wv.y, acc[i]);
            acc[i] = fma_(kv.z, wv.z, acc[i]);
            acc[i] = fma_(kv.w, wv.w, acc[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < NC; i++)
        if (c0 + i <= F) hout[c0 + i] = c0 + i < F ? relu_(add_(acc[i], sB1[c0 + i])) : acc[i];
}

}  // namespace dmg
