// beam_wave_dfm.cuh -- certified fast path for TDM/JTM retrieval with the DeepFM scorer (E = 64, fp32 model).
//
// tdm/src/main/scala/com/mass/tdm/model/DeepFM.scala:11-44:  y = FM(F) + W2 . relu(W1 . Fflat + b1) + b2,  F = [item row x ; T history rows].
// The strict path (shard.cu) walks T + 2 chains of (T + 1) E fma steps per candidate row, because the reference's chains run over
// Fflat in order and nothing of a user can be hoisted out of them without changing the rounding.  Algebraically, per USER
//     s = sum_j K_j,   c = (|s|^2 - sum_j |K_j|^2) / 2,   u_o = W1h_o . Kflat + b1_o          (history half of Linear((T+1)E, T+1))
// and per candidate row only
//     y = x . s + c + sum_o w2_o relu(x . W1x_o + u_o) + b2                                    (12 dot products of length 64)
// i.e. 11x fewer multiply-adds and a kernel bound by the row gather.  The level-synchronous machinery of beam_wave.cuh is reused
// unchanged -- wave_select_kernel (certified cuts, deferred proofs, in-place strict re-scores), wave_final_prep / verify -- with
//     wave_dfm_prologue_kernel      per user: s, c, u_o and the user's terms of the error bound
//     wave_dfm_score_kernel         tiles of (user, 128 rows): rows staged with cp.async, one thread per row, 12 chains from u_o / c
//     dfm_strict_batch128           the oracle-order arithmetic (deepfm_common.cuh) on the band rows of a parked cut
//     wave_dfm_strict_rows_kernel   the same on the deferred band rows and the topk candidates at the end of the search
// so that ids and logits are the strict path's bits (the returned logits ARE strict scores).  N = 12 is no shape for the tensor
// cores; fp32 chains also keep the bound an order of magnitude tighter than bf16 x 3 would.
//
// Bound |fast - strict| <= eps for every row x of a level (u = 2^-24; forward error bounds, worst case):
//   hidden unit o:  strict = a 705-step chain, fast = u_o (a 641-step chain, once per user) continued over the 64 item terms
//       |dz_o| <= (705 + 66 + 2) u X_o + (705 + 641 + 2) u A_o + 2 u |b1_o|,   X_o = sum_k |x_k| |W1x_ok|,  A_o = sum |K| |W1h_o|
//   output:  sum_o |w2_o| |dz_o|  +  2 (F + 3) u (sum_o |w2_o| (X_o + A_o + |b1_o|) + |b2| + |fm|)          (both final dots + adds)
//   FM:  strict (sum_k (x_k + K_0k + ..)^2 - (|x|^2 + sum |K|^2)) / 2 cancels two large sums; with a_k = sum_j |K_jk|, a = |a|_2, q = sum |K|^2
//       |dfm| <= u [ (E + 2T + 4)/2 (|x| + a)^2 + ((T+1)E + 2)/2 (|x|^2 + q) + (E + 1)(|x| a + |c|) + T |x| a + (E + 1) a^2 / 2
//                    + (T E + 1) q / 2 + T a^2 ] + 2 u (|x| a + |c|)
// => eps = tau 1.25 ( max_level sum_k vt_k |x_k|  +  a0 + a1 nx + a2 nx^2 ),  nx = max |x|_2 of the level (level_bounds_kernel with vt),
// a0 / a1 / a2 per user from the prologue.  The largest |fast - strict| / eps observed is reported by dmg_fast_stats.
#pragma once
#include "beam_wave.cuh"
#include "deepfm_common.cuh"

namespace dmg {

// per-user block in WaveParams::uop (floats): [0, 64) s | [64, 80) u_o | [80] c
// ---- K2: per-user prologue ----------------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(128) wave_dfm_prologue_kernel(const WaveParams p, const DfmConsts dc, const float *__restrict__ dense)
{
    constexpr int E = 64;
    __shared__ __align__(16) float sK[kMaxT * E];
    __shared__ float sRed[8];
    __shared__ float sAo[16];
    const int tid = threadIdx.x, user = blockIdx.x, T = p.T, F = T + 1, IN = F * E;
    grid_dep_launch();
    grid_dep_wait();
    float *uop = reinterpret_cast<float *>(p.uop + (size_t)user * WaveGeo::UOP_BYTES);
    for (int i = tid; i < kMaxT * E; i += 128) {
        const int j = i >> 6;
        const int c = j < T ? p.hist[(size_t)user * T + j] : -1;
        sK[i] = c >= 0 ? __ldg(p.emb + (size_t)c * E + (i & 63)) : 0.0f;
    }
    if (tid < 8) sRed[tid] = 0.0f;
    __syncthreads();
    if (tid < E) {                                                 // s_k, a_k and their norms, q
        float s = 0.0f, a = 0.0f, q = 0.0f;
        for (int j = 0; j < T; j++) { const float v = sK[j * E + tid]; s += v; a += fabsf(v); q = fmaf(v, v, q); }
        uop[tid] = s;
        float s2 = s * s, a2 = a * a;
        for (int o = 16; o > 0; o >>= 1) { s2 += __shfl_xor_sync(0xffffffffu, s2, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
        if ((tid & 31) == 0) { atomicAdd(&sRed[0], s2); atomicAdd(&sRed[1], a2); atomicAdd(&sRed[2], q); }
    } else if (tid < E + 16) {                                     // u_o = W1h_o . Kflat + b1_o, A_o = sum |K| |W1h_o|
        const int o = tid - E;
        float acc = 0.0f, A = 0.0f;
        if (o < F) {
            const float *w = dense + (size_t)o * IN + E;
            for (int k = 0; k < T * E; k++) { const float wv = __ldg(w + k), kv = sK[k]; acc = fmaf(kv, wv, acc); A = fmaf(fabsf(kv), fabsf(wv), A); }
            acc += dc.b1[o];
        }
        uop[64 + o] = o < F ? acc : 0.0f;
        sAo[o] = o < F ? A : 0.0f;
    }
    __syncthreads();
    if (tid == 0) {
        const float u = 5.9604645e-8f;
        const float s2 = sRed[0], a = sqrtf(sRed[1]) * 1.0001f, q = sRed[2] * 1.0001f;
        const float c = (s2 - q) * 0.5f, ac = fabsf(c);
        uop[80] = c;
        const float Ef = (float)E, Tf = (float)T, Ff = (float)F;
        float dn = 0.0f, hsum = 0.0f;                              // hidden-unit terms of the user
        for (int o = 0; o < F; o++) {
            dn += dc.aw2[o] * ((705.0f + (Tf * Ef + 1.0f) + 2.0f) * u * sAo[o] + 2.0f * u * fabsf(dc.b1[o]));
            hsum += dc.aw2[o] * (sAo[o] + fabsf(dc.b1[o]));
        }
        const float fd = 2.0f * (Ff + 3.0f) * u;                   // final dots and adds of both paths
        float a0 = dn + fd * (hsum + fabsf(dc.b2) + ac) +
                   u * ((Ef + 2.0f * Tf + 4.0f) * 0.5f * a * a + (Ff * Ef + 2.0f) * 0.5f * q + (Ef + 3.0f) * ac + (Ef + 1.0f) * 0.5f * a * a +
                        (Tf * Ef + 1.0f) * 0.5f * q + Tf * a * a);
        float a1 = u * a * (2.0f * Ef + 3.0f * Tf + 7.0f) + fd * a;
        float a2 = u * ((Ef + 2.0f * Tf + 4.0f) * 0.5f + (Ff * Ef + 2.0f) * 0.5f);
        if (!(a0 == a0) || !(a1 == a1)) { a0 = __int_as_float(0x7f800000); a1 = a0; }
        WaveUser st;
        memset(&st, 0, sizeof(st));
        st.zk = a0; st.kmax = a1; st.hw = a2;
        p.user[user] = st;
        p.count[user] = 0;
    }
}

// ---- strict (oracle-order) scores of n rows of ONE user by a CTA of NT threads ----------------------------------------------------------
// sK: the user's history, flat [T][64] (zero rows for padding); scr: RG * (64 + 4 + 20) floats.  Every row is worked on by
// ceil((T + 2) / 6) threads, each advancing 6 of the row's T + 2 chains together (deepfm_chains: 16-byte loads, the chains stay
// sequential in k, so the bits are deepfm_chain's).
template <int NT, int RG>
__device__ __forceinline__ void dfm_strict_rows(const float *__restrict__ emb, const int32_t *__restrict__ codes, int n, float *__restrict__ out,
                                                const float *__restrict__ sK, const float *__restrict__ dense, const DfmConsts &dc, int T,
                                                float *__restrict__ scr)
{
    constexpr int E = 64, LD = E + 4, HL = 20;
    const int F = T + 1, IN = F * E, nch = (F + 1 + 5) / 6;
    float *sX = scr, *sH = scr + RG * LD;
    const float *w1 = dense, *b1 = dense + (size_t)F * IN, *w2 = b1 + F;
    const int rpg = NT / nch < RG ? NT / nch : RG;                 // rows per pass
    for (int base = 0; base < n; base += rpg) {
        const int nr = n - base < rpg ? n - base : rpg;
        __syncthreads();
        for (int i = threadIdx.x; i < nr * 16; i += NT) {
            const int r = i >> 4, c16 = i & 15;
            *reinterpret_cast<float4 *>(sX + r * LD + c16 * 4) = ldg_row16(emb + (size_t)codes[base + r] * E + c16 * 4);
        }
        __syncthreads();
        if ((int)threadIdx.x < nr * nch) {
            const int r = threadIdx.x / nch, part = threadIdx.x - r * nch;
            deepfm_chains<6>(sX + r * LD, sK, w1, b1, part * 6, E, T, sH + r * HL);
        }
        __syncthreads();
        for (int r = threadIdx.x; r < nr; r += NT) out[base + r] = deepfm_finish(sX + r * LD, sK, sH + r * HL, sH[r * HL + F], w2, dc.b2, E, T);
    }
    __syncthreads();
}

// the band rows of a parked cut (wave_select_kernel, 128 threads): scr = the kernel's strict scratch, history loaded here
__device__ __noinline__ void dfm_strict_batch128(const WaveParams &p, const DfmConsts &dc, const float *__restrict__ dense, int user,
                                                 const int32_t *__restrict__ sRow, int n, float *__restrict__ sOut, float *__restrict__ scr,
                                                 float *__restrict__ sDense, bool dense_loaded)
{
    constexpr int E = 64;
    float *sK = scr;                                               // [kMaxT][64]
    if (!dense_loaded) {                                           // 31 KB once per CTA with a parked cut: the 176-step chains would otherwise
        const int Fd = p.T + 1, n_dense = Fd * Fd * E + 2 * Fd + 1;   // wait for L2 at every step
        for (int i = threadIdx.x; i < n_dense; i += 128) sDense[i] = __ldg(dense + i);
    }
    dense = sDense;
    for (int i = threadIdx.x; i < kMaxT * E; i += 128) {
        const int j = i >> 6;
        const int c = j < p.T ? p.hist[(size_t)user * p.T + j] : -1;
        sK[i] = c >= 0 ? __ldg(p.emb + (size_t)c * E + (i & 63)) : 0.0f;
    }
    __syncthreads();
    dfm_strict_rows<128, 32>(p.emb, sRow, n, sOut, sK, dense, dc, p.T, scr + kMaxT * E);
}

// ---- K1 (score): tiles of (user, <= 128 rows), one thread per row -------------------------------------------------------------------------
// The item half of W1 travels as a KERNEL PARAMETER (2.8 KB of the 4 KB parameter space = constant bank): with the k loop fully
// unrolled every multiply-add reads its weight as a constant-bank operand, so the inner loop is 11 FFMA per k with no load at all
// (the same weights through shared memory made the kernel LSU-bound: three 16-byte broadcast loads per k and warp).
struct DfmW1x { float w[64 * 11]; };    // [k][o] = W1[o][k], o < 11 (T + 1 <= 11 on this path)

static __global__ void __launch_bounds__(128) wave_dfm_score_kernel(const WaveParams p, const DfmConsts dc, const __grid_constant__ DfmW1x wx,
                                                                    int slot, int level)
{
    constexpr int E = 64, LD = 68;
    extern __shared__ __align__(16) unsigned char dfm_raw[];
    float *sX = reinterpret_cast<float *>(dfm_raw);                // [128][LD]
    float *sU = sX + 128 * LD;                                     // [96]: s | u_o | c of the tile's user
    const int tid = threadIdx.x;
    grid_dep_launch();
    grid_dep_wait();
    const int ntiles = *(volatile const int32_t *)(p.tile_count + level);
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int tile = __ldg(p.tile_list + t);
        const int user = tile >> 10, row0 = ((tile >> 8) & 3) * 128, nr = (tile & 255) + 1;
        const int32_t *codes = p.code[slot] + (size_t)user * p.cap + row0;
        __syncthreads();                                           // the previous tile's readers are done with sX / sU
        for (int i = tid; i < nr * 16; i += 128) {
            const int r = i >> 4, c16 = i & 15;
            cp_async16(sX + r * LD + c16 * 4, p.emb + (size_t)codes[r] * E + c16 * 4);
        }
        if (tid < 24) cp_async16(sU + tid * 4, reinterpret_cast<const float *>(p.uop + (size_t)user * WaveGeo::UOP_BYTES) + tid * 4);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        if (tid < nr) {
            float acc[12];
#pragma unroll
            for (int o = 0; o < 11; o++) acc[o] = sU[64 + o];
            acc[11] = sU[80];
            const float *x = sX + tid * LD;
#pragma unroll
            for (int k = 0; k < E; k += 4) {
                const float4 xv = *reinterpret_cast<const float4 *>(x + k), sv = *reinterpret_cast<const float4 *>(sU + k);
                const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ss[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
                for (int q = 0; q < 4; q++) {
#pragma unroll
                    for (int o = 0; o < 11; o++) acc[o] = fmaf(xs[q], wx.w[(k + q) * 11 + o], acc[o]);
                    acc[11] = fmaf(xs[q], ss[q], acc[11]);
                }
            }
            float dnn = 0.0f;
#pragma unroll
            for (int o = 0; o < 11; o++) if (o < dc.F) dnn = fmaf(fmaxf(acc[o], 0.0f), dc.w2[o], dnn);
            p.score[(size_t)user * p.cap + row0 + tid] = acc[11] + (dnn + dc.b2);
        }
    }
}

// ---- K3: strict scores of the deferred band rows and the topk candidates, persistent over users ------------------------------------------
static __global__ void __launch_bounds__(256) wave_dfm_strict_rows_kernel(const WaveParams p, const DfmConsts dc, const float *__restrict__ dense,
                                                                          const WaveFinal wf)
{
    constexpr int E = 64;
    extern __shared__ __align__(16) unsigned char dfm_raw[];
    float *sDense = reinterpret_cast<float *>(dfm_raw);            // [W1 | b1 | W2 | b2]: the chains of a row read W1 from here
    __shared__ __align__(16) float sK[kMaxT * E];
    grid_dep_launch();
    const int Fd = p.T + 1, n_dense = Fd * Fd * E + 2 * Fd + 1;
    float *sScr = sDense + ((n_dense + 3) & ~3);                   // 128 x (68 + 20) floats
    for (int i = threadIdx.x; i < n_dense; i += 256) sDense[i] = __ldg(dense + i);   // weights: constant for the whole chain
    grid_dep_wait();
    dense = sDense;
    for (int user = blockIdx.x; user < p.B; user += gridDim.x) {
        const int n_rows = wf.meta[(size_t)user * 4];
        if (n_rows == 0) continue;
        __syncthreads();
        for (int i = threadIdx.x; i < kMaxT * E; i += 256) {
            const int j = i >> 6;
            const int c = j < p.T ? p.hist[(size_t)user * p.T + j] : -1;
            sK[i] = c >= 0 ? __ldg(p.emb + (size_t)c * E + (i & 63)) : 0.0f;
        }
        __syncthreads();
        dfm_strict_rows<256, 128>(p.emb, wf.row_code + (size_t)user * WaveFinal::RCAP, n_rows, wf.row_strict + (size_t)user * WaveFinal::RCAP, sK, dense, dc,
                                 p.T, sScr);
    }
}

}  // namespace dmg
