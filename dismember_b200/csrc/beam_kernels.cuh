// beam_kernels.cuh -- K1/K2/K3: persistent CTA-per-user beam search with the fused DIN scorer.
//
// Replaces, for a whole batch of users in ONE launch:
//   Recommender._recommend            tdm/src/main/scala/com/mass/tdm/model/Recommender.scala:40-107
//   CandidateSearcher.batchBeamSearch otm/src/main/scala/com/mass/otm/model/CandidateSearcher.scala:15-56
//   DIN.buildModel graph forward      tdm/.../model/DIN.scala:14-43 (EmbeddingShare, Attention, Mask,
//                                     SoftMax, MatMul, Concat, Linear, ReLU of scalann/.../nn)
//   final stable sort + take(topk)    Recommender.scala:37, TDM.scala:21, OTM.scala:17-21
//
// One CTA owns one user at a time and walks every tree level without leaving the SM: the
// beam, candidate codes and scores live in shared memory, the user's T history rows are
// staged once through the TMA bulk-copy engine (cp.async.bulk + mbarrier), candidate rows are
// gathered with 16-byte cp.async into a double-buffered row tile while the previous tile is
// being scored, and only topk results leave the chip.  Users are independent, so there is no
// grid-wide synchronisation and no per-level kernel launch.
//
// "strict" arithmetic (dmg_math.cuh): every GEMM element is one sequential-k fma chain in
// registers, so results are bit-identical to the CPU oracle.
#pragma once
#include "dmg_common.cuh"
#include "dmg_math.cuh"
#include <algorithm>

#include "device_utils.cuh"

namespace dmg {

enum { MODE_TDM_TOPK = 0, MODE_OTM_DUMP = 1, MODE_OTM_TOPK = 2 };

template <typename real> struct BeamParams {
    const real *emb, *wattT, *w1T, *b1, *w2, *b2;
    real scale;
    int T, B;
    int E_run;                  // embed size of the tables above when it differs from the loaded model's (zero-padded copy), else 0
    const int32_t *hist;        // B x T embedding indices, -1 = padding (zero row)
    const uint8_t *hist_mask;   // B x T, 1 = position listed in the Mask input
    int beam;
    const int32_t *beam_user;   // nullable: per-user beam (Recommender.scala:28-31)
    int always_sort;            // OTM sorts every level after the first; TDM only when > beam
    const uint32_t *exists;     // code-existence bitmap, nullptr = complete tree
    int leaf_level;
    int mode, topk;
    const int32_t *leaf_item;   // [2^L] item id of a leaf slot, -1 = none
    const int64_t *cons_off;    // nullable consumed-items CSR
    const int32_t *cons;
    int32_t *out_items;
    real *out_scores;
    int32_t *out_counts;
    int out_stride;
    int cap, capp;              // candidate capacity (>= 2*max beam) and its power of two
    // optional: every level's scored candidates (OTMTree.beamSearchNodes), [user][level][lvl_stride]
    int32_t *lvl_items;
    real *lvl_scores;
    int32_t *lvl_counts;        // [user][n_lvl]
    int lvl_stride, n_lvl;
    // optional: run only the users user_list[0 .. *user_count) (the fast kernel's redo list)
    const int32_t *user_list;
    const int32_t *user_count;
    // > 1: the kernel is launched in clusters of `split` CTAs that run ONE user together -- every CTA keeps the whole
    // beam state (same sorts, same expansion), scores only the tiles t % split == its cluster rank and writes those
    // scores into every peer's shared memory (DSMEM).  Used for the redo list: a lone strict user costs 0.85 ms on one SM.
    int split;
};

// ---- compile-time geometry ----------------------------------------------------------------
template <typename real, int E, int RT = (sizeof(real) == 4 ? 128 : 64)> struct Geo {
    static constexpr int R = RT;                                 // rows per tile (beam search: 128 fp32 / 64 fp64)
    static constexpr int NBUF = sizeof(real) == 4 ? 2 : 1;       // row-tile buffers
    static constexpr int VEC = 16 / sizeof(real);                // reals per 16 B
    static constexpr int LD = E + VEC;                           // padded row stride
    static constexpr int PLD = kMaxT + 1;                        // score/prob row stride
    static constexpr int TX = E / 4;                             // GEMM: column threads (4 cols each)
    static constexpr int TY = kThreads / TX;                     // GEMM: row threads
    static constexpr int TM = R / TY;                            // GEMM: rows per thread
    static constexpr int NP = kThreads / R;                      // attention: threads per row
    static constexpr int MAXJ = (kMaxT + NP - 1) / NP;
    static constexpr int KW = E / NP;                            // attention: k-range per thread
    static_assert(E % 4 == 0 && TX * TY == kThreads && TM * TY == R && TM >= 1, "bad geometry");
    static_assert(KW % 4 == 0 || KW == 2 || KW == 4, "bad attention split");

    static size_t smem_bytes(int cap, int capp)
    {
        size_t reals = (size_t)E * E + 2 * (size_t)E * E + 2 * (size_t)E + 4 + (size_t)kMaxT * E +
                       (size_t)NBUF * R * LD + (size_t)R * LD + (size_t)R * PLD + (size_t)cap;
        size_t b = reals * sizeof(real);
        b = (b + 15) & ~(size_t)15;
        b += (size_t)capp * sizeof(typename KeyOf<real>::type);
        b += (size_t)cap * 2 * sizeof(int32_t);
        b += 64 * sizeof(int32_t);
        b += 16;                                                 // mbarrier
        return b;
    }
};

// ---- the DIN scorer on one tile of R candidate rows -----------------------------------------
// sX: gathered candidate rows [R][LD]; sK: history [T][E]; results -> sScore[0..nrows).
// PER_ROW: every row brings its own history (model.forward on independent rows): sK is [R][T][E], sMask [R][T].
template <typename real, int E, int RT = (sizeof(real) == 4 ? 128 : 64), bool PER_ROW = false>
__device__ __forceinline__ void score_tile(const real *__restrict__ sX, real *__restrict__ sA, real *__restrict__ sP,
                                           const real *__restrict__ sK, const int32_t *__restrict__ sMask,
                                           const real *__restrict__ sWattT, const real *__restrict__ sW1T,
                                           const real *__restrict__ sB1, const real *__restrict__ sW2,
                                           real b2, real scale, int T, int nrows, real *__restrict__ sScoreOut)
{
    using G = Geo<real, E, RT>;
    const int tid = threadIdx.x;

    // (1) attention scores: s[r][j] = scale * sum_k x[r][k] K[j][k]   (MatMul transB + Mask)
    {
        const int r = tid / G::NP, part = tid % G::NP;
        if (r < nrows) {
            real acc[G::MAXJ];
#pragma unroll
            for (int jj = 0; jj < G::MAXJ; jj++) acc[jj] = (real)0;
            const real *xr = sX + r * G::LD;
#pragma unroll 2
            for (int k = 0; k < E; k += 4) {
                real xv[4];
                ld4(xr + k, xv);
#pragma unroll
                for (int jj = 0; jj < G::MAXJ; jj++) {
                    const int j = part + jj * G::NP;
                    if (j < T) {
                        real kv[4];
                        ld4(sK + (PER_ROW ? r * T + j : j) * E + k, kv);
                        acc[jj] = fma_(xv[0], kv[0], acc[jj]);
                        acc[jj] = fma_(xv[1], kv[1], acc[jj]);
                        acc[jj] = fma_(xv[2], kv[2], acc[jj]);
                        acc[jj] = fma_(xv[3], kv[3], acc[jj]);
                    }
                }
            }
#pragma unroll
            for (int jj = 0; jj < G::MAXJ; jj++) {
                const int j = part + jj * G::NP;
                if (j < T) {
                    real s = mul_(acc[jj], scale);
                    if (sMask[PER_ROW ? r * T + j : j]) s = mask_value<real>::get();
                    sP[r * G::PLD + j] = s;
                }
            }
        }
    }
    __syncthreads();
    // (2) softmax over T per row (SoftMax.scala:27-42)
    if (tid < nrows) {
        real *pr = sP + tid * G::PLD;
        real mx = pr[0];
        for (int j = 1; j < T; j++) { real v = pr[j]; mx = (v > mx || v != v) ? v : mx; }
        real sum = (real)0;
        for (int j = 0; j < T; j++) { real e = exp_(sub_(pr[j], mx)); pr[j] = e; sum = add_(sum, e); }
        const real inv = inv_(sum);
        for (int j = 0; j < T; j++) pr[j] = mul_(pr[j], inv);
    }
    __syncthreads();
    // (3) a[r][k] = sum_j p[r][j] K[j][k]   (MatMul)
    {
        const int r = tid / G::NP, part = tid % G::NP;
        if (r < nrows) {
            real acc[G::KW];
#pragma unroll
            for (int kk = 0; kk < G::KW; kk++) acc[kk] = (real)0;
            const real *pr = sP + r * G::PLD;
            for (int j = 0; j < T; j++) {
                const real pj = pr[j];
                const real *kj = sK + (PER_ROW ? r * T + j : j) * E + part * G::KW;
#pragma unroll
                for (int kk = 0; kk < G::KW; kk += G::VEC) {
                    if constexpr (G::VEC == 4) {
                        real kv[4];
                        ld4(kj + kk, kv);
#pragma unroll
                        for (int q = 0; q < 4; q++) acc[kk + q] = fma_(pj, kv[q], acc[kk + q]);
                    } else {
#pragma unroll
                        for (int q = 0; q < G::VEC; q++) acc[kk + q] = fma_(pj, kj[kk + q], acc[kk + q]);
                    }
                }
            }
            real *ar = sA + r * G::LD + part * G::KW;
#pragma unroll
            for (int kk = 0; kk < G::KW; kk++) ar[kk] = acc[kk];
        }
    }
    __syncthreads();
    // (4) att = a . Watt^T   (Linear(E,E), no bias) ; (5) h = relu([x|att] . W1^T + b1)
    const int ty = tid / G::TX, tx = tid % G::TX;
    const int row0 = ty * G::TM;
    const bool active = row0 < nrows;
    real acc[G::TM][4];
#pragma unroll
    for (int i = 0; i < G::TM; i++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[i][c] = (real)0;
    if (active) {
#pragma unroll 4
        for (int k = 0; k < E; k += 4) {
            real bv[4][4];
#pragma unroll
            for (int kk = 0; kk < 4; kk++) ld4(sWattT + (k + kk) * E + tx * 4, bv[kk]);
#pragma unroll
            for (int i = 0; i < G::TM; i++) {
                real av[4];
                ld4(sA + (row0 + i) * G::LD + k, av);
#pragma unroll
                for (int kk = 0; kk < 4; kk++)
#pragma unroll
                    for (int c = 0; c < 4; c++) acc[i][c] = fma_(av[kk], bv[kk][c], acc[i][c]);
            }
        }
    }
    __syncthreads();                       // every read of a[] done before att overwrites it
    if (active) {
#pragma unroll
        for (int i = 0; i < G::TM; i++) st4(sA + (row0 + i) * G::LD + tx * 4, acc[i]);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < G::TM; i++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[i][c] = (real)0;
    if (active) {
#pragma unroll 4
        for (int k = 0; k < E; k += 4) {               // item half of the concat
            real bv[4][4];
#pragma unroll
            for (int kk = 0; kk < 4; kk++) ld4(sW1T + (k + kk) * E + tx * 4, bv[kk]);
#pragma unroll
            for (int i = 0; i < G::TM; i++) {
                real av[4];
                ld4(sX + (row0 + i) * G::LD + k, av);
#pragma unroll
                for (int kk = 0; kk < 4; kk++)
#pragma unroll
                    for (int c = 0; c < 4; c++) acc[i][c] = fma_(av[kk], bv[kk][c], acc[i][c]);
            }
        }
#pragma unroll 4
        for (int k = 0; k < E; k += 4) {               // attention half
            real bv[4][4];
#pragma unroll
            for (int kk = 0; kk < 4; kk++) ld4(sW1T + (E + k + kk) * E + tx * 4, bv[kk]);
#pragma unroll
            for (int i = 0; i < G::TM; i++) {
                real av[4];
                ld4(sA + (row0 + i) * G::LD + k, av);
#pragma unroll
                for (int kk = 0; kk < 4; kk++)
#pragma unroll
                    for (int c = 0; c < 4; c++) acc[i][c] = fma_(av[kk], bv[kk][c], acc[i][c]);
            }
        }
        real b1v[4];
        ld4(sB1 + tx * 4, b1v);
#pragma unroll
        for (int i = 0; i < G::TM; i++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[i][c] = relu_(add_(acc[i][c], b1v[c]));
    }
    __syncthreads();                       // every read of att[] done before h overwrites it
    if (active) {
#pragma unroll
        for (int i = 0; i < G::TM; i++) st4(sA + (row0 + i) * G::LD + tx * 4, acc[i]);
    }
    __syncthreads();
    // (6) logit = h . W2 + b2   (Linear(E,1)), sequential over the hidden units
    if (tid < nrows) {
        const real *hr = sA + tid * G::LD;
        real l = (real)0;
#pragma unroll 4
        for (int o = 0; o < E; o += 4) {
            real hv[4], wv[4];
            ld4(hr + o, hv);
            ld4(sW2 + o, wv);
            l = fma_(hv[0], wv[0], l);
            l = fma_(hv[1], wv[1], l);
            l = fma_(hv[2], wv[2], l);
            l = fma_(hv[3], wv[3], l);
        }
        sScoreOut[tid] = add_(l, b2);
    }
    __syncthreads();
}

template <typename real, int E>
__device__ __forceinline__ void gather_tile(real *__restrict__ sXbuf, const real *__restrict__ emb,
                                            const int32_t *__restrict__ codes, int nrows)
{
    using G = Geo<real, E>;
    constexpr int VPR = E / G::VEC;        // 16-byte vectors per row
    for (int idx = threadIdx.x; idx < nrows * VPR; idx += blockDim.x) {
        const int r = idx / VPR, v = idx % VPR;
        cp_async16(sXbuf + r * G::LD + v * G::VEC, emb + (size_t)codes[r] * E + v * G::VEC);
    }
}

// ---- model.forward on independent rows, tiled (dmg_score_pairs, JTM weights) ------------------------------------------
// Every row brings its own T history indices (Recommender.scala:94, OTMTree.scala:168,198, TreeLearning.scala:168).  RT rows
// per CTA iteration: candidate rows and the RT x T history rows staged with cp.async, weights in shared memory once per
// CTA, then score_tile in PER_ROW mode -- the strict kernel's register-tiled sequential-k chains, so the bits are those
// of the beam search and of the oracle.  Replaces the row-at-a-time din_rows_forward_kernel for E in {16, 32, 64}.
template <typename real, int E> struct RowsGeo {
    static constexpr int RMIN = sizeof(real) == 4 ? 32 : 16;
    static constexpr int RT = (1024 / E > RMIN) ? 1024 / E : RMIN;          // TM = RT / (kThreads / (E/4)) >= 1
    using G = Geo<real, E, RT>;
    static size_t smem_bytes(int T)
    {
        const size_t reals = (size_t)3 * E * E + 2 * E + 4 + (size_t)RT * T * E + 2 * (size_t)RT * G::LD + (size_t)RT * G::PLD + RT;
        return ((reals * sizeof(real) + 15) & ~(size_t)15) + (size_t)RT * T * sizeof(int32_t) + 16;
    }
};
template <typename real, int E>
__global__ void __launch_bounds__(kThreads) din_rows_tiled_kernel(const real *__restrict__ emb, const real *__restrict__ wattT,
                                                                  const real *__restrict__ w1T, const real *__restrict__ b1,
                                                                  const real *__restrict__ w2, const real *__restrict__ b2, real scale,
                                                                  int T, int64_t n, const int32_t *__restrict__ node,
                                                                  const int32_t *__restrict__ seq, const uint8_t *__restrict__ mask,
                                                                  real *__restrict__ out)
{
    using RG = RowsGeo<real, E>;
    using G = typename RG::G;
    constexpr int RT = RG::RT, VPR = E / G::VEC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    real *sWattT = reinterpret_cast<real *>(smem_raw);
    real *sW1T = sWattT + E * E;
    real *sB1 = sW1T + 2 * E * E;
    real *sW2 = sB1 + E;
    real *sB2 = sW2 + E;                          // 4 reals
    real *sK = sB2 + 4;                           // RT x T x E
    real *sX = sK + (size_t)RT * T * E;           // RT x LD
    real *sA = sX + RT * G::LD;
    real *sP = sA + RT * G::LD;
    real *sOut = sP + RT * G::PLD;                // RT
    const size_t off = (((size_t)((unsigned char *)(sOut + RT) - smem_raw)) + 15) & ~(size_t)15;
    int32_t *sMask = reinterpret_cast<int32_t *>(smem_raw + off);            // RT x T
    const int tid = threadIdx.x;
    for (int i = tid; i < E * E; i += kThreads) sWattT[i] = wattT[i];
    for (int i = tid; i < 2 * E * E; i += kThreads) sW1T[i] = w1T[i];
    for (int i = tid; i < E; i += kThreads) { sB1[i] = b1[i]; sW2[i] = w2[i]; }
    if (tid == 0) sB2[0] = b2[0];
    __syncthreads();
    const real b2v = sB2[0];
    for (int64_t g0 = (int64_t)blockIdx.x * RT; g0 < n; g0 += (int64_t)gridDim.x * RT) {
        const int nr = (int)((n - g0) < RT ? (n - g0) : RT);
        __syncthreads();
        for (int idx = tid; idx < nr * (T + 1) * VPR; idx += kThreads) {
            const int v = idx % VPR, slot = (idx / VPR) % (T + 1), r = idx / (VPR * (T + 1));
            const int32_t c = slot == 0 ? node[g0 + r] : seq[(g0 + r) * T + slot - 1];
            real *dst = slot == 0 ? sX + r * G::LD + v * G::VEC : sK + ((size_t)(r * T + slot - 1) * E + v * G::VEC);
            if (c >= 0) cp_async16(dst, emb + (size_t)c * E + v * G::VEC);
            else
#pragma unroll
                for (int q = 0; q < G::VEC; q++) dst[q] = (real)0;       // paddingIdx -> zero row (LookupTable.scala:33-34)
        }
        for (int i = tid; i < nr * T; i += kThreads) sMask[i] = mask[g0 * T + i];
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        score_tile<real, E, RT, true>(sX, sA, sP, sK, sMask, sWattT, sW1T, sB1, sW2, b2v, scale, T, nr, sOut);
        for (int i = tid; i < nr; i += kThreads) out[g0 + i] = sOut[i];
    }
}
template <typename real, int E>
static cudaError_t launch_rows_tiled(const real *emb, const real *wattT, const real *w1T, const real *b1, const real *w2, const real *b2,
                                     real scale, int T, int64_t n, const int32_t *node, const int32_t *seq, const uint8_t *mask, real *out,
                                     int sm_count, size_t smem_per_sm, cudaStream_t st)
{
    using RG = RowsGeo<real, E>;
    const size_t smem = RG::smem_bytes(T);
    auto kern = din_rows_tiled_kernel<real, E>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, smem_per_sm / (smem + 1024)));
    const int64_t tiles = (n + RG::RT - 1) / RG::RT;
    const int grid = (int)std::min<int64_t>(tiles, (int64_t)sm_count * per_sm);
    kern<<<grid, kThreads, smem, st>>>(emb, wattT, w1T, b1, w2, b2, scale, T, n, node, seq, mask, out);
    return cudaGetLastError();
}
// true when the tiled kernel took the rows (E with a tile geometry), false: the caller falls back to din_rows_forward_kernel
template <typename real>
static bool rows_forward_tiled(int E, const real *emb, const real *wattT, const real *w1T, const real *b1, const real *w2, const real *b2,
                               real scale, int T, int64_t n, const int32_t *node, const int32_t *seq, const uint8_t *mask, real *out,
                               int sm_count, size_t smem_per_sm, size_t smem_optin, cudaStream_t st, cudaError_t *err)
{
    *err = cudaSuccess;
    switch (E) {
    case 16: if (RowsGeo<real, 16>::smem_bytes(T) > smem_optin) return false;
             *err = launch_rows_tiled<real, 16>(emb, wattT, w1T, b1, w2, b2, scale, T, n, node, seq, mask, out, sm_count, smem_per_sm, st); return true;
    case 32: if (RowsGeo<real, 32>::smem_bytes(T) > smem_optin) return false;
             *err = launch_rows_tiled<real, 32>(emb, wattT, w1T, b1, w2, b2, scale, T, n, node, seq, mask, out, sm_count, smem_per_sm, st); return true;
    case 64: if (RowsGeo<real, 64>::smem_bytes(T) > smem_optin) return false;
             *err = launch_rows_tiled<real, 64>(emb, wattT, w1T, b1, w2, b2, scale, T, n, node, seq, mask, out, sm_count, smem_per_sm, st); return true;
    default: return false;
    }
}

// ---- thread-block cluster helpers (split mode) ------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_f32(const float *local_smem, uint32_t peer, float v)
{
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_smem)), "r"(peer));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_f32(const double *, uint32_t, double) {}   // split mode is fp32 only

template <typename real, int E>
__global__ void __launch_bounds__(kThreads, 1) beam_search_kernel(const BeamParams<real> p)
{
    using G = Geo<real, E>;
    using KO = KeyOf<real>;
    using KeyT = typename KO::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    real *sWattT = reinterpret_cast<real *>(smem_raw);
    real *sW1T = sWattT + E * E;
    real *sB1 = sW1T + 2 * E * E;
    real *sW2 = sB1 + E;
    real *sB2 = sW2 + E;                         // 4 reals
    real *sK = sB2 + 4;
    real *sX = sK + kMaxT * E;
    real *sA = sX + G::NBUF * G::R * G::LD;
    real *sP = sA + G::R * G::LD;
    real *sScore = sP + G::R * G::PLD;
    size_t off = ((size_t)((unsigned char *)(sScore + p.cap) - smem_raw) + 15) & ~(size_t)15;
    KeyT *sKey = reinterpret_cast<KeyT *>(smem_raw + off);
    int32_t *sCode0 = reinterpret_cast<int32_t *>(sKey + p.capp);
    int32_t *sCode1 = sCode0 + p.cap;
    int32_t *sMisc = sCode1 + p.cap;             // [0..15] history codes, [16..31] mask, [32..39] scan, [40..] scalars
    uint64_t *sBar = reinterpret_cast<uint64_t *>(sMisc + 64);

    const int tid = threadIdx.x;
    const int T = p.T;
    const int split = p.split > 1 ? p.split : 1;
    const uint32_t crank = split > 1 ? cluster_ctarank() : 0u;
    if (p.user_count && (int)blockIdx.x / split >= *p.user_count) return;   // redo launch with nothing (left) to redo (whole cluster)

    // weights -> shared, once per CTA
    for (int i = tid; i < E * E; i += kThreads) sWattT[i] = p.wattT[i];
    for (int i = tid; i < 2 * E * E; i += kThreads) sW1T[i] = p.w1T[i];
    for (int i = tid; i < E; i += kThreads) { sB1[i] = p.b1[i]; sW2[i] = p.w2[i]; }
    if (tid == 0) { sB2[0] = p.b2[0]; mbar_init(sBar, 1); }
    __syncthreads();
    const real b2 = sB2[0];
    uint32_t bar_phase = 0;

    const int n_users = p.user_count ? *p.user_count : p.B;
    for (int ui = blockIdx.x / split; ui < n_users; ui += gridDim.x / split) {
        const int user = p.user_list ? p.user_list[ui] : ui;
        // ---- K2: history tile --------------------------------------------------------------
        if (tid < kMaxT) {
            int c = -1, m = 0;
            if (tid < T) { c = p.hist[(size_t)user * T + tid]; m = p.hist_mask[(size_t)user * T + tid]; }
            sMisc[tid] = c;
            sMisc[16 + tid] = m;
        }
        __syncthreads();
        fence_proxy_async();
        if (tid == 0) {
            uint32_t bytes = 0;
            for (int j = 0; j < T; j++) if (sMisc[j] >= 0) bytes += E * sizeof(real);
            mbar_expect_tx(sBar, bytes);
            for (int j = 0; j < T; j++)
                if (sMisc[j] >= 0) tma_bulk_g2s(sK + j * E, p.emb + (size_t)sMisc[j] * E, E * sizeof(real), sBar);
        }
        for (int i = tid; i < T * E; i += kThreads)
            if (sMisc[i / E] < 0) sK[i] = (real)0;             // paddingIdx -> zero row
        mbar_wait(sBar, bar_phase);
        bar_phase ^= 1;
        __syncthreads();

        // ---- initial beam: every existing code of level s, pred 0 --------------------------
        const int beam = p.beam_user ? p.beam_user[user] : p.beam;
        int s_level = 31 - __clz(beam);                          // floor(log2 beam)
        int32_t *cur = sCode0, *nxt = sCode1;
        int count = 0;
        if (s_level <= p.leaf_level) {
            const int64_t start = ((int64_t)1 << s_level) - 1;
            const int n0 = 1 << s_level;
            for (int base = 0; base < n0; base += kThreads) {
                int i = base + tid;
                int e = (i < n0 && code_exists(p.exists, start + i)) ? 1 : 0;
                int tot;
                int o = block_exscan(e, sMisc + 32, &tot);
                if (e) cur[count + o] = (int32_t)(start + i);
                count += tot;
            }
            for (int i = tid; i < count; i += kThreads) sScore[i] = (real)0;
            __syncthreads();
        }

        // ---- level loop ---------------------------------------------------------------------
        for (int level = s_level; level < p.leaf_level && count > 0; level++) {
            int nb = count;
            if (count > beam || (p.always_sort && level != s_level)) {
                int n2 = 2;
                while (n2 < count) n2 <<= 1;
                for (int i = tid; i < n2; i += kThreads) sKey[i] = i < count ? KO::make(sScore[i], i) : KO::lowest();
                __syncthreads();
                bitonic_sort_desc(sKey, n2);
                nb = count < beam ? count : beam;
                for (int i = tid; i < nb; i += kThreads) nxt[i] = cur[KO::pos(sKey[i])];
                __syncthreads();
                int32_t *t = cur; cur = nxt; nxt = t;
            }
            // children 2c+1, 2c+2 that exist, order preserved (Recommender.scala:88-92)
            int nc = 0;
            if (p.exists == nullptr) {
                for (int i = tid; i < nb; i += kThreads) { int32_t c = cur[i]; nxt[2 * i] = 2 * c + 1; nxt[2 * i + 1] = 2 * c + 2; }
                nc = 2 * nb;
                __syncthreads();
            } else {
                for (int base = 0; base < nb; base += kThreads) {
                    int i = base + tid;
                    int64_t c = i < nb ? cur[i] : 0;
                    int e1 = (i < nb && code_exists(p.exists, 2 * c + 1)) ? 1 : 0;
                    int e2 = (i < nb && code_exists(p.exists, 2 * c + 2)) ? 1 : 0;
                    int tot;
                    int o = block_exscan(e1 + e2, sMisc + 32, &tot);
                    if (e1) nxt[nc + o] = (int32_t)(2 * c + 1);
                    if (e2) nxt[nc + o + e1] = (int32_t)(2 * c + 2);
                    nc += tot;
                }
                __syncthreads();
            }
            { int32_t *t = cur; cur = nxt; nxt = t; }
            count = nc;
            // score the candidates tile by tile, prefetching the next tile's rows
            const int ntiles = (count + G::R - 1) / G::R;
            if (split > 1) {
                // every CTA of the cluster holds the same candidates; this one scores tiles crank, crank + split, ...
                for (int t = crank; t < ntiles; t += split) {
                    const int r0 = t * G::R;
                    const int nrows = count - r0 < G::R ? count - r0 : G::R;
                    __syncthreads();
                    gather_tile<real, E>(sX, p.emb, cur + r0, nrows);
                    cp_async_commit();
                    cp_async_wait<0>();
                    __syncthreads();
                    score_tile<real, E>(sX, sA, sP, sK, sMisc + 16, sWattT, sW1T, sB1, sW2, b2, p.scale, T, nrows, sScore + r0);
                }
                cluster_sync_all();                               // every peer is done reading the previous level's scores
                for (int t = crank; t < ntiles; t += split) {
                    const int r0 = t * G::R;
                    const int nrows = count - r0 < G::R ? count - r0 : G::R;
                    for (int i = threadIdx.x; i < nrows * (split - 1); i += kThreads) {
                        const int r = i % nrows;
                        uint32_t peer = (uint32_t)(i / nrows);
                        peer += peer >= crank ? 1u : 0u;
                        st_cluster_f32(sScore + r0 + r, peer, sScore[r0 + r]);
                    }
                }
                cluster_sync_all();                               // all slices have landed everywhere
            } else
            if (ntiles > 0) {
                gather_tile<real, E>(sX, p.emb, cur, count < G::R ? count : G::R);
                cp_async_commit();
            }
            for (int t = 0; split == 1 && t < ntiles; t++) {
                const int r0 = t * G::R;
                const int nrows = count - r0 < G::R ? count - r0 : G::R;
                real *buf = sX + (G::NBUF == 2 ? (t & 1) : 0) * G::R * G::LD;
                if (G::NBUF == 2 && t + 1 < ntiles) {
                    const int r1 = r0 + G::R;
                    gather_tile<real, E>(sX + ((t + 1) & 1) * G::R * G::LD, p.emb, cur + r1,
                                         count - r1 < G::R ? count - r1 : G::R);
                    cp_async_commit();
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                __syncthreads();
                score_tile<real, E>(buf, sA, sP, sK, sMisc + 16, sWattT, sW1T, sB1, sW2, b2, p.scale, T, nrows,
                                    sScore + r0);
                if (G::NBUF == 1 && t + 1 < ntiles) {
                    const int r1 = r0 + G::R;
                    gather_tile<real, E>(sX, p.emb, cur + r1, count - r1 < G::R ? count - r1 : G::R);
                    cp_async_commit();
                }
            }
            if (p.lvl_items) {
                const int li = level - s_level;
                if (li < p.n_lvl) {
                    const size_t base = ((size_t)user * p.n_lvl + li) * p.lvl_stride;
                    for (int i = tid; i < p.lvl_stride; i += kThreads) {
                        p.lvl_items[base + i] = i < count ? cur[i] : -1;
                        p.lvl_scores[base + i] = i < count ? sScore[i] : (real)0;
                    }
                    if (tid == 0) p.lvl_counts[(size_t)user * p.n_lvl + li] = count;
                }
            }
        }

        // ---- K3: results ----------------------------------------------------------------------
        // TDM: candidates that did not reach the leaf level are dropped (they never become leaves).
        const int64_t leaf_start = ((int64_t)1 << p.leaf_level) - 1;
        const bool at_leaf = (s_level <= p.leaf_level);
        if (p.mode == MODE_OTM_DUMP) {
            for (int i = tid; i < p.out_stride; i += kThreads) {
                p.out_items[(size_t)user * p.out_stride + i] = i < count ? cur[i] : -1;
                p.out_scores[(size_t)user * p.out_stride + i] = i < count ? sScore[i] : (real)0;
            }
            if (tid == 0) p.out_counts[user] = count;
        } else {
            int n2 = 2;
            while (n2 < count) n2 <<= 1;
            const int64_t c0 = p.cons_off ? p.cons_off[user] : 0, c1 = p.cons_off ? p.cons_off[user + 1] : 0;
            for (int i = tid; i < n2; i += kThreads) {
                KeyT k = KO::lowest();
                if (i < count && at_leaf) {
                    int64_t slot = (int64_t)cur[i] - leaf_start;
                    int32_t item = (slot >= 0 && slot < ((int64_t)1 << p.leaf_level)) ? __ldg(p.leaf_item + slot) : -1;
                    bool keep = item >= 0;
                    for (int64_t q = c0; q < c1 && keep; q++) keep = (__ldg(p.cons + q) != item);
                    if (keep) k = KO::make(sScore[i], i);
                }
                sKey[i] = k;
            }
            __syncthreads();
            if (count > 0) bitonic_sort_desc(sKey, n2);
            for (int i = tid; i < p.topk; i += kThreads) {
                int32_t item = -1;
                real sc = (real)0;
                if (i < count && !KO::is_lowest(sKey[i])) {
                    int pos = KO::pos(sKey[i]);
                    item = __ldg(p.leaf_item + ((int64_t)cur[pos] - leaf_start));
                    sc = sScore[pos];
                }
                p.out_items[(size_t)user * p.out_stride + i] = item;
                p.out_scores[(size_t)user * p.out_stride + i] = sc;
            }
            if (tid == 0) {
                int valid = 0;
                const int lim = count < p.topk ? count : p.topk;
                while (valid < lim && !KO::is_lowest(sKey[valid])) valid++;
                p.out_counts[user] = valid;
            }
        }
        __syncthreads();
    }
}

// ---- K2 (host-facing part): item ids -> codes + mask, validity ------------------------------
// TDMTree.idToCode (tdm/src/main/scala/com/mass/tdm/tree/TDMTree.scala:35-56).
static __global__ void tdm_ids_to_codes_kernel(const int32_t *__restrict__ ids, int64_t n, const int32_t *__restrict__ id_code,
                                        int32_t non_leaf_offset, int32_t max_code, int64_t table_rows, int use_mask,
                                        int32_t *__restrict__ codes, uint8_t *__restrict__ mask, int32_t *__restrict__ err_flag,
                                        int32_t *__restrict__ fast_ctl)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x == 0 && fast_ctl)                                   // user scheduler + redo count + per-SM tail owners of the fast kernel
        for (int j = threadIdx.x; j < DMG_FAST_CTL_WORDS; j += blockDim.x) fast_ctl[j] = 0;
    if (i >= n) return;
    const int32_t id = ids[i];
    int32_t code;
    uint8_t m = 0;
    if (id == 0) { m = 1; code = -1; }
    else if (id > 0 && id < non_leaf_offset && id_code[id] >= 0) code = id_code[id];
    else {
        int64_t tmp = (int64_t)id - non_leaf_offset;
        if (tmp > max_code) { m = 1; code = -1; }
        else code = (int32_t)tmp;
    }
    if (code < -1 || (int64_t)code >= table_rows) { atomicExch(err_flag, 1); code = -1; }
    codes[i] = code;
    mask[i] = use_mask ? m : 0;
}

// OTM / generic: sequence entries are already embedding indices; mask where == -1.
static __global__ void seq_to_codes_kernel(const int32_t *__restrict__ seq, int64_t n, int64_t table_rows, int use_mask,
                                    int32_t *__restrict__ codes, uint8_t *__restrict__ mask, int32_t *__restrict__ err_flag)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t c = seq[i];
    if (c < -1 || (int64_t)c >= table_rows) { atomicExch(err_flag, 1); c = -1; }
    codes[i] = c;
    mask[i] = (use_mask && c == -1) ? 1 : 0;
}

}  // namespace dmg
