// shard.cu -- node table sharded by code range across the GPUs of one box ("forced-shard" mode).
//
// The reference has no multi-device path (tdm/.../optim/LocalOptimizer.scala:35-40 clones the model per thread);
// this is the engine's own layout for catalogues whose node table (or table + optimiser state) exceeds one GPU.
// World size G = 2^g.  Tree levels above g are replicated (G - 1 rows); on every level l >= g rank r owns the
// 2^(l-g) consecutive codes under the r-th node of level g, stored level after level in its local table.
// Users are sharded too: every rank brings its own B queries.  One level of TDM beam search
// (tdm/.../model/Recommender.scala:58-99) becomes
//     requester: stable top-b + expansion of the children (shard_select_expand_kernel, the reference's sort keys)
//     requester: bucket (slot, code) by owner (shard_bucket_kernel)            8 B per candidate
//     NCCL all-to-all (ncclSend/ncclRecv group over NVLink)                    requests
//     owner:     DIN forward on its rows with the requester's history tile     strict fp32, bits of the oracle
//     NCCL all-to-all back                                                     4 B per candidate
//     requester: scatter the scores into candidate order (shard_scatter_kernel)
// The history tiles (T rows per user) are fetched once per batch: every rank contributes the rows it owns into a
// zero-filled [G*B, T, E] buffer and an integer-sum all-reduce (ncclUint32: x + 0 is bit-exact, also for -0.0)
// leaves every rank with every user's tile, so an owner scores any requester's candidates without a second hop.
// NCCL is resolved with dlopen at dmg_shard_init: the library has no link-time dependency on it.
#include <algorithm>
#include <cmath>
#include "beam_kernels.cuh"
#include "rows_kernels.cuh"
#include "shard_common.cuh"
#include "deepfm_common.cuh"

using namespace dmg;

namespace dmg {

static int64_t shard_local_rows(int bits, int max_level)
{
    return (((int64_t)1 << bits) - 1) + (((int64_t)1 << (max_level - bits + 1)) - 1);
}

// ---- weights ------------------------------------------------------------------------------------------------
// Same values as dmg_init_din_weights on the unsharded table: element (global row, k) = randn(seed, row*E + k).
static __global__ void shard_randn_rows_kernel(float *__restrict__ dst, ShardGeo g, int64_t local_rows, int E, uint64_t seed, double std)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, n = local_rows * E;
    for (; i < n; i += stride) {
        const int64_t lr = i / E, k = i % E;
        const uint64_t gi = (uint64_t)(shard_global_row(g, lr) * E + k);
        uint64_t hsh = splitmix64(seed ^ splitmix64(gi));
        uint32_t a = (uint32_t)hsh, b = (uint32_t)(hsh >> 32);
        float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);
        float u2 = ((float)(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
        float z = sqrtf(-2.0f * __logf(u1)) * __cosf(6.28318530718f * u2);
        dst[i] = (float)((double)z * std);
    }
}
// dense tail [W_att | W1]: flat indices rows_global*E + i of the unsharded vector
static __global__ void shard_randn_tail_kernel(float *__restrict__ dst, int64_t n, uint64_t first, uint64_t seed, double std)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        uint64_t hsh = splitmix64(seed ^ splitmix64(first + (uint64_t)i));
        uint32_t a = (uint32_t)hsh, b = (uint32_t)(hsh >> 32);
        float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);
        float u2 = ((float)(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
        float z = sqrtf(-2.0f * __logf(u1)) * __cosf(6.28318530718f * u2);
        dst[i] = (float)((double)z * std);
    }
}

// ---- history tiles ------------------------------------------------------------------------------------------
// tiles[gu][j][:] = bits of the embedding row of codes[gu][j] if this rank is its (single) contributor, else 0.
// Replicated levels are contributed by rank 0 only, padding (-1) by nobody.
static __global__ void shard_fill_tiles_kernel(const float *__restrict__ emb, ShardGeo g, const int32_t *__restrict__ codes,
                                               int64_t n_slots, int E, uint32_t *__restrict__ tiles)
{
    const int vec = E / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots * vec; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = i / vec;
        const int v = (int)(i % vec);
        const int32_t c = codes[s];
        uint4 val = make_uint4(0u, 0u, 0u, 0u);
        if (c >= 0) {
            const bool repl = code_level(c) < g.bits;
            if (repl ? g.rank == 0 : shard_owner(g, c) == g.rank)
                val = *reinterpret_cast<const uint4 *>(emb + shard_local_row(g, c) * E + v * 4);
        }
        *reinterpret_cast<uint4 *>(tiles + s * E + v * 4) = val;
    }
}

// ---- requester side: beam state per local user ---------------------------------------------------------------
// cand / score: [B][cap] candidates of the current level in the reference's order; count[B].
// first != 0: fill the start level floor(log2 beam) with every existing code, score 0 (Recommender.scala:51-57).
// Otherwise: keep the best `beam` by the reference's stable descending sort when count > beam (:75-87), then
// expand the children 2c+1, 2c+2 that exist, order preserved (:88-92).
static __global__ void __launch_bounds__(kThreads) shard_select_expand_kernel(int32_t *__restrict__ cand, const float *__restrict__ score,
                                                                              int32_t *__restrict__ count, int cap, int capp, int beam,
                                                                              int level, int first, const uint32_t *__restrict__ exists,
                                                                              const int32_t *__restrict__ beam_user = nullptr, int leaf_level = 0)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *sKey = reinterpret_cast<uint64_t *>(smem_raw);
    int32_t *sCur = reinterpret_cast<int32_t *>(sKey + capp);
    int32_t *sNxt = sCur + cap;
    __shared__ int sScan[8];
    const int u = blockIdx.x, tid = threadIdx.x;
    int32_t *uc = cand + (size_t)u * cap;
    bool fill_then_expand = false;
    if (beam_user) {                                              // per-user widened beams (Recommender.scala:27-33): user u starts at
        beam = beam_user[u];                                      // floor(log2 beam_u) and keeps beam_u candidates per level; one launch
        const int su = 31 - __clz(beam);                          // per level serves every user: not started / start level / running
        if (level < su || (level == leaf_level && su != level)) return;
        first = su == level;
        fill_then_expand = first && level < leaf_level;
    }
    if (first) {
        const int64_t start = ((int64_t)1 << level) - 1;
        const int n0 = 1 << level;
        int cnt = 0;
        for (int base = 0; base < n0; base += kThreads) {
            const int i = base + tid;
            const int e = (i < n0 && code_exists(exists, start + i)) ? 1 : 0;
            int tot;
            const int o = block_exscan(e, sScan, &tot);
            if (e && cnt + o < cap) uc[cnt + o] = (int32_t)(start + i);
            cnt += tot;
        }
        if (tid == 0) count[u] = cnt < cap ? cnt : cap;
        if (!fill_then_expand) return;
        __syncthreads();                                          // the start level holds at most beam_u codes: no cut, expand right away
    }
    const int n = count[u];
    const float *us = score + (size_t)u * cap;
    for (int i = tid; i < n; i += kThreads) sCur[i] = uc[i];
    __syncthreads();
    int nb = n;
    if (n > beam) {
        int n2 = 2;
        while (n2 < n) n2 <<= 1;
        for (int i = tid; i < n2; i += kThreads) sKey[i] = i < n ? KeyOf<float>::make(us[i], i) : KeyOf<float>::lowest();
        __syncthreads();
        bitonic_sort_desc(sKey, n2);
        nb = beam;
        for (int i = tid; i < nb; i += kThreads) sNxt[i] = sCur[KeyOf<float>::pos(sKey[i])];
        __syncthreads();
        for (int i = tid; i < nb; i += kThreads) sCur[i] = sNxt[i];
        __syncthreads();
    }
    int nc = 0;
    for (int base = 0; base < nb; base += kThreads) {
        const int i = base + tid;
        const int64_t c = i < nb ? sCur[i] : 0;
        const int e1 = (i < nb && code_exists(exists, 2 * c + 1)) ? 1 : 0;
        const int e2 = (i < nb && code_exists(exists, 2 * c + 2)) ? 1 : 0;
        int tot;
        const int o = block_exscan(e1 + e2, sScan, &tot);
        if (e1) uc[nc + o] = (int32_t)(2 * c + 1);
        if (e2) uc[nc + o + e1] = (int32_t)(2 * c + 2);
        nc += tot;
    }
    if (tid == 0) count[u] = nc;
}

// Requests (slot = u*cap + pos, code) appended to the region of the code's owner; n_out[o] counts them.
// seg[o * B + u] = (first request of user u in the region of owner o, how many): one user's requests to one owner are
// contiguous, so the owner scores them as tiles that share the user's history.
static __global__ void __launch_bounds__(kThreads) shard_bucket_kernel(const int32_t *__restrict__ cand, const int32_t *__restrict__ count,
                                                                       int cap, ShardGeo g, int2 *__restrict__ region, int64_t region_stride,
                                                                       int32_t *__restrict__ n_out, int2 *__restrict__ seg)
{
    __shared__ int sCnt[32], sBase[32];
    const int u = blockIdx.x, tid = threadIdx.x;
    if (tid < 32) sCnt[tid] = 0;
    __syncthreads();
    const int n = count[u];
    int own[2], loc[2];
    for (int q = 0; q < 2; q++) {
        const int i = tid + q * kThreads;
        own[q] = -1;
        if (i < n) {
            own[q] = shard_owner(g, cand[(size_t)u * cap + i]);
            loc[q] = atomicAdd(&sCnt[own[q]], 1);
        }
    }
    __syncthreads();
    if (tid < g.world) {
        sBase[tid] = sCnt[tid] ? atomicAdd(&n_out[tid], sCnt[tid]) : 0;
        seg[(size_t)tid * gridDim.x + u] = make_int2(sBase[tid], sCnt[tid]);
    }
    __syncthreads();
    for (int q = 0; q < 2; q++) {
        const int i = tid + q * kThreads;
        if (own[q] >= 0) region[(size_t)own[q] * region_stride + sBase[own[q]] + loc[q]] = make_int2(u * cap + i, cand[(size_t)u * cap + i]);
    }
}

static __global__ void shard_scatter_kernel(const int2 *__restrict__ req, const float *__restrict__ reply, int n, float *__restrict__ score)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) score[req[i].x] = reply[i];
}

// ---- owner side: DIN forward (model.forward, DIN.scala:14-43) on requested rows --------------------------------
// Row i: node = local row of req[i].y, history = tile of global user gu_base + req[i].x / cap.  Same phases and the
// same sequential-k fma chains as din_rows_forward_kernel / the oracle => identical bits.
static __global__ void __launch_bounds__(kRowsThreads) shard_score_rows_kernel(
    const float *__restrict__ emb, ShardGeo g, const float *__restrict__ wattT, const float *__restrict__ w1T,
    const float *__restrict__ b1, const float *__restrict__ w2, const float *__restrict__ b2, float scale, int E, int T,
    int n, const int2 *__restrict__ req, int cap, int gu_base, const float *__restrict__ tiles, const uint8_t *__restrict__ mask,
    float *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sQ = reinterpret_cast<float *>(smem_raw);        // RB x E
    float *sK = sQ + kRowsRB * E;                            // RB x T x E
    float *sP = sK + kRowsRB * T * E;                        // RB x (T+1)
    float *sA = sP + kRowsRB * (T + 1);                      // RB x E
    float *sAtt = sA + kRowsRB * E;                          // RB x E
    float *sH = sAtt + kRowsRB * E;                          // RB x E
    __shared__ int sGu[kRowsRB];
    const int tid = threadIdx.x;
    const int PL = T + 1;
    for (int g0 = blockIdx.x * kRowsRB; g0 < n; g0 += gridDim.x * kRowsRB) {
        const int nr = (n - g0) < kRowsRB ? (n - g0) : kRowsRB;
        if (tid < nr) sGu[tid] = gu_base + req[g0 + tid].x / cap;
        __syncthreads();
        for (int idx = tid; idx < nr * (T + 1) * E; idx += kRowsThreads) {
            const int k = idx % E, slot = (idx / E) % (T + 1), r = idx / (E * (T + 1));
            float v;
            if (slot == 0) v = emb[(size_t)shard_local_row(g, req[g0 + r].y) * E + k];
            else v = tiles[((size_t)sGu[r] * T + slot - 1) * E + k];
            if (slot == 0) sQ[r * E + k] = v; else sK[(r * T + slot - 1) * E + k] = v;
        }
        __syncthreads();
        for (int idx = tid; idx < nr * T; idx += kRowsThreads) {
            const int r = idx / T, j = idx % T;
            const float *q = sQ + r * E, *kj = sK + (r * T + j) * E;
            float acc = 0.0f;
            for (int k = 0; k < E; k++) acc = fma_(q[k], kj[k], acc);
            float s = mul_(acc, scale);
            if (mask[(size_t)sGu[r] * T + j]) s = mask_value<float>::get();
            sP[r * PL + j] = s;
        }
        __syncthreads();
        if (tid < nr) {
            float *pr = sP + tid * PL;
            float mx = pr[0];
            for (int j = 1; j < T; j++) { float v = pr[j]; mx = (v > mx || v != v) ? v : mx; }
            float sum = 0.0f;
            for (int j = 0; j < T; j++) { float e = exp_(sub_(pr[j], mx)); pr[j] = e; sum = add_(sum, e); }
            const float inv = inv_(sum);
            for (int j = 0; j < T; j++) pr[j] = mul_(pr[j], inv);
        }
        __syncthreads();
        for (int idx = tid; idx < nr * E; idx += kRowsThreads) {
            const int r = idx / E, k = idx % E;
            float acc = 0.0f;
            for (int j = 0; j < T; j++) acc = fma_(sP[r * PL + j], sK[(r * T + j) * E + k], acc);
            sA[idx] = acc;
        }
        __syncthreads();
        for (int idx = tid; idx < nr * E; idx += kRowsThreads) {
            const int r = idx / E, o = idx % E;
            float acc = 0.0f;
            for (int k = 0; k < E; k++) acc = fma_(sA[r * E + k], __ldg(wattT + (size_t)k * E + o), acc);
            sAtt[idx] = acc;
        }
        __syncthreads();
        for (int idx = tid; idx < nr * E; idx += kRowsThreads) {
            const int r = idx / E, o = idx % E;
            float acc = 0.0f;
            for (int k = 0; k < E; k++) acc = fma_(sQ[r * E + k], __ldg(w1T + (size_t)k * E + o), acc);
            for (int k = 0; k < E; k++) acc = fma_(sAtt[r * E + k], __ldg(w1T + (size_t)(E + k) * E + o), acc);
            sH[idx] = relu_(add_(acc, __ldg(b1 + o)));
        }
        __syncthreads();
        if (tid < nr) {
            float l = 0.0f;
            for (int o = 0; o < E; o++) l = fma_(sH[tid * E + o], __ldg(w2 + o), l);
            out[g0 + tid] = add_(l, __ldg(b2));
        }
        __syncthreads();
    }
}

// Tiled owner scorer: one (requester rank p, user u) segment at a time per CTA -- the user's history tile and mask are
// staged once, the requested rows are gathered R at a time and scored by score_tile (beam_kernels.cuh: 8x4 register
// tiles of sequential-k fma chains, the strict kernel's scorer) => identical bits, ~10x the row rate of the
// row-at-a-time kernel above, which stays as the fallback for embedding sizes without a tile geometry.
struct ShardScoreArgs {
    const float *emb, *wattT, *w1T, *b1, *w2, *b2, *tiles;
    const uint8_t *mask;
    const int2 *req_self, *req_peer;        // [G][stride] requests: own region (rank) / received regions
    const int2 *seg_self, *seg_peer;        // [G][B] segment tables
    float *out_self, *out_peer;             // [G][stride] scores: own region written straight into the reply buffer
    int64_t stride;
    int B, G, T, cap;
    float scale;
    int32_t *work;
    ShardGeo geo;
};
template <int E>
static __global__ void __launch_bounds__(kThreads, 1) shard_score_segments_kernel(const ShardScoreArgs a)
{
    using G = Geo<float, E>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sWattT = reinterpret_cast<float *>(smem_raw);
    float *sW1T = sWattT + E * E;
    float *sB1 = sW1T + 2 * E * E;
    float *sW2 = sB1 + E;
    float *sB2 = sW2 + E;                        // 4 floats
    float *sK = sB2 + 4;                         // kMaxT x E
    float *sX = sK + kMaxT * E;                  // R x LD
    float *sA = sX + G::R * G::LD;
    float *sP = sA + G::R * G::LD;
    float *sOut = sP + G::R * G::PLD;            // R
    int32_t *sMask = reinterpret_cast<int32_t *>(sOut + G::R);   // 16 + [16] next segment
    const int tid = threadIdx.x, T = a.T;
    for (int i = tid; i < E * E; i += kThreads) sWattT[i] = a.wattT[i];
    for (int i = tid; i < 2 * E * E; i += kThreads) sW1T[i] = a.w1T[i];
    for (int i = tid; i < E; i += kThreads) { sB1[i] = a.b1[i]; sW2[i] = a.w2[i]; }
    if (tid == 0) sB2[0] = a.b2[0];
    __syncthreads();
    const float b2 = sB2[0];
    const int n_seg = a.G * a.B;
    for (;;) {
        __syncthreads();
        if (tid == 0) sMask[16] = atomicAdd(a.work, 1);
        __syncthreads();
        const int si = sMask[16];
        if (si >= n_seg) break;
        const int p = si / a.B;
        const bool self = p == a.geo.rank;
        const int2 sg = (self ? a.seg_self : a.seg_peer)[si];
        if (sg.y == 0) continue;
        const int2 *rq = (self ? a.req_self : a.req_peer) + (size_t)p * a.stride + sg.x;
        float *out = (self ? a.out_self : a.out_peer) + (size_t)p * a.stride + sg.x;
        const float *tile = a.tiles + (size_t)si * T * E;          // si = p * B + u = global user
        for (int i = tid; i < T * E; i += kThreads) sK[i] = tile[i];
        if (tid < kMaxT) sMask[tid] = tid < T ? a.mask[(size_t)si * T + tid] : 0;
        for (int r0 = 0; r0 < sg.y; r0 += G::R) {
            const int nrows = sg.y - r0 < G::R ? sg.y - r0 : G::R;
            constexpr int VPR = E / 4;
            __syncthreads();
            for (int idx = tid; idx < nrows * VPR; idx += kThreads) {
                const int r = idx / VPR, v = idx % VPR;
                cp_async16(sX + r * G::LD + v * 4, a.emb + (size_t)shard_local_row(a.geo, rq[r0 + r].y) * E + v * 4);
            }
            cp_async_commit();
            cp_async_wait<0>();
            __syncthreads();
            score_tile<float, E>(sX, sA, sP, sK, sMask, sWattT, sW1T, sB1, sW2, b2, a.scale, T, nrows, sOut);
            __syncthreads();
            for (int i = tid; i < nrows; i += kThreads) out[r0 + i] = sOut[i];
        }
    }
}
template <int E> static size_t shard_score_smem()
{
    using G = Geo<float, E>;
    return ((size_t)3 * E * E + 2 * E + 4 + (size_t)kMaxT * E + 2 * (size_t)G::R * G::LD + (size_t)G::R * G::PLD + G::R) * 4 + 32 * 4;
}

// ---- DeepFM scorer (tdm/.../model/DeepFM.scala:11-44, scalann/.../nn/FM.scala:14-44) -------------------------------
// Features F = [item row ; T history rows]; logit = (|sum_i F_i|^2 - sum |F|^2) / 2 + W2.relu(W1.Fflat + b1) + b2, no mask.
// The Linear chains run over Fflat in order (item first, then the history), so nothing of a user can be hoisted out of
// the rows without changing the rounding: every row walks its 12 chains (T+1 hidden units + the square sum) of (T+1)E
// fma steps.  dense = [W1 (T+1) x (T+1)E | b1 | W2 | b2] staged in shared memory once per CTA.
struct DfmGeo {
    static constexpr int R = 128;
    static size_t smem(int E, int T) { const int F = T + 1; return ((size_t)F * F * E + 2 * F + 4 + (size_t)kMaxT * E + (size_t)R * (E + 4) + (size_t)R * (F + 1)) * 4 + 32 * 4 + 64; }
};
struct ShardDfmArgs {
    const float *emb, *dense, *tiles;
    const int2 *req_self, *req_peer, *seg_self, *seg_peer;
    float *out_self, *out_peer;
    int64_t stride;
    int B, G, T, E;
    int32_t *work;
    ShardGeo geo;
};
static __global__ void __launch_bounds__(kThreads, 1) shard_score_deepfm_kernel(const ShardDfmArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int E = a.E, T = a.T, F = T + 1, LD = E + 4, HL = F + 1;      // LD: rows 16-byte aligned, 8 rows span the 32 banks
    float *sW1 = reinterpret_cast<float *>(smem_raw);
    float *sB1 = sW1 + (size_t)F * F * E;
    float *sW2 = sB1 + F;
    float *sB2 = sW2 + F;                        // 4 floats
    float *sK = sW1 + (((size_t)F * F * E + 2 * F + 4 + 3) & ~(size_t)3);      // kMaxT x E, 16-byte aligned
    float *sX = sK + kMaxT * E;                  // R x LD
    float *sH = sX + DfmGeo::R * LD;             // R x HL: hidden units, [F] = square sum
    int32_t *sCtl = reinterpret_cast<int32_t *>(sH + DfmGeo::R * HL);
    const int tid = threadIdx.x;
    for (int i = tid; i < F * F * E + 2 * F + 1; i += kThreads) sW1[i] = a.dense[i];      // W1 | b1 | W2 | b2 are contiguous
    __syncthreads();
    const float b2 = sB2[0];
    const int n_seg = a.G * a.B;
    for (;;) {
        __syncthreads();
        if (tid == 0) sCtl[0] = atomicAdd(a.work, 1);
        __syncthreads();
        const int si = sCtl[0];
        if (si >= n_seg) break;
        const int p = si / a.B;
        const bool self = p == a.geo.rank;
        const int2 sg = (self ? a.seg_self : a.seg_peer)[si];
        if (sg.y == 0) continue;
        const int2 *rq = (self ? a.req_self : a.req_peer) + (size_t)p * a.stride + sg.x;
        float *out = (self ? a.out_self : a.out_peer) + (size_t)p * a.stride + sg.x;
        const float *tile = a.tiles + (size_t)si * T * E;
        for (int i = tid; i < T * E; i += kThreads) sK[i] = tile[i];
        for (int r0 = 0; r0 < sg.y; r0 += DfmGeo::R) {
            const int nrows = sg.y - r0 < DfmGeo::R ? sg.y - r0 : DfmGeo::R;
            __syncthreads();
            for (int idx = tid; idx < nrows * E; idx += kThreads) {
                const int r = idx / E, k = idx % E;
                sX[r * LD + k] = a.emb[(size_t)shard_local_row(a.geo, rq[r0 + r].y) * E + k];
            }
            __syncthreads();
            // two threads per row, 6 chains each (T = 10: 11 hidden units + the square sum); longer histories loop
            for (int idx = tid; idx < nrows * ((F + 6) / 6); idx += kThreads) {
                const int r = idx % nrows, c0 = (idx / nrows) * 6;
                deepfm_chains<6>(sX + r * LD, sK, sW1, sB1, c0, E, T, sH + r * HL);
            }
            __syncthreads();
            for (int r = tid; r < nrows; r += kThreads)
                out[r0 + r] = deepfm_finish(sX + r * LD, sK, sH + r * HL, sH[r * HL + F], sW2, b2, E, T);
        }
    }
}

// model.forward of the DeepFM graph on independent rows (each row brings its own history): dmg_score_pairs.
static __global__ void __launch_bounds__(128) deepfm_rows_forward_kernel(const float *__restrict__ emb, const float *__restrict__ dense,
                                                                         int E, int T, int64_t n, const int32_t *__restrict__ node,
                                                                         const int32_t *__restrict__ seq, float *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int RB = 8;
    const int F = T + 1, HL = F + 1;
    float *sF = reinterpret_cast<float *>(smem_raw);     // RB x F x E: [item | history] per row
    float *sH = sF + (size_t)RB * F * E;                 // RB x HL
    const float *w1 = dense, *b1 = w1 + (size_t)F * F * E, *w2 = b1 + F, *b2 = w2 + F;
    const int tid = threadIdx.x;
    for (int64_t g0 = (int64_t)blockIdx.x * RB; g0 < n; g0 += (int64_t)gridDim.x * RB) {
        const int nr = (int)((n - g0) < RB ? (n - g0) : RB);
        for (int idx = tid; idx < nr * F * E; idx += 128) {
            const int k = idx % E, slot = (idx / E) % F, r = idx / (E * F);
            const int32_t c = slot == 0 ? node[g0 + r] : seq[(g0 + r) * T + slot - 1];
            sF[idx] = c < 0 ? 0.0f : emb[(size_t)c * E + k];
        }
        __syncthreads();
        for (int idx = tid; idx < nr * (F + 1); idx += 128) {
            const int r = idx % nr, c = idx / nr;
            const float *x = sF + (size_t)r * F * E;
            sH[r * HL + c] = deepfm_chain(x, x + E, w1, b1, c, E, T);
        }
        __syncthreads();
        if (tid < nr) {
            const float *x = sF + (size_t)tid * F * E;
            out[g0 + tid] = deepfm_finish(x, x + E, sH + tid * HL, sH[tid * HL + F], w2, b2[0], E, T);
        }
        __syncthreads();
    }
}

// ---- requester side: results (Recommender.scala:103-106 + recommendItems :18-38) -----------------------------
static __global__ void __launch_bounds__(kThreads) shard_final_topk_kernel(const int32_t *__restrict__ cand, const float *__restrict__ score,
                                                                           const int32_t *__restrict__ count, int cap, int capp, int topk,
                                                                           int leaf_level, int reached_leaf, const int32_t *__restrict__ leaf_item,
                                                                           const int64_t *__restrict__ cons_off, const int32_t *__restrict__ cons,
                                                                           int32_t *__restrict__ out_items, float *__restrict__ out_scores,
                                                                           int32_t *__restrict__ out_counts, const int32_t *__restrict__ beam_user = nullptr)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *sKey = reinterpret_cast<uint64_t *>(smem_raw);
    const int u = blockIdx.x, tid = threadIdx.x;
    if (beam_user) reached_leaf = (31 - __clz(beam_user[u])) <= leaf_level ? 1 : 0;
    const int n = count[u];
    const int32_t *uc = cand + (size_t)u * cap;
    const float *us = score + (size_t)u * cap;
    const int64_t leaf_start = ((int64_t)1 << leaf_level) - 1;
    int n2 = 2;
    while (n2 < n) n2 <<= 1;
    for (int i = tid; i < n2; i += kThreads) {
        uint64_t k = KeyOf<float>::lowest();
        if (i < n && reached_leaf) {
            const int64_t slot = (int64_t)uc[i] - leaf_start;
            const int32_t item = (slot >= 0 && slot < ((int64_t)1 << leaf_level)) ? __ldg(leaf_item + slot) : -1;
            bool keep = item >= 0;
            if (cons_off)
                for (int64_t q = cons_off[u]; q < cons_off[u + 1] && keep; q++) keep = __ldg(cons + q) != item;
            if (keep) k = KeyOf<float>::make(us[i], i);
        }
        sKey[i] = k;
    }
    __syncthreads();
    if (n > 0) bitonic_sort_desc(sKey, n2);
    for (int i = tid; i < topk; i += kThreads) {
        int32_t item = -1;
        float sc = 0.0f;
        if (i < n && !KeyOf<float>::is_lowest(sKey[i])) {
            const int pos = KeyOf<float>::pos(sKey[i]);
            item = __ldg(leaf_item + ((int64_t)uc[pos] - leaf_start));
            sc = us[pos];
        }
        out_items[(size_t)u * topk + i] = item;
        out_scores[(size_t)u * topk + i] = sc;
    }
    if (tid == 0) {
        int valid = 0;
        const int lim = n < topk ? n : topk;
        while (valid < lim && !KeyOf<float>::is_lowest(sKey[valid])) valid++;
        out_counts[u] = valid;
    }
}

}  // namespace dmg

int32_t dmg_tdm_ids_to_codes(dmg_handle_t h, const int32_t *d_ids, int64_t n, int use_mask, int32_t *d_codes, uint8_t *d_mask);   // capi.cu

void dmg_shard_free(dmg_handle_t h)
{
    ShardState *s = h->shard;
    if (!s) return;
    if (s->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(s->comm);
    cudaFree(s->buf.d);
    cudaFreeHost(s->buf.h);
    delete s;
    h->shard = nullptr;
}

int dmg_shard_world(dmg_handle_t h) { return h && h->shard ? h->shard->world : 1; }   // for translation units that do not see ShardState
void dmg_shard_copy_geometry(dmg_handle_t dst, dmg_handle_t src)                       // dmg_clone of a whole-table DeepFM engine
{
    if (dst && src && dst->shard && src->shard) dst->shard->global_rows = src->shard->global_rows;
}

DMG_API int32_t dmg_shard_unique_id(void *out, int32_t nbytes)
{
    if (!out || nbytes < (int32_t)sizeof(ncclUniqueId)) return DMG_ERR_INVALID_ARG;
    if (load_nccl()) return DMG_ERR_UNSUPPORTED;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return DMG_ERR_CUDA;
    memset(out, 0, (size_t)nbytes);
    memcpy(out, &id, sizeof(id));
    return DMG_OK;
}

DMG_API int32_t dmg_shard_init(dmg_handle_t h, int32_t world, int32_t rank, const void *unique_id)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    DMG_TRY(model_is_shared(h, "dmg_shard_init"));
    if (world < 1 || world > 32 || (world & (world - 1)) || rank < 0 || rank >= world)
        return fail(h, DMG_ERR_INVALID_ARG, "dmg_shard_init: world must be a power of two <= 32, 0 <= rank < world");
    if (world > 1 && !unique_id) return fail(h, DMG_ERR_INVALID_ARG, "dmg_shard_init: unique id required for world > 1");
    DMG_CUDA(h, cudaSetDevice(h->device));
    dmg_shard_free(h);
    ShardState *s = new ShardState();
    s->world = world; s->rank = rank;
    while ((1 << s->bits) < world) s->bits++;
    if (world > 1) {
        const char *e = load_nccl();
        if (e) { delete s; return fail(h, DMG_ERR_UNSUPPORTED, "dmg_shard_init: %s", e); }
        ncclUniqueId id;
        memcpy(&id, unique_id, sizeof(id));
        ncclResult_t r = g_nccl.CommInitRank(&s->comm, world, id, rank);
        if (r != ncclSuccess) { delete s; return fail(h, DMG_ERR_CUDA, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r)); }
    }
    h->shard = s;
    return DMG_OK;
}

static int64_t dense_params(int kind, int E, int T)
{
    return kind == 1 ? (int64_t)(T + 1) * (T + 1) * E + 2 * (T + 1) + 1 : (int64_t)3 * E * E + 2 * (int64_t)E + 1;
}

static int32_t shard_alloc_din(dmg_handle_t h, int64_t rows_global, int32_t E, int32_t T, int kind = 0)
{
    DMG_TRY(model_is_shared(h, "loading weights"));
    ShardState *s = h->shard;
    if (!s) return fail(h, DMG_ERR_STATE, "call dmg_shard_init first");
    if (!h->tree.loaded) return fail(h, DMG_ERR_STATE, "load the tree first (the shard layout follows its levels)");
    if (rows_global != h->tree.n_codes) return fail(h, DMG_ERR_INVALID_ARG, "sharded table must have 2^(max_level+1)-1 = %lld rows", (long long)h->tree.n_codes);
    if (h->tree.max_level < s->bits) return fail(h, DMG_ERR_INVALID_ARG, "tree has fewer levels than log2(world)");
    if (T > kMaxT || E % 4 != 0 || E > 256 || E <= 0 || T <= 0) return fail(h, DMG_ERR_UNSUPPORTED, "embed_size / seq_len not supported");
    DinDev &d = h->din;
    cudaFree(d.d_params); cudaFree(d.d_wattT); cudaFree(d.d_w1T); cudaFree(d.d_grad); cudaFree(d.d_m); cudaFree(d.d_v);
    d = DinDev();
    d.dtype = DMG_F32; d.esz = 4; d.E = E; d.T = T;
    d.rows = shard_local_rows(s->bits, h->tree.max_level);
    d.kind = kind;
    d.n_params = d.rows * E + dense_params(kind, E, T);
    s->global_rows = rows_global;
    DMG_CUDA(h, cudaMalloc(&d.d_params, (size_t)d.n_params * 4));
    DMG_CUDA(h, cudaMalloc(&d.d_wattT, sizeof(float) * E * E));
    DMG_CUDA(h, cudaMalloc(&d.d_w1T, sizeof(float) * 2 * E * E));
    return DMG_OK;
}

static int32_t shard_finish_din(dmg_handle_t h)
{
    DinDev &d = h->din;
    const int E = d.E;
    if (d.kind == 1) {                                           // DeepFM: no transposed copies, level-synchronous path only
        DMG_CUDA(h, cudaStreamSynchronize(h->stream));
        h->fast_dirty = true;                                    // the fast path's bound tables and weight image follow the new weights
        d.loaded = true;
        d.sharded = true;
        return DMG_OK;
    }
    transpose_kernel<float><<<(E * E + 255) / 256, 256, 0, h->stream>>>(d.watt<float>(), (float *)d.d_wattT, E, E);
    transpose_kernel<float><<<(2 * E * E + 255) / 256, 256, 0, h->stream>>>(d.w1<float>(), (float *)d.d_w1T, E, 2 * E);
    h->launches += 2;
    DMG_CUDA(h, cudaGetLastError());
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    h->fast_dirty = true;
    d.loaded = true;
    d.sharded = h->shard->world > 1;
    return DMG_OK;
}

DMG_API int32_t dmg_shard_init_din_weights(dmg_handle_t h, int64_t rows_global, int32_t E, int32_t T, uint64_t seed)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    DMG_CUDA(h, cudaSetDevice(h->device));
    DMG_TRY(shard_alloc_din(h, rows_global, E, T));
    DinDev &d = h->din;
    const ShardGeo g = h->shard->geo();
    DMG_CUDA(h, cudaMemsetAsync(d.d_params, 0, (size_t)d.n_params * 4, h->stream));
    shard_randn_rows_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(d.emb<float>(), g, d.rows, E, seed, 0.05);
    shard_randn_tail_kernel<<<64, 256, 0, h->stream>>>(d.watt<float>(), (int64_t)3 * E * E, (uint64_t)(rows_global * E), seed, 0.05);
    randn_fill_kernel<float><<<1, 256, 0, h->stream>>>(d.w2<float>(), E, seed ^ 0x5bd1e995u, 0.05);
    h->launches += 3;
    DMG_CUDA(h, cudaGetLastError());
    return shard_finish_din(h);
}

// params: the FULL compact vector of the unsharded model (Module.parameters(), rows_global rows); only this rank's
// rows are uploaded.
static int32_t shard_load_weights(dmg_handle_t h, int64_t rows_global, int32_t E, int32_t T, const float *params, int kind)
{
    if (!h || !params) return h ? fail(h, DMG_ERR_INVALID_ARG, "null params") : DMG_ERR_INVALID_ARG;
    DMG_CUDA(h, cudaSetDevice(h->device));
    DMG_TRY(shard_alloc_din(h, rows_global, E, T, kind));
    DinDev &d = h->din;
    const ShardGeo g = h->shard->geo();
    // replicated levels, then one contiguous block per level
    int64_t lr = 0;
    while (lr < d.rows) {
        const int64_t gr = shard_global_row(g, lr);
        int64_t n = lr < g.repl_rows ? g.repl_rows : (int64_t)1 << (code_level(gr) - g.bits);
        DMG_CUDA(h, cudaMemcpyAsync(d.emb<float>() + lr * E, params + gr * E, (size_t)n * E * 4, cudaMemcpyHostToDevice, h->stream));
        lr += n;
    }
    DMG_CUDA(h, cudaMemcpyAsync(d.tail<float>(), params + rows_global * E, (size_t)dense_params(kind, E, T) * 4, cudaMemcpyHostToDevice, h->stream));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    return shard_finish_din(h);
}

DMG_API int32_t dmg_shard_load_din_weights(dmg_handle_t h, int64_t rows_global, int32_t E, int32_t T, const float *params)
{
    return shard_load_weights(h, rows_global, E, T, params, 0);
}

// DeepFM (tdm/.../model/DeepFM.scala:11-44): params = [emb | W1 (T+1)x(T+1)E | b1 | W2 | b2].  Runs on the level-synchronous
// path of this file; without a prior dmg_shard_init the table stays whole on this device (world 1).
DMG_API int32_t dmg_load_deepfm_weights(dmg_handle_t h, int64_t rows, int32_t E, int32_t T, const float *params)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    if (!h->shard) DMG_TRY(dmg_shard_init(h, 1, 0, nullptr));
    return shard_load_weights(h, rows, E, T, params, 1);
}

DMG_API int32_t dmg_shard_info(dmg_handle_t h, int64_t *local_rows, int64_t *global_rows, int64_t *exchanged_rows)
{
    if (!h || !h->shard) return DMG_ERR_INVALID_ARG;
    if (local_rows) *local_rows = h->din.loaded ? h->din.rows : 0;
    if (global_rows) *global_rows = h->shard->global_rows;
    if (exchanged_rows) *exchanged_rows = h->shard->exchanged_rows;
    return DMG_OK;
}

// TDM.recommend for this rank's B users over the sharded table (every rank calls it with the same B, beam, topk).
// ---- the collective pieces shared by every sharded entry point ------------------------------------------------------
// B "users" per rank (a user = one history tile + up to cap candidate codes): TDM retrieval brings one per query and
// level, JTM one per (item sample, subtree depth).
struct ShardWork {
    int B = 0, cap = 0;
    int64_t stride = 0;                                         // B * cap
    int32_t *codes_all = nullptr; uint8_t *mask_all = nullptr; float *tiles = nullptr;      // [G*B][T], [G*B][T], [G*B][T][E]
    int32_t *cand = nullptr; float *score = nullptr; int32_t *count = nullptr;              // [B][cap], [B][cap], [B]
    int2 *req = nullptr, *rreq = nullptr; float *rsc = nullptr, *reply = nullptr;           // [G][stride]
    int32_t *nreq = nullptr, *matrix = nullptr; int2 *seg = nullptr, *rseg = nullptr; int32_t *work = nullptr;
    int32_t *h_matrix = nullptr;                                // pinned [G*G]
    int32_t *codes_mine(int rank, int T) const { return codes_all + (size_t)rank * B * T; }
    uint8_t *mask_mine(int rank, int T) const { return mask_all + (size_t)rank * B * T; }
};
static size_t shard_work_bytes(int G, int B, int cap, int T, int E)
{
    const size_t BU = (size_t)G * B, stride = (size_t)B * cap;
    return Carver::need({BU * T * 4, BU * T, BU * T * E * 4, stride * 4, stride * 4, (size_t)B * 4, G * stride * 8, G * stride * 8,
                         G * stride * 4, G * stride * 4, (size_t)G * 4, (size_t)G * G * 4, (size_t)G * B * 8, (size_t)G * B * 8, 256});
}
static void shard_work_carve(Carver &cd, ShardWork &w, int G, int B, int cap, int T, int E)
{
    const size_t BU = (size_t)G * B;
    w.B = B; w.cap = cap; w.stride = (int64_t)B * cap;
    w.codes_all = cd.take<int32_t>(BU * T);
    w.mask_all = cd.take<uint8_t>(BU * T);
    w.tiles = cd.take<float>(BU * T * E);
    w.cand = cd.take<int32_t>((size_t)w.stride);
    w.score = cd.take<float>((size_t)w.stride);
    w.count = cd.take<int32_t>((size_t)B);
    w.req = cd.take<int2>((size_t)G * w.stride);                // my requests, one region per owner
    w.rreq = cd.take<int2>((size_t)G * w.stride);               // requests received, one region per requester
    w.rsc = cd.take<float>((size_t)G * w.stride);               // scores I computed, per requester
    w.reply = cd.take<float>((size_t)G * w.stride);             // scores received, per owner
    w.nreq = cd.take<int32_t>((size_t)G);
    w.matrix = cd.take<int32_t>((size_t)G * G);
    w.seg = cd.take<int2>((size_t)G * B);                       // my segments, one table per owner
    w.rseg = cd.take<int2>((size_t)G * B);                      // segments received, one table per requester
    w.work = cd.take<int32_t>(64);
}

// codes / mask of this rank's users sit in their slice of codes_all / mask_all: replicate them and build every user's
// history tile on every rank (integer-sum all-reduce of the rows each rank owns).
static int32_t shard_history_tiles(dmg_handle_t h, const ShardWork &w)
{
    ShardState *s = h->shard;
    const DinDev &d = h->din;
    const int G = s->world, T = d.T, E = d.E;
    const int64_t BU = (int64_t)G * w.B;
    cudaStream_t st = h->stream;
    if (G > 1) {
        DMG_NCCL(h, g_nccl.AllGather(w.codes_mine(s->rank, T), w.codes_all, (size_t)w.B * T, ncclInt32, s->comm, st));
        DMG_NCCL(h, g_nccl.AllGather(w.mask_mine(s->rank, T), w.mask_all, (size_t)w.B * T, ncclUint8, s->comm, st));
    }
    shard_fill_tiles_kernel<<<h->sm_count * 4, 256, 0, st>>>(d.emb<float>(), s->geo(), w.codes_all, BU * T, E, (uint32_t *)w.tiles);
    h->launches += 1;
    if (G > 1) DMG_NCCL(h, g_nccl.AllReduce(w.tiles, w.tiles, (size_t)BU * T * E, ncclUint32, ncclSum, s->comm, st));
    DMG_CUDA(h, cudaGetLastError());
    return DMG_OK;
}

static int32_t shard_prepare_scorers(dmg_handle_t h)
{
    const DinDev &d = h->din;
    const int E = d.E, T = d.T;
    const size_t row_smem = (size_t)kRowsRB * ((size_t)4 * E + (size_t)T * E + T + 1) * 4;
    DMG_CUDA(h, cudaFuncSetAttribute(shard_score_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_smem));
    if (d.kind == 1) {
        const size_t dfm_smem = DfmGeo::smem(E, T);
        if (dfm_smem > h->smem_optin) return fail(h, DMG_ERR_UNSUPPORTED, "DeepFM weights (%zu B) do not fit shared memory", dfm_smem);
        DMG_CUDA(h, cudaFuncSetAttribute(shard_score_deepfm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dfm_smem));
    } else {
        switch (E) {
        case 16: DMG_CUDA(h, cudaFuncSetAttribute(shard_score_segments_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shard_score_smem<16>())); break;
        case 32: DMG_CUDA(h, cudaFuncSetAttribute(shard_score_segments_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shard_score_smem<32>())); break;
        case 64: DMG_CUDA(h, cudaFuncSetAttribute(shard_score_segments_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shard_score_smem<64>())); break;
        default: break;
        }
    }
    return DMG_OK;
}

// cand / count -> score: requests to the owners, scores back, scattered into candidate order.  One host synchronisation
// (the G x G count matrix sizes the sends).
static int32_t shard_score_level(dmg_handle_t h, const ShardWork &w)
{
    ShardState *s = h->shard;
    const DinDev &d = h->din;
    const int G = s->world, T = d.T, E = d.E, B = w.B, cap = w.cap;
    const int64_t stride = w.stride;
    const ShardGeo geo = s->geo();
    cudaStream_t st = h->stream;
    const bool deepfm = d.kind == 1;
    const bool tiled = deepfm || E == 16 || E == 32 || E == 64;   // DIN: embedding sizes with a tile geometry that fits shared memory
    const float scale = (float)(1.0 / std::sqrt((double)E));
    int32_t *h_matrix = w.h_matrix;
    DMG_CUDA(h, cudaMemsetAsync(w.nreq, 0, (size_t)G * 4, st));
    shard_bucket_kernel<<<B, kThreads, 0, st>>>(w.cand, w.count, cap, geo, w.req, stride, w.nreq, w.seg);
    h->launches += 1;
    if (G > 1) DMG_NCCL(h, g_nccl.AllGather(w.nreq, w.matrix, (size_t)G, ncclInt32, s->comm, st));
    else DMG_CUDA(h, cudaMemcpyAsync(w.matrix, w.nreq, 4, cudaMemcpyDeviceToDevice, st));
    DMG_CUDA(h, cudaMemcpyAsync(h_matrix, w.matrix, (size_t)G * G * 4, cudaMemcpyDeviceToHost, st));
    DMG_CUDA(h, cudaStreamSynchronize(st));
    // h_matrix[src*G + dst] = requests src sends to dst
    if (G > 1) {
        DMG_NCCL(h, g_nccl.GroupStart());
        for (int p = 0; p < G; p++) {
            if (p == s->rank) continue;
            const int ns = h_matrix[s->rank * G + p], nr = h_matrix[p * G + s->rank];
            if (ns) DMG_NCCL(h, g_nccl.Send(w.req + (size_t)p * stride, (size_t)ns * 2, ncclInt32, p, s->comm, st));
            if (nr) DMG_NCCL(h, g_nccl.Recv(w.rreq + (size_t)p * stride, (size_t)nr * 2, ncclInt32, p, s->comm, st));
            if (tiled && ns) DMG_NCCL(h, g_nccl.Send(w.seg + (size_t)p * B, (size_t)B * 2, ncclInt32, p, s->comm, st));
            if (tiled && nr) DMG_NCCL(h, g_nccl.Recv(w.rseg + (size_t)p * B, (size_t)B * 2, ncclInt32, p, s->comm, st));
        }
        DMG_NCCL(h, g_nccl.GroupEnd());
    }
    if (tiled) {
        for (int p = 0; p < G; p++)                              // a requester that sent nothing sent no table either
            if (p != s->rank && h_matrix[p * G + s->rank] == 0) DMG_CUDA(h, cudaMemsetAsync(w.rseg + (size_t)p * B, 0, (size_t)B * 8, st));
        DMG_CUDA(h, cudaMemsetAsync(w.work, 0, 4, st));
        const int grid = std::min(G * B, h->sm_count);
        if (deepfm) {
            ShardDfmArgs da;
            da.emb = d.emb<float>(); da.dense = d.tail<float>(); da.tiles = w.tiles;
            da.req_self = w.req; da.req_peer = w.rreq; da.seg_self = w.seg; da.seg_peer = w.rseg;
            da.out_self = w.reply; da.out_peer = w.rsc; da.stride = stride; da.B = B; da.G = G; da.T = T; da.E = E;
            da.work = w.work; da.geo = geo;
            shard_score_deepfm_kernel<<<grid, kThreads, DfmGeo::smem(E, T), st>>>(da);
        } else {
            ShardScoreArgs sa;
            sa.emb = d.emb<float>(); sa.wattT = (const float *)d.d_wattT; sa.w1T = (const float *)d.d_w1T;
            sa.b1 = d.b1<float>(); sa.w2 = d.w2<float>(); sa.b2 = d.b2<float>(); sa.tiles = w.tiles; sa.mask = w.mask_all;
            sa.req_self = w.req; sa.req_peer = w.rreq; sa.seg_self = w.seg; sa.seg_peer = w.rseg;
            sa.out_self = w.reply; sa.out_peer = w.rsc; sa.stride = stride; sa.B = B; sa.G = G; sa.T = T; sa.cap = cap;
            sa.scale = scale; sa.work = w.work; sa.geo = geo;
            switch (E) {
            case 16: shard_score_segments_kernel<16><<<grid, kThreads, shard_score_smem<16>(), st>>>(sa); break;
            case 32: shard_score_segments_kernel<32><<<grid, kThreads, shard_score_smem<32>(), st>>>(sa); break;
            default: shard_score_segments_kernel<64><<<grid, kThreads, shard_score_smem<64>(), st>>>(sa); break;
            }
        }
        h->launches += 1;
        for (int p = 0; p < G; p++)
            if (p != s->rank) s->exchanged_rows += h_matrix[p * G + s->rank];
    } else {
        const size_t row_smem = (size_t)kRowsRB * ((size_t)4 * E + (size_t)T * E + T + 1) * 4;
        for (int p = 0; p < G; p++) {
            const int nr = h_matrix[p * G + s->rank];
            if (!nr) continue;
            const int2 *rq = p == s->rank ? w.req + (size_t)p * stride : w.rreq + (size_t)p * stride;
            float *ro = p == s->rank ? w.reply + (size_t)p * stride : w.rsc + (size_t)p * stride;
            const int grid = std::min((nr + kRowsRB - 1) / kRowsRB, h->sm_count * 8);
            shard_score_rows_kernel<<<grid, kRowsThreads, row_smem, st>>>(d.emb<float>(), geo, (const float *)d.d_wattT, (const float *)d.d_w1T,
                                                                          d.b1<float>(), d.w2<float>(), d.b2<float>(), scale, E, T, nr, rq, cap,
                                                                          p * B, w.tiles, w.mask_all, ro);
            h->launches += 1;
            if (p != s->rank) s->exchanged_rows += nr;
        }
    }
    if (G > 1) {
        DMG_NCCL(h, g_nccl.GroupStart());
        for (int p = 0; p < G; p++) {
            if (p == s->rank) continue;
            const int ns = h_matrix[s->rank * G + p], nr = h_matrix[p * G + s->rank];
            if (nr) DMG_NCCL(h, g_nccl.Send(w.rsc + (size_t)p * stride, (size_t)nr, ncclFloat32, p, s->comm, st));
            if (ns) DMG_NCCL(h, g_nccl.Recv(w.reply + (size_t)p * stride, (size_t)ns, ncclFloat32, p, s->comm, st));
        }
        DMG_NCCL(h, g_nccl.GroupEnd());
    }
    for (int p = 0; p < G; p++) {
        const int ns = h_matrix[s->rank * G + p];
        if (!ns) continue;
        shard_scatter_kernel<<<(ns + 255) / 256, 256, 0, st>>>(w.req + (size_t)p * stride, w.reply + (size_t)p * stride, ns, w.score);
        h->launches += 1;
    }
    DMG_CUDA(h, cudaGetLastError());
    return DMG_OK;
}

static int32_t shard_tdm_retrieve_impl(dmg_handle_t h, int32_t B, const int32_t *item_seq, int32_t beam, int32_t topk,
                                       int32_t use_mask, const int64_t *cons_off, const int32_t *cons,
                                       int32_t *out_items, float *out_logits, int32_t *out_counts, const int32_t *beam_user = nullptr)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    ShardState *s = h->shard;
    if (!s) return fail(h, DMG_ERR_STATE, "call dmg_shard_init first");
    if (!h->tree.loaded || h->tree.complete || !h->din.loaded) return fail(h, DMG_ERR_STATE, "TDM tree and sharded DIN weights must be loaded first");
    if (B <= 0 || beam <= 0 || topk <= 0 || !item_seq || !out_items || !out_logits || !out_counts) return fail(h, DMG_ERR_INVALID_ARG, "bad arguments");
    DMG_CUDA(h, cudaSetDevice(h->device));
    const DinDev &d = h->din;
    const TreeDev &t = h->tree;
    const int G = s->world, T = d.T, E = d.E, L = t.max_level;
    int max_beam = beam;
    if (beam_user)
        for (int u = 0; u < B; u++) {
            if (beam_user[u] < beam) return fail(h, DMG_ERR_INVALID_ARG, "per-user beams only widen the configured beam");
            max_beam = std::max(max_beam, beam_user[u]);
        }
    const int cap = std::max(((2 * max_beam + 7) / 8) * 8, ((topk + 7) / 8) * 8);
    if (cap > 2 * kThreads) return fail(h, DMG_ERR_UNSUPPORTED, "beam %d too wide for the level-synchronous path (2*beam <= %d)", max_beam, 2 * kThreads);
    int capp = 2;
    while (capp < cap) capp <<= 1;
    const int s_level = (int)std::floor(std::log2((double)beam) + 1e-9);

    const size_t need = shard_work_bytes(G, B, cap, T, E) +
                        Carver::need({(size_t)B * T * 4, (size_t)B * T, (size_t)B * topk * 4, (size_t)B * topk * 4, (size_t)B * 4,
                                      cons_off ? (size_t)(B + 1) * 8 : 0, cons_off ? (size_t)cons_off[B] * 4 : 0, beam_user ? (size_t)B * 4 : 0});
    DMG_TRY(ensure_dev(h, s->buf, need));
    DMG_TRY(ensure_host(h, s->buf, (size_t)B * T * 4 + (size_t)G * G * 4 + (size_t)B * topk * 8 + (size_t)B * 4 + 1024));
    Carver cd(s->buf.d);
    ShardWork w;
    shard_work_carve(cd, w, G, B, cap, T, E);
    int32_t *d_seq = cd.take<int32_t>((size_t)B * T);
    uint8_t *d_mask = cd.take<uint8_t>((size_t)B * T);
    int32_t *d_items = cd.take<int32_t>((size_t)B * topk);
    float *d_logits = cd.take<float>((size_t)B * topk);
    int32_t *d_cnt_out = cd.take<int32_t>((size_t)B);
    int64_t *d_cons_off = cons_off ? cd.take<int64_t>((size_t)B + 1) : nullptr;
    int32_t *d_cons = cons_off ? cd.take<int32_t>((size_t)cons_off[B]) : nullptr;
    int32_t *d_beam_user = beam_user ? cd.take<int32_t>((size_t)B) : nullptr;
    char *hp = (char *)s->buf.h;
    int32_t *h_seq = (int32_t *)hp; hp += (size_t)B * T * 4;
    w.h_matrix = (int32_t *)hp; hp += (((size_t)G * G * 4 + 255) & ~(size_t)255);
    int32_t *h_items = (int32_t *)hp; hp += (size_t)B * topk * 4;
    float *h_logits = (float *)hp; hp += (size_t)B * topk * 4;
    int32_t *h_cnt = (int32_t *)hp;
    cudaStream_t st = h->stream;

    // ---- K2: ids -> codes + mask, replicated to every rank; history tiles by integer all-reduce ----------------
    memcpy(h_seq, item_seq, (size_t)B * T * 4);
    DMG_CUDA(h, cudaMemcpyAsync(d_seq, h_seq, (size_t)B * T * 4, cudaMemcpyHostToDevice, st));
    if (cons_off) {                                              // small, once per call: straight from the caller's arrays, waited for below
        DMG_CUDA(h, cudaMemcpyAsync(d_cons_off, cons_off, (size_t)(B + 1) * 8, cudaMemcpyHostToDevice, st));
        if (cons_off[B]) DMG_CUDA(h, cudaMemcpyAsync(d_cons, cons, (size_t)cons_off[B] * 4, cudaMemcpyHostToDevice, st));
        if (beam_user) DMG_CUDA(h, cudaMemcpyAsync(d_beam_user, beam_user, (size_t)B * 4, cudaMemcpyHostToDevice, st));
        DMG_CUDA(h, cudaStreamSynchronize(st));
    }
    // TDMTree.idToCode validates against the table size: use the global row count here
    {
        const int64_t local_rows = h->din.rows;
        h->din.rows = s->global_rows;
        const int32_t rc = dmg_tdm_ids_to_codes(h, d_seq, (int64_t)B * T, use_mask, w.codes_mine(s->rank, T), d_mask);
        h->din.rows = local_rows;
        DMG_TRY(rc);
    }
    DMG_CUDA(h, cudaMemcpyAsync(w.mask_mine(s->rank, T), d_mask, (size_t)B * T, cudaMemcpyDeviceToDevice, st));
    DMG_TRY(shard_history_tiles(h, w));

    // ---- level loop -----------------------------------------------------------------------------------------------
    const size_t sel_smem = (size_t)capp * 8 + (size_t)cap * 8;
    DMG_TRY(shard_prepare_scorers(h));
    const bool reached = s_level <= L;
    if (d_beam_user) {                                           // every user from its own start level: one launch per level serves them all
        DMG_CUDA(h, cudaMemsetAsync(w.count, 0, (size_t)B * 4, st));
        for (int level = s_level; level <= L; level++) {
            shard_select_expand_kernel<<<B, kThreads, sel_smem, st>>>(w.cand, w.score, w.count, cap, capp, beam, level, 0, t.d_exists, d_beam_user, L);
            h->launches += 1;
            if (level < L) DMG_TRY(shard_score_level(h, w));
        }
    } else {
        if (reached) {
            shard_select_expand_kernel<<<B, kThreads, sel_smem, st>>>(w.cand, w.score, w.count, cap, capp, beam, s_level, 1, t.d_exists);
            h->launches += 1;
        } else {
            DMG_CUDA(h, cudaMemsetAsync(w.count, 0, (size_t)B * 4, st));
        }
        for (int level = s_level; level < L; level++) {
            shard_select_expand_kernel<<<B, kThreads, sel_smem, st>>>(w.cand, w.score, w.count, cap, capp, beam, level, 0, t.d_exists);
            h->launches += 1;
            DMG_TRY(shard_score_level(h, w));
        }
    }
    shard_final_topk_kernel<<<B, kThreads, (size_t)capp * 8, st>>>(w.cand, w.score, w.count, cap, capp, topk, L, reached ? 1 : 0, t.d_leaf_item,
                                                                   d_cons_off, d_cons, d_items, d_logits, d_cnt_out, d_beam_user);
    h->launches += 1;
    DMG_CUDA(h, cudaGetLastError());
    DMG_CUDA(h, cudaMemcpyAsync(h_items, d_items, (size_t)B * topk * 4, cudaMemcpyDeviceToHost, st));
    DMG_CUDA(h, cudaMemcpyAsync(h_logits, d_logits, (size_t)B * topk * 4, cudaMemcpyDeviceToHost, st));
    DMG_CUDA(h, cudaMemcpyAsync(h_cnt, d_cnt_out, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    DMG_CUDA(h, cudaStreamSynchronize(st));
    memcpy(out_items, h_items, (size_t)B * topk * 4);
    memcpy(out_logits, h_logits, (size_t)B * topk * 4);
    memcpy(out_counts, h_cnt, (size_t)B * 4);
    return DMG_OK;
}

DMG_API int32_t dmg_shard_tdm_retrieve(dmg_handle_t h, int32_t B, const int32_t *item_seq, int32_t beam, int32_t topk,
                                       int32_t use_mask, int32_t *out_items, float *out_logits, int32_t *out_counts)
{
    return shard_tdm_retrieve_impl(h, B, item_seq, beam, topk, use_mask, nullptr, nullptr, out_items, out_logits, out_counts);
}

// dmg_tdm_retrieve / dmg_score_pairs with a DeepFM model loaded (called from capi.cu)
int32_t dmg_deepfm_tdm_retrieve(dmg_handle_t h, int32_t B, const int32_t *item_seq, int32_t beam, int32_t topk,
                                const int64_t *cons_off, const int32_t *cons, int32_t widen_beam,
                                int32_t *out_items, float *out_logits, int32_t *out_counts)
{
    if (h->shard && h->shard->world > 1 && cons_off) return fail(h, DMG_ERR_UNSUPPORTED, "consumed items with a sharded table");
    std::vector<int32_t> bu;
    if (cons_off && widen_beam) {                                // Recommender.recommendItems widens for every model (Recommender.scala:27-33)
        bu.resize(B);
        for (int u = 0; u < B; u++) bu[u] = std::max((int32_t)((cons_off[u + 1] - cons_off[u] + topk) / 2), beam);
    }
    return shard_tdm_retrieve_impl(h, B, item_seq, beam, topk, 0, cons_off, cons, out_items, out_logits, out_counts, bu.empty() ? nullptr : bu.data());
}

int32_t dmg_deepfm_score_pairs(dmg_handle_t h, int64_t n, const int32_t *node, const int32_t *seq, float *out)
{
    const DinDev &d = h->din;
    if (h->shard && h->shard->world > 1) return fail(h, DMG_ERR_UNSUPPORTED, "dmg_score_pairs on a sharded table");
    if (n <= 0) return DMG_OK;
    DMG_CUDA(h, cudaSetDevice(h->device));
    const int T = d.T, E = d.E, F = T + 1;
    for (int64_t i = 0; i < n * (T + 1); i++) {
        const int32_t c = i < n ? node[i] : seq[i - n];
        if (c < -1 || c >= d.rows) return fail(h, DMG_ERR_INDEX, "dmg_score_pairs: embeddingLookup failed, index outside [0, %lld)", (long long)d.rows);
    }
    DMG_TRY(ensure_dev(h, h->s_in, Carver::need({(size_t)n * 4, (size_t)n * T * 4, (size_t)n * 4})));
    Carver cd(h->s_in.d);
    int32_t *dn = cd.take<int32_t>((size_t)n), *ds = cd.take<int32_t>((size_t)n * T);
    float *dout = cd.take<float>((size_t)n);
    DMG_CUDA(h, cudaMemcpyAsync(dn, node, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
    DMG_CUDA(h, cudaMemcpyAsync(ds, seq, (size_t)n * T * 4, cudaMemcpyHostToDevice, h->stream));
    const size_t smem = ((size_t)8 * F * E + (size_t)8 * (F + 1)) * 4;
    DMG_CUDA(h, cudaFuncSetAttribute(deepfm_rows_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<int64_t>((n + 7) / 8, (int64_t)h->sm_count * 8);
    deepfm_rows_forward_kernel<<<grid, 128, smem, h->stream>>>(d.emb<float>(), d.tail<float>(), E, T, n, dn, ds, dout);
    h->launches += 1;
    DMG_CUDA(h, cudaGetLastError());
    DMG_CUDA(h, cudaMemcpyAsync(out, dout, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    return DMG_OK;
}

// ---- JTM tree learning on the sharded table (BASELINE config 4) -----------------------------------------------------
// dmg_jtm_item_weights with the node table split over the ranks: every rank brings ITS OWN items (any split of the
// catalogue), the scorer rows travel to the owners of the nodes exactly like retrieval candidates.  A "user" of the
// exchange is (item sample, depth d below the item's current node): its history is the sample re-coded for level
// old_level + d (JTMTree.idToCodeWithMask, jtm/.../tree/JTMTree.scala:86-113), its candidates the 2^d descendants at that
// depth.  Items are processed in chunks of whole items; chunk shape and count are agreed over NCCL (max over ranks).
static __global__ void shard_jtm_users_kernel(int n_samples, int gap, int T, int cap, const int32_t *__restrict__ sample_seq,
                                              const int32_t *__restrict__ sample_parent, int old_level,
                                              const int32_t *__restrict__ id_code, int32_t non_leaf_offset, int32_t max_code,
                                              int hierarchical, int min_level, int use_mask, int n_users_total,
                                              int32_t *__restrict__ codes, uint8_t *__restrict__ mask, int32_t *__restrict__ cand,
                                              int32_t *__restrict__ count)
{
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < n_users_total; u += gridDim.x * blockDim.x) {
        const int sl = u / gap, d = u % gap + 1;
        if (sl >= n_samples) {                                      // padding users of a short chunk
            count[u] = 0;
            for (int j = 0; j < T; j++) { codes[(size_t)u * T + j] = -1; mask[(size_t)u * T + j] = 0; }
            continue;
        }
        const int level = old_level + d;
        for (int j = 0; j < T; j++) {                               // JTMTree.idToCodeWithMask :86-113
            const int32_t id = sample_seq[(size_t)sl * T + j];
            int32_t c;
            uint8_t m = 0;
            if (id == 0) { c = -1; m = 1; }
            else if (id > 0 && id < non_leaf_offset && id_code[id] >= 0) {
                c = id_code[id];
                if (hierarchical && level >= min_level) {           // getAncestorAtLevel :36-43
                    const int64_t lim = ((int64_t)1 << (level + 1)) - 1;
                    int64_t cc = c;
                    while (cc >= lim) cc = (cc - 1) >> 1;
                    c = (int32_t)cc;
                }
            } else {
                const int64_t tmp = (int64_t)id - non_leaf_offset;
                c = tmp > max_code ? -1 : (int32_t)tmp;             // NB: not added to the mask (JTMTree.scala:104-107)
            }
            codes[(size_t)u * T + j] = c;
            mask[(size_t)u * T + j] = use_mask ? m : 0;
        }
        const int64_t first = ((int64_t)sample_parent[sl] + 1) * ((int64_t)1 << d) - 1;   // leftmost descendant at depth d
        const int n = 1 << d;
        for (int j = 0; j < n; j++) cand[(size_t)u * cap + j] = (int32_t)(first + j);
        count[u] = n;
    }
}

// node sums over an item's samples in order (Tensor.sum), then per child the in-order sum along child -> parent
// (TreeLearning.scala:163-172); -1e6 for an item without samples (:160).
static __global__ void shard_jtm_reduce_kernel(int n_items, const int32_t *__restrict__ item_off /*samples of the chunk*/, int gap, int cap,
                                               const float *__restrict__ score, float *__restrict__ out_weights)
{
    const int n_child = 1 << gap;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < n_items * n_child; g += gridDim.x * blockDim.x) {
        const int item = g / n_child, ch = g % n_child;
        const int s0 = item_off[item], s1 = item_off[item + 1];
        if (s1 == s0) { out_weights[g] = -1e6f; continue; }
        int idx = n_child - 1 + ch;                                 // heap index inside the parent's subtree, root = 0
        float w = 0.0f;
        while (idx > 0) {
            const int depth = 31 - __clz(idx + 1), j = idx + 1 - (1 << depth);
            float acc = 0.0f;
            for (int sm = s0; sm < s1; sm++) acc = __fadd_rn(acc, score[((size_t)sm * gap + depth - 1) * cap + j]);
            w = __fadd_rn(w, acc);
            idx = (idx - 1) >> 1;
        }
        out_weights[g] = w;
    }
}

DMG_API int32_t dmg_shard_jtm_item_weights(dmg_handle_t h, int32_t n_items, const int64_t *sample_off, const int32_t *sample_seq,
                                           const int32_t *parent_code, int32_t old_level, int32_t level, int32_t hierarchical,
                                           int32_t min_level, int32_t use_mask, float *out_weights)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    ShardState *s = h->shard;
    if (!s) return fail(h, DMG_ERR_STATE, "call dmg_shard_init first");
    if (!h->tree.loaded || h->tree.complete || !h->din.loaded || h->din.kind != 0)
        return fail(h, DMG_ERR_STATE, "TDM/JTM tree and sharded DIN weights must be loaded first");
    const int gap = level - old_level;
    if (n_items < 0 || gap < 1 || gap > 8 || old_level < 0 || level > h->tree.max_level || (n_items > 0 && (!sample_off || !parent_code || !out_weights)))
        return fail(h, DMG_ERR_INVALID_ARG, "dmg_shard_jtm_item_weights: bad arguments (1 <= level - old_level <= 8, level <= max_level)");
    DMG_CUDA(h, cudaSetDevice(h->device));
    const DinDev &d = h->din;
    const TreeDev &t = h->tree;
    const int G = s->world, T = d.T, E = d.E, cap = 1 << gap, n_child = 1 << gap;
    cudaStream_t st = h->stream;
    for (int32_t i = 0; i < n_items; i++) {
        const int64_t pc = parent_code[i];
        if (pc < ((int64_t)1 << old_level) - 1 || pc > ((int64_t)2 << old_level) - 2)
            return fail(h, DMG_ERR_INVALID_ARG, "parent_code[%d] = %lld is not a node of level %d", i, (long long)pc, old_level);
    }
    // chunk shape: whole items, at most ns_cap samples per chunk; agreed over the ranks
    int64_t max_item = 0;
    for (int32_t i = 0; i < n_items; i++) max_item = std::max(max_item, sample_off[i + 1] - sample_off[i]);
    int32_t agree[2] = {(int32_t)std::max<int64_t>(max_item, std::max(1, 8192 / gap)), 0};
    auto pack = [&](int ns_cap, std::vector<int32_t> &starts) {
        starts.assign(1, 0);
        int64_t used = 0, items = 0;                              // a chunk holds at most ns_cap samples and ns_cap items
        for (int32_t i = 0; i < n_items; i++) {
            const int64_t ns = sample_off[i + 1] - sample_off[i];
            if ((used + ns > ns_cap && used > 0) || items >= ns_cap) { starts.push_back(i); used = 0; items = 0; }
            used += ns;
            items++;
        }
        if (starts.back() != n_items) starts.push_back(n_items);
    };
    std::vector<int32_t> starts;
    int32_t *d_agree = nullptr;
    if (G > 1) {
        DMG_TRY(ensure_dev(h, s->buf, 256));
        d_agree = (int32_t *)s->buf.d;
        DMG_CUDA(h, cudaMemcpyAsync(d_agree, agree, 4, cudaMemcpyHostToDevice, st));
        DMG_NCCL(h, g_nccl.AllReduce(d_agree, d_agree, 1, ncclInt32, ncclMax, s->comm, st));
        DMG_CUDA(h, cudaMemcpyAsync(agree, d_agree, 4, cudaMemcpyDeviceToHost, st));
        DMG_CUDA(h, cudaStreamSynchronize(st));
    }
    const int ns_cap = agree[0];
    pack(ns_cap, starts);
    int32_t n_chunks = (int32_t)starts.size() - 1;
    if (G > 1) {
        DMG_CUDA(h, cudaMemcpyAsync(d_agree, &n_chunks, 4, cudaMemcpyHostToDevice, st));
        DMG_NCCL(h, g_nccl.AllReduce(d_agree, d_agree, 1, ncclInt32, ncclMax, s->comm, st));
        DMG_CUDA(h, cudaMemcpyAsync(&n_chunks, d_agree, 4, cudaMemcpyDeviceToHost, st));
        DMG_CUDA(h, cudaStreamSynchronize(st));
    }
    const int B = ns_cap * gap;                                   // users per rank and chunk
    const size_t need = shard_work_bytes(G, B, cap, T, E) +
                        Carver::need({(size_t)ns_cap * T * 4, (size_t)ns_cap * 4, ((size_t)ns_cap + 2) * 4, (size_t)ns_cap * n_child * 4});
    DMG_TRY(ensure_dev(h, s->buf, need));
    DMG_TRY(ensure_host(h, s->buf, (size_t)G * G * 4 + 256 + (size_t)ns_cap * T * 4 + (size_t)ns_cap * 4 + ((size_t)ns_cap + 2) * 4 + (size_t)ns_cap * n_child * 4 + 1024));
    Carver cd(s->buf.d);
    ShardWork w;
    shard_work_carve(cd, w, G, B, cap, T, E);
    int32_t *d_sseq = cd.take<int32_t>((size_t)ns_cap * T);
    int32_t *d_spar = cd.take<int32_t>((size_t)ns_cap);
    int32_t *d_ioff = cd.take<int32_t>((size_t)ns_cap + 2);
    float *d_wout = cd.take<float>((size_t)ns_cap * n_child);
    char *hp = (char *)s->buf.h;
    w.h_matrix = (int32_t *)hp; hp += (((size_t)G * G * 4 + 255) & ~(size_t)255);
    int32_t *h_sseq = (int32_t *)hp; hp += (((size_t)ns_cap * T * 4 + 255) & ~(size_t)255);
    int32_t *h_spar = (int32_t *)hp; hp += (((size_t)ns_cap * 4 + 255) & ~(size_t)255);
    int32_t *h_ioff = (int32_t *)hp; hp += ((((size_t)ns_cap + 2) * 4 + 255) & ~(size_t)255);
    float *h_wout = (float *)hp;
    DMG_TRY(shard_prepare_scorers(h));
    for (int32_t c = 0; c < n_chunks; c++) {
        const int32_t i0 = c + 1 < (int32_t)starts.size() ? starts[c] : n_items, i1 = c + 1 < (int32_t)starts.size() ? starts[c + 1] : n_items;
        const int32_t ni = i1 - i0;
        const int64_t sb = ni > 0 ? sample_off[i0] : 0;
        const int32_t ns = ni > 0 ? (int32_t)(sample_off[i1] - sb) : 0;
        if (ns > 0 && !sample_seq) return fail(h, DMG_ERR_INVALID_ARG, "sample_seq is null");
        if (ns > 0) memcpy(h_sseq, sample_seq + sb * T, (size_t)ns * T * 4);
        h_ioff[0] = 0;
        for (int32_t i = 0; i < ni; i++) {
            h_ioff[i + 1] = (int32_t)(sample_off[i0 + i + 1] - sb);
            for (int32_t q = h_ioff[i]; q < h_ioff[i + 1]; q++) h_spar[q] = parent_code[i0 + i];
        }
        if (ns > 0) {
            DMG_CUDA(h, cudaMemcpyAsync(d_sseq, h_sseq, (size_t)ns * T * 4, cudaMemcpyHostToDevice, st));
            DMG_CUDA(h, cudaMemcpyAsync(d_spar, h_spar, (size_t)ns * 4, cudaMemcpyHostToDevice, st));
        }
        DMG_CUDA(h, cudaMemcpyAsync(d_ioff, h_ioff, ((size_t)ni + 1) * 4, cudaMemcpyHostToDevice, st));
        shard_jtm_users_kernel<<<std::max(1, std::min((B + 255) / 256, h->sm_count * 8)), 256, 0, st>>>(
            ns, gap, T, cap, d_sseq, d_spar, old_level, t.d_id_code, t.non_leaf_offset, t.max_code, hierarchical, min_level, use_mask, B,
            w.codes_mine(s->rank, T), w.mask_mine(s->rank, T), w.cand, w.count);
        h->launches += 1;
        DMG_TRY(shard_history_tiles(h, w));
        DMG_TRY(shard_score_level(h, w));
        if (ni > 0) {
            shard_jtm_reduce_kernel<<<std::min((ni * n_child + 255) / 256, h->sm_count * 8), 256, 0, st>>>(ni, d_ioff, gap, cap, w.score, d_wout);
            h->launches += 1;
            DMG_CUDA(h, cudaMemcpyAsync(h_wout, d_wout, (size_t)ni * n_child * 4, cudaMemcpyDeviceToHost, st));
            DMG_CUDA(h, cudaStreamSynchronize(st));
            memcpy(out_weights + (size_t)i0 * n_child, h_wout, (size_t)ni * n_child * 4);
        }
        DMG_CUDA(h, cudaGetLastError());
    }
    return DMG_OK;
}
