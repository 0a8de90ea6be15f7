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
#include <dlfcn.h>
#include <nccl.h>

#include "beam_kernels.cuh"
#include "rows_kernels.cuh"

using namespace dmg;

namespace dmg {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static const char *load_nccl()
{
    if (g_nccl.lib) return nullptr;
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return "libnccl.so.2 not found (dlopen)";
#define DMG_NCCL_SYM(field, name)                                                   \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(lib, name));      \
    if (!g_nccl.field) return "libnccl is missing " name;
    DMG_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    DMG_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    DMG_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    DMG_NCCL_SYM(AllGather, "ncclAllGather")
    DMG_NCCL_SYM(AllReduce, "ncclAllReduce")
    DMG_NCCL_SYM(Send, "ncclSend")
    DMG_NCCL_SYM(Recv, "ncclRecv")
    DMG_NCCL_SYM(GroupStart, "ncclGroupStart")
    DMG_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    DMG_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef DMG_NCCL_SYM
    g_nccl.lib = lib;
    return nullptr;
}

#define DMG_NCCL(h, expr)                                                                              \
    do {                                                                                               \
        ncclResult_t r_ = (expr);                                                                      \
        if (r_ != ncclSuccess)                                                                         \
            return dmg::fail(h, DMG_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, g_nccl.GetErrorString(r_), \
                             __FILE__, __LINE__);                                                      \
    } while (0)

struct ShardGeo {
    int bits, rank, world;
    int64_t repl_rows;            // 2^bits - 1 replicated rows (levels < bits)
};

struct ShardState {
    int world = 1, rank = 0, bits = 0;
    ncclComm_t comm = nullptr;
    int64_t global_rows = 0;
    int64_t exchanged_rows = 0;   // candidates scored for another rank (statistics)
    Scratch buf;                  // level-loop buffers
    ShardGeo geo() const { ShardGeo g; g.bits = bits; g.rank = rank; g.world = world; g.repl_rows = ((int64_t)1 << bits) - 1; return g; }
};

__host__ __device__ __forceinline__ int code_level(int64_t c)
{
#ifdef __CUDA_ARCH__
    return 63 - __clzll((unsigned long long)(c + 1));
#else
    int l = 0;
    while (((int64_t)2 << l) <= c + 1) l++;
    return l;
#endif
}
// owner of code c; replicated levels report `self`
__host__ __device__ __forceinline__ int shard_owner(const ShardGeo &g, int64_t c)
{
    const int l = code_level(c);
    if (l < g.bits) return g.rank;
    return (int)((c - (((int64_t)1 << l) - 1)) >> (l - g.bits));
}
// row of code c in the local table of its owner (or of anyone, on a replicated level)
__host__ __device__ __forceinline__ int64_t shard_local_row(const ShardGeo &g, int64_t c)
{
    const int l = code_level(c);
    if (l < g.bits) return c;
    const int sh = l - g.bits;
    return g.repl_rows + (((int64_t)1 << sh) - 1) + ((c - (((int64_t)1 << l) - 1)) & (((int64_t)1 << sh) - 1));
}
// inverse: global code of local row lr on rank g.rank
__host__ __device__ __forceinline__ int64_t shard_global_row(const ShardGeo &g, int64_t lr)
{
    if (lr < g.repl_rows) return lr;
    const int64_t t = lr - g.repl_rows + 1;
    const int sh = code_level(t - 1);                       // floor(log2 t)
    const int l = g.bits + sh;
    return (((int64_t)1 << l) - 1) + ((int64_t)g.rank << sh) + (t - ((int64_t)1 << sh));
}
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
                                                                              int level, int first, const uint32_t *__restrict__ exists)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *sKey = reinterpret_cast<uint64_t *>(smem_raw);
    int32_t *sCur = reinterpret_cast<int32_t *>(sKey + capp);
    int32_t *sNxt = sCur + cap;
    __shared__ int sScan[8];
    const int u = blockIdx.x, tid = threadIdx.x;
    int32_t *uc = cand + (size_t)u * cap;
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
        return;
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
    static size_t smem(int E, int T) { const int F = T + 1; return ((size_t)F * F * E + 2 * F + 4 + (size_t)kMaxT * E + (size_t)R * (E + 1) + (size_t)R * (F + 1)) * 4 + 32 * 4; }
};
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
    const int E = a.E, T = a.T, F = T + 1, LD = E + 1, HL = F + 1;
    float *sW1 = reinterpret_cast<float *>(smem_raw);
    float *sB1 = sW1 + (size_t)F * F * E;
    float *sW2 = sB1 + F;
    float *sB2 = sW2 + F;                        // 4 floats
    float *sK = sB2 + 4;                         // kMaxT x E
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
            for (int idx = tid; idx < nrows * (F + 1); idx += kThreads) {
                const int r = idx % nrows, c = idx / nrows;
                sH[r * HL + c] = deepfm_chain(sX + r * LD, sK, sW1, sB1, c, E, T);
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
                                                                           int32_t *__restrict__ out_counts)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *sKey = reinterpret_cast<uint64_t *>(smem_raw);
    const int u = blockIdx.x, tid = threadIdx.x;
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
static int32_t shard_tdm_retrieve_impl(dmg_handle_t h, int32_t B, const int32_t *item_seq, int32_t beam, int32_t topk,
                                       int32_t use_mask, const int64_t *cons_off, const int32_t *cons,
                                       int32_t *out_items, float *out_logits, int32_t *out_counts)
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
    const ShardGeo geo = s->geo();
    const int cap = std::max(((2 * beam + 7) / 8) * 8, ((topk + 7) / 8) * 8);
    if (cap > 2 * kThreads) return fail(h, DMG_ERR_UNSUPPORTED, "beam %d too wide for the sharded path (2*beam <= %d)", beam, 2 * kThreads);
    int capp = 2;
    while (capp < cap) capp <<= 1;
    const int s_level = (int)std::floor(std::log2((double)beam) + 1e-9);
    const int64_t BU = (int64_t)G * B, stride = (int64_t)B * cap;

    const size_t need = Carver::need({(size_t)B * T * 4, (size_t)BU * T * 4, (size_t)B * T, (size_t)BU * T, (size_t)BU * T * E * 4,
                                      (size_t)stride * 4, (size_t)stride * 4, (size_t)B * 4, (size_t)G * stride * 8, (size_t)G * stride * 8,
                                      (size_t)G * stride * 4, (size_t)G * stride * 4, (size_t)G * 4, (size_t)G * G * 4,
                                      (size_t)B * topk * 4, (size_t)B * topk * 4, (size_t)B * 4, (size_t)G * B * 8, (size_t)G * B * 8, 256, cons_off ? (size_t)(B + 1) * 8 : 0, cons_off ? (size_t)cons_off[B] * 4 : 0});
    DMG_TRY(ensure_dev(h, s->buf, need));
    DMG_TRY(ensure_host(h, s->buf, (size_t)B * T * 4 + (size_t)G * G * 4 + (size_t)B * topk * 8 + (size_t)B * 4 + 1024));
    Carver cd(s->buf.d);
    int32_t *d_seq = cd.take<int32_t>((size_t)B * T);
    int32_t *d_codes_all = cd.take<int32_t>((size_t)BU * T);
    uint8_t *d_mask = cd.take<uint8_t>((size_t)B * T);
    uint8_t *d_mask_all = cd.take<uint8_t>((size_t)BU * T);
    float *d_tiles = cd.take<float>((size_t)BU * T * E);
    int32_t *d_cand = cd.take<int32_t>((size_t)stride);
    float *d_score = cd.take<float>((size_t)stride);
    int32_t *d_count = cd.take<int32_t>((size_t)B);
    int2 *d_req = cd.take<int2>((size_t)G * stride);          // my requests, one region per owner
    int2 *d_rreq = cd.take<int2>((size_t)G * stride);         // requests received, one region per requester
    float *d_rsc = cd.take<float>((size_t)G * stride);        // scores I computed, per requester
    float *d_reply = cd.take<float>((size_t)G * stride);      // scores received, per owner
    int32_t *d_nreq = cd.take<int32_t>((size_t)G);
    int32_t *d_matrix = cd.take<int32_t>((size_t)G * G);
    int32_t *d_items = cd.take<int32_t>((size_t)B * topk);
    float *d_logits = cd.take<float>((size_t)B * topk);
    int32_t *d_cnt_out = cd.take<int32_t>((size_t)B);
    int2 *d_seg = cd.take<int2>((size_t)G * B);               // my segments, one table per owner
    int2 *d_rseg = cd.take<int2>((size_t)G * B);              // segments received, one table per requester
    int32_t *d_work = cd.take<int32_t>(64);
    int64_t *d_cons_off = cons_off ? cd.take<int64_t>((size_t)B + 1) : nullptr;
    int32_t *d_cons = cons_off ? cd.take<int32_t>((size_t)cons_off[B]) : nullptr;
    char *hp = (char *)s->buf.h;
    int32_t *h_seq = (int32_t *)hp; hp += (size_t)B * T * 4;
    int32_t *h_matrix = (int32_t *)hp; hp += (((size_t)G * G * 4 + 255) & ~(size_t)255);
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
        DMG_CUDA(h, cudaStreamSynchronize(st));
    }
    int32_t *d_codes_mine = d_codes_all + (size_t)s->rank * B * T;
    uint8_t *d_mask_mine = d_mask_all + (size_t)s->rank * B * T;
    // TDMTree.idToCode validates against the table size: use the global row count here
    {
        const int64_t local_rows = h->din.rows;
        h->din.rows = s->global_rows;
        const int32_t rc = dmg_tdm_ids_to_codes(h, d_seq, (int64_t)B * T, use_mask, d_codes_mine, d_mask);
        h->din.rows = local_rows;
        DMG_TRY(rc);
    }
    DMG_CUDA(h, cudaMemcpyAsync(d_mask_mine, d_mask, (size_t)B * T, cudaMemcpyDeviceToDevice, st));
    if (G > 1) {
        DMG_NCCL(h, g_nccl.AllGather(d_codes_mine, d_codes_all, (size_t)B * T, ncclInt32, s->comm, st));
        DMG_NCCL(h, g_nccl.AllGather(d_mask_mine, d_mask_all, (size_t)B * T, ncclUint8, s->comm, st));
    }
    shard_fill_tiles_kernel<<<h->sm_count * 4, 256, 0, st>>>(d.emb<float>(), geo, d_codes_all, BU * T, E, (uint32_t *)d_tiles);
    h->launches += 1;
    if (G > 1) DMG_NCCL(h, g_nccl.AllReduce(d_tiles, d_tiles, (size_t)BU * T * E, ncclUint32, ncclSum, s->comm, st));

    // ---- level loop -----------------------------------------------------------------------------------------------
    const size_t sel_smem = (size_t)capp * 8 + (size_t)cap * 8;
    const size_t row_smem = (size_t)kRowsRB * ((size_t)4 * E + (size_t)T * E + T + 1) * 4;
    DMG_CUDA(h, cudaFuncSetAttribute(shard_score_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_smem));
    const float scale = (float)(1.0 / std::sqrt((double)E));
    const bool deepfm = d.kind == 1;
    const bool tiled = deepfm || E == 16 || E == 32 || E == 64;   // DIN: embedding sizes with a tile geometry that fits shared memory
    const size_t dfm_smem = DfmGeo::smem(E, T);
    if (deepfm) {
        if (dfm_smem > h->smem_optin) return fail(h, DMG_ERR_UNSUPPORTED, "DeepFM weights (%zu B) do not fit shared memory", dfm_smem);
        DMG_CUDA(h, cudaFuncSetAttribute(shard_score_deepfm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dfm_smem));
    } else if (tiled) {
        switch (E) {
        case 16: DMG_CUDA(h, cudaFuncSetAttribute(shard_score_segments_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shard_score_smem<16>())); break;
        case 32: DMG_CUDA(h, cudaFuncSetAttribute(shard_score_segments_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shard_score_smem<32>())); break;
        case 64: DMG_CUDA(h, cudaFuncSetAttribute(shard_score_segments_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shard_score_smem<64>())); break;
        default: break;
        }
    }
    bool reached = s_level <= L;
    if (reached) {
        shard_select_expand_kernel<<<B, kThreads, sel_smem, st>>>(d_cand, d_score, d_count, cap, capp, beam, s_level, 1, t.d_exists);
        h->launches += 1;
    } else {
        DMG_CUDA(h, cudaMemsetAsync(d_count, 0, (size_t)B * 4, st));
    }
    for (int level = s_level; level < L; level++) {
        shard_select_expand_kernel<<<B, kThreads, sel_smem, st>>>(d_cand, d_score, d_count, cap, capp, beam, level, 0, t.d_exists);
        DMG_CUDA(h, cudaMemsetAsync(d_nreq, 0, (size_t)G * 4, st));
        shard_bucket_kernel<<<B, kThreads, 0, st>>>(d_cand, d_count, cap, geo, d_req, stride, d_nreq, d_seg);
        h->launches += 2;
        if (G > 1) DMG_NCCL(h, g_nccl.AllGather(d_nreq, d_matrix, (size_t)G, ncclInt32, s->comm, st));
        else DMG_CUDA(h, cudaMemcpyAsync(d_matrix, d_nreq, 4, cudaMemcpyDeviceToDevice, st));
        DMG_CUDA(h, cudaMemcpyAsync(h_matrix, d_matrix, (size_t)G * G * 4, cudaMemcpyDeviceToHost, st));
        DMG_CUDA(h, cudaStreamSynchronize(st));
        // h_matrix[src*G + dst] = requests src sends to dst
        if (G > 1) {
            DMG_NCCL(h, g_nccl.GroupStart());
            for (int p = 0; p < G; p++) {
                if (p == s->rank) continue;
                const int ns = h_matrix[s->rank * G + p], nr = h_matrix[p * G + s->rank];
                if (ns) DMG_NCCL(h, g_nccl.Send(d_req + (size_t)p * stride, (size_t)ns * 2, ncclInt32, p, s->comm, st));
                if (nr) DMG_NCCL(h, g_nccl.Recv(d_rreq + (size_t)p * stride, (size_t)nr * 2, ncclInt32, p, s->comm, st));
                if (tiled && ns) DMG_NCCL(h, g_nccl.Send(d_seg + (size_t)p * B, (size_t)B * 2, ncclInt32, p, s->comm, st));
                if (tiled && nr) DMG_NCCL(h, g_nccl.Recv(d_rseg + (size_t)p * B, (size_t)B * 2, ncclInt32, p, s->comm, st));
            }
            DMG_NCCL(h, g_nccl.GroupEnd());
        }
        if (tiled) {
            for (int p = 0; p < G; p++)                          // a requester that sent nothing sent no table either
                if (p != s->rank && h_matrix[p * G + s->rank] == 0) DMG_CUDA(h, cudaMemsetAsync(d_rseg + (size_t)p * B, 0, (size_t)B * 8, st));
            DMG_CUDA(h, cudaMemsetAsync(d_work, 0, 4, st));
            if (deepfm) {
                ShardDfmArgs da;
                da.emb = d.emb<float>(); da.dense = d.tail<float>(); da.tiles = d_tiles;
                da.req_self = d_req; da.req_peer = d_rreq; da.seg_self = d_seg; da.seg_peer = d_rseg;
                da.out_self = d_reply; da.out_peer = d_rsc; da.stride = stride; da.B = B; da.G = G; da.T = T; da.E = E;
                da.work = d_work; da.geo = geo;
                shard_score_deepfm_kernel<<<std::min(G * B, h->sm_count), kThreads, dfm_smem, st>>>(da);
                h->launches += 1;
                for (int p = 0; p < G; p++)
                    if (p != s->rank) s->exchanged_rows += h_matrix[p * G + s->rank];
                goto scored;
            }
            ShardScoreArgs sa;
            sa.emb = d.emb<float>(); sa.wattT = (const float *)d.d_wattT; sa.w1T = (const float *)d.d_w1T;
            sa.b1 = d.b1<float>(); sa.w2 = d.w2<float>(); sa.b2 = d.b2<float>(); sa.tiles = d_tiles; sa.mask = d_mask_all;
            sa.req_self = d_req; sa.req_peer = d_rreq; sa.seg_self = d_seg; sa.seg_peer = d_rseg;
            sa.out_self = d_reply; sa.out_peer = d_rsc; sa.stride = stride; sa.B = B; sa.G = G; sa.T = T; sa.cap = cap;
            sa.scale = scale; sa.work = d_work; sa.geo = geo;
            const int grid = std::min(G * B, h->sm_count);
            switch (E) {
            case 16: shard_score_segments_kernel<16><<<grid, kThreads, shard_score_smem<16>(), st>>>(sa); break;
            case 32: shard_score_segments_kernel<32><<<grid, kThreads, shard_score_smem<32>(), st>>>(sa); break;
            case 64: shard_score_segments_kernel<64><<<grid, kThreads, shard_score_smem<64>(), st>>>(sa); break;
            default: break;
            }
            h->launches += 1;
            for (int p = 0; p < G; p++)
                if (p != s->rank) s->exchanged_rows += h_matrix[p * G + s->rank];
        } else
        for (int p = 0; p < G; p++) {
            const int nr = h_matrix[p * G + s->rank];
            if (!nr) continue;
            const int2 *rq = p == s->rank ? d_req + (size_t)p * stride : d_rreq + (size_t)p * stride;
            float *ro = p == s->rank ? d_reply + (size_t)p * stride : d_rsc + (size_t)p * stride;
            const int grid = std::min((nr + kRowsRB - 1) / kRowsRB, h->sm_count * 8);
            shard_score_rows_kernel<<<grid, kRowsThreads, row_smem, st>>>(d.emb<float>(), geo, (const float *)d.d_wattT, (const float *)d.d_w1T,
                                                                          d.b1<float>(), d.w2<float>(), d.b2<float>(), scale, E, T, nr, rq, cap,
                                                                          p * B, d_tiles, d_mask_all, ro);
            h->launches += 1;
            if (p != s->rank) s->exchanged_rows += nr;
        }
    scored:
        if (G > 1) {
            DMG_NCCL(h, g_nccl.GroupStart());
            for (int p = 0; p < G; p++) {
                if (p == s->rank) continue;
                const int ns = h_matrix[s->rank * G + p], nr = h_matrix[p * G + s->rank];
                if (nr) DMG_NCCL(h, g_nccl.Send(d_rsc + (size_t)p * stride, (size_t)nr, ncclFloat32, p, s->comm, st));
                if (ns) DMG_NCCL(h, g_nccl.Recv(d_reply + (size_t)p * stride, (size_t)ns, ncclFloat32, p, s->comm, st));
            }
            DMG_NCCL(h, g_nccl.GroupEnd());
        }
        for (int p = 0; p < G; p++) {
            const int ns = h_matrix[s->rank * G + p];
            if (!ns) continue;
            shard_scatter_kernel<<<(ns + 255) / 256, 256, 0, st>>>(d_req + (size_t)p * stride, d_reply + (size_t)p * stride, ns, d_score);
            h->launches += 1;
        }
        DMG_CUDA(h, cudaGetLastError());
    }
    shard_final_topk_kernel<<<B, kThreads, (size_t)capp * 8, st>>>(d_cand, d_score, d_count, cap, capp, topk, L, reached ? 1 : 0, t.d_leaf_item,
                                                                   d_cons_off, d_cons, d_items, d_logits, d_cnt_out);
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
    if (cons_off && widen_beam)
        return fail(h, DMG_ERR_UNSUPPORTED, "DeepFM: the eval variant with per-user widened beams runs on the DIN path only");
    if (h->shard && h->shard->world > 1 && cons_off) return fail(h, DMG_ERR_UNSUPPORTED, "consumed items with a sharded table");
    return shard_tdm_retrieve_impl(h, B, item_seq, beam, topk, 0, cons_off, cons, out_items, out_logits, out_counts);
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
