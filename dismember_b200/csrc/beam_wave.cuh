// beam_wave.cuh -- level-synchronous ("wave") tensor-core beam search, E = 64, fp32 model.
//
// Same search, same certified cuts and the same outputs as beam_search_fast_kernel (beam_fast.cuh): ids and logits
// bit-identical to the strict path / CPU oracle.  What changes is the shape of the work.  The persistent kernel walks ONE
// user per CTA through every level -- a serial chain select -> expand -> gather -> MMA -> softmax -> MMA -> epilogue whose
// length, not the machine, bounds it.  Here the batch advances level by level, the way the
// reference's own batched searcher does (otm/.../model/CandidateSearcher.scala:25-55; TDM: Recommender.scala:58-99 per user):
//
//   wave_prologue_kernel   per user, once: history rows -> K / H = M.K tensor-core operands (bf16 hi/lo, UMMA layout),
//                          softmax mask, the user's terms of the error bound                         (K2)
//   per level:
//     wave_select_kernel   one WARP per user: certified cut of the previous level's scores (threshold search on 16 keys
//                          per lane, ballots), band bookkeeping, children 2c+1 / 2c+2 in candidate order (K1 cut + expand)
//     wave_score_kernel    persistent, warp-specialised, 3 CTAs per SM, tiles of (user, 128 candidate rows) from a dynamic
//                          counter.  A control warp gathers the rows with the TMA (cp.async.bulk.tensor ... tile::gather4
//                          from a bf16 hi|lo copy of the node table straight into the 128-byte-swizzled UMMA operand
//                          tile: no register staging, no conversion instructions), issues the tcgen05.mma chains and
//                          refills the tile while four consumer warps run softmax (tcgen05.ld of S) and the
//                          ReLU / W2 epilogue of the previous MMAs                                    (K1 score)
//   wave_final_kernel      per user: ONE strict batch over every band row deferred at a cut and over the topk candidates,
//                          proofs of the deferred cuts, topk by the reference's (score desc, position asc) order  (K3)
//
// Beams, scores, band lists and per-user state live in global memory between the kernels (about 9 KB per user: L2
// resident).  The hi|lo table costs the same 256 bytes per row as the fp32 table it is derived from (x = hi + lo +
// O(2^-18 x)); the strict re-scores read the fp32 table.
#pragma once
#include <cuda.h>

#include "beam_fast.cuh"

namespace dmg {

// Programmatic dependent launch: the kernels of a batch form one chain on the stream; each lets the next one start (launch_dependents)
// as soon as it is resident and waits (grid_dep_wait) for its predecessor to complete before it touches anything the chain produced,
// so that launch latency, barrier set-up, TMEM allocation and the weight image load of kernel n + 1 run under the tail of kernel n.
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

struct WaveUser {                   // per-user search state (64 B)
    float eps;                      // bound on |fast - strict| of the scores in WaveParams::score (level being cut next)
    float kmax, zk, hw;             // user terms of the bound (DESIGN.md "certified cuts")
    uint32_t maskbits;              // Mask input, bit j = history slot j masked
    int32_t vcount, nseg;           // deferred verification list
    int32_t flags;                  // 1 redo, 2 scored, 4 every slot masked
    int32_t redo_why;
    int32_t pad[7];
};
enum { WU_REDO = 1, WU_SCORED = 2, WU_ALLMASK = 4 };

struct WaveGeo {
    static constexpr int THREADS = 256;                      // 2 warpgroups (warp w: TMEM lanes 32 (w & 3) ..), one per pipeline stage
    // shared memory of the scorer, two pipeline stages s = tile & 1 (base 1024-byte aligned)
    static constexpr int XH = 0, XL = 16384, X_STAGE = 32768; // X[s]: [128 rows][128 B] bf16 hi / lo, SWIZZLE_128B (TMA gather4 destination)
    static constexpr int BH = 65536, BL = 77824;             // B operand [96 n][64 k]: n < 64 W1x outputs, 64 + 16 s + j = history slot j of stage s;
    static constexpr int B_LBO = 1536;                       //   K-major, no swizzle, LBO 1536 (96 rows x 16 B), SBO 128
    static constexpr int HH = 90112, HL = 92160, H_STAGE = 4096;  // H[s]: B operand [64 o][16 j]: LBO 1024, SBO 128
    static constexpr int PH = 98304, PL = 102400, P_STAGE = 8192; // P[s]: A operand [128 rows][16 j]: LBO 2048, SBO 128
    static constexpr int ADDV = 114688, ADDV_STAGE = 128;    // [tile & 3]: [16] additive softmax mask + [4] user flags (two slots per stage: a warp may
    static constexpr int BAR = 115200;                       //   still read tile t's while warp 0 of its group writes tile t + 2's); 9 mbarriers
    static constexpr int TMEMP = BAR + 13 * 8;
    static constexpr int SMEM = TMEMP + 24;
    static constexpr int UOP_BYTES = 8320;                   // per-user operand image [KH | KL | HH | HL | addv | flags] (8272, padded)
    static constexpr int VCAP = FastGeo::VCAP, MAX_UNC = FastGeo::MAX_UNC;
    static constexpr int NWIN = 16;                          // narrow park: at most this many band rows are settled strictly in place
};
enum { WB_W1 = 0, WB_HFULL = 1, WB_XFULL = 3, WB_M1 = 5, WB_M2 = 7 };   // + stage

struct WaveParams {
    int B, T, cap, beam;
    const int32_t *beam_user;
    const float *emb;
    const int32_t *hist;
    const uint8_t *hist_mask;
    const uint32_t *exists;
    int leaf_level, sparse_from;
    float scale;
    int32_t *code[2];               // [B][cap] candidate codes, ping-pong by level parity
    float *score;                   // [B][cap] fast scores of the current candidates
    int32_t *count;                 // [B]
    WaveUser *user;
    unsigned char *uop;             // [B][UOP_BYTES]
    int32_t *v_code; float *v_fast; uint32_t *v_meta; float *v_segeps;     // [B][VCAP] x3, [B][32]
    const float *mT, *zvec, *lvl_vx, *lvl_nx, *b1;
    float cA, cZ, cH, cGamma, tau;
    unsigned long long *stats;
    int32_t *redo_list, *redo_count, *host_flags;
    int32_t *tile_list;             // [B * ceil(cap / 128)] tiles of the level being scored: user << 12 | first row >> 7 << 8 | rows - 1 ... see wave_tile_pack
    int32_t *tile_count;            // per level: tiles appended by the select kernel
    const unsigned char *w1img;     // [W1x hi | W1x lo] as the n < 64 rows of the B operand, 2 x 12288 B
    const unsigned char *split;     // bf16 hi|lo table, 256 B per code
    int kind;                       // 0 DIN (this file), 1 DeepFM (beam_wave_dfm.cuh): the bound and the strict re-scores differ
    const float *dfm_dense;         // DeepFM: [W1 | b1 | W2 | b2]
};
struct WaveW2 { float w2[64]; float b2; };
struct DfmConsts {                      // DeepFM scorer constants (beam_wave_dfm.cuh): host-computed, passed by value
    float w2[16], b1[16];               // T + 1 <= 16 hidden units
    float b2;
    float aw2[16];                      // |w2_o|
    int F;
};

// ---- bf16 hi|lo copy of the node table: row c = [64 hi | 64 lo] bf16 (256 B), i.e. a [2 rows][64] bf16 matrix ------
static __global__ void wave_split_table_kernel(const float *__restrict__ emb, int64_t rows, unsigned char *__restrict__ out)
{
    const int64_t n = rows * 8;                               // 8 chunks of 8 floats per row
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i >> 3;
        const int c = (int)(i & 7);
        const float4 a = ldg_row16(emb + r * 64 + c * 8), b = ldg_row16(emb + r * 64 + c * 8 + 4);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint4 hi, lo;
        split8(v, hi, lo);
        *reinterpret_cast<uint4 *>(out + r * 256 + c * 16) = hi;
        *reinterpret_cast<uint4 *>(out + r * 256 + 128 + c * 16) = lo;
    }
}

// W1 item half [o][k] -> rows n < 64 of the B operand image [hi | lo], chunk kc at kc * 1536, row o at o * 16
static __global__ void wave_w1_image_kernel(const float *__restrict__ w1, unsigned char *__restrict__ img)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 64 * 8) return;
    const int o = i & 63, kc = i >> 6;
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = __ldg(w1 + o * 128 + kc * 8 + q);
    uint4 hi, lo;
    split8(v, hi, lo);
    *reinterpret_cast<uint4 *>(img + kc * 1536 + o * 16) = hi;
    *reinterpret_cast<uint4 *>(img + 12288 + kc * 1536 + o * 16) = lo;
}
// a tile of the level being scored: user, first candidate row (multiple of 128), rows (1..128)
__device__ __forceinline__ int32_t wave_tile_pack(int user, int row0, int nr) { return (user << 10) | ((row0 >> 7) << 8) | (nr - 1); }

// ---- K2: per-user prologue ----------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) wave_prologue_kernel(const WaveParams p)
{
    constexpr int E = 64, LD = 68;
    __shared__ __align__(16) float sKf[kMaxT * LD];
    __shared__ float sHmax[4 * E];
    __shared__ int sCode[kMaxT];
    __shared__ int sRed[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, user = blockIdx.x, T = p.T;
    grid_dep_launch();
    grid_dep_wait();
    unsigned char *uop = p.uop + (size_t)user * WaveGeo::UOP_BYTES;
    if (tid < kMaxT) {
        int c = -1, m = 0;
        if (tid < T) { c = p.hist[(size_t)user * T + tid]; m = p.hist_mask[(size_t)user * T + tid]; }
        sCode[tid] = c;
        const uint32_t mb = __ballot_sync(0x0000ffffu, m != 0);
        reinterpret_cast<float *>(uop + 8192)[tid] = (tid < T && !m) ? 0.0f : -3.4028234663852886e+38f;
        if (tid == 0) { sRed[0] = (int)mb; sRed[1] = 0; }
    }
    __syncthreads();
    {                                                         // history rows, 16 lanes x 16 B per row
        const int j = tid >> 4, c16 = tid & 15;
        const int c = sCode[j];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c >= 0) v = ldg_row16(p.emb + (size_t)c * E + c16 * 4);
        *reinterpret_cast<float4 *>(sKf + j * LD + c16 * 4) = v;
    }
    __syncthreads();
    if (tid < 128) {                                          // K as B operand of S = X . K^T: [16 j][64 k], LBO 256
        const int j = tid & 15, kc = tid >> 4;
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; q++) v[q] = sKf[j * LD + kc * 8 + q];
        uint4 hi, lo;
        split8(v, hi, lo);
        *reinterpret_cast<uint4 *>(uop + kc * 256 + j * 16) = hi;
        *reinterpret_cast<uint4 *>(uop + 2048 + kc * 256 + j * 16) = lo;
    } else if (tid < 192) {                                   // ZK = sum_k z_k max_j |K_jk|
        const int k = tid - 128;
        float kab = 0.0f;
#pragma unroll
        for (int j = 0; j < kMaxT; j++) { const float v = fabsf(sKf[j * LD + k]); kab = (v > kab || v != v) ? v : kab; }
        float zk = __ldg(p.zvec + k) * kab;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) zk += __shfl_xor_sync(0xffffffffu, zk, o);
        if (lane == 0) sRed[2 + (warp & 1)] = __float_as_int(zk);
    } else if (tid < 192 + kMaxT) {                           // Kmax = max_j |K_j|_2
        const int j = tid - 192;
        float n2 = 0.0f;
        for (int k = 0; k < E; k++) n2 = fmaf(sKf[j * LD + k], sKf[j * LD + k], n2);
        float nj = sqrtf(n2) * 1.0001f;
        if (!(nj == nj)) nj = __int_as_float(0x7f800000);
        atomicMax(&sRed[1], __float_as_int(nj));
    }
    {                                                         // H[j][o] = sum_k M[o][k] K[j][k]; row 15 carries b1
        const int o = tid & 63, jg = tid >> 6;
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        const float *kr = sKf + (jg * 4) * LD;
#pragma unroll 8
        for (int k = 0; k < E; k++) {
            const float m = __ldg(p.mT + k * E + o);
#pragma unroll
            for (int q = 0; q < 4; q++) acc[q] = fmaf(m, kr[q * LD + k], acc[q]);
        }
        float hm = fmaxf(fmaxf(fabsf(acc[0]), fabsf(acc[1])), fmaxf(fabsf(acc[2]), fabsf(acc[3])));
        if (acc[0] != acc[0] || acc[1] != acc[1] || acc[2] != acc[2] || acc[3] != acc[3]) hm = __int_as_float(0x7f800000);
        sHmax[jg * E + o] = hm;
        if (jg == 3) acc[3] = __ldg(p.b1 + o);
        uint2 hi, lo;
        split_pair(acc[0], acc[1], hi.x, lo.x);
        split_pair(acc[2], acc[3], hi.y, lo.y);
        const int off = (jg >> 1) * 1024 + o * 16 + (jg & 1) * 8;
        *reinterpret_cast<uint2 *>(uop + 4096 + off) = hi;
        *reinterpret_cast<uint2 *>(uop + 6144 + off) = lo;
    }
    __syncthreads();
    if (tid < E) {
        float hw = fmaxf(fmaxf(sHmax[tid], sHmax[E + tid]), fmaxf(sHmax[2 * E + tid], sHmax[3 * E + tid])) * fabsf(__ldg(p.b1 + E + tid));   // w2 follows b1
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) hw += __shfl_xor_sync(0xffffffffu, hw, o);
        if (lane == 0) sRed[4 + warp] = __float_as_int(hw);
    }
    __syncthreads();
    if (tid == 0) {
        WaveUser st;
        memset(&st, 0, sizeof(st));
        st.kmax = __int_as_float(sRed[1]);
        st.zk = (__int_as_float(sRed[2]) + __int_as_float(sRed[3])) * 1.0001f;
        st.hw = (__int_as_float(sRed[4]) + __int_as_float(sRed[5])) * 1.0001f;
        st.maskbits = (uint32_t)sRed[0];
        const uint32_t full = (1u << T) - 1u;
        st.flags = (((uint32_t)sRed[0] & full) == full) ? WU_ALLMASK : 0;
        reinterpret_cast<int32_t *>(uop + 8192 + 64)[0] = st.flags;
        p.user[user] = st;
        p.count[user] = 0;
    }
}

// ---- K1 (cut + expand): one warp per user ---------------------------------------------------------------------------
// Lane l owns the candidates i = 32 j + l, j < 16 (cap <= 512).  `level` is the tree level of the current candidates;
// the children written to code[(slot ^ 1)] sit on level + 1.
template <int NJ>
__device__ __forceinline__ void warp_select(const uint32_t (&k)[NJ], int kk, uint32_t lo, uint32_t hi, int clo,
                                            uint32_t &kdn, uint32_t &kup, int &iters)
{
    int chi = 0, it = 0;
#pragma unroll 1
    while (clo != kk && hi - lo > 255u && hi > lo) {
        uint32_t mid = lo + ((hi - lo) >> 1) + 1u;
        if (it < 12) {
            const float flo = key_to_float(lo), fhi = key_to_float(hi);
            const float fr = (it & 1) ? 0.5f : ((float)(kk - chi) - 0.5f) / (float)(clo - chi);
            const float fm = fhi - (fhi - flo) * fr;
            if (fm == fm) mid = order_key(fm);
        }
        mid = min(max(mid, lo + 1u), hi);
        int c = 0;
#pragma unroll
        for (int j = 0; j < NJ; j++) c += k[j] >= mid ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        it++;
        if (c >= kk) { lo = mid; clo = c; } else { hi = mid - 1u; chi = c; }
    }
    iters = it;
    if (clo == kk) {
        uint32_t a = 0xffffffffu, b = 0u;
#pragma unroll
        for (int j = 0; j < NJ; j++) { a = min(a, k[j] >= lo ? k[j] : 0xffffffffu); b = max(b, k[j] < lo ? k[j] : 0u); }
        kdn = __reduce_min_sync(0xffffffffu, a);
        kup = __reduce_max_sync(0xffffffffu, b);
        if (kup == 0u) kup = kdn;
    } else {
        kdn = lo; kup = hi;
    }
}

// children of the surviving candidates in candidate order, the tiles and the bound of the scores about to be computed
template <int NJ, int KIND>
__device__ __forceinline__ void wave_expand_finish(const WaveParams &p, WaveUser *st, int user, int level, int lane, const int32_t (&code)[NJ],
                                                   const uint32_t *sKeepW, int count, int32_t *__restrict__ nxt, int flags,
                                                   unsigned long long st_cut, unsigned long long st_recut, unsigned long long st_iters)
{
    const uint32_t lt = (1u << lane) - 1u;
    int out = 0;
    const uint32_t *bm = (level + 1 >= p.sparse_from) ? p.exists : nullptr;
    const int nj = (count + 31) >> 5;
#pragma unroll
    for (int j = 0; j < NJ; j++) {
        if (j < nj) {
            const bool kp = sKeepW[j] >> lane & 1u;
            const int64_t c = code[j];
            const bool a1 = kp && code_exists(bm, 2 * c + 1), a2 = kp && code_exists(bm, 2 * c + 2);
            const uint32_t m1 = __ballot_sync(0xffffffffu, a1), m2 = __ballot_sync(0xffffffffu, a2);
            int o = out + __popc(m1 & lt) + __popc(m2 & lt);
            if (a1) nxt[o++] = (int32_t)(2 * c + 1);
            if (a2) nxt[o] = (int32_t)(2 * c + 2);
            out += __popc(m1) + __popc(m2);
        }
    }
    const int nt = (out + 127) >> 7;
    if (lane < nt) {                                              // this user's tiles of the next level
        int base = 0;
        if (lane == 0) base = atomicAdd(p.tile_count + level + 1, nt);
        base = __shfl_sync((1u << nt) - 1u, base, 0);
        const int row0 = lane * 128;
        p.tile_list[base + lane] = wave_tile_pack(user, row0, out - row0 < 128 ? out - row0 : 128);
    }
    if (lane == 0) {
        p.count[user] = out;
        // eps of the scores about to be computed (children sit on tree level `level + 1`)
        const float u = 5.9604645e-8f;
        const float vx = __ldg(p.lvl_vx + level + 1), nx = __ldg(p.lvl_nx + level + 1);
        const float smax = p.scale * nx * st->kmax;
        const float ds = 6.1035156e-5f * smax;
        const float dp1 = 2.1f * ds + 2.0f * (float)(p.T + 8) * u + 2.0f * 9.5367432e-7f * (1.0f + 2.0f * smax);
        float eps = p.cA * vx + p.cZ * st->zk + (p.cH + dp1) * st->hw * 1.05f + p.cGamma;
        if (!(ds < 0.04f) || !(eps < 1e30f)) eps = __int_as_float(0x7f800000);
        if (KIND == 1) {                                          // DeepFM (beam_wave_dfm.cuh): vt . |x| + a0 + a1 nx + a2 nx^2, 25 % on top
            eps = 1.25f * (vx + st->zk + st->kmax * nx + st->hw * nx * nx);
            if (!(eps < 1e30f)) eps = __int_as_float(0x7f800000);
        }
        st->eps = eps * p.tau;
        if (out > 0) st->flags = flags | WU_SCORED;
        if (p.stats) {
            if (st_cut) { atomicAdd(&p.stats[0], st_cut); atomicAdd(&p.stats[7], st_iters); }
            if (st_recut) atomicAdd(&p.stats[1], st_recut);
            atomicAdd(&p.stats[3], (unsigned long long)out);
        }
    }
}

#ifndef DMG_WAVE_PARK_FRAC
#define DMG_WAVE_PARK_FRAC 0.0078125f                            // gap at the cut below eps / 128: settle the band strictly in place
#endif
struct WaveStrictW { const float *wattT, *w1T, *b1, *w2; float b2; };
__device__ void dfm_strict_batch128(const WaveParams &p, const DfmConsts &dc, const float *__restrict__ dense, int user,
                                    const int32_t *__restrict__ sRow, int n, float *__restrict__ sOut, float *__restrict__ scr,
                                    float *__restrict__ sDense, bool dense_loaded);   // beam_wave_dfm.cuh

// 4 users per CTA.  Phase 1: every warp cuts its user; a cut whose band cannot be deferred (fast-score gap at the cut below
// eps / 128) is parked.  Phase 2: the whole CTA scores the band rows of its parked users strictly (sequential-k fma chains,
// the oracle's bits).  Phase 3: the parked warps finish their cuts with the strict order.
// Per-lane state is the 16 candidate codes / scores; the surviving set is a 16-word bitmap in shared memory (word j, bit
// lane = candidate 32 j + lane) and the band rows are compacted into shared-memory lists that the lanes then walk in
// parallel -- rolled loops, so that the kernel stays a few thousand instructions (it runs at low occupancy: instruction
// fetch is what bounds a long unrolled body).
template <int NJ, int KIND = 0>
static __global__ void __launch_bounds__(128) wave_select_kernel(const WaveParams p, const WaveStrictW sw, int level, int slot, const DfmConsts dc)
{
    constexpr int MU = WaveGeo::MAX_UNC;
    __shared__ __align__(16) float sScr[FastGeo::STRICT_SCR / 4];
    __shared__ uint32_t sKeyU[4][MU];                             // band rows: order key of the fast score,
    __shared__ int sUPos[4][MU];                                  //   candidate position,
    __shared__ int32_t sLCode[4][MU];                             //   code,
    __shared__ float sLStr[4][MU];                                //   fast score, then (parked cuts) the strict score
    __shared__ int sFr[4][MU];                                    //   rank by fast score
    __shared__ uint32_t sKeepW[4][NJ];
    __shared__ int sGap[4][2];
    __shared__ int sPark[4][4];                                   // rows to score strictly (0 = not parked), need, narrow?, n_unc
    __shared__ int32_t sWCode[4][WaveGeo::NWIN];                  // narrow park: the band rows within eps / 32 of the cut,
    __shared__ float sWStr[4][WaveGeo::NWIN];                     //   their strict scores,
    __shared__ int sWIdx[4][WaveGeo::NWIN];                       //   their index in the band lists (then: chosen?)
    extern __shared__ __align__(16) unsigned char sel_dyn[];      // DeepFM only: the dense weights of a parked cut's strict chains
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int user = blockIdx.x * 4 + warp;
    const uint32_t lt = (1u << lane) - 1u;
    if (lane == 0) sPark[warp][0] = 0;
    grid_dep_launch();
    grid_dep_wait();
    WaveUser *st = p.user + (user < p.B ? user : 0);
    int flags = 0, beam = 1, s_level = 0;
    bool live = user < p.B;
    if (live) {
        flags = st->flags;
        beam = p.beam_user ? __ldg(p.beam_user + user) : p.beam;
        s_level = 31 - __clz(beam);
        if ((flags & WU_REDO) || level < s_level || s_level >= p.leaf_level) {
            if (lane == 0) p.count[user] = 0;
            live = false;
        }
    }
    const int32_t *cur = p.code[slot] + (size_t)(live ? user : 0) * p.cap;
    int32_t *nxt = p.code[slot ^ 1] + (size_t)(live ? user : 0) * p.cap;
    int32_t code[NJ];
    int count = 0;
    unsigned long long st_cut = 0, st_recut = 0, st_iters = 0;
    bool parked = false;
    if (live) {
        if (level == s_level) {                                   // initial beam: every existing code of the level, no scores yet
            const int64_t start = ((int64_t)1 << s_level) - 1;
            count = 1 << s_level;
#pragma unroll
            for (int j = 0; j < NJ; j++) {
                const int i = 32 * j + lane;
                code[j] = (int32_t)(start + i);
                const uint32_t m = __ballot_sync(0xffffffffu, i < count && code_exists(p.exists, start + i));
                if (lane == 0) sKeepW[warp][j] = m;
            }
        } else {
            count = p.count[user];
            float f[NJ];
            const float *sc = p.score + (size_t)user * p.cap;
#pragma unroll
            for (int j = 0; j < NJ; j++) {
                const int i = 32 * j + lane;
                code[j] = i < count ? cur[i] : 0;
                f[j] = i < count ? sc[i] : 0.0f;
            }
            if (count <= beam) {
#pragma unroll
                for (int j = 0; j < NJ; j++) {
                    const uint32_t m = __ballot_sync(0xffffffffu, 32 * j + lane < count);
                    if (lane == 0) sKeepW[warp][j] = m;
                }
            } else {
                st_cut = 1;
                uint32_t key[NJ];
                uint32_t mn = 0xffffffffu, mx = 0u;
#pragma unroll
                for (int j = 0; j < NJ; j++) {
                    key[j] = 32 * j + lane < count ? order_key(f[j]) : 0u;
                    mn = min(mn, key[j] ? key[j] : 0xffffffffu); mx = max(mx, key[j]);
                }
                mn = __reduce_min_sync(0xffffffffu, mn);
                mx = __reduce_max_sync(0xffffffffu, mx);
                uint32_t kdn, kup;
                int iters;
                warp_select(key, beam, mn, mx, count, kdn, kup, iters);
                st_iters = iters;
                const float eps_level = st->eps;
                const float band = 2.0f * eps_level * 1.0001f + 1e-30f;
                const float up = key_to_float(kup) + band, dn = key_to_float(kdn) - band;
                int n_keep = 0, n_unc = 0;
#pragma unroll
                for (int j = 0; j < NJ; j++) {
                    const bool valid = 32 * j + lane < count;
                    const bool isunc = valid && !(f[j] > up) && !(f[j] < dn);
                    const uint32_t mk = __ballot_sync(0xffffffffu, valid && !(f[j] < dn)), mu = __ballot_sync(0xffffffffu, isunc);
                    if (lane == 0) sKeepW[warp][j] = mk;
                    if (isunc) {
                        const int e = n_unc + __popc(mu & lt);
                        if (e < MU) { sKeyU[warp][e] = key[j]; sUPos[warp][e] = 32 * j + lane; sLCode[warp][e] = code[j]; sLStr[warp][e] = f[j]; }
                    }
                    n_keep += __popc(mk); n_unc += __popc(mu);
                }
                __syncwarp();
                if (n_keep != beam) {                             // some uncertain row must go
                    st_recut = 1;
                    if (n_unc > MU) {
                        if (lane == 0) { st->flags = flags | WU_REDO; st->redo_why = 0; p.count[user] = 0; }
                        live = false;
                    } else {
                        const int need = beam - (n_keep - n_unc);
#pragma unroll 1
                        for (int e = lane; e < n_unc; e += 32) {  // rank of every band row by its FAST score, the gap at the cut
                            const uint32_t ke = sKeyU[warp][e];
                            const int pe = sUPos[warp][e];
                            int fr = 0;
#pragma unroll 4
                            for (int q = 0; q < n_unc; q++) {
                                const uint32_t kq = sKeyU[warp][q];
                                fr += (kq > ke || (kq == ke && sUPos[warp][q] < pe)) ? 1 : 0;
                            }
                            sFr[warp][e] = fr;
                            if (fr == need - 1) sGap[warp][0] = __float_as_int(sLStr[warp][e]);
                            if (fr == need) sGap[warp][1] = __float_as_int(sLStr[warp][e]);
                        }
                        __syncwarp();
                        const float gap = __int_as_float(sGap[warp][0]) - __int_as_float(sGap[warp][1]);
                        const int vcount = st->vcount, nseg = st->nseg;
                        // Observed |fast - strict| stays below 1 % of eps: with a gap of eps/128 or more the fast order is taken
                        // now and PROVEN by wave_final_kernel; a narrower gap is settled strictly right here.
                        const bool defer = gap >= DMG_WAVE_PARK_FRAC * eps_level && vcount + n_unc <= WaveGeo::VCAP && nseg < 32 && eps_level < 1e30f;
                        if (defer) {
#pragma unroll 1
                            for (int e = lane; e < n_unc; e += 32) {
                                const uint32_t ch = sFr[warp][e] < need ? 1u : 0u;
                                const size_t g = (size_t)user * WaveGeo::VCAP + vcount + e;
                                p.v_code[g] = sLCode[warp][e]; p.v_fast[g] = sLStr[warp][e];
                                p.v_meta[g] = (uint32_t)vcount | ((uint32_t)n_unc << 8) | ((uint32_t)need << 16) | (ch << 24) | ((uint32_t)nseg << 25);
                                const int pe = sUPos[warp][e];
                                if (!ch) atomicAnd(&sKeepW[warp][pe >> 5], ~(1u << (pe & 31)));
                            }
                            __syncwarp();
                            if (lane == 0) { p.v_segeps[(size_t)user * 32 + nseg] = eps_level; st->vcount = vcount + n_unc; st->nseg = nseg + 1; }
                        } else {
                            // park.  Narrow form (the usual one): only the band rows within eps / 32 of the cut -- those whose fast
                            // order cannot be trusted -- are scored strictly in place (phase 2); the rest of the band keeps its fast
                            // order and the WHOLE band is deferred to the end-of-search proof like any other cut (phase 3), which
                            // checks the chosen set against the strict ranks whatever way it was chosen.  Full form (no room on the
                            // deferred list, or a crowded window): strict scores of the whole band decide, nothing to prove.
                            const bool room = vcount + n_unc <= WaveGeo::VCAP && nseg < 32 && eps_level < 1e30f;
                            const float mid = 0.5f * (__int_as_float(sGap[warp][0]) + __int_as_float(sGap[warp][1])), win = 0.03125f * eps_level;
                            int n_w = 0;
                            if (room) {
#pragma unroll 1
                                for (int e0 = 0; e0 < n_unc; e0 += 32) {
                                    const int e = e0 + lane;
                                    const bool in = e < n_unc && fabsf(sLStr[warp][e] - mid) <= win;
                                    const uint32_t mw = __ballot_sync(0xffffffffu, in);
                                    if (in) {
                                        const int i = n_w + __popc(mw & lt);
                                        if (i < WaveGeo::NWIN) { sWCode[warp][i] = sLCode[warp][e]; sWIdx[warp][i] = e; }
                                    }
                                    n_w += __popc(mw);
                                }
                            }
                            const bool narrow = room && n_w >= 2 && n_w <= WaveGeo::NWIN;
                            if (lane == 0) { sPark[warp][0] = narrow ? n_w : n_unc; sPark[warp][1] = need; sPark[warp][2] = narrow ? 1 : 0; sPark[warp][3] = n_unc; }
                            parked = true;
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (live && !parked) wave_expand_finish<NJ, KIND>(p, st, user, level, lane, code, sKeepW[warp], count, nxt, flags, st_cut, st_recut, st_iters);
    }
    __syncthreads();
    // ---- phase 2: strict scores of the parked bands, whole CTA ----
    bool any = false, dense_loaded = false;
#pragma unroll 1
    for (int w = 0; w < 4; w++) {
        const int n = sPark[w][0];
        if (n == 0) continue;
        any = true;
        const int pu = blockIdx.x * 4 + w;
        const bool narrow = sPark[w][2] != 0;
        const int32_t *rows = narrow ? sWCode[w] : sLCode[w];
        float *strict = narrow ? sWStr[w] : sLStr[w];
        if (KIND == 1) {                                           // DeepFM: the oracle-order chains of deepfm_common.cuh
            if (!narrow && tid < n) sKeyU[w][tid] = __float_as_uint(sLStr[w][tid]);
            __syncthreads();
            dfm_strict_batch128(p, dc, p.dfm_dense, pu, rows, n, strict, sScr, reinterpret_cast<float *>(sel_dyn), dense_loaded);
            dense_loaded = true;
            continue;
        }
        for (int i = tid; i < kMaxT * 64; i += 128) {              // history rows (fp32), zero rows for padding
            const int j = i >> 6, k = i & 63;
            const int c = j < p.T ? p.hist[(size_t)pu * p.T + j] : -1;
            sScr[j * FastGeo::KLD + k] = c >= 0 ? __ldg(p.emb + (size_t)c * 64 + k) : 0.0f;
        }
        if (!narrow && tid < n) sKeyU[w][tid] = __float_as_uint(sLStr[w][tid]);   // keep the fast scores (error statistics)
        __syncthreads();
        strict_score_batch128(p.emb, rows, n, strict, sScr, p.user[pu].maskbits, p.T, p.scale, sw.wattT, sw.w1T, sw.b1, sw.w2, sw.b2);
    }
    if (!any) return;
    __syncthreads();
    // ---- phase 3: the parked warps finish their cuts with the strict order ----
    if (parked && sPark[warp][2]) {                               // narrow park: strict order inside the window, fast order outside, proof later
        const int n_w = sPark[warp][0], need = sPark[warp][1], n_unc = sPark[warp][3];
        const float eps_level = st->eps;
        int above = 0;                                            // band rows with a fast rank above the cut that are NOT in the window: chosen
#pragma unroll 1
        for (int e0 = 0; e0 < n_unc; e0 += 32) {
            const int e = e0 + lane;
            above += __popc(__ballot_sync(0xffffffffu, e < n_unc && sFr[warp][e] < need));
        }
        int in_above = 0, tie = 0, mych = 0;
        float ratio = 0.0f;
        if (lane < n_w) {
            const int e = sWIdx[warp][lane];
            in_above = sFr[warp][e] < need ? 1 : 0;
        }
        above -= __popc(__ballot_sync(0xffffffffu, in_above != 0));
        const int need_w = need - above;                          // rows the window contributes to the kept set
        if (lane < n_w) {
            const float mine = sWStr[warp][lane];
            const uint32_t ks = order_key(mine);
            int g = 0, t = 0;
            for (int q = 0; q < n_w; q++) {
                const uint32_t kq = order_key(sWStr[warp][q]);
                g += kq > ks ? 1 : 0;
                t += (kq == ks && q != lane) ? 1 : 0;
            }
            tie = (g < need_w && g + t >= need_w) ? 1 : 0;
            mych = g + t < need_w ? 1 : 0;
            if (eps_level > 0.0f && eps_level < 1e30f) ratio = fabsf(mine - sLStr[warp][sWIdx[warp][lane]]) / eps_level;
        }
        tie = __any_sync(0xffffffffu, tie);
        __syncwarp();
        if (lane < n_w) sFr[warp][sWIdx[warp][lane]] = mych ? -1 : 0x7fffffff;     // the window rows' fast ranks give way to the strict decision
        __syncwarp();
        if (p.stats) {
            for (int o = 16; o > 0; o >>= 1) ratio = fmaxf(ratio, __shfl_xor_sync(0xffffffffu, ratio, o));
            if (lane == 0) {
                if (ratio > 0.0f) atomicMax(reinterpret_cast<unsigned int *>(&p.stats[4]), __float_as_uint(ratio));
                atomicAdd(&p.stats[2], (unsigned long long)n_w); atomicAdd(&p.stats[6], 1ull);
            }
        }
        if (tie) {
            if (lane == 0) {
                st->flags = flags | WU_REDO; st->redo_why = 1; p.count[user] = 0;
                if (p.stats) { atomicAdd(&p.stats[0], st_cut); atomicAdd(&p.stats[1], st_recut); }
            }
        } else {
            const int vcount = st->vcount, nseg = st->nseg;
#pragma unroll 1
            for (int e = lane; e < n_unc; e += 32) {
                const uint32_t ch = sFr[warp][e] < need ? 1u : 0u;
                const size_t g = (size_t)user * WaveGeo::VCAP + vcount + e;
                p.v_code[g] = sLCode[warp][e]; p.v_fast[g] = sLStr[warp][e];
                p.v_meta[g] = (uint32_t)vcount | ((uint32_t)n_unc << 8) | ((uint32_t)need << 16) | (ch << 24) | ((uint32_t)nseg << 25);
                const int pe = sUPos[warp][e];
                if (!ch) atomicAnd(&sKeepW[warp][pe >> 5], ~(1u << (pe & 31)));
            }
            __syncwarp();
            if (lane == 0) { p.v_segeps[(size_t)user * 32 + nseg] = eps_level; st->vcount = vcount + n_unc; st->nseg = nseg + 1; }
            __syncwarp();
            wave_expand_finish<NJ, KIND>(p, st, user, level, lane, code, sKeepW[warp], count, nxt, flags, st_cut, st_recut, st_iters);
        }
    } else if (parked) {
        const int n_unc = sPark[warp][0], need = sPark[warp][1];
        const float eps_level = st->eps;
        int tie = 0;
        float ratio = 0.0f;
#pragma unroll 1
        for (int e = lane; e < n_unc; e += 32) {
            const float mine = sLStr[warp][e];
            const uint32_t ks = order_key(mine);
            // strict rank = rows strictly above (+ an undetermined share of exact ties: the reference orders ties by candidate
            // position, which this path does not track).  Only a tie that STRADDLES the cut is undecidable here.
            int g = 0, t = 0;
#pragma unroll 4
            for (int q = 0; q < n_unc; q++) {
                const uint32_t kq = order_key(sLStr[warp][q]);
                g += kq > ks ? 1 : 0;
                t += (kq == ks && q != e) ? 1 : 0;
            }
            tie |= (g < need && g + t >= need) ? 1 : 0;
            const int pe = sUPos[warp][e];
            if (!(g + t < need)) atomicAnd(&sKeepW[warp][pe >> 5], ~(1u << (pe & 31)));
            if (eps_level > 0.0f && eps_level < 1e30f) ratio = fmaxf(ratio, fabsf(mine - __uint_as_float(sKeyU[warp][e])) / eps_level);
        }
        tie = __any_sync(0xffffffffu, tie);
        if (p.stats) {
            for (int o = 16; o > 0; o >>= 1) ratio = fmaxf(ratio, __shfl_xor_sync(0xffffffffu, ratio, o));
            if (lane == 0) {
                if (ratio > 0.0f) atomicMax(reinterpret_cast<unsigned int *>(&p.stats[4]), __float_as_uint(ratio));
                atomicAdd(&p.stats[2], (unsigned long long)n_unc); atomicAdd(&p.stats[6], 1ull);
            }
        }
        __syncwarp();
        if (tie) {
            if (lane == 0) {
                st->flags = flags | WU_REDO; st->redo_why = 1; p.count[user] = 0;
                if (p.stats) { atomicAdd(&p.stats[0], st_cut); atomicAdd(&p.stats[1], st_recut); }
            }
        } else {
            wave_expand_finish<NJ, KIND>(p, st, user, level, lane, code, sKeepW[warp], count, nxt, flags, st_cut, st_recut, st_iters);
        }
    }
}

// ---- K1 (score): persistent warp-specialised tile scorer ------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// two 32-column accumulator loads, one wait
__device__ __forceinline__ void tmem_ld32x2(uint32_t taddr, float (&a)[32], float (&b)[32])
{
    uint32_t r[32], q[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
          "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]), "=r"(q[16]),
          "=r"(q[17]), "=r"(q[18]), "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]),
          "=r"(q[25]), "=r"(q[26]), "=r"(q[27]), "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
        : "r"(taddr + 32) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) { a[i] = __uint_as_float(r[i]); b[i] = __uint_as_float(q[i]); }
}
// rows c0..c3 of the bf16 hi|lo table (tensor map: [rows][128] bf16, row c = [64 hi | 64 lo], box 64 x 1) -> 4 rows of the hi tile
// (columns 0..63) and 4 rows of the lo tile (columns 64..127): the same four row coordinates for both instructions
__device__ __forceinline__ void tma_gather4_hilo(uint32_t dst_hi, uint32_t dst_lo, const void *tmap, uint32_t bar, int c0, int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%2, {%4, %6, %7, %8, %9}], [%3];\n\t"
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%1], [%2, {%5, %6, %7, %8, %9}], [%3];"
        ::"r"(dst_hi), "r"(dst_lo), "l"(tmap), "r"(bar), "r"(0), "r"(64), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

#ifdef DMG_WAVE_TIMING
#define WTICK(i) do { if (gtid == 0) { const long long t_ = clock64(); wacc[i] += t_ - wlast; wlast = t_; } } while (0)
#else
#define WTICK(i) do { } while (0)
#endif
constexpr uint32_t kIdescBf16M128N96 = (1u << 4) | (1u << 7) | (1u << 10) | ((96u >> 3) << 17) | ((128u >> 4) << 24);

// Tiles of (user, <= 128 candidate rows) come from the list the select kernel wrote; CTA b takes tiles b, b + grid, ...
// Two pipeline stages per CTA, each a warpgroup g (warps 4 g .. 4 g + 3, TMEM lanes 32 (w & 3) ..) with its own X, K rows, H, P
// and TMEM accumulators, working through the tiles t = g, g + 2, ..; 2 CTAs per SM (113 KB of shared memory and 256 TMEM
// columns each), i.e. four independent tile chains per SM -- while one waits for the tensor pipe or the TMA the others compute.
// Per tile:
//   [Hacc | S] = X . [W1x | K0 | K1]^T   128 x 96 x 64 tcgen05.mma, bf16 hi*hi + hi*lo + lo*hi, fp32 in TMEM (the history rows of stage
//                                         g sit in rows 64 + 16 g .. of the B operand), issued by one elected lane of the group's
//                                         second warp as soon as X[g] has landed and the group has drained its accumulators;
//   refill X[g] with the group's next tile: every warp gathers its own 32 rows with the TMA (tile::gather4 from the bf16 hi|lo
//                                         table; one elected lane issues the 16 instructions from warp-uniform coordinates);
//   Mask + SoftMax on S, one row per thread, log2 domain -> P[g] (A operand); group barrier;
//   Hacc += P . H                         128 x 64 x 16 (row 15 of H carries b1), one elected lane of the third warp; under it every
//                                         warp copies its quarter of the next user's K rows (the first warp also mask + flags);
//   read Hacc; group barrier; [second warp: the next tile's first chain]; logit = relu(Hacc) . W2 + b2 from registers.
// DBG: ablation switches for profiling (results are wrong when != 0): 1 no row gather, 2 no softmax math, 4 no epilogue math, 8 no K copy, 16 no MMA
template <int DBG = 0>
static __global__ void __launch_bounds__(WaveGeo::THREADS, 2)
wave_score_kernel(const __grid_constant__ CUtensorMap tmap, const WaveParams p, const WaveW2 w, int slot, int level)
{
    using G = WaveGeo;
    constexpr int dbg = DBG;
    extern __shared__ __align__(1024) unsigned char sm[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm + G::BAR);
    const int tid = threadIdx.x, lane = tid & 31;
    // warp-uniform for the compiler as well: every TMA / MMA / mbarrier operand derived from it goes to uniform registers, and the
    // single-lane issue below (elect.sync) compiles to back-to-back UTMALDG / UTCHMMA instead of one ELECT .. R2UR.BROADCAST ..
    // BRA.U.ANY loop per instruction (~100 cycles each: 1.7 K cycles per refill_x and 1.3 K per 12-MMA chain before this)
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t sbase = smem_u32(sm);
    if (sbase & 1023u) __trap();                                 // the swizzled X tiles need 1024-byte alignment
    grid_dep_launch();

    if (tid == 0) {
        mbar_init(&bar[WB_W1], 1);
        for (int s = 0; s < 2; s++) {
            mbar_init(&bar[WB_HFULL + s], 1);
            mbar_init(&bar[WB_XFULL + s], 4);
            mbar_init(&bar[WB_M1 + s], 1);
            mbar_init(&bar[WB_M2 + s], 1);
        }
        mbar_expect_tx(&bar[WB_W1], 24576);
        tma_bulk_g2s(sm + G::BH, p.w1img, 24576, &bar[WB_W1]);
    }
    if (warp == 0) tmem_alloc(reinterpret_cast<uint32_t *>(sm + G::TMEMP), 256);     // stage g: Hacc [128 g, +64), S [128 g + 64 + 16 g, +16)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t *>(sm + G::TMEMP), 0);
    grid_dep_wait();                                             // everything above ran under the previous kernel's tail
    // the first two tile entries of this thread's group are fetched together with the level's tile count (one L2 round trip less on the
    // launch ramp); entries at or past the count are ignored below
    const int g_early = (tid >> 7) & 1, list_cap = p.B * ((p.cap + 127) >> 7);
    const int i0 = (int)blockIdx.x + g_early * (int)gridDim.x, i2 = i0 + 2 * (int)gridDim.x;
    const int raw0 = i0 < list_cap ? __ldg(p.tile_list + i0) : -1, raw2 = i2 < list_cap ? __ldg(p.tile_list + i2) : -1;
    const int ntiles = __shfl_sync(0xffffffffu, *(volatile const int32_t *)(p.tile_count + level), 0);
    const int n_my = ntiles > (int)blockIdx.x ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    const int g = warp >> 2, wq = warp & 3, gtid = tid & 127;             // group = pipeline stage, warp within the group, row of the tile
    const bool issuer1 = wq == 1, issuer2 = wq == 2;             // the WARPS issuing the first / second MMA chain of the group's tiles (one elected lane each)
    const uint32_t tmem_lane = (uint32_t)(wq * 32) << 16;
    const uint32_t td = tmem_base + g * 128, tm = td + tmem_lane;
    const float scale2 = p.scale * 1.4426950408889634f, inv_T = 1.0f / (float)p.T;
    // operand descriptors: K-major; X tiles SWIZZLE_128B (SBO 1024, LBO unused), everything else no swizzle (SBO 128)
    auto nsdesc = [&](uint32_t off, uint32_t lbo) -> uint64_t {
        return ((uint64_t)(0x4000u | (128u >> 4)) << 32) | (uint64_t)(((sbase + off) >> 4) | ((lbo >> 4) << 16));
    };
    auto swdesc = [&](uint32_t off) -> uint64_t {
        return ((uint64_t)2 << 61) | ((uint64_t)(0x4000u | (1024u >> 4)) << 32) | (uint64_t)(((sbase + off) >> 4) | (1u << 16));
    };
    auto group_sync = [&]() {
        if (g == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
        else asm volatile("bar.sync 2, 128;" ::: "memory");
    };
    // tile_raw: the load only (per lane, in flight); uni(): the same value made warp-uniform for the compiler at the point of use
    auto tile_raw = [&](int t) -> int { return t < n_my ? __ldg(p.tile_list + blockIdx.x + t * gridDim.x) : -1; };
    auto uni = [&](int v) -> int { return __shfl_sync(0xffffffffu, v, 0); };
    auto codes_of = [&](int tile) -> int4 {                      // this lane's 4 candidate codes of the tile (lanes 0..7 of a warp)
        int4 c = make_int4(0, 0, 0, 0);
        if (tile >= 0) {
            const int mine = (tile & 255) + 1 - wq * 32;
            if (4 * lane < mine) c = __ldg(reinterpret_cast<const int4 *>(p.code[slot] + (size_t)(tile >> 10) * p.cap + ((tile >> 8) & 3) * 128 + wq * 32) + lane);
        }
        return c;                                                // still in flight: first touched in refill_x
    };
    // the gathered operand of tile `tile` -> WB_XFULL[g]: this warp's 32 candidate rows, one tile::gather4 pair (hi, lo) per 4 rows.  The
    // barrier's expect_tx also counts the bytes refill_k sends later.  Everything goes through the TMA: no register staging, no proxy fence.
    auto refill_x = [&](int tile, int4 c) {                      // whole warp, converged; lane i < 8 holds the codes of rows 4 i .. 4 i + 3
        const int nr = (tile & 255) + 1;
        const int mine = nr - wq * 32 < 32 ? nr - wq * 32 : 32;              // rows of this warp (may be <= 0)
        const int nl = mine > 0 ? (mine + 3) >> 2 : 0;
        const uint32_t kbytes = (dbg & 8) ? 0u : (wq == 0 ? 1024u + 80u : 1024u);
        const bool leader = elect_one();
        if (leader) mbar_expect_tx(&bar[WB_XFULL + g], ((dbg & 1) ? 0u : (uint32_t)nl * 1024u) + kbytes);   // includes refill_k's bytes
        if (!(dbg & 1)) {
            const uint32_t woff = (uint32_t)(g * G::X_STAGE + wq * 32 * 128);
            const uint32_t xh = sbase + G::XH + woff, xl = sbase + G::XL + woff, xb = smem_u32(&bar[WB_XFULL + g]);
            if (4 * lane + 1 >= mine) c.y = c.x;                 // rows past the tile's last one re-fetch a valid row (their scores are never read)
            if (4 * lane + 2 >= mine) c.z = c.x;
            if (4 * lane + 3 >= mine) c.w = c.x;
            if (nl == 8) {                                       // a full warp quarter (every tile but a user's last): straight-line issue
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int cx = __shfl_sync(0xffffffffu, c.x, i), cy = __shfl_sync(0xffffffffu, c.y, i);
                    const int cz = __shfl_sync(0xffffffffu, c.z, i), cw = __shfl_sync(0xffffffffu, c.w, i);
                    if (leader) tma_gather4_hilo(xh + i * 512, xl + i * 512, &tmap, xb, cx, cy, cz, cw);
                }
            } else {
#pragma unroll 1
                for (int i = 0; i < nl; i++) {
                    const int cx = __shfl_sync(0xffffffffu, c.x, i), cy = __shfl_sync(0xffffffffu, c.y, i);
                    const int cz = __shfl_sync(0xffffffffu, c.z, i), cw = __shfl_sync(0xffffffffu, c.w, i);
                    if (leader) tma_gather4_hilo(xh + i * 512, xl + i * 512, &tmap, xb, cx, cy, cz, cw);
                }
            }
        }
        __syncwarp();
    };
    // the small operands of the same tile (counted by refill_x's expect_tx): a quarter of the user's K rows per warp (256-byte bulk copies
    // into rows 64 + 16 g .. of the B operand) and, first warp, the softmax mask + flags.  Issued while the second chain of the
    // current tile runs: they land long before the gathered rows do.
    auto refill_k = [&](int tile, int aslot) {
        if (!(dbg & 8)) {
            const unsigned char *uop = p.uop + (size_t)(tile >> 10) * G::UOP_BYTES;
            if (elect_one()) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int q = wq * 4 + i, kc = q & 7;                    // piece q: hi (q < 8) / lo, k-chunk kc: 16 rows x 16 B
                    tma_bulk_g2s(sm + (q >> 3 ? G::BL : G::BH) + kc * G::B_LBO + (64 + 16 * g) * 16, uop + q * 256, 256, &bar[WB_XFULL + g]);
                }
                if (wq == 0) tma_bulk_g2s(sm + G::ADDV + aslot * G::ADDV_STAGE, uop + 8192, 80, &bar[WB_XFULL + g]);
            }
            __syncwarp();
        }
    };
    auto refill_h = [&](int tile) {                              // H operand of the tile's user -> WB_HFULL[g]
        if (wq == 3) {
            if (elect_one()) {
                mbar_expect_tx(&bar[WB_HFULL + g], 4096);
                tma_bulk_g2s(sm + G::HH + g * G::H_STAGE, p.uop + (size_t)(tile >> 10) * G::UOP_BYTES + 4096, 4096, &bar[WB_HFULL + g]);
            }
            __syncwarp();
        }
    };
    auto issue_m1 = [&](int t) {                                 // first chain of tile t (stage g): called by ONE WARP of the group, converged
        mbar_wait(&bar[WB_XFULL + g], (t >> 1) & 1);
        tc_fence_after();
        const uint64_t dXh = swdesc(G::XH + g * G::X_STAGE), dXl = swdesc(G::XL + g * G::X_STAGE);
        const uint64_t dBh = nsdesc(G::BH, G::B_LBO), dBl = nsdesc(G::BL, G::B_LBO);
        if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
                if (dbg & 16) break;
                const uint64_t ah = dXh + ks * 2, al = dXl + ks * 2;
                const uint64_t bh = dBh + ks * (2 * G::B_LBO / 16), bl = dBl + ks * (2 * G::B_LBO / 16);
                umma_bf16(td, ah, bh, kIdescBf16M128N96, ks > 0);
                umma_bf16(td, ah, bl, kIdescBf16M128N96, 1);
                umma_bf16(td, al, bh, kIdescBf16M128N96, 1);
            }
            umma_commit(&bar[WB_M1 + g]);
        }
        __syncwarp();
    };

#ifdef DMG_WAVE_TIMING
    long long wacc[12] = {0}, wlast = clock64();
#endif
    int cur = uni(g < n_my ? raw0 : -1), nxt = uni(g + 2 < n_my ? raw2 : -1);
    if (cur >= 0) {
        refill_x(cur, codes_of(cur));
        refill_k(cur, g);
        refill_h(cur);
        if (issuer1) { mbar_wait(&bar[WB_W1], 0); issue_m1(g); }
        __syncwarp();
    }
    int4 cn = codes_of(nxt);                                     // candidate codes of the NEXT tile, always one tile ahead
    for (int t = g; t < n_my; t += 2) {
        const uint32_t par = (t >> 1) & 1;
        const int nr = (cur & 255) + 1;
        const bool active = wq * 32 < nr;
        WTICK(0);
        mbar_wait(&bar[WB_XFULL + g], par);                      // acquire the first warp's mask / flags of this tile
        WTICK(1);
        mbar_wait(&bar[WB_M1 + g], par);
        tc_fence_after();
        WTICK(2);
        if (nxt >= 0) refill_x(nxt, cn);                         // X[g] and K slot g are free: the first chain of tile t has completed
        WTICK(3);
        if (active) {
            float sc[16];
            tmem_ld16(tm + 64 + 16 * g, sc);
            const float *sAddv = reinterpret_cast<const float *>(sm + G::ADDV + (t & 3) * G::ADDV_STAGE);
            if (dbg & 2) {
            } else if (reinterpret_cast<const int *>(sAddv)[16] & WU_ALLMASK) {
#pragma unroll
                for (int j = 0; j < 16; j++) sc[j] = j < p.T ? inv_T : 0.0f;
            } else {
                float mx = -3.4028234663852886e+38f;
#pragma unroll
                for (int j = 0; j < 16; j++) { sc[j] = fmaf(sc[j], scale2, sAddv[j]); mx = fmaxf(mx, sc[j]); }
                float sum = 0.0f;
#pragma unroll
                for (int j = 0; j < 16; j++) { sc[j] = ex2_approx(sc[j] - mx); sum += sc[j]; }
                const float inv = 1.0f / sum;
#pragma unroll
                for (int j = 0; j < 16; j++) sc[j] *= inv;
            }
            sc[15] = 1.0f;
            unsigned char *ph = sm + G::PH + g * G::P_STAGE + gtid * 16, *pl = sm + G::PL + g * G::P_STAGE + gtid * 16;
            uint4 hi, lo;
            split8(*reinterpret_cast<float(*)[8]>(&sc[0]), hi, lo);
            *reinterpret_cast<uint4 *>(ph) = hi;
            *reinterpret_cast<uint4 *>(pl) = lo;
            split8(*reinterpret_cast<float(*)[8]>(&sc[8]), hi, lo);
            *reinterpret_cast<uint4 *>(ph + 2048) = hi;
            *reinterpret_cast<uint4 *>(pl + 2048) = lo;
            fence_proxy_async();
        }
        WTICK(4);
        tc_fence_before();
        group_sync();                                            // the group's P rows are complete
        WTICK(5);
        if (issuer2) {                                           // second chain of tile t
            mbar_wait(&bar[WB_HFULL + g], par);
            tc_fence_after();
            const uint64_t dPh = nsdesc(G::PH + g * G::P_STAGE, 2048), dPl = nsdesc(G::PL + g * G::P_STAGE, 2048);
            const uint64_t dHh = nsdesc(G::HH + g * G::H_STAGE, 1024), dHl = nsdesc(G::HL + g * G::H_STAGE, 1024);
            if (elect_one()) {
                if (!(dbg & 16)) {
                    umma_bf16(td, dPh, dHh, kIdescBf16M128N64, 1);
                    umma_bf16(td, dPh, dHl, kIdescBf16M128N64, 1);
                    umma_bf16(td, dPl, dHh, kIdescBf16M128N64, 1);
                }
                umma_commit(&bar[WB_M2 + g]);
            }
        }
        __syncwarp();
        const int nx2 = tile_raw(t + 4);                         // the list entry two tiles ahead: its L2 round trip hides under the second chain
        if (nxt >= 0) refill_k(nxt, (t + 2) & 3);                // under the second chain
        WTICK(6);
        mbar_wait(&bar[WB_M2 + g], par);
        tc_fence_after();
        WTICK(7);
        const int nn = uni(nx2);
        cn = codes_of(nn);                                       // candidate codes of tile t + 4: in flight until the next iteration's refill_x
        if (nxt >= 0) refill_h(nxt);                             // H[g] is free: the second chain of tile t has completed
        float h0[32], h1[32];
        if (active) tmem_ld32x2(tm, h0, h1);
        WTICK(8);
        tc_fence_before();
        group_sync();                                            // the group has drained its accumulators
        WTICK(9);
        if (issuer1 && nxt >= 0) issue_m1(t + 2);                // the next tile's first chain runs under the epilogue below
        __syncwarp();
        WTICK(10);
        if (active && !(dbg & 4)) {
            // four interleaved chains: term o sees 18 - o/4 roundings, inside the (64 - o) + 2 the bound allows the fast path
            float l0 = 0.0f, l1 = 0.0f, l2 = 0.0f, l3 = 0.0f;
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                l0 = fmaf(fmaxf(h0[c], 0.0f), w.w2[c], l0);
                l1 = fmaf(fmaxf(h0[c + 1], 0.0f), w.w2[c + 1], l1);
                l2 = fmaf(fmaxf(h0[c + 2], 0.0f), w.w2[c + 2], l2);
                l3 = fmaf(fmaxf(h0[c + 3], 0.0f), w.w2[c + 3], l3);
            }
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                l0 = fmaf(fmaxf(h1[c], 0.0f), w.w2[32 + c], l0);
                l1 = fmaf(fmaxf(h1[c + 1], 0.0f), w.w2[33 + c], l1);
                l2 = fmaf(fmaxf(h1[c + 2], 0.0f), w.w2[34 + c], l2);
                l3 = fmaf(fmaxf(h1[c + 3], 0.0f), w.w2[35 + c], l3);
            }
            if (gtid < nr) p.score[(size_t)(cur >> 10) * p.cap + ((cur >> 8) & 3) * 128 + gtid] = ((l0 + l1) + (l2 + l3)) + w.b2;
        }
        cur = nxt; nxt = nn;
        WTICK(11);
    }
#ifdef DMG_WAVE_TIMING
    if (p.stats && gtid == 0 && g == 0)
        for (int i = 0; i < 12; i++) atomicAdd(&p.stats[32 + i], (unsigned long long)wacc[i]);
#endif
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ---- K3: strict verification of the deferred cuts + topk, three throughput kernels ---------------------------------------
//   wave_final_prep_kernel    one warp per user: consumed / empty-slot filter, the band of the topk-th fast score, the list of rows
//                             that need strict scores (band rows deferred at the cuts, then the topk candidates) in 16-row chunks
//   wave_strict_rows_kernel   one CTA per chunk: the strict scorer (sequential-k fma chains: the oracle's bits) on <= 16 rows
//   wave_final_verify_kernel  one warp per user: proofs of the deferred cuts, topk by the reference's (score desc, position asc)
//                             order (Recommender.scala:37, TDM.scala:21), outputs, users for the strict kernel
struct WaveFinal {
    int32_t *row_code; float *row_strict;       // [B][RCAP]: deferred band rows first, then the topk candidates
    int32_t *fin_pos;                           // [B][MAX_FINAL] candidate positions of the topk candidates
    int32_t *meta;                              // [B][4]: rows, topk candidates, kk, redo reason + 1 (0 = none)
    int32_t *chunk_list, *chunk_count;          // user << 5 | chunk
    static constexpr int RCAP = FastGeo::VCAP + FastGeo::MAX_FINAL;
};

template <int NJ>
static __global__ void __launch_bounds__(128) wave_final_prep_kernel(const WaveParams p, const BeamParams<float> bp, const WaveFinal wf, int slot)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int user = blockIdx.x * 4 + warp;
    grid_dep_launch();
    grid_dep_wait();
    if (user >= p.B) return;
    const uint32_t lt = (1u << lane) - 1u;
    const WaveUser st = p.user[user];
    int32_t *meta = wf.meta + (size_t)user * 4;
    int why = (st.flags & WU_REDO) ? st.redo_why + 1 : 0;
    const int beam = p.beam_user ? __ldg(p.beam_user + user) : p.beam;
    const int s_level = 31 - __clz(beam);
    if (!why && s_level == p.leaf_level) why = 3;                  // nothing is scored (beam >= 2^leaf_level): every score ties -> strict kernel
    int n_rows = 0, na = 0, kk = 0;
    if (!why) {
        const int count = (s_level < p.leaf_level) ? p.count[user] : 0;
        const int32_t *cur = p.code[slot] + (size_t)user * p.cap;
        const float *sc = p.score + (size_t)user * p.cap;
        const int64_t leaf_start = ((int64_t)1 << p.leaf_level) - 1;
        const int64_t c0 = bp.cons_off ? bp.cons_off[user] : 0, c1 = bp.cons_off ? bp.cons_off[user + 1] : 0;
        float f[NJ];
        uint32_t key[NJ];
        int32_t code[NJ];
        int valid = 0;
        uint32_t mn = 0xffffffffu, mx = 0u;
#pragma unroll
        for (int j = 0; j < NJ; j++) {
            const int i = 32 * j + lane;
            bool keep = false;
            code[j] = 0; f[j] = 0.0f;
            if (i < count) {
                code[j] = cur[i]; f[j] = sc[i];
                const int64_t sl = (int64_t)code[j] - leaf_start;
                const int32_t item = (sl >= 0 && sl < ((int64_t)1 << p.leaf_level)) ? __ldg(bp.leaf_item + sl) : -1;
                keep = item >= 0;
                for (int64_t q = c0; q < c1 && keep; q++) keep = (__ldg(bp.cons + q) != item);
            }
            key[j] = keep ? order_key(f[j]) : 0u;
            valid += __popc(__ballot_sync(0xffffffffu, keep));
            mn = min(mn, key[j] ? key[j] : 0xffffffffu); mx = max(mx, key[j]);
        }
        kk = valid < bp.topk ? valid : bp.topk;
        if (kk > 0 && !(st.flags & WU_SCORED)) why = 3;
        if (kk > 0 && !why) {
            mn = __reduce_min_sync(0xffffffffu, mn);
            mx = __reduce_max_sync(0xffffffffu, mx);
            uint32_t kdn, kup;
            int iters;
            warp_select<NJ>(key, kk, mn, mx, valid, kdn, kup, iters);
            const float dn = key_to_float(kdn) - (2.0f * st.eps * 1.0001f + 1e-30f);
            int32_t *rc = wf.row_code + (size_t)user * WaveFinal::RCAP + st.vcount;
            int32_t *fp = wf.fin_pos + (size_t)user * FastGeo::MAX_FINAL;
#pragma unroll
            for (int j = 0; j < NJ; j++) {
                const bool c = key[j] != 0u && !(f[j] < dn);
                const uint32_t m = __ballot_sync(0xffffffffu, c);
                const int e = na + __popc(m & lt);
                if (c && e < FastGeo::MAX_FINAL) { rc[e] = code[j]; fp[e] = 32 * j + lane; }
                na += __popc(m);
            }
            if (na > FastGeo::MAX_FINAL) why = 4;
        }
        if (!why) {
            for (int e = lane; e < st.vcount; e += 32) wf.row_code[(size_t)user * WaveFinal::RCAP + e] = p.v_code[(size_t)user * WaveGeo::VCAP + e];
            n_rows = st.vcount + na;
            const int nc = (n_rows + 15) >> 4;
            if (lane < nc) {
                int base = 0;
                if (lane == 0) base = atomicAdd(wf.chunk_count, nc);
                base = __shfl_sync(__activemask(), base, 0);
                wf.chunk_list[base + lane] = (user << 5) | lane;
            }
        }
    }
    if (lane == 0) { meta[0] = n_rows; meta[1] = na; meta[2] = kk; meta[3] = why; }
}

// Persistent: weights in shared memory once per CTA, then user after user -- history tile, <= 64 listed rows per pass, the strict
// kernel's tile scorer (score_tile: 4 x 4 register tiles of sequential-k fma chains, beam_kernels.cuh).
struct WaveStrictGeo {
    using G = Geo<float, 64, 64>;
    static constexpr int RT = 64;
    static size_t smem_bytes() { return (size_t)(3 * 64 * 64 + 2 * 64 + kMaxT * 64 + 2 * RT * G::LD + RT * G::PLD + RT) * 4 + kMaxT * 4 + 64; }
};
static __global__ void __launch_bounds__(kThreads, 2) wave_strict_rows_kernel(const WaveParams p, const WaveStrictW sw, const WaveFinal wf)
{
    using SG = WaveStrictGeo;
    using G = SG::G;
    constexpr int E = 64, RT = SG::RT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sWattT = reinterpret_cast<float *>(smem_raw), *sW1T = sWattT + E * E, *sB1 = sW1T + 2 * E * E, *sW2 = sB1 + E;
    float *sK = sW2 + E, *sX = sK + kMaxT * E, *sA = sX + RT * G::LD, *sP = sA + RT * G::LD, *sOut = sP + RT * G::PLD;
    int32_t *sMask = reinterpret_cast<int32_t *>(sOut + RT);
    const int tid = threadIdx.x;
    grid_dep_launch();
    for (int i = tid; i < E * E; i += kThreads) sWattT[i] = __ldg(sw.wattT + i);     // weights: constant for the whole chain
    for (int i = tid; i < 2 * E * E; i += kThreads) sW1T[i] = __ldg(sw.w1T + i);
    if (tid < E) { sB1[tid] = __ldg(sw.b1 + tid); sW2[tid] = __ldg(sw.w2 + tid); }
    grid_dep_wait();
    for (int user = blockIdx.x; user < p.B; user += gridDim.x) {
        const int n_rows = wf.meta[(size_t)user * 4];
        if (n_rows == 0) continue;
        const uint32_t mb = p.user[user].maskbits;
        __syncthreads();
        for (int i = tid; i < kMaxT * 16; i += kThreads) {         // history rows (fp32), zero rows for padding
            const int j = i >> 4, v = i & 15;
            const int c = j < p.T ? p.hist[(size_t)user * p.T + j] : -1;
            if (c >= 0) cp_async16(sK + j * E + v * 4, p.emb + (size_t)c * E + v * 4);
            else *reinterpret_cast<float4 *>(sK + j * E + v * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (tid < kMaxT) sMask[tid] = (mb >> tid) & 1u;
        const int32_t *rc = wf.row_code + (size_t)user * WaveFinal::RCAP;
        float *rs = wf.row_strict + (size_t)user * WaveFinal::RCAP;
        for (int r0 = 0; r0 < n_rows; r0 += RT) {
            const int nr = n_rows - r0 < RT ? n_rows - r0 : RT;
            if (r0) __syncthreads();
            gather_tile<float, E>(sX, p.emb, rc + r0, nr);            // row stride G::LD does not depend on the tile height
            cp_async_commit();
            cp_async_wait<0>();
            __syncthreads();
            score_tile<float, E, RT>(sX, sA, sP, sK, sMask, sWattT, sW1T, sB1, sW2, sw.b2, p.scale, p.T, nr, sOut);
            for (int i = tid; i < nr; i += kThreads) rs[r0 + i] = sOut[i];
        }
    }
}

static __global__ void __launch_bounds__(128) wave_final_verify_kernel(const WaveParams p, const BeamParams<float> bp, const WaveFinal wf, int slot)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int user = blockIdx.x * 4 + warp;
    grid_dep_launch();
    grid_dep_wait();
    if (user >= p.B) return;
    const int32_t *meta = wf.meta + (size_t)user * 4;
    const int n_rows = meta[0], na = meta[1], kk = meta[2];
    int why = meta[3];
    float ratio = 0.0f;
    if (!why) {
        const WaveUser st = p.user[user];
        const int vcount = st.vcount;
        const float *rs = wf.row_strict + (size_t)user * WaveFinal::RCAP;
        const int64_t leaf_start = ((int64_t)1 << p.leaf_level) - 1;
        int bad = 0;
        for (int e = lane; e < vcount; e += 32) {                  // the deferred cuts: the strict order must pick the same rows
            const uint32_t m = p.v_meta[(size_t)user * WaveGeo::VCAP + e];
            const int s0 = m & 255u, n = (m >> 8) & 255u, need = (m >> 16) & 255u, chosen = (m >> 24) & 1u;
            const float eps_e = p.v_segeps[(size_t)user * 32 + (m >> 25)];
            const float mine = rs[e];
            if (eps_e > 0.0f && eps_e < 1e30f) ratio = fmaxf(ratio, fabsf(mine - p.v_fast[(size_t)user * WaveGeo::VCAP + e]) / eps_e);
            const uint32_t ks = order_key(mine);
            int rank = 0, ties = 0;
            for (int q = s0; q < s0 + n; q++) {
                const uint32_t kq = order_key(rs[q]);
                rank += kq > ks ? 1 : 0;
                ties += (kq == ks && q != e) ? 1 : 0;
            }
            bad |= (rank < need && rank + ties >= need) ? 1 : 0;   // exact strict tie across the cut: the reference decides by position
            bad |= ((rank + ties < need ? 1 : 0) != chosen) ? 2 : 0;
        }
        const int32_t *fp = wf.fin_pos + (size_t)user * FastGeo::MAX_FINAL;
        const int32_t *cur = p.code[slot] + (size_t)user * p.cap;
        const float *sc = p.score + (size_t)user * p.cap;
        for (int t = lane; t < na; t += 32) {
            const float mine = rs[vcount + t];
            const int ps = fp[t];
            if (st.eps > 0.0f && st.eps < 1e30f) ratio = fmaxf(ratio, fabsf(mine - sc[ps]) / st.eps);
            const uint32_t ks = order_key(mine);
            int rank = 0, ties = 0;
            for (int q = 0; q < na; q++) {
                const uint32_t kq = order_key(rs[vcount + q]);
                rank += kq > ks ? 1 : 0;
                ties += (kq == ks && q != t) ? 1 : 0;
            }
            bad |= (ties > 0 && rank < kk) ? 1 : 0;                // a tie that reaches the output: its order is positional
            if (rank < kk) {
                bp.out_items[(size_t)user * bp.out_stride + rank] = __ldg(bp.leaf_item + ((int64_t)cur[ps] - leaf_start));
                bp.out_scores[(size_t)user * bp.out_stride + rank] = mine;
            }
        }
        const uint32_t anybad = __ballot_sync(0xffffffffu, bad != 0), proof = __ballot_sync(0xffffffffu, (bad & 2) != 0);
        if (anybad) why = proof ? 6 : 5;
        if (!why) {
            for (int i = kk + lane; i < bp.topk; i += 32) {
                bp.out_items[(size_t)user * bp.out_stride + i] = -1;
                bp.out_scores[(size_t)user * bp.out_stride + i] = 0.0f;
            }
            if (lane == 0) bp.out_counts[user] = kk;
        }
    }
    if (why && lane == 0) {
        p.redo_list[atomicAdd(p.redo_count, 1)] = user;
        *reinterpret_cast<volatile int32_t *>(p.host_flags + 1) = 1;
        if (p.stats) { atomicAdd(&p.stats[5], 1ull); atomicAdd(&p.stats[24 + ((why - 1) & 7)], 1ull); }
    }
    if (p.stats) {
        for (int o = 16; o > 0; o >>= 1) ratio = fmaxf(ratio, __shfl_xor_sync(0xffffffffu, ratio, o));
        if (lane == 0 && ratio > 0.0f) atomicMax(reinterpret_cast<unsigned int *>(&p.stats[4]), __float_as_uint(ratio));
        if (lane == 0 && n_rows) atomicAdd(&p.stats[2], (unsigned long long)n_rows);
    }
}

}  // namespace dmg
