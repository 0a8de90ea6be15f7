// beam_fast.cuh -- tensor-core ("fast") variant of the persistent beam-search kernel, E = 64, fp32 model.
//
// Same search, same outputs as beam_search_kernel<float,64> (ids and scores bit-identical to the
// strict path / CPU oracle), but the two dense contractions of the DIN scorer
//     att = a . W_att^T   (128 x 64 x 64)          h = [x | att] . W1^T   (128 x 128 x 64)
// run on the 5th-generation tensor cores: tcgen05.mma (kind::f16, bf16 operands staged in shared
// memory in the canonical K-major no-swizzle layout, fp32 accumulators in TMEM, read back with
// tcgen05.ld).  fp32 operands are split into bf16 hi + lo parts and three products are
// accumulated (hi*hi + hi*lo + lo*hi), giving ~2^-16 relative accuracy per product.
//
// Exactness is restored by CERTIFIED CUTS: every row carries a rigorous bound eps on
// |fast - strict| (eps_row = alpha*|x| + beta*|a| + gamma from the weight norms);
// at a beam cut only candidates whose fast score lies within 2*eps of the cut can be on the wrong
// side, and exactly those are re-scored with the strict sequential-fma scorer (score_tile) and
// re-ranked with the reference's (score desc, position asc) key.  The attention contractions
// (scores = X.K^T, a = P.K) run on the tensor cores too (N = 16 / K = 16 MMAs against per-user
// bf16 copies of the history tile); the softmax stays in fp32 registers with the spec'd exp.
// The final topk is always re-scored strictly, so returned logits are the oracle's bits.
#pragma once
#include <cuda_bf16.h>

#include "beam_kernels.cuh"

namespace dmg {

struct FastExtra {
    // eps_row = alpha*|x| + beta*(|a| + da) + gamma + zeta*da,  da = Kmax*(2.1*cs*|x|*Kmax + ca)
    float alpha, beta, gamma, zeta, cs, ca;
    const float *watt, *w1;         // row-major [out][in] fp32 (converted to bf16 hi/lo per CTA)
    unsigned long long *stats;      // [0] cuts, [1] cuts needing a strict re-score, [2] rows re-scored, [3] rows scored, [4] max |fast-strict|/eps (float bits)
};

// ---- tcgen05 / TMEM wrappers ------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16])
{
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   canonical layout ((8,n),2):((16 B, SBO),(LBO)) -- 8 rows x 16 B core matrices,
//   SBO = byte stride between 8-row groups, LBO = byte stride between the two 16-byte K chunks.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                       // version = 1 (Blackwell)
    return d;                                     // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE
}
// instruction descriptor: D fp32, A/B bf16, both K-major, M = 128, N = 64
constexpr uint32_t kIdescBf16M128N64 = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t kIdescBf16M128N16 = (1u << 4) | (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);

// fp32 -> bf16 hi + bf16 lo (round to nearest), packed pairs
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16 &hi, __nv_bfloat16 &lo)
{
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(__fsub_rn(x, __bfloat162float(hi)));
}
__device__ __forceinline__ void split8(const float (&v)[8], uint4 &hi, uint4 &lo)
{
    __nv_bfloat16 h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; i++) split_bf16(v[i], h[i], l[i]);
    hi = *reinterpret_cast<uint4 *>(h);
    lo = *reinterpret_cast<uint4 *>(l);
}

struct FastGeo {
    static constexpr int E = 64, R = 128, LD = 68, PLD = kMaxT + 1;
    static constexpr int A_BYTES = R * E * 2;                 // one bf16 operand tile 128 x 64 = 16 KB
    static constexpr int A_LBO = R * 16, B_LBO = E * 16, SBO = 128;
    static constexpr int WATT_BYTES = E * E * 2, W1_BYTES = E * 2 * E * 2;
    static constexpr int KB_BYTES = 16 * E * 2, KB_LBO = 16 * 16;      // history as B operand of the scores MMA: [16 j][64 k]
    static constexpr int KT_BYTES = E * 16 * 2, KT_LBO = E * 16;       // transposed history, B operand of a = P.K: [64 e][16 j]
    static constexpr int P_BYTES = R * 16 * 2;                         // probabilities as A operand: [128][16 j]
    static size_t smem_bytes(int cap, int capp)
    {
        size_t b = 0;
        b += 2 * (size_t)WATT_BYTES + 2 * (size_t)W1_BYTES;   // bf16 hi/lo weights
        b += 4 * (size_t)A_BYTES;                             // X hi/lo, a|att hi/lo (aliased by the strict scratch: sA + sP)
        b += 2 * (size_t)KB_BYTES + 2 * (size_t)KT_BYTES + 2 * (size_t)P_BYTES;
        b += (size_t)kMaxT * E * 4;                           // history tile
        b += 2 * (size_t)R * LD * 4;                          // fp32 row tiles (double buffer)
        b += 2 * (size_t)E * 4 + 16;                          // b1, w2, b2
        b += 4 * (size_t)R * 4;                               // |x|, |a|, partial logits (2)
        b += (size_t)cap * 4 * 2;                             // fast scores, strict scores
        b = (b + 15) & ~(size_t)15;
        b += 2 * (size_t)capp * 8;                            // sort keys + scratch keys
        b += (size_t)cap * 2 * 4;                             // candidate codes (ping-pong)
        b += 64 * 4 + 64;                                     // misc ints, mbarriers, tmem address
        return b;
    }
};

__global__ void __launch_bounds__(kThreads, 1) beam_search_fast_kernel(const BeamParams<float> p, const FastExtra fx)
{
    using G = FastGeo;
    using KO = KeyOf<float>;
    constexpr int E = G::E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *sp = smem_raw;
    unsigned char *sWattH = sp; sp += G::WATT_BYTES;
    unsigned char *sWattL = sp; sp += G::WATT_BYTES;
    unsigned char *sW1H = sp; sp += G::W1_BYTES;
    unsigned char *sW1L = sp; sp += G::W1_BYTES;
    unsigned char *sAxH = sp; sp += G::A_BYTES;               // X hi
    unsigned char *sAxL = sp; sp += G::A_BYTES;               // X lo
    unsigned char *sAaH = sp; sp += G::A_BYTES;               // a, then att, hi
    unsigned char *sAaL = sp; sp += G::A_BYTES;               // lo
    float *sStrictA = reinterpret_cast<float *>(sAxH);        // strict scorer scratch [R][LD] aliases the operand tiles
    float *sP = sStrictA + G::R * G::LD;                      // ... followed by its [R][PLD] score/probability scratch
    static_assert((G::R * G::LD + G::R * G::PLD) * 4 <= 4 * G::A_BYTES, "strict scratch must fit in the operand tiles");
    unsigned char *sKbH = sp; sp += G::KB_BYTES;
    unsigned char *sKbL = sp; sp += G::KB_BYTES;
    unsigned char *sKtH = sp; sp += G::KT_BYTES;
    unsigned char *sKtL = sp; sp += G::KT_BYTES;
    unsigned char *sPH = sp; sp += G::P_BYTES;
    unsigned char *sPL = sp; sp += G::P_BYTES;
    float *sK = reinterpret_cast<float *>(sp); sp += kMaxT * E * 4;
    float *sX = reinterpret_cast<float *>(sp); sp += 2 * G::R * G::LD * 4;
    float *sB1 = reinterpret_cast<float *>(sp); sp += E * 4;
    float *sW2 = reinterpret_cast<float *>(sp); sp += E * 4;
    float *sB2 = reinterpret_cast<float *>(sp); sp += 16;
    float *sNx = reinterpret_cast<float *>(sp); sp += G::R * 4;
    float *sNa = reinterpret_cast<float *>(sp); sp += G::R * 4;
    float *sPart = reinterpret_cast<float *>(sp); sp += 2 * G::R * 4;
    float *sScore = reinterpret_cast<float *>(sp); sp += p.cap * 4;
    float *sStrict = reinterpret_cast<float *>(sp); sp += p.cap * 4;
    sp = smem_raw + (((size_t)(sp - smem_raw) + 15) & ~(size_t)15);
    uint64_t *sKey = reinterpret_cast<uint64_t *>(sp); sp += (size_t)p.capp * 8;
    uint64_t *sKey2 = reinterpret_cast<uint64_t *>(sp); sp += (size_t)p.capp * 8;
    int32_t *sCode0 = reinterpret_cast<int32_t *>(sp); sp += p.cap * 4;
    int32_t *sCode1 = reinterpret_cast<int32_t *>(sp); sp += p.cap * 4;
    int32_t *sMisc = reinterpret_cast<int32_t *>(sp); sp += 64 * 4;   // [0..15] hist, [16..31] mask, [32..39] scan, [40] eps bits, [41] tmem
    uint64_t *sBar = reinterpret_cast<uint64_t *>(sp);                // [0] history TMA, [1] MMA

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = p.T;

    // ---- one-time setup: weights -> bf16 hi/lo in UMMA K-major layout, TMEM allocation ----------
    for (int i = tid; i < E * (E / 8); i += kThreads) {           // W_att: 64 outputs x 8 k-chunks
        const int o = i % E, kc = i / E;
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; q++) v[q] = fx.watt[o * E + kc * 8 + q];
        uint4 hi, lo;
        split8(v, hi, lo);
        *reinterpret_cast<uint4 *>(sWattH + kc * G::B_LBO + o * 16) = hi;
        *reinterpret_cast<uint4 *>(sWattL + kc * G::B_LBO + o * 16) = lo;
    }
    for (int i = tid; i < E * (2 * E / 8); i += kThreads) {       // W1: 64 outputs x 16 k-chunks
        const int o = i % E, kc = i / E;
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; q++) v[q] = fx.w1[o * 2 * E + kc * 8 + q];
        uint4 hi, lo;
        split8(v, hi, lo);
        *reinterpret_cast<uint4 *>(sW1H + kc * G::B_LBO + o * 16) = hi;
        *reinterpret_cast<uint4 *>(sW1L + kc * G::B_LBO + o * 16) = lo;
    }
    for (int i = tid; i < E; i += kThreads) { sB1[i] = p.b1[i]; sW2[i] = p.w2[i]; }
    if (tid == 0) { sB2[0] = p.b2[0]; mbar_init(&sBar[0], 1); mbar_init(&sBar[1], 1); }
    if (warp == 0) tmem_alloc(reinterpret_cast<uint32_t *>(&sMisc[41]), 256);   // scores [0,16) a [64,128) att [128,192) h [192,256)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(&sMisc[41]);
    const float b2 = sB2[0];
    uint32_t hist_phase = 0, mma_phase = 0;
    unsigned long long st_cuts = 0, st_recuts = 0, st_rerows = 0, st_rows = 0;

    // operand descriptors (constant per CTA)
    const uint32_t aXH = smem_u32(sAxH), aXL = smem_u32(sAxL), aAH = smem_u32(sAaH), aAL = smem_u32(sAaL);
    const uint32_t bWaH = smem_u32(sWattH), bWaL = smem_u32(sWattL), bW1H = smem_u32(sW1H), bW1L = smem_u32(sW1L);
    const uint32_t bKbH = smem_u32(sKbH), bKbL = smem_u32(sKbL), bKtH = smem_u32(sKtH), bKtL = smem_u32(sKtL);
    const uint32_t aPH = smem_u32(sPH), aPL = smem_u32(sPL);

    // ---- strict scorer on `n` candidate positions listed in sKey2[0..n) (as positions) ----------
    // gathers the rows again, scores them with score_tile (sequential-k fma chains, weights read
    // through the generic path from global memory) and leaves the strict logits in sStrict[pos].
    float st_ratio = 0.0f, eps_now = 0.0f;
    auto strict_rescore = [&](const int32_t *codes, const int *positions_from_keys, int n, const uint64_t *keys) {
        (void)positions_from_keys;
        for (int base = 0; base < n; base += G::R) {
            const int nr = n - base < G::R ? n - base : G::R;
            for (int idx = tid; idx < nr * (E / 4); idx += kThreads) {
                const int r = idx / (E / 4), v = idx % (E / 4);
                const int pos = KO::pos(keys[base + r]);
                cp_async16(sX + r * G::LD + v * 4, p.emb + (size_t)codes[pos] * E + v * 4);
            }
            cp_async_commit();
            cp_async_wait<0>();
            __syncthreads();
            score_tile<float, 64>(sX, sStrictA, sP, sK, sMisc + 16, p.wattT, p.w1T, sB1, sW2, b2, p.scale, T, nr, sPart);
            for (int r = tid; r < nr; r += kThreads) {
                const int pos = KO::pos(keys[base + r]);
                sStrict[pos] = sPart[r];
                if (eps_now > 0.0f) {                              // observed |fast - strict| / eps (must stay < 1)
                    const float ratio = fabsf(sPart[r] - sScore[pos]) / eps_now;
                    if (ratio > st_ratio) st_ratio = ratio;
                }
            }
            __syncthreads();
        }
    };

    for (int user = blockIdx.x; user < p.B; user += gridDim.x) {
        // ---- K2: history tile through the TMA bulk-copy engine ------------------------------------
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
            for (int j = 0; j < T; j++) if (sMisc[j] >= 0) bytes += E * sizeof(float);
            mbar_expect_tx(&sBar[0], bytes);
            for (int j = 0; j < T; j++)
                if (sMisc[j] >= 0) tma_bulk_g2s(sK + j * E, p.emb + (size_t)sMisc[j] * E, E * sizeof(float), &sBar[0]);
        }
        for (int i = tid; i < T * E; i += kThreads)
            if (sMisc[i / E] < 0) sK[i] = 0.0f;
        mbar_wait(&sBar[0], hist_phase);
        hist_phase ^= 1;
        __syncthreads();
        // history tile -> bf16 hi/lo B operands (rows j >= T are zero) and Kmax = max_j |K_j|
        if (tid < 128) {
            const int j = tid & 15, kc = tid >> 4;
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; q++) v[q] = j < T ? sK[j * E + kc * 8 + q] : 0.0f;
            uint4 hi, lo;
            split8(v, hi, lo);
            *reinterpret_cast<uint4 *>(sKbH + kc * G::KB_LBO + j * 16) = hi;
            *reinterpret_cast<uint4 *>(sKbL + kc * G::KB_LBO + j * 16) = lo;
        } else {
            const int t2 = tid - 128, e = t2 & 63, jc = t2 >> 6;
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; q++) v[q] = (jc * 8 + q) < T ? sK[(jc * 8 + q) * E + e] : 0.0f;
            uint4 hi, lo;
            split8(v, hi, lo);
            *reinterpret_cast<uint4 *>(sKtH + jc * G::KT_LBO + e * 16) = hi;
            *reinterpret_cast<uint4 *>(sKtL + jc * G::KT_LBO + e * 16) = lo;
        }
        if (tid == 0) sMisc[44] = 0;
        __syncthreads();
        if (tid < T) {
            float n2 = 0.0f;
            for (int k = 0; k < E; k++) n2 = fmaf(sK[tid * E + k], sK[tid * E + k], n2);
            atomicMax(&sMisc[44], __float_as_int(sqrtf(n2) * 1.0001f));
        }
        __syncthreads();
        const float kmax = __int_as_float(sMisc[44]);

        const int beam = p.beam_user ? p.beam_user[user] : p.beam;
        const int s_level = 31 - __clz(beam);
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
            for (int i = tid; i < count; i += kThreads) sScore[i] = 0.0f;
            __syncthreads();
        }
        float eps_level = 0.0f;                                   // max eps_row of the current candidates

        for (int level = s_level; level < p.leaf_level && count > 0; level++) {
            int nb = count;
            if (count > beam) {
                // ---- certified cut ---------------------------------------------------------------
                int n2 = 2;
                while (n2 < count) n2 <<= 1;
                for (int i = tid; i < n2; i += kThreads) sKey[i] = i < count ? KO::make(sScore[i], i) : KO::lowest();
                __syncthreads();
                bitonic_sort_desc(sKey, n2);
                nb = beam;
                st_cuts++;
                const float pivot = sScore[KO::pos(sKey[beam - 1])];
                const float band = 2.0f * eps_level * 1.0001f + 1e-30f;
                // hi = last rank whose fast score >= pivot - band ; lo = first rank with score <= pivot + band
                int hi_local = -1, lo_local = 1 << 30;
                for (int i = tid; i < count; i += kThreads) {
                    const float s = sScore[KO::pos(sKey[i])];
                    if (!(s < pivot - band)) hi_local = i > hi_local ? i : hi_local;
                    if (!(s > pivot + band)) lo_local = i < lo_local ? i : lo_local;
                }
                if (tid == 0) { sMisc[42] = -1; sMisc[43] = 1 << 30; }
                __syncthreads();
                atomicMax(&sMisc[42], hi_local);
                atomicMin(&sMisc[43], lo_local);
                __syncthreads();
                const int hi = sMisc[42], lo = sMisc[43];
                if (hi >= beam) {                                  // an outsider can overtake an insider: settle it strictly
                    st_recuts++;
                    eps_now = eps_level;
                    const int na = hi - lo + 1;
                    st_rerows += na;
                    strict_rescore(cur, nullptr, na, sKey + lo);
                    int m2 = 2;
                    while (m2 < na) m2 <<= 1;
                    for (int i = tid; i < m2; i += kThreads) {
                        uint64_t k = KO::lowest();
                        if (i < na) { const int pos = KO::pos(sKey[lo + i]); k = KO::make(sStrict[pos], pos); }
                        sKey2[i] = k;
                    }
                    __syncthreads();
                    bitonic_sort_desc(sKey2, m2);
                    for (int i = tid; i < beam - lo; i += kThreads) sKey[lo + i] = sKey2[i];
                    __syncthreads();
                }
                for (int i = tid; i < nb; i += kThreads) nxt[i] = cur[KO::pos(sKey[i])];
                __syncthreads();
                int32_t *t = cur; cur = nxt; nxt = t;
            }
            // ---- children (order preserved) ---------------------------------------------------------
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
            st_rows += count;
            if (tid == 0) sMisc[40] = 0;
            // ---- fast scoring, tile by tile ------------------------------------------------------------
            const int ntiles = (count + G::R - 1) / G::R;
            if (ntiles > 0) {
                gather_tile<float, 64>(sX, p.emb, cur, count < G::R ? count : G::R);
                cp_async_commit();
            }
            for (int t = 0; t < ntiles; t++) {
                const int r0 = t * G::R;
                const int nrows = count - r0 < G::R ? count - r0 : G::R;
                const float *buf = sX + (t & 1) * G::R * G::LD;
                if (t + 1 < ntiles) {
                    const int r1 = r0 + G::R;
                    gather_tile<float, 64>(sX + ((t + 1) & 1) * G::R * G::LD, p.emb, cur + r1, count - r1 < G::R ? count - r1 : G::R);
                    cp_async_commit();
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                __syncthreads();
                // (1) x -> bf16 hi/lo operand tile, |x|
                {
                    const int r = tid >> 1, part = tid & 1;
                    float nx = 0.0f;
                    if (r < nrows) {
                        const float *xr = buf + r * G::LD + part * 32;
#pragma unroll
                        for (int c = 0; c < 4; c++) {
                            float xv[8];
                            ld4(xr + c * 8, *reinterpret_cast<float(*)[4]>(&xv[0]));
                            ld4(xr + c * 8 + 4, *reinterpret_cast<float(*)[4]>(&xv[4]));
#pragma unroll
                            for (int q = 0; q < 8; q++) nx = fmaf(xv[q], xv[q], nx);
                            uint4 hi, lo;
                            const int off = (part * 4 + c) * G::A_LBO + r * 16;
                            split8(xv, hi, lo);
                            *reinterpret_cast<uint4 *>(sAxH + off) = hi;
                            *reinterpret_cast<uint4 *>(sAxL + off) = lo;
                        }
                    }
                    nx += __shfl_xor_sync(0xffffffffu, nx, 1);
                    if (part == 0 && r < nrows) sNx[r] = sqrtf(nx) * 1.0001f;
                }
                fence_proxy_async();
                tc_fence_before();
                __syncthreads();
                // (2) attention scores S = X . K^T  (M 128, N 16, K 64) on the tensor cores
                if (tid == 0) {
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) {
                        const uint64_t ah = umma_desc(aXH + ks * 2 * G::A_LBO, G::A_LBO, G::SBO);
                        const uint64_t al = umma_desc(aXL + ks * 2 * G::A_LBO, G::A_LBO, G::SBO);
                        const uint64_t bh = umma_desc(bKbH + ks * 2 * G::KB_LBO, G::KB_LBO, G::SBO);
                        const uint64_t bl = umma_desc(bKbL + ks * 2 * G::KB_LBO, G::KB_LBO, G::SBO);
                        umma_bf16(tmem_base, ah, bh, kIdescBf16M128N16, ks > 0);
                        umma_bf16(tmem_base, ah, bl, kIdescBf16M128N16, 1);
                        umma_bf16(tmem_base, al, bh, kIdescBf16M128N16, 1);
                    }
                    umma_commit(&sBar[1]);
                }
                mbar_wait(&sBar[1], mma_phase);
                mma_phase ^= 1;
                tc_fence_after();
                // (3) Mask + SoftMax per row in registers (warps 0-3 own the 128 TMEM lanes), P -> bf16 hi/lo
                if (warp < 4) {
                    const int row = warp * 32 + lane;
                    float sc[16];
                    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16), sc);
                    float mx = -3.4028234663852886e+38f;
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        float v = mul_(sc[j], p.scale);
                        if (j < T && sMisc[16 + j]) v = mask_value<float>::get();
                        sc[j] = v;
                        if (j < T) mx = v > mx ? v : mx;
                    }
                    float sum = 0.0f;
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const float e = j < T ? exp_(sub_(sc[j], mx)) : 0.0f;
                        sc[j] = e;
                        sum = add_(sum, e);
                    }
                    const float inv = inv_(sum);
#pragma unroll
                    for (int j = 0; j < 16; j++) sc[j] = mul_(sc[j], inv);
                    uint4 hi, lo;
                    split8(*reinterpret_cast<float(*)[8]>(&sc[0]), hi, lo);
                    *reinterpret_cast<uint4 *>(sPH + row * 16) = hi;
                    *reinterpret_cast<uint4 *>(sPL + row * 16) = lo;
                    split8(*reinterpret_cast<float(*)[8]>(&sc[8]), hi, lo);
                    *reinterpret_cast<uint4 *>(sPH + G::A_LBO + row * 16) = hi;
                    *reinterpret_cast<uint4 *>(sPL + G::A_LBO + row * 16) = lo;
                }
                fence_proxy_async();
                tc_fence_before();
                __syncthreads();
                // (3b) a = P . K  (M 128, N 64, K 16): one k-step
                if (tid == 0) {
                    tc_fence_after();
                    const uint64_t ah = umma_desc(aPH, G::A_LBO, G::SBO), al = umma_desc(aPL, G::A_LBO, G::SBO);
                    const uint64_t bh = umma_desc(bKtH, G::KT_LBO, G::SBO), bl = umma_desc(bKtL, G::KT_LBO, G::SBO);
                    umma_bf16(tmem_base + 64, ah, bh, kIdescBf16M128N64, 0);
                    umma_bf16(tmem_base + 64, ah, bl, kIdescBf16M128N64, 1);
                    umma_bf16(tmem_base + 64, al, bh, kIdescBf16M128N64, 1);
                    umma_commit(&sBar[1]);
                }
                mbar_wait(&sBar[1], mma_phase);
                mma_phase ^= 1;
                tc_fence_after();
                // (3c) a: TMEM -> registers -> bf16 hi/lo operand tile, |a|
                {
                    const int row = (warp & 3) * 32 + lane, cbase = (warp >> 2) * 32;
                    float v[32];
                    tmem_ld32(tmem_base + 64 + ((uint32_t)((warp & 3) * 32) << 16) + cbase, v);
                    float na = 0.0f;
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        float av[8];
#pragma unroll
                        for (int q = 0; q < 8; q++) { av[q] = v[c * 8 + q]; na = fmaf(av[q], av[q], na); }
                        uint4 hi, lo;
                        split8(av, hi, lo);
                        const int off = ((cbase >> 3) + c) * G::A_LBO + row * 16;
                        *reinterpret_cast<uint4 *>(sAaH + off) = hi;
                        *reinterpret_cast<uint4 *>(sAaL + off) = lo;
                    }
                    sPart[(warp >> 2) * G::R + row] = na;
                }
                fence_proxy_async();
                tc_fence_before();
                __syncthreads();
                if (tid < nrows) sNa[tid] = sqrtf(sPart[tid] + sPart[G::R + tid]) * 1.0001f;
                // (4) att = a . Watt^T on the tensor cores: 4 k-steps x (hi*hi + hi*lo + lo*hi)
                if (tid == 0) {
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) {
                        const uint64_t ah = umma_desc(aAH + ks * 2 * G::A_LBO, G::A_LBO, G::SBO);
                        const uint64_t al = umma_desc(aAL + ks * 2 * G::A_LBO, G::A_LBO, G::SBO);
                        const uint64_t bh = umma_desc(bWaH + ks * 2 * G::B_LBO, G::B_LBO, G::SBO);
                        const uint64_t bl = umma_desc(bWaL + ks * 2 * G::B_LBO, G::B_LBO, G::SBO);
                        umma_bf16(tmem_base + 128, ah, bh, kIdescBf16M128N64, ks > 0);
                        umma_bf16(tmem_base + 128, ah, bl, kIdescBf16M128N64, 1);
                        umma_bf16(tmem_base + 128, al, bh, kIdescBf16M128N64, 1);
                    }
                    umma_commit(&sBar[1]);
                }
                mbar_wait(&sBar[1], mma_phase);
                mma_phase ^= 1;
                tc_fence_after();
                // (5) att: TMEM -> registers -> bf16 hi/lo operand tile (overwrites a)
                {
                    const int row = (warp & 3) * 32 + lane, cbase = (warp >> 2) * 32;
                    float v[32];
                    tmem_ld32(tmem_base + 128 + ((uint32_t)((warp & 3) * 32) << 16) + cbase, v);
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        float av[8];
#pragma unroll
                        for (int q = 0; q < 8; q++) av[q] = v[c * 8 + q];
                        uint4 hi, lo;
                        split8(av, hi, lo);
                        const int off = ((cbase >> 3) + c) * G::A_LBO + row * 16;
                        *reinterpret_cast<uint4 *>(sAaH + off) = hi;
                        *reinterpret_cast<uint4 *>(sAaL + off) = lo;
                    }
                }
                fence_proxy_async();
                tc_fence_before();
                __syncthreads();
                // (6) h = [x | att] . W1^T : 8 k-steps (4 on x, 4 on att)
                if (tid == 0) {
                    tc_fence_after();
                    const uint32_t d2 = tmem_base + 192;
#pragma unroll
                    for (int ks = 0; ks < 8; ks++) {
                        const uint32_t aH = ks < 4 ? aXH + ks * 2 * G::A_LBO : aAH + (ks - 4) * 2 * G::A_LBO;
                        const uint32_t aL = ks < 4 ? aXL + ks * 2 * G::A_LBO : aAL + (ks - 4) * 2 * G::A_LBO;
                        const uint64_t ah = umma_desc(aH, G::A_LBO, G::SBO);
                        const uint64_t al = umma_desc(aL, G::A_LBO, G::SBO);
                        const uint64_t bh = umma_desc(bW1H + ks * 2 * G::B_LBO, G::B_LBO, G::SBO);
                        const uint64_t bl = umma_desc(bW1L + ks * 2 * G::B_LBO, G::B_LBO, G::SBO);
                        umma_bf16(d2, ah, bh, kIdescBf16M128N64, ks > 0);
                        umma_bf16(d2, ah, bl, kIdescBf16M128N64, 1);
                        umma_bf16(d2, al, bh, kIdescBf16M128N64, 1);
                    }
                    umma_commit(&sBar[1]);
                }
                mbar_wait(&sBar[1], mma_phase);
                mma_phase ^= 1;
                tc_fence_after();
                // (7) epilogue: h = relu(acc + b1); logit = h . W2 + b2 ; eps_row
                {
                    const int row = (warp & 3) * 32 + lane, cbase = (warp >> 2) * 32;
                    float v[32];
                    tmem_ld32(tmem_base + 192 + ((uint32_t)((warp & 3) * 32) << 16) + cbase, v);
                    float part = 0.0f;
#pragma unroll
                    for (int c = 0; c < 32; c++) {
                        float h = v[c] + sB1[cbase + c];
                        h = h > 0.0f ? h : 0.0f;
                        part = fmaf(h, sW2[cbase + c], part);
                    }
                    sPart[(warp >> 2) * G::R + row] = part;
                }
                tc_fence_before();
                __syncthreads();
                if (tid < nrows) {
                    sScore[r0 + tid] = sPart[tid] + sPart[G::R + tid] + b2;
                    const float ds = fx.cs * sNx[tid] * kmax;                 // bound on the attention-score error
                    const float da = kmax * (2.1f * ds + fx.ca);              // bound on |a_fast - a_strict|
                    float eps = fx.alpha * sNx[tid] + fx.beta * (sNa[tid] + da) + fx.gamma + fx.zeta * da;
                    if (!(ds < 0.01f)) eps = 3.0e38f;                         // outside the linearised regime: certify nothing
                    atomicMax(&sMisc[40], __float_as_int(eps));
                }
                __syncthreads();
            }
            eps_level = __int_as_float(sMisc[40]);
        }

        // ---- K3: topk, always settled with strict scores -----------------------------------------------
        const int64_t leaf_start = ((int64_t)1 << p.leaf_level) - 1;
        const bool at_leaf = (s_level <= p.leaf_level);
        {
            int n2 = 2;
            while (n2 < count) n2 <<= 1;
            const int64_t c0 = p.cons_off ? p.cons_off[user] : 0, c1 = p.cons_off ? p.cons_off[user + 1] : 0;
            for (int i = tid; i < n2; i += kThreads) {
                uint64_t k = KO::lowest();
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
            // valid = entries that survived the filters
            int vloc = 0;
            for (int i = tid; i < count; i += kThreads) vloc += KO::is_lowest(sKey[i]) ? 0 : 1;
            if (tid == 0) sMisc[42] = 0;
            __syncthreads();
            atomicAdd(&sMisc[42], vloc);
            __syncthreads();
            const int valid = sMisc[42];
            const int kk = valid < p.topk ? valid : p.topk;
            int na = 0;
            if (kk > 0) {
                const bool scored = (s_level < p.leaf_level);         // s_level == leaf_level: all scores are the exact 0 of the start level
                if (scored) {
                    const float pivot = sScore[KO::pos(sKey[kk - 1])];
                    const float band = 2.0f * eps_level * 1.0001f + 1e-30f;
                    int hi_local = -1;
                    for (int i = tid; i < valid; i += kThreads)
                        if (!(sScore[KO::pos(sKey[i])] < pivot - band)) hi_local = i > hi_local ? i : hi_local;
                    if (tid == 0) sMisc[43] = -1;
                    __syncthreads();
                    atomicMax(&sMisc[43], hi_local);
                    __syncthreads();
                    na = sMisc[43] + 1;                                // ranks [0, na) may end up in the topk
                    st_rerows += na;
                    eps_now = eps_level;
                    strict_rescore(cur, nullptr, na, sKey);
                } else {
                    na = valid;
                    for (int i = tid; i < na; i += kThreads) sStrict[KO::pos(sKey[i])] = sScore[KO::pos(sKey[i])];
                    __syncthreads();
                }
                int m2 = 2;
                while (m2 < na) m2 <<= 1;
                for (int i = tid; i < m2; i += kThreads) {
                    uint64_t k = KO::lowest();
                    if (i < na) { const int pos = KO::pos(sKey[i]); k = KO::make(sStrict[pos], pos); }
                    sKey2[i] = k;
                }
                __syncthreads();
                bitonic_sort_desc(sKey2, m2);
            }
            for (int i = tid; i < p.topk; i += kThreads) {
                int32_t item = -1;
                float sc = 0.0f;
                if (i < kk) {
                    const int pos = KO::pos(sKey2[i]);
                    item = __ldg(p.leaf_item + ((int64_t)cur[pos] - leaf_start));
                    sc = sStrict[pos];
                }
                p.out_items[(size_t)user * p.out_stride + i] = item;
                p.out_scores[(size_t)user * p.out_stride + i] = sc;
            }
            if (tid == 0) p.out_counts[user] = kk;
        }
        __syncthreads();
    }
    if (fx.stats) {
        for (int o = 16; o > 0; o >>= 1) st_ratio = fmaxf(st_ratio, __shfl_xor_sync(0xffffffffu, st_ratio, o));
        if (lane == 0) atomicMax(reinterpret_cast<unsigned int *>(&fx.stats[4]), __float_as_uint(st_ratio));
        if (tid == 0) {
            atomicAdd(&fx.stats[0], st_cuts); atomicAdd(&fx.stats[1], st_recuts);
            atomicAdd(&fx.stats[2], st_rerows); atomicAdd(&fx.stats[3], st_rows);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}

}  // namespace dmg
