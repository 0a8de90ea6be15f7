// beam_fast.cuh -- tensor-core ("fast") persistent beam-search kernel, E = 64, fp32 model.
//
// Same search and the same outputs as beam_search_kernel<float,64> (ids and logits bit-identical
// to the strict path / CPU oracle).  What changes is how the <= 2*beam candidate rows of a level
// are scored and how the beam is cut:
//
//  * the attention branch of the DIN scorer is linear after the softmax, so it is collapsed per
//    user:   h = W1x.x + W1a.Watt.(sum_j p_j K_j) + b1 = W1x.x + sum_j p_j H_j + b1,
//    H_j = M.K_j,  M = W1a.Watt (64x64, folded on the host in double).  Per candidate row the
//    tensor cores then run two contractions only:
//        S    = X . K^T                      (256 x 16 x 64)    attention scores
//        Hacc = X . W1x^T  +  P . H          (256 x 64 x 80)    hidden layer
//    as tcgen05.mma kind::f16 with bf16 hi/lo split operands (hi*hi + hi*lo + lo*hi, ~2^-16),
//    operands in shared memory in the canonical K-major no-swizzle layout, fp32 accumulators in
//    TMEM, read back with tcgen05.ld for the softmax (registers, one row per thread) and the
//    ReLU / W2 epilogue.
//  * candidate rows are gathered from the node table with coalesced 128-bit loads (16 lanes per
//    256-byte row) straight into registers, split into bf16 hi/lo and stored as UMMA operands --
//    no fp32 staging tile; the user's history rows are staged by the TMA bulk-copy engine.
//  * the beam cut is a 4-pass radix SELECT of the beam-th largest fast score, not a sort.  With a
//    rigorous bound eps on |fast - strict| for the level (DESIGN.md "certified cuts"), rows with
//    fast > pivot + 2 eps are certainly inside the beam, rows with fast < pivot - 2 eps certainly
//    outside; only rows within the band are re-scored by the strict sequential-fma scorer (one
//    warp per row) and ranked by the reference's (score desc, position asc) key.
//  * the order of the beam only matters for tie-breaking.  If two strictly re-scored rows tie
//    exactly at a decision point (or the band is implausibly wide) the user is appended to a
//    redo list and re-run by the strict kernel, which reproduces the reference's stable sorts.
//  * the final topk is always re-scored strictly, so returned logits are the oracle's bits.
//
// Two CTAs of 256 threads per SM (about 110 KB of shared memory and 256 TMEM columns each): while
// one CTA waits on a gather or selects its beam, the other one keeps the tensor/ALU pipes busy.
#pragma once
#include <cuda_bf16.h>

#include "beam_kernels.cuh"

namespace dmg {

struct FastParams {
    float b1[64], w2[64];           // read as constant-bank operands in the epilogue
    float b2;
    const float *mT;                // M~^T [k][o] fp32, M = W1a.Watt (host, double -> fp32)
    const float *w1;                // W1 row-major [o][2E] fp32 (item half -> bf16 hi/lo operand)
    const float *lvl_vx;            // [32] per tree level: max over nodes of sum_k v_k |x_k|
    const float *lvl_nx;            // [32] per tree level: max |x|_2
    const float *zvec;              // [64] z_k = sum_o sum_m |w2_o| |W1a_om| |Watt_mk|
    float cA, cZ, cH, cGamma;       // bound constants: eps = tau (cA VX + cZ ZK + (cH + |dp|_1) HW + cGamma)
    float tau;                      // fraction of the worst-case bound used as the band
    int32_t sparse_from;            // lowest tree level with a missing code: expansion probes the bitmap only from there
    int32_t *redo_list;             // users to re-run with the strict kernel
    int32_t *redo_count;
    int32_t *host_flags;            // mapped pinned memory: [1] is raised with the first redo user, so a synchronous caller launches the
                                    // strict redo kernel only for the batches that need it
    int32_t *work_counter;          // dynamic user scheduler: [0] users below tail_start, [2] tail users, [8 + smid] tail owner of an SM
    int32_t tail_start;             // users >= tail_start (the last, partial round of the batch; == B when unused) run ONE per SM
    unsigned long long *stats;      // [0] cuts [1] cuts re-scored [2] rows re-scored [3] rows scored [4] max ratio bits [5] redo users
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
// instruction descriptor: D fp32, A/B bf16, both K-major, M = 128, N = 64 / 16
constexpr uint32_t kIdescBf16M128N64 = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t kIdescBf16M128N80 = (1u << 4) | (1u << 7) | (1u << 10) | ((80u >> 3) << 17) | ((128u >> 4) << 24);

// fp32 pair -> packed bf16 hi pair and bf16 lo pair (x = hi + lo + O(2^-18 |x|))
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t &hi, uint32_t &lo)
{
    __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
    hi = *reinterpret_cast<uint32_t *>(&h);
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    __nv_bfloat162 l = __floats2bfloat162_rn(__fsub_rn(x0, h0), __fsub_rn(x1, h1));
    lo = *reinterpret_cast<uint32_t *>(&l);
}
__device__ __forceinline__ void split8(const float (&v)[8], uint4 &hi, uint4 &lo)
{
    split_pair(v[0], v[1], hi.x, lo.x);
    split_pair(v[2], v[3], hi.y, lo.y);
    split_pair(v[4], v[5], hi.z, lo.z);
    split_pair(v[6], v[7], hi.w, lo.w);
}
__device__ __forceinline__ float4 ldg_row16(const float *p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float key_to_float(uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}

struct FastGeo {
    static constexpr int E = 64, R = 256, THREADS = 256, KLD = 68;
    static constexpr int X_LBO = R * 16 + 16;                 // +16: the two rows of a warp store hit disjoint banks
    static constexpr int X_BYTES = 8 * X_LBO;                 // one bf16 operand tile 256 x 64 (hi or lo)
    static constexpr int P_LBO = R * 16, P_BYTES = 2 * P_LBO; // probabilities [256][16]
    static constexpr int B_LBO = 80 * 16, B_BYTES = 8 * B_LBO; // B operand [80 n][64 k]: n < 64 W1x outputs, n >= 64 history slots
    static constexpr int H_LBO = E * 16, H_BYTES = 2 * H_LBO;       // H as B operand [64 o][16 j]
    static constexpr int SBO = 128;
    static constexpr int MAX_UNC = 96;                        // widest band re-scored in place at a cut
    static constexpr int MAX_FINAL = 256;                     // widest topk candidate set re-scored in place
    static constexpr int SB = 32;                             // rows per strict re-score batch
    static constexpr int STRICT_SCR = (16 + 3 * SB) * KLD * 4 + SB * 17 * 4;   // history, x, a|h, att, scores
    // aliases inside the P operand region while no MMA is in flight
    static constexpr int VCAP = 160;                          // band rows a user may defer to its end-of-search verification
    static constexpr int OFF_HIST = 0, OFF_SEL = 1024, OFF_KEYU = 2048, OFF_UPOS = 4096, OFF_CLS = 6144,
                         OFF_LCODE = 8192, OFF_LSTR = OFF_LCODE + (VCAP + MAX_FINAL) * 4;
    static size_t smem_bytes(int cap)
    {
        size_t b = 2 * (size_t)X_BYTES + 2 * (size_t)P_BYTES + 2 * (size_t)B_BYTES + 2 * (size_t)H_BYTES;
        b += (size_t)cap * 4 * 3;                             // fast scores, candidate codes (ping-pong)
        b += (size_t)VCAP * 12 + 32 * 4;                      // deferred verification list (code, fast score, meta), eps per segment
        b += 128 * 4 + 64;                                    // misc ints, mbarriers
        return b;
    }
    static constexpr int max_cap() { return 512; }            // OFF_CLS + 512 <= P region, 2 candidates per thread
};
static_assert(2 * FastGeo::P_BYTES >= FastGeo::OFF_LSTR + (FastGeo::VCAP + FastGeo::MAX_FINAL) * 4 && FastGeo::OFF_CLS + 512 <= FastGeo::OFF_LCODE,
              "cut scratch must fit in the P operand region");
static_assert(2 * FastGeo::X_BYTES >= FastGeo::STRICT_SCR, "strict scratch must fit in the X operand region");

// ---- strict scorer on a batch of rows, whole CTA (bit-identical to score_tile / the oracle) ---------
// scr: [history 16 x KLD (already loaded, zero rows for padding) | x SB x KLD | a,h SB x KLD | att SB x KLD | s,p SB x 17]
// Rows with the codes sRow[base .. base+nr) -> sOut[base ..), nr <= 4*RPT.  Every dot product is one sequential-k
// fma chain; thread (o = tid & 63, rg = tid >> 6) owns output column o of rows rg, rg+4, .. (RPT of them) and
// keeps the 64 weights of that column in registers, loaded once per batch.
// NT = threads of the CTA (a multiple of 64): NT / 64 row groups, NB = (NT / 64) * RPT rows per call.
template <int RPT, int NT = FastGeo::THREADS>
__device__ __forceinline__ void strict_score_rows(const float *__restrict__ emb, const int32_t *__restrict__ sRow,
                                                  int base, int nr, float *__restrict__ sOut,
                                                  float *__restrict__ scr, uint32_t maskbits, int T, float scale,
                                                  const float *__restrict__ wattT, const float *__restrict__ w1T,
                                                  const float *__restrict__ b1, const float *__restrict__ w2, float b2)
{
    constexpr int E = 64, LD = FastGeo::KLD, SB = FastGeo::SB, PLD = 17, GS = NT / 64, NB = GS * RPT;
    static_assert(NB * 8 <= NT && NB <= SB, "row batch does not fit the CTA");
    float *sKf = scr, *sXs = sKf + 16 * LD, *sA = sXs + SB * LD, *sT = sA + SB * LD, *sPs = sT + SB * LD;
    const int tid = threadIdx.x, o = tid & 63, rg = tid >> 6;
    float w[E];
#pragma unroll
    for (int k = 0; k < E; k++) w[k] = __ldg(wattT + k * E + o);
    for (int idx = tid; idx < NB * 16; idx += NT) {
        const int r = idx >> 4, c = idx & 15;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nr) v = __ldg(reinterpret_cast<const float4 *>(emb + (size_t)sRow[base + r] * E) + c);
        *reinterpret_cast<float4 *>(sXs + r * LD + c * 4) = v;
    }
    __syncthreads();
    if (tid < NB * 8) {                                         // (1) scores: MatMul transB, Mask
        const int r = tid >> 3, j0 = tid & 7, j1 = j0 + 8;
        const float *xr = sXs + r * LD, *k0 = sKf + j0 * LD, *k1 = sKf + j1 * LD;
        float a0 = 0.0f, a1 = 0.0f;
#pragma unroll 8
        for (int k = 0; k < E; k++) { const float xv = xr[k]; a0 = fma_(xv, k0[k], a0); a1 = fma_(xv, k1[k], a1); }
        if (j0 < T) sPs[r * PLD + j0] = ((maskbits >> j0) & 1u) ? mask_value<float>::get() : mul_(a0, scale);
        if (j1 < T) sPs[r * PLD + j1] = ((maskbits >> j1) & 1u) ? mask_value<float>::get() : mul_(a1, scale);
    }
    __syncthreads();
    if (tid < NB) {                                             // (2) SoftMax
        float *pr = sPs + tid * PLD;
        float mx = pr[0];
        for (int j = 1; j < T; j++) { const float v = pr[j]; mx = (v > mx || v != v) ? v : mx; }
        float sum = 0.0f;
        for (int j = 0; j < T; j++) { const float e = exp_(sub_(pr[j], mx)); pr[j] = e; sum = add_(sum, e); }
        const float inv = inv_(sum);
        for (int j = 0; j < T; j++) pr[j] = mul_(pr[j], inv);
    }
    __syncthreads();
    if (tid < NB * 8) {                                         // (3) a = p . K
        const int r = tid >> 3, kq = tid & 7;
        float acc[8];
#pragma unroll
        for (int q = 0; q < 8; q++) acc[q] = 0.0f;
        for (int j = 0; j < T; j++) {
            const float pj = sPs[r * PLD + j];
            float kv[8];
            ld4(sKf + j * LD + kq * 8, *reinterpret_cast<float(*)[4]>(&kv[0]));
            ld4(sKf + j * LD + kq * 8 + 4, *reinterpret_cast<float(*)[4]>(&kv[4]));
#pragma unroll
            for (int q = 0; q < 8; q++) acc[q] = fma_(pj, kv[q], acc[q]);
        }
        st4(sA + r * LD + kq * 8, *reinterpret_cast<float(*)[4]>(&acc[0]));
        st4(sA + r * LD + kq * 8 + 4, *reinterpret_cast<float(*)[4]>(&acc[4]));
    }
    __syncthreads();
    float acc[RPT];
#pragma unroll
    for (int i = 0; i < RPT; i++) acc[i] = 0.0f;
#pragma unroll
    for (int k = 0; k < E; k += 4) {                            // (4) att = a . Watt^T
#pragma unroll
        for (int i = 0; i < RPT; i++) {
            float av[4];
            ld4(sA + (rg + GS * i) * LD + k, av);
            acc[i] = fma_(av[0], w[k], acc[i]);
            acc[i] = fma_(av[1], w[k + 1], acc[i]);
            acc[i] = fma_(av[2], w[k + 2], acc[i]);
            acc[i] = fma_(av[3], w[k + 3], acc[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < RPT; i++) { sT[(rg + GS * i) * LD + o] = acc[i]; acc[i] = 0.0f; }
#pragma unroll
    for (int k = 0; k < E; k++) w[k] = __ldg(w1T + k * E + o);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < E; k += 4) {                            // (5) h = relu([x | att] . W1^T + b1): item half
#pragma unroll
        for (int i = 0; i < RPT; i++) {
            float xv[4];
            ld4(sXs + (rg + GS * i) * LD + k, xv);
            acc[i] = fma_(xv[0], w[k], acc[i]);
            acc[i] = fma_(xv[1], w[k + 1], acc[i]);
            acc[i] = fma_(xv[2], w[k + 2], acc[i]);
            acc[i] = fma_(xv[3], w[k + 3], acc[i]);
        }
    }
#pragma unroll
    for (int k = 0; k < E; k++) w[k] = __ldg(w1T + (E + k) * E + o);
#pragma unroll
    for (int k = 0; k < E; k += 4) {                            //     attention half
#pragma unroll
        for (int i = 0; i < RPT; i++) {
            float tv[4];
            ld4(sT + (rg + GS * i) * LD + k, tv);
            acc[i] = fma_(tv[0], w[k], acc[i]);
            acc[i] = fma_(tv[1], w[k + 1], acc[i]);
            acc[i] = fma_(tv[2], w[k + 2], acc[i]);
            acc[i] = fma_(tv[3], w[k + 3], acc[i]);
        }
    }
    {
        const float bo = __ldg(b1 + o);
#pragma unroll
        for (int i = 0; i < RPT; i++) sA[(rg + GS * i) * LD + o] = relu_(add_(acc[i], bo));   // a[] was last read before the previous barrier
    }
    __syncthreads();
    if (tid < nr) {                                             // (6) logit = h . W2 + b2
        const float *hr = sA + tid * LD;
        float l = 0.0f;
#pragma unroll 8
        for (int q = 0; q < E; q++) l = fma_(hr[q], __ldg(w2 + q), l);
        sOut[base + tid] = add_(l, b2);
    }
    __syncthreads();
}

__device__ __noinline__ void strict_score_batch(const float *__restrict__ emb, const int32_t *__restrict__ sRow,
                                                int n, float *__restrict__ sOut,
                                                float *__restrict__ scr, uint32_t maskbits, int T, float scale,
                                                const float *__restrict__ wattT, const float *__restrict__ w1T,
                                                const float *__restrict__ b1, const float *__restrict__ w2, float b2)
{
    int base = 0;
    while (base < n) {
        const int rem = n - base;
        if (rem > 16) {
            strict_score_rows<8>(emb, sRow, base, rem < 32 ? rem : 32, sOut, scr, maskbits, T, scale, wattT, w1T, b1, w2, b2);
            base += 32;
        } else if (rem > 8) {
            strict_score_rows<4>(emb, sRow, base, rem, sOut, scr, maskbits, T, scale, wattT, w1T, b1, w2, b2);
            base += 16;
        } else if (rem > 4) {
            strict_score_rows<2>(emb, sRow, base, rem, sOut, scr, maskbits, T, scale, wattT, w1T, b1, w2, b2);
            base += 8;
        } else {
            strict_score_rows<1>(emb, sRow, base, rem, sOut, scr, maskbits, T, scale, wattT, w1T, b1, w2, b2);
            base += 4;
        }
    }
}

// the same for a CTA of 128 threads (wave_select_kernel): 16 / 8 / 4 / 2 rows per call
__device__ __noinline__ void strict_score_batch128(const float *__restrict__ emb, const int32_t *__restrict__ sRow,
                                                   int n, float *__restrict__ sOut,
                                                   float *__restrict__ scr, uint32_t maskbits, int T, float scale,
                                                   const float *__restrict__ wattT, const float *__restrict__ w1T,
                                                   const float *__restrict__ b1, const float *__restrict__ w2, float b2)
{
    int base = 0;
    while (base < n) {
        const int rem = n - base;
        if (rem > 8) {
            strict_score_rows<8, 128>(emb, sRow, base, rem < 16 ? rem : 16, sOut, scr, maskbits, T, scale, wattT, w1T, b1, w2, b2);
            base += 16;
        } else if (rem > 4) {
            strict_score_rows<4, 128>(emb, sRow, base, rem, sOut, scr, maskbits, T, scale, wattT, w1T, b1, w2, b2);
            base += 8;
        } else if (rem > 2) {
            strict_score_rows<2, 128>(emb, sRow, base, rem, sOut, scr, maskbits, T, scale, wattT, w1T, b1, w2, b2);
            base += 4;
        } else {
            strict_score_rows<1, 128>(emb, sRow, base, rem, sOut, scr, maskbits, T, scale, wattT, w1T, b1, w2, b2);
            base += 2;
        }
    }
}

__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- beam cut: block-wide selection of the k best fast scores -------------------------------------------------
// The CTA holds <= 512 order keys two per thread (k0, k1; 0 marks an invalid slot).  block_select finds a threshold t
// with EXACTLY k keys >= t and returns kdn = smallest key >= t (the k-th largest) and kup = largest key < t (the
// (k+1)-th): a row of the upper group is certainly inside the strict top k iff fast > float(kup) + 2 eps, a row of
// the lower group certainly outside iff fast < float(kdn) - 2 eps -- the pairwise condition against the extreme of
// the other group, tighter than a band around a pivot.  With ties at the cut (no such t) it returns an interval
// [kdn, kup] narrower than 2^8 key steps (2^-15 relative) that contains the k-th largest key; the same two
// comparisons then give a (slightly wider) valid band.
// Search: thresholds alternate between linear interpolation of the bracket counts and the bracket midpoint, both in
// SCORE space (key space is logarithmic: its midpoints crawl through the exponents); per step two ballots per warp,
// one 8-entry shared-memory exchange and one barrier (double-buffered); ~8 steps for 400 keys.  After 12 steps, or
// with non-finite scores, plain key-space bisection guarantees termination.  All threads call it with the same
// lo/hi/clo and return the same.
// Measured alternatives (B200, cycles per cut): one warp with 16 keys per lane ~7 k (a lone warp issues ~0.2
// instr/clk on dependent code); shared-memory histograms, radix or linear buckets, as long (ATOMS retires ~2 cycles
// per lane); 7 thresholds per step: 3.6 steps but 1.8 k cycles each (register pressure at the 128-register cap).
__device__ __forceinline__ void block_minmax(uint32_t k0, uint32_t k1, int *sRed, uint32_t &lo, uint32_t &hi, int &nvalid)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t mn = min(k0 ? k0 : 0xffffffffu, k1 ? k1 : 0xffffffffu), mx = max(k0, k1);
    mn = __reduce_min_sync(0xffffffffu, mn);
    mx = __reduce_max_sync(0xffffffffu, mx);
    const int nv = __popc(__ballot_sync(0xffffffffu, k0 != 0u)) + __popc(__ballot_sync(0xffffffffu, k1 != 0u));
    if (lane == 0) { sRed[warp] = (int)mn; sRed[8 + warp] = (int)mx; sRed[16 + warp] = nv; }
    __syncthreads();
    const uint4 a = *reinterpret_cast<const uint4 *>(sRed), b = *reinterpret_cast<const uint4 *>(sRed + 4);
    const uint4 c = *reinterpret_cast<const uint4 *>(sRed + 8), d = *reinterpret_cast<const uint4 *>(sRed + 12);
    const int4 e = *reinterpret_cast<const int4 *>(sRed + 16), f = *reinterpret_cast<const int4 *>(sRed + 20);
    lo = min(min(min(a.x, a.y), min(a.z, a.w)), min(min(b.x, b.y), min(b.z, b.w)));
    hi = max(max(max(c.x, c.y), max(c.z, c.w)), max(max(d.x, d.y), max(d.z, d.w)));
    nvalid = (e.x + e.y) + (e.z + e.w) + (f.x + f.y) + (f.z + f.w);
    __syncthreads();                                           // sRed is reused by the search
}

// lo <= every valid key <= hi, clo = number of valid keys >= k.  sRed: 128 ints (two 64-int count buffers).
template <typename StepFn>
__device__ __forceinline__ void block_select(uint32_t k0, uint32_t k1, int k, uint32_t lo, uint32_t hi, int clo, int *sRed,
                                             uint32_t &kdn, uint32_t &kup, StepFn on_step)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int chi = 0;                                               // clo = count(keys >= lo) >= k, chi = count(keys >= hi + 1) < k
    int *sCnt = sRed;
    int it = 0;
#pragma unroll 1
    while (clo != k && hi - lo > 255u && hi > lo) {
        uint32_t mid = lo + ((hi - lo) >> 1) + 1u;
        if (it < 12) {
            const float flo = key_to_float(lo), fhi = key_to_float(hi);
            const float fr = (it & 1) ? 0.5f : ((float)(k - chi) - 0.5f) / (float)(clo - chi);
            const float fm = fhi - (fhi - flo) * fr;
            if (fm == fm) mid = order_key(fm);
        }
        mid = min(max(mid, lo + 1u), hi);
        const int c = __syncthreads_count(k0 >= mid) + __syncthreads_count(k1 >= mid);     // barrier.red.popc: no shared-memory round trip
        it++;
        if (c >= k) { lo = mid; clo = c; } else { hi = mid - 1u; chi = c; }
    }
    on_step(it);
    if (clo == k) {
        // k keys >= lo: the k-th largest is the smallest of them, the (k+1)-th the largest key below lo
        uint32_t a = min(k0 >= lo ? k0 : 0xffffffffu, k1 >= lo ? k1 : 0xffffffffu);
        uint32_t b = max(k0 < lo ? k0 : 0u, k1 < lo ? k1 : 0u);
        a = __reduce_min_sync(0xffffffffu, a);
        b = __reduce_max_sync(0xffffffffu, b);
        if (lane == 0) { sCnt[warp] = (int)a; sCnt[8 + warp] = (int)b; }
        __syncthreads();
        const uint4 p0 = *reinterpret_cast<const uint4 *>(sCnt), p1 = *reinterpret_cast<const uint4 *>(sCnt + 4);
        const uint4 q0 = *reinterpret_cast<const uint4 *>(sCnt + 8), q1 = *reinterpret_cast<const uint4 *>(sCnt + 12);
        kdn = min(min(min(p0.x, p0.y), min(p0.z, p0.w)), min(min(p1.x, p1.y), min(p1.z, p1.w)));
        kup = max(max(max(q0.x, q0.y), max(q0.z, q0.w)), max(max(q1.x, q1.y), max(q1.z, q1.w)));
        if (kup == 0u) kup = kdn;                              // no valid key below the cut (k == number of valid keys)
    } else {
        kdn = lo; kup = hi;
    }
}

// ---- (A) gather NIT x RS rows of a tile: 16 lanes x 16 B per row, every load issued before the first use --
// codes: candidate codes of the tile (shared), row = q*RS + rowbase; rows past nrows re-read the last valid row
// (their operand rows are never consumed).  dst: operand base + this thread's chunk / row offset.
template <int NIT, int RS>
__device__ __forceinline__ void gather_convert(const float *__restrict__ emb_chunk, const int32_t *__restrict__ codes, int nrows,
                                               int rowbase, unsigned char *__restrict__ dstH, unsigned char *__restrict__ dstL)
{
    float4 v[NIT];
#pragma unroll
    for (int q = 0; q < NIT; q++) {
        int row = q * RS + rowbase;
        row = row < nrows ? row : nrows - 1;
        v[q] = ldg_row16(emb_chunk + (size_t)codes[row] * 64);
    }
#pragma unroll
    for (int q = 0; q < NIT; q++) {
        uint2 hi, lo;
        split_pair(v[q].x, v[q].y, hi.x, lo.x);
        split_pair(v[q].z, v[q].w, hi.y, lo.y);
        *reinterpret_cast<uint2 *>(dstH + q * RS * 16) = hi;
        *reinterpret_cast<uint2 *>(dstL + q * RS * 16) = lo;
    }
}
__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Phase timers (cycles of thread 0, summed over CTAs) -> stats[8 + i]; compiled in with -DDMG_FAST_TIMING.
#ifdef DMG_FAST_TIMING
#define DMG_TICK(i) do { if (tid == 0) { const long long t_ = clock64(); tacc[i] += t_ - tlast; tlast = t_; } } while (0)
#else
#define DMG_TICK(i) do { } while (0)
#endif
enum { TK_PROLOGUE = 0, TK_SELECT, TK_RESCORE, TK_EXPAND, TK_GATHER, TK_SOFTMAX, TK_EPILOGUE, TK_FINAL, TK_SCHED, TK_SELWARP, TK_MMAWAIT, TK_PHWAIT, TK_FINPREP, TK_FINSTRICT, TK_SELA, TK_SELB, TK_N };

__global__ void __launch_bounds__(FastGeo::THREADS, 2) beam_search_fast_kernel(const BeamParams<float> p, const FastParams fp)
{
    using G = FastGeo;
    constexpr int E = G::E;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *sp = smem_raw;
    unsigned char *sXh = sp; sp += G::X_BYTES;
    unsigned char *sXl = sp; sp += G::X_BYTES;
    unsigned char *sPh = sp; sp += G::P_BYTES;
    unsigned char *sPl = sp; sp += G::P_BYTES;
    unsigned char *sBh = sp; sp += G::B_BYTES;
    unsigned char *sBl = sp; sp += G::B_BYTES;
    unsigned char *sHh = sp; sp += G::H_BYTES;
    unsigned char *sHl = sp; sp += G::H_BYTES;
    float *sScore = reinterpret_cast<float *>(sp); sp += (size_t)p.cap * 4;
    int32_t *sCode0 = reinterpret_cast<int32_t *>(sp); sp += (size_t)p.cap * 4;
    int32_t *sCode1 = reinterpret_cast<int32_t *>(sp); sp += (size_t)p.cap * 4;
    int32_t *sVCode = reinterpret_cast<int32_t *>(sp); sp += G::VCAP * 4;   // deferred verification list
    float *sVFast = reinterpret_cast<float *>(sp); sp += G::VCAP * 4;
    uint32_t *sVMeta = reinterpret_cast<uint32_t *>(sp); sp += G::VCAP * 4;  // start | n << 8 | need << 16 | chosen << 24 | segment << 25
    float *sSegEps = reinterpret_cast<float *>(sp); sp += 32 * 4;
    int32_t *sMisc = reinterpret_cast<int32_t *>(sp); sp += 128 * 4;   // [0..15] hist codes, [16] mask bits, [32..39] scan, [40..] scalars
    uint64_t *sBar = reinterpret_cast<uint64_t *>(sp);                 // [0] history TMA, [1] S MMAs, [2] all MMAs of a pass
    // aliases (valid only while no MMA is in flight)
    float *sKf = reinterpret_cast<float *>(sXh);                       // history fp32 [16][KLD]
    int *sRed = reinterpret_cast<int *>(sPh + G::OFF_HIST);           // 256 ints: block_select exchange
    int *sSel = reinterpret_cast<int *>(sPh + G::OFF_SEL);
    uint32_t *sKeyU = reinterpret_cast<uint32_t *>(sPh + G::OFF_KEYU);
    int *sUPos = reinterpret_cast<int *>(sPh + G::OFF_UPOS);
    int32_t *sLCode = reinterpret_cast<int32_t *>(sPh + G::OFF_LCODE);  // codes handed to the strict scorer
    float *sLStr = reinterpret_cast<float *>(sPh + G::OFF_LSTR);        // ... and their strict logits
    uint8_t *sCls = reinterpret_cast<uint8_t *>(sPh + G::OFF_CLS);
    float *sHmax = reinterpret_cast<float *>(sPh);                     // prologue only: [4][64] max_j |H_jo| partials

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = p.T;

    // ---- one-time setup: W1x -> bf16 hi/lo B operand, barriers, TMEM ------------------------------
    for (int i = tid; i < E * (E / 8); i += G::THREADS) {         // 64 outputs x 8 k-chunks
        const int o = i % E, kc = i / E;
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; q++) v[q] = __ldg(fp.w1 + o * 2 * E + kc * 8 + q);
        uint4 hi, lo;
        split8(v, hi, lo);
        *reinterpret_cast<uint4 *>(sBh + kc * G::B_LBO + o * 16) = hi;
        *reinterpret_cast<uint4 *>(sBl + kc * G::B_LBO + o * 16) = lo;
    }
    if (tid == 0) { for (int i = 0; i < 5; i++) mbar_init(&sBar[i], 1); }      // [0] history TMA; tile t: [1 + 2t] X.[W1x|K]^T done, [2 + 2t] P.H done
    if (warp == 0) tmem_alloc(reinterpret_cast<uint32_t *>(&sMisc[41]), 256);   // tile t: Hacc [128t, 128t+64), S [128t+64, 128t+80)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(&sMisc[41]);
    const uint32_t tmem_lane = (uint32_t)((warp & 3) * 32) << 16;
    const int tile_of_warp = warp >> 2;
    const float scale2 = p.scale * 1.4426950408889634f, inv_T = 1.0f / (float)T;
    float *sAddv = reinterpret_cast<float *>(sMisc + 64);               // [16] additive softmax mask of the current user
    uint32_t hist_phase = 0, s_phase = 0;                        // s_phase: phase of this warpgroup's two MMA mbarriers
    unsigned long long st_cuts = 0, st_recuts = 0, st_rerows = 0, st_rows = 0, st_redo = 0, st_sync = 0, st_iters = 0;
    float st_ratio = 0.0f;
#ifdef DMG_FAST_TIMING
    long long tacc[TK_N] = {0}, tlast = clock64();
#endif

    // operand descriptors = one live register (the CTA's shared-memory base) + compile-time offsets;
    // a k-step / M-tile offset is added to the 16-byte start-address field
    const uint32_t sbase16 = smem_u32(smem_raw) >> 4;
    auto mkdesc = [&](uint32_t byte_off, uint32_t lbo) -> uint64_t {
        return ((uint64_t)(0x4000u | (G::SBO >> 4)) << 32) | (uint64_t)((sbase16 + (byte_off >> 4)) | ((lbo >> 4) << 16));
    };
    constexpr uint32_t OFF_XH = 0, OFF_XL = G::X_BYTES, OFF_PH = 2 * G::X_BYTES, OFF_PL = OFF_PH + G::P_BYTES,
                       OFF_BH = OFF_PL + G::P_BYTES, OFF_BL = OFF_BH + G::B_BYTES, OFF_HH = OFF_BL + G::B_BYTES,
                       OFF_HL = OFF_HH + G::H_BYTES;

    // history rows (fp32) -> sKf through the TMA bulk-copy engine; padding rows are zero
    auto load_history = [&]() {
        fence_proxy_async();
        if (tid == 0) {
            uint32_t bytes = 0;
            for (int j = 0; j < T; j++) if (sMisc[j] >= 0) bytes += E * sizeof(float);
            mbar_expect_tx(&sBar[0], bytes);
            for (int j = 0; j < T; j++)
                if (sMisc[j] >= 0) tma_bulk_g2s(sKf + j * G::KLD, p.emb + (size_t)sMisc[j] * E, E * sizeof(float), &sBar[0]);
        }
        for (int i = tid; i < kMaxT * E; i += G::THREADS) {
            const int j = i / E;
            if (j >= T || sMisc[j] < 0) sKf[j * G::KLD + (i % E)] = 0.0f;
        }
        mbar_wait(&sBar[0], hist_phase);
        hist_phase ^= 1;
        __syncthreads();
    };

    // strict logits of the n rows whose codes sit in sLCode[0..n) -> sLStr[0..n)
    auto strict_rescore = [&](int n) {
        __syncthreads();
        load_history();
        strict_score_batch(p.emb, sLCode, n, sLStr, sKf, (uint32_t)sMisc[16], T, p.scale, p.wattT, p.w1T, p.b1, p.w2, fp.b2);
    };
    auto track_ratio = [&](float strict, float fast, float eps_now) {
        if (eps_now > 0.0f && eps_now < 1e30f) {
            const float ratio = fabsf(strict - fast) / eps_now;
            if (ratio > st_ratio) st_ratio = ratio;
        }
    };

    if (tid == 0) sMisc[43] = 0;                                // tail owner flag (thread 0 only)
    for (;;) {
        // ---- next user (dynamic scheduler) ------------------------------------------------------
        __syncthreads();
        if (tid == 0) {
            // The last round of a batch holds fewer users than SMs: the first CTA of an SM to run out of main-round users
            // becomes its tail owner and runs tail users alone (a chain is ~20 % faster without a co-resident CTA), the
            // other CTA of the SM retires.
            int tail_owner = sMisc[43];
            int u = tail_owner ? p.B : atomicAdd(fp.work_counter, 1);
            if (u >= fp.tail_start) {
                u = p.B;
                if (fp.tail_start < p.B) {
                    if (!tail_owner) {
                        uint32_t smid;
                        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                        sMisc[43] = tail_owner = atomicExch(fp.work_counter + 8 + (smid & 255), 1) == 0;
                    }
                    if (tail_owner) u = fp.tail_start + atomicAdd(fp.work_counter + 2, 1);
                }
            }
            sMisc[42] = u;
        }
        __syncthreads();
        const int user = sMisc[42];
        if (user >= p.B) break;
        DMG_TICK(TK_SCHED);

        // ---- K2: history tile, its operand forms, H = K . M^T -------------------------------------
        if (tid < kMaxT) {
            int c = -1, m = 0;
            if (tid < T) { c = p.hist[(size_t)user * T + tid]; m = p.hist_mask[(size_t)user * T + tid]; }
            sMisc[tid] = c;
            const uint32_t mb = __ballot_sync(0x0000ffffu, m != 0);
            if (tid == 0) { sMisc[16] = (int)mb; sMisc[44] = 0; sMisc[45] = 0; }
            sAddv[tid] = (tid < T && !m) ? 0.0f : -3.4028234663852886e+38f;
        }
        __syncthreads();
        load_history();
        if (tid < 128) {                                        // history as B operand of S = X . K^T : [16 j][64 k]
            const int j = tid & 15, kc = tid >> 4;
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; q++) v[q] = sKf[j * G::KLD + kc * 8 + q];
            uint4 hi, lo;
            split8(v, hi, lo);
            *reinterpret_cast<uint4 *>(sBh + kc * G::B_LBO + (64 + j) * 16) = hi;
            *reinterpret_cast<uint4 *>(sBl + kc * G::B_LBO + (64 + j) * 16) = lo;
        } else if (tid < 192) {                                 // ZK = sum_k z_k max_j |K_jk|  (two warp partials)
            const int k = tid - 128;
            float kab = 0.0f;
#pragma unroll
            for (int j = 0; j < kMaxT; j++) { const float v = fabsf(sKf[j * G::KLD + k]); kab = (v > kab || v != v) ? v : kab; }   // NaN sticks
            float zk = __ldg(fp.zvec + k) * kab;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) zk += __shfl_xor_sync(0xffffffffu, zk, o);
            if (lane == 0) sMisc[46 + (warp & 1)] = __float_as_int(zk);
        } else if (tid < 192 + kMaxT) {                         // Kmax = max_j |K_j|_2
            const int j = tid - 192;
            float n2 = 0.0f;
            for (int k = 0; k < E; k++) n2 = fmaf(sKf[j * G::KLD + k], sKf[j * G::KLD + k], n2);
            float nj = sqrtf(n2) * 1.0001f;
            if (!(nj == nj)) nj = __int_as_float(0x7f800000);
            atomicMax(&sMisc[44], __float_as_int(nj));
        }
        {                                                       // H[j][o] = sum_k M[o][k] K[j][k], 4 j per thread
            const int o = tid & 63, jg = tid >> 6;
            float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            const float *kr = sKf + (jg * 4) * G::KLD;
#pragma unroll 8
            for (int k = 0; k < E; k++) {
                const float m = __ldg(fp.mT + k * E + o);
#pragma unroll
                for (int q = 0; q < 4; q++) acc[q] = fmaf(m, kr[q * G::KLD + k], acc[q]);
            }
            float hm = fmaxf(fmaxf(fabsf(acc[0]), fabsf(acc[1])), fmaxf(fabsf(acc[2]), fabsf(acc[3])));
            if (acc[0] != acc[0] || acc[1] != acc[1] || acc[2] != acc[2] || acc[3] != acc[3]) hm = __int_as_float(0x7f800000);
            sHmax[jg * E + o] = hm;
            if (jg == 3) acc[3] = fp.b1[o];                     // T <= 15: row 15 of H carries b1 (P column 15 is 1)
            uint2 hi, lo;
            split_pair(acc[0], acc[1], hi.x, lo.x);
            split_pair(acc[2], acc[3], hi.y, lo.y);
            const int off = (jg >> 1) * G::H_LBO + o * 16 + (jg & 1) * 8;
            *reinterpret_cast<uint2 *>(sHh + off) = hi;
            *reinterpret_cast<uint2 *>(sHl + off) = lo;
        }
        __syncthreads();
        if (tid < E) {                                          // HW = sum_o |w2_o| max_j |H_jo|  (two warp partials)
            float hw = fmaxf(fmaxf(sHmax[tid], sHmax[E + tid]), fmaxf(sHmax[2 * E + tid], sHmax[3 * E + tid])) * fabsf(fp.w2[tid]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) hw += __shfl_xor_sync(0xffffffffu, hw, o);
            if (lane == 0) sMisc[48 + warp] = __float_as_int(hw);
        }
        __syncthreads();
        const float kmax = __int_as_float(sMisc[44]);
        const float zk_user = (__int_as_float(sMisc[46]) + __int_as_float(sMisc[47])) * 1.0001f;
        const float hw_user = (__int_as_float(sMisc[48]) + __int_as_float(sMisc[49])) * 1.0001f;
        const bool all_masked = ((uint32_t)sMisc[16] & ((1u << T) - 1u)) == ((1u << T) - 1u);

        const int beam = p.beam_user ? p.beam_user[user] : p.beam;
        const int s_level = 31 - __clz(beam);
        int32_t *cur = sCode0, *nxt = sCode1;
        int count = 0;
        if (s_level <= p.leaf_level) {
            const int64_t start = ((int64_t)1 << s_level) - 1;
            const int n0 = 1 << s_level;
            for (int base = 0; base < n0; base += G::THREADS) {
                const int i = base + tid;
                const int e = (i < n0 && code_exists(p.exists, start + i)) ? 1 : 0;
                int tot;
                const int o = block_exscan(e, sMisc + 32, &tot);
                if (e) cur[count + o] = (int32_t)(start + i);
                count += tot;
            }
            __syncthreads();
        }
        DMG_TICK(TK_PROLOGUE);
        float eps_level = 0.0f;                                 // bound on |fast - strict| of the scores in sScore
        int vcount = 0, nseg = 0;                               // deferred verification list of this user
        bool redo = false;
        int redo_why = 0;                                       // 0 wide band, 1 tie in place, 2 unscored, 3 wide final, 4 tie at the end, 5 proof failed
        bool scored = false;

        for (int level = s_level; level < p.leaf_level && count > 0; level++) {
            // ---- which candidates stay (certified cut), then their children, order preserved --------
            // thread t owns candidates 2t and 2t+1
            const int i0 = 2 * tid, i1 = 2 * tid + 1;
            const bool cut = count > beam;
            int cls0 = i0 < count ? 1 : 0, cls1 = i1 < count ? 1 : 0;      // 0 out, 1 in, 2 uncertain
            if (cut) {
                st_cuts++;
                const float f0 = i0 < count ? sScore[i0] : 0.0f, f1 = i1 < count ? sScore[i1] : 0.0f;
                if (tid == 0) { sSel[4] = 0; sSel[5] = 0; }
                uint32_t kdn, kup;
                {
                    const uint32_t key0 = i0 < count ? order_key(f0) : 0u, key1 = i1 < count ? order_key(f1) : 0u;
                    const uint32_t lo = (uint32_t)sMisc[50], hi = (uint32_t)sMisc[51];     // from the epilogue of this level's passes
                    DMG_TICK(TK_SELA);
                    block_select(key0, key1, beam, lo, hi, count, sRed, kdn, kup, [&](int n) { st_iters += n; });
                    DMG_TICK(TK_SELB);
                }
                const float band = 2.0f * eps_level * 1.0001f + 1e-30f;
                const float up = key_to_float(kup) + band, dn = key_to_float(kdn) - band;
                if (i0 < count) cls0 = f0 > up ? 1 : (f0 < dn ? 0 : 2);
                if (i1 < count) cls1 = f1 > up ? 1 : (f1 < dn ? 0 : 2);
                int slot0 = -1, slot1 = -1;
                if (cls0 == 2) { slot0 = atomicAdd(&sSel[5], 1); if (slot0 < G::MAX_UNC) { sUPos[slot0] = i0; sKeyU[slot0] = order_key(f0); } }
                if (cls1 == 2) { slot1 = atomicAdd(&sSel[5], 1); if (slot1 < G::MAX_UNC) { sUPos[slot1] = i1; sKeyU[slot1] = order_key(f1); } }
                {
                    const int mine = ((cls0 != 0) + (cls1 != 0)) | (((cls0 == 2) + (cls1 == 2)) << 16);
                    const int wsum = __reduce_add_sync(0xffffffffu, mine);
                    if (lane == 0 && wsum) atomicAdd(&sSel[4], wsum);
                }
                __syncthreads();
                const int n_keep = sSel[4] & 0xffff, n_unc = sSel[4] >> 16;
                DMG_TICK(TK_SELECT);
                if (n_keep != beam) {                            // some uncertain row must go
                    st_recuts++;
                    if (n_unc > G::MAX_UNC) { redo = true; redo_why = 0; break; }
                    const int need = beam - (n_keep - n_unc);    // uncertain rows that still fit (1 <= need < n_unc)
                    // rank of every band row by its FAST score (owner thread), and the fast-score gap at the cut
                    int frank0 = 0, frank1 = 0;
                    if (cls0 == 2 || cls1 == 2) {
                        const uint32_t k0 = order_key(f0), k1 = order_key(f1);
                        for (int q = 0; q < n_unc; q++) {
                            const uint32_t kq = sKeyU[q];
                            const int pq = sUPos[q];
                            frank0 += (kq > k0 || (kq == k0 && pq < i0)) ? 1 : 0;
                            frank1 += (kq > k1 || (kq == k1 && pq < i1)) ? 1 : 0;
                        }
                        if (cls0 == 2 && frank0 == need - 1) sSel[2] = __float_as_int(f0);
                        if (cls0 == 2 && frank0 == need) sSel[3] = __float_as_int(f0);
                        if (cls1 == 2 && frank1 == need - 1) sSel[2] = __float_as_int(f1);
                        if (cls1 == 2 && frank1 == need) sSel[3] = __float_as_int(f1);
                    }
                    __syncthreads();
                    const float gap = __int_as_float(sSel[2]) - __int_as_float(sSel[3]);
                    // Observed |fast - strict| stays below 1 % of eps: with a gap of eps/32 or more the fast order is
                    // taken now and PROVEN at the end of the search (one strict batch over every deferred band row);
                    // a narrower gap is settled strictly right here.
                    const bool defer = gap >= 0.03125f * eps_level && vcount + n_unc <= G::VCAP && nseg < 32 && eps_level < 1e30f;
                    if (defer) {
                        if (cls0 == 2) {
                            cls0 = frank0 < need ? 1 : 0;
                            const int e = vcount + slot0;
                            sVCode[e] = cur[i0]; sVFast[e] = f0;
                            sVMeta[e] = (uint32_t)vcount | ((uint32_t)n_unc << 8) | ((uint32_t)need << 16) | ((uint32_t)cls0 << 24) | ((uint32_t)nseg << 25);
                        }
                        if (cls1 == 2) {
                            cls1 = frank1 < need ? 1 : 0;
                            const int e = vcount + slot1;
                            sVCode[e] = cur[i1]; sVFast[e] = f1;
                            sVMeta[e] = (uint32_t)vcount | ((uint32_t)n_unc << 8) | ((uint32_t)need << 16) | ((uint32_t)cls1 << 24) | ((uint32_t)nseg << 25);
                        }
                        if (tid == 0) sSegEps[nseg] = eps_level;
                        vcount += n_unc;
                        nseg++;
                    } else {
                        st_sync++;
                        st_rerows += n_unc;
                        if (cls0 == 2) sLCode[slot0] = cur[i0];
                        if (cls1 == 2) sLCode[slot1] = cur[i1];
                        strict_rescore(n_unc);
                        int tie = 0;
                        if (cls0 == 2 || cls1 == 2) {
                            const uint32_t k0 = cls0 == 2 ? order_key(sLStr[slot0]) : 0u, k1 = cls1 == 2 ? order_key(sLStr[slot1]) : 0u;
                            if (cls0 == 2) track_ratio(sLStr[slot0], f0, eps_level);
                            if (cls1 == 2) track_ratio(sLStr[slot1], f1, eps_level);
                            // strict rank of a band row = rows strictly above it (+ an undetermined share of its exact ties: the
                            // reference orders ties by candidate position, which this path does not track).  Only a tie that
                            // STRADDLES the cut is undecidable here; ties on one side of it change nothing.
                            int g0 = 0, g1 = 0, e0 = 0, e1 = 0;
                            for (int q = 0; q < n_unc; q++) {
                                const uint32_t kq = order_key(sLStr[q]);
                                const int pq = sUPos[q];
                                g0 += kq > k0 ? 1 : 0;
                                g1 += kq > k1 ? 1 : 0;
                                e0 += (kq == k0 && pq != i0) ? 1 : 0;
                                e1 += (kq == k1 && pq != i1) ? 1 : 0;
                            }
                            if (cls0 == 2) { tie |= (g0 < need && g0 + e0 >= need) ? 1 : 0; cls0 = g0 + e0 < need ? 1 : 0; }
                            if (cls1 == 2) { tie |= (g1 < need && g1 + e1 >= need) ? 1 : 0; cls1 = g1 + e1 < need ? 1 : 0; }
                        }
                        if (__syncthreads_or(tie)) { redo = true; redo_why = 1; break; }
                    }
                    DMG_TICK(TK_RESCORE);
                }
            }
            int nc;
            {
                const int64_t c0 = cls0 ? cur[i0] : 0, c1 = cls1 ? cur[i1] : 0;
                const uint32_t *bm = (level + 1 >= fp.sparse_from) ? p.exists : nullptr;      // full levels: every child exists
                const int a1 = (cls0 && code_exists(bm, 2 * c0 + 1)) ? 1 : 0, a2 = (cls0 && code_exists(bm, 2 * c0 + 2)) ? 1 : 0;
                const int b1 = (cls1 && code_exists(bm, 2 * c1 + 1)) ? 1 : 0, b2 = (cls1 && code_exists(bm, 2 * c1 + 2)) ? 1 : 0;
                int o = block_exscan(a1 + a2 + b1 + b2, sMisc + 32, &nc);
                if (a1) nxt[o++] = (int32_t)(2 * c0 + 1);
                if (a2) nxt[o++] = (int32_t)(2 * c0 + 2);
                if (b1) nxt[o++] = (int32_t)(2 * c1 + 1);
                if (b2) nxt[o++] = (int32_t)(2 * c1 + 2);
                if (tid == 0) { sMisc[50] = -1; sMisc[51] = 0; }  // min / max order key of the scores of the level about to be scored (epilogue atomics)
            }
            __syncthreads();
            { int32_t *t = cur; cur = nxt; nxt = t; }
            count = nc;
            st_rows += count;
            scored = true;
            DMG_TICK(TK_EXPAND);

            // ---- eps of this level's scores (children sit on tree level `level + 1`) ------------------
            {
                const float u = 5.9604645e-8f;                   // 2^-24
                const float vx = __ldg(fp.lvl_vx + level + 1), nx = __ldg(fp.lvl_nx + level + 1);
                const float smax = p.scale * nx * kmax;
                const float ds = 6.1035156e-5f * smax;           // 2^-14: |s_fast - s_strict| <= ds
                const float dp1 = 2.1f * ds + 2.0f * (float)(T + 8) * u + 2.0f * 9.5367432e-7f * (1.0f + 2.0f * smax);
                float eps = fp.cA * vx + fp.cZ * zk_user + (fp.cH + dp1) * hw_user * 1.05f + fp.cGamma;
                if (!(ds < 0.04f) || !(eps < 1e30f)) eps = __int_as_float(0x7f800000);
                eps_level = eps * fp.tau;
            }

            // ---- fast scoring, 256 rows per pass ---------------------------------------------------------
            // Two independent scoring chains per CTA: warpgroup g (TMEM lanes of warp w are 32 (w & 3) ..) owns M tile g --
            // its own rows of X and P in shared memory, its own accumulators in TMEM, its own mbarriers and a 128-thread
            // named barrier -- and runs  gather -> convert -> MMA X.[W1x|K]^T -> softmax -> MMA P.H -> epilogue  without a
            // CTA-wide barrier; the groups meet again at the end of the level.  First up to 128 rows each, then the rest of
            // the level split in halves (400 candidates: 128 + 72 rows per group), so both chains carry the same work.
            {
                const int grp = warp >> 2, gw = warp & 3, gtid = tid & 127;
                const uint64_t xo = (uint64_t)(grp * 128);                    // this group's 128 rows x 16 B, in 16-byte descriptor units
                const uint32_t tm = tmem_base + grp * 128;
                for (int it = 0; it < 2; it++) {
                    int rbeg, nr;
                    if (it == 0) { rbeg = grp * 128; nr = count - rbeg < 128 ? count - rbeg : 128; }
                    else {
                        const int rem = count - 256;
                        if (rem <= 0) break;
                        const int h = (rem + 1) >> 1;
                        rbeg = 256 + grp * h;
                        nr = grp == 0 ? h : rem - h;
                    }
                    if (nr <= 0) continue;
                    // (A) gather rows -> bf16 hi/lo operand tile; every load is in flight before the first use
                    {
                        const int chunk = lane & 15, rowbase = gw * 2 + (lane >> 4);
                        const uint32_t st_off = (uint32_t)((chunk >> 1) * G::X_LBO + (chunk & 1) * 8 + (grp * 128 + rowbase) * 16);
                        if (nr > 80) gather_convert<16, 8>(p.emb + chunk * 4, cur + rbeg, nr, rowbase, sXh + st_off, sXl + st_off);
                        else if (nr > 64) gather_convert<10, 8>(p.emb + chunk * 4, cur + rbeg, nr, rowbase, sXh + st_off, sXl + st_off);
                        else gather_convert<8, 8>(p.emb + chunk * 4, cur + rbeg, nr, rowbase, sXh + st_off, sXl + st_off);
                    }
                    fence_proxy_async();
                    tc_fence_before();
                    if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
                    else asm volatile("bar.sync 2, 128;" ::: "memory");
                    DMG_TICK(TK_GATHER);
                    // (B) [Hacc | S] = X . [W1x | K]^T (N = 80: one pass over the A operand): 4 k-steps x (hi*hi + hi*lo + lo*hi).
                    // The group's first warp runs the descriptor arithmetic warp-uniformly; one elected lane issues.
                    if (gw == 0) {
                        tc_fence_after();
                        const bool leader = elect_one();
                        const uint64_t dXh = mkdesc(OFF_XH, G::X_LBO) + xo, dXl = mkdesc(OFF_XL, G::X_LBO) + xo;
                        const uint64_t dBh = mkdesc(OFF_BH, G::B_LBO), dBl = mkdesc(OFF_BL, G::B_LBO);
#pragma unroll
                        for (int ks = 0; ks < 4; ks++) {
                            const uint64_t ah = dXh + ks * (2 * G::X_LBO / 16), al = dXl + ks * (2 * G::X_LBO / 16);
                            const uint64_t bh = dBh + ks * (2 * G::B_LBO / 16), bl = dBl + ks * (2 * G::B_LBO / 16);
                            if (leader) {
                                umma_bf16(tm, ah, bh, kIdescBf16M128N80, ks > 0);
                                umma_bf16(tm, ah, bl, kIdescBf16M128N80, 1);
                                umma_bf16(tm, al, bh, kIdescBf16M128N80, 1);
                            }
                        }
                        if (leader) umma_commit(&sBar[1 + 2 * grp]);
                        __syncwarp();
                    }
                    // (C) Mask + SoftMax per row in registers (log2 domain: t_j = S_j * scale*log2e + addv_j, addv_j = -FLT_MAX on
                    // padded / masked slots), P -> bf16 hi/lo A operand [256][16]; column 15 = 1 multiplies the b1 row of H
                    if (gw * 32 < nr) {
                        mbar_wait(&sBar[1 + 2 * grp], s_phase);
                        tc_fence_after();
                        DMG_TICK(TK_MMAWAIT);
                        float sc[16];
                        tmem_ld16(tm + tmem_lane + 64, sc);
                        if (all_masked) {                         // SoftMax of T equal values: exactly 1/T each
#pragma unroll
                            for (int j = 0; j < 16; j++) sc[j] = j < T ? inv_T : 0.0f;
                        } else {
                            float mx = -3.4028234663852886e+38f;
#pragma unroll
                            for (int j = 0; j < 16; j++) {
                                sc[j] = fmaf(sc[j], scale2, sAddv[j]);
                                mx = fmaxf(mx, sc[j]);
                            }
                            float sum = 0.0f;
#pragma unroll
                            for (int j = 0; j < 16; j++) { sc[j] = ex2_approx(sc[j] - mx); sum += sc[j]; }
                            const float inv = 1.0f / sum;
#pragma unroll
                            for (int j = 0; j < 16; j++) sc[j] *= inv;
                        }
                        sc[15] = 1.0f;
                        uint4 hi, lo;
                        split8(*reinterpret_cast<float(*)[8]>(&sc[0]), hi, lo);
                        *reinterpret_cast<uint4 *>(sPh + tid * 16) = hi;
                        *reinterpret_cast<uint4 *>(sPl + tid * 16) = lo;
                        split8(*reinterpret_cast<float(*)[8]>(&sc[8]), hi, lo);
                        *reinterpret_cast<uint4 *>(sPh + G::P_LBO + tid * 16) = hi;
                        *reinterpret_cast<uint4 *>(sPl + G::P_LBO + tid * 16) = lo;
                    }
                    fence_proxy_async();
                    tc_fence_before();
                    if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");             // the group's P rows are complete
                    else asm volatile("bar.sync 2, 128;" ::: "memory");
                    DMG_TICK(TK_SOFTMAX);
                    // (D) Hacc += P . H of this group's tile (one k-step of 16; row 15 of H holds b1)
                    if (gw == 0) {
                        tc_fence_after();
                        const bool leader = elect_one();
                        const uint64_t dHh = mkdesc(OFF_HH, G::H_LBO), dHl = mkdesc(OFF_HL, G::H_LBO);
                        const uint64_t ah = mkdesc(OFF_PH, G::P_LBO) + xo, al = mkdesc(OFF_PL, G::P_LBO) + xo;
                        if (leader) {
                            umma_bf16(tm, ah, dHh, kIdescBf16M128N64, 1);
                            umma_bf16(tm, ah, dHl, kIdescBf16M128N64, 1);
                            umma_bf16(tm, al, dHh, kIdescBf16M128N64, 1);
                            umma_commit(&sBar[2 + 2 * grp]);
                        }
                        __syncwarp();
                    }
                    // (E) epilogue: logit = relu(Hacc) . W2 + b2 (b1 already inside Hacc), one row per thread
                    if (gw * 32 < nr) {
                        mbar_wait(&sBar[2 + 2 * grp], s_phase);
                        tc_fence_after();
                        DMG_TICK(TK_PHWAIT);
                        float logit = 0.0f;
#pragma unroll
                        for (int hf = 0; hf < 2; hf++) {
                            float hv[32];
                            tmem_ld32(tm + tmem_lane + hf * 32, hv);
#pragma unroll
                            for (int c = 0; c < 32; c++) logit = fmaf(fmaxf(hv[c], 0.0f), fp.w2[hf * 32 + c], logit);
                        }
                        const float sc_out = logit + fp.b2;
                        if (gtid < nr) sScore[rbeg + gtid] = sc_out;
                        {
                            const uint32_t ky = order_key(sc_out);
                            const uint32_t wmn = __reduce_min_sync(0xffffffffu, gtid < nr ? ky : 0xffffffffu);
                            const uint32_t wmx = __reduce_max_sync(0xffffffffu, gtid < nr ? ky : 0u);
                            if (lane == 0) { atomicMin(reinterpret_cast<unsigned int *>(&sMisc[50]), wmn); atomicMax(reinterpret_cast<unsigned int *>(&sMisc[51]), wmx); }
                        }
                    }
                    s_phase ^= 1;                                 // both mbarriers of the group completed one phase
                    tc_fence_before();                            // the next tile's MMAs overwrite the accumulators just read
                    DMG_TICK(TK_EPILOGUE);
                }
                tc_fence_before();
                __syncthreads();
            }
        }

        // ---- K3: topk, always settled with strict scores -----------------------------------------------
        if (!redo) {
            const int64_t leaf_start = ((int64_t)1 << p.leaf_level) - 1;
            const bool at_leaf = (s_level <= p.leaf_level);
            const int64_t c0 = p.cons_off ? p.cons_off[user] : 0, c1 = p.cons_off ? p.cons_off[user + 1] : 0;
            int valid = 0;
            for (int base = 0; base < count; base += G::THREADS) {
                const int i = base + tid;
                bool keep = false;
                if (i < count && at_leaf) {
                    const int64_t slot = (int64_t)cur[i] - leaf_start;
                    const int32_t item = (slot >= 0 && slot < ((int64_t)1 << p.leaf_level)) ? __ldg(p.leaf_item + slot) : -1;
                    keep = item >= 0;
                    for (int64_t q = c0; q < c1 && keep; q++) keep = (__ldg(p.cons + q) != item);
                }
                if (i < count) sKeyU[i] = keep ? order_key(sScore[i]) : 0u;
                valid += __syncthreads_count(keep);
            }
            const int kk = valid < p.topk ? valid : p.topk;
            if (kk > 0 && !scored) { redo = true; redo_why = 2; }                 // beam >= 2^leaf_level: nothing was scored, every score ties
            int na = 0;
            if (kk > 0 && !redo) {
                uint32_t kdn, kup;
                {
                    const uint32_t key0 = 2 * tid < count ? sKeyU[2 * tid] : 0u, key1 = 2 * tid + 1 < count ? sKeyU[2 * tid + 1] : 0u;
                    uint32_t lo, hi;
                    int nv;
                    block_minmax(key0, key1, sRed, lo, hi, nv);
                    block_select(key0, key1, kk, lo, hi, nv, sRed, kdn, kup, [](int) {});
                }
                const float dn = key_to_float(kdn) - (2.0f * eps_level * 1.0001f + 1e-30f);
                if (tid == 0) sMisc[45] = 0;
                __syncthreads();
                for (int i = tid; i < count; i += G::THREADS)
                    if (sKeyU[i] != 0u && !(sScore[i] < dn)) {
                        const int slot = atomicAdd(&sMisc[45], 1);
                        if (slot < G::MAX_FINAL) sUPos[slot] = i;
                    }
                __syncthreads();
                na = sMisc[45];
                if (na > G::MAX_FINAL) { redo = true; redo_why = 3; }
            }
            DMG_TICK(TK_FINPREP);
            if (!redo && vcount + na > 0) {
                // ONE strict batch: every band row deferred at a cut, then the topk candidates
                for (int e = tid; e < vcount; e += G::THREADS) sLCode[e] = sVCode[e];
                for (int q = tid; q < na; q += G::THREADS) sLCode[vcount + q] = cur[sUPos[q]];
                st_rerows += vcount + na;
                strict_rescore(vcount + na);
                DMG_TICK(TK_FINSTRICT);
                int bad = 0;
                for (int e = tid; e < vcount; e += G::THREADS) {   // the deferred cuts: strict order must pick the same rows
                    const uint32_t meta = sVMeta[e];
                    const int st = meta & 255u, n = (meta >> 8) & 255u, need = (meta >> 16) & 255u, chosen = (meta >> 24) & 1u;
                    track_ratio(sLStr[e], sVFast[e], sSegEps[meta >> 25]);
                    const uint32_t ks = order_key(sLStr[e]);
                    int rank = 0, ties = 0;
                    for (int q = st; q < st + n; q++) {
                        const uint32_t kq = order_key(sLStr[q]);
                        rank += kq > ks ? 1 : 0;
                        ties += (kq == ks && q != e) ? 1 : 0;
                    }
                    bad |= (rank < need && rank + ties >= need) ? 1 : 0;   // exact strict tie across the cut: the reference decides by position
                    bad |= ((rank + ties < need ? 1 : 0) != chosen) ? 2 : 0;
                }
                if (tid < na) {
                    const float mine = sLStr[vcount + tid];
                    const int ps = sUPos[tid];
                    track_ratio(mine, sScore[ps], eps_level);
                    const uint32_t ks = order_key(mine);
                    int rank = 0, ties = 0;
                    for (int q = 0; q < na; q++) {
                        const uint32_t kq = order_key(sLStr[vcount + q]);
                        rank += kq > ks ? 1 : 0;
                        ties += (kq == ks && q != tid) ? 1 : 0;
                    }
                    bad |= (ties > 0 && rank < kk) ? 1 : 0;          // a tie that reaches the output: its order is positional
                    if (rank < kk) {
                        p.out_items[(size_t)user * p.out_stride + rank] = __ldg(p.leaf_item + ((int64_t)cur[ps] - leaf_start));
                        p.out_scores[(size_t)user * p.out_stride + rank] = mine;
                    }
                }
                if (__syncthreads_or(bad)) { redo = true; redo_why = 4 + (__syncthreads_or(bad & 2) ? 1 : 0); }
            }
            if (!redo) {
                for (int i = kk + tid; i < p.topk; i += G::THREADS) {
                    p.out_items[(size_t)user * p.out_stride + i] = -1;
                    p.out_scores[(size_t)user * p.out_stride + i] = 0.0f;
                }
                if (tid == 0) p.out_counts[user] = kk;
            }
        }
        if (redo) {
            st_redo++;
            if (tid == 0) {
                fp.redo_list[atomicAdd(fp.redo_count, 1)] = user;
                *reinterpret_cast<volatile int32_t *>(fp.host_flags + 1) = 1;
                if (fp.stats) atomicAdd(&fp.stats[24 + redo_why], 1ull);
            }
        }
        DMG_TICK(TK_FINAL);
    }
#ifdef DMG_FAST_TIMING
    if (fp.stats && tid == 0)
        for (int i = 0; i < TK_N; i++) atomicAdd(&fp.stats[8 + i], (unsigned long long)tacc[i]);
#endif
    if (fp.stats) {
        for (int o = 16; o > 0; o >>= 1) st_ratio = fmaxf(st_ratio, __shfl_xor_sync(0xffffffffu, st_ratio, o));
        if (lane == 0) atomicMax(reinterpret_cast<unsigned int *>(&fp.stats[4]), __float_as_uint(st_ratio));
        if (tid == 0) {
            atomicAdd(&fp.stats[0], st_cuts); atomicAdd(&fp.stats[1], st_recuts);
            atomicAdd(&fp.stats[2], st_rerows); atomicAdd(&fp.stats[3], st_rows);
            atomicAdd(&fp.stats[5], st_redo); atomicAdd(&fp.stats[6], st_sync); atomicAdd(&fp.stats[7], st_iters);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ---- per-level maxima of the node-dependent bound terms (run once per weight load) ----------------
// lvl_vx[l] = max over nodes c of level l of sum_k v[k] |emb[c][k]|, lvl_nx[l] = max |emb[c]|_2 (both rounded up).
static __global__ void level_bounds_kernel(const float *__restrict__ emb, int64_t rows, const float *__restrict__ v,
                                           float *__restrict__ lvl_vx, float *__restrict__ lvl_nx)
{
    __shared__ int s_vx[32], s_nx[32];
    if (threadIdx.x < 32) { s_vx[threadIdx.x] = 0; s_nx[threadIdx.x] = 0; }
    __syncthreads();
    const int lane16 = threadIdx.x & 15;
    const float4 vv = *reinterpret_cast<const float4 *>(v + lane16 * 4);
    const int64_t groups = (int64_t)gridDim.x * (blockDim.x >> 4);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 4) + (threadIdx.x >> 4); r < rows + 15; r += groups) {
        const bool ok = r < rows;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) x = ldg_row16(emb + r * 64 + lane16 * 4);
        float a = fmaf(vv.x, fabsf(x.x), fmaf(vv.y, fabsf(x.y), fmaf(vv.z, fabsf(x.z), vv.w * fabsf(x.w))));
        float n = fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, x.w * x.w)));
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); n += __shfl_xor_sync(0xffffffffu, n, o); }
        if (ok && lane16 == 0) {
            a *= 1.0001f;
            n = sqrtf(n) * 1.0001f;
            if (!(a == a)) a = __int_as_float(0x7f800000);
            if (!(n == n)) n = __int_as_float(0x7f800000);
            const int lvl = 63 - __clzll((unsigned long long)(r + 1));
            atomicMax(&s_vx[lvl & 31], __float_as_int(a));
            atomicMax(&s_nx[lvl & 31], __float_as_int(n));
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        atomicMax(reinterpret_cast<int *>(lvl_vx) + threadIdx.x, s_vx[threadIdx.x]);
        atomicMax(reinterpret_cast<int *>(lvl_nx) + threadIdx.x, s_nx[threadIdx.x]);
    }
}

}  // namespace dmg
