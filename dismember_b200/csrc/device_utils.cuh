// device_utils.cuh -- device helpers shared by the kernels: sort keys and the block-wide bitonic
// sort that reproduce the reference's stable descending sorts, vector shared-memory access,
// cp.async / TMA bulk-copy / mbarrier wrappers, block scan.
#pragma once
#include "dmg_common.cuh"
#include "dmg_math.cuh"

namespace dmg {

// ---- sort keys: (score desc, candidate position asc) == the reference's stable sort ------
struct Key128 { uint64_t hi, lo; };
__device__ __forceinline__ bool key_less(uint64_t a, uint64_t b) { return a < b; }
__device__ __forceinline__ bool key_less(const Key128 &a, const Key128 &b)
{
    return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo);
}
template <typename real> struct KeyOf;
template <> struct KeyOf<float> {
    using type = uint64_t;
    static __device__ __forceinline__ type make(float s, int pos)
    {
        return ((uint64_t)order_key(s) << 32) | (uint32_t)(0xFFFFFFFFu - (uint32_t)pos);
    }
    static __device__ __forceinline__ type lowest() { return 0; }
    static __device__ __forceinline__ bool is_lowest(type k) { return k == 0; }
    static __device__ __forceinline__ int pos(type k) { return (int)(0xFFFFFFFFu - (uint32_t)k); }
};
template <> struct KeyOf<double> {
    using type = Key128;
    static __device__ __forceinline__ type make(double s, int pos)
    {
        Key128 k; k.hi = order_key(s); k.lo = (uint64_t)(0xFFFFFFFFu - (uint32_t)pos); return k;
    }
    static __device__ __forceinline__ type lowest() { Key128 k; k.hi = 0; k.lo = 0; return k; }
    static __device__ __forceinline__ bool is_lowest(const type &k) { return k.hi == 0 && k.lo == 0; }
    static __device__ __forceinline__ int pos(const type &k) { return (int)(0xFFFFFFFFu - (uint32_t)k.lo); }
};

template <typename K> __device__ void bitonic_sort_desc(K *keys, int n)
{
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    K a = keys[i], b = keys[ixj];
                    bool sw = ((i & k) == 0) ? key_less(a, b) : key_less(b, a);
                    if (sw) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
}

// 64-bit keys, n <= 2*blockDim.x (blockDim.x = kThreads): every thread keeps its two elements (tid, tid + 256)
// in registers; partner exchanges inside a warp (j < 32) are shuffles, j = 256 is register-local, and only
// the j in {32, 64, 128} steps go through shared memory with a barrier: 9 block barriers pairs instead of 45
// for 512 keys.  Result (descending) is written back to keys[0..n).
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m)
{
    uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, m), hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), m);
    return ((uint64_t)hi << 32) | lo;
}
__device__ inline void bitonic_sort_desc(uint64_t *keys, int n)
{
    const int tid = threadIdx.x;
    if (n > 2 * kThreads) {                                   // wide beams: plain shared-memory network
        for (int k = 2; k <= n; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < n; i += blockDim.x) {
                    int ixj = i ^ j;
                    if (ixj > i) {
                        uint64_t a = keys[i], b = keys[ixj];
                        bool sw = ((i & k) == 0) ? (a < b) : (b < a);
                        if (sw) { keys[i] = b; keys[ixj] = a; }
                    }
                }
                __syncthreads();
            }
        return;
    }
    const int i0 = tid, i1 = tid + kThreads;
    uint64_t k0 = i0 < n ? keys[i0] : 0, k1 = i1 < n ? keys[i1] : 0;
    __syncthreads();                                          // everyone has read its keys: smem is free for exchanges
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j == kThreads) {                              // partner is my other register
                // i0 has (i & j) == 0 -> lower; descending block iff (i0 & k) == 0 (k == 2*j here -> always)
                const bool desc = (i0 & k) == 0;
                const uint64_t mx = k0 > k1 ? k0 : k1, mn = k0 > k1 ? k1 : k0;
                k0 = desc ? mx : mn;
                k1 = desc ? mn : mx;
            } else if (j < 32) {
                const uint64_t p0 = shfl_xor_u64(k0, j), p1 = shfl_xor_u64(k1, j);
                const bool lower = (i0 & j) == 0;             // same for i0 and i1 (j < 256)
                const bool tmax0 = lower == ((i0 & k) == 0), tmax1 = lower == ((i1 & k) == 0);
                k0 = tmax0 ? (k0 > p0 ? k0 : p0) : (k0 > p0 ? p0 : k0);
                k1 = tmax1 ? (k1 > p1 ? k1 : p1) : (k1 > p1 ? p1 : k1);
            } else {                                          // cross-warp partner through shared memory
                if (i0 < n) keys[i0] = k0;
                if (i1 < n) keys[i1] = k1;
                __syncthreads();
                const uint64_t p0 = i0 < n ? keys[i0 ^ j] : 0;
                const uint64_t p1 = i1 < n ? keys[i1 ^ j] : 0;
                __syncthreads();
                const bool lower = (i0 & j) == 0;
                const bool tmax0 = lower == ((i0 & k) == 0), tmax1 = lower == ((i1 & k) == 0);
                k0 = tmax0 ? (k0 > p0 ? k0 : p0) : (k0 > p0 ? p0 : k0);
                k1 = tmax1 ? (k1 > p1 ? k1 : p1) : (k1 > p1 ? p1 : k1);
            }
        }
    }
    if (i0 < n) keys[i0] = k0;
    if (i1 < n) keys[i1] = k1;
    __syncthreads();
}

// ---- vector smem access -------------------------------------------------------------------
__device__ __forceinline__ void ld4(const float *p, float (&v)[4])
{
    float4 t = *reinterpret_cast<const float4 *>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void ld4(const double *p, double (&v)[4])
{
    double2 a = *reinterpret_cast<const double2 *>(p), b = *reinterpret_cast<const double2 *>(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void st4(float *p, const float (&v)[4])
{
    *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st4(double *p, const double (&v)[4])
{
    *reinterpret_cast<double2 *>(p) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2 *>(p + 2) = make_double2(v[2], v[3]);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void *dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// TMA bulk copy (non-tensor form) global -> shared, completion on an mbarrier.
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ bool code_exists(const uint32_t *bm, int64_t c)
{
    return bm == nullptr || ((__ldg(bm + (c >> 5)) >> (c & 31)) & 1u);
}

// Block-wide exclusive scan of one small int per thread (kThreads threads). Returns the
// exclusive prefix; *total receives the block sum.  sWarp: >= 8 ints of shared scratch.
__device__ __forceinline__ int block_exscan(int v, int *sWarp, int *total)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) sWarp[w] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < kThreads / 32; i++) {
        int s = sWarp[i];
        if (i < w) base += s;
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return base + inc - v;
}

}  // namespace dmg
