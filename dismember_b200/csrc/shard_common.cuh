// shard_common.cuh -- what the sharded entry points of shard.cu and dr.cu share: the NCCL functions resolved with dlopen
// (no link-time dependency on libnccl) and the per-handle communicator state.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include "dmg_common.cuh"

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

inline NcclApi g_nccl;

inline const char *load_nccl()
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
struct ShardState {
    int world = 1, rank = 0, bits = 0;
    ncclComm_t comm = nullptr;
    int64_t global_rows = 0;
    int64_t exchanged_rows = 0;   // candidates scored for another rank (statistics)
    Scratch buf;                  // level-loop buffers
    ShardGeo geo() const { ShardGeo g; g.bits = bits; g.rank = rank; g.world = world; g.repl_rows = ((int64_t)1 << bits) - 1; return g; }
};

}  // namespace dmg
