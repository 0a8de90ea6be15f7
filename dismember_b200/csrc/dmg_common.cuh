// dmg_common.cuh -- handle, error plumbing and small device helpers.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/dismember_gpu.h"

namespace dmg {

constexpr int kMaxT = 16;          // history length supported by the fused kernels (reference: 10)
constexpr int kThreads = 256;      // CTA size of the beam-search / scorer kernels
#define DMG_FAST_CTL_WORDS 264     // fast kernel control words: [0] user counter, [1] redo count, [2] tail counter, [8 + smid] tail owners

// ---- device-side index structures -------------------------------------------------------
struct TreeDev {
    bool loaded = false;
    bool complete = false;         // OTM / synthetic complete tree: every code exists
    int max_level = 0;             // leaf level L
    int64_t n_codes = 0;           // 2^(L+1)-1
    uint32_t *d_exists = nullptr;  // bitmap over codes (null when complete)
    int32_t *d_leaf_item = nullptr;  // [2^L] item id held by leaf slot, -1 = empty
    int32_t *d_id_code = nullptr;  // TDM: [non_leaf_offset] item id -> code, -1 = unknown
    int32_t non_leaf_offset = 0;   // DistTree.scala:35
    int32_t max_code = -1;         // DistTree.scala:36
    int64_t n_items = 0;
    int sparse_from = 0;           // lowest level with a missing code (max_level + 1 when every level is full)
    double *d_cdf = nullptr;       // per-level inclusive prefix sums of Node.probality over the codes (NegativeSampler.levelProbs); null = no probabilities
};

struct DinDev {
    bool loaded = false;
    int dtype = DMG_F32;
    int64_t rows = 0;
    int E = 0, T = 0;
    int scale_E = 0;               // != 0: the attention scale is 1 / sqrt(scale_E) (a zero-padded copy of a narrower model, capi.cu: build_padded_model)
    size_t esz = 4;
    void *d_params = nullptr;      // compact vector [emb | Watt | W1 | b1 | W2 | b2]
    void *d_wattT = nullptr;       // k-major copies  WattT[k][o], W1T[k][o]
    void *d_w1T = nullptr;
    int64_t n_params = 0;
    int kind = 0;                  // 0 = DIN [emb|Watt|W1|b1|W2|b2], 1 = DeepFM [emb|W1 (T+1)x(T+1)E|b1 T+1|W2 T+1|b2] (shard.cu)
    bool sharded = false;          // rows are this rank's slice of a table split by dmg_shard_init (shard.cu): shard entry points only
    // training state (allocated lazily)
    void *d_grad = nullptr, *d_m = nullptr, *d_v = nullptr;
    // wave path (beam_wave.cuh): bf16 hi|lo copy of the table (256 B per row) and the W1x operand image; owned by the model's owner handle
    unsigned char *d_split = nullptr, *d_w1img = nullptr;
    template <typename real> real *emb() const { return (real *)d_params; }
    template <typename real> real *tail() const { return (real *)d_params + rows * E; }      // dense parameters behind the table
    template <typename real> real *watt() const { return (real *)d_params + rows * E; }
    template <typename real> real *w1() const { return watt<real>() + (int64_t)E * E; }
    template <typename real> real *b1() const { return w1<real>() + (int64_t)2 * E * E; }
    template <typename real> real *w2() const { return b1<real>() + E; }
    template <typename real> real *b2() const { return w2<real>() + E; }
};

struct DrDev {
    bool loaded = false, paths_loaded = false;
    bool sharded = false;          // item-indexed tables hold this rank's item range only (dmg_shard_dr_load)
    int64_t local_items = 0, item_base = 0, item_chunk = 0;   // rows [item_base, item_base + local_items) of the item tables
    int num_item = 0, K = 0, D = 0, T = 0, E = 0;
    double *d_layer_emb = nullptr;
    std::vector<double *> d_layer_w, d_layer_b;     // device pointers per layer
    std::vector<double *> d_layer_wT;               // in-major copies [in][K]
    double *d_rr_emb = nullptr, *d_rr_w = nullptr, *d_rr_b = nullptr, *d_sm_w = nullptr, *d_sm_b = nullptr;
    int64_t *d_path_off = nullptr;
    int32_t *d_path_items = nullptr;
    std::vector<int64_t> h_path_off;                // kept for output sizing
    // training (dr_train.cu): tensors in the order layer_emb, (layer_w[d], layer_b[d]) d < D, rr_emb, rr_w (kept [T E][E]), rr_b, sm_w, sm_b
    std::vector<double *> tr_g, tr_s, tr_r;         // gradient, Adam first / second moment per tensor (allocated by the first training call)
    int32_t *d_item_paths = nullptr;                // itemPathMapping [num_item][P][D]
    int P = 0;
    bool tr_dirty = false;                          // gradients left in place by a step with apply = 0: zero them before the next one
};

struct ShardState;                  // shard.cu

// One captured retrieval step (capi.cu: tdm_enqueue): K2 + the whole level-synchronous chain of a batch as ONE graph launch.
struct StepGraph {
    uint64_t key[16];               // every argument and every pointer the captured kernels were given
    int seen = 0;                   // calls with this key so far (the second one captures: all scratch is allocated by then)
    bool bad = false;               // capture failed once: stay on plain launches
    cudaGraphExec_t exec = nullptr;
    int64_t launches = 0;           // kernels inside
    unsigned char redo[256];        // the strict redo launch the synchronous callers decide on (BeamParams<float>)
    bool has_redo = false;
};

struct Scratch {                    // grow-only device / pinned buffers
    void *d = nullptr; size_t d_bytes = 0;
    void *h = nullptr; size_t h_bytes = 0;
};

}  // namespace dmg

struct dmg_handle_s {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    size_t smem_optin = 0, smem_per_sm = 0;
    std::string err;
    int64_t launches = 0;
    dmg::TreeDev tree;
    dmg::DinDev din;
    dmg::DinDev din_pad;             // E = 16 / 32 Float DIN models: zero-padded E = 64 copy that the tensor-core retrieval path runs on
    dmg::DrDev dr;
    dmg::Scratch s_in, s_out, s_work, s_wave;
    bool wave_ok = false;            // level-synchronous tensor-core path (beam_wave.cuh) usable with the current tables
    int wave_mode = 0;               // 0 = TMA gather4 row gather, 1 = cp.async row gather
    alignas(64) unsigned char wave_tmap[128];   // CUtensorMap over the hi|lo table
    int32_t *d_flags = nullptr;     // [0] = index error flag: device alias of h_flags (mapped pinned memory: kernels raise it with no copy back)
    int32_t *h_flags = nullptr;
    int arithmetic = DMG_ARITH_FAST;   // same ids and logit bits as DMG_ARITH_STRICT (certified cuts); strict is the opt-out
    bool fast_ok = false;            // tensor-core scorer available for the loaded weights
    bool fast_dirty = true;          // weights changed since the bound tables were computed
    float fast_cA = 0, fast_cZ = 0, fast_cH = 0, fast_cGamma = 0;   // certified-cut bound constants
    float fast_tau = 1.0f;           // fraction of the worst-case bound used as the certification band
    float *d_fast_tab = nullptr;     // [0,4096) M^T  [4096,4160) v  [4160,4192) lvl_vx  [4192,4224) lvl_nx  [4224,4288) z
    std::vector<float> fast_host;    // host copy of b1 | w2 | b2 (kernel parameters of the fast kernel)
    int32_t *d_fast_ctl = nullptr;   // [0] dynamic user counter, [1] redo count, [2] tail user counter, [8 + smid] tail owner flags
    int32_t *d_redo_list = nullptr;  // users the fast kernel hands to the strict kernel
    int64_t redo_cap = 0;
    unsigned long long *d_fast_stats = nullptr;
    dmg::ShardState *shard = nullptr;   // node-table sharding over NCCL (shard.cu)
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    std::vector<dmg::StepGraph> graphs;  // CUDA graphs of recent retrieval steps (one cudaGraphLaunch instead of ~31 kernel launches)
    bool last_enqueue_wave = false;
    cudaEvent_t sync_ev = nullptr;       // blocking-sync event: a host thread waiting for its batch sleeps instead of spinning on a core
    alignas(8) unsigned char dfm_consts[256];   // DfmConsts of the loaded DeepFM model (beam_wave_dfm.cuh), valid when !fast_dirty
    int sync_mode = 0;                   // dmg_set_sync_mode: 0 spin (lowest latency), 1 sleep (hosts with more waiting threads than cores)
    dmg_handle_s *parent = nullptr;     // dmg_clone: tree / weight tables are the parent's (read-only here, never freed here)
    std::atomic<int> n_clones{0};       // live clones: the model of this handle is frozen until they are destroyed
};

namespace dmg {

inline int32_t fail(dmg_handle_t h, int32_t code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    return code;
}

// Entry points that replace or update the model refuse clones and handles with live clones (dmg_clone).
inline int32_t model_is_shared(dmg_handle_t h, const char *what)
{
    if (h && h->parent) return fail(h, DMG_ERR_STATE, "%s: this handle is a clone (dmg_clone) and shares its parent's model read-only", what);
    if (h && h->n_clones.load() > 0) return fail(h, DMG_ERR_STATE, "%s: %d clone(s) share this model -- destroy them first", what, h->n_clones.load());
    return DMG_OK;
}

#define DMG_CUDA(h, expr)                                                                         \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return dmg::fail(h, DMG_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                    \
                             cudaGetErrorString(e_), __FILE__, __LINE__);                         \
    } while (0)

#define DMG_TRY(expr)                                                                             \
    do {                                                                                          \
        int32_t rc_ = (expr);                                                                     \
        if (rc_ != DMG_OK) return rc_;                                                            \
    } while (0)

inline int32_t ensure_dev(dmg_handle_t h, Scratch &s, size_t bytes)
{
    if (bytes <= s.d_bytes) return DMG_OK;
    if (s.d) cudaFree(s.d);
    s.d = nullptr; s.d_bytes = 0;
    size_t want = bytes + bytes / 4 + 256;
    DMG_CUDA(h, cudaMalloc(&s.d, want));
    s.d_bytes = want;
    return DMG_OK;
}

inline int32_t ensure_host(dmg_handle_t h, Scratch &s, size_t bytes)
{
    if (bytes <= s.h_bytes) return DMG_OK;
    if (s.h) cudaFreeHost(s.h);
    s.h = nullptr; s.h_bytes = 0;
    size_t want = bytes + bytes / 4 + 256;
    DMG_CUDA(h, cudaMallocHost(&s.h, want));
    s.h_bytes = want;
    return DMG_OK;
}

// Bump allocator over one scratch block (256 B aligned slices).
struct Carver {
    char *base; size_t off = 0;
    explicit Carver(void *p) : base((char *)p) {}
    template <typename T> T *take(size_t n)
    {
        T *p = (T *)(base + off);
        off += (n * sizeof(T) + 255) & ~(size_t)255;
        return p;
    }
    static size_t need(std::initializer_list<size_t> bytes)
    {
        size_t t = 0;
        for (size_t b : bytes) t += (b + 255) & ~(size_t)255;
        return t;
    }
};

}  // namespace dmg
