// capi.cu -- the C ABI of include/dismember_gpu.h: handle lifecycle, index/weight upload and
// the retrieval entry points (TDM/JTM, OTM, model.forward).  Deep Retrieval lives in dr.cu,
// training / JTM in train.cu.
#include <algorithm>
#include <cmath>
#include <mutex>

#include "beam_kernels.cuh"
#include "beam_fast.cuh"
#include "beam_wave.cuh"
#include "beam_wave_dfm.cuh"
#include "rows_kernels.cuh"

using namespace dmg;

static std::string g_create_error;

// Stream-ordered upload: a blocking cudaMemcpy from pageable memory may return before the DMA
// lands and is NOT ordered with the handle's non-blocking stream, so every upload goes through
// the stream the kernels run on and is waited for (the caller may free `src` on return).
static int32_t h2d(dmg_handle_t h, void *dst, const void *src, size_t bytes)
{
    if (bytes == 0) return DMG_OK;
    DMG_CUDA(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    return DMG_OK;
}

DMG_API const char *dmg_version(void) { return "dismember-b200 0.1 (sm_100a)"; }

DMG_API const char *dmg_last_error(dmg_handle_t h) { return h ? h->err.c_str() : g_create_error.c_str(); }

DMG_API int32_t dmg_create(int32_t device, dmg_handle_t *out)
{
    if (!out) return DMG_ERR_INVALID_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) +
                         " (this engine has no CPU fallback)";
        return DMG_ERR_CUDA;
    }
    if (device < 0 || device >= n) { g_create_error = "device ordinal out of range"; return DMG_ERR_INVALID_ARG; }
    dmg_handle_t h = new dmg_handle_s();
    h->device = device;
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaHostAlloc((void **)&h->h_flags, 64 * sizeof(int32_t), cudaHostAllocMapped)) != cudaSuccess ||
        (e = cudaHostGetDevicePointer((void **)&h->d_flags, h->h_flags, 0)) != cudaSuccess) {
        g_create_error = std::string("dmg_create: ") + cudaGetErrorString(e);
        delete h;
        return DMG_ERR_CUDA;
    }
    memset(h->h_flags, 0, 64 * sizeof(int32_t));
    if (prop.major < 10) {
        g_create_error = "dmg_create: device is not sm_100 (Blackwell) -- kernels are built for sm_100a only";
        cudaStreamDestroy(h->own_stream); cudaFreeHost(h->h_flags); delete h;
        return DMG_ERR_UNSUPPORTED;
    }
    h->stream = h->own_stream;
    h->sm_count = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    h->smem_per_sm = prop.sharedMemPerMultiprocessor;
    *out = h;
    return DMG_OK;
}

static void free_tree(TreeDev &t)
{
    cudaFree(t.d_exists); cudaFree(t.d_leaf_item); cudaFree(t.d_id_code); cudaFree(t.d_cdf);
    t = TreeDev();
}
static void free_din(DinDev &d)
{
    cudaFree(d.d_params); cudaFree(d.d_wattT); cudaFree(d.d_w1T); cudaFree(d.d_grad); cudaFree(d.d_m); cudaFree(d.d_v);
    cudaFree(d.d_split); cudaFree(d.d_w1img);
    d = DinDev();
}
static int32_t compute_fast_bounds(dmg_handle_t h);
static int32_t compute_dfm_bounds(dmg_handle_t h);
int dmg_shard_world(dmg_handle_t h);   // shard.cu
void dmg_shard_copy_geometry(dmg_handle_t dst, dmg_handle_t src);   // shard.cu
void dmg_free_dr(DrDev &d);     // dr.cu
void dmg_shard_free(dmg_handle_t h);   // shard.cu
int32_t dmg_deepfm_tdm_retrieve(dmg_handle_t h, int32_t B, const int32_t *item_seq, int32_t beam, int32_t topk, const int64_t *cons_off,
                                const int32_t *cons, int32_t widen_beam, int32_t *out_items, float *out_logits, int32_t *out_counts);   // shard.cu
int32_t dmg_deepfm_score_pairs(dmg_handle_t h, int64_t n, const int32_t *node, const int32_t *seq, float *out);                         // shard.cu
int32_t dmg_deepfm64_score_pairs(dmg_handle_t h, int64_t n, const int32_t *node, const int32_t *seq, double *out);                      // otm_deepfm.cu
int32_t dmg_deepfm64_otm_run(dmg_handle_t h, int32_t B, const int32_t *leaf_seq, int32_t beam, int topk_mode, int32_t topk, int32_t *out_ids,
                             double *out_scores, int32_t *out_counts, int32_t *lvl_ids, double *lvl_scores, int32_t *lvl_counts);       // otm_deepfm.cu

DMG_API int32_t dmg_destroy(dmg_handle_t h)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    if (h->n_clones.load() > 0) return fail(h, DMG_ERR_STATE, "dmg_destroy: %d clone(s) still share this model", h->n_clones.load());
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    dmg_shard_free(h);
    if (h->parent) h->parent->n_clones.fetch_sub(1);           // the tables are the parent's
    else {
        free_tree(h->tree);
        free_din(h->din);
        free_din(h->din_pad);
        dmg_free_dr(h->dr);
    }
    for (Scratch *s : {&h->s_in, &h->s_out, &h->s_work, &h->s_wave}) { cudaFree(s->d); cudaFreeHost(s->h); }
    for (auto &ev : h->prof_events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    for (StepGraph &g : h->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (h->sync_ev) cudaEventDestroy(h->sync_ev);
    cudaFreeHost(h->h_flags);
    cudaFree(h->d_fast_stats); cudaFree(h->d_fast_tab); cudaFree(h->d_fast_ctl); cudaFree(h->d_redo_list);
    cudaStreamDestroy(h->own_stream);
    delete h;
    return DMG_OK;
}

// A second handle on the same device that shares src's tree index and weight tables read-only: own stream, own scratch,
// own scheduler / redo state.  The GPU counterpart of the reference's per-thread model clones, which share one weight
// storage (tdm/.../optim/LocalOptimizer.scala:35-40; the evaluator splits the users over threads,
// tdm/.../evaluation/Evaluator.scala:29-37): one clone per host thread keeps several batches in flight, and the tail of
// one batch's persistent kernel overlaps the head of the next.
DMG_API int32_t dmg_clone(dmg_handle_t src, dmg_handle_t *out)
{
    if (!src || !out) return DMG_ERR_INVALID_ARG;
    *out = nullptr;
    if (src->parent) src = src->parent;                          // clones of clones hang off the owner
    const bool whole_deepfm = src->shard && dmg_shard_world(src) == 1 && src->din.loaded && src->din.kind == 1 && src->din.dtype == DMG_F32;
    if (src->shard && !whole_deepfm)
        return fail(src, DMG_ERR_UNSUPPORTED, "dmg_clone: sharded handles own NCCL and exchange state -- create one handle per rank");
    if (src->din.d_grad) return fail(src, DMG_ERR_STATE, "dmg_clone: this handle holds training state; clone an inference handle");
    dmg_handle_t h = nullptr;
    const int32_t rc = dmg_create(src->device, &h);
    if (rc != DMG_OK) return fail(src, rc, "dmg_clone: %s", dmg_last_error(nullptr));
    // the owner builds the shared tables of the tensor-core paths (bound tables, bf16 hi|lo copy of the node table) before
    // they are copied: clones only read them
    if (src->arithmetic == DMG_ARITH_FAST && src->fast_dirty && src->din.loaded && src->din.dtype == DMG_F32 && src->din.kind == 0 &&
        (src->din.E == 64 || src->din.E == 32 || src->din.E == 16)) {
        const int32_t rb = compute_fast_bounds(src);
        if (rb != DMG_OK) { dmg_destroy(h); return rb; }
    }
    if (whole_deepfm) {                                          // a Float DeepFM model on one GPU: the clone gets its own (world 1) exchange state
        if (src->arithmetic == DMG_ARITH_FAST && src->fast_dirty) {
            const int32_t rb = compute_dfm_bounds(src);
            if (rb != DMG_OK) { dmg_destroy(h); return rb; }
        }
        const int32_t rs = dmg_shard_init(h, 1, 0, nullptr);
        if (rs != DMG_OK) { dmg_destroy(h); return rs; }
        dmg_shard_copy_geometry(h, src);
    }
    cudaStreamSynchronize(src->stream);                          // uploads of the model are complete before another stream reads it
    h->tree = src->tree;
    h->din = src->din;
    h->din_pad = src->din_pad;
    h->dr = src->dr;
    h->arithmetic = src->arithmetic;
    h->sync_mode = src->sync_mode;
    h->fast_tau = src->fast_tau;
    h->fast_dirty = true;                                        // its own bound tables, computed on first use
    h->parent = src;
    src->n_clones.fetch_add(1);
    *out = h;
    return DMG_OK;
}

DMG_API int32_t dmg_set_stream(dmg_handle_t h, void *cuda_stream)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    return DMG_OK;
}

DMG_API int32_t dmg_synchronize(dmg_handle_t h)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    DMG_CUDA(h, cudaSetDevice(h->device));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    if (h->h_flags && *(volatile int32_t *)h->h_flags) {          // raised by a *_dev call since the last check (mapped pinned memory)
        h->h_flags[0] = 0;
        return fail(h, DMG_ERR_INDEX, "embeddingLookup failed in a *_dev call: index outside [-1, %lld)", (long long)h->din.rows);
    }
    return DMG_OK;
}

DMG_API int64_t dmg_launch_count(dmg_handle_t h) { return h ? h->launches : 0; }

DMG_API int32_t dmg_set_profiling(dmg_handle_t h, int32_t on)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    h->profiling = on != 0;
    return DMG_OK;
}

DMG_API int32_t dmg_kernel_time(dmg_handle_t h, double *total_ms, int64_t *n_launches)
{
    if (!h || !total_ms || !n_launches) return DMG_ERR_INVALID_ARG;
    DMG_CUDA(h, cudaSetDevice(h->device));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    double tot = 0.0;
    for (auto &ev : h->prof_events) {
        float ms = 0.f;
        DMG_CUDA(h, cudaEventElapsedTime(&ms, ev.first, ev.second));
        tot += ms;
        cudaEventDestroy(ev.first); cudaEventDestroy(ev.second);
    }
    *total_ms = tot;
    *n_launches = (int64_t)h->prof_events.size();
    h->prof_events.clear();
    return DMG_OK;
}

// ------------------------------------------------------------------------------------- trees
DMG_API int32_t dmg_load_tree_tdm(dmg_handle_t h, int32_t max_level, int64_t n_nodes, const int32_t *codes,
                                  const int32_t *node_ids, const uint8_t *is_leaf, int64_t n_items,
                                  const int32_t *leaf_ids, const int32_t *leaf_codes, const float *prob)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    DMG_TRY(model_is_shared(h, "dmg_load_tree_tdm"));
    if (max_level < 0 || max_level > 29 || n_nodes <= 0 || n_items <= 0 || !codes || !node_ids || !is_leaf ||
        !leaf_ids || !leaf_codes)
        return fail(h, DMG_ERR_INVALID_ARG, "dmg_load_tree_tdm: bad arguments (max_level must be in [0,29])");
    DMG_CUDA(h, cudaSetDevice(h->device));
    const int64_t n_codes = ((int64_t)1 << (max_level + 1)) - 1;
    const int64_t leaf_start = ((int64_t)1 << max_level) - 1, n_slots = (int64_t)1 << max_level;
    std::vector<uint32_t> bm((size_t)((n_codes + 31) / 32), 0u);
    std::vector<int32_t> leaf_item((size_t)n_slots, -1);
    for (int64_t i = 0; i < n_nodes; i++) {
        const int64_t c = codes[i];
        if (c < 0 || c >= n_codes)
            return fail(h, DMG_ERR_INVALID_ARG, "dmg_load_tree_tdm: node code %lld outside [0, 2^(max_level+1)-1)", (long long)c);
        bm[(size_t)(c >> 5)] |= 1u << (c & 31);
        if (is_leaf[i]) {
            if (c < leaf_start)
                return fail(h, DMG_ERR_UNSUPPORTED,
                            "dmg_load_tree_tdm: leaf code %lld lies above max_level %d; every writer of the format "
                            "(TreeBuilder.flattenLeaves, JTMTree.writeTree) sinks leaves to max_level", (long long)c, max_level);
            leaf_item[(size_t)(c - leaf_start)] = node_ids[i];
        } else if (c >= leaf_start) {
            return fail(h, DMG_ERR_UNSUPPORTED, "dmg_load_tree_tdm: non-leaf node %lld at max_level", (long long)c);
        }
    }
    int32_t mx_id = -1, mx_code = -1;
    for (int64_t i = 0; i < n_items; i++) { mx_id = std::max(mx_id, leaf_ids[i]); mx_code = std::max(mx_code, leaf_codes[i]); }
    if (mx_id < 0) return fail(h, DMG_ERR_INVALID_ARG, "dmg_load_tree_tdm: no non-negative leaf id");
    const int32_t offset = mx_id + 1;                                  // DistTree.scala:35
    std::vector<int32_t> id_code((size_t)offset, -1);
    for (int64_t i = 0; i < n_items; i++)
        if (leaf_ids[i] >= 0) id_code[(size_t)leaf_ids[i]] = leaf_codes[i];
    free_tree(h->tree);
    TreeDev &t = h->tree;
    DMG_CUDA(h, cudaMalloc(&t.d_exists, bm.size() * sizeof(uint32_t)));
    DMG_CUDA(h, cudaMalloc(&t.d_leaf_item, leaf_item.size() * sizeof(int32_t)));
    DMG_CUDA(h, cudaMalloc(&t.d_id_code, id_code.size() * sizeof(int32_t)));
    DMG_TRY(h2d(h, t.d_exists, bm.data(), bm.size() * sizeof(uint32_t)));
    DMG_TRY(h2d(h, t.d_leaf_item, leaf_item.data(), leaf_item.size() * sizeof(int32_t)));
    DMG_TRY(h2d(h, t.d_id_code, id_code.data(), id_code.size() * sizeof(int32_t)));
    if (prob) {                                                        // TreeNode.probality -> per-level cumulative weights (NegativeSampler.scala:59-66)
        std::vector<double> cdf((size_t)n_codes, 0.0);
        for (int64_t i = 0; i < n_nodes; i++) {
            if (!(prob[i] >= 0.0f)) return fail(h, DMG_ERR_INVALID_ARG, "dmg_load_tree_tdm: node probabilities must be >= 0");
            cdf[(size_t)codes[i]] = (double)prob[i];
        }
        for (int l = 0; l <= max_level; l++) {
            const int64_t a = ((int64_t)1 << l) - 1, b = ((int64_t)2 << l) - 1;
            for (int64_t c = a + 1; c < b; c++) cdf[(size_t)c] += cdf[(size_t)c - 1];
        }
        DMG_CUDA(h, cudaMalloc(&t.d_cdf, cdf.size() * sizeof(double)));
        DMG_TRY(h2d(h, t.d_cdf, cdf.data(), cdf.size() * sizeof(double)));
    }
    t.loaded = true; t.complete = false; t.max_level = max_level; t.n_codes = n_codes;
    t.non_leaf_offset = offset; t.max_code = mx_code; t.n_items = n_items;
    t.sparse_from = max_level + 1;                                     // levels above it are full: no bitmap probe on expansion
    for (int l = max_level; l >= 0; l--) {
        const int64_t a = ((int64_t)1 << l) - 1, b = ((int64_t)2 << l) - 1;
        bool full = true;
        for (int64_t w = a; w < b && full;) {
            if ((w & 31) == 0 && w + 32 <= b) { full = bm[(size_t)(w >> 5)] == 0xffffffffu; w += 32; }
            else { full = (bm[(size_t)(w >> 5)] >> (w & 31)) & 1u; w++; }
        }
        if (!full) t.sparse_from = l;
    }
    return DMG_OK;
}

DMG_API int32_t dmg_load_tree_complete(dmg_handle_t h, int32_t leaf_level, int64_t n_items, const int32_t *item_ids,
                                       const int32_t *leaf_ids)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    DMG_TRY(model_is_shared(h, "dmg_load_tree_complete"));
    if (leaf_level < 0 || leaf_level > 29 || n_items <= 0 || !item_ids || !leaf_ids)
        return fail(h, DMG_ERR_INVALID_ARG, "dmg_load_tree_complete: bad arguments");
    DMG_CUDA(h, cudaSetDevice(h->device));
    const int64_t leaf_start = ((int64_t)1 << leaf_level) - 1, n_slots = (int64_t)1 << leaf_level;
    std::vector<int32_t> leaf_item((size_t)n_slots, -1);
    for (int64_t i = 0; i < n_items; i++) {
        const int64_t s = (int64_t)leaf_ids[i] - leaf_start;
        if (s < 0 || s >= n_slots)
            return fail(h, DMG_ERR_INVALID_ARG, "dmg_load_tree_complete: leaf id %d not on level %d", leaf_ids[i], leaf_level);
        leaf_item[(size_t)s] = item_ids[i];
    }
    free_tree(h->tree);
    TreeDev &t = h->tree;
    DMG_CUDA(h, cudaMalloc(&t.d_leaf_item, leaf_item.size() * sizeof(int32_t)));
    DMG_TRY(h2d(h, t.d_leaf_item, leaf_item.data(), leaf_item.size() * sizeof(int32_t)));
    t.loaded = true; t.complete = true; t.max_level = leaf_level; t.n_codes = ((int64_t)1 << (leaf_level + 1)) - 1;
    t.sparse_from = leaf_level + 1;
    t.n_items = n_items;
    return DMG_OK;
}

// ----------------------------------------------------------------------------------- weights
static int32_t compute_fast_bounds(dmg_handle_t h);
template <typename real> static int32_t make_transposes(dmg_handle_t h)
{
    DinDev &d = h->din;
    const int E = d.E;
    DMG_CUDA(h, cudaMalloc(&d.d_wattT, sizeof(real) * E * E));
    DMG_CUDA(h, cudaMalloc(&d.d_w1T, sizeof(real) * 2 * E * E));
    transpose_kernel<real><<<(E * E + 255) / 256, 256, 0, h->stream>>>(d.watt<real>(), (real *)d.d_wattT, E, E);
    transpose_kernel<real><<<(2 * E * E + 255) / 256, 256, 0, h->stream>>>(d.w1<real>(), (real *)d.d_w1T, E, 2 * E);
    h->launches += 2;
    DMG_CUDA(h, cudaGetLastError());
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    return DMG_OK;
}

int32_t dmg_refresh_transposes(dmg_handle_t h)       // used by train.cu after an optimiser step
{
    DinDev &d = h->din;
    const int E = d.E;
    if (d.dtype == DMG_F32) {
        transpose_kernel<float><<<(E * E + 255) / 256, 256, 0, h->stream>>>(d.watt<float>(), (float *)d.d_wattT, E, E);
        transpose_kernel<float><<<(2 * E * E + 255) / 256, 256, 0, h->stream>>>(d.w1<float>(), (float *)d.d_w1T, E, 2 * E);
    } else {
        transpose_kernel<double><<<(E * E + 255) / 256, 256, 0, h->stream>>>(d.watt<double>(), (double *)d.d_wattT, E, E);
        transpose_kernel<double><<<(2 * E * E + 255) / 256, 256, 0, h->stream>>>(d.w1<double>(), (double *)d.d_w1T, E, 2 * E);
    }
    h->launches += 2;
    DMG_CUDA(h, cudaGetLastError());
    h->fast_dirty = true;                              // bound tables are rebuilt at the next fast retrieval
    return DMG_OK;
}

static int32_t alloc_din(dmg_handle_t h, int32_t dtype, int64_t rows, int32_t E, int32_t T)
{
    DMG_TRY(model_is_shared(h, "loading weights"));
    if (dtype != DMG_F32 && dtype != DMG_F64) return fail(h, DMG_ERR_INVALID_ARG, "dtype must be DMG_F32 or DMG_F64");
    if (rows <= 0 || E <= 0 || T <= 0) return fail(h, DMG_ERR_INVALID_ARG, "rows, E, T must be positive");
    if (T > kMaxT) return fail(h, DMG_ERR_UNSUPPORTED, "seq_len %d > %d", T, kMaxT);
    if (E % 4 != 0 || E > 256) return fail(h, DMG_ERR_UNSUPPORTED, "embed_size %d: must be a multiple of 4, <= 256", E);
    DMG_CUDA(h, cudaSetDevice(h->device));
    free_din(h->din);
    DinDev &d = h->din;
    d.dtype = dtype; d.rows = rows; d.E = E; d.T = T; d.esz = dtype == DMG_F32 ? 4 : 8;
    d.n_params = rows * E + (int64_t)E * E + (int64_t)2 * E * E + 2 * (int64_t)E + 1;
    DMG_CUDA(h, cudaMalloc(&d.d_params, (size_t)d.n_params * d.esz));
    return DMG_OK;
}

DMG_API int32_t dmg_load_din_weights(dmg_handle_t h, int32_t dtype, int64_t rows, int32_t E, int32_t T, const void *params)
{
    if (!h || !params) return h ? fail(h, DMG_ERR_INVALID_ARG, "dmg_load_din_weights: null params") : DMG_ERR_INVALID_ARG;
    DMG_TRY(alloc_din(h, dtype, rows, E, T));
    DinDev &d = h->din;
    DMG_TRY(h2d(h, d.d_params, params, (size_t)d.n_params * d.esz));
    DMG_TRY(dtype == DMG_F32 ? make_transposes<float>(h) : make_transposes<double>(h));
    h->fast_dirty = true;
    d.loaded = true;
    return DMG_OK;
}

DMG_API int32_t dmg_init_din_weights(dmg_handle_t h, int32_t dtype, int64_t rows, int32_t E, int32_t T, uint64_t seed)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    DMG_TRY(alloc_din(h, dtype, rows, E, T));
    DinDev &d = h->din;
    const int64_t n_rand = rows * E + (int64_t)E * E + (int64_t)2 * E * E;       // emb, W_att, W1
    const int grid = h->sm_count * 8;
    DMG_CUDA(h, cudaMemsetAsync(d.d_params, 0, (size_t)d.n_params * d.esz, h->stream));
    if (dtype == DMG_F32) {
        randn_fill_kernel<float><<<grid, 256, 0, h->stream>>>((float *)d.d_params, n_rand, seed, 0.05);
        randn_fill_kernel<float><<<1, 256, 0, h->stream>>>(d.w2<float>(), E, seed ^ 0x5bd1e995u, 0.05);
    } else {
        randn_fill_kernel<double><<<grid, 256, 0, h->stream>>>((double *)d.d_params, n_rand, seed, 0.05);
        randn_fill_kernel<double><<<1, 256, 0, h->stream>>>(d.w2<double>(), E, seed ^ 0x5bd1e995u, 0.05);
    }
    h->launches += 2;
    DMG_CUDA(h, cudaGetLastError());
    DMG_TRY(dtype == DMG_F32 ? make_transposes<float>(h) : make_transposes<double>(h));
    h->fast_dirty = true;
    d.loaded = true;
    return DMG_OK;
}

DMG_API int32_t dmg_din_shape(dmg_handle_t h, int64_t *rows, int32_t *E, int32_t *T, int32_t *dtype)
{
    if (!h || !rows || !E || !T || !dtype) return DMG_ERR_INVALID_ARG;
    if (!h->din.loaded) return fail(h, DMG_ERR_STATE, "no DIN / DeepFM weights loaded");
    *rows = h->din.rows; *E = h->din.E; *T = h->din.T; *dtype = h->din.dtype;
    return DMG_OK;
}

DMG_API int32_t dmg_download_din_weights(dmg_handle_t h, void *params, int64_t n)
{
    if (!h || !params) return DMG_ERR_INVALID_ARG;
    if (!h->din.loaded) return fail(h, DMG_ERR_STATE, "no DIN weights loaded");
    if (n != h->din.n_params) return fail(h, DMG_ERR_INVALID_ARG, "expected %lld elements", (long long)h->din.n_params);
    DMG_CUDA(h, cudaSetDevice(h->device));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    DMG_CUDA(h, cudaMemcpy(params, h->din.d_params, (size_t)n * h->din.esz, cudaMemcpyDeviceToHost));
    return DMG_OK;
}

// --------------------------------------------------------------------------------- retrieval
template <typename real, int E>
static int32_t launch_beam_E(dmg_handle_t h, const BeamParams<real> &p)
{
    using G = Geo<real, E>;
    const size_t smem = G::smem_bytes(p.cap, p.capp);
    if (smem > h->smem_optin)
        return fail(h, DMG_ERR_UNSUPPORTED, "beam too large: needs %zu B of shared memory per CTA (limit %zu)", smem, h->smem_optin);
    auto kern = beam_search_kernel<real, E>;
    DMG_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int split = p.split > 1 ? p.split : 1;
    const int grid = split * std::min(p.B, std::max(h->sm_count / split, 1));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->profiling) {
        DMG_CUDA(h, cudaEventCreate(&e0));
        DMG_CUDA(h, cudaEventCreate(&e1));
        DMG_CUDA(h, cudaEventRecord(e0, h->stream));
    }
    if (split > 1) {                                             // clusters of `split` CTAs share one user (beam_kernels.cuh)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = h->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)split; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        DMG_CUDA(h, cudaLaunchKernelEx(&cfg, kern, p));
    } else
    kern<<<grid, kThreads, smem, h->stream>>>(p);
    h->launches += 1;
    DMG_CUDA(h, cudaGetLastError());
    if (h->profiling) {
        DMG_CUDA(h, cudaEventRecord(e1, h->stream));
        h->prof_events.emplace_back(e0, e1);
    }
    return DMG_OK;
}

template <typename real> static int32_t launch_beam(dmg_handle_t h, const BeamParams<real> &p, int E)
{
    switch (E) {
        case 16: return launch_beam_E<real, 16>(h, p);
        case 32: return launch_beam_E<real, 32>(h, p);
        case 64: return launch_beam_E<real, 64>(h, p);
        default:
            return fail(h, DMG_ERR_UNSUPPORTED, "beam search kernels are built for embed_size 16, 32 and 64 (got %d)", E);
    }
}

// Tables and constants of the certified-cut bound, rebuilt lazily after every weight change:
//   M = W1a.Watt (double -> fp32, stored k-major), v[k] = sum_o |w2[o]| |W1x[o][k]|,
//   per-level maxima of v.|x| and |x|_2 over the node table, and
//   eps = tau * ( cA * max v.|x|  +  cBq * Kmax * (cCq + |dp|_1)  +  cGamma ).
// Tables of the level-synchronous path (beam_wave.cuh): the bf16 hi|lo copy of the node table and the W1x operand image
// belong to the model's owner (a clone reads its parent's, which dmg_clone has brought up to date); every handle encodes
// its own TMA descriptor over the copy.  No room for the copy (it is as large as the table) => the persistent kernel runs.
static int32_t wave_prepare(dmg_handle_t h)
{
    DinDev &d = h->din_pad.loaded ? h->din_pad : h->din;
    if (!h->parent) {
        if (!d.d_split && cudaMalloc(&d.d_split, (size_t)d.rows * 256) != cudaSuccess) { cudaGetLastError(); d.d_split = nullptr; return DMG_OK; }
        if (!d.d_w1img) { DMG_CUDA(h, cudaMalloc(&d.d_w1img, 24576)); DMG_CUDA(h, cudaMemsetAsync(d.d_w1img, 0, 24576, h->stream)); }
        wave_split_table_kernel<<<h->sm_count * 16, 256, 0, h->stream>>>(d.emb<float>(), d.rows, d.d_split);
        wave_w1_image_kernel<<<2, 256, 0, h->stream>>>(d.w1<float>(), d.d_w1img);
        h->launches += 2;
        DMG_CUDA(h, cudaGetLastError());
        DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    } else {
        const DinDev &pd = h->parent->din_pad.loaded ? h->parent->din_pad : h->parent->din;
        if (h->parent->fast_dirty || !pd.d_split) return DMG_OK;
        d.d_split = pd.d_split;
        d.d_w1img = pd.d_w1img;
    }
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) { cudaGetLastError(); return DMG_OK; }
        encode = (encode_fn)fn;
    }
    const cuuint64_t gdim[2] = {128, (cuuint64_t)d.rows};        // [rows][128] bf16: row c = [64 hi | 64 lo]
    const cuuint64_t gstr[1] = {256};
    const cuuint32_t box[2] = {64, 1}, estr[2] = {1, 1};         // tile::gather4 fetches 4 such boxes (half rows) per instruction: columns 0.. = hi, 64.. = lo
    const CUresult cr = encode(reinterpret_cast<CUtensorMap *>(h->wave_tmap), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d.d_split, gdim, gstr, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return DMG_OK;
    const char *m = getenv("DMG_WAVE_GATHER");
    h->wave_mode = (m && !strcmp(m, "cpasync")) ? 1 : 0;
    h->wave_ok = true;
    return DMG_OK;
}

// ---- E = 16 / 32 models on the E = 64 tensor-core path --------------------------------------------------------------------------------
// Every configuration the reference ships uses embed_size 16 (configs/*.conf).  A narrower Float DIN model zero-padded to E = 64 --
// table rows [x | 0], Watt / W1 / b1 / W2 with zero rows and columns, the item half and the attention half of W1 at columns 0 and 64
// -- runs through the SAME sequential-k chains with exact zeros appended (fma(0, 0, acc) == acc), and relu(0 + 0) . 0 adds nothing to
// the logit: the strict arithmetic of the padded model is the narrow model's bit for bit, PROVIDED the attention scale stays
// 1 / sqrt(embed_size) of the original (DinDev::scale_E).  So retrieval in FAST arithmetic builds that copy once per weight load and
// runs the E = 64 path on it: certified cuts, strict re-scores and the redo kernel included.  It costs 4x (E = 16) the table's memory
// and gather traffic, and takes E = 16 retrieval from 0.62 M (strict SIMT kernel) to the E = 64 path's rate.
static __global__ void pad_table_kernel(const float *__restrict__ src, int64_t rows, int E, float *__restrict__ dst)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * 64; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i >> 6;
        const int k = (int)(i & 63);
        dst[i] = k < E ? src[r * E + k] : 0.0f;
    }
}
// dense tail [Watt E x E | W1 E x 2E | b1 | W2 | b2] -> [64 x 64 | 64 x 128 | 64 | 64 | 1]
static __global__ void pad_dense_kernel(const float *__restrict__ src, int E, float *__restrict__ dst)
{
    const int n = 3 * 64 * 64 + 2 * 64 + 1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float v = 0.0f;
        if (i < 64 * 64) {                                            // Watt[o][k]
            const int o = i >> 6, k = i & 63;
            if (o < E && k < E) v = src[o * E + k];
        } else if (i < 3 * 64 * 64) {                                 // W1[o][c], c < 64 the item half, c >= 64 the attention half
            const int q = i - 64 * 64, o = q >> 7, c = q & 127, half = c >> 6, k = c & 63;
            if (o < E && k < E) v = src[E * E + o * 2 * E + half * E + k];
        } else {
            const int q = i - 3 * 64 * 64;                            // b1 | W2 | b2
            if (q < 64) { if (q < E) v = src[3 * E * E + q]; }
            else if (q < 128) { if (q - 64 < E) v = src[3 * E * E + E + (q - 64)]; }
            else v = src[3 * E * E + 2 * E];
        }
        dst[i] = v;
    }
}
static int32_t build_padded_model(dmg_handle_t h)
{
    const DinDev &d = h->din;
    DinDev &p = h->din_pad;
    if (h->parent) return DMG_OK;                                     // a clone reads its parent's copy (dmg_clone brought it up to date)
    const int64_t n_params = d.rows * 64 + 3 * 64 * 64 + 2 * 64 + 1;
    if (p.loaded && (p.rows != d.rows || p.n_params != n_params || p.kind != 0)) free_din(p);
    if (!p.d_params) {
        if (cudaMalloc(&p.d_params, (size_t)n_params * 4) != cudaSuccess) { cudaGetLastError(); p = DinDev(); return DMG_OK; }   // no room: strict kernel
        DMG_CUDA(h, cudaMalloc(&p.d_wattT, sizeof(float) * 64 * 64));
        DMG_CUDA(h, cudaMalloc(&p.d_w1T, sizeof(float) * 2 * 64 * 64));
    }
    p.dtype = DMG_F32; p.esz = 4; p.rows = d.rows; p.E = 64; p.T = d.T; p.scale_E = d.E; p.n_params = n_params; p.kind = 0;
    pad_table_kernel<<<h->sm_count * 16, 256, 0, h->stream>>>(d.emb<float>(), d.rows, d.E, p.emb<float>());
    pad_dense_kernel<<<64, 256, 0, h->stream>>>(d.tail<float>(), d.E, p.tail<float>());
    transpose_kernel<float><<<(64 * 64 + 255) / 256, 256, 0, h->stream>>>(p.watt<float>(), (float *)p.d_wattT, 64, 64);
    transpose_kernel<float><<<(2 * 64 * 64 + 255) / 256, 256, 0, h->stream>>>(p.w1<float>(), (float *)p.d_w1T, 64, 128);
    h->launches += 4;
    DMG_CUDA(h, cudaGetLastError());
    p.loaded = true;
    return DMG_OK;
}
// the model the retrieval kernels read: the zero-padded copy when it is built and current, else the loaded one
static const DinDev &retrieval_din(const dmg_handle_t h)
{
    return (h->din_pad.loaded && h->arithmetic == DMG_ARITH_FAST && h->fast_ok && !h->fast_dirty) ? h->din_pad : h->din;
}
static bool wants_padded_model(const dmg_handle_t h)
{
    const DinDev &d = h->din;
    return h->arithmetic == DMG_ARITH_FAST && d.loaded && d.dtype == DMG_F32 && d.kind == 0 && !d.sharded && (d.E == 16 || d.E == 32) && d.T <= 15 &&
           !getenv("DMG_NO_PAD");
}

static int32_t compute_fast_bounds(dmg_handle_t h)
{
    h->fast_ok = false;
    h->fast_dirty = false;
    if (wants_padded_model(h)) DMG_TRY(build_padded_model(h));
    else if (!h->parent && h->din_pad.loaded) free_din(h->din_pad);
    DinDev &d = h->din_pad.loaded ? h->din_pad : h->din;
    if (d.dtype != DMG_F32 || d.E != 64 || d.kind != 0) return DMG_OK;
    const int E = d.E;
    const size_t n_dense = (size_t)3 * E * E + 2 * E + 1;
    std::vector<float> w(n_dense);
    DMG_CUDA(h, cudaMemcpyAsync(w.data(), d.watt<float>(), n_dense * 4, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    const float *watt = w.data(), *w1 = watt + E * E, *b1 = w1 + 2 * E * E, *w2 = b1 + E, *b2 = w2 + E;
    std::vector<float> tab(4288, 0.0f);                          // M^T | v | lvl_vx | lvl_nx | z
    // u = fp32 unit roundoff; c_mma = bf16 hi/lo split (3 * 2^-18 = 1.14e-5 per product) + fp32 accumulation inside
    // the tensor core (allowance 0.86e-5 of sum |terms|).  A length-n fma chain perturbs its k-th term (0-based) by at
    // most (n - k) roundings, so every weight below carries the position of its term in the strict chains.
    const double u = std::ldexp(1.0, -24), c_mma = 2.0e-5, safety = 1.05;
    double gam = 0;
    std::vector<double> z(E, 0.0), vt(E, 0.0), fo(E);
    for (int o = 0; o < E; o++) fo[o] = (2.0 * (E - o) + 4.0) * u;            // logit = h . W2 + b2, both paths
    for (int o = 0; o < E; o++) {
        const double aw2 = std::fabs((double)w2[o]);
        gam += aw2 * std::fabs((double)b1[o]) * (c_mma + 2 * u + fo[o]);
        for (int k = 0; k < E; k++) {
            double acc = 0;
            for (int m = 0; m < E; m++) {
                const double t = (double)w1[o * 2 * E + E + m] * (double)watt[m * E + k];
                acc += t;
                // strict: att term m sits at position 64+m of the h chain, a_k at position k of the att chain, a is a T-chain;
                // fast:   H = M.K is a 64-chain over k in fp32 and M was rounded once
                z[k] += aw2 * std::fabs(t) * ((E - m) + 2.0 * (E - k) + 18.0) * u;
            }
            tab[(size_t)k * E + o] = (float)acc;                  // M^T[k][o]
            // main branch: MMA, position k of the 128-chain, the b1 add, the final dot
            vt[k] += aw2 * std::fabs((double)w1[o * 2 * E + k]) * (c_mma + (2.0 * E - k) * u + 2 * u + fo[o]);
        }
    }
    for (int k = 0; k < E; k++) {
        tab[4096 + k] = (float)(vt[k] * (1.0 + 1e-6));
        tab[4224 + k] = (float)(z[k] * (1.0 + 1e-6));
    }
    h->fast_cA = (float)safety;                                   // eps = tau (cA max v.|x| + cZ z.Kabs + (cH + |dp|_1) HW + cGamma)
    h->fast_cZ = (float)safety;
    h->fast_cH = (float)(c_mma + 134 * u);                        // attention branch on the tensor cores + final dot
    h->fast_cGamma = (float)(safety * (gam + 4 * u * std::fabs((double)b2[0])) + 1e-30);
    h->fast_host.assign(b1, b1 + 2 * E + 1);                     // b1 | w2 | b2
    if (!h->d_fast_tab) DMG_CUDA(h, cudaMalloc(&h->d_fast_tab, tab.size() * sizeof(float)));
    if (!h->d_fast_ctl) {
        DMG_CUDA(h, cudaMalloc(&h->d_fast_ctl, DMG_FAST_CTL_WORDS * sizeof(int32_t)));
        DMG_CUDA(h, cudaMemsetAsync(h->d_fast_ctl, 0, DMG_FAST_CTL_WORDS * sizeof(int32_t), h->stream));
    }
    DMG_CUDA(h, cudaMemcpyAsync(h->d_fast_tab, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    level_bounds_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(d.emb<float>(), d.rows, h->d_fast_tab + 4096, h->d_fast_tab + 4160,
                                                                h->d_fast_tab + 4192);
    h->launches += 1;
    DMG_CUDA(h, cudaGetLastError());
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));               // tab / w are stack-owned
    h->fast_ok = std::isfinite(h->fast_cA) && std::isfinite(h->fast_cGamma);
    h->wave_ok = false;
    const char *impl = getenv("DMG_FAST_IMPL");
    if (h->fast_ok && !(impl && !strcmp(impl, "v1"))) DMG_TRY(wave_prepare(h));
    return DMG_OK;
}

template <typename real> static void fill_scorer(const DinDev &d, BeamParams<real> &p)
{
    p.emb = d.emb<real>(); p.wattT = (const real *)d.d_wattT; p.w1T = (const real *)d.d_w1T;
    p.b1 = d.b1<real>(); p.w2 = d.w2<real>(); p.b2 = d.b2<real>();
    p.scale = (real)(1.0 / std::sqrt((double)(d.scale_E ? d.scale_E : d.E)));   // Mask.scala:12
    p.T = d.T;
}

static int pow2_ge(int n) { int p = 2; while (p < n) p <<= 1; return p; }
static int lower_log2(int n) { int l = 0; while ((2 << l) <= n) l++; return l; }   // == floor(log(n)/log(2)) for n >= 1

// One handle per host thread keeps several batches in flight.  cudaStreamSynchronize spins on a core, which is the lowest latency
// while the host has cores to spare (one GPU: 2.27 M users/s e2e against 2.07 M with a sleeping wait); a host that drives 8 GPUs
// with 8 threads each on 32 cores is better off with waiting threads that sleep on a cudaEventBlockingSync event (18.0 M against
// 14.4 M users/s): dmg_set_sync_mode(h, 1).
static int32_t wait_stream(dmg_handle_t h)
{
    if (h->sync_mode == 0) { DMG_CUDA(h, cudaStreamSynchronize(h->stream)); return DMG_OK; }
    if (!h->sync_ev) DMG_CUDA(h, cudaEventCreateWithFlags(&h->sync_ev, cudaEventBlockingSync | cudaEventDisableTiming));
    DMG_CUDA(h, cudaEventRecord(h->sync_ev, h->stream));
    DMG_CUDA(h, cudaEventSynchronize(h->sync_ev));
    return DMG_OK;
}

static int32_t check_flag(dmg_handle_t h, const char *what)
{
    DMG_TRY(wait_stream(h));
    const int32_t flag = *(volatile int32_t *)h->h_flags;        // mapped pinned memory: nothing to copy
    if (flag) {
        h->h_flags[0] = 0;
        return fail(h, DMG_ERR_INDEX, "%s: embeddingLookup failed, index outside [0, %lld)", what, (long long)h->din.rows);
    }
    return DMG_OK;
}

// K2 on its own (used by the sampler in train.cu): TDMTree.idToCode on device-resident ids.
int32_t dmg_tdm_ids_to_codes(dmg_handle_t h, const int32_t *d_ids, int64_t n, int use_mask, int32_t *d_codes, uint8_t *d_mask)
{
    const TreeDev &t = h->tree;
    tdm_ids_to_codes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(d_ids, n, t.d_id_code, t.non_leaf_offset, t.max_code,
                                                                              h->din.rows, use_mask, d_codes, d_mask, h->d_flags, nullptr);
    h->launches += 1;
    DMG_CUDA(h, cudaGetLastError());
    return DMG_OK;
}

// One kernel of a programmatic-dependent-launch chain (beam_wave.cuh: grid_dep_launch / grid_dep_wait).
template <typename... KArgs, typename... Args>
static cudaError_t launch_chain(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, bool pdl, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// The level-synchronous tensor-core path (beam_wave.cuh): prologue, then select + score per tree level, then the strict
// final.  p / fx are the persistent kernel's parameter blocks (same tables, same outputs, same redo list).
static int32_t wave_enqueue(dmg_handle_t h, const BeamParams<float> &p, const FastParams &fx, int max_beam, int stop_level = -1,
                            WaveParams *wp_out = nullptr, int *slot_out = nullptr)
{
    using WG = WaveGeo;
    const DinDev &d = retrieval_din(h);
    const int B = p.B, cap = p.cap;
    DMG_TRY(ensure_dev(h, h->s_wave, Carver::need({(size_t)B * cap * 4, (size_t)B * cap * 4, (size_t)B * cap * 4, (size_t)B * 4,
                                                   (size_t)B * sizeof(WaveUser), (size_t)B * WG::UOP_BYTES, (size_t)B * WG::VCAP * 4,
                                                   (size_t)B * WG::VCAP * 4, (size_t)B * WG::VCAP * 4, (size_t)B * 32 * 4,
                                                   (size_t)B * ((cap + 127) / 128) * 4, (size_t)B * WaveFinal::RCAP * 4, (size_t)B * WaveFinal::RCAP * 4,
                                                   (size_t)B * FastGeo::MAX_FINAL * 4, (size_t)B * 4 * 4, (size_t)B * 27 * 4})));
    Carver c(h->s_wave.d);
    WaveParams wp;
    memset(&wp, 0, sizeof(wp));
    wp.B = B; wp.T = p.T; wp.cap = cap; wp.beam = p.beam; wp.beam_user = p.beam_user;
    wp.emb = p.emb; wp.hist = p.hist; wp.hist_mask = p.hist_mask; wp.exists = p.exists;
    wp.leaf_level = p.leaf_level; wp.sparse_from = fx.sparse_from; wp.scale = p.scale;
    wp.code[0] = c.take<int32_t>((size_t)B * cap); wp.code[1] = c.take<int32_t>((size_t)B * cap);
    wp.score = c.take<float>((size_t)B * cap);
    wp.count = c.take<int32_t>(B);
    wp.user = c.take<WaveUser>(B);
    wp.uop = c.take<unsigned char>((size_t)B * WG::UOP_BYTES);
    wp.v_code = c.take<int32_t>((size_t)B * WG::VCAP); wp.v_fast = c.take<float>((size_t)B * WG::VCAP);
    wp.v_meta = c.take<uint32_t>((size_t)B * WG::VCAP); wp.v_segeps = c.take<float>((size_t)B * 32);
    wp.tile_list = c.take<int32_t>((size_t)B * ((cap + 127) / 128));
    WaveFinal wf;
    wf.row_code = c.take<int32_t>((size_t)B * WaveFinal::RCAP); wf.row_strict = c.take<float>((size_t)B * WaveFinal::RCAP);
    wf.fin_pos = c.take<int32_t>((size_t)B * FastGeo::MAX_FINAL); wf.meta = c.take<int32_t>((size_t)B * 4);
    wf.chunk_list = c.take<int32_t>((size_t)B * 27);
    wf.chunk_count = h->d_fast_ctl + 8 + 40;                     // zeroed by tdm_ids_to_codes_kernel
    wp.tile_count = h->d_fast_ctl + 8;                           // [level]: tiles listed (zeroed by tdm_ids_to_codes_kernel)
    wp.mT = fx.mT; wp.zvec = fx.zvec; wp.lvl_vx = fx.lvl_vx; wp.lvl_nx = fx.lvl_nx; wp.b1 = p.b1;
    wp.cA = fx.cA; wp.cZ = fx.cZ; wp.cH = fx.cH; wp.cGamma = fx.cGamma; wp.tau = fx.tau;
    wp.stats = fx.stats; wp.redo_list = fx.redo_list; wp.redo_count = fx.redo_count; wp.host_flags = fx.host_flags;
    wp.w1img = d.d_w1img; wp.split = d.d_split;
    WaveW2 w2;
    memcpy(w2.w2, fx.w2, sizeof(w2.w2));
    w2.b2 = fx.b2;
    WaveStrictW sw;
    sw.wattT = p.wattT; sw.w1T = p.w1T; sw.b1 = p.b1; sw.w2 = p.w2; sw.b2 = fx.b2;
    const CUtensorMap &tmap = *reinterpret_cast<const CUtensorMap *>(h->wave_tmap);
    auto score_kernel = wave_score_kernel<0>;
    auto select_kernel = cap <= 256 ? wave_select_kernel<8> : (cap <= 416 ? wave_select_kernel<13> : wave_select_kernel<16>);
    DMG_CUDA(h, cudaFuncSetAttribute(score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG::SMEM));
    DMG_CUDA(h, cudaFuncSetAttribute(score_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    DMG_CUDA(h, cudaFuncSetAttribute(wave_strict_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WaveStrictGeo::smem_bytes()));
    const bool pdl = !getenv("DMG_WAVE_NO_PDL");
    DMG_CUDA(h, launch_chain(wave_prologue_kernel, B, 256, 0, h->stream, false, wp));   // first of the chain: ordinary stream order behind K2
    h->launches += 1;
    const int tpu = (cap + 127) / 128, ntiles = B * tpu;
    // One scorer CTA per SM when several batches are in flight on this model (dmg_clone: the scorers of two batches then share
    // every SM and one batch's select / strict kernels fill the other's gaps), two when the handle works alone (shortest batch).
    const bool shared_model = h->parent != nullptr || h->n_clones.load() > 0;
    const int ctas_per_sm = getenv("DMG_WAVE_SCORE_CTAS") ? std::max(1, std::min(2, atoi(getenv("DMG_WAVE_SCORE_CTAS")))) : (shared_model ? 1 : 2);
    const int grid = std::min(ntiles, ctas_per_sm * h->sm_count);
    int slot = 0;
    const int s_min = lower_log2(p.beam);                        // per-user beams only widen (Recommender.scala:28-31)
    (void)max_beam;
    const int last_level = stop_level >= 0 ? std::min(stop_level, p.leaf_level) : p.leaf_level;
    for (int level = s_min; level < last_level; level++) {
        DMG_CUDA(h, launch_chain(select_kernel, (B + 3) / 4, 128, 0, h->stream, pdl, wp, sw, level, slot, DfmConsts()));
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (h->profiling) {                                      // dmg_set_profiling: every launch of the dominant kernel timed alone
            DMG_CUDA(h, cudaEventCreate(&e0));
            DMG_CUDA(h, cudaEventCreate(&e1));
            DMG_CUDA(h, cudaEventRecord(e0, h->stream));
        }
        DMG_CUDA(h, launch_chain(score_kernel, grid, WG::THREADS, WG::SMEM, h->stream, pdl && !h->profiling, tmap, wp, w2, slot ^ 1, level + 1));
        if (h->profiling) {
            DMG_CUDA(h, cudaEventRecord(e1, h->stream));
            h->prof_events.emplace_back(e0, e1);
        }
#ifdef DMG_WAVE_ABLATION                                         // profiling build only (DMG_NVCC_EXTRA=-DDMG_WAVE_ABLATION): the DBG instantiations stay out of the product
        if (getenv("DMG_WAVE_ABLATE") && level == atoi(getenv("DMG_WAVE_ABLATE"))) {
            // profiling aid: replay this level's scorer with parts switched off (scores go to a scratch buffer)
            static float *dummy = nullptr;
            {
                cudaError_t e0 = cudaStreamSynchronize(h->stream);
                fprintf(stderr, "[wave ablate] before replay: %s\n", cudaGetErrorString(e0));
            }
            if (!dummy) { cudaError_t em = cudaMalloc(&dummy, (size_t)B * cap * 4); fprintf(stderr, "[wave ablate] malloc: %s %p\n", cudaGetErrorString(em), (void *)dummy); }
            WaveParams wq = wp;
            wq.score = dummy;
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            std::vector<int> masks = {0, 0, 1, 2, 4, 8, 16, 3, 6, 7, 15, 31};
            if (const char *ml = getenv("DMG_WAVE_MASKS")) { masks.clear(); for (const char *q = ml; *q; q++) if (*q >= '0' && *q <= '9') { masks.push_back(atoi(q)); while (q[1] >= '0' && q[1] <= '9') q++; } }
            for (int mask : masks) {
                float best = 1e9f;
                for (int rep = 0; rep < 5; rep++) {
                    cudaEventRecord(a, h->stream);
                    void (*kq)(const CUtensorMap, const WaveParams, const WaveW2, int, int) = nullptr;
                    switch (mask) {
                        case 0: kq = wave_score_kernel<0>; break;   case 1: kq = wave_score_kernel<1>; break;
                        case 2: kq = wave_score_kernel<2>; break;   case 4: kq = wave_score_kernel<4>; break;
                        case 8: kq = wave_score_kernel<8>; break;   case 16: kq = wave_score_kernel<16>; break;
                        case 3: kq = wave_score_kernel<3>; break;   case 6: kq = wave_score_kernel<6>; break;
                        case 7: kq = wave_score_kernel<7>; break;   case 15: kq = wave_score_kernel<15>; break;
                        case 30: kq = wave_score_kernel<30>; break;   case 14: kq = wave_score_kernel<14>; break;
                        case 22: kq = wave_score_kernel<22>; break;
                        default: kq = wave_score_kernel<31>; break;
                    }
                    cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, WG::SMEM);
                    kq<<<grid, WG::THREADS, WG::SMEM, h->stream>>>(tmap, wq, w2, slot ^ 1, level + 1);
                    cudaEventRecord(b, h->stream);
                    cudaEventSynchronize(b);
                    float ms = 0.f;
                    cudaError_t ee = cudaEventElapsedTime(&ms, a, b);
                    if (ee != cudaSuccess) { fprintf(stderr, "[wave ablate] error %s\n", cudaGetErrorString(ee)); break; }
                    best = std::min(best, ms);
                }
                fprintf(stderr, "[wave ablate] level %d mask %2d: %.1f us\n", level, mask, best * 1e3f);
            }
            cudaEventDestroy(a); cudaEventDestroy(b);
        }
#endif
        slot ^= 1;
        h->launches += 2;
    }
    if (wp_out) { *wp_out = wp; *slot_out = slot; }
    if (stop_level < 0) {
        auto prep_kernel = cap <= 256 ? wave_final_prep_kernel<8> : (cap <= 416 ? wave_final_prep_kernel<13> : wave_final_prep_kernel<16>);
        DMG_CUDA(h, launch_chain(prep_kernel, (B + 3) / 4, 128, 0, h->stream, pdl, wp, p, wf, slot));
        DMG_CUDA(h, launch_chain(wave_strict_rows_kernel, std::min(B, 2 * h->sm_count), kThreads, WaveStrictGeo::smem_bytes(), h->stream, pdl, wp, sw, wf));
        DMG_CUDA(h, launch_chain(wave_final_verify_kernel, (B + 3) / 4, 128, 0, h->stream, pdl, wp, p, wf, slot));
        h->launches += 3;
    }
    DMG_CUDA(h, cudaGetLastError());
    return DMG_OK;
}

// Enqueue K2 + K1 for a TDM batch whose inputs already sit on the device.
static int32_t tdm_enqueue_raw(dmg_handle_t h, int32_t B, const int32_t *d_seq, int32_t beam, int max_beam,
                               const int32_t *d_beam_user, int32_t topk, int32_t use_mask, const int64_t *d_cons_off,
                               const int32_t *d_cons, int32_t *d_items, float *d_logits, int32_t *d_counts,
                               BeamParams<float> *redo_out = nullptr, int probe_level = -1, WaveParams *probe_wp = nullptr,
                               int *probe_slot = nullptr)
{
    h->last_enqueue_wave = false;
    // redo_out: a caller that synchronises anyway takes the strict redo launch into its own hands (h_flags[1] tells it
    // whether the batch has redo users); redo_out->B == 0 on return when there is no such launch.
    if (redo_out) redo_out->B = 0;
    if (h->fast_dirty && h->arithmetic == DMG_ARITH_FAST && h->din.dtype == DMG_F32 && h->din.kind == 0)
        DMG_TRY(compute_fast_bounds(h));                         // bound tables; for an E = 16 / 32 model also its zero-padded E = 64 copy
    const DinDev &d = retrieval_din(h);
    const TreeDev &t = h->tree;
    const int T = d.T;
    DMG_TRY(ensure_dev(h, h->s_work, Carver::need({(size_t)B * T * 4, (size_t)B * T})));
    Carver cw(h->s_work.d);
    int32_t *d_codes = cw.take<int32_t>((size_t)B * T);
    uint8_t *d_mask = cw.take<uint8_t>((size_t)B * T);
    const int64_t n = (int64_t)B * T;
    const int cap = std::max(std::max(((2 * max_beam + 7) / 8) * 8, 8), ((topk + 7) / 8) * 8);
    bool use_fast = h->arithmetic == DMG_ARITH_FAST && d.dtype == DMG_F32 && d.E == 64 && d.T <= 15 && cap <= FastGeo::max_cap() &&
                    2 * (FastGeo::smem_bytes(cap) + 1024) <= (size_t)h->smem_per_sm;
    if (use_fast && h->fast_dirty) DMG_TRY(compute_fast_bounds(h));
    use_fast = use_fast && h->fast_ok;
    if (use_fast && h->redo_cap < B) {
        cudaFree(h->d_redo_list);
        h->d_redo_list = nullptr; h->redo_cap = 0;
        DMG_CUDA(h, cudaMalloc(&h->d_redo_list, ((size_t)B + 256) * sizeof(int32_t)));
        h->redo_cap = (int64_t)B + 256;
    }
    tdm_ids_to_codes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(
        d_seq, n, t.d_id_code, t.non_leaf_offset, t.max_code, d.rows, use_mask, d_codes, d_mask, h->d_flags,
        use_fast ? h->d_fast_ctl : nullptr);
    h->launches += 1;
    DMG_CUDA(h, cudaGetLastError());
    BeamParams<float> p;
    memset(&p, 0, sizeof(p));
    fill_scorer(d, p);
    p.E_run = d.E != h->din.E ? d.E : 0;
    p.B = B; p.hist = d_codes; p.hist_mask = d_mask; p.beam = beam; p.beam_user = d_beam_user;
    p.always_sort = 0; p.exists = t.complete ? nullptr : t.d_exists; p.leaf_level = t.max_level;
    p.mode = MODE_TDM_TOPK; p.topk = topk; p.leaf_item = t.d_leaf_item; p.cons_off = d_cons_off; p.cons = d_cons;
    p.out_items = d_items; p.out_scores = d_logits; p.out_counts = d_counts; p.out_stride = topk;
    p.cap = cap;
    p.capp = pow2_ge(p.cap);
    if (use_fast) {
        if (!h->d_fast_stats) {
            DMG_CUDA(h, cudaMalloc(&h->d_fast_stats, 64 * sizeof(unsigned long long)));
            DMG_CUDA(h, cudaMemsetAsync(h->d_fast_stats, 0, 64 * sizeof(unsigned long long), h->stream));
        }
        FastParams fx;
        memcpy(fx.b1, h->fast_host.data(), 64 * sizeof(float));
        memcpy(fx.w2, h->fast_host.data() + 64, 64 * sizeof(float));
        fx.b2 = h->fast_host[128];
        fx.mT = h->d_fast_tab; fx.w1 = d.w1<float>();
        fx.lvl_vx = h->d_fast_tab + 4160; fx.lvl_nx = h->d_fast_tab + 4192;
        fx.cA = h->fast_cA; fx.cZ = h->fast_cZ; fx.cH = h->fast_cH; fx.cGamma = h->fast_cGamma; fx.tau = h->fast_tau;
        fx.zvec = h->d_fast_tab + 4224;
        fx.redo_list = h->d_redo_list; fx.redo_count = h->d_fast_ctl + 1; fx.work_counter = h->d_fast_ctl;
        fx.host_flags = h->d_flags;
        fx.stats = h->d_fast_stats;
        fx.sparse_from = t.sparse_from;
        if (probe_level >= 0) {
            if (!h->wave_ok) return fail(h, DMG_ERR_UNSUPPORTED, "dmg_wave_probe: the level-synchronous path is not available for this model");
            return wave_enqueue(h, p, fx, max_beam, probe_level, probe_wp, probe_slot);
        }
        if (h->wave_ok) {
            DMG_TRY(wave_enqueue(h, p, fx, max_beam));
            h->last_enqueue_wave = true;
        } else {
        const size_t smem = FastGeo::smem_bytes(p.cap);
        DMG_CUDA(h, cudaFuncSetAttribute(beam_search_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int grid = std::min(B, 2 * h->sm_count);           // two co-resident CTAs per SM
        const int tail = B % grid;                               // last, partial round: one user per SM when it fits
        fx.tail_start = (grid == 2 * h->sm_count && tail > 0 && tail <= h->sm_count && !getenv("DMG_NO_TAIL")) ? B - tail : B;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (h->profiling) {
            DMG_CUDA(h, cudaEventCreate(&e0));
            DMG_CUDA(h, cudaEventCreate(&e1));
            DMG_CUDA(h, cudaEventRecord(e0, h->stream));
        }
        beam_search_fast_kernel<<<grid, FastGeo::THREADS, smem, h->stream>>>(p, fx);
        h->launches += 1;
        DMG_CUDA(h, cudaGetLastError());
        if (h->profiling) {
            DMG_CUDA(h, cudaEventRecord(e1, h->stream));
            h->prof_events.emplace_back(e0, e1);
        }
        }
        // users the fast kernel could not certify (exact ties at a cut, implausibly wide band): strict kernel
        p.user_list = h->d_redo_list; p.user_count = h->d_fast_ctl + 1;
        p.split = 4;                                             // a redo user runs on a cluster of 4 SMs: its <= 4 row tiles per level in parallel
        p.B = std::min(B, 32);                                   // cluster slots launched (the kernel strides over the list)
        if (redo_out) { *redo_out = p; return DMG_OK; }
        h->h_flags[1] = 0;                                       // nobody reads it on this path
        const bool prof = h->profiling;
        h->profiling = false;
        const int32_t rc = launch_beam<float>(h, p, d.E);
        h->profiling = prof;
        return rc;
    }
    return launch_beam<float>(h, p, d.E);
}

static void free_step_graphs(dmg_handle_t h)
{
    for (StepGraph &g : h->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    h->graphs.clear();
}

// K2 + K1 + K3 of a batch.  The level-synchronous path is ~31 kernel launches per batch; a host that drives several GPUs (or
// several handles) spends its time in cudaLaunchKernel, so a step whose arguments repeat -- the serving loop: same batch shape,
// same staging buffers -- is captured once (programmatic dependent launches become programmatic graph edges) and replayed with one
// cudaGraphLaunch.  The key holds every argument and every pointer baked into the kernels' parameters; anything else (first call,
// profiling, probes, the persistent / strict kernels, a failed capture) takes the plain launches.
static int32_t tdm_enqueue(dmg_handle_t h, int32_t B, const int32_t *d_seq, int32_t beam, int max_beam,
                           const int32_t *d_beam_user, int32_t topk, int32_t use_mask, const int64_t *d_cons_off,
                           const int32_t *d_cons, int32_t *d_items, float *d_logits, int32_t *d_counts,
                           BeamParams<float> *redo_out = nullptr, int probe_level = -1, WaveParams *probe_wp = nullptr,
                           int *probe_slot = nullptr)
{
    static const bool no_graph = getenv("DMG_NO_GRAPH") != nullptr;
    const bool eligible = !no_graph && probe_level < 0 && !h->profiling && h->arithmetic == DMG_ARITH_FAST && !h->fast_dirty && h->fast_ok &&
                          h->wave_ok;
    if (!eligible) {
        if (h->fast_dirty) free_step_graphs(h);                  // the model changed: every captured pointer may be stale
        return tdm_enqueue_raw(h, B, d_seq, beam, max_beam, d_beam_user, topk, use_mask, d_cons_off, d_cons, d_items, d_logits, d_counts,
                               redo_out, probe_level, probe_wp, probe_slot);
    }
    uint64_t tau_bits = 0;
    memcpy(&tau_bits, &h->fast_tau, sizeof(float));
    const uint64_t key[16] = {(uint64_t)B, (uint64_t)beam, (uint64_t)max_beam, (uint64_t)topk, (uint64_t)use_mask, (uint64_t)(uintptr_t)d_seq,
                              (uint64_t)(uintptr_t)d_beam_user, (uint64_t)(uintptr_t)d_cons_off, (uint64_t)(uintptr_t)d_cons,
                              (uint64_t)(uintptr_t)d_items, (uint64_t)(uintptr_t)d_logits, (uint64_t)(uintptr_t)d_counts,
                              (uint64_t)(uintptr_t)h->s_wave.d ^ ((uint64_t)(uintptr_t)h->s_work.d << 1), (uint64_t)(uintptr_t)h->din.d_params ^ tau_bits,
                              (uint64_t)(uintptr_t)h->tree.d_exists ^ ((uint64_t)(uintptr_t)h->d_redo_list << 1),
                              (uint64_t)(redo_out ? 1 : 0) | ((uint64_t)(uintptr_t)h->stream << 1)};
    StepGraph *g = nullptr;
    for (StepGraph &c : h->graphs) if (!memcmp(c.key, key, sizeof(key))) { g = &c; break; }
    if (!g) {
        if (h->graphs.size() >= 16) { if (h->graphs.front().exec) cudaGraphExecDestroy(h->graphs.front().exec); h->graphs.erase(h->graphs.begin()); }
        h->graphs.emplace_back();
        g = &h->graphs.back();
        memcpy(g->key, key, sizeof(key));
    }
    static_assert(sizeof(BeamParams<float>) <= sizeof(g->redo), "StepGraph::redo too small");
    if (g->exec) {
        DMG_CUDA(h, cudaGraphLaunch(g->exec, h->stream));
        h->launches += g->launches;
        if (redo_out) { if (g->has_redo) memcpy(redo_out, g->redo, sizeof(BeamParams<float>)); else redo_out->B = 0; }
        return DMG_OK;
    }
    g->seen++;
    if (g->bad || g->seen < 2) {
        const int32_t rc = tdm_enqueue_raw(h, B, d_seq, beam, max_beam, d_beam_user, topk, use_mask, d_cons_off, d_cons, d_items, d_logits, d_counts, redo_out);
        if (rc == DMG_OK && !h->last_enqueue_wave) g->bad = true;   // not the level-synchronous path: nothing to capture
        // tdm_enqueue_raw may have grown a scratch buffer: the key of the NEXT call will differ and start over
        return rc;
    }
    // second call with the same key: capture it
    const int64_t l0 = h->launches;
    BeamParams<float> redo;
    redo.B = 0;
    if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        g->bad = true;
        return tdm_enqueue_raw(h, B, d_seq, beam, max_beam, d_beam_user, topk, use_mask, d_cons_off, d_cons, d_items, d_logits, d_counts, redo_out);
    }
    const int32_t rc = tdm_enqueue_raw(h, B, d_seq, beam, max_beam, d_beam_user, topk, use_mask, d_cons_off, d_cons, d_items, d_logits, d_counts,
                                       redo_out ? &redo : nullptr);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
    cudaGraphExec_t exec = nullptr;
    if (rc != DMG_OK || ce != cudaSuccess || !graph || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
        cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        g->bad = true;
        h->launches = l0;
        return tdm_enqueue_raw(h, B, d_seq, beam, max_beam, d_beam_user, topk, use_mask, d_cons_off, d_cons, d_items, d_logits, d_counts, redo_out);
    }
    cudaGraphDestroy(graph);
    g->exec = exec;
    g->launches = h->launches - l0;
    if (getenv("DMG_GRAPH_DEBUG")) fprintf(stderr, "[dmg] retrieval step captured: B=%d beam=%d, %lld kernels in one graph\n", B, beam, (long long)g->launches);
    g->has_redo = redo_out != nullptr && redo.B != 0;
    if (g->has_redo) memcpy(g->redo, &redo, sizeof(redo));
    DMG_CUDA(h, cudaGraphLaunch(exec, h->stream));
    if (redo_out) { if (g->has_redo) *redo_out = redo; else redo_out->B = 0; }
    return DMG_OK;
}

DMG_API int32_t dmg_set_sync_mode(dmg_handle_t h, int32_t mode)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    if (mode != 0 && mode != 1) return fail(h, DMG_ERR_INVALID_ARG, "mode must be 0 (spin) or 1 (sleep)");
    h->sync_mode = mode;
    return DMG_OK;
}

DMG_API int32_t dmg_set_arithmetic(dmg_handle_t h, int32_t mode)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    if (mode != DMG_ARITH_STRICT && mode != DMG_ARITH_FAST) return fail(h, DMG_ERR_INVALID_ARG, "mode must be DMG_ARITH_STRICT or DMG_ARITH_FAST");
    h->arithmetic = mode;
    return DMG_OK;
}

DMG_API int32_t dmg_set_fast_tolerance(dmg_handle_t h, double tau)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    if (!(tau >= 0.0) || tau > 1.0) return fail(h, DMG_ERR_INVALID_ARG, "tau must be in [0, 1]");
    h->fast_tau = (float)tau;
    return DMG_OK;
}

DMG_API int32_t dmg_fast_stats(dmg_handle_t h, uint64_t *out6)
{
    if (!h || !out6) return DMG_ERR_INVALID_ARG;
    for (int i = 0; i < 7; i++) out6[i] = 0;
    if (!h->d_fast_stats) return DMG_OK;
    DMG_CUDA(h, cudaSetDevice(h->device));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    DMG_CUDA(h, cudaMemcpy(out6, h->d_fast_stats, 7 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
#ifdef DMG_FAST_TIMING
    {
        uint64_t t[24];
        DMG_CUDA(h, cudaMemcpy(t, h->d_fast_stats + 8, sizeof(t), cudaMemcpyDeviceToHost));
        static const char *names[] = {"prologue", "select", "rescore", "expand", "gather", "softmax", "epilogue", "final", "sched",
                                      "sel_warp", "mma_wait", "ph_wait", "fin_prep", "fin_strict", "sel_load", "sel_loop"};
        double tot = 0;
        for (int i = 0; i < TK_N; i++) tot += (double)t[i];
        uint64_t why[6];
        cudaMemcpy(why, h->d_fast_stats + 24, sizeof(why), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[fast timing] redo reasons: wide band %llu, tie in place %llu, unscored %llu, wide final %llu, tie at end %llu, proof failed %llu\n",
                (unsigned long long)why[0], (unsigned long long)why[1], (unsigned long long)why[2], (unsigned long long)why[3], (unsigned long long)why[4], (unsigned long long)why[5]);
        uint64_t iters = 0;
        cudaMemcpy(&iters, h->d_fast_stats + 7, 8, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[fast timing] select iterations per cut %.2f\n", (double)iters / (double)(out6[0] ? out6[0] : 1));
        for (int i = 0; i < TK_N; i++) fprintf(stderr, "[fast timing] %-9s %6.2f %%  %.3e cyc\n", names[i], 100.0 * t[i] / (tot > 0 ? tot : 1), (double)t[i]);
    }
#endif
#ifdef DMG_WAVE_TIMING
    {
        uint64_t t[12];
        DMG_CUDA(h, cudaMemcpy(t, h->d_fast_stats + 32, sizeof(t), cudaMemcpyDeviceToHost));
        static const char *names[] = {"loop top", "wait XFULL", "wait M1", "refill_x", "softmax", "group sync 1", "issue M2", "wait M2", "read Hacc", "group sync 2", "issue M1", "epilogue"};
        double tot = 0;
        for (int i = 0; i < 12; i++) tot += (double)t[i];
        for (int i = 0; i < 12; i++) fprintf(stderr, "[wave timing] %-13s %6.2f %%  %.3e cyc\n", names[i], 100.0 * t[i] / (tot > 0 ? tot : 1), (double)t[i]);
    }
#endif
    DMG_CUDA(h, cudaMemset(h->d_fast_stats, 0, 64 * sizeof(uint64_t)));
    return DMG_OK;
}

static int32_t tdm_precheck(dmg_handle_t h, int32_t B, int32_t beam, int32_t topk)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    if (!h->tree.loaded || !h->din.loaded) return fail(h, DMG_ERR_STATE, "tree and DIN weights must be loaded first");
    if (h->din.sharded) return fail(h, DMG_ERR_STATE, "the node table is sharded (dmg_shard_init): use the dmg_shard_* entry points");
    if (h->tree.complete || !h->tree.d_id_code) return fail(h, DMG_ERR_STATE, "dmg_tdm_retrieve needs a tree loaded with dmg_load_tree_tdm");
    if (h->din.dtype != DMG_F32) return fail(h, DMG_ERR_STATE, "TDM/JTM scorer is Module[Float]: load DMG_F32 weights");
    if (B <= 0 || beam <= 0 || topk <= 0) return fail(h, DMG_ERR_INVALID_ARG, "B, beam and topk must be positive");  // require(candidateNum > 0)
    if (h->din.rows < h->tree.n_codes)
        return fail(h, DMG_ERR_INVALID_ARG, "node table has %lld rows, tree needs %lld", (long long)h->din.rows, (long long)h->tree.n_codes);
    return DMG_OK;
}

// The *_dev entry points copy the caller's ids (B x T x 4 bytes, device to device) into the handle's own staging buffer first: the
// captured step (tdm_enqueue) then reads a pointer that does not change from call to call, whatever buffer the caller passes.
static int32_t stage_dev_seq(dmg_handle_t h, int32_t B, const int32_t **d_seq)
{
    const size_t bytes = (size_t)B * h->din.T * 4;
    DMG_TRY(ensure_dev(h, h->s_in, bytes));
    DMG_CUDA(h, cudaMemcpyAsync(h->s_in.d, *d_seq, bytes, cudaMemcpyDeviceToDevice, h->stream));
    *d_seq = (const int32_t *)h->s_in.d;
    return DMG_OK;
}

DMG_API int32_t dmg_tdm_retrieve_dev(dmg_handle_t h, int32_t B, const int32_t *d_item_seq, int32_t beam, int32_t topk,
                                     int32_t use_mask, int32_t *d_out_items, float *d_out_logits, int32_t *d_out_counts)
{
    DMG_TRY(tdm_precheck(h, B, beam, topk));
    if (!d_item_seq || !d_out_items || !d_out_logits || !d_out_counts) return fail(h, DMG_ERR_INVALID_ARG, "null device pointer");
    DMG_CUDA(h, cudaSetDevice(h->device));
    DMG_TRY(stage_dev_seq(h, B, &d_item_seq));
    return tdm_enqueue(h, B, d_item_seq, beam, beam, nullptr, topk, use_mask, nullptr, nullptr, d_out_items, d_out_logits, d_out_counts);
}

// The strict redo launch of a synchronous call: only when the fast kernel raised h_flags[1] (exact ties at a cut -- rare).
static int32_t tdm_redo_if_flagged(dmg_handle_t h, const BeamParams<float> &redo, bool *ran)
{
    *ran = false;
    if (redo.B <= 0 || !*(volatile int32_t *)(h->h_flags + 1)) return DMG_OK;
    h->h_flags[1] = 0;
    const bool prof = h->profiling;
    h->profiling = false;
    const int32_t rc = launch_beam<float>(h, redo, redo.E_run ? redo.E_run : h->din.E);
    h->profiling = prof;
    *ran = rc == DMG_OK;
    return rc;
}

DMG_API int32_t dmg_tdm_retrieve_dev_sync(dmg_handle_t h, int32_t B, const int32_t *d_item_seq, int32_t beam, int32_t topk,
                                          int32_t use_mask, int32_t *d_out_items, float *d_out_logits, int32_t *d_out_counts)
{
    DMG_TRY(tdm_precheck(h, B, beam, topk));
    if (!d_item_seq || !d_out_items || !d_out_logits || !d_out_counts) return fail(h, DMG_ERR_INVALID_ARG, "null device pointer");
    DMG_CUDA(h, cudaSetDevice(h->device));
    BeamParams<float> redo;
    DMG_TRY(stage_dev_seq(h, B, &d_item_seq));
    DMG_TRY(tdm_enqueue(h, B, d_item_seq, beam, beam, nullptr, topk, use_mask, nullptr, nullptr, d_out_items, d_out_logits, d_out_counts, &redo));
    DMG_TRY(check_flag(h, "dmg_tdm_retrieve_dev_sync"));
    bool ran = false;
    DMG_TRY(tdm_redo_if_flagged(h, redo, &ran));
    if (ran) DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    return DMG_OK;
}

// ---- DeepFM scorer: certified fast path (beam_wave_dfm.cuh) -----------------------------------------------------------------------------
// Bound tables of the loaded DeepFM model: vt (per-dimension weight of |x| in the hidden-unit and final-dot error terms) and, through
// level_bounds_kernel, the per-level maxima of vt . |x| and |x|_2; the transposed item half of W1 for the fast scorer.
// DeepFM with embed_size 16 / 32: the same zero-padding argument as for DIN (build_padded_model) -- features [x | 0], W1 with zero
// columns at the padded positions of every feature row; the FM sums and the Linear chains only gain exact zeros.
static __global__ void pad_dfm_dense_kernel(const float *__restrict__ src, int E, int F, float *__restrict__ dst)
{
    const int n = F * F * 64 + 2 * F + 1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float v;
        if (i < F * F * 64) {
            const int o = i / (F * 64), q = i - o * F * 64, sl = q >> 6, k = q & 63;
            v = k < E ? src[(size_t)o * F * E + sl * E + k] : 0.0f;
        } else {
            v = src[(size_t)F * F * E + (i - F * F * 64)];             // b1 | W2 | b2
        }
        dst[i] = v;
    }
}
static int32_t build_padded_dfm(dmg_handle_t h)
{
    const DinDev &d = h->din;
    DinDev &p = h->din_pad;
    const int F = d.T + 1;
    const int64_t n_params = d.rows * 64 + (int64_t)F * F * 64 + 2 * F + 1;
    if (p.loaded && (p.rows != d.rows || p.n_params != n_params)) free_din(p);
    if (!p.d_params && cudaMalloc(&p.d_params, (size_t)n_params * 4) != cudaSuccess) { cudaGetLastError(); p = DinDev(); return DMG_OK; }
    p.dtype = DMG_F32; p.esz = 4; p.rows = d.rows; p.E = 64; p.T = d.T; p.scale_E = d.E; p.n_params = n_params; p.kind = 1;
    pad_table_kernel<<<h->sm_count * 16, 256, 0, h->stream>>>(d.emb<float>(), d.rows, d.E, p.emb<float>());
    pad_dfm_dense_kernel<<<64, 256, 0, h->stream>>>(d.tail<float>(), d.E, F, p.tail<float>());
    h->launches += 2;
    DMG_CUDA(h, cudaGetLastError());
    p.loaded = true;
    return DMG_OK;
}

static int32_t compute_dfm_bounds(dmg_handle_t h)
{
    h->fast_ok = false;
    h->fast_dirty = false;
    if (h->parent) {
        // a clone reads its parent's tables (and zero-padded copy), brought up to date by dmg_clone
    } else if (h->din.kind == 1 && h->din.dtype == DMG_F32 && (h->din.E == 16 || h->din.E == 32) && h->din.T <= 10 && !getenv("DMG_NO_PAD")) {
        DMG_TRY(build_padded_dfm(h));
    } else if (h->din_pad.loaded) {
        free_din(h->din_pad);
    }
    DinDev &d = h->din_pad.loaded ? h->din_pad : h->din;
    if (d.kind != 1 || d.dtype != DMG_F32 || d.E != 64 || d.T > 10) return DMG_OK;   // T + 1 <= 11 hidden units in the fast scorer
    const int E = d.E, T = d.T, F = T + 1, IN = F * E;
    const size_t n_dense = (size_t)F * IN + 2 * F + 1;
    std::vector<float> w(n_dense);
    DMG_CUDA(h, cudaMemcpyAsync(w.data(), d.tail<float>(), n_dense * 4, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    const float *w1 = w.data(), *b1 = w1 + (size_t)F * IN, *w2 = b1 + F, *b2 = w2 + F;
    DfmConsts dc;
    memset(&dc, 0, sizeof(dc));
    dc.F = F; dc.b2 = b2[0];
    for (int o = 0; o < F; o++) { dc.w2[o] = w2[o]; dc.b1[o] = b1[o]; dc.aw2[o] = std::fabs(w2[o]); }
    static_assert(sizeof(DfmConsts) <= sizeof(h->dfm_consts), "dfm_consts too small");
    memcpy(h->dfm_consts, &dc, sizeof(dc));
    h->fast_host.assign(64 * 11, 0.0f);                          // W1x^T [k][o] for the fast scorer's parameter block
    for (int k = 0; k < E; k++)
        for (int o = 0; o < F && o < 11; o++) h->fast_host[(size_t)k * 11 + o] = w1[(size_t)o * IN + k];
    std::vector<float> tab(4288, 0.0f);                          // [4096, 4160) vt | lvl_vx | lvl_nx
    const double u = std::ldexp(1.0, -24);
    bool finite = true;
    for (int k = 0; k < E; k++) {
        double vt = 0.0;
        for (int o = 0; o < F; o++) {
            vt += std::fabs((double)w2[o]) * std::fabs((double)w1[(size_t)o * IN + k]) * ((705.0 + 66.0 + 2.0) * u + 2.0 * (F + 3.0) * u);
        }
        tab[4096 + k] = (float)(vt * (1.0 + 1e-6));
        finite = finite && std::isfinite(vt);
    }
    if (!h->d_fast_tab) DMG_CUDA(h, cudaMalloc(&h->d_fast_tab, tab.size() * sizeof(float)));
    if (!h->d_fast_ctl) {
        DMG_CUDA(h, cudaMalloc(&h->d_fast_ctl, DMG_FAST_CTL_WORDS * sizeof(int32_t)));
        DMG_CUDA(h, cudaMemsetAsync(h->d_fast_ctl, 0, DMG_FAST_CTL_WORDS * sizeof(int32_t), h->stream));
    }
    DMG_CUDA(h, cudaMemcpyAsync(h->d_fast_tab, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    level_bounds_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(d.emb<float>(), d.rows, h->d_fast_tab + 4096, h->d_fast_tab + 4160, h->d_fast_tab + 4192);
    h->launches += 1;
    DMG_CUDA(h, cudaGetLastError());
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    h->fast_ok = finite;
    return DMG_OK;
}

int32_t dmg_deepfm_tdm_retrieve(dmg_handle_t h, int32_t B, const int32_t *item_seq, int32_t beam, int32_t topk, const int64_t *cons_off,
                                const int32_t *cons, int32_t widen_beam, int32_t *out_items, float *out_logits, int32_t *out_counts);   // shard.cu

// dmg_tdm_retrieve with a DeepFM model in FAST arithmetic: the level-synchronous chain of beam_wave.cuh with the DeepFM prologue /
// scorer / strict re-scores; users the fast path hands back (exact ties across a decision point, failed proofs) re-run on the strict
// level-synchronous path (shard.cu).  Returns DMG_ERR_UNSUPPORTED-free: `*handled` = false when the shape is outside the fast path.
static int32_t dfm_fast_retrieve(dmg_handle_t h, int32_t B, const int32_t *item_seq, int32_t beam, int32_t topk, const int64_t *consumed_off,
                                 const int32_t *consumed_items, int32_t widen_beam, int32_t *out_items, float *out_logits, int32_t *out_counts,
                                 bool *handled)
{
    using WG = WaveGeo;
    *handled = false;
    const TreeDev &t = h->tree;
    if (h->arithmetic != DMG_ARITH_FAST || (h->din.E != 64 && h->din.E != 32 && h->din.E != 16) || h->din.T > 10 || dmg_shard_world(h) > 1 || !t.loaded ||
        t.complete || !t.d_id_code || getenv("DMG_DFM_STRICT"))
        return DMG_OK;
    if (B <= 0 || beam <= 0 || topk <= 0) return DMG_OK;          // the strict path reports the argument errors
    DMG_CUDA(h, cudaSetDevice(h->device));
    if (h->fast_dirty) DMG_TRY(compute_dfm_bounds(h));
    if (!h->fast_ok) return DMG_OK;
    DinDev &d = h->din_pad.loaded ? h->din_pad : h->din;          // the zero-padded E = 64 copy of an E = 16 / 32 model
    const int T = d.T;
    const int64_t n_cons = consumed_off ? consumed_off[B] : 0;
    const bool per_user = consumed_off && widen_beam;
    int max_beam = beam;
    if (per_user)
        for (int u = 0; u < B; u++) max_beam = std::max(max_beam, (int)((consumed_off[u + 1] - consumed_off[u] + topk) / 2));
    const int cap = std::max(std::max(((2 * max_beam + 7) / 8) * 8, 8), ((topk + 7) / 8) * 8);
    if (cap > 512 || topk > FastGeo::MAX_FINAL / 2) return DMG_OK;
    DfmConsts dc;
    memcpy(&dc, h->dfm_consts, sizeof(dc));
    // staging: [seq | cons_off | cons | beam_user]
    const size_t in_bytes = Carver::need({(size_t)B * T * 4, consumed_off ? (size_t)(B + 1) * 8 : 0, (size_t)n_cons * 4, per_user ? (size_t)B * 4 : 0});
    DMG_TRY(ensure_host(h, h->s_in, in_bytes));
    DMG_TRY(ensure_dev(h, h->s_in, in_bytes));
    Carver ch(h->s_in.h), cdv(h->s_in.d);
    int32_t *hs = ch.take<int32_t>((size_t)B * T), *ds = cdv.take<int32_t>((size_t)B * T);
    int64_t *ho = ch.take<int64_t>(consumed_off ? B + 1 : 0), *dof = cdv.take<int64_t>(consumed_off ? B + 1 : 0);
    int32_t *hc = ch.take<int32_t>((size_t)n_cons), *dcs = cdv.take<int32_t>((size_t)n_cons);
    int32_t *hb = ch.take<int32_t>(per_user ? B : 0), *db = cdv.take<int32_t>(per_user ? B : 0);
    memcpy(hs, item_seq, (size_t)B * T * 4);
    if (consumed_off) {
        memcpy(ho, consumed_off, (size_t)(B + 1) * 8);
        if (n_cons) memcpy(hc, consumed_items, (size_t)n_cons * 4);
        if (per_user)
            for (int u = 0; u < B; u++) hb[u] = std::max((int)((consumed_off[u + 1] - consumed_off[u] + topk) / 2), beam);   // Recommender.scala:28-31
    }
    DMG_CUDA(h, cudaMemcpyAsync(h->s_in.d, h->s_in.h, ch.off, cudaMemcpyHostToDevice, h->stream));
    const size_t out_bytes = Carver::need({(size_t)B * topk * 4, (size_t)B * topk * 4, (size_t)B * 4, (size_t)B * 4, 64});
    DMG_TRY(ensure_host(h, h->s_out, out_bytes));
    DMG_TRY(ensure_dev(h, h->s_out, out_bytes));
    Carver oh(h->s_out.h), od(h->s_out.d);
    int32_t *h_items = oh.take<int32_t>((size_t)B * topk), *d_items = od.take<int32_t>((size_t)B * topk);
    float *h_log = oh.take<float>((size_t)B * topk), *d_log = od.take<float>((size_t)B * topk);
    int32_t *h_cnt = oh.take<int32_t>(B), *d_cnt = od.take<int32_t>(B);
    int32_t *h_redo = oh.take<int32_t>(B), *d_redo = od.take<int32_t>(B);
    int32_t *h_nredo = oh.take<int32_t>(1), *d_nredo = od.take<int32_t>(1);
    (void)d_nredo;
    // work: codes + mask of the histories, then the wave state
    DMG_TRY(ensure_dev(h, h->s_work, Carver::need({(size_t)B * T * 4, (size_t)B * T})));
    Carver cw(h->s_work.d);
    int32_t *d_codes = cw.take<int32_t>((size_t)B * T);
    uint8_t *d_mask = cw.take<uint8_t>((size_t)B * T);
    DMG_TRY(ensure_dev(h, h->s_wave, Carver::need({(size_t)B * cap * 4, (size_t)B * cap * 4, (size_t)B * cap * 4, (size_t)B * 4,
                                                   (size_t)B * sizeof(WaveUser), (size_t)B * WG::UOP_BYTES, (size_t)B * WG::VCAP * 4,
                                                   (size_t)B * WG::VCAP * 4, (size_t)B * WG::VCAP * 4, (size_t)B * 32 * 4,
                                                   (size_t)B * ((cap + 127) / 128) * 4, (size_t)B * WaveFinal::RCAP * 4, (size_t)B * WaveFinal::RCAP * 4,
                                                   (size_t)B * FastGeo::MAX_FINAL * 4, (size_t)B * 4 * 4, (size_t)B * 27 * 4})));
    if (!h->d_fast_stats) {
        DMG_CUDA(h, cudaMalloc(&h->d_fast_stats, 64 * sizeof(unsigned long long)));
        DMG_CUDA(h, cudaMemsetAsync(h->d_fast_stats, 0, 64 * sizeof(unsigned long long), h->stream));
    }
    Carver c(h->s_wave.d);
    WaveParams wp;
    memset(&wp, 0, sizeof(wp));
    wp.kind = 1; wp.dfm_dense = d.tail<float>();
    wp.B = B; wp.T = T; wp.cap = cap; wp.beam = beam; wp.beam_user = per_user ? db : nullptr;
    wp.emb = d.emb<float>(); wp.hist = d_codes; wp.hist_mask = d_mask; wp.exists = t.d_exists;
    wp.leaf_level = t.max_level; wp.sparse_from = t.sparse_from; wp.scale = 1.0f;
    wp.code[0] = c.take<int32_t>((size_t)B * cap); wp.code[1] = c.take<int32_t>((size_t)B * cap);
    wp.score = c.take<float>((size_t)B * cap);
    wp.count = c.take<int32_t>(B);
    wp.user = c.take<WaveUser>(B);
    wp.uop = c.take<unsigned char>((size_t)B * WG::UOP_BYTES);
    wp.v_code = c.take<int32_t>((size_t)B * WG::VCAP); wp.v_fast = c.take<float>((size_t)B * WG::VCAP);
    wp.v_meta = c.take<uint32_t>((size_t)B * WG::VCAP); wp.v_segeps = c.take<float>((size_t)B * 32);
    wp.tile_list = c.take<int32_t>((size_t)B * ((cap + 127) / 128));
    WaveFinal wf;
    wf.row_code = c.take<int32_t>((size_t)B * WaveFinal::RCAP); wf.row_strict = c.take<float>((size_t)B * WaveFinal::RCAP);
    wf.fin_pos = c.take<int32_t>((size_t)B * FastGeo::MAX_FINAL); wf.meta = c.take<int32_t>((size_t)B * 4);
    wf.chunk_list = c.take<int32_t>((size_t)B * 27);
    wf.chunk_count = h->d_fast_ctl + 8 + 40;
    wp.tile_count = h->d_fast_ctl + 8;
    wp.lvl_vx = h->d_fast_tab + 4160; wp.lvl_nx = h->d_fast_tab + 4192;
    wp.tau = h->fast_tau;
    wp.stats = h->d_fast_stats; wp.redo_list = d_redo; wp.redo_count = h->d_fast_ctl + 1; wp.host_flags = h->d_flags;
    BeamParams<float> bp;
    memset(&bp, 0, sizeof(bp));
    bp.B = B; bp.topk = topk; bp.leaf_item = t.d_leaf_item; bp.cons_off = consumed_off ? dof : nullptr; bp.cons = consumed_off ? dcs : nullptr;
    bp.out_items = d_items; bp.out_scores = d_log; bp.out_counts = d_cnt; bp.out_stride = topk; bp.leaf_level = t.max_level; bp.cap = cap;
    WaveStrictW sw;
    memset(&sw, 0, sizeof(sw));
    const int64_t n = (int64_t)B * T;
    tdm_ids_to_codes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(ds, n, t.d_id_code, t.non_leaf_offset, t.max_code, d.rows, 0, d_codes,
                                                                              d_mask, h->d_flags, h->d_fast_ctl);
    h->launches += 1;
    const bool pdl = !getenv("DMG_WAVE_NO_PDL");
    DMG_CUDA(h, launch_chain(wave_dfm_prologue_kernel, B, 128, 0, h->stream, false, wp, dc, (const float *)d.tail<float>()));
    auto select_kernel = cap <= 256 ? wave_select_kernel<8, 1> : (cap <= 416 ? wave_select_kernel<13, 1> : wave_select_kernel<16, 1>);
    const size_t score_smem = (size_t)(128 * 68 + 96) * 4;
    const int ntiles = B * ((cap + 127) / 128), grid = std::min(ntiles, 6 * h->sm_count);
    DfmW1x wx;
    memcpy(wx.w, h->fast_host.data(), sizeof(wx.w));
    const size_t strict_smem = ((size_t)(T + 1) * (T + 1) * 64 + 2 * (T + 1) + 1 + 4 + 128 * (68 + 20)) * 4;
    DMG_CUDA(h, cudaFuncSetAttribute(wave_dfm_strict_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)strict_smem));
    int slot = 0;
    const int s_min = lower_log2(beam);
    const size_t sel_smem = ((size_t)(T + 1) * (T + 1) * 64 + 2 * (T + 1) + 1 + 4) * 4;     // the dense weights for parked cuts
    DMG_CUDA(h, cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem));
    for (int level = s_min; level < t.max_level; level++) {
        DMG_CUDA(h, launch_chain(select_kernel, (B + 3) / 4, 128, sel_smem, h->stream, pdl, wp, sw, level, slot, dc));
        DMG_CUDA(h, launch_chain(wave_dfm_score_kernel, grid, 128, score_smem, h->stream, pdl, wp, dc, wx, slot ^ 1, level + 1));
        slot ^= 1;
        h->launches += 2;
    }
    auto prep_kernel = cap <= 256 ? wave_final_prep_kernel<8> : (cap <= 416 ? wave_final_prep_kernel<13> : wave_final_prep_kernel<16>);
    DMG_CUDA(h, launch_chain(prep_kernel, (B + 3) / 4, 128, 0, h->stream, pdl, wp, bp, wf, slot));
    DMG_CUDA(h, launch_chain(wave_dfm_strict_rows_kernel, std::min(B, 2 * h->sm_count), 256, strict_smem, h->stream, pdl, wp, dc, (const float *)d.tail<float>(), wf));
    DMG_CUDA(h, launch_chain(wave_final_verify_kernel, (B + 3) / 4, 128, 0, h->stream, pdl, wp, bp, wf, slot));
    h->launches += 4;
    DMG_CUDA(h, cudaGetLastError());
    DMG_CUDA(h, cudaMemcpyAsync(h->s_out.h, h->s_out.d, od.off, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaMemcpyAsync(h_nredo, h->d_fast_ctl + 1, 4, cudaMemcpyDeviceToHost, h->stream));
    DMG_TRY(check_flag(h, "dmg_tdm_retrieve"));
    h->h_flags[1] = 0;
    memcpy(out_items, h_items, (size_t)B * topk * 4);
    memcpy(out_logits, h_log, (size_t)B * topk * 4);
    memcpy(out_counts, h_cnt, (size_t)B * 4);
    const int n_redo = *h_nredo;
    if (n_redo > 0) {                                            // the strict level-synchronous path for the users handed back
        std::vector<int32_t> rs((size_t)n_redo * T), ri((size_t)n_redo * topk), rcn(n_redo), rc_flat;
        std::vector<float> rl((size_t)n_redo * topk);
        std::vector<int64_t> ro;
        if (consumed_off) ro.push_back(0);
        for (int q = 0; q < n_redo; q++) {
            const int u = h_redo[q];
            memcpy(rs.data() + (size_t)q * T, item_seq + (size_t)u * T, (size_t)T * 4);
            if (consumed_off) {
                rc_flat.insert(rc_flat.end(), consumed_items + consumed_off[u], consumed_items + consumed_off[u + 1]);
                ro.push_back((int64_t)rc_flat.size());
            }
        }
        DMG_TRY(dmg_deepfm_tdm_retrieve(h, n_redo, rs.data(), beam, topk, consumed_off ? ro.data() : nullptr, consumed_off ? rc_flat.data() : nullptr,
                                        widen_beam, ri.data(), rl.data(), rcn.data()));
        for (int q = 0; q < n_redo; q++) {
            const int u = h_redo[q];
            memcpy(out_items + (size_t)u * topk, ri.data() + (size_t)q * topk, (size_t)topk * 4);
            memcpy(out_logits + (size_t)u * topk, rl.data() + (size_t)q * topk, (size_t)topk * 4);
            out_counts[u] = rcn[q];
        }
    }
    *handled = true;
    return DMG_OK;
}

DMG_API int32_t dmg_tdm_retrieve(dmg_handle_t h, int32_t B, const int32_t *item_seq, int32_t beam, int32_t topk,
                                 int32_t use_mask, const int64_t *consumed_off, const int32_t *consumed_items,
                                 int32_t widen_beam, int32_t *out_items, float *out_logits, int32_t *out_counts)
{
    if (h && h->din.loaded && h->din.kind == 1 && h->din.dtype == DMG_F32) {   // DeepFM scorer: level-synchronous path (shard.cu)
        if (!item_seq || !out_items || !out_logits || !out_counts) return fail(h, DMG_ERR_INVALID_ARG, "null host pointer");
        if (consumed_off && !consumed_items && B > 0 && consumed_off[B] > 0) return fail(h, DMG_ERR_INVALID_ARG, "consumed_items is null");
        bool handled = false;
        DMG_TRY(dfm_fast_retrieve(h, B, item_seq, beam, topk, consumed_off, consumed_items, widen_beam, out_items, out_logits, out_counts, &handled));
        if (handled) return DMG_OK;
        return dmg_deepfm_tdm_retrieve(h, B, item_seq, beam, topk, consumed_off, consumed_items, widen_beam, out_items, out_logits, out_counts);
    }
    DMG_TRY(tdm_precheck(h, B, beam, topk));
    if (!item_seq || !out_items || !out_logits || !out_counts) return fail(h, DMG_ERR_INVALID_ARG, "null host pointer");
    if (consumed_off && !consumed_items && consumed_off[B] > 0) return fail(h, DMG_ERR_INVALID_ARG, "consumed_items is null");
    DMG_CUDA(h, cudaSetDevice(h->device));
    const int T = h->din.T;
    const int64_t n_cons = consumed_off ? consumed_off[B] : 0;
    const bool per_user = consumed_off && widen_beam;
    // pinned staging: [seq | cons_off | cons | beam_user]
    const size_t b_seq = (size_t)B * T * 4, b_off = consumed_off ? (size_t)(B + 1) * 8 : 0, b_cons = (size_t)n_cons * 4,
                 b_beam = per_user ? (size_t)B * 4 : 0;
    const size_t in_bytes = Carver::need({b_seq, b_off, b_cons, b_beam});
    DMG_TRY(ensure_host(h, h->s_in, in_bytes));
    DMG_TRY(ensure_dev(h, h->s_in, in_bytes));
    Carver ch(h->s_in.h), cd(h->s_in.d);
    int32_t *hs = ch.take<int32_t>((size_t)B * T), *ds = cd.take<int32_t>((size_t)B * T);
    int64_t *ho = ch.take<int64_t>(consumed_off ? B + 1 : 0), *dof = cd.take<int64_t>(consumed_off ? B + 1 : 0);
    int32_t *hc = ch.take<int32_t>((size_t)n_cons), *dc = cd.take<int32_t>((size_t)n_cons);
    int32_t *hb = ch.take<int32_t>(per_user ? B : 0), *db = cd.take<int32_t>(per_user ? B : 0);
    memcpy(hs, item_seq, b_seq);
    int max_beam = beam;
    if (consumed_off) {
        memcpy(ho, consumed_off, b_off);
        if (n_cons) memcpy(hc, consumed_items, b_cons);
        if (per_user)
            for (int u = 0; u < B; u++) {                                   // Recommender.scala:28-31
                const int w = (int)((consumed_off[u + 1] - consumed_off[u] + topk) / 2);
                hb[u] = std::max(w, beam);
                max_beam = std::max(max_beam, hb[u]);
            }
    }
    DMG_CUDA(h, cudaMemcpyAsync(h->s_in.d, h->s_in.h, ch.off, cudaMemcpyHostToDevice, h->stream));
    const size_t out_bytes = Carver::need({(size_t)B * topk * 4, (size_t)B * topk * 4, (size_t)B * 4});
    DMG_TRY(ensure_host(h, h->s_out, out_bytes));
    DMG_TRY(ensure_dev(h, h->s_out, out_bytes));
    Carver oh(h->s_out.h), od(h->s_out.d);
    int32_t *h_items = oh.take<int32_t>((size_t)B * topk), *d_items = od.take<int32_t>((size_t)B * topk);
    float *h_log = oh.take<float>((size_t)B * topk), *d_log = od.take<float>((size_t)B * topk);
    int32_t *h_cnt = oh.take<int32_t>(B), *d_cnt = od.take<int32_t>(B);
    BeamParams<float> redo;
    DMG_TRY(tdm_enqueue(h, B, ds, beam, max_beam, per_user ? db : nullptr, topk, use_mask, consumed_off ? dof : nullptr,
                        consumed_off ? dc : nullptr, d_items, d_log, d_cnt, &redo));
    DMG_CUDA(h, cudaMemcpyAsync(h->s_out.h, h->s_out.d, od.off, cudaMemcpyDeviceToHost, h->stream));
    DMG_TRY(check_flag(h, "dmg_tdm_retrieve"));
    bool ran = false;
    DMG_TRY(tdm_redo_if_flagged(h, redo, &ran));
    if (ran) {                                                   // results again, now with the redo users
        DMG_CUDA(h, cudaMemcpyAsync(h->s_out.h, h->s_out.d, od.off, cudaMemcpyDeviceToHost, h->stream));
        DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    memcpy(out_items, h_items, (size_t)B * topk * 4);
    memcpy(out_logits, h_log, (size_t)B * topk * 4);
    memcpy(out_counts, h_cnt, (size_t)B * 4);
    return DMG_OK;
}

// Diagnostic: run the level-synchronous tensor-core search until the candidates of tree level `level` are scored and
// return them with their FAST scores and the bound eps on |fast - strict| (tests compare with model.forward).
DMG_API int32_t dmg_wave_probe(dmg_handle_t h, int32_t B, const int32_t *item_seq, int32_t beam, int32_t use_mask, int32_t level,
                               int32_t cap, int32_t *out_codes, float *out_scores, int32_t *out_counts, float *out_eps)
{
    DMG_TRY(tdm_precheck(h, B, beam, 1));
    if (!item_seq || !out_codes || !out_scores || !out_counts || !out_eps || level < 0) return fail(h, DMG_ERR_INVALID_ARG, "bad arguments");
    if (h->arithmetic != DMG_ARITH_FAST) return fail(h, DMG_ERR_STATE, "dmg_wave_probe needs DMG_ARITH_FAST");
    DMG_CUDA(h, cudaSetDevice(h->device));
    const int T = h->din.T;
    const size_t b_seq = (size_t)B * T * 4;
    DMG_TRY(ensure_host(h, h->s_in, b_seq));
    DMG_TRY(ensure_dev(h, h->s_in, b_seq));
    memcpy(h->s_in.h, item_seq, b_seq);
    DMG_CUDA(h, cudaMemcpyAsync(h->s_in.d, h->s_in.h, b_seq, cudaMemcpyHostToDevice, h->stream));
    const size_t out_bytes = Carver::need({(size_t)B * 4, (size_t)B * 4, (size_t)B * 4});
    DMG_TRY(ensure_dev(h, h->s_out, out_bytes));
    Carver od(h->s_out.d);
    int32_t *d_items = od.take<int32_t>(B);
    float *d_log = od.take<float>(B);
    int32_t *d_cnt = od.take<int32_t>(B);
    WaveParams wp;
    int slot = 0;
    DMG_TRY(tdm_enqueue(h, B, (const int32_t *)h->s_in.d, beam, beam, nullptr, 1, use_mask, nullptr, nullptr, d_items, d_log, d_cnt, nullptr,
                        level, &wp, &slot));
    if (wp.cap != cap) return fail(h, DMG_ERR_INVALID_ARG, "dmg_wave_probe: cap must be %d for this beam", wp.cap);
    std::vector<WaveUser> us(B);
    DMG_CUDA(h, cudaMemcpyAsync(out_codes, wp.code[slot], (size_t)B * cap * 4, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaMemcpyAsync(out_scores, wp.score, (size_t)B * cap * 4, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaMemcpyAsync(out_counts, wp.count, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaMemcpyAsync(us.data(), wp.user, (size_t)B * sizeof(WaveUser), cudaMemcpyDeviceToHost, h->stream));
    DMG_TRY(check_flag(h, "dmg_wave_probe"));
    for (int u = 0; u < B; u++) out_eps[u] = (us[u].flags & WU_REDO) ? -1.0f : us[u].eps;
    return DMG_OK;
}

// OTM: batchBeamSearch (dump) and recommend (topk) share one path.
static int32_t otm_run(dmg_handle_t h, int32_t B, const int32_t *leaf_seq, int32_t beam, int32_t use_mask, int mode,
                       int32_t topk, int32_t *out_ids, double *out_scores, int32_t *out_counts,
                       int32_t *lvl_ids = nullptr, double *lvl_scores = nullptr, int32_t *lvl_counts = nullptr)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    if (!h->tree.loaded || !h->din.loaded) return fail(h, DMG_ERR_STATE, "tree and DIN weights must be loaded first");
    const bool deepfm64 = h->din.kind == 1 && h->din.dtype == DMG_F64;      // DeepModel[Double] = DeepFM: level-synchronous path (otm_deepfm.cu)
    if (h->din.sharded && !deepfm64) return fail(h, DMG_ERR_STATE, "the node table is sharded (dmg_shard_init): use the dmg_shard_* entry points");
    if (!h->tree.complete) return fail(h, DMG_ERR_STATE, "OTM needs a complete tree (dmg_load_tree_complete)");
    if (h->din.dtype != DMG_F64) return fail(h, DMG_ERR_STATE, "OTM scorer is DeepModel[Double]: load DMG_F64 weights");
    if (B <= 0 || beam <= 0 || (mode == MODE_OTM_TOPK && topk <= 0) || !leaf_seq || !out_ids || !out_scores || !out_counts)
        return fail(h, DMG_ERR_INVALID_ARG, "bad arguments");
    if (h->din.rows < h->tree.n_codes)
        return fail(h, DMG_ERR_INVALID_ARG, "node table has %lld rows, tree needs %lld", (long long)h->din.rows, (long long)h->tree.n_codes);
    if (deepfm64)                                                 // no mask input (otm/.../model/DeepFM.scala:17-18)
        return dmg_deepfm64_otm_run(h, B, leaf_seq, beam, mode == MODE_OTM_TOPK, topk, out_ids, out_scores, out_counts, lvl_ids, lvl_scores, lvl_counts);
    DMG_CUDA(h, cudaSetDevice(h->device));
    const DinDev &d = h->din;
    const TreeDev &t = h->tree;
    const int T = d.T;
    const int s = lower_log2(beam);
    const int width = 2 * std::max(beam, 1 << s);                      // entries per user in dump mode
    const int stride = mode == MODE_OTM_DUMP ? width : topk;
    const size_t b_seq = (size_t)B * T * 4;
    DMG_TRY(ensure_host(h, h->s_in, b_seq));
    DMG_TRY(ensure_dev(h, h->s_in, b_seq));
    memcpy(h->s_in.h, leaf_seq, b_seq);
    DMG_CUDA(h, cudaMemcpyAsync(h->s_in.d, h->s_in.h, b_seq, cudaMemcpyHostToDevice, h->stream));
    DMG_TRY(ensure_dev(h, h->s_work, Carver::need({(size_t)B * T * 4, (size_t)B * T})));
    Carver cw(h->s_work.d);
    int32_t *d_codes = cw.take<int32_t>((size_t)B * T);
    uint8_t *d_mask = cw.take<uint8_t>((size_t)B * T);
    const int64_t n = (int64_t)B * T;
    seq_to_codes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>((const int32_t *)h->s_in.d, n, d.rows, use_mask,
                                                                            d_codes, d_mask, h->d_flags);
    h->launches += 1;
    DMG_CUDA(h, cudaGetLastError());
    const size_t out_bytes = Carver::need({(size_t)B * stride * 4, (size_t)B * stride * 8, (size_t)B * 4});
    DMG_TRY(ensure_host(h, h->s_out, out_bytes));
    DMG_TRY(ensure_dev(h, h->s_out, out_bytes));
    Carver oh(h->s_out.h), od(h->s_out.d);
    int32_t *h_ids = oh.take<int32_t>((size_t)B * stride), *d_ids = od.take<int32_t>((size_t)B * stride);
    double *h_sc = oh.take<double>((size_t)B * stride), *d_sc = od.take<double>((size_t)B * stride);
    int32_t *h_cnt = oh.take<int32_t>(B), *d_cnt = od.take<int32_t>(B);
    BeamParams<double> p;
    memset(&p, 0, sizeof(p));
    fill_scorer(d, p);
    p.B = B; p.hist = d_codes; p.hist_mask = d_mask; p.beam = beam; p.beam_user = nullptr;
    p.always_sort = 1; p.exists = nullptr; p.leaf_level = t.max_level;
    p.mode = mode; p.topk = topk; p.leaf_item = t.d_leaf_item;
    p.out_items = d_ids; p.out_scores = d_sc; p.out_counts = d_cnt; p.out_stride = stride;
    p.cap = std::max(((width + 7) / 8) * 8, 8);
    p.cap = std::max(p.cap, ((std::max(topk, 1) + 7) / 8) * 8);
    p.capp = pow2_ge(p.cap);
    const int n_lvl = std::max(t.max_level - s, 0);
    int32_t *d_li = nullptr, *d_lc = nullptr;
    double *d_ls = nullptr;
    if (lvl_ids && n_lvl > 0) {
        const size_t lb = Carver::need({(size_t)B * n_lvl * width * 4, (size_t)B * n_lvl * width * 8, (size_t)B * n_lvl * 4});
        void *blk = nullptr;
        DMG_CUDA(h, cudaMalloc(&blk, lb));
        Carver c2(blk);
        d_li = c2.take<int32_t>((size_t)B * n_lvl * width);
        d_ls = c2.take<double>((size_t)B * n_lvl * width);
        d_lc = c2.take<int32_t>((size_t)B * n_lvl);
        p.lvl_items = d_li; p.lvl_scores = d_ls; p.lvl_counts = d_lc; p.lvl_stride = width; p.n_lvl = n_lvl;
    }
    int32_t rc_launch = launch_beam<double>(h, p, d.E);
    if (rc_launch != DMG_OK) { if (d_li) cudaFree(d_li); return rc_launch; }
    DMG_CUDA(h, cudaMemcpyAsync(h->s_out.h, h->s_out.d, od.off, cudaMemcpyDeviceToHost, h->stream));
    if (d_li) {
        cudaMemcpyAsync(lvl_ids, d_li, (size_t)B * n_lvl * width * 4, cudaMemcpyDeviceToHost, h->stream);
        cudaMemcpyAsync(lvl_scores, d_ls, (size_t)B * n_lvl * width * 8, cudaMemcpyDeviceToHost, h->stream);
        cudaMemcpyAsync(lvl_counts, d_lc, (size_t)B * n_lvl * 4, cudaMemcpyDeviceToHost, h->stream);
        cudaStreamSynchronize(h->stream);
        cudaFree(d_li);
    }
    DMG_TRY(check_flag(h, "dmg_otm"));
    memcpy(out_ids, h_ids, (size_t)B * stride * 4);
    memcpy(out_scores, h_sc, (size_t)B * stride * 8);
    memcpy(out_counts, h_cnt, (size_t)B * 4);
    return DMG_OK;
}

DMG_API int32_t dmg_otm_beam_search(dmg_handle_t h, int32_t B, const int32_t *leaf_seq, int32_t beam, int32_t use_mask,
                                    int32_t *out_ids, double *out_scores, int32_t *out_counts)
{
    return otm_run(h, B, leaf_seq, beam, use_mask, MODE_OTM_DUMP, 0, out_ids, out_scores, out_counts);
}

DMG_API int32_t dmg_otm_beam_search_levels(dmg_handle_t h, int32_t B, const int32_t *leaf_seq, int32_t beam, int32_t use_mask,
                                           int32_t *out_ids, double *out_scores, int32_t *out_counts)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    if (!out_ids || !out_scores || !out_counts) return fail(h, DMG_ERR_INVALID_ARG, "null output");
    if (!h->tree.loaded || beam <= 0 || B <= 0) return fail(h, DMG_ERR_INVALID_ARG, "bad arguments");
    const int s = lower_log2(beam);
    const int width = 2 * std::max(beam, 1 << s);
    std::vector<int32_t> last_ids((size_t)B * width), last_cnt(B);
    std::vector<double> last_sc((size_t)B * width);
    return otm_run(h, B, leaf_seq, beam, use_mask, MODE_OTM_DUMP, 0, last_ids.data(), last_sc.data(), last_cnt.data(),
                   out_ids, out_scores, out_counts);
}

DMG_API int32_t dmg_otm_retrieve(dmg_handle_t h, int32_t B, const int32_t *leaf_seq, int32_t beam, int32_t topk,
                                 int32_t use_mask, int32_t *out_items, double *out_scores, int32_t *out_counts)
{
    return otm_run(h, B, leaf_seq, beam, use_mask, MODE_OTM_TOPK, topk, out_items, out_scores, out_counts);
}

// model.forward on n rows
// model.forward on n device-resident rows (DIN): logits of the loaded dtype into d_out
static int32_t score_pairs_enqueue(dmg_handle_t h, int64_t n, const int32_t *dn, const int32_t *ds, const uint8_t *d_mask, void *d_out)
{
    const DinDev &d = h->din;
    const int T = d.T, E = d.E;
    const size_t smem = (size_t)kRowsRB * ((size_t)4 * E + (size_t)T * E + T + 1) * d.esz;
    const int grid = (int)std::min<int64_t>((n + kRowsRB - 1) / kRowsRB, (int64_t)h->sm_count * 8);
    cudaError_t terr = cudaSuccess;
    const bool tiled = d.dtype == DMG_F32
        ? rows_forward_tiled<float>(E, d.emb<float>(), (const float *)d.d_wattT, (const float *)d.d_w1T, d.b1<float>(), d.w2<float>(), d.b2<float>(),
                                    (float)(1.0 / std::sqrt((double)E)), T, n, dn, ds, d_mask, (float *)d_out, h->sm_count, h->smem_per_sm,
                                    h->smem_optin, h->stream, &terr)
        : rows_forward_tiled<double>(E, d.emb<double>(), (const double *)d.d_wattT, (const double *)d.d_w1T, d.b1<double>(), d.w2<double>(), d.b2<double>(),
                                     1.0 / std::sqrt((double)E), T, n, dn, ds, d_mask, (double *)d_out, h->sm_count, h->smem_per_sm,
                                     h->smem_optin, h->stream, &terr);
    DMG_CUDA(h, terr);
    if (tiled) {
    } else if (d.dtype == DMG_F32) {
        auto kern = din_rows_forward_kernel<float>;
        DMG_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kRowsThreads, smem, h->stream>>>(d.emb<float>(), (const float *)d.d_wattT, (const float *)d.d_w1T,
                                                      d.b1<float>(), d.w2<float>(), d.b2<float>(),
                                                      (float)(1.0 / std::sqrt((double)E)), E, T, n, dn, ds, d_mask,
                                                      (float *)d_out);
    } else {
        auto kern = din_rows_forward_kernel<double>;
        DMG_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kRowsThreads, smem, h->stream>>>(d.emb<double>(), (const double *)d.d_wattT, (const double *)d.d_w1T,
                                                      d.b1<double>(), d.w2<double>(), d.b2<double>(),
                                                      1.0 / std::sqrt((double)E), E, T, n, dn, ds, d_mask,
                                                      (double *)d_out);
    }
    h->launches += 1;
    DMG_CUDA(h, cudaGetLastError());
    return DMG_OK;
}

int32_t dmg_sanitize_indices(dmg_handle_t h, const int32_t *src, int32_t *dst, int64_t n, int64_t rows);   // train.cu

/* device-buffer variant of dmg_score_pairs (DIN scorers): d_mask = rows x T mask bytes (nullable), nothing copied or synchronised;
 * a bad index raises the handle's flag (dmg_synchronize returns DMG_ERR_INDEX) and scores as padding */
DMG_API int32_t dmg_score_pairs_dev(dmg_handle_t h, int64_t n, const int32_t *d_node, const int32_t *d_seq, const uint8_t *d_mask, void *d_out)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    if (!h->din.loaded || h->din.kind != 0) return fail(h, DMG_ERR_STATE, "DIN weights must be loaded first");
    if (h->din.sharded) return fail(h, DMG_ERR_STATE, "the node table is sharded (dmg_shard_init): use the dmg_shard_* entry points");
    if (n < 0 || (n > 0 && (!d_node || !d_seq || !d_out))) return fail(h, DMG_ERR_INVALID_ARG, "bad arguments");
    if (n == 0) return DMG_OK;
    DMG_CUDA(h, cudaSetDevice(h->device));
    const int T = h->din.T;
    DMG_TRY(ensure_dev(h, h->s_work, Carver::need({(size_t)n * 4, (size_t)n * T * 4, (size_t)n * T})));
    Carver cw(h->s_work.d);
    int32_t *dn = cw.take<int32_t>((size_t)n), *ds = cw.take<int32_t>((size_t)n * T);
    uint8_t *zmask = cw.take<uint8_t>((size_t)n * T);
    DMG_TRY(dmg_sanitize_indices(h, d_node, dn, n, h->din.rows));
    DMG_TRY(dmg_sanitize_indices(h, d_seq, ds, n * T, h->din.rows));
    if (!d_mask) DMG_CUDA(h, cudaMemsetAsync(zmask, 0, (size_t)n * T, h->stream));
    return score_pairs_enqueue(h, n, dn, ds, d_mask ? d_mask : zmask, d_out);
}

DMG_API int32_t dmg_score_pairs(dmg_handle_t h, int64_t n, const int32_t *node, const int32_t *seq, const int32_t *mask_flat,
                                int64_t n_mask, void *out)
{
    if (!h) return DMG_ERR_INVALID_ARG;
    if (!h->din.loaded) return fail(h, DMG_ERR_STATE, "DIN weights must be loaded first");
    if (h->din.kind == 1) {                                      // DeepFM takes no mask input (DeepFM.scala:14-15)
        if (n < 0 || (n > 0 && (!node || !seq || !out))) return fail(h, DMG_ERR_INVALID_ARG, "bad arguments");
        if (h->din.dtype == DMG_F64) return dmg_deepfm64_score_pairs(h, n, node, seq, (double *)out);   // otm_deepfm.cu
        return dmg_deepfm_score_pairs(h, n, node, seq, (float *)out);
    }
    if (h->din.sharded) return fail(h, DMG_ERR_STATE, "the node table is sharded (dmg_shard_init): use the dmg_shard_* entry points");
    if (n < 0 || n_mask < 0 || (n > 0 && (!node || !seq || !out)) || (n_mask > 0 && !mask_flat))
        return fail(h, DMG_ERR_INVALID_ARG, "bad arguments");
    if (n == 0) return DMG_OK;
    DMG_CUDA(h, cudaSetDevice(h->device));
    const DinDev &d = h->din;
    const int T = d.T, E = d.E;
    const size_t b_node = (size_t)n * 4, b_seq = (size_t)n * T * 4, b_mask = (size_t)n_mask * 4;
    const size_t in_bytes = Carver::need({b_node, b_seq, b_mask});
    DMG_TRY(ensure_host(h, h->s_in, in_bytes));
    DMG_TRY(ensure_dev(h, h->s_in, in_bytes));
    Carver ch(h->s_in.h), cd(h->s_in.d);
    int32_t *hn = ch.take<int32_t>((size_t)n), *dn = cd.take<int32_t>((size_t)n);
    int32_t *hs = ch.take<int32_t>((size_t)n * T), *ds = cd.take<int32_t>((size_t)n * T);
    int32_t *hm = ch.take<int32_t>((size_t)n_mask), *dm = cd.take<int32_t>((size_t)n_mask);
    memcpy(hn, node, b_node); memcpy(hs, seq, b_seq);
    if (n_mask) memcpy(hm, mask_flat, b_mask);
    DMG_CUDA(h, cudaMemcpyAsync(h->s_in.d, h->s_in.h, ch.off, cudaMemcpyHostToDevice, h->stream));
    DMG_TRY(ensure_dev(h, h->s_work, Carver::need({(size_t)n * T})));
    uint8_t *d_mask = (uint8_t *)h->s_work.d;
    DMG_CUDA(h, cudaMemsetAsync(d_mask, 0, (size_t)n * T, h->stream));
    if (n_mask) {
        mask_scatter_kernel<<<(unsigned)((n_mask + 255) / 256), 256, 0, h->stream>>>(dm, n_mask, n * T, d_mask, h->d_flags);
        h->launches += 1;
    }
    check_index_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(dn, n, d.rows, h->d_flags);
    check_index_kernel<<<(unsigned)((n * T + 255) / 256), 256, 0, h->stream>>>(ds, n * T, d.rows, h->d_flags);
    h->launches += 2;
    DMG_CUDA(h, cudaGetLastError());
    int32_t flag = 0;
    DMG_CUDA(h, cudaMemcpyAsync(&flag, h->d_flags, 4, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    if (flag) {
        DMG_CUDA(h, cudaMemsetAsync(h->d_flags, 0, 4, h->stream));
        return fail(h, DMG_ERR_INDEX, "dmg_score_pairs: embeddingLookup failed, index outside [0, %lld) or bad mask position", (long long)d.rows);
    }
    const size_t out_bytes = (size_t)n * d.esz;
    DMG_TRY(ensure_host(h, h->s_out, out_bytes));
    DMG_TRY(ensure_dev(h, h->s_out, out_bytes));
    DMG_TRY(score_pairs_enqueue(h, n, dn, ds, d_mask, h->s_out.d));
    DMG_CUDA(h, cudaGetLastError());
    DMG_CUDA(h, cudaMemcpyAsync(h->s_out.h, h->s_out.d, out_bytes, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    memcpy(out, h->s_out.h, out_bytes);
    return DMG_OK;
}

