// otm_deepfm.cu -- OTM with the DeepFM scorer in fp64 (second half).
//
// otm/src/main/scala/com/mass/otm/model/DeepFM.scala:12-48 is the TDM DeepFM graph instantiated for Double: features
// F = [item row ; T history rows] ((T+1) x E, padding rows zero, NO mask input),
//   FM  (scalann/.../nn/FM.scala:14-44)  (|sum_i F_i|^2 - sum |F|^2) / 2
//   DNN Linear((T+1)E, T+1) -> ReLU -> Linear(T+1, 1) on Fflat      (Linear.scala:19-56: bias after the product)
//   Add (nn/Add.scala)                   fm + dnn
// compact parameter vector [emb rows*E | W1 (T+1) x (T+1)E | b1 T+1 | W2 T+1 | b2 1] (Graph.scala:37-48).
//
// The search is CandidateSearcher.batchBeamSearch as the reference runs it (otm/.../model/CandidateSearcher.scala:15-56):
// level-synchronous, ONE model.forward over the B x 2b candidate rows of a level, then per user a stable descending
// sort, take(beam) and both children of every survivor (OTMTree.initializeBeam, otm/.../tree/OTMTree.scala:16-23).
// The Linear chains run over Fflat in order (item first, then the history), so nothing of a user can be hoisted out of
// a row without changing the rounding: every row walks T+2 sequential fma chains of (T+1)E steps -- fp64 FMA bound.
// Arithmetic = oracle/oracle.c orc_deepfm_row_f64, bit for bit (dmg_math.cuh).
#include <algorithm>
#include <cmath>

#include "dmg_common.cuh"
#include "dmg_math.cuh"

namespace dmg {
namespace {

constexpr int kRB = 8;          // rows per CTA iteration of the scorer

// chain c of a row whose features sit in x[(T+1)E] (item | history): c <= T hidden unit c, c == T + 1 the square sum
__device__ __forceinline__ double dfm64_chain(const double *x, const double *__restrict__ w1, const double *__restrict__ b1, int c, int E, int T)
{
    const int n = (T + 1) * E;
    double acc = 0.0;
    if (c <= T) {
        const double *w = w1 + (size_t)c * n;
        for (int k = 0; k < n; k++) acc = fma_(x[k], __ldg(w + k), acc);
        return relu_(add_(acc, __ldg(b1 + c)));
    }
    for (int k = 0; k < n; k++) acc = fma_(x[k], x[k], acc);
    return acc;
}

__device__ __forceinline__ double dfm64_finish(const double *x, const double *hrow, double square_sum, const double *__restrict__ w2,
                                               double b2, int E, int T)
{
    double sum_square = 0.0;
    for (int k = 0; k < E; k++) {
        double b = add_(0.0, x[k]);                               // vAdd from a zero buffer, feature 0 = the item
        for (int j = 0; j < T; j++) b = add_(b, x[(j + 1) * E + k]);
        sum_square = fma_(b, b, sum_square);
    }
    const double fm = __ddiv_rn(sub_(sum_square, square_sum), 2.0);
    double dnn = 0.0;
    for (int o = 0; o <= T; o++) dnn = fma_(hrow[o], __ldg(w2 + o), dnn);
    return add_(fm, add_(dnn, b2));
}

// model.forward on n_rows rows.  Row r belongs to group g = r / rows_per_group (a user of the search, or the row itself
// for dmg_score_pairs): its node code is node[g * node_stride + r % rows_per_group], its history hist[g * T ..], its
// logit goes to out[g * node_stride + r % rows_per_group].
__global__ void __launch_bounds__(128) dfm64_rows_kernel(const double *__restrict__ emb, const double *__restrict__ dense, int E, int T,
                                                         int64_t n_rows, const int32_t *__restrict__ node, int64_t node_stride,
                                                         int rows_per_group, const int32_t *__restrict__ hist, double *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int F = T + 1, HL = F + 1;
    double *sF = reinterpret_cast<double *>(smem_raw);   // kRB x F x E
    double *sH = sF + (size_t)kRB * F * E;               // kRB x HL: hidden units, then the square sum
    const double *w1 = dense, *b1 = w1 + (size_t)F * F * E, *w2 = b1 + F, *b2 = w2 + F;
    const int tid = threadIdx.x;
    for (int64_t g0 = (int64_t)blockIdx.x * kRB; g0 < n_rows; g0 += (int64_t)gridDim.x * kRB) {
        const int nr = (int)((n_rows - g0) < kRB ? (n_rows - g0) : kRB);
        for (int idx = tid; idx < nr * F * E; idx += 128) {
            const int k = idx % E, slot = (idx / E) % F, r = idx / (E * F);
            const int64_t row = g0 + r, g = row / rows_per_group, i = row % rows_per_group;
            const int32_t c = slot == 0 ? node[g * node_stride + i] : hist[g * T + slot - 1];
            sF[idx] = c < 0 ? 0.0 : emb[(size_t)c * E + k];
        }
        __syncthreads();
        for (int idx = tid; idx < nr * (F + 1); idx += 128) {
            const int r = idx % nr, c = idx / nr;
            sH[r * HL + c] = dfm64_chain(sF + (size_t)r * F * E, w1, b1, c, E, T);
        }
        __syncthreads();
        if (tid < nr) {
            const int64_t row = g0 + tid, g = row / rows_per_group, i = row % rows_per_group;
            out[g * node_stride + i] = dfm64_finish(sF + (size_t)tid * F * E, sH + tid * HL, sH[tid * HL + F], w2, __ldg(b2), E, T);
        }
        __syncthreads();
    }
}

// The same rows when every group is a USER of the search: its n candidate rows share one history, so the history tile is
// staged once per tile and only the item rows differ.  W1 sits in shared memory (row stride padded by one double: the
// T+2 chains of a warp read T+2 different banks), thread (chain c, row group rg) advances chain c of R rows together:
// one weight and one history load feed R independent fma chains (each still its own sequential-k chain: same bits).
// Work item = (user, tile of NG*R rows), persistent CTAs, 2 per SM.
template <int R>
__global__ void __launch_bounds__(128, 2) dfm64_user_rows_kernel(const double *__restrict__ emb, const double *__restrict__ dense, int E, int T,
                                                                 int B, int n, const int32_t *__restrict__ node, int64_t node_stride,
                                                                 const int32_t *__restrict__ hist, double *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int F = T + 1, NC = F + 1, n_in = F * E, WS = n_in + 1;
    const int NG = 128 / NC, TILE = NG * R;
    double *sW = reinterpret_cast<double *>(smem_raw);   // F x WS
    double *sB1 = sW + (size_t)F * WS;                   // F
    double *sW2 = sB1 + F;                               // F
    double *sK = sW2 + F;                                // T x E
    double *sX = sK + (size_t)T * E;                     // TILE x E (item rows, then the FM buffers)
    double *sH = sX + (size_t)TILE * E;                  // TILE x NC: hidden units, then the square sum
    const double *w1 = dense, *b1 = w1 + (size_t)F * n_in, *w2 = b1 + F, *b2 = w2 + F;
    const int tid = threadIdx.x;
    for (int i = tid; i < F * n_in; i += 128) sW[(i / n_in) * WS + i % n_in] = w1[i];
    for (int i = tid; i < F; i += 128) { sB1[i] = b1[i]; sW2[i] = w2[i]; }
    const double b2v = __ldg(b2);
    const int tiles = (n + TILE - 1) / TILE;
    const int c = tid % NC, rg = tid / NC;
    int staged_user = -1;
    for (int64_t item = blockIdx.x; item < (int64_t)B * tiles; item += gridDim.x) {
        const int user = (int)(item / tiles), r0 = (int)(item % tiles) * TILE;
        const int nr = n - r0 < TILE ? n - r0 : TILE;
        __syncthreads();                                 // the previous item's sX / sH are free (and sW is complete)
        if (user != staged_user) {
            for (int i = tid; i < T * E; i += 128) {
                const int32_t hc = hist[(size_t)user * T + i / E];
                sK[i] = hc < 0 ? 0.0 : emb[(size_t)hc * E + i % E];
            }
            staged_user = user;
        }
        for (int i = tid; i < nr * E; i += 128) {
            const int32_t nc = node[(size_t)user * node_stride + r0 + i / E];
            sX[i] = nc < 0 ? 0.0 : emb[(size_t)nc * E + i % E];
        }
        __syncthreads();
        if (rg < NG) {
            double acc[R];
            const double *x[R];
#pragma unroll
            for (int i = 0; i < R; i++) { acc[i] = 0.0; x[i] = sX + (size_t)(rg * R + i < nr ? rg * R + i : 0) * E; }
            if (c < F) {
                const double *w = sW + (size_t)c * WS;
                for (int k = 0; k < E; k++) {
                    const double wk = w[k];
#pragma unroll
                    for (int i = 0; i < R; i++) acc[i] = fma_(x[i][k], wk, acc[i]);
                }
                for (int k = 0; k < T * E; k++) {
                    const double wk = w[E + k], kk = sK[k];
#pragma unroll
                    for (int i = 0; i < R; i++) acc[i] = fma_(kk, wk, acc[i]);
                }
                const double bc = sB1[c];
#pragma unroll
                for (int i = 0; i < R; i++) acc[i] = relu_(add_(acc[i], bc));
            } else {
                for (int k = 0; k < E; k++) {
#pragma unroll
                    for (int i = 0; i < R; i++) acc[i] = fma_(x[i][k], x[i][k], acc[i]);
                }
                for (int k = 0; k < T * E; k++) {
                    const double kk = sK[k];
#pragma unroll
                    for (int i = 0; i < R; i++) acc[i] = fma_(kk, kk, acc[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < R; i++)
                if (rg * R + i < nr) sH[(rg * R + i) * NC + c] = acc[i];
        }
        __syncthreads();
        for (int i = tid; i < nr * E; i += 128) {        // FM buffer of (row, k): vAdd from zero, item first, then the history rows
            const int k = i % E;
            double bsum = add_(0.0, sX[i]);
            for (int j = 0; j < T; j++) bsum = add_(bsum, sK[j * E + k]);
            sX[i] = bsum;
        }
        __syncthreads();
        if (tid < nr) {
            const double *bb = sX + (size_t)tid * E, *hrow = sH + tid * NC;
            double sum_square = 0.0;
            for (int k = 0; k < E; k++) sum_square = fma_(bb[k], bb[k], sum_square);
            const double fm = __ddiv_rn(sub_(sum_square, hrow[F]), 2.0);
            double dnn = 0.0;
            for (int o = 0; o < F; o++) dnn = fma_(hrow[o], sW2[o], dnn);
            out[(size_t)user * node_stride + r0 + tid] = add_(fm, add_(dnn, b2v));
        }
    }
}

// OTMTree.initializeBeam: all nodes of the start level, score 0
__global__ void otm_dfm_init_kernel(int B, int n0, int width, int32_t start, int32_t *__restrict__ ids, double *__restrict__ sc)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * n0) return;
    const int u = (int)(i / n0), j = (int)(i % n0);
    ids[(size_t)u * width + j] = start + j;
    sc[(size_t)u * width + j] = 0.0;
}

// One CTA per user: sortBy(_.score)(reverse).take(nb) -- stable, Double.compare order -- then both children of every
// survivor in order (CandidateSearcher.scala:26-38).  Ranks by counting: rank_i = #{key_j > key_i} + #{j < i, key_j == key_i}.
__global__ void __launch_bounds__(256) otm_dfm_expand_kernel(int n, int nb, int select, int width, const int32_t *__restrict__ ids,
                                                             const double *__restrict__ sc, int32_t *__restrict__ nxt)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *sKey = reinterpret_cast<uint64_t *>(smem_raw);
    const size_t base = (size_t)blockIdx.x * width;
    if (!select) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int32_t c = ids[base + i];
            nxt[base + 2 * i] = 2 * c + 1;
            nxt[base + 2 * i + 1] = 2 * c + 2;
        }
        return;
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) sKey[i] = order_key(sc[base + i]);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint64_t k = sKey[i];
        int rank = 0;
        for (int j = 0; j < n; j++) {
            const uint64_t kj = sKey[j];
            rank += (kj > k) || (kj == k && j < i);
        }
        if (rank < nb) {
            const int32_t c = ids[base + i];
            nxt[base + 2 * rank] = 2 * c + 1;
            nxt[base + 2 * rank + 1] = 2 * c + 2;
        }
    }
}

// candidates of a level, padded to `width` (-1 / 0) like the fused kernel's dump: dst slot (user * n_slots + slot)
__global__ void otm_dfm_dump_kernel(int B, int n, int width, int n_slots, int slot, const int32_t *__restrict__ ids,
                                    const double *__restrict__ sc, int32_t *__restrict__ out_ids, double *__restrict__ out_sc,
                                    int32_t *__restrict__ out_counts)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * width) return;
    const int u = (int)(i / width), j = (int)(i % width);
    const size_t dst = ((size_t)u * n_slots + slot) * width + j;
    out_ids[dst] = j < n ? ids[i] : -1;
    out_sc[dst] = j < n ? sc[i] : 0.0;
    if (j == 0) out_counts[(size_t)u * n_slots + slot] = n;
}

// OTM.recommend (otm/.../model/OTM.scala:14-23): keep the leaf ids that map back to an item, stable sort desc, topk.
__global__ void __launch_bounds__(256) otm_dfm_topk_kernel(int n, int width, int topk, int leaf_level, const int32_t *__restrict__ leaf_item,
                                                           const int32_t *__restrict__ ids, const double *__restrict__ sc,
                                                           int32_t *__restrict__ out_items, double *__restrict__ out_scores,
                                                           int32_t *__restrict__ out_counts)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *sKey = reinterpret_cast<uint64_t *>(smem_raw);
    int32_t *sItem = reinterpret_cast<int32_t *>(sKey + n);
    const int user = blockIdx.x;
    const size_t base = (size_t)user * width;
    const int64_t leaf_start = ((int64_t)1 << leaf_level) - 1, n_leaf = (int64_t)1 << leaf_level;
    __shared__ int s_valid;
    if (threadIdx.x == 0) s_valid = 0;
    for (int i = threadIdx.x; i < topk; i += blockDim.x) {
        out_items[(size_t)user * topk + i] = -1;
        out_scores[(size_t)user * topk + i] = 0.0;
    }
    int mine = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int64_t slot = (int64_t)ids[base + i] - leaf_start;
        const int32_t item = (slot >= 0 && slot < n_leaf) ? __ldg(leaf_item + slot) : -1;
        sItem[i] = item;
        sKey[i] = order_key(sc[base + i]);
        mine += item >= 0;
    }
    __syncthreads();
    if (mine) atomicAdd(&s_valid, mine);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (sItem[i] < 0) continue;
        const uint64_t k = sKey[i];
        int rank = 0;
        for (int j = 0; j < n; j++) {
            const uint64_t kj = sKey[j];
            rank += sItem[j] >= 0 && ((kj > k) || (kj == k && j < i));
        }
        if (rank < topk) {
            out_items[(size_t)user * topk + rank] = sItem[i];
            out_scores[(size_t)user * topk + rank] = sc[base + i];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) out_counts[user] = s_valid < topk ? s_valid : topk;
}

int32_t launch_rows(dmg_handle_t h, int64_t n_rows, const int32_t *node, int64_t node_stride, int rows_per_group, const int32_t *hist,
                    double *out)
{
    const DinDev &d = h->din;
    const int F = d.T + 1;
    const size_t smem = ((size_t)kRB * F * d.E + (size_t)kRB * (F + 1)) * sizeof(double);
    if (smem > h->smem_optin) return fail(h, DMG_ERR_UNSUPPORTED, "DeepFM(Double): embed_size %d x seq_len %d does not fit shared memory", d.E, d.T);
    DMG_CUDA(h, cudaFuncSetAttribute(dfm64_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<int64_t>((n_rows + kRB - 1) / kRB, (int64_t)h->sm_count * 4);   // multiples of the SM count, 4 CTAs of 46 KB each
    dfm64_rows_kernel<<<grid, 128, smem, h->stream>>>(d.emb<double>(), d.tail<double>(), d.E, d.T, n_rows, node, node_stride,
                                                      rows_per_group, hist, out);
    h->launches += 1;
    DMG_CUDA(h, cudaGetLastError());
    return DMG_OK;
}

// the n candidate rows of each of B users (node[user * node_stride + i], one history per user)
int32_t launch_user_rows(dmg_handle_t h, int B, int n, const int32_t *node, int64_t node_stride, const int32_t *hist, double *out)
{
    const DinDev &d = h->din;
    constexpr int R = 4;
    const int F = d.T + 1, NC = F + 1, NG = 128 / NC, TILE = NG * R;
    const size_t smem = ((size_t)F * (F * d.E + 1) + 2 * F + (size_t)d.T * d.E + (size_t)TILE * d.E + (size_t)TILE * NC) * sizeof(double);
    if (NG < 1 || 2 * (smem + 1024) > h->smem_per_sm)             // large E x T: the generic kernel (weights through L1)
        return launch_rows(h, (int64_t)B * n, node, node_stride, n, hist, out);
    DMG_CUDA(h, cudaFuncSetAttribute(dfm64_user_rows_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t items = (int64_t)B * ((n + TILE - 1) / TILE);
    const int grid = (int)std::min<int64_t>(items, (int64_t)h->sm_count * 2);      // persistent, two CTAs per SM
    dfm64_user_rows_kernel<R><<<grid, 128, smem, h->stream>>>(d.emb<double>(), d.tail<double>(), d.E, d.T, B, n, node, node_stride, hist, out);
    h->launches += 1;
    DMG_CUDA(h, cudaGetLastError());
    return DMG_OK;
}

int lower_log2(int n) { int l = 0; while ((2 << l) <= n) l++; return l; }

}  // namespace
}  // namespace dmg

using namespace dmg;

// dmg_score_pairs with a DeepFM(Double) model loaded (called from capi.cu): model.forward, no mask input
int32_t dmg_deepfm64_score_pairs(dmg_handle_t h, int64_t n, const int32_t *node, const int32_t *seq, double *out)
{
    const DinDev &d = h->din;
    if (n <= 0) return DMG_OK;
    DMG_CUDA(h, cudaSetDevice(h->device));
    const int T = d.T;
    for (int64_t i = 0; i < n * (T + 1); i++) {                   // LookupTable.scala:46-52
        const int32_t c = i < n ? node[i] : seq[i - n];
        if (c < -1 || c >= d.rows) return fail(h, DMG_ERR_INDEX, "dmg_score_pairs: embeddingLookup failed, index outside [0, %lld)", (long long)d.rows);
    }
    DMG_TRY(ensure_dev(h, h->s_in, Carver::need({(size_t)n * 4, (size_t)n * T * 4, (size_t)n * 8})));
    Carver cd(h->s_in.d);
    int32_t *dn = cd.take<int32_t>((size_t)n), *ds = cd.take<int32_t>((size_t)n * T);
    double *dout = cd.take<double>((size_t)n);
    DMG_CUDA(h, cudaMemcpyAsync(dn, node, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
    DMG_CUDA(h, cudaMemcpyAsync(ds, seq, (size_t)n * T * 4, cudaMemcpyHostToDevice, h->stream));
    DMG_TRY(launch_rows(h, n, dn, 1, 1, ds, dout));
    DMG_CUDA(h, cudaMemcpyAsync(out, dout, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    return DMG_OK;
}

// dmg_otm_beam_search / dmg_otm_beam_search_levels / dmg_otm_retrieve with a DeepFM(Double) model loaded (called from
// capi.cu after its argument checks).  topk_mode: OTM.recommend; otherwise the candidates of the last scored level.
int32_t dmg_deepfm64_otm_run(dmg_handle_t h, int32_t B, const int32_t *leaf_seq, int32_t beam, int topk_mode, int32_t topk,
                             int32_t *out_ids, double *out_scores, int32_t *out_counts, int32_t *lvl_ids, double *lvl_scores,
                             int32_t *lvl_counts)
{
    const DinDev &d = h->din;
    const TreeDev &t = h->tree;
    const int T = d.T, L = t.max_level;
    DMG_CUDA(h, cudaSetDevice(h->device));
    for (int64_t i = 0; i < (int64_t)B * T; i++)
        if (leaf_seq[i] < -1 || leaf_seq[i] >= d.rows)
            return fail(h, DMG_ERR_INDEX, "dmg_otm: embeddingLookup failed, index outside [0, %lld)", (long long)d.rows);
    const int s = lower_log2(beam);                               // otm/package.scala:15
    const int n0 = 1 << s;
    const int width = 2 * std::max(beam, n0);
    const int stride = topk_mode ? topk : width;
    const int n_lvl = std::max(L - s, 0);
    if (width > 8192) return fail(h, DMG_ERR_UNSUPPORTED, "DeepFM(Double) search: beam %d too wide", beam);
    const size_t out_bytes = Carver::need({(size_t)B * stride * 4, (size_t)B * stride * 8, (size_t)B * 4});
    const size_t work_bytes = Carver::need({(size_t)B * T * 4, (size_t)B * width * 4, (size_t)B * width * 4, (size_t)B * width * 8});
    DMG_TRY(ensure_dev(h, h->s_out, out_bytes));
    DMG_TRY(ensure_dev(h, h->s_work, work_bytes));
    Carver od(h->s_out.d), cw(h->s_work.d);
    int32_t *d_oi = od.take<int32_t>((size_t)B * stride);
    double *d_os = od.take<double>((size_t)B * stride);
    int32_t *d_oc = od.take<int32_t>(B);
    int32_t *d_seq = cw.take<int32_t>((size_t)B * T);
    int32_t *d_ids = cw.take<int32_t>((size_t)B * width), *d_nxt = cw.take<int32_t>((size_t)B * width);
    double *d_sc = cw.take<double>((size_t)B * width);
    DMG_CUDA(h, cudaMemcpyAsync(d_seq, leaf_seq, (size_t)B * T * 4, cudaMemcpyHostToDevice, h->stream));
    int32_t *d_li = nullptr, *d_lc = nullptr;
    double *d_ls = nullptr;
    void *lvl_blk = nullptr;
    if (lvl_ids && n_lvl > 0) {
        DMG_CUDA(h, cudaMalloc(&lvl_blk, Carver::need({(size_t)B * n_lvl * width * 4, (size_t)B * n_lvl * width * 8, (size_t)B * n_lvl * 4})));
        Carver c2(lvl_blk);
        d_li = c2.take<int32_t>((size_t)B * n_lvl * width);
        d_ls = c2.take<double>((size_t)B * n_lvl * width);
        d_lc = c2.take<int32_t>((size_t)B * n_lvl);
    }
    int32_t rc = DMG_OK;
    int n = 0;
    if (s <= L) {                                                 // deeper start level than the tree: no candidates
        n = n0;
        const int64_t tot = (int64_t)B * n0;
        otm_dfm_init_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(B, n0, width, (int32_t)(((int64_t)1 << s) - 1), d_ids, d_sc);
        h->launches += 1;
        for (int level = s; level < L && rc == DMG_OK; level++) {
            const int select = level != s;
            const int nb = select ? std::min(n, beam) : n;
            otm_dfm_expand_kernel<<<B, 256, (size_t)n * 8, h->stream>>>(n, nb, select, width, d_ids, d_sc, d_nxt);
            h->launches += 1;
            n = 2 * nb;
            rc = launch_user_rows(h, B, n, d_nxt, width, d_seq, d_sc);
            std::swap(d_ids, d_nxt);
            if (rc == DMG_OK && d_li) {
                const int64_t tw = (int64_t)B * width;
                otm_dfm_dump_kernel<<<(unsigned)((tw + 255) / 256), 256, 0, h->stream>>>(B, n, width, n_lvl, level - s, d_ids, d_sc, d_li, d_ls, d_lc);
                h->launches += 1;
            }
        }
    }
    if (rc == DMG_OK) {
        if (topk_mode) {
            otm_dfm_topk_kernel<<<B, 256, (size_t)std::max(n, 1) * 12, h->stream>>>(n, width, topk, L, t.d_leaf_item, d_ids, d_sc, d_oi, d_os, d_oc);
        } else {
            const int64_t tw = (int64_t)B * width;
            otm_dfm_dump_kernel<<<(unsigned)((tw + 255) / 256), 256, 0, h->stream>>>(B, n, width, 1, 0, d_ids, d_sc, d_oi, d_os, d_oc);
        }
        h->launches += 1;
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(out_ids, d_oi, (size_t)B * stride * 4, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(out_scores, d_os, (size_t)B * stride * 8, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(out_counts, d_oc, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess && d_li) {
            e = cudaMemcpyAsync(lvl_ids, d_li, (size_t)B * n_lvl * width * 4, cudaMemcpyDeviceToHost, h->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(lvl_scores, d_ls, (size_t)B * n_lvl * width * 8, cudaMemcpyDeviceToHost, h->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(lvl_counts, d_lc, (size_t)B * n_lvl * 4, cudaMemcpyDeviceToHost, h->stream);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = fail(h, DMG_ERR_CUDA, "DeepFM(Double) search: %s", cudaGetErrorString(e));
    } else {
        cudaStreamSynchronize(h->stream);
    }
    if (lvl_blk) cudaFree(lvl_blk);
    return rc;
}

// DeepModel[Double] = DeepFM (otm/.../model/DeepFM.scala:12-48): params = [emb | W1 (T+1)x(T+1)E | b1 | W2 | b2] as doubles.
// Afterwards dmg_otm_beam_search / dmg_otm_beam_search_levels / dmg_otm_retrieve / dmg_score_pairs run this scorer.
DMG_API int32_t dmg_load_deepfm_weights_f64(dmg_handle_t h, int64_t rows, int32_t E, int32_t T, const double *params)
{
    if (!h || !params) return h ? fail(h, DMG_ERR_INVALID_ARG, "dmg_load_deepfm_weights_f64: null params") : DMG_ERR_INVALID_ARG;
    DMG_TRY(model_is_shared(h, "dmg_load_deepfm_weights_f64"));
    if (rows <= 0 || E <= 0 || T <= 0) return fail(h, DMG_ERR_INVALID_ARG, "rows, E, T must be positive");
    if (T > kMaxT) return fail(h, DMG_ERR_UNSUPPORTED, "seq_len %d > %d", T, kMaxT);
    DMG_CUDA(h, cudaSetDevice(h->device));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    DinDev &d = h->din;
    cudaFree(d.d_params); cudaFree(d.d_wattT); cudaFree(d.d_w1T); cudaFree(d.d_grad); cudaFree(d.d_m); cudaFree(d.d_v);
    d = DinDev();
    const int64_t F = T + 1;
    d.dtype = DMG_F64; d.esz = 8; d.E = E; d.T = T; d.rows = rows; d.kind = 1;
    d.n_params = rows * E + F * F * E + 2 * F + 1;
    DMG_CUDA(h, cudaMalloc(&d.d_params, (size_t)d.n_params * 8));
    DMG_CUDA(h, cudaMemcpyAsync(d.d_params, params, (size_t)d.n_params * 8, cudaMemcpyHostToDevice, h->stream));
    DMG_CUDA(h, cudaStreamSynchronize(h->stream));
    h->fast_dirty = true;
    d.sharded = true;                                            // same convention as the Float DeepFM: no fused-kernel / training entry points
    d.loaded = true;
    return DMG_OK;
}
