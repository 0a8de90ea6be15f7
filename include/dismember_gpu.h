/*
 * dismember_gpu.h -- C ABI of the B200-native retrieval engine (libdismember_gpu.so).
 *
 * Drop-in boundary for ONE path of massquantity/dismember: level-synchronous
 * beam search over the TDM/JTM/OTM binary tree and Deep Retrieval's K^D paths,
 * the DIN scorer behind it, its training step and the JTM re-assignment scores.
 * The reference has no FFI of its own on this path -- the seam is the Scala call
 * `model.forward(features)` made once per tree level by the searchers and, one
 * level up, recommend/_recommend/batchBeamSearch.  Every entry point below names
 * the reference interface it replaces (paths relative to the reference root).
 * The JNI shim (jni/com_mass_gpu_DismemberGPU.c) and the ctypes binding
 * (dismember_b200/_capi.py) bind exactly these symbols; see INTEGRATION.md.
 *
 * Conventions
 *  - every function returns int32 status: 0 = DMG_OK, <0 = error; the message
 *    is available from dmg_last_error(h).  No exception crosses the boundary.
 *  - plain pointers and sizes only.  Unless a name ends in _dev, pointers are
 *    HOST memory owned by the caller; the library copies in/out and returns
 *    after the results are complete.  *_dev variants take DEVICE pointers,
 *    enqueue on the handle's stream and return without synchronising.
 *  - a handle is bound to one CUDA device + one stream and is NOT thread-safe:
 *    one handle per host thread, mirroring "one model clone per thread"
 *    (tdm/src/main/scala/com/mass/tdm/optim/LocalOptimizer.scala:35-40).
 *  - there is no CPU fallback: without a CUDA device dmg_create fails.
 */
#ifndef DISMEMBER_GPU_H
#define DISMEMBER_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define DMG_API __declspec(dllexport)
#else
#define DMG_API __attribute__((visibility("default")))
#endif

typedef struct dmg_handle_s *dmg_handle_t;

enum {
    DMG_OK = 0,
    DMG_ERR_INVALID_ARG = -1,  /* IllegalArgumentException / require(...) failures          */
    DMG_ERR_CUDA = -2,         /* CUDA runtime error (message holds cudaGetErrorString)      */
    DMG_ERR_INDEX = -3,        /* ArrayIndexOutOfBoundsException of LookupTable.scala:46-52  */
    DMG_ERR_STATE = -4,        /* tree / weights not loaded yet                              */
    DMG_ERR_UNSUPPORTED = -5,  /* shape outside what the kernels are built for               */
    DMG_ERR_NOMEM = -6
};

enum { DMG_F32 = 0, DMG_F64 = 1 };

/* ---- lifecycle ------------------------------------------------------------------------ */
DMG_API int32_t dmg_create(int32_t device, dmg_handle_t *out);
DMG_API int32_t dmg_destroy(dmg_handle_t h);
/* A second handle on src's device that SHARES src's tree index and weight tables read-only (own
 * stream, scratch and scheduler state): the GPU form of the reference's per-thread model clones over
 * one weight storage (tdm/src/main/scala/com/mass/tdm/optim/LocalOptimizer.scala:35-40;
 * otm/src/main/scala/com/mass/otm/optim/LocalOptimizer.scala:225-230 cloneModule + putWeights, pinned by
 * otm/src/test/scala/CloneModelSpec.scala: "cloned models share same weights storage") that the
 * evaluator hands its user slices to (tdm/.../evaluation/Evaluator.scala:29-37).  One clone per host
 * thread keeps several batches in flight on one GPU: the tail of one batch's persistent kernel
 * overlaps the head of the next.  While clones live, loaders and training entry points on src (and
 * always on a clone) return DMG_ERR_STATE; destroy the clones before src. */
DMG_API int32_t dmg_clone(dmg_handle_t src, dmg_handle_t *out);
DMG_API const char *dmg_last_error(dmg_handle_t h);      /* h may be NULL: last create error   */
DMG_API const char *dmg_version(void);
/* Use a caller-owned cudaStream_t (e.g. the framework's current stream) instead of the
 * handle's own; pass NULL to go back.  Lets callers time with their own CUDA events. */
DMG_API int32_t dmg_set_stream(dmg_handle_t h, void *cuda_stream);
DMG_API int32_t dmg_synchronize(dmg_handle_t h);
/* How the synchronous retrieval calls wait for their batch: 0 = spin (default, lowest latency), 1 = sleep on a blocking event
 * (hosts where the waiting threads of all handles / processes outnumber the cores, e.g. 8 GPUs x 8 serving threads on 32 cores).
 * Clones inherit the mode of their parent at dmg_clone. */
DMG_API int32_t dmg_set_sync_mode(dmg_handle_t h, int32_t mode);
/* number of kernels this handle has launched so far (bench.py's gpu_launches). */
DMG_API int64_t dmg_launch_count(dmg_handle_t h);
/* Optional per-kernel timing of the dominant (beam-search) kernel: when on, every launch is
 * bracketed by CUDA events on the launching stream.  dmg_kernel_time synchronises, returns
 * the accumulated milliseconds and launch count since the last call, and resets both. */
DMG_API int32_t dmg_set_profiling(dmg_handle_t h, int32_t on);
DMG_API int32_t dmg_kernel_time(dmg_handle_t h, double *total_ms, int64_t *n_launches);

/* ---- index structures ----------------------------------------------------------------- */
/* TDM/JTM tree = the maps DistTree.loadData/loadItems build
 * (tdm/src/main/scala/com/mass/tdm/tree/DistTree.scala:25-87): one entry per stored node
 * (codes/node_ids/is_leaf = codeNodeMap) and the Part_* leaf (id, code) pairs (idCodeMap).
 * nonLeafOffset = max(leaf id)+1 and maxCode = max(leaf code) are derived as in :35-36.
 * All writers of the format put every leaf at max_level (TreeBuilder.flattenLeaves
 * TreeBuilder.scala:133-140, JTMTree.writeTree JTMTree.scala:115-182); a tree with a leaf
 * above max_level is rejected with DMG_ERR_UNSUPPORTED.  prob (nullable) = Node.probality of
 * every stored node (tree.proto:4-9), the weights of NegativeSampler's withProb sampling. */
DMG_API int32_t dmg_load_tree_tdm(dmg_handle_t h, int32_t max_level, int64_t n_nodes,
                                  const int32_t *codes, const int32_t *node_ids,
                                  const uint8_t *is_leaf, int64_t n_items,
                                  const int32_t *leaf_ids, const int32_t *leaf_codes,
                                  const float *prob);

/* OTM: complete binary tree; itemIdMapping item -> leaf node id
 * (otm/src/main/scala/com/mass/otm/model/OTM.scala:6-12, Serialization.loadMapping).
 * leaf_level = upperLog2(n_items) (OTM.scala:12). */
DMG_API int32_t dmg_load_tree_complete(dmg_handle_t h, int32_t leaf_level, int64_t n_items,
                                       const int32_t *item_ids, const int32_t *leaf_ids);

/* ---- DIN scorer weights --------------------------------------------------------------- */
/* The compact parameter vector of Module.parameters()/adjustParameters()
 * (scalann/.../nn/graphnn/Graph.scala:37-48, nn/abstractnn/AbstractModule.scala:163):
 *   [ emb rows*E | W_att E*E | W1 E*2E | b1 E | W2 E | b2 1 ]  row-major [out,in],
 * dtype DMG_F32 (tdm/jtm DIN.buildModel[Float]) or DMG_F64 (otm DIN.buildModel[Double]).
 * rows = numIndex of EmbeddingShare (tdm DIN.scala:18: 2^(maxLevel+1)-1).  T = seq_len. */
DMG_API int32_t dmg_load_din_weights(dmg_handle_t h, int32_t dtype, int64_t rows, int32_t E,
                                     int32_t T, const void *params);
/* Same, node table initialised on the device like EmbeddingShare/Linear do at construction
 * (randn(0, 0.05), biases 0: EmbeddingShare.scala:21, Linear.scala:12-13), counter-based RNG. */
DMG_API int32_t dmg_init_din_weights(dmg_handle_t h, int32_t dtype, int64_t rows, int32_t E,
                                     int32_t T, uint64_t seed);
/* Shape of the loaded scorer: node-table rows, embed_size, seq_len, DMG_F32 / DMG_F64 (a binding sizes its arrays with it). */
DMG_API int32_t dmg_din_shape(dmg_handle_t h, int64_t *rows, int32_t *E, int32_t *T, int32_t *dtype);
/* Copy the compact vector back (Serialization.saveModel side).  n = element count. */
DMG_API int32_t dmg_download_din_weights(dmg_handle_t h, void *params, int64_t n);

/* Scorer arithmetic of dmg_tdm_retrieve.  Both modes return the SAME item ids and logits (the
 * CPU oracle's bits).  DMG_ARITH_STRICT scores every candidate with sequential-k fp32 fma chains;
 * DMG_ARITH_FAST (E = 64 fp32 models, beam <= 256) runs the dense contractions on the tcgen05 tensor
 * cores with bf16 hi/lo split operands, bounds |fast - strict| per tree level, and re-scores in strict
 * arithmetic only the candidates whose fast score is within that bound of a beam cut (certified cuts)
 * plus the final topk; a user whose strictly re-scored candidates tie exactly at a decision point is
 * re-run by the strict kernel.  dmg_fast_stats fills 7 values since the last call: {cuts, cuts that
 * needed a strict re-score, rows re-scored strictly, rows scored on the tensor cores, float bits of
 * the largest observed |fast - strict| / bound (must stay < 1), users re-run by the strict kernel,
 * cuts whose band was settled strictly in place instead of being deferred to the end-of-search proof}. */
enum { DMG_ARITH_STRICT = 0, DMG_ARITH_FAST = 1 };
DMG_API int32_t dmg_set_arithmetic(dmg_handle_t h, int32_t mode);
DMG_API int32_t dmg_fast_stats(dmg_handle_t h, uint64_t *out7);
/* Certification band = tau x the worst-case bound.  tau = 1 (default) is provable: rounding errors
 * would all have to align.  Smaller tau trades the proof for a statistical margin (the largest error
 * ever observed is reported by dmg_fast_stats as a fraction of the worst-case bound). */
DMG_API int32_t dmg_set_fast_tolerance(dmg_handle_t h, double tau);
/* Diagnostic of the level-synchronous tensor-core search (DMG_ARITH_FAST): run Recommender._recommend's level loop
 * (Recommender.scala:58-99) for B users until the candidates of tree level `level` are scored and return them:
 * out_codes / out_scores [B x cap] (cap = 2*beam rounded up to 8), out_counts [B], out_eps [B] = the bound on
 * |fast - strict| the certified cuts use for these scores (-1: the user was handed to the strict kernel). */
DMG_API int32_t dmg_wave_probe(dmg_handle_t h, int32_t B, const int32_t *item_seq, int32_t beam, int32_t use_mask,
                               int32_t level, int32_t cap, int32_t *out_codes, float *out_scores,
                               int32_t *out_counts, float *out_eps);

/* ---- retrieval ------------------------------------------------------------------------ */
/* Recommender.recommendItems / TDM.recommend over a batch of users
 * (tdm/.../model/Recommender.scala:18-107, TDM.scala:17-22; batched caller
 * tdm/.../evaluation/Evaluator.scala:51-66).
 *   item_seq       B x T item ids, 0 = padding (TDMTree.idToCode, TDMTree.scala:35-56)
 *   beam           candidateNum; consumed_off != NULL && widen_beam: per user
 *                  max((|consumed|+topk)/2, beam) as in Recommender.scala:28-31
 *   consumed_off   nullable, B+1 offsets into consumed_items (item ids to drop, :104)
 *   out_items      B x topk item ids, -1 padded;  out_logits B x topk raw logits (apply
 *                  TDM.sigmoid on the host);  out_counts B = number of valid entries.
 * Ties are broken exactly like the reference's stable sorts (earlier candidate first). */
DMG_API int32_t dmg_tdm_retrieve(dmg_handle_t h, int32_t B, const int32_t *item_seq, int32_t beam,
                                 int32_t topk, int32_t use_mask, const int64_t *consumed_off,
                                 const int32_t *consumed_items, int32_t widen_beam,
                                 int32_t *out_items, float *out_logits, int32_t *out_counts);
DMG_API int32_t dmg_tdm_retrieve_dev(dmg_handle_t h, int32_t B, const int32_t *d_item_seq,
                                     int32_t beam, int32_t topk, int32_t use_mask,
                                     int32_t *d_out_items, float *d_out_logits,
                                     int32_t *d_out_counts);
/* Same device buffers, but returns after the results are complete (one stream synchronisation).  The
 * strict re-run of the users the tensor-core kernel could not certify is then launched only for the
 * batches that have such users; with one handle per host thread (dmg_clone) the next batch's kernel
 * fills the SMs this batch's tail leaves idle. */
DMG_API int32_t dmg_tdm_retrieve_dev_sync(dmg_handle_t h, int32_t B, const int32_t *d_item_seq,
                                          int32_t beam, int32_t topk, int32_t use_mask,
                                          int32_t *d_out_items, float *d_out_logits,
                                          int32_t *d_out_counts);

/* CandidateSearcher.batchBeamSearch (otm/.../model/CandidateSearcher.scala:15-56): per user
 * the 2*beam candidates of the leaf level with their Double scores, in candidate order.
 *   leaf_seq  B x T leaf node ids, -1 = padding.  out_ids/out_scores: B x 2*max(beam,2^s). */
DMG_API int32_t dmg_otm_beam_search(dmg_handle_t h, int32_t B, const int32_t *leaf_seq,
                                    int32_t beam, int32_t use_mask, int32_t *out_ids,
                                    double *out_scores, int32_t *out_counts);
/* OTMTree.beamSearchNodes (otm/.../tree/OTMTree.scala:67-91,174-212): the same search, returning the
 * scored candidates of EVERY level startLevel+1..leafLevel (the rows the OTM trainer fits, one Adam
 * step per level).  out_ids/out_scores: B x (leafLevel-startLevel) x 2*max(beam,2^s); out_counts:
 * B x (leafLevel-startLevel). */
DMG_API int32_t dmg_otm_beam_search_levels(dmg_handle_t h, int32_t B, const int32_t *leaf_seq,
                                           int32_t beam, int32_t use_mask, int32_t *out_ids,
                                           double *out_scores, int32_t *out_counts);
/* OTM.recommend over a batch (otm/.../model/OTM.scala:14-23): leaf candidates that map back
 * to an item, stable sort desc, topk.  out_scores = raw logits (sigmoid on the host). */
DMG_API int32_t dmg_otm_retrieve(dmg_handle_t h, int32_t B, const int32_t *leaf_seq, int32_t beam,
                                 int32_t topk, int32_t use_mask, int32_t *out_items,
                                 double *out_scores, int32_t *out_counts);

/* model.forward(Table(item, seq, mask)) on n independent rows -- the seam itself
 * (Recommender.scala:94, otm CandidateSearcher.scala:41,77, OTMTree.scala:168,198,
 * jtm TreeLearning.scala:168).  node[n], seq[n*T] = embedding indices, -1 = padding;
 * mask_flat = flat positions row*T+j (nullable).  out: n logits of the loaded dtype. */
DMG_API int32_t dmg_score_pairs(dmg_handle_t h, int64_t n, const int32_t *node, const int32_t *seq,
                                const int32_t *mask_flat, int64_t n_mask, void *out);

/* Device-buffer variant (DIN scorers): d_node[n], d_seq[n*T], d_mask = n x T mask BYTES (1 = masked; nullable), d_out[n]
 * of the loaded dtype.  Enqueues on the handle's stream and returns; an index outside [-1, rows) scores as padding
 * and makes the next dmg_synchronize return DMG_ERR_INDEX. */
DMG_API int32_t dmg_score_pairs_dev(dmg_handle_t h, int64_t n, const int32_t *d_node,
                                    const int32_t *d_seq, const uint8_t *d_mask, void *d_out);

/* ---- Deep Retrieval ------------------------------------------------------------------- */
/* LayerModel + RerankModel parameters (deep-retrieval/.../model/LayerModel.scala:22-39,
 * RerankModel.scala:20-41), all Double, row-major [out,in]:
 *   layer_emb (num_item + K*(D-1)) x E ; layer_w[d] K x (T+d)E ; layer_b[d] K ;
 *   rr_emb num_item x E ; rr_w E x T*E ; rr_b E ; sm_w num_item x E ; sm_b num_item. */
DMG_API int32_t dmg_dr_load(dmg_handle_t h, int32_t num_item, int32_t K, int32_t D, int32_t T,
                            int32_t E, const double *layer_emb, const double *const *layer_w,
                            const double *const *layer_b, const double *rr_emb, const double *rr_w,
                            const double *rr_b, const double *sm_w, const double *sm_b);
/* MappingOp.pathItemMapping as CSR over path keys sum_d c_d K^(D-1-d)
 * (MappingOp.scala:17-28): path_off[K^D + 1], path_items[path_off[K^D]]. */
DMG_API int32_t dmg_dr_load_paths(dmg_handle_t h, const int64_t *path_off, const int32_t *path_items);
/* CandidateSearcher.beamSearch (dr CandidateSearcher.scala:22-60): seq = B x T item indices
 * (-1 padding); out_paths B x beam x D, out_probs B x beam, out_counts B. */
DMG_API int32_t dmg_dr_beam_search(dmg_handle_t h, int32_t B, const int32_t *seq, int32_t beam,
                                   int32_t *out_paths, double *out_probs, int32_t *out_counts);
/* DeepRetrieval.recommend (DeepRetrieval.scala:26-46): beam search -> path items -> rerank
 * -> stable sort desc -> topk.  out_items = item indices (map with idItemMapping on host). */
DMG_API int32_t dmg_dr_retrieve(dmg_handle_t h, int32_t B, const int32_t *seq, int32_t beam,
                                int32_t topk, int32_t *out_items, double *out_scores,
                                int32_t *out_counts);

/* Deep Retrieval training.  dmg_dr_load_item_paths: itemPathMapping (item index -> its P = numPathPerItem paths,
 * [num_item][P][D] node indices in [0, K)), kept on the device.
 * dmg_dr_train_step = the body of LocalOptimizer.optimize's mini-batch loop
 * (deep-retrieval/.../optim/LocalOptimizer.scala:62-84): n samples (seq[n*T] item indices, -1 padding; target[n] item index).
 *   layer model: MiniBatch.transformLayerData (dataset/MiniBatch.scala:19-50) -> LayerModel forward -> CrossEntropyLayer ->
 *     backward -> syncGradients over `parallelism` thread chunks (:139-187; 1 = one chunk) -> Adam (eps 1e-8), step_t 1-based;
 *   rerank model (rerank_step_t >= 1; 0 = epoch > reRankStoppingEpoch, skipped): RerankModel forward -> SampledSoftmaxLoss
 *     (scalann/.../nn/SampledSoftmaxLoss.scala:49-153, batchMode = false; its Adam over the softmax weights / biases, eps 1e-7,
 *     on gradients that are never zeroed: nn/mixin/ParameterOptimizer.scala:28-88) -> backward -> Adam.
 *   sampled[n*(num_sampled+1)] = the reference's `sampledValues` (positive first); NULL = SampledSoftmaxLoss.uniformSampler on
 *     the device (distinct uniform negatives != positive, ascending; counter-based generator seeded by `seed`).
 *   apply = 0 leaves the parameters alone and the gradients in place for dmg_dr_download(which = 1).
 * out_layer_loss[D], out_rerank_loss (NaN when skipped).  All Double. */
DMG_API int32_t dmg_dr_load_item_paths(dmg_handle_t h, int32_t P, const int32_t *item_paths);
DMG_API int32_t dmg_dr_train_step(dmg_handle_t h, int32_t n, const int32_t *seq, const int32_t *target,
                                  const int32_t *sampled, int32_t num_sampled, uint64_t seed, double lr,
                                  int32_t step_t, int32_t rerank_step_t, int32_t parallelism, int32_t apply,
                                  double *out_layer_loss, double *out_rerank_loss);
/* Parameters (which = 0) or gradients (which = 1) back to the host in dmg_dr_load's layout; any pointer may be NULL
 * (what LayerModel / RerankModel.getParameters and DeepRetrieval.saveModel read). */
DMG_API int32_t dmg_dr_download(dmg_handle_t h, int32_t which, double *layer_emb, double *const *layer_w,
                                double *const *layer_b, double *rr_emb, double *rr_w, double *rr_b,
                                double *sm_w, double *sm_b);

/* k-means tree rebuild: RecursiveCluster.run, clusterType = "kmeans"
 * (tdm/.../cluster/RecursiveCluster.scala:34-214) without the file: recursive balanced bisection of n points (emb[n*E] Double,
 * row-major) by 2-means (best of `iters` = clusterIterNum runs; smile-core 2.6.0 KMeans.fit restated, counter-based generator
 * keyed by `seed`), squaredDistance to the first centroid, Utils.argPartition at the median; out_codes[n] = the node code of
 * every point (TreeBuilder.build flattens and writes them: dismember_b200/formats/tree_file.py build_tree). */
DMG_API int32_t dmg_kmeans_tree(dmg_handle_t h, int32_t n, int32_t E, const double *emb, int32_t iters,
                                uint64_t seed, int32_t *out_codes);

/* ---- training ------------------------------------------------------------------------- */
/* One step of LocalOptimizer.optimize on an already expanded batch
 * (tdm/.../optim/LocalOptimizer.scala:58-120,139-187; otm/.../optim/LocalOptimizer.scala:73-80):
 * zeroGradParameters, DIN forward, BCECriterionWithLogits (mean), backward with scatter-add
 * into the node table, dense Adam over the whole flat vector (scalann/.../optim/Adam.scala:19-73,
 * beta 0.9/0.999, eps 1e-8).  rows x (node, seq[T], label); step_t = 1-based timestep.
 * out_loss: one value of the loaded dtype. */
DMG_API int32_t dmg_train_step(dmg_handle_t h, int64_t rows, const int32_t *node, const int32_t *seq,
                               const int32_t *mask_flat, int64_t n_mask, const void *labels,
                               double lr, int32_t step_t, void *out_loss);
/* Device-buffer variant of dmg_train_step: rows already on the device (d_mask = rows x T mask bytes, nullable;
 * d_labels / d_out_loss of the loaded dtype), no copy and no synchronisation inside; index errors as in
 * dmg_score_pairs_dev. */
DMG_API int32_t dmg_train_step_dev(dmg_handle_t h, int64_t rows, const int32_t *d_node, const int32_t *d_seq,
                                   const uint8_t *d_mask, const void *d_labels, double lr, int32_t step_t,
                                   void *d_out_loss);
/* forward + backward only: gradient of the compact vector (testing / syncGradients). */
DMG_API int32_t dmg_din_gradients(dmg_handle_t h, int64_t rows, const int32_t *node,
                                  const int32_t *seq, const int32_t *mask_flat, int64_t n_mask,
                                  const void *labels, void *out_loss, void *out_grad, int64_t n_grad);
/* NegativeSampler.sample + MiniBatch.convert (tdm/.../utils/NegativeSampler.scala:76-158,
 * tdm/.../dataset/MiniBatch.scala:49-88): per target item the ancestor positives and
 * layer_neg[l] negatives per level >= start_level, ascending code order per level (BitSet.toList):
 * uniform over the level's existing codes (with_prob = 0, sampleFromUniformDistribution :146-158)
 * or drawn from the level's Node.probality weights with at most layer_neg[l] + tolerance draws and
 * the reference's uniform fallback (with_prob = 1, sampleFromCategoricalDistribution :116-144;
 * `tolerance` = the conf key sample_tolerance).  The reference seeds from nanoTime, so only the
 * distribution and the ordering are reproducible.  out arrays sized n_targets * layer_sum. */
DMG_API int32_t dmg_tdm_sample_expand(dmg_handle_t h, int32_t n_targets, const int32_t *target_items,
                                      const int32_t *item_seq, const int32_t *layer_neg,
                                      int32_t start_level, int32_t with_prob, int32_t tolerance,
                                      uint64_t seed, int32_t *out_node,
                                      int32_t *out_seq, float *out_label, int32_t *out_rows);

/* OTMTree.optimalPseudoTargets (otm/.../tree/OTMTree.scala:27-46 with computeTargets :104-129 and computeChildrenScores :131-165):
 * bottom-up pseudo targets of B users for the levels start_level + 1 .. leaf_level.  leaf_seq B x T leaf node ids (-1 = padding),
 * target CSR of leaf node ids.  Per level and user the (node id, target) list sorted by id in M slots (-1 padding):
 * out_ids / out_vals [leaf_level - start_level][B][M], out_counts [leaf_level - start_level][B], levels ascending.
 * use_mask = 0 keeps the reference's quirk of scoring the sibling tensor for both predictions (:157-161). */
DMG_API int32_t dmg_otm_pseudo_targets(dmg_handle_t h, int32_t B, const int32_t *leaf_seq, const int64_t *target_off,
                                       const int32_t *target_leaves, int32_t start_level, int32_t use_mask,
                                       int32_t M, int32_t *out_ids, double *out_vals, int32_t *out_counts);

/* ---- JTM tree learning ---------------------------------------------------------------- */
/* TreeLearning.aggregateWeights for a whole level step (jtm/.../optim/TreeLearning.scala:137-174):
 * item i currently sits under node parent_code[i] of level old_level and owns the training samples
 * [sample_off[i], sample_off[i+1]) (histories sample_seq, T item ids each, 0 = padding).  For each of
 * its 2^(level-old_level) candidate children (left to right), weight = sum over the nodes on the
 * child -> parent path (parent excluded, child first) of Tensor.sum of the logits of the item's samples
 * scored against that node, histories coded by JTMTree.idToCodeWithMask(seq, node level, hierarchical,
 * min_level) (JTMTree.scala:86-113); -1e6 for an item without samples (:160).
 * out_weights: n_items x 2^(level-old_level) floats, bit-identical to the in-order fp32 sums. */
DMG_API int32_t dmg_jtm_item_weights(dmg_handle_t h, int32_t n_items, const int64_t *sample_off,
                                     const int32_t *sample_seq, const int32_t *parent_code,
                                     int32_t old_level, int32_t level, int32_t hierarchical,
                                     int32_t min_level, int32_t use_mask, float *out_weights);

/* TreeLearning.getChildrenProjection / reBalance for one level step (jtm/.../optim/TreeLearning.scala:48-97,
 * 217-265): greedy, sequential per parent, host code inside the library (the scorer work is
 * dmg_jtm_item_weights).  parent_code[i]: node of old_level the item sits under; old_child[i]:
 * JTMTree.getAncestorAtLevel(item, level) in the current tree; weights: n_items x n_child (children left to
 * right); out_node[i]: new node on `level` (the parent code if every candidate child was already full).
 * Items of one parent are processed in array order. */
DMG_API int32_t dmg_jtm_assign_level(dmg_handle_t h, int32_t n_items, const int32_t *parent_code,
                                     const int32_t *old_child, int32_t n_child, const float *weights,
                                     int32_t max_assign, int32_t *out_node);

/* Metrics.computeMetrics (tdm/.../evaluation/Metrics.scala:5-25) for B users: rec_items B x topk (first rec_counts[u]
 * valid = k of the reference), labels as CSR; out_metrics B x 3 doubles (precision, recall, NDCG), to be summed by the
 * caller in user order like EvalResult (Evaluator.scala:62-66).  log() differs from java.lang.Math.log by <= 1 ulp. */
DMG_API int32_t dmg_eval_metrics(dmg_handle_t h, int32_t B, int32_t topk, const int32_t *rec_items,
                                 const int32_t *rec_counts, const int64_t *label_off, const int32_t *labels,
                                 double *out_metrics);

/* DeepFM scorer, the other `model.deep_model` of the TDM/JTM tasks (tdm/.../model/DeepFM.scala:11-44,
 * scalann/.../nn/FM.scala:14-44): params = the compact vector of Module.parameters()
 * [emb rows x E | W1 (T+1) x (T+1)E | b1 T+1 | W2 T+1 | b2 1], fp32.  Afterwards dmg_tdm_retrieve
 * (no mask; consumed items supported, widen_beam not) and dmg_score_pairs (mask arguments ignored)
 * score with the DeepFM graph; results are bit-identical to the oracle's restatement.  After
 * dmg_shard_init(world > 1) only this rank's rows are uploaded and dmg_shard_tdm_retrieve uses it. */
DMG_API int32_t dmg_load_deepfm_weights(dmg_handle_t h, int64_t rows, int32_t E, int32_t T,
                                        const float *params);
/* OTM's DeepModel[Double] = DeepFM (otm/src/main/scala/com/mass/otm/model/DeepFM.scala:12-48: the same
 * graph for Double), same parameter layout as doubles.  Afterwards dmg_otm_beam_search,
 * dmg_otm_beam_search_levels, dmg_otm_retrieve and dmg_score_pairs score with it (no mask input:
 * use_mask / mask arguments are ignored), level-synchronously like CandidateSearcher.batchBeamSearch
 * (otm/.../model/CandidateSearcher.scala:15-56); bit-identical to the oracle's restatement. */
DMG_API int32_t dmg_load_deepfm_weights_f64(dmg_handle_t h, int64_t rows, int32_t E, int32_t T,
                                            const double *params);

/* ---- node table sharded across the GPUs of one box ------------------------------------ */
/* The reference has no multi-device path: model replicas are per-thread clones
 * (tdm/.../optim/LocalOptimizer.scala:35-40) and users are split over threads
 * (tdm/.../evaluation/Evaluator.scala:28-37).  These entry points are the engine's layout for
 * a node table that does not fit one GPU (BASELINE.json configs 4-5): world = 2^g ranks, one
 * handle per rank / device; tree levels above g replicated, every deeper level split into
 * world contiguous code ranges (rank r owns the sub-trees under the r-th node of level g).
 * One process per GPU: rank 0 calls dmg_shard_unique_id and hands the 128 bytes to every rank
 * (any transport), every rank calls dmg_shard_init, loads the SAME tree with dmg_load_tree_tdm
 * and then its slice of the weights.  NCCL (libnccl.so.2) is resolved at dmg_shard_init. */
DMG_API int32_t dmg_shard_unique_id(void *out128, int32_t nbytes);
DMG_API int32_t dmg_shard_init(dmg_handle_t h, int32_t world, int32_t rank, const void *unique_id128);
/* Same values as dmg_init_din_weights on the unsharded table (counter-based generator). */
DMG_API int32_t dmg_shard_init_din_weights(dmg_handle_t h, int64_t rows_global, int32_t E, int32_t T,
                                           uint64_t seed);
/* params = the FULL compact vector of Module.parameters() (Graph.scala:37); only the rows this
 * rank owns are uploaded. fp32 (TDM/JTM scorer). */
DMG_API int32_t dmg_shard_load_din_weights(dmg_handle_t h, int64_t rows_global, int32_t E, int32_t T,
                                           const float *params);
DMG_API int32_t dmg_shard_info(dmg_handle_t h, int64_t *local_rows, int64_t *global_rows,
                                int64_t *exchanged_rows);
/* TDM.recommend (tdm/.../model/TDM.scala:17-22, Recommender.scala:18-107) for this rank's B
 * users; collective: every rank calls it with the same B, beam, topk.  Per level the candidate
 * (slot, code) pairs go to the owners of the codes over NCCL send/recv, scores come back
 * (12 bytes per remote candidate); strict fp32 arithmetic, results bit-identical to
 * dmg_tdm_retrieve on the unsharded table. */
DMG_API int32_t dmg_shard_tdm_retrieve(dmg_handle_t h, int32_t B, const int32_t *item_seq,
                                       int32_t beam, int32_t topk, int32_t use_mask,
                                       int32_t *out_items, float *out_logits, int32_t *out_counts);

/* dmg_jtm_item_weights over the sharded table (BASELINE config 4): collective; every rank passes ITS OWN items (any
 * split of the catalogue) with the same old_level / level / flags; the (sample, node) scorer rows go to the owners of
 * the nodes like retrieval candidates.  Same values as dmg_jtm_item_weights on the unsharded table. */
DMG_API int32_t dmg_shard_jtm_item_weights(dmg_handle_t h, int32_t n_items, const int64_t *sample_off,
                                           const int32_t *sample_seq, const int32_t *parent_code,
                                           int32_t old_level, int32_t level, int32_t hierarchical,
                                           int32_t min_level, int32_t use_mask, float *out_weights);

/* Deep Retrieval with the item-indexed tables (layer-embedding item rows, rerank embedding, softmax weights / biases)
 * split by contiguous item-id range over the ranks of dmg_shard_init (BASELINE config 5); the K (D-1) path-node rows
 * and every Linear are replicated.  dmg_shard_dr_load takes the WHOLE tables (same arguments as dmg_dr_load) and uploads
 * this rank's range; dmg_dr_load_paths as usual (replicated).  dmg_shard_dr_retrieve = DeepRetrieval.recommend
 * (deep-retrieval/.../model/DeepRetrieval.scala:26-46) for this rank's B users, collective: history rows by integer
 * all-reduce, beam search locally, rerank candidates scored by the owners of the items (12 B out, 8 B back per
 * candidate).  Results are bit-identical to dmg_dr_retrieve on the whole tables. */
DMG_API int32_t dmg_shard_dr_load(dmg_handle_t h, int32_t num_item, int32_t K, int32_t D, int32_t T, int32_t E,
                                  const double *layer_emb, const double *const *layer_w,
                                  const double *const *layer_b, const double *rr_emb, const double *rr_w,
                                  const double *rr_b, const double *sm_w, const double *sm_b);
DMG_API int32_t dmg_shard_dr_retrieve(dmg_handle_t h, int32_t B, const int32_t *seq, int32_t beam,
                                      int32_t topk, int32_t *out_items, double *out_scores,
                                      int32_t *out_counts);

/* Training step on the SHARDED node table (tdm/.../optim/LocalOptimizer.scala:139-187): collective; every rank passes
 * ITS rows of the mini-batch with GLOBAL node codes (node[rows], seq[rows*T], -1 = padding).  Embedding rows are fetched from their
 * owners per occurrence, forward / BCE (mean over the GLOBAL batch) / backward run locally, the embedding gradients are applied
 * by the row owners, and only the dense scorer weights (3 E^2 + 2 E + 1 scalars) plus the replicated top rows are all-reduced;
 * every rank then runs the dense Adam over its shard.  Float model.  out_loss: mean loss of the global batch. */
DMG_API int32_t dmg_shard_train_step(dmg_handle_t h, int64_t rows, const int32_t *node, const int32_t *seq,
                                     const int32_t *mask_flat, int64_t n_mask, const float *labels, double lr,
                                     int32_t step_t, float *out_loss);

/* Synthetic Deep Retrieval model generated on the device (benchmarks, BASELINE config 5): Tensor.randn(0, 0.05) of every
 * table as counter-based values of the GLOBAL element index, and J hashed paths per item folded into MappingOp.pathItemMapping's
 * shape (one item per path, MappingOp.scala:23-28) as the CSR over the K^D path keys.  On a handle with dmg_shard_init the
 * item-indexed tables hold this rank's item range (then dmg_shard_dr_retrieve), otherwise the whole tables (dmg_dr_retrieve);
 * both hold the same values, so the two give bit-identical results. */
DMG_API int32_t dmg_dr_init_synthetic(dmg_handle_t h, int32_t num_item, int32_t K, int32_t D, int32_t T, int32_t E,
                                      int32_t J, uint64_t seed);

/* Data-parallel training step over the replicas of one box: LocalOptimizer.trainBatch / syncGradients
 * (tdm/.../optim/LocalOptimizer.scala:139-187) with GPUs in the place of threads.  Every rank holds the whole model (same
 * weights, loaded with dmg_load_din_weights / dmg_init_din_weights after dmg_shard_init, which provides the communicator),
 * passes ITS rows of the mini-batch; gradients are averaged with one ncclAllReduce(ncclAvg) and every rank applies the same
 * dense Adam step.  Collective.  out_loss: mean loss of this rank's rows. */
DMG_API int32_t dmg_dp_train_step(dmg_handle_t h, int64_t rows, const int32_t *node, const int32_t *seq,
                                  const int32_t *mask_flat, int64_t n_mask, const void *labels, double lr,
                                  int32_t step_t, void *out_loss);

#ifdef __cplusplus
}
#endif
#endif /* DISMEMBER_GPU_H */
