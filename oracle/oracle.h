/*
 * oracle/oracle.h -- C interface of the CPU oracle (TEST INFRASTRUCTURE).
 * See oracle.c for the provenance of every function.
 */
#ifndef DMG_ORACLE_H
#define DMG_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_tree orc_tree;
typedef struct orc_tdm_model orc_tdm_model;
typedef struct orc_otm_model orc_otm_model;
typedef struct orc_dr_model orc_dr_model;

orc_tree *orc_tree_create(int max_level, int64_t n_nodes, const int32_t *codes, const int32_t *node_ids,
                          const uint8_t *is_leaf, int64_t n_items, const int32_t *leaf_ids,
                          const int32_t *leaf_codes);
void orc_tree_destroy(orc_tree *t);
void orc_tdm_id_to_code(const orc_tree *t, int T, const int32_t *ids, int32_t *codes, uint8_t *masked);

/* params = compact DIN vector [emb | W_att | W1 | b1 | W2 | b2]; NOT copied. */
orc_tdm_model *orc_tdm_model_create(int64_t rows, int E, int T, const float *params);
/* params = compact DeepFM vector [emb | W1 (T+1)x(T+1)E | b1 | W2 | b2] (tdm/.../model/DeepFM.scala:11-44); NOT copied. */
orc_tdm_model *orc_tdm_deepfm_create(int64_t rows, int E, int T, const float *params);
void orc_tdm_model_destroy(orc_tdm_model *m);
orc_otm_model *orc_otm_model_create(int64_t rows, int E, int T, const double *params);
orc_otm_model *orc_otm_deepfm_create(int64_t rows, int E, int T, const double *params);   /* otm/.../model/DeepFM.scala:12-48 */
void orc_otm_model_destroy(orc_otm_model *m);

int orc_din_forward_f32_api(const orc_tdm_model *m, int64_t n, const int32_t *node, const int32_t *seq,
                            const int32_t *mask_flat, int64_t n_mask, float *out);
int orc_din_forward_f64_api(const orc_otm_model *m, int64_t n, const int32_t *node, const int32_t *seq,
                            const int32_t *mask_flat, int64_t n_mask, double *out);

int orc_tdm_recommend_raw(const orc_tree *t, const orc_tdm_model *m, const int32_t *seq_ids, int beam,
                          int use_mask, const int32_t *consumed, int n_consumed, int32_t *out_items,
                          float *out_logits, int cap);
int orc_tdm_recommend(const orc_tree *t, const orc_tdm_model *m, const int32_t *seq_ids, int beam, int topk,
                      int use_mask, const int32_t *consumed, int n_consumed, int widen_beam,
                      int32_t *out_items, float *out_logits, double *out_prob);
int orc_tdm_retrieve_batch(const orc_tree *t, const orc_tdm_model *m, int B, const int32_t *seq_ids, int beam,
                           int topk, int use_mask, const int64_t *cons_off, const int32_t *cons, int widen_beam,
                           int n_threads, int32_t *out_items, float *out_logits, int32_t *out_counts);

/* oracle_tuned.c: the re-associated, batched CPU form (a reported baseline, never a checker) */
typedef struct orc_tuned orc_tuned;
void orc_tree_view(const orc_tree *t, int *max_level, int64_t *n_codes, const uint8_t **exists, const uint8_t **is_leaf,
                   const int32_t **node_id);
int orc_tdm_model_view(const orc_tdm_model *m, int64_t *rows, int *E, int *T, const float **emb, const float **watt,
                       const float **w1, const float **b1, const float **w2, const float **b2);
orc_tuned *orc_tuned_create(const orc_tree *t, const orc_tdm_model *model, int n_threads);
void orc_tuned_destroy(orc_tuned *m);
int orc_tuned_retrieve_batch(const orc_tuned *m, int B, const int32_t *seq_ids, int beam, int topk, int use_mask,
                             int n_threads, int32_t *out_items, float *out_logits, int32_t *out_counts);

int orc_otm_beam_search(const orc_otm_model *m, const int32_t *seq, int leaf_level, int beam, int use_mask,
                        int32_t *out_ids, double *out_scores);
int orc_otm_recommend(const orc_otm_model *m, const int32_t *seq_leaf_ids, int leaf_level, int beam, int topk,
                      int use_mask, const int32_t *leaf_item, int32_t *out_items, double *out_scores,
                      double *out_prob);
int orc_otm_retrieve_batch(const orc_otm_model *m, int B, const int32_t *seq_leaf_ids, int leaf_level, int beam,
                           int topk, int use_mask, const int32_t *leaf_item, int n_threads, int32_t *out_items,
                           double *out_scores, int32_t *out_counts);

/* OTMTree.optimalPseudoTargets / computeTargets / computeChildrenScores (otm/.../tree/OTMTree.scala:27-46, 104-172): bottom-up pseudo
 * targets of B users.  seqs B x T leaf ids (-1 pad), targets CSR of leaf node ids, M >= max targets per user.  Outputs for the levels
 * start_level + 1 .. leaf_level (n_lvl = leaf_level - start_level of them, ascending): out_ids / out_vals [n_lvl][B][M], ids ascending,
 * -1 padding, out_cnt [n_lvl][B]. */
int orc_otm_pseudo_targets(const orc_otm_model *m, int B, int T, const int32_t *seqs, const int64_t *target_off,
                           const int32_t *targets, int leaf_level, int start_level, int use_mask, int M,
                           int32_t *out_ids, double *out_vals, int32_t *out_cnt);

orc_dr_model *orc_dr_model_create(int num_item, int K, int D, int T, int E, const double *layer_emb,
                                  const double *const *layer_w, const double *const *layer_b,
                                  const double *rr_emb, const double *rr_w, const double *rr_b,
                                  const double *sm_w, const double *sm_b);
void orc_dr_model_destroy(orc_dr_model *m);
int orc_dr_beam_search(const orc_dr_model *m, const int32_t *seq, int beam, int32_t *out_paths, double *out_prob);
int orc_dr_rerank(const orc_dr_model *m, const int32_t *seq, int n_cand, const int32_t *cand, double *out);
int orc_dr_recommend(const orc_dr_model *m, const int32_t *seq, int beam, int topk, const int64_t *path_off,
                     const int32_t *path_items, int32_t *out_ids, double *out_scores, double *out_prob);

int orc_din_gradients_f32(int64_t rows, int E, int T, const float *params, int64_t n, const int32_t *node,
                          const int32_t *seq, const int32_t *mask_flat, int64_t n_mask, const float *labels,
                          float *grad, float *loss);
int orc_din_gradients_f64(int64_t rows, int E, int T, const double *params, int64_t n, const int32_t *node,
                          const int32_t *seq, const int32_t *mask_flat, int64_t n_mask, const double *labels,
                          double *grad, double *loss);
int orc_deepfm_gradients_f32(int64_t rows, int E, int T, const float *params, int64_t n, const int32_t *node, const int32_t *seq,
                             const float *labels, float *grad, float *loss_out);
void orc_adam_f32(float *w, const float *g, float *s, float *r, int64_t n, double lr, int t);
void orc_adam_f64(double *w, const double *g, double *s, double *r, int64_t n, double lr, int t);

void orc_softmax_f32(int n, int dim, const float *in, float *out);
void orc_softmax_grad_f32(int n, int dim, const float *y, const float *go, float *gi);
float orc_expf_api(float x);
double orc_exp_api(double x);

/* ---- Deep Retrieval training step (oracle_dr_train.c; deep-retrieval/.../optim/LocalOptimizer.scala:58-194) ---- */
double orc_cross_entropy_f64(int64_t R, int C, const double *logits, const int32_t *target, double *grad);
int orc_dr_layer_grad(int num_item, int K, int D, int T, int E, const double *emb, const double *const *w, const double *const *b,
                      int n, const int32_t *seq, const int32_t *target, const int32_t *item_paths, int P, int parallelism,
                      double *g_emb, double *const *g_w, double *const *g_b, double *loss);
double orc_sampled_softmax_f64(int n, int E, int S, const double *u, const double *sm_w, const double *sm_b, const int32_t *sampled,
                               double *gu, double *g_sm_w, double *g_sm_b);
int orc_dr_rerank_grad(int num_item, int T, int E, const double *rr_emb, const double *rr_w, const double *rr_b, const double *sm_w,
                       const double *sm_b, int n, const int32_t *seq, const int32_t *sampled, int S, double *g_rr_emb, double *g_rr_w,
                       double *g_rr_b, double *g_sm_w, double *g_sm_b, double *loss);
void orc_adam_eps_f64(double *w, const double *g, double *s, double *r, int64_t n, double lr, double eps, int t);

/* ---- k-means tree rebuild (oracle_cluster.c; tdm/.../cluster/RecursiveCluster.scala:34-214, utils/Utils.scala:130-199) ---- */
int orc_arg_partition(double *elems, int n, int position, int32_t *indices);
int orc_kmeans_tree(int n, int E, const double *emb, int iters, uint64_t seed, int32_t *codes);

#ifdef __cplusplus
}
#endif
#endif
