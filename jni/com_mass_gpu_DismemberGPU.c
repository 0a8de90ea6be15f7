/*
 * jni/com_mass_gpu_DismemberGPU.c -- thin JNI -> C-ABI shim (reference-side binding).
 *
 * Static native methods of `object com.mass.gpu.DismemberGPU` (see INTEGRATION.md for the Scala
 * declaration), in the static "array + length" style the code base already uses for MKL
 * (com.intel.analytics.bigdl.mkl.MKL called from
 * scalann/src/main/scala/com/mass/scalann/tensor/TensorNumeric.scala:217-465).
 * It cannot be compiled in this image (no JDK => no jni.h); it is kept free of logic so that
 * what is tested through the ctypes binding (dismember_b200/_capi.py) is what Scala would run:
 * every function pins the primitive arrays, forwards to ONE dmg_* call and turns a non-zero
 * status into a RuntimeException / ArrayIndexOutOfBoundsException carrying dmg_last_error().
 *
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../include \
 *       com_mass_gpu_DismemberGPU.c -L../dismember_b200 -ldismember_gpu -o libdismember_jni.so
 */
#include <jni.h>
#include <stdint.h>

#include "dismember_gpu.h"

static void throw_status(JNIEnv *env, dmg_handle_t h, int32_t rc)
{
    const char *cls = rc == DMG_ERR_INDEX ? "java/lang/ArrayIndexOutOfBoundsException"
                    : rc == DMG_ERR_INVALID_ARG ? "java/lang/IllegalArgumentException"
                    : "java/lang/RuntimeException";
    (*env)->ThrowNew(env, (*env)->FindClass(env, cls), dmg_last_error(h));
}

#define H(handle) ((dmg_handle_t)(intptr_t)(handle))
#define PIN(arr) ((arr) ? (*env)->GetPrimitiveArrayCritical(env, (arr), NULL) : NULL)
#define UNPIN(arr, p, mode) do { if (arr) (*env)->ReleasePrimitiveArrayCritical(env, (arr), (p), (mode)); } while (0)

JNIEXPORT jlong JNICALL Java_com_mass_gpu_DismemberGPU_00024_create(JNIEnv *env, jobject self, jint device)
{
    dmg_handle_t h = NULL;
    int32_t rc = dmg_create(device, &h);
    if (rc) { throw_status(env, NULL, rc); return 0; }
    return (jlong)(intptr_t)h;
}

JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_destroy(JNIEnv *env, jobject self, jlong handle)
{
    dmg_destroy(H(handle));
}

/* One handle per evaluator / serving thread over ONE copy of the tables: the GPU form of model.cloneModule() sharing the
 * weight storage (tdm/.../optim/LocalOptimizer.scala:35-40, tdm/.../evaluation/Evaluator.scala:29-37). */
JNIEXPORT jlong JNICALL Java_com_mass_gpu_DismemberGPU_00024_cloneHandle(JNIEnv *env, jobject self, jlong handle)
{
    dmg_handle_t h = NULL;
    int32_t rc = dmg_clone(H(handle), &h);
    if (rc) { throw_status(env, H(handle), rc); return 0; }
    return (jlong)(intptr_t)h;
}

/* 0 = strict fp32 chains, 1 = tensor-core scorer with certified cuts (same ids and logit bits). */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_setArithmetic(JNIEnv *env, jobject self, jlong handle, jint mode)
{
    int32_t rc = dmg_set_arithmetic(H(handle), mode);
    if (rc) throw_status(env, H(handle), rc);
}

/* TDMOp.initTree: arrays built from DistTree's maps (codeNodeMap, idCodeMap). */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_loadTreeTdm(
    JNIEnv *env, jobject self, jlong handle, jint maxLevel, jintArray codes, jintArray nodeIds, jbyteArray isLeaf,
    jintArray leafIds, jintArray leafCodes)
{
    jsize n = (*env)->GetArrayLength(env, codes), m = (*env)->GetArrayLength(env, leafIds);
    void *c = PIN(codes), *i = PIN(nodeIds), *l = PIN(isLeaf), *li = PIN(leafIds), *lc = PIN(leafCodes);
    int32_t rc = dmg_load_tree_tdm(H(handle), maxLevel, n, c, i, l, m, li, lc);
    UNPIN(leafCodes, lc, JNI_ABORT); UNPIN(leafIds, li, JNI_ABORT); UNPIN(isLeaf, l, JNI_ABORT);
    UNPIN(nodeIds, i, JNI_ABORT); UNPIN(codes, c, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_loadTreeComplete(
    JNIEnv *env, jobject self, jlong handle, jint leafLevel, jintArray itemIds, jintArray leafIds)
{
    jsize n = (*env)->GetArrayLength(env, itemIds);
    void *a = PIN(itemIds), *b = PIN(leafIds);
    int32_t rc = dmg_load_tree_complete(H(handle), leafLevel, n, a, b);
    UNPIN(leafIds, b, JNI_ABORT); UNPIN(itemIds, a, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

/* model.adjustParameters()._1.storage().array() : Array[Float] (tdm/jtm) */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_loadDinWeightsFloat(
    JNIEnv *env, jobject self, jlong handle, jlong rows, jint embedSize, jint seqLen, jfloatArray params)
{
    void *p = PIN(params);
    int32_t rc = dmg_load_din_weights(H(handle), DMG_F32, rows, embedSize, seqLen, p);
    UNPIN(params, p, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

/* : Array[Double] (otm) */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_loadDinWeightsDouble(
    JNIEnv *env, jobject self, jlong handle, jlong rows, jint embedSize, jint seqLen, jdoubleArray params)
{
    void *p = PIN(params);
    int32_t rc = dmg_load_din_weights(H(handle), DMG_F64, rows, embedSize, seqLen, p);
    UNPIN(params, p, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

/* Recommender.recommendItems for a batch; consumedOff may be null. */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_tdmRetrieve(
    JNIEnv *env, jobject self, jlong handle, jint batch, jintArray itemSeq, jint beam, jint topk, jboolean useMask,
    jlongArray consumedOff, jintArray consumed, jboolean widenBeam, jintArray outItems, jfloatArray outLogits,
    jintArray outCounts)
{
    void *s = PIN(itemSeq), *co = PIN(consumedOff), *cc = PIN(consumed), *oi = PIN(outItems), *ol = PIN(outLogits),
         *oc = PIN(outCounts);
    int32_t rc = dmg_tdm_retrieve(H(handle), batch, s, beam, topk, useMask, co, cc, widenBeam, oi, ol, oc);
    UNPIN(outCounts, oc, 0); UNPIN(outLogits, ol, 0); UNPIN(outItems, oi, 0);
    UNPIN(consumed, cc, JNI_ABORT); UNPIN(consumedOff, co, JNI_ABORT); UNPIN(itemSeq, s, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

/* CandidateSearcher.batchBeamSearch */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_otmBeamSearch(
    JNIEnv *env, jobject self, jlong handle, jint batch, jintArray leafSeq, jint beam, jboolean useMask,
    jintArray outIds, jdoubleArray outScores, jintArray outCounts)
{
    void *s = PIN(leafSeq), *oi = PIN(outIds), *os = PIN(outScores), *oc = PIN(outCounts);
    int32_t rc = dmg_otm_beam_search(H(handle), batch, s, beam, useMask, oi, os, oc);
    UNPIN(outCounts, oc, 0); UNPIN(outScores, os, 0); UNPIN(outIds, oi, 0); UNPIN(leafSeq, s, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

/* OTM.recommend for a batch */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_otmRetrieve(
    JNIEnv *env, jobject self, jlong handle, jint batch, jintArray leafSeq, jint beam, jint topk, jboolean useMask,
    jintArray outItems, jdoubleArray outScores, jintArray outCounts)
{
    void *s = PIN(leafSeq), *oi = PIN(outItems), *os = PIN(outScores), *oc = PIN(outCounts);
    int32_t rc = dmg_otm_retrieve(H(handle), batch, s, beam, topk, useMask, oi, os, oc);
    UNPIN(outCounts, oc, 0); UNPIN(outScores, os, 0); UNPIN(outItems, oi, 0); UNPIN(leafSeq, s, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

/* model.forward(Table(item, seq, mask)) : Float model */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_scorePairsFloat(
    JNIEnv *env, jobject self, jlong handle, jintArray node, jintArray seq, jintArray mask, jfloatArray out)
{
    jsize n = (*env)->GetArrayLength(env, node), nm = mask ? (*env)->GetArrayLength(env, mask) : 0;
    void *a = PIN(node), *b = PIN(seq), *c = PIN(mask), *o = PIN(out);
    int32_t rc = dmg_score_pairs(H(handle), n, a, b, c, nm, o);
    UNPIN(out, o, 0); UNPIN(mask, c, JNI_ABORT); UNPIN(seq, b, JNI_ABORT); UNPIN(node, a, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_scorePairsDouble(
    JNIEnv *env, jobject self, jlong handle, jintArray node, jintArray seq, jintArray mask, jdoubleArray out)
{
    jsize n = (*env)->GetArrayLength(env, node), nm = mask ? (*env)->GetArrayLength(env, mask) : 0;
    void *a = PIN(node), *b = PIN(seq), *c = PIN(mask), *o = PIN(out);
    int32_t rc = dmg_score_pairs(H(handle), n, a, b, c, nm, o);
    UNPIN(out, o, 0); UNPIN(mask, c, JNI_ABORT); UNPIN(seq, b, JNI_ABORT); UNPIN(node, a, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

/* CandidateSearcher.beamSearch (Deep Retrieval) */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_drBeamSearch(
    JNIEnv *env, jobject self, jlong handle, jint batch, jintArray seq, jint beam, jintArray outPaths,
    jdoubleArray outProbs, jintArray outCounts)
{
    void *s = PIN(seq), *op = PIN(outPaths), *opr = PIN(outProbs), *oc = PIN(outCounts);
    int32_t rc = dmg_dr_beam_search(H(handle), batch, s, beam, op, opr, oc);
    UNPIN(outCounts, oc, 0); UNPIN(outProbs, opr, 0); UNPIN(outPaths, op, 0); UNPIN(seq, s, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_drRetrieve(
    JNIEnv *env, jobject self, jlong handle, jint batch, jintArray seq, jint beam, jint topk, jintArray outItems,
    jdoubleArray outScores, jintArray outCounts)
{
    void *s = PIN(seq), *oi = PIN(outItems), *os = PIN(outScores), *oc = PIN(outCounts);
    int32_t rc = dmg_dr_retrieve(H(handle), batch, s, beam, topk, oi, os, oc);
    UNPIN(outCounts, oc, 0); UNPIN(outScores, os, 0); UNPIN(outItems, oi, 0); UNPIN(seq, s, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

/* LocalOptimizer step on an expanded batch (Float model) */
JNIEXPORT jfloat JNICALL Java_com_mass_gpu_DismemberGPU_00024_trainStepFloat(
    JNIEnv *env, jobject self, jlong handle, jintArray node, jintArray seq, jintArray mask, jfloatArray labels,
    jdouble lr, jint stepT)
{
    jsize n = (*env)->GetArrayLength(env, node), nm = mask ? (*env)->GetArrayLength(env, mask) : 0;
    float loss = 0.0f;
    void *a = PIN(node), *b = PIN(seq), *c = PIN(mask), *l = PIN(labels);
    int32_t rc = dmg_train_step(H(handle), n, a, b, c, nm, l, lr, stepT, &loss);
    UNPIN(labels, l, JNI_ABORT); UNPIN(mask, c, JNI_ABORT); UNPIN(seq, b, JNI_ABORT); UNPIN(node, a, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
    return loss;
}

/* ---- node table sharded across the GPUs of one box (dmg_shard_*) ---- */
JNIEXPORT jbyteArray JNICALL Java_com_mass_gpu_DismemberGPU_00024_shardUniqueId(JNIEnv *env, jobject self)
{
    jbyte id[128];
    if (dmg_shard_unique_id(id, 128)) { throw_status(env, NULL, DMG_ERR_UNSUPPORTED); return NULL; }
    jbyteArray out = (*env)->NewByteArray(env, 128);
    (*env)->SetByteArrayRegion(env, out, 0, 128, id);
    return out;
}

JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_shardInit(
    JNIEnv *env, jobject self, jlong handle, jint world, jint rank, jbyteArray uniqueId)
{
    void *id = PIN(uniqueId);
    int32_t rc = dmg_shard_init(H(handle), world, rank, id);
    UNPIN(uniqueId, id, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_shardLoadDinWeightsFloat(
    JNIEnv *env, jobject self, jlong handle, jlong rowsGlobal, jint embedSize, jint seqLen, jfloatArray params)
{
    void *p = PIN(params);
    int32_t rc = dmg_shard_load_din_weights(H(handle), rowsGlobal, embedSize, seqLen, p);
    UNPIN(params, p, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_shardTdmRetrieve(
    JNIEnv *env, jobject self, jlong handle, jint batch, jintArray itemSeq, jint beam, jint topk, jboolean useMask,
    jintArray outItems, jfloatArray outLogits, jintArray outCounts)
{
    void *s = PIN(itemSeq), *oi = PIN(outItems), *ol = PIN(outLogits), *oc = PIN(outCounts);
    int32_t rc = dmg_shard_tdm_retrieve(H(handle), batch, s, beam, topk, useMask ? 1 : 0, oi, ol, oc);
    UNPIN(outCounts, oc, 0); UNPIN(outLogits, ol, 0); UNPIN(outItems, oi, 0); UNPIN(itemSeq, s, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

/* TreeLearning.reBalance for one level step (host code inside the library) */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_jtmAssignLevel(
    JNIEnv *env, jobject self, jlong handle, jintArray parentCode, jintArray oldChild, jint nChild, jfloatArray weights,
    jint maxAssign, jintArray outNode)
{
    jsize n = (*env)->GetArrayLength(env, parentCode);
    void *p = PIN(parentCode), *o = PIN(oldChild), *w = PIN(weights), *out = PIN(outNode);
    int32_t rc = dmg_jtm_assign_level(H(handle), n, p, o, nChild, w, maxAssign, out);
    UNPIN(outNode, out, 0); UNPIN(weights, w, JNI_ABORT); UNPIN(oldChild, o, JNI_ABORT); UNPIN(parentCode, p, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

/* TreeLearning.aggregateWeights for a level step; out: nItems x 2^(level - oldLevel) */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_jtmItemWeights(
    JNIEnv *env, jobject self, jlong handle, jlongArray sampleOff, jintArray sampleSeq, jintArray parentCode, jint oldLevel,
    jint level, jboolean hierarchical, jint minLevel, jboolean useMask, jfloatArray outWeights)
{
    jsize n = (*env)->GetArrayLength(env, parentCode);
    void *o = PIN(sampleOff), *s = PIN(sampleSeq), *p = PIN(parentCode), *w = PIN(outWeights);
    int32_t rc = dmg_jtm_item_weights(H(handle), n, o, s, p, oldLevel, level, hierarchical ? 1 : 0, minLevel, useMask ? 1 : 0, w);
    UNPIN(outWeights, w, 0); UNPIN(parentCode, p, JNI_ABORT); UNPIN(sampleSeq, s, JNI_ABORT); UNPIN(sampleOff, o, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

/* NegativeSampler.sample + MiniBatch.convert on the device; returns the number of rows written */
JNIEXPORT jint JNICALL Java_com_mass_gpu_DismemberGPU_00024_tdmSampleExpand(
    JNIEnv *env, jobject self, jlong handle, jintArray targets, jintArray itemSeq, jintArray layerNeg, jint startLevel, jlong seed,
    jintArray outNode, jintArray outSeq, jfloatArray outLabel)
{
    jsize n = (*env)->GetArrayLength(env, targets);
    int32_t rows = 0;
    void *t = PIN(targets), *s = PIN(itemSeq), *l = PIN(layerNeg), *on = PIN(outNode), *os = PIN(outSeq), *ol = PIN(outLabel);
    int32_t rc = dmg_tdm_sample_expand(H(handle), n, t, s, l, startLevel, (uint64_t)seed, on, os, ol, &rows);
    UNPIN(outLabel, ol, 0); UNPIN(outSeq, os, 0); UNPIN(outNode, on, 0);
    UNPIN(layerNeg, l, JNI_ABORT); UNPIN(itemSeq, s, JNI_ABORT); UNPIN(targets, t, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
    return rows;
}

/* Module.parameters() back to the JVM (Serialization.saveModel) */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_downloadDinWeightsFloat(JNIEnv *env, jobject self, jlong handle, jfloatArray params)
{
    jsize n = (*env)->GetArrayLength(env, params);
    void *p = PIN(params);
    int32_t rc = dmg_download_din_weights(H(handle), p, n);
    UNPIN(params, p, 0);
    if (rc) throw_status(env, H(handle), rc);
}

/* model.deep_model = "DeepFM" (TDM.scala:43) */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_loadDeepFmWeightsFloat(
    JNIEnv *env, jobject self, jlong handle, jlong rows, jint embedSize, jint seqLen, jfloatArray params)
{
    void *p = PIN(params);
    int32_t rc = dmg_load_deepfm_weights(H(handle), rows, embedSize, seqLen, p);
    UNPIN(params, p, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

/* Metrics.computeMetrics per user; out: batch x 3 (precision, recall, ndcg) */
/* OTM with deepModel = "DeepFM": DeepModel[Double] (otm/.../model/DeepFM.scala:12-48) */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_loadDeepFmWeightsDouble(
    JNIEnv *env, jobject self, jlong handle, jlong rows, jint embedSize, jint seqLen, jdoubleArray params)
{
    jdouble *p = PIN(params);
    int32_t rc = dmg_load_deepfm_weights_f64(H(handle), rows, embedSize, seqLen, p);
    UNPIN(params, p, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_evalMetrics(
    JNIEnv *env, jobject self, jlong handle, jint batch, jint topk, jintArray recItems, jintArray recCounts, jlongArray labelOff,
    jintArray labels, jdoubleArray out)
{
    void *r = PIN(recItems), *c = PIN(recCounts), *o = PIN(labelOff), *l = PIN(labels), *m = PIN(out);
    int32_t rc = dmg_eval_metrics(H(handle), batch, topk, r, c, o, l, m);
    UNPIN(out, m, 0); UNPIN(labels, l, JNI_ABORT); UNPIN(labelOff, o, JNI_ABORT); UNPIN(recCounts, c, JNI_ABORT); UNPIN(recItems, r, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}
