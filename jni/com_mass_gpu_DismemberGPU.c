/*
 * jni/com_mass_gpu_DismemberGPU.c -- thin JNI -> C-ABI shim (reference-side binding).
 *
 * Static native methods of `object com.mass.gpu.DismemberGPU` (see INTEGRATION.md for the Scala
 * declaration), in the static "array + length" style the code base already uses for MKL
 * (com.intel.analytics.bigdl.mkl.MKL called from
 * scalann/src/main/scala/com/mass/scalann/tensor/TensorNumeric.scala:217-465).
 * It cannot be compiled in this image (no JDK => no jni.h); it is kept free of logic so that
 * what is tested through the ctypes binding (dismember_b200/_capi.py) is what Scala would run:
 * every function measures the arrays it is handed (IllegalArgumentException when one is too short for
 * what the call reads or writes), maps them with Get<Type>ArrayElements (never the critical variants: the
 * calls block on the GPU), forwards to ONE dmg_* call and turns a non-zero status into a RuntimeException /
 * ArrayIndexOutOfBoundsException carrying dmg_last_error().  tests/jni_stub/jni.h + tests/test_capi_symbols.py
 * keep it syntactically checked (gcc -fsyntax-only) although no JDK is present.
 *
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../include \
 *       com_mass_gpu_DismemberGPU.c -L../dismember_b200 -ldismember_gpu -o libdismember_jni.so
 */
#include <jni.h>
#include <stddef.h>
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
/* Every dmg_* call synchronises the GPU, and JNI forbids blocking inside a Get/ReleasePrimitiveArrayCritical region (the GC
 * locker would stall every other thread of the JVM for the length of the batch): the arrays are accessed with the ordinary
 * Get<Type>ArrayElements calls.  Inputs are released with JNI_ABORT (no copy back), outputs with 0. */
#define PIN_I(arr) ((arr) ? (void *)(*env)->GetIntArrayElements(env, (arr), NULL) : NULL)
#define PIN_F(arr) ((arr) ? (void *)(*env)->GetFloatArrayElements(env, (arr), NULL) : NULL)
#define PIN_D(arr) ((arr) ? (void *)(*env)->GetDoubleArrayElements(env, (arr), NULL) : NULL)
#define PIN_L(arr) ((arr) ? (void *)(*env)->GetLongArrayElements(env, (arr), NULL) : NULL)
#define PIN_B(arr) ((arr) ? (void *)(*env)->GetByteArrayElements(env, (arr), NULL) : NULL)
#define UNPIN_I(arr, p, mode) do { if (arr) (*env)->ReleaseIntArrayElements(env, (arr), (jint *)(p), (mode)); } while (0)
#define UNPIN_F(arr, p, mode) do { if (arr) (*env)->ReleaseFloatArrayElements(env, (arr), (jfloat *)(p), (mode)); } while (0)
#define UNPIN_D(arr, p, mode) do { if (arr) (*env)->ReleaseDoubleArrayElements(env, (arr), (jdouble *)(p), (mode)); } while (0)
#define UNPIN_L(arr, p, mode) do { if (arr) (*env)->ReleaseLongArrayElements(env, (arr), (jlong *)(p), (mode)); } while (0)
#define UNPIN_B(arr, p, mode) do { if (arr) (*env)->ReleaseByteArrayElements(env, (arr), (jbyte *)(p), (mode)); } while (0)

/* The C ABI writes batch x topk (2 x beam, rows x T ..) elements through the raw pointers: an undersized Java array would
 * corrupt the heap silently, so every array is measured first.  need() throws IllegalArgumentException and returns 0. */
static int need(JNIEnv *env, jarray arr, jlong n, const char *what)
{
    if (n < 0 || (n > 0 && !arr) || (arr && (jlong)(*env)->GetArrayLength(env, arr) < n)) {
        (*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/IllegalArgumentException"), what);
        return 0;
    }
    return 1;
}
/* seq_len of the loaded model (0 when none): the row length of every history array */
static jlong seq_len(dmg_handle_t h)
{
    int64_t rows = 0; int32_t e = 0, t = 0, dt = 0;
    return dmg_din_shape(h, &rows, &e, &t, &dt) == DMG_OK ? t : 0;
}

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
    jintArray leafIds, jintArray leafCodes, jfloatArray prob)
{
    jsize n = (*env)->GetArrayLength(env, codes), m = (*env)->GetArrayLength(env, leafIds);
    if (!need(env, nodeIds, n, "nodeIds: one per node") || !need(env, isLeaf, n, "isLeaf: one per node") ||
        !need(env, leafCodes, m, "leafCodes: one per leaf id") || (prob && !need(env, prob, n, "prob: one per node")))
        return;
    void *c = PIN_I(codes), *i = PIN_I(nodeIds), *l = PIN_B(isLeaf), *li = PIN_I(leafIds), *lc = PIN_I(leafCodes), *pr = PIN_F(prob);
    int32_t rc = dmg_load_tree_tdm(H(handle), maxLevel, n, c, i, l, m, li, lc, pr);
    UNPIN_F(prob, pr, JNI_ABORT);
    UNPIN_I(leafCodes, lc, JNI_ABORT); UNPIN_I(leafIds, li, JNI_ABORT); UNPIN_B(isLeaf, l, JNI_ABORT);
    UNPIN_I(nodeIds, i, JNI_ABORT); UNPIN_I(codes, c, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_loadTreeComplete(
    JNIEnv *env, jobject self, jlong handle, jint leafLevel, jintArray itemIds, jintArray leafIds)
{
    jsize n = (*env)->GetArrayLength(env, itemIds);
    void *a = PIN_I(itemIds), *b = PIN_I(leafIds);
    int32_t rc = dmg_load_tree_complete(H(handle), leafLevel, n, a, b);
    UNPIN_I(leafIds, b, JNI_ABORT); UNPIN_I(itemIds, a, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

/* model.adjustParameters()._1.storage().array() : Array[Float] (tdm/jtm) */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_loadDinWeightsFloat(
    JNIEnv *env, jobject self, jlong handle, jlong rows, jint embedSize, jint seqLen, jfloatArray params)
{
    void *p = PIN_F(params);
    int32_t rc = dmg_load_din_weights(H(handle), DMG_F32, rows, embedSize, seqLen, p);
    UNPIN_F(params, p, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

/* : Array[Double] (otm) */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_loadDinWeightsDouble(
    JNIEnv *env, jobject self, jlong handle, jlong rows, jint embedSize, jint seqLen, jdoubleArray params)
{
    void *p = PIN_D(params);
    int32_t rc = dmg_load_din_weights(H(handle), DMG_F64, rows, embedSize, seqLen, p);
    UNPIN_D(params, p, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

/* Recommender.recommendItems for a batch; consumedOff may be null. */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_tdmRetrieve(
    JNIEnv *env, jobject self, jlong handle, jint batch, jintArray itemSeq, jint beam, jint topk, jboolean useMask,
    jlongArray consumedOff, jintArray consumed, jboolean widenBeam, jintArray outItems, jfloatArray outLogits,
    jintArray outCounts)
{
    const jlong T = seq_len(H(handle));
    if (batch <= 0 || topk <= 0 || !need(env, itemSeq, (jlong)batch * T, "itemSeq: batch x seq_len ints") ||
        (consumedOff && !need(env, consumedOff, (jlong)batch + 1, "consumedOff: batch + 1 longs")) ||
        !need(env, outItems, (jlong)batch * topk, "outItems: batch x topk ints") ||
        !need(env, outLogits, (jlong)batch * topk, "outLogits: batch x topk floats") || !need(env, outCounts, batch, "outCounts: batch ints"))
        return;
    if (consumedOff) {                                              /* the CSR's last offset bounds the consumed-items array */
        jlong last = 0;
        (*env)->GetLongArrayRegion(env, consumedOff, batch, 1, &last);
        if (!need(env, consumed, last, "consumed: consumedOff(batch) ints")) return;
    }
    void *s = PIN_I(itemSeq), *co = PIN_L(consumedOff), *cc = PIN_I(consumed), *oi = PIN_I(outItems), *ol = PIN_F(outLogits),
         *oc = PIN_I(outCounts);
    int32_t rc = dmg_tdm_retrieve(H(handle), batch, s, beam, topk, useMask, co, cc, widenBeam, oi, ol, oc);
    UNPIN_I(outCounts, oc, 0); UNPIN_F(outLogits, ol, 0); UNPIN_I(outItems, oi, 0);
    UNPIN_I(consumed, cc, JNI_ABORT); UNPIN_L(consumedOff, co, JNI_ABORT); UNPIN_I(itemSeq, s, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

/* CandidateSearcher.batchBeamSearch */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_otmBeamSearch(
    JNIEnv *env, jobject self, jlong handle, jint batch, jintArray leafSeq, jint beam, jboolean useMask,
    jintArray outIds, jdoubleArray outScores, jintArray outCounts)
{
    const jlong T = seq_len(H(handle));
    {
        jlong width = 2; while (width * 2 <= beam) width *= 2;       /* 2 x max(beam, 2^floor(log2 beam)) entries per user */
        width = 2 * (beam > width ? beam : width);
        if (batch <= 0 || beam <= 0 || !need(env, leafSeq, (jlong)batch * T, "leafSeq: batch x seq_len ints") ||
            !need(env, outIds, batch * width, "outIds: batch x 2 max(beam, 2^floor(log2 beam)) ints") ||
            !need(env, outScores, batch * width, "outScores: batch x 2 max(beam, 2^floor(log2 beam)) doubles") ||
            !need(env, outCounts, batch, "outCounts: batch ints"))
            return;
    }
    void *s = PIN_I(leafSeq), *oi = PIN_I(outIds), *os = PIN_D(outScores), *oc = PIN_I(outCounts);
    int32_t rc = dmg_otm_beam_search(H(handle), batch, s, beam, useMask, oi, os, oc);
    UNPIN_I(outCounts, oc, 0); UNPIN_D(outScores, os, 0); UNPIN_I(outIds, oi, 0); UNPIN_I(leafSeq, s, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

/* OTM.recommend for a batch */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_otmRetrieve(
    JNIEnv *env, jobject self, jlong handle, jint batch, jintArray leafSeq, jint beam, jint topk, jboolean useMask,
    jintArray outItems, jdoubleArray outScores, jintArray outCounts)
{
    const jlong T = seq_len(H(handle));
    if (batch <= 0 || topk <= 0 || !need(env, leafSeq, (jlong)batch * T, "leafSeq: batch x seq_len ints") ||
        !need(env, outItems, (jlong)batch * topk, "outItems: batch x topk ints") ||
        !need(env, outScores, (jlong)batch * topk, "outScores: batch x topk doubles") || !need(env, outCounts, batch, "outCounts: batch ints"))
        return;
    void *s = PIN_I(leafSeq), *oi = PIN_I(outItems), *os = PIN_D(outScores), *oc = PIN_I(outCounts);
    int32_t rc = dmg_otm_retrieve(H(handle), batch, s, beam, topk, useMask, oi, os, oc);
    UNPIN_I(outCounts, oc, 0); UNPIN_D(outScores, os, 0); UNPIN_I(outItems, oi, 0); UNPIN_I(leafSeq, s, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

/* model.forward(Table(item, seq, mask)) : Float model */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_scorePairsFloat(
    JNIEnv *env, jobject self, jlong handle, jintArray node, jintArray seq, jintArray mask, jfloatArray out)
{
    jsize n = (*env)->GetArrayLength(env, node), nm = mask ? (*env)->GetArrayLength(env, mask) : 0;
    const jlong T = seq_len(H(handle));
    if (!need(env, seq, (jlong)n * T, "seq: rows x seq_len ints") || !need(env, out, n, "out: one logit per row")) return;
    void *a = PIN_I(node), *b = PIN_I(seq), *c = PIN_I(mask), *o = PIN_F(out);
    int32_t rc = dmg_score_pairs(H(handle), n, a, b, c, nm, o);
    UNPIN_F(out, o, 0); UNPIN_I(mask, c, JNI_ABORT); UNPIN_I(seq, b, JNI_ABORT); UNPIN_I(node, a, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_scorePairsDouble(
    JNIEnv *env, jobject self, jlong handle, jintArray node, jintArray seq, jintArray mask, jdoubleArray out)
{
    jsize n = (*env)->GetArrayLength(env, node), nm = mask ? (*env)->GetArrayLength(env, mask) : 0;
    const jlong T = seq_len(H(handle));
    if (!need(env, seq, (jlong)n * T, "seq: rows x seq_len ints") || !need(env, out, n, "out: one logit per row")) return;
    void *a = PIN_I(node), *b = PIN_I(seq), *c = PIN_I(mask), *o = PIN_D(out);
    int32_t rc = dmg_score_pairs(H(handle), n, a, b, c, nm, o);
    UNPIN_D(out, o, 0); UNPIN_I(mask, c, JNI_ABORT); UNPIN_I(seq, b, JNI_ABORT); UNPIN_I(node, a, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

/* CandidateSearcher.beamSearch (Deep Retrieval) */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_drBeamSearch(
    JNIEnv *env, jobject self, jlong handle, jint batch, jintArray seq, jint beam, jintArray outPaths,
    jdoubleArray outProbs, jintArray outCounts)
{
    void *s = PIN_I(seq), *op = PIN_I(outPaths), *opr = PIN_D(outProbs), *oc = PIN_I(outCounts);
    int32_t rc = dmg_dr_beam_search(H(handle), batch, s, beam, op, opr, oc);
    UNPIN_I(outCounts, oc, 0); UNPIN_D(outProbs, opr, 0); UNPIN_I(outPaths, op, 0); UNPIN_I(seq, s, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_drRetrieve(
    JNIEnv *env, jobject self, jlong handle, jint batch, jintArray seq, jint beam, jint topk, jintArray outItems,
    jdoubleArray outScores, jintArray outCounts)
{
    if (batch <= 0 || topk <= 0 || !need(env, outItems, (jlong)batch * topk, "outItems: batch x topk ints") ||
        !need(env, outScores, (jlong)batch * topk, "outScores: batch x topk doubles") || !need(env, outCounts, batch, "outCounts: batch ints"))
        return;
    void *s = PIN_I(seq), *oi = PIN_I(outItems), *os = PIN_D(outScores), *oc = PIN_I(outCounts);
    int32_t rc = dmg_dr_retrieve(H(handle), batch, s, beam, topk, oi, os, oc);
    UNPIN_I(outCounts, oc, 0); UNPIN_D(outScores, os, 0); UNPIN_I(outItems, oi, 0); UNPIN_I(seq, s, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

/* how the synchronous calls wait for their batch: 0 spin, 1 sleep (dmg_set_sync_mode) */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_setSyncMode(JNIEnv *env, jobject self, jlong handle, jint mode)
{
    int32_t rc = dmg_set_sync_mode(H(handle), mode);
    if (rc) throw_status(env, H(handle), rc);
}

/* RecursiveCluster.run (kmeans): node code per embedding row */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_kmeansTree(
    JNIEnv *env, jobject self, jlong handle, jint n, jint embedSize, jdoubleArray embeddings, jint clusterIterNum, jlong seed, jintArray outCodes)
{
    if (n <= 0 || embedSize <= 0 || !need(env, embeddings, (jlong)n * embedSize, "embeddings: n x embedSize doubles") ||
        !need(env, outCodes, n, "outCodes: n ints"))
        return;
    double *e = PIN_D(embeddings);
    void *o = PIN_I(outCodes);
    int32_t rc = dmg_kmeans_tree(H(handle), n, embedSize, e, clusterIterNum, (uint64_t)seed, o);
    UNPIN_I(outCodes, o, 0); UNPIN_D(embeddings, e, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

/* Deep Retrieval LocalOptimizer: itemPathMapping upload, then one mini-batch iteration (layer model + rerank model);
 * outLosses = numLayer layer losses followed by the rerank loss (NaN when reRankStepT == 0) */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_drLoadItemPaths(
    JNIEnv *env, jobject self, jlong handle, jint numPathPerItem, jintArray itemPaths)
{
    if (numPathPerItem <= 0 || !need(env, itemPaths, numPathPerItem, "itemPaths: numItem x numPathPerItem x numLayer ints")) return;
    void *p = PIN_I(itemPaths);
    int32_t rc = dmg_dr_load_item_paths(H(handle), numPathPerItem, p);
    UNPIN_I(itemPaths, p, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_drTrainStep(
    JNIEnv *env, jobject self, jlong handle, jint batch, jint seqLen, jint numLayer, jintArray seq, jintArray target, jint numSampled,
    jlong seed, jdouble lr, jint stepT, jint reRankStepT, jint parallelism, jdoubleArray outLosses)
{
    if (batch <= 0 || seqLen <= 0 || numLayer <= 0 || !need(env, seq, (jlong)batch * seqLen, "seq: batch x seq_len ints") ||
        !need(env, target, batch, "target: batch ints") || !need(env, outLosses, (jlong)numLayer + 1, "outLosses: numLayer + 1 doubles"))
        return;
    void *s = PIN_I(seq), *t = PIN_I(target);
    double *o = PIN_D(outLosses);
    int32_t rc = dmg_dr_train_step(H(handle), batch, s, t, NULL, numSampled, (uint64_t)seed, lr, stepT, reRankStepT, parallelism, 1, o, o + numLayer);
    UNPIN_D(outLosses, o, 0); UNPIN_I(target, t, JNI_ABORT); UNPIN_I(seq, s, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

/* OTMTree.optimalPseudoTargets for a mini-batch; outputs [leafLevel - startLevel][batch][maxLabels] */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_otmPseudoTargets(
    JNIEnv *env, jobject self, jlong handle, jint batch, jintArray leafSeq, jlongArray targetOff, jintArray targets, jint startLevel,
    jint leafLevel, jboolean useMask, jint maxLabels, jintArray outIds, jdoubleArray outVals, jintArray outCounts)
{
    const jlong T = seq_len(H(handle)), nl = (jlong)leafLevel - startLevel;
    jlong last = 0;
    if (batch <= 0 || nl <= 0 || maxLabels <= 0 || !need(env, leafSeq, (jlong)batch * T, "leafSeq: batch x seq_len ints") ||
        !need(env, targetOff, (jlong)batch + 1, "targetOff: batch + 1 longs"))
        return;
    (*env)->GetLongArrayRegion(env, targetOff, batch, 1, &last);
    if (!need(env, targets, last, "targets: targetOff(batch) ints") || !need(env, outIds, nl * batch * maxLabels, "outIds: levels x batch x maxLabels ints") ||
        !need(env, outVals, nl * batch * maxLabels, "outVals: levels x batch x maxLabels doubles") || !need(env, outCounts, nl * batch, "outCounts: levels x batch ints"))
        return;
    void *s = PIN_I(leafSeq), *o = PIN_L(targetOff), *t = PIN_I(targets), *oi = PIN_I(outIds), *ov = PIN_D(outVals), *oc = PIN_I(outCounts);
    int32_t rc = dmg_otm_pseudo_targets(H(handle), batch, s, o, t, startLevel, useMask ? 1 : 0, maxLabels, oi, ov, oc);
    UNPIN_I(outCounts, oc, 0); UNPIN_D(outVals, ov, 0); UNPIN_I(outIds, oi, 0);
    UNPIN_I(targets, t, JNI_ABORT); UNPIN_L(targetOff, o, JNI_ABORT); UNPIN_I(leafSeq, s, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

/* LocalOptimizer step on an expanded batch (Float model) */
JNIEXPORT jfloat JNICALL Java_com_mass_gpu_DismemberGPU_00024_trainStepFloat(
    JNIEnv *env, jobject self, jlong handle, jintArray node, jintArray seq, jintArray mask, jfloatArray labels,
    jdouble lr, jint stepT)
{
    jsize n = (*env)->GetArrayLength(env, node), nm = mask ? (*env)->GetArrayLength(env, mask) : 0;
    float loss = 0.0f;
    const jlong T = seq_len(H(handle));
    if (!need(env, seq, (jlong)n * T, "seq: rows x seq_len ints") || !need(env, labels, n, "labels: one per row")) return 0.0f;
    void *a = PIN_I(node), *b = PIN_I(seq), *c = PIN_I(mask), *l = PIN_F(labels);
    int32_t rc = dmg_train_step(H(handle), n, a, b, c, nm, l, lr, stepT, &loss);
    UNPIN_F(labels, l, JNI_ABORT); UNPIN_I(mask, c, JNI_ABORT); UNPIN_I(seq, b, JNI_ABORT); UNPIN_I(node, a, JNI_ABORT);
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
    void *id = PIN_B(uniqueId);
    int32_t rc = dmg_shard_init(H(handle), world, rank, id);
    UNPIN_B(uniqueId, id, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_shardLoadDinWeightsFloat(
    JNIEnv *env, jobject self, jlong handle, jlong rowsGlobal, jint embedSize, jint seqLen, jfloatArray params)
{
    void *p = PIN_F(params);
    int32_t rc = dmg_shard_load_din_weights(H(handle), rowsGlobal, embedSize, seqLen, p);
    UNPIN_F(params, p, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_shardTdmRetrieve(
    JNIEnv *env, jobject self, jlong handle, jint batch, jintArray itemSeq, jint beam, jint topk, jboolean useMask,
    jintArray outItems, jfloatArray outLogits, jintArray outCounts)
{
    const jlong T = seq_len(H(handle));
    if (batch <= 0 || topk <= 0 || !need(env, itemSeq, (jlong)batch * T, "itemSeq: batch x seq_len ints") ||
        !need(env, outItems, (jlong)batch * topk, "outItems: batch x topk ints") ||
        !need(env, outLogits, (jlong)batch * topk, "outLogits: batch x topk floats") || !need(env, outCounts, batch, "outCounts: batch ints"))
        return;
    void *s = PIN_I(itemSeq), *oi = PIN_I(outItems), *ol = PIN_F(outLogits), *oc = PIN_I(outCounts);
    int32_t rc = dmg_shard_tdm_retrieve(H(handle), batch, s, beam, topk, useMask ? 1 : 0, oi, ol, oc);
    UNPIN_I(outCounts, oc, 0); UNPIN_F(outLogits, ol, 0); UNPIN_I(outItems, oi, 0); UNPIN_I(itemSeq, s, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

/* TreeLearning.reBalance for one level step (host code inside the library) */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_jtmAssignLevel(
    JNIEnv *env, jobject self, jlong handle, jintArray parentCode, jintArray oldChild, jint nChild, jfloatArray weights,
    jint maxAssign, jintArray outNode)
{
    jsize n = (*env)->GetArrayLength(env, parentCode);
    if (nChild <= 0 || !need(env, oldChild, n, "oldChild: one per item") || !need(env, weights, (jlong)n * nChild, "weights: items x nChild floats") ||
        !need(env, outNode, n, "outNode: one per item"))
        return;
    void *p = PIN_I(parentCode), *o = PIN_I(oldChild), *w = PIN_F(weights), *out = PIN_I(outNode);
    int32_t rc = dmg_jtm_assign_level(H(handle), n, p, o, nChild, w, maxAssign, out);
    UNPIN_I(outNode, out, 0); UNPIN_F(weights, w, JNI_ABORT); UNPIN_I(oldChild, o, JNI_ABORT); UNPIN_I(parentCode, p, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

/* TreeLearning.aggregateWeights for a level step; out: nItems x 2^(level - oldLevel) */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_jtmItemWeights(
    JNIEnv *env, jobject self, jlong handle, jlongArray sampleOff, jintArray sampleSeq, jintArray parentCode, jint oldLevel,
    jint level, jboolean hierarchical, jint minLevel, jboolean useMask, jfloatArray outWeights)
{
    jsize n = (*env)->GetArrayLength(env, parentCode);
    void *o = PIN_L(sampleOff), *s = PIN_I(sampleSeq), *p = PIN_I(parentCode), *w = PIN_F(outWeights);
    int32_t rc = dmg_jtm_item_weights(H(handle), n, o, s, p, oldLevel, level, hierarchical ? 1 : 0, minLevel, useMask ? 1 : 0, w);
    UNPIN_F(outWeights, w, 0); UNPIN_I(parentCode, p, JNI_ABORT); UNPIN_I(sampleSeq, s, JNI_ABORT); UNPIN_L(sampleOff, o, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

/* NegativeSampler.sample + MiniBatch.convert on the device; returns the number of rows written */
JNIEXPORT jint JNICALL Java_com_mass_gpu_DismemberGPU_00024_tdmSampleExpand(
    JNIEnv *env, jobject self, jlong handle, jintArray targets, jintArray itemSeq, jintArray layerNeg, jint startLevel,
    jboolean withProb, jint tolerance, jlong seed, jintArray outNode, jintArray outSeq, jfloatArray outLabel)
{
    jsize n = (*env)->GetArrayLength(env, targets);
    int32_t rows = 0;
    const jlong T = seq_len(H(handle));
    {
        /* layerSum rows per target (NegativeSampler.scala:57): the outputs must hold targets x layerSum (x seq_len) entries */
        jsize nl = layerNeg ? (*env)->GetArrayLength(env, layerNeg) : 0;
        jlong layer_sum = 0;
        for (jsize q = startLevel > 0 ? startLevel : 0; q < nl; q++) { jint v = 0; (*env)->GetIntArrayRegion(env, layerNeg, q, 1, &v); layer_sum += 1 + v; }
        if (!need(env, itemSeq, (jlong)n * T, "itemSeq: targets x seq_len ints") || !need(env, outNode, n * layer_sum, "outNode: targets x layerSum ints") ||
            !need(env, outSeq, n * layer_sum * T, "outSeq: targets x layerSum x seq_len ints") ||
            !need(env, outLabel, n * layer_sum, "outLabel: targets x layerSum floats"))
            return 0;
    }
    void *t = PIN_I(targets), *s = PIN_I(itemSeq), *l = PIN_I(layerNeg), *on = PIN_I(outNode), *os = PIN_I(outSeq), *ol = PIN_F(outLabel);
    int32_t rc = dmg_tdm_sample_expand(H(handle), n, t, s, l, startLevel, withProb ? 1 : 0, tolerance, (uint64_t)seed, on, os, ol, &rows);
    UNPIN_F(outLabel, ol, 0); UNPIN_I(outSeq, os, 0); UNPIN_I(outNode, on, 0);
    UNPIN_I(layerNeg, l, JNI_ABORT); UNPIN_I(itemSeq, s, JNI_ABORT); UNPIN_I(targets, t, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
    return rows;
}

/* Module.parameters() back to the JVM (Serialization.saveModel) */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_downloadDinWeightsFloat(JNIEnv *env, jobject self, jlong handle, jfloatArray params)
{
    jsize n = (*env)->GetArrayLength(env, params);
    void *p = PIN_F(params);
    int32_t rc = dmg_download_din_weights(H(handle), p, n);
    UNPIN_F(params, p, 0);
    if (rc) throw_status(env, H(handle), rc);
}

/* model.deep_model = "DeepFM" (TDM.scala:43) */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_loadDeepFmWeightsFloat(
    JNIEnv *env, jobject self, jlong handle, jlong rows, jint embedSize, jint seqLen, jfloatArray params)
{
    void *p = PIN_F(params);
    int32_t rc = dmg_load_deepfm_weights(H(handle), rows, embedSize, seqLen, p);
    UNPIN_F(params, p, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

/* Metrics.computeMetrics per user; out: batch x 3 (precision, recall, ndcg) */
/* OTM with deepModel = "DeepFM": DeepModel[Double] (otm/.../model/DeepFM.scala:12-48) */
JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_loadDeepFmWeightsDouble(
    JNIEnv *env, jobject self, jlong handle, jlong rows, jint embedSize, jint seqLen, jdoubleArray params)
{
    jdouble *p = PIN_D(params);
    int32_t rc = dmg_load_deepfm_weights_f64(H(handle), rows, embedSize, seqLen, p);
    UNPIN_D(params, p, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}

JNIEXPORT void JNICALL Java_com_mass_gpu_DismemberGPU_00024_evalMetrics(
    JNIEnv *env, jobject self, jlong handle, jint batch, jint topk, jintArray recItems, jintArray recCounts, jlongArray labelOff,
    jintArray labels, jdoubleArray out)
{
    void *r = PIN_I(recItems), *c = PIN_I(recCounts), *o = PIN_L(labelOff), *l = PIN_I(labels), *m = PIN_D(out);
    int32_t rc = dmg_eval_metrics(H(handle), batch, topk, r, c, o, l, m);
    UNPIN_D(out, m, 0); UNPIN_I(labels, l, JNI_ABORT); UNPIN_L(labelOff, o, JNI_ABORT); UNPIN_I(recCounts, c, JNI_ABORT); UNPIN_I(recItems, r, JNI_ABORT);
    if (rc) throw_status(env, H(handle), rc);
}
