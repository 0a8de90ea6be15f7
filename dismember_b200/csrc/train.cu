// train.cu -- training / JTM entry points (placeholder until the kernels land).
#include "dmg_common.cuh"
using namespace dmg;
DMG_API int32_t dmg_train_step(dmg_handle_t h, int64_t, const int32_t *, const int32_t *, const int32_t *, int64_t, const void *, double, int32_t, void *) { return fail(h, DMG_ERR_UNSUPPORTED, "not built yet"); }
DMG_API int32_t dmg_din_gradients(dmg_handle_t h, int64_t, const int32_t *, const int32_t *, const int32_t *, int64_t, const void *, void *, void *, int64_t) { return fail(h, DMG_ERR_UNSUPPORTED, "not built yet"); }
DMG_API int32_t dmg_tdm_sample_expand(dmg_handle_t h, int32_t, const int32_t *, const int32_t *, const int32_t *, int32_t, uint64_t, int32_t *, int32_t *, float *, int32_t *) { return fail(h, DMG_ERR_UNSUPPORTED, "not built yet"); }
DMG_API int32_t dmg_jtm_item_weights(dmg_handle_t h, int32_t, const int64_t *, const int32_t *, const int32_t *, int32_t, int32_t, float *) { return fail(h, DMG_ERR_UNSUPPORTED, "not built yet"); }
