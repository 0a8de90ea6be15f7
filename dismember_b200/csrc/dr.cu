// dr.cu -- Deep Retrieval entry points (placeholder until the kernels land).
#include "dmg_common.cuh"
using namespace dmg;
void dmg_free_dr(DrDev &d) { d = DrDev(); }
DMG_API int32_t dmg_dr_load(dmg_handle_t h, int32_t, int32_t, int32_t, int32_t, int32_t, const double *, const double *const *,
                            const double *const *, const double *, const double *, const double *, const double *, const double *)
{ return fail(h, DMG_ERR_UNSUPPORTED, "dmg_dr_load: not built yet"); }
DMG_API int32_t dmg_dr_load_paths(dmg_handle_t h, const int64_t *, const int32_t *) { return fail(h, DMG_ERR_UNSUPPORTED, "not built yet"); }
DMG_API int32_t dmg_dr_beam_search(dmg_handle_t h, int32_t, const int32_t *, int32_t, int32_t *, double *, int32_t *) { return fail(h, DMG_ERR_UNSUPPORTED, "not built yet"); }
DMG_API int32_t dmg_dr_retrieve(dmg_handle_t h, int32_t, const int32_t *, int32_t, int32_t, int32_t *, double *, int32_t *) { return fail(h, DMG_ERR_UNSUPPORTED, "not built yet"); }
