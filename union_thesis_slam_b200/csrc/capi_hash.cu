// capi_hash.cu -- C ABI of the voxel-block hash map (include/tsdf_b200.h, vh_*).
// PLACEHOLDER while the dense path is brought up on the GPU: every entry point fails loudly.
#include "host_common.h"

#define NOT_YET                                              \
    tsdf_set_error("%s: hash map not implemented yet", __func__); \
    return TSDF_ERR_BAD_ARG

extern "C" {
int vh_create(const int64_t *, const float *, double, double, int64_t, int64_t, int, int, int, tsdf_hash **) { NOT_YET; }
int vh_destroy(tsdf_hash *) { return TSDF_OK; }
int vh_integrate(tsdf_hash *, const uint8_t *, const double *, int, int, const double *, const double *, double) { NOT_YET; }
int vh_integrate_device(tsdf_hash *, const uint8_t *, const double *, int, int, const double *, const double *, double) { NOT_YET; }
int vh_stats(tsdf_hash *, int64_t *) { NOT_YET; }
int vh_num_blocks(tsdf_hash *, int64_t *) { NOT_YET; }
int vh_export_blocks(tsdf_hash *, int64_t, int64_t *, int32_t *, float *) { NOT_YET; }
int vh_to_dense(tsdf_hash *, float *, float *, float *) { NOT_YET; }
int vh_to_dense_device(tsdf_hash *, float *, float *, float *) { NOT_YET; }
int vh_remove_blocks(tsdf_hash *, const int32_t *, int64_t, int64_t *) { NOT_YET; }
int vh_sync(tsdf_hash *) { NOT_YET; }
int vh_last_kernel_ms(tsdf_hash *, float *) { NOT_YET; }
}
