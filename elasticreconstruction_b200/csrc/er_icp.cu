// er_icp.cu -- placeholder entry points so the ABI is complete while the ICP kernels are being written.
#include "er_common.h"
extern "C" {
er_cloud* er_cloud_upload(int, const float*, const float*, size_t, double) { er::set_error("ICP path not built yet"); return nullptr; }
void er_cloud_free(er_cloud*) {}
size_t er_cloud_size(const er_cloud*) { return 0; }
int er_icp_count_inliers(er_cloud*, er_cloud*, const double*, double, int*) { er::set_error("ICP path not built yet"); return ER_ERR_ARG; }
int er_icp_align(er_cloud*, er_cloud*, const double*, double, int, double, int, double*, int*, double*) { er::set_error("ICP path not built yet"); return ER_ERR_ARG; }
int er_icp_align_batch(er_cloud* const*, int, const int*, const int*, int, const double*, double, int, double, int, double*, int*, double*) { er::set_error("ICP path not built yet"); return ER_ERR_ARG; }
int er_icp_correspondences(er_cloud*, er_cloud*, const double*, double, double, int32_t*, size_t, size_t*, double*) { er::set_error("ICP path not built yet"); return ER_ERR_ARG; }
int er_icp_event_record(int, int) { return ER_ERR_ARG; }
int er_icp_event_elapsed_ms(int, int, int, float*) { return ER_ERR_ARG; }
uint64_t er_icp_launch_count(int) { return 0; }
int er_icp_sync(int) { return ER_ERR_ARG; }
}
