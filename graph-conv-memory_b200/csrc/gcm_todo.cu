// Entry points declared in include/gcm_b200.h whose kernels have not landed yet.
// Each returns GCM_ERR_UNSUPPORTED (never a silent fallback).
#include "gcm_common.cuh"

#define GCM_TODO(name) \
  gcm_set_error(name ": not implemented yet"); \
  return GCM_ERR_UNSUPPORTED

extern "C" int gcm_sparse_write_flatten(float*, const float*, const int64_t*, const int64_t*, const int64_t*,
                                        const int64_t*, int, int, int, int, float*, int64_t*, int32_t*, void*) {
  GCM_TODO("gcm_sparse_write_flatten");
}
extern "C" int gcm_sparse_temporal_edges(const int64_t*, const int64_t*, const int64_t*, int, const int32_t*, int,
                                         int64_t, int32_t*, const int64_t*, int64_t*, int64_t, void*) {
  GCM_TODO("gcm_sparse_temporal_edges");
}
extern "C" int gcm_sparse_radius_edges(const float*, const int64_t*, const int64_t*, const int64_t*, int, int, int,
                                       int, int, int, float, int64_t, int32_t*, const int64_t*, int64_t*, int64_t,
                                       void*) {
  GCM_TODO("gcm_sparse_radius_edges");
}
extern "C" int gcm_sparse_graphconv_fwd(const float*, const int64_t*, const int64_t*, const float*, int64_t, int, int,
                                        const float*, const float*, int, float*, float*, void*) {
  GCM_TODO("gcm_sparse_graphconv_fwd");
}
extern "C" int gcm_sparse_graphconv_bwd(const float*, const float*, const float*, const float*, const int64_t*,
                                        const int64_t*, const float*, int64_t, int, int, const float*, const float*,
                                        int, float*, float*, float*, float*, float*, void*) {
  GCM_TODO("gcm_sparse_graphconv_bwd");
}
