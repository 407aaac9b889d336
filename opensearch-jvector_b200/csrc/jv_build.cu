// jv_build.cu — device-side fixtures for the "next" rows: PQ codebook training and
// Vamana graph construction.  (placeholder: filled in after the query path is parity-green)
#include "jv_internal.h"

extern "C" {

int32_t jv_pq_train_dev(int32_t, const float *, int64_t, int32_t, int32_t, int32_t, int32_t, int32_t, uint64_t, float *, float *) {
    jv::set_error("jv_pq_train_dev: not implemented yet");
    return JV_ERR_UNSUPPORTED;
}
int32_t jv_pq_train(int32_t, const float *, int64_t, int32_t, int32_t, int32_t, int32_t, int32_t, uint64_t, float *, float *) {
    jv::set_error("jv_pq_train: not implemented yet");
    return JV_ERR_UNSUPPORTED;
}
int32_t jv_graph_build_dev(int32_t, const float *, int64_t, int32_t, int32_t, int32_t, int32_t, float, float, int32_t *, int32_t *) {
    jv::set_error("jv_graph_build_dev: not implemented yet");
    return JV_ERR_UNSUPPORTED;
}
int32_t jv_graph_build(int32_t, const float *, int64_t, int32_t, int32_t, int32_t, int32_t, float, float, int32_t *, int32_t *) {
    jv::set_error("jv_graph_build: not implemented yet");
    return JV_ERR_UNSUPPORTED;
}
}
