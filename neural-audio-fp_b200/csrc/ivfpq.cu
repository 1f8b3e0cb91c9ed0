// IVF-PQ index -- placeholder translation unit until the kernels land.
#include "index.h"

namespace nafp {
int ivfpq_create(nafp_index*, int, int, int) {
    set_error("IVF-PQ kernels are not built yet");
    return NAFP_ERR_UNSUPPORTED;
}
void ivfpq_destroy(nafp_index*) {}
int ivfpq_train(nafp_index*, const float*, int64_t, int64_t) { return NAFP_ERR_UNSUPPORTED; }
int ivfpq_add_rows(nafp_index*, int64_t, int64_t) { return NAFP_ERR_UNSUPPORTED; }
int ivfpq_search_dev(nafp_index*, const float*, int64_t, int, float*, int64_t*) { return NAFP_ERR_UNSUPPORTED; }
}  // namespace nafp

extern "C" int nafp_index_is_trained(nafp_index* idx) {
    if (!idx) return 0;
    return idx->type == NAFP_INDEX_FLAT_L2 ? 1 : 0;
}
