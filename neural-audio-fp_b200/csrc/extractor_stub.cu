// Encoder entry points -- placeholder until encoder.cu lands.
#include "common.h"
namespace nafp {
void encoder_destroy(nafp_ctx*) {}
}
using namespace nafp;
#define STUB(...) { set_error("encoder kernels are not built yet"); return NAFP_ERR_UNSUPPORTED; }
extern "C" {
int nafp_weights_load(nafp_ctx*, const float* const*, const float* const*, const float* const*, const float* const*, const float*, const float*, const float*, const float*) STUB()
int nafp_encoder_forward(nafp_ctx*, const float*, int64_t, float*) STUB()
int nafp_fingerprint(nafp_ctx*, const float*, int64_t, int64_t, float*) STUB()
int nafp_fingerprint_host(nafp_ctx*, const float*, int64_t, int64_t, float*) STUB()
int nafp_fingerprint_pcm16_host(nafp_ctx*, const int16_t*, int64_t, int64_t, float*) STUB()
int nafp_encoder_activation_host(nafp_ctx*, int, int64_t, float*) STUB()
}
