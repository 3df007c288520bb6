// In-cube ("cuts") scores of the ML recommender.
//
// Replaces the cuts walk of reference src/scripts/ml_recommend.py:105-108 / web/ml_recommend_web.py:61-64
// (`results[idx]` for every in-cube idx, in cubelist order) for a whole batch: the score of every CSR entry of every
// cube is picked out of the cube's logit row, with the same float32 sigmoid the select kernels rank by.
#include "cc_common.cuh"

namespace cc {

// one CTA per cube; out[p] for the CSR entries p of the cube (ids outside [0, C) give 0; the reference would raise)
__global__ void __launch_bounds__(128)
cuts_gather_kernel(const float* __restrict__ scores, int64_t ld, int32_t num_cards, const int64_t* __restrict__ row_ptr,
                   const int32_t* __restrict__ idx, int apply_sigmoid, float* __restrict__ out) {
  const int cube = blockIdx.x;
  const float* row = scores + int64_t(cube) * ld;
  const int64_t pe = row_ptr[cube + 1];
  for (int64_t p = row_ptr[cube] + threadIdx.x; p < pe; p += blockDim.x) {
    const int32_t c = idx[p];
    float v = 0.f;
    if (c >= 0 && c < num_cards) {
      v = row[c];
      if (apply_sigmoid) v = sigmoid_f32(v);
    }
    out[p] = v;
  }
}

}  // namespace cc

using namespace cc;

extern "C" {

int cc_cuts_gather_f32(const float* scores, int64_t ld, int32_t num_cards, int32_t batch, const int64_t* row_ptr,
                       const int32_t* idx, int apply_sigmoid, float* out, void* stream) {
  CC_NVTX("cc_cuts_gather_f32");
  CC_REQUIRE(scores && row_ptr && out, "cc_cuts_gather_f32: null pointer");
  CC_REQUIRE(num_cards > 0 && batch >= 0 && ld >= num_cards, "cc_cuts_gather_f32: bad sizes");
  if (batch == 0) return CC_OK;
  cuts_gather_kernel<<<batch, 128, 0, as_stream(stream)>>>(scores, ld, num_cards, row_ptr, idx, apply_sigmoid, out);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

}  // extern "C"
