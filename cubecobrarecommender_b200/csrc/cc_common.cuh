// Shared helpers for the cubecobra_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/cubecobra_b200.h"

namespace cc {

// ---- error convention: int return, message kept per host thread -------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define CC_CHECK_CUDA(expr)                                                     \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) return cc::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define CC_CHECK_LAUNCH()                                                          \
  do {                                                                             \
    cudaError_t _e = cudaPeekAtLastError();                                        \
    if (_e != cudaSuccess) return cc::cuda_fail(_e, "launch", __FILE__, __LINE__); \
  } while (0)

#define CC_REQUIRE(cond, ...)        \
  do {                               \
    if (!(cond)) {                   \
      cc::set_error(__VA_ARGS__);    \
      return CC_ERR_ARGUMENT;        \
    }                                \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// NVTX range around every C-ABI entry point that enqueues work (header-only NVTX v3: a no-op unless a tool such as
// Nsight Systems / ncu --nvtx injects itself), so a timeline shows which host call a kernel belongs to.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
#define CC_NVTX(name) cc::NvtxRange cc_nvtx_range_(name)

int sm_count();

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

// ---- device helpers ---------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// round-to-nearest fp32 -> tf32 (the tensor core's kind::tf32 truncates the low 13 mantissa bits, which
// biases every product towards zero; operands are rounded once, where they are produced)
__device__ __forceinline__ float rn_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// the same rounding (nearest, ties away from zero) as two integer operations: add half a tf32 ulp to the
// magnitude bits, clear the 13 low bits (cvt.rna.tf32.f32 expands to more, with Inf/NaN handling the gradients
// and activations rounded in the GEMM epilogues never need)
__device__ __forceinline__ float rn_tf32_bits(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

// sigmoid of a logit as the recommenders report it (reference ml_recommend.py:78-80 reads float32 sigmoid
// outputs); ONE definition, so the fused select and the standalone sigmoid pass give identical bits
__device__ __forceinline__ float sigmoid_f32(float z) { return 1.f / (1.f + expf(-z)); }

// streaming 128-bit loads that do not pollute L1 (data is read once per CTA)
__device__ __forceinline__ uint4 ld_nc_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_nc_f4(const void* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
#endif

}  // namespace cc
