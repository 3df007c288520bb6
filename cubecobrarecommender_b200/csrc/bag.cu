// Encoder first layer on sparse binary cubes: embedding-bag gather-sum and its backward.
//
// Replaces the dense `Dense(512)` on a 97%-zero (B, C) matrix of reference
// src/ml/model.py:27,36 (and :122 for the one-hot rows of I, where x@W1 == W1[r]):
// for a 0/1 row, x @ W1 is the sum of the W1 rows of the cube's cards.
//
// Layout: W1 float32 [C][H] row-major (Keras (in,out)); one CTA per cube, thread t owns
// the 128-bit column group t (H/4 threads), 8 independent 16-byte loads in flight.
#include "cc_common.cuh"

namespace cc {

constexpr int BAG_CHUNK = 256;

__global__ void bag_fwd_kernel(const float* __restrict__ w, int64_t ldw, int32_t h4,
                               const int32_t* __restrict__ idx, const int64_t* __restrict__ row_start,
                               const int32_t* __restrict__ row_len, const float* __restrict__ bias,
                               float* __restrict__ out, int64_t ldo, int relu, int round_tf32) {
  // lets a dependent tcgen05 GEMM / chain launch (programmatic stream serialisation) be scheduled and run its prologue
  // while this grid drains; it still waits (griddepcontrol.wait) for this grid's completion before touching memory
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  __shared__ int32_t sidx[BAG_CHUNK];
  const int b = blockIdx.x, t = threadIdx.x;
  const int len = row_len[b];
  const int32_t* list = idx + row_start[b];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int base = 0; base < len; base += BAG_CHUNK) {
    const int cnt = min(BAG_CHUNK, len - base);
    for (int i = t; i < cnt; i += blockDim.x) sidx[i] = list[base + i];
    __syncthreads();
    if (t < h4) {
      int k = 0;
      for (; k + 8 <= cnt; k += 8) {
        float4 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q)
          v[q] = __ldg(reinterpret_cast<const float4*>(w + int64_t(sidx[k + q]) * ldw) + t);
#pragma unroll
        for (int q = 0; q < 8; ++q) { acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w; }
      }
      for (; k < cnt; ++k) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(w + int64_t(sidx[k]) * ldw) + t);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    __syncthreads();
  }
  if (t < h4) {
    if (bias) {
      const float4 bv = __ldg(reinterpret_cast<const float4*>(bias) + t);
      acc.x += bv.x; acc.y += bv.y; acc.z += bv.z; acc.w += bv.w;
    }
    if (relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
    if (round_tf32) {
      acc.x = rn_tf32(acc.x); acc.y = rn_tf32(acc.y);
      acc.z = rn_tf32(acc.z); acc.w = rn_tf32(acc.w);
    }
    reinterpret_cast<float4*>(out + int64_t(b) * ldo)[t] = acc;
  }
}

__device__ __forceinline__ void red_add_v4(float* p, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// dW1[c,:] += g[b,:] for every card c of cube b  (g = gradient of the pre-activation,
// already multiplied by the ReLU mask).  128-bit vector reductions into L2.
__global__ void bag_bwd_kernel(const float* __restrict__ g, int64_t ldg, int32_t h4,
                               const int32_t* __restrict__ idx, const int64_t* __restrict__ row_start,
                               const int32_t* __restrict__ row_len, float* __restrict__ dw, int64_t ldw) {
  __shared__ int32_t sidx[BAG_CHUNK];
  const int b = blockIdx.x, t = threadIdx.x;
  const int len = row_len[b];
  const int32_t* list = idx + row_start[b];
  float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t < h4) gv = reinterpret_cast<const float4*>(g + int64_t(b) * ldg)[t];
  // rows of g that are entirely zero (dead ReLU) still have to be skipped per element
  const bool nz = (gv.x != 0.f) | (gv.y != 0.f) | (gv.z != 0.f) | (gv.w != 0.f);
  for (int base = 0; base < len; base += BAG_CHUNK) {
    const int cnt = min(BAG_CHUNK, len - base);
    for (int i = t; i < cnt; i += blockDim.x) sidx[i] = list[base + i];
    __syncthreads();
    if (t < h4 && nz) {
      for (int k = 0; k < cnt; ++k) red_add_v4(dw + int64_t(sidx[k]) * ldw + 4 * t, gv);
    }
    __syncthreads();
  }
}

}  // namespace cc

using namespace cc;

extern "C" {

int cc_bag_fwd(const float* w, int64_t ldw, int32_t hidden, const int32_t* idx, const int64_t* row_start,
               const int32_t* row_len, int32_t batch, const float* bias, float* out, int64_t ldo, int relu,
               int round_tf32, void* stream) {
  CC_NVTX("cc_bag_fwd");
  CC_REQUIRE(w && idx && row_start && row_len && out, "cc_bag_fwd: null pointer");
  CC_REQUIRE(hidden > 0 && hidden % 4 == 0 && hidden <= 4096 && ldw % 4 == 0 && ldo % 4 == 0 && ldw >= hidden &&
                 ldo >= hidden, "cc_bag_fwd: hidden=%d ldw=%lld ldo=%lld must be multiples of 4", hidden,
             (long long)ldw, (long long)ldo);
  CC_REQUIRE((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(bias)) % 16 == 0,
             "cc_bag_fwd: pointers must be 16-byte aligned");
  if (batch == 0) return CC_OK;
  const int h4 = hidden / 4;
  const int threads = ((h4 + 31) / 32) * 32;
  bag_fwd_kernel<<<batch, threads, 0, as_stream(stream)>>>(w, ldw, h4, idx, row_start, row_len, bias, out, ldo, relu, round_tf32);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

int cc_bag_bwd(const float* g, int64_t ldg, int32_t hidden, const int32_t* idx, const int64_t* row_start,
               const int32_t* row_len, int32_t batch, float* dw, int64_t ldw, void* stream) {
  CC_NVTX("cc_bag_bwd");
  CC_REQUIRE(g && idx && row_start && row_len && dw, "cc_bag_bwd: null pointer");
  CC_REQUIRE(hidden > 0 && hidden % 4 == 0 && hidden <= 4096 && ldw % 4 == 0 && ldg % 4 == 0, "cc_bag_bwd: bad sizes");
  CC_REQUIRE((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(dw)) % 16 == 0,
             "cc_bag_bwd: pointers must be 16-byte aligned");
  if (batch == 0) return CC_OK;
  const int h4 = hidden / 4;
  const int threads = ((h4 + 31) / 32) * 32;
  bag_bwd_kernel<<<batch, threads, 0, as_stream(stream)>>>(g, ldg, h4, idx, row_start, row_len, dw, ldw);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

}  // extern "C"
