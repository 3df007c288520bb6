// Exact-fp32 GEMM (FFMA, no tensor cores) with the Dense-layer epilogues.
//
// This is the "fp32" precision mode of the DAE (reference src/ml/model.py:27-33,58-64:
// Dense = act(x @ kernel + bias)); the tcgen05 kernels in gemm_tc.cu are the tf32/bf16
// modes.  C[M,N] = epi(op(A) op(B)), all row-major:
//   transa=0: A is [M,K] (lda)   transa=1: A is [K,M] (lda), used transposed
//   transb=0: B is [K,N] (ldb)   transb=1: B is [N,K] (ldb), used transposed
#include "cc_common.cuh"

namespace cc {

constexpr int BM = 128, BN = 128, BK = 16, GEMM_THREADS = 256;

struct GemmArgs {
  const float* a; const float* b; float* c;
  int64_t lda, ldb, ldc;
  int m, n, k;
  const float* bias;      // [N] or null
  const float* mask;      // [M, ldmask] or null: output multiplied by (mask > 0)
  int64_t ldmask;
  int relu, accumulate;
};

template <bool TA, bool TB>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_simt_kernel(const GemmArgs g) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  // each thread moves 8 elements of each operand tile per k-step
  float ra[8], rb[8];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int mm, kk;
      if (TA) { mm = tid & 127; kk = (tid >> 7) + 2 * i; }     // M contiguous
      else    { kk = tid & 15;  mm = (tid >> 4) + 16 * i; }    // K contiguous
      const int gm = m0 + mm, gk = k0 + kk;
      float v = 0.f;
      if (gm < g.m && gk < g.k) v = TA ? __ldg(g.a + int64_t(gk) * g.lda + gm) : __ldg(g.a + int64_t(gm) * g.lda + gk);
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int nn, kk;
      if (!TB) { nn = tid & 127; kk = (tid >> 7) + 2 * i; }    // N contiguous
      else     { kk = tid & 15;  nn = (tid >> 4) + 16 * i; }   // K contiguous
      const int gn = n0 + nn, gk = k0 + kk;
      float v = 0.f;
      if (gn < g.n && gk < g.k) v = TB ? __ldg(g.b + int64_t(gn) * g.ldb + gk) : __ldg(g.b + int64_t(gk) * g.ldb + gn);
      rb[i] = v;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (TA) As[(tid >> 7) + 2 * i][tid & 127] = ra[i];
      else    As[tid & 15][(tid >> 4) + 16 * i] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (!TB) Bs[(tid >> 7) + 2 * i][tid & 127] = rb[i];
      else     Bs[tid & 15][(tid >> 4) + 16 * i] = rb[i];
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  fetch(0);
  for (int k0 = 0; k0 < g.k; k0 += BK) {
    stash();
    __syncthreads();
    if (k0 + BK < g.k) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= g.m) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (gn >= g.n) continue;
      float v = acc[i][j];
      if (g.bias) v += __ldg(g.bias + gn);
      if (g.relu) v = fmaxf(v, 0.f);
      if (g.mask) v = (__ldg(g.mask + int64_t(gm) * g.ldmask + gn) > 0.f) ? v : 0.f;
      float* p = g.c + int64_t(gm) * g.ldc + gn;
      *p = g.accumulate ? *p + v : v;
    }
  }
}

// column sums: out[n] (+)= sum_m x[m, n]   (bias gradients).  Deterministic two-stage.
constexpr int CS_ROWS = 128;
__global__ void __launch_bounds__(128)
colsum_partial_kernel(const float* __restrict__ x, int64_t ld, int m, int n, float* __restrict__ partial) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int i0 = blockIdx.y * CS_ROWS, i1 = min(i0 + CS_ROWS, m);
  float s = 0.f;
  for (int i = i0; i < i1; ++i) s += x[int64_t(i) * ld + j];
  partial[int64_t(blockIdx.y) * n + j] = s;
}
__global__ void colsum_final_kernel(const float* __restrict__ partial, int chunks, int n, float* __restrict__ out,
                                    int accumulate) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  float s = 0.f;
  for (int c = 0; c < chunks; ++c) s += partial[int64_t(c) * n + j];
  out[j] = accumulate ? out[j] + s : s;
}

// y = x * (act > 0)   (ReLU backward on a separate buffer)
__global__ void relu_mask_kernel(float* __restrict__ x, int64_t ldx, const float* __restrict__ act, int64_t lda,
                                 int m, int n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= int64_t(m) * n) return;
  const int r = int(i / n), c = int(i % n);
  if (!(act[int64_t(r) * lda + c] > 0.f)) x[int64_t(r) * ldx + c] = 0.f;
}

}  // namespace cc

using namespace cc;

extern "C" {

int cc_gemm_f32_simt(int transa, int transb, int m, int n, int k, const float* a, int64_t lda, const float* b,
                     int64_t ldb, float* c, int64_t ldc, const float* bias, int relu, const float* mask,
                     int64_t ldmask, int accumulate, void* stream) {
  CC_NVTX("cc_gemm_f32_simt");
  CC_REQUIRE(a && b && c, "cc_gemm_f32_simt: null pointer");
  CC_REQUIRE(m >= 0 && n >= 0 && k >= 0, "cc_gemm_f32_simt: negative size");
  if (m == 0 || n == 0) return CC_OK;
  GemmArgs g{a, b, c, lda, ldb, ldc, m, n, k, bias, mask, ldmask, relu, accumulate};
  dim3 grid(ceil_div(n, BN), ceil_div(m, BM));
  cudaStream_t st = as_stream(stream);
  if (transa) {
    if (transb) gemm_simt_kernel<true, true><<<grid, GEMM_THREADS, 0, st>>>(g);
    else        gemm_simt_kernel<true, false><<<grid, GEMM_THREADS, 0, st>>>(g);
  } else {
    if (transb) gemm_simt_kernel<false, true><<<grid, GEMM_THREADS, 0, st>>>(g);
    else        gemm_simt_kernel<false, false><<<grid, GEMM_THREADS, 0, st>>>(g);
  }
  CC_CHECK_LAUNCH();
  return CC_OK;
}

int64_t cc_colsum_workspace_bytes(int m, int n) { return int64_t(ceil_div(m, CS_ROWS)) * n * 4; }

int cc_colsum_f32(const float* x, int64_t ld, int m, int n, float* workspace, float* out, int accumulate,
                  void* stream) {
  CC_NVTX("cc_colsum_f32");
  CC_REQUIRE(x && workspace && out && m >= 0 && n > 0, "cc_colsum_f32: bad arguments");
  cudaStream_t st = as_stream(stream);
  const int chunks = ceil_div(m, CS_ROWS);
  if (chunks > 0) {
    colsum_partial_kernel<<<dim3(ceil_div(n, 128), chunks), 128, 0, st>>>(x, ld, m, n, workspace);
    CC_CHECK_LAUNCH();
  }
  colsum_final_kernel<<<ceil_div(n, 256), 256, 0, st>>>(workspace, chunks, n, out, accumulate);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

int cc_relu_mask_f32(float* x, int64_t ldx, const float* act, int64_t lda, int m, int n, void* stream) {
  CC_NVTX("cc_relu_mask_f32");
  CC_REQUIRE(x && act, "cc_relu_mask_f32: null pointer");
  if (m == 0 || n == 0) return CC_OK;
  const int64_t total = int64_t(m) * n;
  relu_mask_kernel<<<(unsigned)ceil_div<int64_t>(total, 256), 256, 0, as_stream(stream)>>>(x, ldx, act, lda, m, n);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

}  // extern "C"
