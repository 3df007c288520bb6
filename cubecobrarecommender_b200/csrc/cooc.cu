// Co-occurrence graph build: bit-pack -> popcount count -> row normalise.
//
// Replaces reference src/non_ml/utils.py:75-92 (create_adjacency_matrix),
// src/ml/train.py:69-71 (M-hat) and src/ml/generator.py:30 (neg_sampler).
//
// Data layout in HBM
//   bits   uint32 [Kw_pad][Cpad]   word w, card c: bit (k & 31) of bits[k>>5][c] is set
//                                  iff cube k contains card c ("word-major": a tile of
//                                  128 cards x KC words is KC contiguous 512-byte rows,
//                                  so tiles land in shared memory with no transpose)
//   counts int32  [C][ld]          counts[i][j] = |{k : i in cube k and j in cube k}|
//   M      float64 [C][C]          counts[i][j]/counts[i][i]   (utils.py:85-89)
//   M-hat  float32 [C][ld_mhat]    diag<-1, row / row sum      (train.py:69-71)
#include <thread>
#include <vector>

#include "cc_common.cuh"

namespace cc {

// ------------------------------------------------------------------ bit-pack
__global__ void bitpack_kernel(const int64_t* __restrict__ indptr,
                               const int32_t* __restrict__ indices, int64_t num_cubes,
                               int32_t num_cards, uint32_t* __restrict__ bits, int64_t cpad,
                               int* __restrict__ bad) {
  // one warp per cube; duplicates collapse because OR is idempotent (utils.py:71)
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= num_cubes) return;
  const int64_t beg = indptr[warp], end = indptr[warp + 1];
  uint32_t* row = bits + (warp >> 5) * cpad;
  const uint32_t bit = 1u << (warp & 31);
  for (int64_t p = beg + lane; p < end; p += 32) {
    const int32_t c = indices[p];
    if (c < 0 || c >= num_cards) { atomicExch(bad, 1); continue; }
    atomicOr(row + c, bit);
  }
}

// ------------------------------------------------------------- popcount count
// One CTA = one 128x128 tile of counts on or above the diagonal; 256 threads, each
// 8x8 outputs.  The K (cube) dimension streams through shared memory in stages of
// KC words with cp.async double buffering; every word pair costs AND + POPC + IADD.
constexpr int TILE = 128;
constexpr int KC = 16;
constexpr int COUNT_THREADS = 256;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__global__ void __launch_bounds__(COUNT_THREADS, 2)
cooc_count_kernel(const uint32_t* __restrict__ bits, int64_t kw_pad, int64_t cpad, int32_t num_cards,
                  int32_t* __restrict__ counts, int64_t ld, int accumulate) {
  const int ti = blockIdx.y, tj = blockIdx.x;
  if (tj < ti) return;  // lower triangle is written as the transpose of the upper one
  __shared__ __align__(16) uint32_t sA[2][KC][TILE];
  __shared__ __align__(16) uint32_t sB[2][KC][TILE];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const uint32_t* gA = bits + int64_t(ti) * TILE;
  const uint32_t* gB = bits + int64_t(tj) * TILE;

  auto load_stage = [&](int buf, int64_t w0) {
    // KC rows x 32 16-byte chunks per operand
#pragma unroll
    for (int it = 0; it < (KC * 32) / COUNT_THREADS; ++it) {
      const int chunk = tid + it * COUNT_THREADS;
      const int r = chunk >> 5, c4 = (chunk & 31) << 2;
      cp_async16(&sA[buf][r][c4], gA + (w0 + r) * cpad + c4);
      cp_async16(&sB[buf][r][c4], gB + (w0 + r) * cpad + c4);
    }
    cp_async_commit();
  };

  int acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0;

  const int64_t stages = kw_pad / KC;
  load_stage(0, 0);
  for (int64_t s = 0; s < stages; ++s) {
    const int buf = int(s & 1);
    if (s + 1 < stages) { load_stage(buf ^ 1, (s + 1) * KC); cp_async_wait<1>(); }
    else                { cp_async_wait<0>(); }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < KC; ++w) {
      // rows  i = ty*8 .. ty*8+7          (two broadcast LDS.128)
      // cols  j = tx*4 .. +3 and 64 + tx*4 .. +3   (two conflict-free LDS.128)
      const uint4 a0 = *reinterpret_cast<const uint4*>(&sA[buf][w][ty * 8]);
      const uint4 a1 = *reinterpret_cast<const uint4*>(&sA[buf][w][ty * 8 + 4]);
      const uint4 b0 = *reinterpret_cast<const uint4*>(&sB[buf][w][tx * 4]);
      const uint4 b1 = *reinterpret_cast<const uint4*>(&sB[buf][w][64 + tx * 4]);
      const uint32_t a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const uint32_t b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] += __popc(a[i] & b[j]);
    }
    __syncthreads();
  }

  // write the tile, and its transpose when off the diagonal
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gi = ti * TILE + ty * 8 + i;
    if (gi >= num_cards) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int gj0 = tj * TILE + h * 64 + tx * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gj = gj0 + j;
        if (gj >= num_cards) continue;
        const int v = acc[i][h * 4 + j];
        int32_t* p = counts + int64_t(gi) * ld + gj;
        *p = accumulate ? *p + v : v;
        if (ti != tj) {
          int32_t* q = counts + int64_t(gj) * ld + gi;
          *q = accumulate ? *q + v : v;
        }
      }
    }
  }
}

// --------------------------------------------------------------- row normalise
// One CTA per card i.  Emits, from the int32 counts row:
//   M[i,:]      float64  = cnt/cnt[i,i] if cnt[i,i] != 0 else cnt      (utils.py:85-89)
//   rowsum[i]   float64  = sum_j y[i,j], y = M with diagonal forced to 1 (train.py:69-70)
//   Mhat[i,:]   float32  = y[i,:]/rowsum[i]                              (train.py:71)
// Streaming form: 16-byte loads of the counts row, 32-byte runs of float64 stores.  Only M needs the IEEE
// division cnt/cnt[i,i] (bit-exact float64 output); M-hat is float32 (bar 1e-6 relative), so its two divisions
// collapse into one multiply by 1/(cnt[i,i]*rowsum) computed once per row.
template <bool VEC>
__global__ void __launch_bounds__(256)
row_normalise_kernel(const int32_t* __restrict__ counts, int64_t ld, int32_t num_cards,
                     double* __restrict__ m64, int64_t ld_m, float* __restrict__ mhat, int64_t ld_mhat,
                     double* __restrict__ rowsum, int has_force_diag, double force_diag, int32_t row0) {
  // row0 != 0: the buffers hold the row block [row0, row0 + gridDim.x) of the matrices (a rank's shard of a
  // reduce-scattered build); local row blockIdx.x is card i = row0 + blockIdx.x
  const int i = row0 + blockIdx.x;
  const int32_t* row = counts + int64_t(blockIdx.x) * ld;
  const int32_t d = row[i];
  const double dd = double(d);
  __shared__ double red[8];
  double s = 0.0;
  double* mrow = m64 ? m64 + int64_t(blockIdx.x) * ld_m : nullptr;
  const int nvec = VEC ? (num_cards >> 2) : 0;
  for (int q = threadIdx.x; q < nvec; q += blockDim.x) {
    const int4 c4 = *reinterpret_cast<const int4*>(row + 4 * q);
    const int cc[4] = {c4.x, c4.y, c4.z, c4.w};
    double v[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const double c = double(cc[t]);
      v[t] = d != 0 ? c / dd : c;
      s += (4 * q + t == i) ? 1.0 : v[t];
      if (has_force_diag && 4 * q + t == i) v[t] = force_diag;
    }
    if (mrow) {
      *reinterpret_cast<double2*>(mrow + 4 * q) = make_double2(v[0], v[1]);
      *reinterpret_cast<double2*>(mrow + 4 * q + 2) = make_double2(v[2], v[3]);
    }
  }
  for (int j = 4 * nvec + threadIdx.x; j < num_cards; j += blockDim.x) {
    const double c = double(row[j]);
    const double v = d != 0 ? c / dd : c;
    if (mrow) mrow[j] = (has_force_diag && j == i) ? force_diag : v;
    s += (j == i) ? 1.0 : v;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    t = warp_sum(t);
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  const double total = red[0];
  if (threadIdx.x == 0 && rowsum) rowsum[blockIdx.x] = total;
  if (mhat) {
    const double scale = d != 0 ? 1.0 / (dd * total) : 1.0 / total;     // y[i,j]/rowsum = cnt * scale off the diagonal
    const float diag = float(1.0 / total);
    float* hrow = mhat + int64_t(blockIdx.x) * ld_mhat;
    for (int q = threadIdx.x; q < nvec; q += blockDim.x) {
      const int4 c4 = *reinterpret_cast<const int4*>(row + 4 * q);
      float4 o = make_float4(float(double(c4.x) * scale), float(double(c4.y) * scale), float(double(c4.z) * scale),
                             float(double(c4.w) * scale));
      const int t = i - 4 * q;
      if (t == 0) o.x = diag; else if (t == 1) o.y = diag; else if (t == 2) o.z = diag; else if (t == 3) o.w = diag;
      *reinterpret_cast<float4*>(hrow + 4 * q) = o;
    }
    for (int j = 4 * nvec + threadIdx.x; j < num_cards; j += blockDim.x)
      hrow[j] = (j == i) ? diag : float(double(row[j]) * scale);
  }
}

// column mass of M-hat in float64: partial[chunk][j] = sum_{i in chunk} y[i,j]/rowsum[i]
constexpr int COL_ROWS = 256;
// y[i,j]/rowsum[i] = cnt[i,j] * scale_i off the diagonal with scale_i = 1/(cnt[i,i]*rowsum[i]) (1/rowsum[i] for an unseen
// card), and 1/rowsum[i] on it: one reciprocal per ROW, staged in shared memory, instead of two float64 divisions per
// element (the kernel was bound by the FP64 divider, not by HBM)
__global__ void __launch_bounds__(128)
col_mass_partial_kernel(const int32_t* __restrict__ counts, int64_t ld, int32_t num_cards,
                        const double* __restrict__ rowsum, double* __restrict__ partial, int32_t row0, int32_t nrows) {
  // counts / rowsum hold the row block [row0, row0 + nrows) (row0 = 0, nrows = num_cards: the whole matrix); i below
  // is the LOCAL row, its card (= the column of its diagonal element) is row0 + i
  __shared__ double s_scale[COL_ROWS], s_diag[COL_ROWS];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i0 = blockIdx.y * COL_ROWS;
  const int i1 = min(i0 + COL_ROWS, nrows);
  for (int r = threadIdx.x; r < i1 - i0; r += blockDim.x) {
    const int i = i0 + r;
    const int32_t d = counts[int64_t(i) * ld + row0 + i];
    const double rs = rowsum[i];
    s_diag[r] = 1.0 / rs;
    s_scale[r] = d != 0 ? 1.0 / (double(d) * rs) : 1.0 / rs;
  }
  __syncthreads();
  if (j >= num_cards) return;
  double s = 0.0;
#pragma unroll 8
  for (int i = i0; i < i1; ++i) {
    const double c = double(counts[int64_t(i) * ld + j]);
    s += (j == row0 + i) ? s_diag[i - i0] : c * s_scale[i - i0];
  }
  partial[int64_t(blockIdx.y) * num_cards + j] = s;
}
__global__ void col_mass_final_kernel(const double* __restrict__ partial, int chunks, int32_t num_cards,
                                      double* __restrict__ neg_sampler) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= num_cards) return;
  double s = 0.0;
  for (int c = 0; c < chunks; ++c) s += partial[int64_t(c) * num_cards + j];
  // M-hat.sum() == number of rows (each row sums to 1) up to rounding; the reference
  // divides by the actual total (generator.py:30), so do the same deterministic sum
  neg_sampler[j] = s;
}
__global__ void col_mass_scale_kernel(double* __restrict__ neg_sampler, int32_t num_cards) {
  // single CTA: total = sum_j neg[j] (fixed order), then neg /= total
  __shared__ double red[32];
  __shared__ double total_s;
  double s = 0.0;
  for (int j = threadIdx.x; j < num_cards; j += blockDim.x) s += neg_sampler[j];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    t = warp_sum(t);
    if (threadIdx.x == 0) total_s = t;
  }
  __syncthreads();
  const double total = total_s;
  for (int j = threadIdx.x; j < num_cards; j += blockDim.x) neg_sampler[j] /= total;
}

}  // namespace cc

using namespace cc;

extern "C" {

int64_t cc_bits_words(int64_t num_cubes) { return ceil_div<int64_t>(ceil_div<int64_t>(num_cubes, 32), KC) * KC; }
int64_t cc_bits_cpad(int32_t num_cards) { return ceil_div<int64_t>(num_cards, TILE) * TILE; }

int cc_bitpack_cubes(const int64_t* indptr, const int32_t* indices, int64_t num_cubes, int32_t num_cards,
                     uint32_t* bits, int* bad_flag, void* stream) {
  CC_NVTX("cc_bitpack_cubes");
  CC_REQUIRE(num_cubes >= 0 && num_cards > 0, "cc_bitpack_cubes: bad sizes K=%lld C=%d", (long long)num_cubes, num_cards);
  CC_REQUIRE(bits && bad_flag, "cc_bitpack_cubes: null output");
  cudaStream_t st = as_stream(stream);
  const int64_t kw = cc_bits_words(num_cubes), cpad = cc_bits_cpad(num_cards);
  CC_CHECK_CUDA(cudaMemsetAsync(bits, 0, size_t(kw) * cpad * sizeof(uint32_t), st));
  CC_CHECK_CUDA(cudaMemsetAsync(bad_flag, 0, sizeof(int), st));
  if (num_cubes == 0) return CC_OK;
  const int threads = 256;
  const int64_t blocks = ceil_div<int64_t>(num_cubes * 32, threads);
  bitpack_kernel<<<(unsigned)blocks, threads, 0, st>>>(indptr, indices, num_cubes, num_cards, bits, cpad, bad_flag);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

int cc_cooc_count(const uint32_t* bits, int64_t num_cubes, int32_t num_cards, int32_t* counts, int64_t ld,
                  int accumulate, void* stream) {
  CC_NVTX("cc_cooc_count");
  CC_REQUIRE(bits && counts && num_cards > 0 && ld >= num_cards, "cc_cooc_count: bad arguments");
  const int64_t kw = cc_bits_words(num_cubes), cpad = cc_bits_cpad(num_cards);
  const int tiles = int(cpad / TILE);
  cudaStream_t st = as_stream(stream);
  if (kw == 0) {
    if (!accumulate) CC_CHECK_CUDA(cudaMemset2DAsync(counts, ld * 4, 0, size_t(num_cards) * 4, num_cards, st));
    return CC_OK;
  }
  dim3 grid(tiles, tiles);
  cooc_count_kernel<<<grid, COUNT_THREADS, 0, st>>>(bits, kw, cpad, num_cards, counts, ld, accumulate);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

int cc_row_normalise(const int32_t* counts, int64_t ld, int32_t num_cards, double* m64, int64_t ld_m,
                     float* mhat, int64_t ld_mhat, double* rowsum, int has_force_diag, double force_diag,
                     void* stream) {
  CC_NVTX("cc_row_normalise");
  return cc_row_normalise_rows(counts, ld, 0, num_cards, num_cards, m64, ld_m, mhat, ld_mhat, rowsum, has_force_diag,
                               force_diag, stream);
}

int cc_row_normalise_rows(const int32_t* counts, int64_t ld, int32_t row0, int32_t nrows, int32_t num_cards, double* m64,
                          int64_t ld_m, float* mhat, int64_t ld_mhat, double* rowsum, int has_force_diag,
                          double force_diag, void* stream) {
  CC_NVTX("cc_row_normalise_rows");
  CC_REQUIRE(counts && num_cards > 0 && ld >= num_cards, "cc_row_normalise: bad arguments");
  CC_REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= num_cards, "cc_row_normalise: row block [%d, %d) outside [0, %d)",
             row0, row0 + nrows, num_cards);
  if (nrows == 0) return CC_OK;
  CC_REQUIRE(!m64 || ld_m >= num_cards, "cc_row_normalise: ld_m too small");
  CC_REQUIRE(!mhat || ld_mhat >= num_cards, "cc_row_normalise: ld_mhat too small");
  // 16-byte vector path when every row of every buffer starts 16-byte aligned
  const bool vec = ld % 4 == 0 && (reinterpret_cast<uintptr_t>(counts) & 15) == 0 &&
                   (!m64 || (ld_m % 2 == 0 && (reinterpret_cast<uintptr_t>(m64) & 15) == 0)) &&
                   (!mhat || (ld_mhat % 4 == 0 && (reinterpret_cast<uintptr_t>(mhat) & 15) == 0));
  if (vec)
    row_normalise_kernel<true><<<nrows, 256, 0, as_stream(stream)>>>(counts, ld, num_cards, m64, ld_m, mhat, ld_mhat,
                                                                    rowsum, has_force_diag, force_diag, row0);
  else
    row_normalise_kernel<false><<<nrows, 256, 0, as_stream(stream)>>>(counts, ld, num_cards, m64, ld_m, mhat,
                                                                     ld_mhat, rowsum, has_force_diag, force_diag, row0);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

int64_t cc_col_mass_workspace_bytes(int32_t num_cards) {
  return int64_t(ceil_div(num_cards, COL_ROWS)) * num_cards * int64_t(sizeof(double));
}

int cc_col_mass(const int32_t* counts, int64_t ld, int32_t num_cards, const double* rowsum, double* workspace,
                double* neg_sampler, void* stream) {
  CC_NVTX("cc_col_mass");
  int rc = cc_col_mass_rows(counts, ld, 0, num_cards, num_cards, rowsum, workspace, neg_sampler, stream);
  if (rc != CC_OK) return rc;
  return cc_col_mass_scale(neg_sampler, num_cards, stream);
}

// Row-block form for a reduce-scattered build: col_mass[j] = sum over the block's rows of M-hat[i][j], NOT yet
// normalised -- the caller sums the blocks' vectors over the ranks (all_reduce of C doubles), then cc_col_mass_scale.
int cc_col_mass_rows(const int32_t* counts, int64_t ld, int32_t row0, int32_t nrows, int32_t num_cards, const double* rowsum,
                     double* workspace, double* col_mass, void* stream) {
  CC_NVTX("cc_col_mass_rows");
  CC_REQUIRE(counts && rowsum && workspace && col_mass, "cc_col_mass: null pointer");
  CC_REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= num_cards, "cc_col_mass: row block outside the matrix");
  cudaStream_t st = as_stream(stream);
  if (nrows == 0) { CC_CHECK_CUDA(cudaMemsetAsync(col_mass, 0, size_t(num_cards) * sizeof(double), st)); return CC_OK; }
  const int chunks = ceil_div(nrows, COL_ROWS);
  dim3 grid(ceil_div(num_cards, 128), chunks);
  col_mass_partial_kernel<<<grid, 128, 0, st>>>(counts, ld, num_cards, rowsum, workspace, row0, nrows);
  CC_CHECK_LAUNCH();
  col_mass_final_kernel<<<ceil_div(num_cards, 256), 256, 0, st>>>(workspace, chunks, num_cards, col_mass);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

int cc_col_mass_scale(double* col_mass, int32_t num_cards, void* stream) {
  CC_NVTX("cc_col_mass_scale");
  CC_REQUIRE(col_mass && num_cards > 0, "cc_col_mass_scale: bad arguments");
  col_mass_scale_kernel<<<1, 1024, 0, as_stream(stream)>>>(col_mass, num_cards);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

// Device -> pageable host delivery of a large result (3.5 GB of float64 at C = 21k).  A plain cudaMemcpy into freshly
// allocated pageable memory is bound by first-touch page faults and the driver's single staging path (0.8-3 s
// measured); here T host threads each own a slice: D2H into their own pinned staging buffer on their own stream,
// then memcpy into the destination -- the page faults and the copies run in parallel.
static int copy_to_host_parallel(void* dst_host, const void* src_dev, size_t bytes) {
  if (bytes == 0) return CC_OK;
  unsigned hw = std::thread::hardware_concurrency();
  int T = int(hw ? (hw < 8 ? hw : 8) : 4);
  const size_t CH = size_t(16) << 20;                       // staging chunk per thread
  if (bytes < 4 * CH) T = 1;
  int dev = 0;
  CC_CHECK_CUDA(cudaGetDevice(&dev));
  std::vector<int> rcs(T, 0);
  std::vector<std::thread> th;
  const size_t per = ((bytes + T - 1) / T + 255) & ~size_t(255);
  for (int t = 0; t < T; ++t) {
    th.emplace_back([&, t]() {
      const size_t lo = size_t(t) * per, hi = lo + per < bytes ? lo + per : bytes;
      if (lo >= hi) return;
      cudaSetDevice(dev);
      cudaStream_t st = nullptr; void* pin[2] = {nullptr, nullptr}; cudaEvent_t ev[2] = {nullptr, nullptr};
      auto fail = [&](cudaError_t e) { rcs[t] = int(e) ? int(e) : -1; };
      cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
      for (int b = 0; b < 2 && e == cudaSuccess; ++b) { e = cudaHostAlloc(&pin[b], CH, cudaHostAllocDefault); if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming); }
      if (e != cudaSuccess) fail(e);
      // two staging buffers: the DMA of chunk i+1 overlaps the memcpy of chunk i
      size_t off = lo; int b = 0; size_t pend_off[2] = {0, 0}, pend_n[2] = {0, 0}; bool pend[2] = {false, false};
      while (rcs[t] == 0 && (off < hi || pend[0] || pend[1])) {
        if (pend[b]) {
          e = cudaEventSynchronize(ev[b]);
          if (e != cudaSuccess) { fail(e); break; }
          memcpy(static_cast<char*>(dst_host) + pend_off[b], pin[b], pend_n[b]);
          pend[b] = false;
        }
        if (off < hi) {
          const size_t n = hi - off < CH ? hi - off : CH;
          e = cudaMemcpyAsync(pin[b], static_cast<const char*>(src_dev) + off, n, cudaMemcpyDeviceToHost, st);
          if (e == cudaSuccess) e = cudaEventRecord(ev[b], st);
          if (e != cudaSuccess) { fail(e); break; }
          pend_off[b] = off; pend_n[b] = n; pend[b] = true; off += n;
        }
        b ^= 1;
      }
      if (st) cudaStreamSynchronize(st);
      for (int q = 0; q < 2; ++q) { if (pin[q]) cudaFreeHost(pin[q]); if (ev[q]) cudaEventDestroy(ev[q]); }
      if (st) cudaStreamDestroy(st);
    });
  }
  for (auto& x : th) x.join();
  for (int t = 0; t < T; ++t)
    if (rcs[t] != 0) { set_error("copy_to_host_parallel: CUDA error %d in worker %d", rcs[t], t); return CC_ERR_CUDA; }
  return CC_OK;
}

// Host-buffer entry: the drop-in for utils.create_adjacency_matrix (utils.py:75-92) on
// CSR cubes.  Copies H2D, builds on `device`'s current context, copies M back (float64).
int cc_create_adjacency_matrix_host(const int64_t* indptr_host, const int32_t* indices_host, int64_t num_cubes,
                                    int32_t num_cards, int has_force_diag, double force_diag, double* m_host,
                                    int32_t* counts_host /* nullable */) {
  CC_REQUIRE(indptr_host && indices_host && m_host, "cc_create_adjacency_matrix_host: null pointer");
  CC_REQUIRE(num_cards > 0 && num_cubes >= 0, "cc_create_adjacency_matrix_host: bad sizes");
  const int64_t nnz = indptr_host[num_cubes];
  const int64_t kw = cc_bits_words(num_cubes), cpad = cc_bits_cpad(num_cards);
  int64_t* d_indptr = nullptr; int32_t* d_indices = nullptr; uint32_t* d_bits = nullptr;
  int32_t* d_counts = nullptr; double* d_m = nullptr; int* d_bad = nullptr;
  // tensor-core contraction when the counts rows meet TMA's 16-byte rule, bit-packed popcount tiles otherwise
  const bool tensor = num_cards % 4 == 0;
  const int64_t ws_bytes = tensor ? cc_cooc_tc_workspace_bytes(num_cubes, num_cards) : size_t(kw > 0 ? kw : 1) * cpad * 4;
  int rc = CC_OK;
  cudaStream_t st = nullptr;
#define CC_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { rc = cuda_fail(_e, #expr, __FILE__, __LINE__); goto done; } } while (0)
  CC_TRY(cudaStreamCreate(&st));
  CC_TRY(cudaMalloc(&d_indptr, size_t(num_cubes + 1) * 8));
  CC_TRY(cudaMalloc(&d_indices, size_t(nnz > 0 ? nnz : 1) * 4));
  CC_TRY(cudaMalloc(&d_bits, size_t(ws_bytes)));
  CC_TRY(cudaMalloc(&d_counts, size_t(num_cards) * num_cards * 4));
  CC_TRY(cudaMalloc(&d_m, size_t(num_cards) * num_cards * 8));
  CC_TRY(cudaMalloc(&d_bad, sizeof(int)));
  CC_TRY(cudaMemcpyAsync(d_indptr, indptr_host, size_t(num_cubes + 1) * 8, cudaMemcpyHostToDevice, st));
  CC_TRY(cudaMemcpyAsync(d_indices, indices_host, size_t(nnz) * 4, cudaMemcpyHostToDevice, st));
  if (tensor) {
    CC_TRY(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    if ((rc = cc_cooc_count_tc(d_indptr, d_indices, num_cubes, num_cards, d_bits, ws_bytes, d_counts, num_cards, 0, d_bad,
                               st)) != CC_OK) goto done;
  } else {
    if ((rc = cc_bitpack_cubes(d_indptr, d_indices, num_cubes, num_cards, d_bits, d_bad, st)) != CC_OK) goto done;
    if ((rc = cc_cooc_count(d_bits, num_cubes, num_cards, d_counts, num_cards, 0, st)) != CC_OK) goto done;
  }
  if ((rc = cc_row_normalise(d_counts, num_cards, num_cards, d_m, num_cards, nullptr, 0, nullptr, has_force_diag,
                             force_diag, st)) != CC_OK) goto done;
  {
    int bad = 0;
    CC_TRY(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    CC_TRY(cudaStreamSynchronize(st));
    if (bad) { set_error("cc_create_adjacency_matrix_host: card index out of range [0,%d)", num_cards); rc = CC_ERR_ARGUMENT; goto done; }
  }
  if ((rc = copy_to_host_parallel(m_host, d_m, size_t(num_cards) * num_cards * 8)) != CC_OK) goto done;
  if (counts_host && (rc = copy_to_host_parallel(counts_host, d_counts, size_t(num_cards) * num_cards * 4)) != CC_OK) goto done;
done:
#undef CC_TRY
  cudaFree(d_indptr); cudaFree(d_indices); cudaFree(d_bits); cudaFree(d_counts); cudaFree(d_m); cudaFree(d_bad);
  if (st) cudaStreamDestroy(st);
  return rc;
}

}  // extern "C"
