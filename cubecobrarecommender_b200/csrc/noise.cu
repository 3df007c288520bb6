// Noise function F: 1->0 flips plus frequency-proportional negative sampling.
//
// Replaces reference src/ml/generator.py:74-103 (DataGenerator.generate_data) and the
// reg-row draw of generator.py:47-51.  Counter-based Philox4x32-10 per draw (no state),
// and an alias table of `neg_sampler` (generator.py:30): drawing from
// neg_sampler[excludes]/sum (generator.py:93-94) is exactly "draw from the full
// neg_sampler, reject cards that are in the cube".
//
// Per cube (one CTA):
//   noise ~ clip(N(noise, std), 0.05, 0.8)              generator.py:86-90
//   flip  = int(size * noise)                            generator.py:91
//   flip_include   = flip uniform draws (with replacement) from the cube    :92
//   flip_exclude   = flip draws from the cards outside the cube, p ~ neg    :93-94
//   y_flip_include = flip//4 uniform draws from the flip_include *array*    :95
//   x = cube - 1[flip_include] + 1[flip_exclude];  y = cube - 1[y_flip_include]   :96-101
// Outputs are sparse: x as a sorted index list (kept cards, then added cards), y as a
// bit row (bit c of y_bits[b][c>>5]) for the fused sigmoid-BCE epilogue.
#include "cc_common.cuh"

namespace cc {

struct Philox {
  uint32_t c[4];
  __device__ __forceinline__ Philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    c[0] = c0; c[1] = c1; c[2] = c2; c[3] = c3;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
      const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
      c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
  }
};

__device__ __forceinline__ float u01(uint32_t x) { return (float(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }  // (0,1)

enum : uint32_t { STREAM_NOISE = 1, STREAM_INCLUDE = 2, STREAM_YFLIP = 3, STREAM_REG = 4, STREAM_EXCLUDE = 16 };

__device__ __forceinline__ int alias_draw(const float* __restrict__ prob, const int32_t* __restrict__ alias,
                                          int32_t n, uint32_t r0, uint32_t r1) {
  const int i = int((uint64_t(r0) * uint64_t(n)) >> 32);
  return (u01(r1) < prob[i]) ? i : alias[i];
}

// ordered block compaction helper: returns the exclusive offset of this thread's `cnt`
// items inside the block-wide sequence (thread order), and the block total.
__device__ __forceinline__ int block_excl_scan(int cnt, int* warp_tot /* smem[32] */, int& total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  int base = 0, tot = 0;
  for (int w = 0; w < nw; ++w) { const int t = warp_tot[w]; if (w < wid) base += t; tot += t; }
  __syncthreads();
  total = tot;
  return base + inc - cnt;
}

constexpr int NOISE_THREADS = 128;

__global__ void __launch_bounds__(NOISE_THREADS)
noise_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
             const int32_t* __restrict__ batch_ids, int32_t batch, int32_t num_cards,
             const float* __restrict__ alias_prob, const int32_t* __restrict__ alias_idx,
             float noise_mean, float noise_std, uint64_t seed, const int64_t* __restrict__ step_ptr,
             int32_t max_size, int32_t x_stride,
             int32_t* __restrict__ x_idx, int32_t* __restrict__ x_len, uint32_t* __restrict__ y_bits,
             int64_t y_words, int32_t* __restrict__ flips_out, int* __restrict__ overflow,
             float* __restrict__ x_dense, int64_t ld_dense, int dense_bf16) {
  extern __shared__ uint32_t sm[];
  const int W = (num_cards + 31) >> 5;
  const int FW = (max_size + 31) >> 5;
  uint32_t* cube_mask = sm;            // W   cards in the cube
  uint32_t* y_mask = cube_mask + W;    // W   cards in y
  uint32_t* add_mask = y_mask + W;     // W   cards added to x
  uint32_t* rem_flag = add_mask + W;   // FW  positions removed from x
  int32_t* flip_pos = reinterpret_cast<int32_t*>(rem_flag + FW);  // max flips (<= 0.8*max_size)
  __shared__ int warp_tot[32];

  const int b = blockIdx.x;
  if (b >= batch) return;
  const int64_t cube = batch_ids ? int64_t(batch_ids[b]) : int64_t(b);
  const int64_t beg = indptr[cube];
  const int s = int(indptr[cube + 1] - beg);
  const int32_t* inc = indices + beg;
  const uint64_t step = step_ptr ? uint64_t(*step_ptr) : 0ull;
  const uint32_t k0 = uint32_t(seed), k1 = uint32_t(seed >> 32);
  const uint32_t c1 = uint32_t(cube), c2 = uint32_t(step) ^ (uint32_t(cube >> 32) << 16), c3hi = uint32_t(step >> 32) << 8;

  for (int w = threadIdx.x; w < 3 * W + FW; w += blockDim.x) sm[w] = 0;
  __syncthreads();
  if (s > max_size) {
    // a cube larger than the engine was sized for: flag it AND leave a well-defined empty row behind (x empty, y and
    // the dense row all zero) -- never the previous step's data -- so a caller that only checks the flag at the end of
    // an epoch has trained on nothing worse than an empty cube meanwhile
    if (threadIdx.x == 0) { atomicExch(overflow, 1); x_len[b] = 0; if (flips_out) flips_out[b] = 0; }
    if (y_bits) {
      uint32_t* yo = y_bits + int64_t(b) * y_words;
      for (int w = threadIdx.x; w < y_words; w += blockDim.x) yo[w] = 0u;
    }
    if (x_dense) {
      const int groups = int(ld_dense >> 2);
      if (dense_bf16) {
        uint2* xo2 = reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(x_dense) + int64_t(b) * ld_dense);
        for (int g = threadIdx.x; g < groups; g += blockDim.x) xo2[g] = make_uint2(0u, 0u);
      } else {
        float4* xo4 = reinterpret_cast<float4*>(x_dense + int64_t(b) * ld_dense);
        for (int g = threadIdx.x; g < groups; g += blockDim.x) xo4[g] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    return;
  }
  for (int p = threadIdx.x; p < s; p += blockDim.x) {
    const int32_t c = inc[p];
    atomicOr(&cube_mask[c >> 5], 1u << (c & 31));
  }
  // per-cube noise level (every thread evaluates the same counter)
  Philox pn(k0, k1, 0u, c1, c2, c3hi | STREAM_NOISE);
  const float z = sqrtf(-2.0f * logf(u01(pn.c[0]))) * cospif(2.0f * u01(pn.c[1]));
  double nz = double(noise_mean) + double(noise_std) * double(z);
  nz = fmin(fmax(nz, 0.05), 0.8);
  const int flip = int(double(s) * nz);
  __syncthreads();
  for (int w = threadIdx.x; w < W; w += blockDim.x) y_mask[w] = cube_mask[w];
  // flip_include: uniform with replacement over the cube's positions
  for (int d = threadIdx.x; d < flip; d += blockDim.x) {
    Philox pr(k0, k1, uint32_t(d), c1, c2, c3hi | STREAM_INCLUDE);
    const int pos = int((uint64_t(pr.c[0]) * uint64_t(s)) >> 32);
    flip_pos[d] = pos;
    atomicOr(&rem_flag[pos >> 5], 1u << (pos & 31));
  }
  __syncthreads();
  // y_flip_include: flip//4 uniform draws from the flip_include array
  for (int d = threadIdx.x; d < flip / 4; d += blockDim.x) {
    Philox pr(k0, k1, uint32_t(d), c1, c2, c3hi | STREAM_YFLIP);
    const int e = int((uint64_t(pr.c[0]) * uint64_t(flip)) >> 32);
    const int32_t c = inc[flip_pos[e]];
    atomicAnd(&y_mask[c >> 5], ~(1u << (c & 31)));
  }
  // flip_exclude: alias draw from neg_sampler, rejecting cards in the cube
  for (int d = threadIdx.x; d < flip; d += blockDim.x) {
    for (uint32_t attempt = 0; attempt < 4096u; ++attempt) {
      Philox pr(k0, k1, uint32_t(d), c1, c2, c3hi | (STREAM_EXCLUDE + attempt));
      const int c = alias_draw(alias_prob, alias_idx, num_cards, pr.c[0], pr.c[1]);
      if ((cube_mask[c >> 5] >> (c & 31)) & 1u) continue;
      atomicOr(&add_mask[c >> 5], 1u << (c & 31));
      break;
    }
  }
  __syncthreads();
  // ---- emit x: kept cards (cube order), then added cards (ascending id) ----
  int32_t* xo = x_idx + int64_t(b) * x_stride;
  int written = 0;
  for (int base = 0; base < s; base += blockDim.x) {
    const int p = base + threadIdx.x;
    const int keep = (p < s) && !((rem_flag[p >> 5] >> (p & 31)) & 1u);
    int tot;
    const int off = block_excl_scan(keep, warp_tot, tot);
    if (keep && written + off < x_stride) xo[written + off] = inc[p];
    written += tot;
  }
  for (int base = 0; base < W; base += blockDim.x) {
    const int w = base + threadIdx.x;
    uint32_t bits = (w < W) ? add_mask[w] : 0u;
    int tot;
    int off = written + block_excl_scan(__popc(bits), warp_tot, tot);
    while (bits) {
      const int bpos = __ffs(bits) - 1;
      bits &= bits - 1;
      if (off < x_stride) xo[off] = (w << 5) + bpos;
      ++off;
    }
    written += tot;
  }
  if (threadIdx.x == 0) {
    x_len[b] = min(written, x_stride);
    if (written > x_stride) atomicExch(overflow, 2);
    if (flips_out) flips_out[b] = flip;
  }
  if (y_bits) {
    uint32_t* yo = y_bits + int64_t(b) * y_words;
    for (int w = threadIdx.x; w < y_words; w += blockDim.x) yo[w] = (w < W) ? y_mask[w] : 0u;
  }
  if (x_dense) {
    // dense 0/1 row of x (the MN-major operand of the tensor-core dW1 = x^T g1 GEMM): x = (cube & ~removed) | added.
    // removed cards are cleared from cube_mask first (positions -> cards), then the row is expanded 4 columns a store
    for (int p = threadIdx.x; p < s; p += blockDim.x)
      if ((rem_flag[p >> 5] >> (p & 31)) & 1u) atomicAnd(&cube_mask[inc[p] >> 5], ~(1u << (inc[p] & 31)));
    __syncthreads();
    const int groups = int(ld_dense >> 2);
    if (dense_bf16) {      // the same 0/1 rows as bf16 (1.0 = 0x3F80): 8 bytes per group of four cards
      uint2* xo2 = reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(x_dense) + int64_t(b) * ld_dense);
      for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        const int w = g >> 3, sh = (g & 7) << 2;
        const uint32_t bits = (w < W) ? ((cube_mask[w] | add_mask[w]) >> sh) & 0xfu : 0u;
        xo2[g] = make_uint2(((bits & 1u) ? 0x3F80u : 0u) | ((bits & 2u) ? 0x3F800000u : 0u),
                            ((bits & 4u) ? 0x3F80u : 0u) | ((bits & 8u) ? 0x3F800000u : 0u));
      }
    } else {
      float4* xo4 = reinterpret_cast<float4*>(x_dense + int64_t(b) * ld_dense);
      for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        const int w = g >> 3, sh = (g & 7) << 2;
        const uint32_t bits = (w < W) ? ((cube_mask[w] | add_mask[w]) >> sh) & 0xfu : 0u;
        xo4[g] = make_float4(float(bits & 1u), float((bits >> 1) & 1u), float((bits >> 2) & 1u), float((bits >> 3) & 1u));
      }
    }
  }
}

// reg_indices = choice(C, n, p=neg_sampler) (generator.py:47-51), with replacement
__global__ void reg_rows_kernel(const float* __restrict__ alias_prob, const int32_t* __restrict__ alias_idx,
                                int32_t num_cards, int32_t n, uint64_t seed, const int64_t* __restrict__ step_ptr,
                                int32_t* __restrict__ rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t step = step_ptr ? uint64_t(*step_ptr) : 0ull;
  Philox pr(uint32_t(seed), uint32_t(seed >> 32), uint32_t(i), 0xC0FFEEu, uint32_t(step),
            (uint32_t(step >> 32) << 8) | STREAM_REG);
  rows[i] = alias_draw(alias_prob, alias_idx, num_cards, pr.c[0], pr.c[1]);
}

// cube rows -> bit rows (the noise-free y / in-cube mask for inference)
__global__ void cubes_to_bits_kernel(const int32_t* __restrict__ idx, const int64_t* __restrict__ row_start,
                                     const int32_t* __restrict__ row_len, int32_t num_cards,
                                     uint32_t* __restrict__ bits, int64_t words) {
  const int b = blockIdx.x;
  uint32_t* o = bits + int64_t(b) * words;
  const int32_t* r = idx + row_start[b];
  for (int p = threadIdx.x; p < row_len[b]; p += blockDim.x) {
    const int32_t c = r[p];
    if (c >= 0 && c < num_cards) atomicOr(&o[c >> 5], 1u << (c & 31));
  }
}

__global__ void step_increment_kernel(int64_t* step) { *step += 1; }

}  // namespace cc

using namespace cc;

extern "C" {

// Vose alias table of a float64 probability vector (host helper; O(n)).
int cc_alias_build_host(const double* p, int32_t n, float* prob, int32_t* alias) {
  CC_REQUIRE(p && prob && alias && n > 0, "cc_alias_build_host: bad arguments");
  double total = 0.0;
  for (int i = 0; i < n; ++i) { CC_REQUIRE(p[i] >= 0.0, "cc_alias_build_host: negative probability"); total += p[i]; }
  CC_REQUIRE(total > 0.0, "cc_alias_build_host: zero mass");
  double* scaled = new double[n];
  int32_t* small = new int32_t[n];
  int32_t* large = new int32_t[n];
  int ns = 0, nl = 0;
  for (int i = 0; i < n; ++i) {
    scaled[i] = p[i] / total * n;
    if (scaled[i] < 1.0) small[ns++] = i; else large[nl++] = i;
  }
  while (ns > 0 && nl > 0) {
    const int s = small[--ns], l = large[--nl];
    prob[s] = float(scaled[s]); alias[s] = l;
    scaled[l] = (scaled[l] + scaled[s]) - 1.0;
    if (scaled[l] < 1.0) small[ns++] = l; else large[nl++] = l;
  }
  while (nl > 0) { const int l = large[--nl]; prob[l] = 1.0f; alias[l] = l; }
  while (ns > 0) { const int s = small[--ns]; prob[s] = 1.0f; alias[s] = s; }
  delete[] scaled; delete[] small; delete[] large;
  return CC_OK;
}

int64_t cc_noise_smem_bytes(int32_t num_cards, int32_t max_size) {
  const int64_t W = (num_cards + 31) / 32, FW = (max_size + 31) / 32;
  return (3 * W + FW) * 4 + int64_t(max_size) * 4;
}

int cc_noise(const int64_t* indptr, const int32_t* indices, const int32_t* batch_ids, int32_t batch,
             int32_t num_cards, const float* alias_prob, const int32_t* alias_idx, float noise_mean, float noise_std,
             uint64_t seed, const int64_t* step_ptr, int32_t max_size, int32_t x_stride, int32_t* x_idx,
             int32_t* x_len, uint32_t* y_bits, int64_t y_words, int32_t* flips_out, int* overflow_flag,
             float* x_dense, int64_t ld_dense, void* stream) {
  CC_NVTX("cc_noise");
  return cc_noise_ex(indptr, indices, batch_ids, batch, num_cards, alias_prob, alias_idx, noise_mean, noise_std, seed, step_ptr,
                     max_size, x_stride, x_idx, x_len, y_bits, y_words, flips_out, overflow_flag, x_dense, ld_dense, 0, stream);
}

int cc_noise_ex(const int64_t* indptr, const int32_t* indices, const int32_t* batch_ids, int32_t batch,
                int32_t num_cards, const float* alias_prob, const int32_t* alias_idx, float noise_mean, float noise_std,
                uint64_t seed, const int64_t* step_ptr, int32_t max_size, int32_t x_stride, int32_t* x_idx,
                int32_t* x_len, uint32_t* y_bits, int64_t y_words, int32_t* flips_out, int* overflow_flag,
                void* x_dense_v, int64_t ld_dense, int dense_bf16, void* stream) {
  CC_NVTX("cc_noise_ex");
  float* x_dense = static_cast<float*>(x_dense_v);
  CC_REQUIRE(indptr && indices && alias_prob && alias_idx && x_idx && x_len && overflow_flag, "cc_noise: null pointer");
  CC_REQUIRE(!x_dense || (ld_dense % 4 == 0 && ld_dense >= num_cards && (reinterpret_cast<uintptr_t>(x_dense) & 15) == 0),
             "cc_noise: x_dense needs ld_dense % 4 == 0, ld_dense >= num_cards and a 16-byte aligned base");
  CC_REQUIRE(batch >= 0 && num_cards > 0 && max_size > 0 && x_stride > 0, "cc_noise: bad sizes");
  CC_REQUIRE(!y_bits || y_words * 32 >= num_cards, "cc_noise: y_words too small");
  if (batch == 0) return CC_OK;
  const int64_t smem = cc_noise_smem_bytes(num_cards, max_size);
  CC_REQUIRE(smem <= 200 * 1024, "cc_noise: C=%d max_size=%d need %lld bytes of shared memory", num_cards, max_size,
             (long long)smem);
  CC_CHECK_CUDA(cudaFuncSetAttribute(noise_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  noise_kernel<<<batch, NOISE_THREADS, smem, as_stream(stream)>>>(
      indptr, indices, batch_ids, batch, num_cards, alias_prob, alias_idx, noise_mean, noise_std, seed, step_ptr,
      max_size, x_stride, x_idx, x_len, y_bits, y_words, flips_out, overflow_flag, x_dense, ld_dense, dense_bf16);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

int cc_sample_reg_rows(const float* alias_prob, const int32_t* alias_idx, int32_t num_cards, int32_t n, uint64_t seed,
                       const int64_t* step_ptr, int32_t* rows, void* stream) {
  CC_NVTX("cc_sample_reg_rows");
  CC_REQUIRE(alias_prob && alias_idx && rows && num_cards > 0 && n >= 0, "cc_sample_reg_rows: bad arguments");
  if (n == 0) return CC_OK;
  reg_rows_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(alias_prob, alias_idx, num_cards, n, seed,
                                                                  step_ptr, rows);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

int cc_cubes_to_bits(const int32_t* idx, const int64_t* row_start, const int32_t* row_len, int32_t batch,
                     int32_t num_cards, uint32_t* bits, int64_t words, void* stream) {
  CC_NVTX("cc_cubes_to_bits");
  CC_REQUIRE(idx && row_start && row_len && bits && words * 32 >= num_cards, "cc_cubes_to_bits: bad arguments");
  if (batch == 0) return CC_OK;
  cudaStream_t st = as_stream(stream);
  CC_CHECK_CUDA(cudaMemsetAsync(bits, 0, size_t(batch) * words * 4, st));
  cubes_to_bits_kernel<<<batch, 128, 0, st>>>(idx, row_start, row_len, num_cards, bits, words);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

int cc_step_increment(int64_t* step_ptr, void* stream) {
  CC_NVTX("cc_step_increment");
  CC_REQUIRE(step_ptr, "cc_step_increment: null pointer");
  step_increment_kernel<<<1, 1, 0, as_stream(stream)>>>(step_ptr);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

}  // extern "C"
