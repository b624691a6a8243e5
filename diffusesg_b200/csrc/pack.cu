// Weight packing kernels: run once per weight update (dsg_model_finalize), never on the per-step path.
#include "common.cuh"
#include "kernels.h"

namespace dsg {
namespace {

__global__ void pack_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long numel,
                                 long long n_scaled, float scale) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < numel;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dst[i] = __float2bfloat16_rn(i < n_scaled ? src[i] * scale : src[i]);
}

__global__ void scale_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, long long numel,
                                  long long n_scaled, float scale) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < numel;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dst[i] = i < n_scaled ? src[i] * scale : src[i];
}

__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int ld, int col0,
                                 int ncols, int dst_pitch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * ncols) return;
  const int c = i / R, r = i - c * R;
  dst[c * dst_pitch + r] = src[static_cast<size_t>(r) * ld + col0 + c];
}

__global__ void bias_expand_kernel(const float* __restrict__ table, const int64_t* __restrict__ index,
                                   float* __restrict__ out, int TT, int heads, int table_rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= TT * heads) return;
  const int h = i / TT, pq = i - h * TT;
  long long row = index[pq];
  if (row < 0) row = 0;
  if (row >= table_rows) row = table_rows - 1;
  out[i] = table[row * heads + h];
}

__global__ void small_mm_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int n,
                                int trans_b) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * n) return;
  const int o = idx / n, i = idx - o * n;
  float s = 0.f;
  for (int k = 0; k < n; ++k) s = fmaf(A[o * n + k], trans_b ? B[i * n + k] : B[k * n + i], s);
  C[idx] = s;
}

__global__ void small_mv_kernel(const float* __restrict__ A, const float* __restrict__ x, const float* __restrict__ b,
                                float* __restrict__ y, int n) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n) return;
  float s = 0.f;
  for (int k = 0; k < n; ++k) s = fmaf(A[o * n + k], x[k], s);
  y[o] = s + b[o];
}

// staleness probe: sets *flag when any 32-bit word of a differs from b (never clears it)
__global__ void compare_words_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, long long words,
                                     int* __restrict__ flag) {
  bool diff = false;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < words;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    diff |= a[i] != b[i];
  if (__any_sync(0xffffffffu, diff) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

unsigned blocks_for(long long n, int per = 256, long long cap = 148 * 8) {
  long long b = (n + per - 1) / per;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<unsigned>(b);
}

}  // namespace

int launch_compare_words(const void* a, const void* b, int64_t words, int* flag, cudaStream_t st) {
  if (words <= 0) return DSG_OK;
  compare_words_kernel<<<blocks_for(words), 256, 0, st>>>(static_cast<const uint32_t*>(a), static_cast<const uint32_t*>(b),
                                                          words, flag);
  DSG_LAUNCH_CHECK();
  return DSG_OK;
}

int launch_pack_bf16(const float* src, bf16* dst, int64_t numel, int64_t n_scaled, float scale, cudaStream_t st) {
  pack_bf16_kernel<<<blocks_for(numel), 256, 0, st>>>(src, dst, numel, n_scaled, scale);
  DSG_LAUNCH_CHECK();
  return DSG_OK;
}

int launch_scale_copy(const float* src, float* dst, int64_t numel, int64_t n_scaled, float scale, cudaStream_t st) {
  scale_copy_kernel<<<blocks_for(numel), 256, 0, st>>>(src, dst, numel, n_scaled, scale);
  DSG_LAUNCH_CHECK();
  return DSG_OK;
}

int launch_transpose(const float* src, float* dst, int R, int ld, int col0, int ncols, int dst_pitch, cudaStream_t st) {
  transpose_kernel<<<(R * ncols + 255) / 256, 256, 0, st>>>(src, dst, R, ld, col0, ncols, dst_pitch);
  DSG_LAUNCH_CHECK();
  return DSG_OK;
}

int launch_bias_expand(const float* table, const int64_t* index, float* out, int T, int heads, int table_rows,
                       cudaStream_t st) {
  const int total = T * T * heads;
  bias_expand_kernel<<<(total + 255) / 256, 256, 0, st>>>(table, index, out, T * T, heads, table_rows);
  DSG_LAUNCH_CHECK();
  return DSG_OK;
}

int launch_small_mm(const float* A, const float* B, float* C, int n, int trans_b, cudaStream_t st) {
  small_mm_kernel<<<(n * n + 255) / 256, 256, 0, st>>>(A, B, C, n, trans_b);
  DSG_LAUNCH_CHECK();
  return DSG_OK;
}

int launch_small_mv(const float* A, const float* x, const float* b, float* y, int n, cudaStream_t st) {
  small_mv_kernel<<<(n + 127) / 128, 128, 0, st>>>(A, x, b, y, n);
  DSG_LAUNCH_CHECK();
  return DSG_OK;
}

}  // namespace dsg
