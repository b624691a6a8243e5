// Microbenchmark: TMA gather rate of window boxes from the [B res, res, 3C] qkv view as a function of the inner (channel)
// extent of the box: 64 B (one head, what the attention kernels load), 128 B (two heads), and of the box shape.
// Prints bytes / clk / SM and GB/s.  Question: is a CTA's TMA rate bound per 64-byte row rather than per byte?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../diffusesg_b200/csrc -o tma_rows_bench tma_rows_bench.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace dsg {
void set_last_error(const char*, ...) {}
void count_launch(int) {}
}  // namespace dsg
using namespace dsg;

DSG_DEVICE void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

constexpr int kStages = 6;

// each CTA walks boxes: box index -> (head group, window); one elected thread issues, the same thread waits (ring)
__global__ void __launch_bounds__(64, 1)
bench(const __grid_constant__ CUtensorMap tm, int box_bytes, int inner_elems, int groups, int nwx, int wtok, int n_boxes,
      int boxes_per_stage, int lanes, long long* clk_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t full[kStages];
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (lanes > 1 && threadIdx.x < 32) {
    // the boxes of a stage are issued by `lanes` lanes in parallel (lane j takes boxes j, j + lanes, ...)
    const int lane = threadIdx.x;
    const long long t0 = clock64();
    const int per_cta = n_boxes / gridDim.x;
    const int n_it = per_cta / boxes_per_stage;
    for (int it = 0; it < n_it + kStages; ++it) {
      if (it >= kStages) mbar_wait(&full[it % kStages], ((it / kStages) - 1) & 1);
      if (it < n_it) {
        const int st = it % kStages;
        if (lane == 0) mbar_expect_tx(&full[st], box_bytes * boxes_per_stage);
        __syncwarp();
        for (int j = lane; j < boxes_per_stage; j += lanes) {
          if (lane < lanes) {
            const int box = (it * boxes_per_stage + j) * gridDim.x + blockIdx.x;
            const int g = box % groups, win = box / groups;
            const int wx = win % nwx, wy = win / nwx;
            tma_load_3d(smem + (st * boxes_per_stage + j) * box_bytes, &tm, &full[st], g * inner_elems, wx * wtok, wy * wtok);
          }
        }
        __syncwarp();
      }
    }
    if (lane == 0) clk_out[blockIdx.x] = clock64() - t0;
  } else if (lanes <= 1 && threadIdx.x == 0) {
    const long long t0 = clock64();
    const int per_cta = n_boxes / gridDim.x;
    const int n_it = per_cta / boxes_per_stage;
    for (int it = 0; it < n_it + kStages; ++it) {
      if (it >= kStages) mbar_wait(&full[it % kStages], ((it / kStages) - 1) & 1);
      if (it < n_it) {
        const int st = it % kStages;
        mbar_expect_tx(&full[st], box_bytes * boxes_per_stage);
        for (int j = 0; j < boxes_per_stage; ++j) {
          const int box = (it * boxes_per_stage + j) * gridDim.x + blockIdx.x;
          const int g = box % groups, win = box / groups;
          const int wx = win % nwx, wy = win / nwx;  // wy runs over B * nwy
          tma_load_3d(smem + (st * boxes_per_stage + j) * box_bytes, &tm, &full[st], g * inner_elems, wx * wtok, wy * wtok);
        }
      }
    }
    clk_out[blockIdx.x] = clock64() - t0;
  }
}

// GEMM-like: every CTA walks `per_cta` boxes of 16 KB (box b -> channel group b % 4, 16 x 8-token tile b / 4 of a small,
// L2-resident tensor); `lanes` lanes of warp 0 issue the boxes of a stage in parallel
__global__ void __launch_bounds__(64, 1)
bench16(const __grid_constant__ CUtensorMap tm, int box_bytes, int distinct, int per_cta, int boxes_per_stage, int lanes,
        long long* clk_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t full[kStages];
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const long long t0 = clock64();
    const int n_it = per_cta / boxes_per_stage;
    for (int it = 0; it < n_it + kStages; ++it) {
      if (it >= kStages) mbar_wait(&full[it % kStages], ((it / kStages) - 1) & 1);
      if (it < n_it) {
        const int st = it % kStages;
        if (lane == 0) mbar_expect_tx(&full[st], box_bytes * boxes_per_stage);
        __syncwarp();
        for (int j = lane; j < boxes_per_stage; j += lanes) {
          if (lane < lanes) {
            const int box = ((it * boxes_per_stage + j) + blockIdx.x * 37) % distinct;
            const int g = box & 3, tile = box >> 2;
            const int tx = tile & 3, ty = tile >> 2;
            tma_load_3d(smem + (st * boxes_per_stage + j) * box_bytes, &tm, &full[st], g * 64, tx * 16, ty * 8);
          }
        }
        __syncwarp();
      }
    }
    if (lane == 0) clk_out[blockIdx.x] = clock64() - t0;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fnp);
  const int B = 512, res = 64, C3 = 288;  // VG first stage: [B res, res, 3C] bf16
  const size_t bytes = static_cast<size_t>(B) * res * res * C3 * 2;
  void* d = nullptr;
  cudaMalloc(&d, bytes);
  cudaMemset(d, 0, bytes);
  long long* d_clk = nullptr;
  cudaMalloc(&d_clk, 148 * sizeof(long long));
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(bench16, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Case { const char* name; int inner; int wtok; CUtensorMapSwizzle sw; int per_stage; };
  const Case cases[] = {
      {"8x8 tokens x 32 ch (64 B rows, SW64)  [pair kernel]", 32, 8, CU_TENSOR_MAP_SWIZZLE_64B, 6},
      {"8x8 tokens x 64 ch (128 B rows, SW128)", 64, 8, CU_TENSOR_MAP_SWIZZLE_128B, 3},
      {"4x4 tokens x 32 ch (64 B rows, SW64)  [shifted]", 32, 4, CU_TENSOR_MAP_SWIZZLE_64B, 24},
      {"4x4 tokens x 64 ch (128 B rows, SW128)", 64, 4, CU_TENSOR_MAP_SWIZZLE_128B, 12},
      {"8x8 tokens x 96 ch (192 B rows, no swizzle)", 96, 8, CU_TENSOR_MAP_SWIZZLE_NONE, 2},
  };
  // ---- L2-resident operands, GEMM-sized boxes (128 rows x 128 B = 16 KB): is one issuing thread the limit, or the
  //      L2 -> SM path?  (tensor of 19 MB, every CTA walks all of it)
  {
    const int Bs = 8;
    CUtensorMap tm;
    const cuuint64_t gdim[3] = {static_cast<cuuint64_t>(C3), static_cast<cuuint64_t>(res), static_cast<cuuint64_t>(Bs) * res};
    const cuuint64_t gstr[2] = {static_cast<cuuint64_t>(C3) * 2, static_cast<cuuint64_t>(C3) * 2 * res};
    const cuuint32_t box[3] = {64, 16, 8};
    const cuuint32_t es[3] = {1, 1, 1};
    enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    for (int lanes : {1, 2, 4, 8}) {
      const int per_stage = 2;                 // two 16 KB boxes per stage, like a GEMM k-block
      const int nwx = res / 16;                // 4 channel groups of 64, 4 x (Bs * 8) boxes of 16 x 8 tokens
      const long long distinct = 4LL * nwx * Bs * (res / 8);
      const long long n_boxes = distinct * 148 / (148LL * per_stage) * (148LL * per_stage) ;  // every CTA ~ all boxes
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      float ms = 0;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        bench16<<<148, 64, 200 * 1024>>>(tm, 16384, static_cast<int>(distinct), static_cast<int>(n_boxes / 148), per_stage, lanes, d_clk);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
      }
      cudaEventElapsedTime(&ms, e0, e1);
      std::vector<long long> clk(148);
      cudaMemcpy(clk.data(), d_clk, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
      long long cmax = 0;
      for (long long v : clk) cmax = v > cmax ? v : cmax;
      const double total = static_cast<double>(n_boxes) * 16384;
      printf("L2-resident 16 KB boxes, %d issuing lanes: %7.1f us  %6.0f GB/s  %5.1f B/clk/SM  %6.1f clk per box per SM (err %s)\n",
             lanes, ms * 1e3, total / ms / 1e6, total / 148 / cmax, cmax / (static_cast<double>(n_boxes) / 148),
             cudaGetErrorString(cudaGetLastError()));
    }
  }
  for (int lanes : {1, 8, 24})
  for (const Case& c : cases) {
    CUtensorMap tm;
    const cuuint64_t gdim[3] = {static_cast<cuuint64_t>(C3), static_cast<cuuint64_t>(res), static_cast<cuuint64_t>(B) * res};
    const cuuint64_t gstr[2] = {static_cast<cuuint64_t>(C3) * 2, static_cast<cuuint64_t>(C3) * 2 * res};
    const cuuint32_t box[3] = {static_cast<cuuint32_t>(c.inner), static_cast<cuuint32_t>(c.wtok), static_cast<cuuint32_t>(c.wtok)};
    const cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", c.name, (int)r); continue; }
    const int box_bytes = c.inner * 2 * c.wtok * c.wtok;
    const int groups = C3 / c.inner;  // channel groups per token (q | k | v heads)
    const int nwx = res / c.wtok;
    const long long windows = static_cast<long long>(B) * nwx * nwx;
    long long n_boxes = windows * groups;
    n_boxes -= n_boxes % (148LL * c.per_stage);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      bench<<<148, 64, 200 * 1024>>>(tm, box_bytes, c.inner, groups, nwx, c.wtok, static_cast<int>(n_boxes), c.per_stage, lanes, d_clk);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> clk(148);
    cudaMemcpy(clk.data(), d_clk, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
    long long cmax = 0;
    for (long long v : clk) cmax = v > cmax ? v : cmax;
    const double total = static_cast<double>(n_boxes) * box_bytes;
    const double rows = static_cast<double>(n_boxes) * c.wtok * c.wtok;
    printf("lanes %2d  %-52s %7.1f us  %6.0f GB/s  %5.1f B/clk/SM  %5.2f clk per token row per SM  (err %s)\n", lanes, c.name, ms * 1e3,
           total / ms / 1e6, total / 148 / cmax, cmax / (rows / 148), cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
