// Microbenchmark: issue-to-completion cost of back-to-back tcgen05.mma (bf16, M = 128, K = 16) as a function of N,
// operand source (A from shared memory = SS, A from tensor memory = TS) and number of CTAs.  Prints clk per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../diffusesg_b200/csrc -o mma_bench mma_bench.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace dsg {
void set_last_error(const char*, ...) {}
void count_launch(int) {}
}  // namespace dsg
using namespace dsg;

DSG_DEVICE void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) bench(long long* out, int reps, int per_commit) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;               // [128 x 64] bf16 SW128 = 16 KB
  uint8_t* sB = smem + 16384;       // [256 x 64] bf16 SW128 = 32 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = uniform_warp_id();
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = uniform_u32(slot);
  if (warp == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(N);
    uint32_t ph = 0;
    long long t0 = 0, t1 = 0;
    for (int pass = 0; pass < 2; ++pass) {  // pass 0 warms up
      t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        if (elect_one()) {
          const uint64_t da = umma_desc_sw128(smem_u32(sA));
          const uint64_t db = umma_desc_sw128(smem_u32(sB));
          for (int i = 0; i < per_commit; ++i) {
            const int k = i & 3;
            if (TS) umma_ts(tm + (i & 1) * N, tm + 2 * N + k * 8, db + 2 * k, idesc, 1u);
            else umma_bf16_ss(tm + (i & 1) * N, da + 2 * k, db + 2 * k, idesc, 1u);
          }
          umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, ph);
        ph ^= 1;
      }
      t1 = clock64();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}

template <int N, bool TS>
void run(int grid, int reps, int per_commit) {
  long long* d;
  cudaMalloc(&d, grid * sizeof(long long));
  const int smem = 1024 + 16384 + 32768;
  cudaFuncSetAttribute(bench<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  bench<N, TS><<<grid, 128, smem>>>(d, reps, per_commit);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d TS=%d: %s\n", N, (int)TS, cudaGetErrorString(e)); exit(1); }
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (long long v : h) mx = v > mx ? v : mx;
  const double per = double(mx) / (double(reps) * per_commit);
  printf("grid %3d  %s  N=%3d  per_commit=%3d : %7.1f clk/MMA   (ideal N/2 = %d)  -> %.0f%% of nominal\n", grid,
         TS ? "TS" : "SS", N, per_commit, per, N / 2, 100.0 * (N / 2) / per);
  cudaFree(d);
}

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 148;
  const int reps = 2000;
  for (int pc : {4, 16, 64}) {
    run<64, false>(grid, reps, pc);
    run<96, false>(grid, reps, pc);
    run<128, false>(grid, reps, pc);
    run<192, false>(grid, reps, pc);
    run<256, false>(grid, reps, pc);
    run<64, true>(grid, reps, pc);
    run<96, true>(grid, reps, pc);
    run<128, true>(grid, reps, pc);
    run<192, true>(grid, reps, pc);
  }
  return 0;
}
