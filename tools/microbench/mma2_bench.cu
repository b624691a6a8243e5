// Microbenchmark: tcgen05.mma.cta_group::2 (CTA pair, M = 256) mechanics and rate.  Each CTA of the pair holds its
// 128 rows of A and its half (N / 2 rows) of B in shared memory; the leader CTA issues, commits multicast to both.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../diffusesg_b200/csrc -o mma2_bench mma2_bench.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace dsg {
void set_last_error(const char*, ...) {}
void count_launch(int) {}
}  // namespace dsg
using namespace dsg;

DSG_DEVICE uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
DSG_DEVICE void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
DSG_DEVICE void tmem_alloc2(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
DSG_DEVICE void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}
DSG_DEVICE void umma2_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
DSG_DEVICE void umma2_commit_mc(uint64_t* bar) {  // arrives on the barrier at this offset in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3)) : "memory");
}
__host__ __device__ constexpr uint32_t idesc2_bf16(int n) {  // M = 256 across the pair
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | ((256u >> 4) << 24);
}

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) bench2(long long* out, float* probe, int reps, int per_commit) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;               // [128 x 64] bf16 SW128 = 16 KB (this CTA's rows)
  uint8_t* sB = smem + 16384;       // [N / 2 x 64] bf16 SW128 (this CTA's half of B)
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = uniform_warp_id();
  const uint32_t rank = cluster_ctarank();
  // A = 1.0 everywhere, B = 1.0 everywhere (bf16 0x3F80): D = K_total per element -> checkable
  for (int i = threadIdx.x; i < (16384 + 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3F803F80u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc2(&slot);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync();
  tcgen05_fence_after();
  const uint32_t tm = uniform_u32(slot);
  long long t0 = 0, t1 = 0;
  uint32_t ph = 0;
  if (warp == 0) {
    constexpr uint32_t idesc = idesc2_bf16(N);
    for (int pass = 0; pass < 2; ++pass) {
      t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        if (rank == 0) {
          if (elect_one()) {
            const uint64_t da = umma_desc_sw128(smem_u32(sA));
            const uint64_t db = umma_desc_sw128(smem_u32(sB));
            for (int i = 0; i < per_commit; ++i) {
              const int k = i & 3;
              umma2_ss(tm, da + 2 * k, db + 2 * k, idesc, (pass | r | i) != 0);
            }
            umma2_commit_mc(&bar);
          }
          __syncwarp();
        }
        mbar_wait(&bar, ph);  // both CTAs wait on their own copy
        ph ^= 1;
      }
      t1 = clock64();
    }
    tcgen05_fence_after();
    uint32_t v[16];
    tmem_ld_32x16(tm, v);
    tmem_ld_wait();
    if (threadIdx.x == 0) {
      out[blockIdx.x] = t1 - t0;
      probe[blockIdx.x] = __uint_as_float(v[0]);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 0) tmem_dealloc2(tm);
}

template <int N>
void run(int grid, int reps, int per_commit) {
  long long* d;
  float* pr;
  cudaMalloc(&d, grid * sizeof(long long));
  cudaMalloc(&pr, grid * sizeof(float));
  const int smem = 1024 + 16384 + 16384;
  cudaFuncSetAttribute(bench2<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  bench2<N><<<grid, 128, smem>>>(d, pr, reps, per_commit);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d: %s\n", N, cudaGetErrorString(e)); exit(1); }
  std::vector<long long> h(grid);
  std::vector<float> hp(grid);
  cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaMemcpy(hp.data(), pr, grid * sizeof(float), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (long long v : h) mx = v > mx ? v : mx;
  const double per = double(mx) / (double(reps) * per_commit);
  printf("grid %3d  2-CTA SS  M=256 N=%3d  per_commit=%3d : %7.1f clk/MMA (ideal N/2 = %d -> %.0f%%)   D[0][0] rank0 = %.0f rank1 = %.0f (expect %d)\n",
         grid, N, per_commit, per, N / 2, 100.0 * (N / 2) / per, hp[0], hp[1], 2 * reps * per_commit * 16);
  cudaFree(d);
  cudaFree(pr);
}

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 148;
  for (int pc : {4, 64}) {
    run<64>(grid, 200, pc);
    run<128>(grid, 200, pc);
    run<192>(grid, 200, pc);
    run<256>(grid, 200, pc);
  }
  return 0;
}
