// Shared device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM wrappers (inline PTX),
// small math helpers.  Everything here is header-only and internal to libdsg_b200.so.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dsg_b200.h"  // DSG_OK / DSG_ERR_* codes

#define DSG_DEVICE __device__ __forceinline__

namespace dsg {

// ------------------------------------------------------------------------------------------------
// error plumbing (host)
// ------------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
void count_launch(int n = 1);

#define DSG_CUDA_CHECK(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::dsg::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                            __FILE__, __LINE__);                                          \
      return ::DSG_ERR_CUDA;                                                         \
    }                                                                                     \
  } while (0)

#define DSG_LAUNCH_CHECK()                                                                \
  do {                                                                                    \
    ::dsg::count_launch();                                                                \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) {                                                              \
      ::dsg::set_last_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),  \
                            __FILE__, __LINE__);                                          \
      return ::DSG_ERR_CUDA;                                                         \
    }                                                                                     \
  } while (0)

#define DSG_REQUIRE(cond, ...)                                                            \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      ::dsg::set_last_error(__VA_ARGS__);                                                 \
      return ::DSG_ERR_INVALID;                                                      \
    }                                                                                     \
  } while (0)

// ------------------------------------------------------------------------------------------------
// generic device helpers
// ------------------------------------------------------------------------------------------------
DSG_DEVICE uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// Warp-uniform values the compiler can PROVE uniform (a shuffle broadcast): branching on them keeps the MMA / TMA
// issue code on the uniform datapath.  With `threadIdx.x >> 5` and `if (lane == 0)` ptxas wraps every tcgen05.mma
// in an ELECT + 3x R2UR.BROADCAST + BRA.U.ANY loop (~95 clk per MMA, measured), which starves the tensor pipe.
DSG_DEVICE int uniform_warp_id() { return __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0); }
DSG_DEVICE uint32_t uniform_u32(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
DSG_DEVICE bool elect_one() {  // true in exactly one lane of a converged warp
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

DSG_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// MUFU-backed approximations (1-2 ulp): the IEEE-rounded expf / division sequences cost ~10x more issue slots
DSG_DEVICE float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
DSG_DEVICE float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

DSG_DEVICE float silu_f(float x) { return x * rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }

// ------------------------------------------------------------------------------------------------
// packed fp32 pairs (FFMA2 / FMUL2 / FADD2 on sm_100): two lanes of work per issue slot.  The epilogue warps of the
// tcgen05 kernels are issue-bound, not latency-bound, so halving the instruction count is a direct speed-up.
// ------------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;
DSG_DEVICE f32x2 f2_pack(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
DSG_DEVICE f32x2 f2_splat(float v) { return f2_pack(v, v); }
DSG_DEVICE void f2_unpack(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
DSG_DEVICE f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
DSG_DEVICE f32x2 f2_mul(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
DSG_DEVICE f32x2 f2_add(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
DSG_DEVICE float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// erf GELU, the reference's nn.GELU() default (model/diffusesg/diffusesg.py:10,15), evaluated as
//   gelu(x) = x/2 (1 + tanh(x (a + b x^2 + c x^4)))
// with (a, b, c) a minimax fit to the ERF form (not the "tanh GELU" constants): |fit - gelu_erf| <= 2.6e-5
// everywhere, 2.6e-5 relative rms under N(0,1) inputs.  MUFU.TANH adds <= 2^-11 relative error on tanh.  Both are far
// below the bf16 rounding every call site applies to the result (1.7e-3 relative rms).  Two elements cost 6 packed
// FP32 instructions + 2 MUFU; erff() costs ~40 per element.
DSG_DEVICE f32x2 gelu_erf2(f32x2 x) {
  const f32x2 x2 = f2_mul(x, x);
  f32x2 q = f2_fma(x2, f2_splat(-3.51516789e-04f), f2_splat(3.70056460e-02f));
  q = f2_fma(x2, q, f2_splat(7.97507884e-01f));
  float u0, u1;
  f2_unpack(f2_mul(x, q), u0, u1);
  const f32x2 t = f2_pack(tanh_approx(u0), tanh_approx(u1));
  const f32x2 h = f2_mul(x, f2_splat(0.5f));
  return f2_fma(h, t, h);
}
DSG_DEVICE float gelu_erf(float x) {
  const float x2 = x * x;
  const float q = fmaf(x2, fmaf(x2, -3.51516789e-04f, 3.70056460e-02f), 7.97507884e-01f);
  const float h = 0.5f * x;
  return fmaf(h, tanh_approx(x * q), h);
}

DSG_DEVICE uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
DSG_DEVICE uint32_t pack_bf16x2(f32x2 v) {
  float lo, hi;
  f2_unpack(v, lo, hi);
  return pack_bf16x2(lo, hi);
}
// bf16x2(gelu(acc + bias)) of two adjacent accumulator columns
DSG_DEVICE uint32_t gelu_bias_bf16x2(uint32_t acc_lo, uint32_t acc_hi, float b_lo, float b_hi) {
  return pack_bf16x2(gelu_erf2(f2_add(f2_pack(__uint_as_float(acc_lo), __uint_as_float(acc_hi)), f2_pack(b_lo, b_hi))));
}

// ------------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------------
DSG_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

DSG_DEVICE void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

DSG_DEVICE void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

DSG_DEVICE void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}

DSG_DEVICE void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// One potentially-blocking probe (the hardware may suspend the thread for a short, implementation-defined time).
// NB: the variant with an explicit suspend-time hint was measured to sleep for the WHOLE hint (~2 us) instead of
// waking on phase completion, which put 1-2 k cycles of wake-up latency on every producer/consumer hand-off.
DSG_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// Bounded wait: a pipeline bug must surface as a trap (launch error), never as a hung GPU.  The bound counts
// probes, not clock reads, to keep the wait loop at a few instructions.
DSG_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t probes = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++probes > 400000000u) {  // seconds, even if every probe returned immediately
      printf("dsg: mbarrier wait timed out (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// TMA (bulk tensor copy global -> shared, completion on an mbarrier)
// ------------------------------------------------------------------------------------------------
DSG_DEVICE void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

DSG_DEVICE void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c_inner),
      "r"(c_outer)
      : "memory");
}

// ------------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA issue, commit, TMEM loads
// ------------------------------------------------------------------------------------------------
template <uint32_t kCols>
DSG_DEVICE void tmem_alloc(uint32_t* smem_result) {  // whole warp
  static_assert(kCols >= 32 && kCols <= 512 && (kCols & (kCols - 1)) == 0, "TMEM columns: power of two in [32,512]");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

template <uint32_t kCols>
DSG_DEVICE void tmem_dealloc(uint32_t taddr) {  // whole warp (the allocating one)
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

DSG_DEVICE void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
DSG_DEVICE void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor for a K-major bf16 tile stored as rows of 128 bytes with the
// 128-byte swizzle (what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B):
//   start address >> 4 | LBO (unused for swizzled K-major, =1) | SBO = 1024 B (8 rows) | version 1 | SWIZZLE_128B
DSG_DEVICE uint64_t umma_desc_sw128(uint32_t smem_addr) {
  const uint64_t lo = ((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16);
  const uint64_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return lo | (hi << 32);
}

// MN-major operand tile (the non-contracted index is the contiguous one: a weight-gradient GEMM reads dY [tokens, N_out]
// and X [tokens, K_in] as they lie in memory).  Canonical layout, 128-byte swizzle, in 16-byte units
// ((8, n), (8, k)) : ((1, LBO), (8, SBO)): an atom is 64 MN elements x 8 K rows = what TMA writes for a box of 64 columns
// (128 B) x 8 rows; K groups of 8 rows are SBO = 1024 B apart, the next 64 MN elements (the next TMA box) LBO bytes.
DSG_DEVICE uint64_t umma_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  const uint64_t lo = ((smem_addr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(lbo_bytes >> 4) << 16);
  const uint64_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return lo | (hi << 32);
}
// Instruction descriptor: bf16 x bf16 -> fp32, BOTH operands MN-major (bits 15, 16), M=128, N=n.
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>(n >> 3) << 17) | ((128u >> 4) << 24);
}

// Instruction descriptor: bf16 x bf16 -> fp32, both operands K-major, M=128, N=n.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | ((128u >> 4) << 24);
}

DSG_DEVICE void umma_bf16_ss(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// mbarrier arrives once every tcgen05.mma previously issued by this thread has completed.
DSG_DEVICE void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------------------------------------
// CTA pairs (cta_group::2): two CTAs of a 2-cluster issue ONE M = 256 MMA; each holds its 128 rows of A, half of B
// (N / 2 rows) and the accumulator rows of its own 128 rows.  Only the leader (cluster rank 0) issues; commits are
// multicast to the barrier at the same shared-memory offset in both CTAs.  (tools/microbench/mma2_bench.cu)
// ------------------------------------------------------------------------------------------------
DSG_DEVICE uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
DSG_DEVICE void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared-memory object of this CTA) in the CTA of rank `rank`
DSG_DEVICE uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
DSG_DEVICE void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
template <uint32_t kCols>
DSG_DEVICE void tmem_alloc_pair(uint32_t* smem_result) {  // the same warp of BOTH CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
DSG_DEVICE void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_pair(int n) {  // M = 256 across the pair
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | ((256u >> 4) << 24);
}
DSG_DEVICE void umma_bf16_ss_pair(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
DSG_DEVICE void umma_commit_pair(uint64_t* bar) {  // arrives in BOTH CTAs once the issued MMAs have completed
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
               : "memory");
}
// TMA load into THIS CTA's shared memory whose completion is signalled on a barrier of the pair's leader
DSG_DEVICE void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint32_t leader_bar_cluster_addr, int c_inner,
                                 int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar_cluster_addr), "r"(c_inner),
      "r"(c_outer)
      : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives row (lane base + t).
DSG_DEVICE void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// 32 lanes x 16 consecutive fp32 columns
DSG_DEVICE void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

DSG_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// legacy warp MMA (m16n8k16 bf16 -> fp32), used by the small-tile window attention kernel
// ------------------------------------------------------------------------------------------------
DSG_DEVICE void mma_m16n8k16_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}


// ---- per-device one-time setup -----------------------------------------------------------------------------------
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count belong to a DEVICE, not to the process: a second
// GPU driven from the same process (nn.DataParallel, a notebook) needs its own opt-in.  `once.first()` is true the
// first time a call site runs with a given device current.
inline int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
  return dev;
}
struct PerDeviceOnce {
  unsigned long long seen[4] = {0, 0, 0, 0};  // 256 device ordinals
  bool first() {
    const int d = current_device() & 255;
    const unsigned long long bit = 1ull << (d & 63);
    const unsigned long long old = __atomic_fetch_or(&seen[d >> 6], bit, __ATOMIC_RELAXED);
    return (old & bit) == 0;
  }
};
inline int device_sm_count() {
  static int cache[256];
  const int d = current_device() & 255;
  int n = __atomic_load_n(&cache[d], __ATOMIC_RELAXED);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 148; }
    __atomic_store_n(&cache[d], n, __ATOMIC_RELAXED);
  }
  return n;
}

}  // namespace dsg
