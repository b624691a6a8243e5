// Fused head of a Swin block for sm_100a (one kernel per block, C = 96):
//
//     x   = silu(shift + x (1 + scale))              (model/diffusesg/diffusesg.py:238-242 of the reference: FiLM, shortcut)
//     y   = LayerNorm1(x)                            (:243)
//     qkv = y W_qkv^T + b_qkv                        (:115, WindowAttention.qkv; q rows pre-scaled in the packed weights)
//
// Unfused this is the FiLM + LayerNorm row kernel (reads x, writes x and the bf16 y) and the qkv GEMM (reads y, writes
// qkv): 3.6 GB of HBM traffic at batch 512.  Here a 128-token tile reads x once, writes x and qkv once (2.8 GB), the
// LayerNorm output goes to the MMA through TENSOR MEMORY (bf16 pairs, TS-mode A operand) and W_qkv (55 KB) stays
// resident in shared memory for the whole kernel, so the kernel has no weight traffic and is HBM-bound.
//
//   warp 16  loader        W_qkv once; x tile in (TMA -> sX), L2 prefetch of the tile after
//   warp 17  MMA issuer    acc[128 x 288] = y . W_qkv^T as three N = 96 groups (TS mode, K = 96)
//   warp 18  storer        x tile out (sXo -> TMA store), qkv parts out (two 24 KB staging buffers -> TMA store);
//                          a staging buffer is handed back one item late (cp.async.bulk.wait_group.read 1)
//   warps 0..15 workers    F: FiLM + SiLU -> sXo, two-pass LayerNorm (values in registers) -> y in tensor memory
//                          E: acc part + bias -> bf16 -> swizzled staging            (3 parts of 96 columns per tile)
//   worker order:          F(0); then per tile t:  F(t + 1), E(t)     (the MMA of tile t + 1 runs under E(t) / F(t + 2))
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace dsg {
namespace {

constexpr int kHeadThreads = 19 * 32;
constexpr float kHeadLnEps = 1e-5f;

template <int C>
struct HeadCfg {
  static constexpr int N3 = 3 * C;                 // qkv width
  static constexpr int KBN = C / 32;               // 32-column k-blocks (64-byte swizzle)
  static constexpr int W_KB = C * 64;              // one [C x 32] k-block of one qkv part
  static constexpr int W_BYTES = 3 * KBN * W_KB;   // [part][k-block][C rows x 32]: 54 KB
  static constexpr int XB = C / 32;                // [128 x 32] fp32 boxes of an x tile
  static constexpr int X_BYTES = XB * 16384;
  static constexpr int QB = C / 32;                // [128 x 32] bf16 boxes of a qkv part
  static constexpr int Q_BYTES = QB * 8192;        // one part: 24 KB
  static constexpr int CW = C / 4;                 // columns per worker warp
  static constexpr int PAR_FLOATS = N3 + 2 * C;    // b_qkv, gamma, beta
  static constexpr int PART_BYTES = 4 * 128 * 4;   // LayerNorm partial sums: float [4 column groups][128 rows]
  static constexpr int SMEM_BYTES = 1024 + W_BYTES + 2 * X_BYTES + 2 * Q_BYTES + PART_BYTES + PAR_FLOATS * 4 + 256;
  static constexpr int Y_COL = 3 * C;              // TMEM: acc part p @ p C | y (bf16 pairs)
  static_assert(C % 32 == 0 && CW % 8 == 0, "head: C must be a multiple of 32");
  static_assert(Y_COL + C / 2 <= 512 && SMEM_BYTES <= 227 * 1024, "head budget");
};

struct HeadParams {
  const float* film;   // row b: scale[C] then shift[C]
  int film_ld;
  int cond_uniform;    // every sample uses film row 0
  int tokens_per_sample;
  const float* gamma;  // [C]  norm1
  const float* beta;   // [C]
  const float* bqkv;   // [3C]
  int M;
};

DSG_DEVICE uint64_t umma_desc_sw64_h(uint32_t smem_addr) {  // 64-byte swizzle, SBO = 512 B (see blocktail.cu)
  const uint64_t lo = ((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16);
  const uint64_t hi = (512u >> 4) | (1u << 14) | (4u << 29);
  return lo | (hi << 32);
}
DSG_DEVICE void umma_bf16_ts_h(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
DSG_DEVICE void tmem_ld_32x8_h(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
DSG_DEVICE void tmem_st_32x4_h(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
               : "memory");
}
DSG_DEVICE void tmem_st_wait_h() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
DSG_DEVICE void tma_store_2d_h(const CUtensorMap* map, const void* smem_src, int c_inner, int c_outer) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c_inner), "r"(c_outer)
               : "memory");
}

template <int C>
__global__ void __launch_bounds__(kHeadThreads, 1)
block_head_kernel(const __grid_constant__ CUtensorMap tmXin, const __grid_constant__ CUtensorMap tmXout,
                  const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmQ, const HeadParams p) {
  using G = HeadCfg<C>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem;                      // resident W_qkv
  uint8_t* sX = sW + G::W_BYTES;           // x tile in
  uint8_t* sXo = sX + G::X_BYTES;          // x tile out
  uint8_t* sQ = sXo + G::X_BYTES;          // two qkv part staging buffers
  float* sPart = reinterpret_cast<float*>(sQ + 2 * G::Q_BYTES);
  float* sBq = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sPart) + G::PART_BYTES);
  float* sGam = sBq + G::N3;
  float* sBet = sGam + C;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBet + C);
  uint64_t* w_full = bars;         // loader -> MMA: W_qkv resident
  uint64_t* xin_full = bars + 1;   // loader -> workers
  uint64_t* xin_free = bars + 2;   // workers -> loader: the x tile is in registers
  uint64_t* xo_ready = bars + 3;   // workers -> storer: the x tile out is in sXo
  uint64_t* xo_free = bars + 4;    // storer -> workers: the TMA store has read sXo
  uint64_t* y_ready = bars + 5;    // workers -> MMA: y of the tile in tensor memory
  uint64_t* acc_full = bars + 6;   // MMA -> workers: qkv accumulator complete
  uint64_t* acc_empty = bars + 7;  // workers -> MMA: accumulator drained
  uint64_t* q_ready = bars + 8;    // [2] workers -> storer
  uint64_t* q_free = bars + 10;    // [2] storer -> workers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = uniform_warp_id();
  const int lane = threadIdx.x & 31;
  constexpr int kLoadWarp = 16, kMmaWarp = 17, kStoreWarp = 18;
  const int num_tiles = (p.M + 127) / 128;
  const int my_tiles = (num_tiles > static_cast<int>(blockIdx.x)) ? (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == kLoadWarp && lane == 0) {
    tma_prefetch_desc(&tmXin);
    tma_prefetch_desc(&tmXout);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmQ);
    mbar_init(w_full, 1);
    mbar_init(xin_full, 1);
    mbar_init(xin_free, 16);
    mbar_init(xo_ready, 16);
    mbar_init(xo_free, 1);
    mbar_init(y_ready, 16);
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 16);
    for (int b = 0; b < 2; ++b) { mbar_init(&q_ready[b], 16); mbar_init(&q_free[b], 1); }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc<512>(tmem_slot);
  for (int i = threadIdx.x; i < G::N3; i += kHeadThreads) sBq[i] = p.bqkv[i];
  for (int i = threadIdx.x; i < C; i += kHeadThreads) { sGam[i] = p.gamma[i]; sBet[i] = p.beta[i]; }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);

  if (warp == kLoadWarp) {
    // ------------------------------------------------------------------ loader
    if (elect_one()) {
      mbar_expect_tx(w_full, G::W_BYTES);
      for (int part = 0; part < 3; ++part)
        for (int kb = 0; kb < G::KBN; ++kb)
          tma_load_2d(sW + (part * G::KBN + kb) * G::W_KB, &tmW, w_full, kb * 32, part * C);
      auto load_x = [&](int tl) {
        const int tile = blockIdx.x + tl * gridDim.x;
        mbar_expect_tx(xin_full, G::X_BYTES);
        for (int xb = 0; xb < G::XB; ++xb) tma_load_2d(sX + xb * 16384, &tmXin, xin_full, xb * 32, tile * 128);
      };
      auto prefetch_x = [&](int tl) {
        const int tile = blockIdx.x + tl * gridDim.x;
        for (int xb = 0; xb < G::XB; ++xb)
          asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
                       ::"l"(reinterpret_cast<uint64_t>(&tmXin)), "r"(xb * 32), "r"(tile * 128) : "memory");
      };
      if (my_tiles > 0) load_x(0);
      if (my_tiles > 1) prefetch_x(1);
      for (int t = 0; t + 1 < my_tiles; ++t) {
        mbar_wait(xin_free, t & 1);
        load_x(t + 1);
        if (t + 2 < my_tiles) prefetch_x(t + 2);
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(C);
    if (my_tiles > 0) mbar_wait(w_full, 0);
    for (int t = 0; t < my_tiles; ++t) {
      mbar_wait(y_ready, t & 1);
      mbar_wait(acc_empty, (t & 1) ^ 1);  // E(t - 1) has drained the accumulator
      tcgen05_fence_after();
      if (elect_one()) {
        for (int part = 0; part < 3; ++part)
          for (int kb = 0; kb < G::KBN; ++kb) {
            const uint64_t db = umma_desc_sw64_h(smem_u32(sW + (part * G::KBN + kb) * G::W_KB));
            for (int k = 0; k < 2; ++k)
              umma_bf16_ts_h(tmem_base + part * C, tmem_base + G::Y_COL + kb * 16 + k * 8, db + 2 * k, idesc, (kb | k) != 0);
          }
        umma_commit(acc_full);
      }
      __syncwarp();
    }
  } else if (warp == kStoreWarp) {
    // ------------------------------------------------------------------ storer
    // items in worker order: xo(0), then per tile t: xo(t + 1), q(t, 0), q(t, 1), q(t, 2).  A qkv staging buffer is
    // handed back when the NEXT item has been issued (bulk groups finish reading in order; its next user is two items
    // away); the x tile has one buffer whose next user is the very next F phase, so it is drained at once.
    if (elect_one()) {
      uint64_t* prev_free = nullptr;
      auto issued = [&](uint64_t* this_free, bool lag) {
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (lag) {
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          if (prev_free != nullptr) mbar_arrive(prev_free);
          prev_free = this_free;
        } else {
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          if (prev_free != nullptr) mbar_arrive(prev_free);
          mbar_arrive(this_free);
          prev_free = nullptr;
        }
      };
      auto store_xo = [&](int tl) {
        const int tile = blockIdx.x + tl * gridDim.x;
        mbar_wait(xo_ready, tl & 1);
        for (int xb = 0; xb < G::XB; ++xb) tma_store_2d_h(&tmXout, sXo + xb * 16384, xb * 32, tile * 128);
        issued(xo_free, false);
      };
      if (my_tiles > 0) store_xo(0);
      int k = 0;  // running qkv part counter -> staging buffer / parity
      for (int t = 0; t < my_tiles; ++t) {
        const int tile = blockIdx.x + t * gridDim.x;
        if (t + 1 < my_tiles) store_xo(t + 1);
        for (int part = 0; part < 3; ++part, ++k) {
          const int b = k & 1;
          mbar_wait(&q_ready[b], (k >> 1) & 1);
          for (int qb = 0; qb < G::QB; ++qb)
            tma_store_2d_h(&tmQ, sQ + b * G::Q_BYTES + qb * 8192, part * C + qb * 32, tile * 128);
          issued(&q_free[b], true);
        }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else {
    // ------------------------------------------------------------------ workers
    const int q = warp & 3;            // TMEM lane quarter of this warp
    const int cg = warp >> 2;          // column group 0..3
    const int r_t = q * 32 + lane;     // row owned by this thread
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int c0 = cg * G::CW;
    auto x_off = [&](int c) -> uint32_t {  // 16-byte chunk of columns [c, c + 4) of this row in an fp32 tile
      return static_cast<uint32_t>((c >> 5) * 16384 + r_t * 128 + (((((c & 31) >> 2)) ^ (r_t & 7)) << 4));
    };

    // ---- F(tl): FiLM + SiLU -> sXo;  LayerNorm -> y in tensor memory
    auto phase_f = [&](int tl) {
      const long long row = static_cast<long long>(blockIdx.x + tl * gridDim.x) * 128 + r_t;
      const long long rc = row < p.M ? row : p.M - 1;
      const float* fs = p.film + static_cast<size_t>(p.cond_uniform ? 0 : rc / p.tokens_per_sample) * p.film_ld + c0;
      mbar_wait(xin_full, tl & 1);
      float4 v[G::CW / 4];
#pragma unroll
      for (int i = 0; i < G::CW; i += 4) v[i >> 2] = *reinterpret_cast<const float4*>(sX + x_off(c0 + i));
      __syncwarp();
      if (lane == 0) mbar_arrive(xin_free);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < G::CW; i += 4) {
        const float4 sc = __ldg(reinterpret_cast<const float4*>(fs + i));
        const float4 sh = __ldg(reinterpret_cast<const float4*>(fs + C + i));
        float4& x = v[i >> 2];
        x.x = silu_f(fmaf(x.x, sc.x + 1.f, sh.x)); x.y = silu_f(fmaf(x.y, sc.y + 1.f, sh.y));
        x.z = silu_f(fmaf(x.z, sc.z + 1.f, sh.z)); x.w = silu_f(fmaf(x.w, sc.w + 1.f, sh.w));
        s += (x.x + x.y) + (x.z + x.w);
      }
      sPart[cg * 128 + r_t] = s;
      if (tl > 0) mbar_wait(xo_free, (tl - 1) & 1);  // the store of the previous x tile has read sXo
#pragma unroll
      for (int i = 0; i < G::CW; i += 4) *reinterpret_cast<float4*>(sXo + x_off(c0 + i)) = v[i >> 2];
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(xo_ready);
      asm volatile("bar.sync 1, 512;" ::: "memory");
      const float mean = ((sPart[r_t] + sPart[128 + r_t]) + (sPart[256 + r_t] + sPart[384 + r_t])) * (1.0f / C);
      float qq = 0.f;
#pragma unroll
      for (int i = 0; i < G::CW; i += 4) {
        float4& x = v[i >> 2];
        x.x -= mean; x.y -= mean; x.z -= mean; x.w -= mean;
        qq += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w);
      }
      asm volatile("bar.sync 2, 512;" ::: "memory");  // every thread has read the sums
      sPart[cg * 128 + r_t] = qq;
      asm volatile("bar.sync 1, 512;" ::: "memory");
      const float rstd =
          rsqrtf(((sPart[r_t] + sPart[128 + r_t]) + (sPart[256 + r_t] + sPart[384 + r_t])) * (1.0f / C) + kHeadLnEps);
      // y is single-buffered in tensor memory: the MMA of the previous tile must have read it
      if (tl > 0) mbar_wait(acc_full, (tl - 1) & 1);
      tcgen05_fence_after();
#pragma unroll
      for (int i = 0; i < G::CW; i += 8) {
        const float4 g0 = *reinterpret_cast<const float4*>(&sGam[c0 + i]), g1 = *reinterpret_cast<const float4*>(&sGam[c0 + i + 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&sBet[c0 + i]), b1 = *reinterpret_cast<const float4*>(&sBet[c0 + i + 4]);
        const float4 xa = v[i >> 2], xb = v[(i >> 2) + 1];
        uint32_t pk[4];
        pk[0] = pack_bf16x2(fmaf(xa.x * rstd, g0.x, b0.x), fmaf(xa.y * rstd, g0.y, b0.y));
        pk[1] = pack_bf16x2(fmaf(xa.z * rstd, g0.z, b0.z), fmaf(xa.w * rstd, g0.w, b0.w));
        pk[2] = pack_bf16x2(fmaf(xb.x * rstd, g1.x, b1.x), fmaf(xb.y * rstd, g1.y, b1.y));
        pk[3] = pack_bf16x2(fmaf(xb.z * rstd, g1.z, b1.z), fmaf(xb.w * rstd, g1.w, b1.w));
        tmem_st_32x4_h(t_lane + G::Y_COL + ((c0 + i) >> 1), pk);
      }
      tmem_st_wait_h();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(y_ready);
    };

    // ---- E(tl): accumulator parts + bias -> bf16 -> swizzled staging (64-byte rows, 32 columns per box)
    int k = 0;  // running qkv part counter (same sequence as the storer's)
    auto phase_e = [&](int tl) {
      mbar_wait(acc_full, tl & 1);
      tcgen05_fence_after();
#pragma unroll 1
      for (int part = 0; part < 3; ++part, ++k) {
        const int b = k & 1;
        uint32_t v[G::CW];
#pragma unroll
        for (int i = 0; i < G::CW; i += 8) tmem_ld_32x8_h(t_lane + part * C + c0 + i, v + i);
        tmem_ld_wait();
        if (part == 2) {  // accumulator drained by this warp
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty);
        }
        if (k >= 2) mbar_wait(&q_free[b], ((k >> 1) - 1) & 1);  // the store that last used this buffer has read it
        uint8_t* buf = sQ + b * G::Q_BYTES;
#pragma unroll
        for (int i = 0; i < G::CW; i += 8) {
          const int c = c0 + i;  // column inside the part
          const float4 ba = *reinterpret_cast<const float4*>(&sBq[part * C + c]);
          const float4 bb = *reinterpret_cast<const float4*>(&sBq[part * C + c + 4]);
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(v[i]) + ba.x, __uint_as_float(v[i + 1]) + ba.y);
          w.y = pack_bf16x2(__uint_as_float(v[i + 2]) + ba.z, __uint_as_float(v[i + 3]) + ba.w);
          w.z = pack_bf16x2(__uint_as_float(v[i + 4]) + bb.x, __uint_as_float(v[i + 5]) + bb.y);
          w.w = pack_bf16x2(__uint_as_float(v[i + 6]) + bb.z, __uint_as_float(v[i + 7]) + bb.w);
          // box (c / 32): 64-byte rows, 16-byte chunk index ^= (row >> 1) & 3
          *reinterpret_cast<uint4*>(buf + (c >> 5) * 8192 + r_t * 64 + (((((c & 31) >> 3)) ^ ((r_t >> 1) & 3)) << 4)) = w;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&q_ready[b]);
      }
    };

    if (my_tiles > 0) phase_f(0);
    for (int tl = 0; tl < my_tiles; ++tl) {
      if (tl + 1 < my_tiles) phase_f(tl + 1);
      phase_e(tl);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == kMmaWarp) tmem_dealloc<512>(tmem_base);
}

}  // namespace

bool block_head_supported(int C) { return C == 96; }

int launch_block_head(const CUtensorMap* tmXin, const CUtensorMap* tmXout, const CUtensorMap* tmW, const CUtensorMap* tmQ,
                      const float* film, int film_ld, int cond_uniform, int tokens_per_sample, const float* gamma,
                      const float* beta, const float* bqkv, long long rows, int C, cudaStream_t st) {
  DSG_REQUIRE(block_head_supported(C) && rows > 0 && rows < 2147483647LL, "block_head: C=%d rows=%lld", C, rows);
  using G = HeadCfg<96>;
  static PerDeviceOnce configured;
  if (configured.first()) {
    DSG_CUDA_CHECK(cudaFuncSetAttribute(block_head_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
  }
  const int sms = device_sm_count();
  HeadParams p{film, film_ld, cond_uniform, tokens_per_sample, gamma, beta, bqkv, static_cast<int>(rows)};
  const int tiles = (p.M + 127) / 128;
  block_head_kernel<96><<<tiles < sms ? tiles : sms, kHeadThreads, G::SMEM_BYTES, st>>>(*tmXin, *tmXout, *tmW, *tmQ, p);
  DSG_LAUNCH_CHECK();
  return DSG_OK;
}

}  // namespace dsg
