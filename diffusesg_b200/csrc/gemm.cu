// tcgen05 / TMEM / TMA GEMM for sm_100a:   out[M, N] = epilogue(A[M, K] . W[N, K]^T)
//
// A (activations) and W (nn.Linear weight, [out, in]) are row-major bf16 with K contiguous, i.e. both
// operands are "K-major" in UMMA terms.  One persistent CTA per SM walks 128 x BN output tiles:
//
//   warp 0      TMA producer: 128 x 64 A box + BN x 64 W box per stage (128-byte swizzle), kStages ring
//   warp 1      MMA issuer: one lane issues tcgen05.mma (M=128, N=BN, K=16) x 4 per stage; accumulators
//               live in TMEM, double buffered so the epilogue of tile i overlaps the main loop of i+1
//   warps 2..5  epilogue: tcgen05.ld (32 lanes x 32 columns per warp), bias / GELU / residual /
//               fused adj read-out head, vectorised global stores
//
// Replaces the cuBLAS sgemm behind every nn.Linear / 1x1 conv of the reference denoiser
// (model/diffusesg/diffusesg.py:20-24 Mlp, :115 qkv, :137 proj, :334 reduction, :385/:402 breakup linears,
// :705-709 read_out, :806-809 adj head).
#include "common.cuh"
#include "kernels.h"

namespace dsg {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;                    // 64 bf16 = one 128-byte swizzle row
constexpr int kGemmThreads = 192;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB

template <int BN>
struct Cfg {
  static constexpr int kStages = (BN == 192) ? 4 : 6;
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int ACC_STRIDE = (BN == 192) ? 256 : 128;  // TMEM columns between the two accumulators
  static constexpr uint32_t TMEM_COLS = 2 * ACC_STRIDE;       // 512 / 256
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + kStages * STAGE_BYTES + 256 /*barriers*/;
};

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmParams p) {
  using C = Cfg<BN>;
  constexpr int kStages = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * A_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  __shared__ float s_w2t[(EPI == EPI_ADJ_HEAD) ? 96 * 8 : 1];
  __shared__ float s_b2[8];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (p.M + BM - 1) / BM;
  const int num_n = p.N / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  if (EPI == EPI_ADJ_HEAD) {
    for (int i = threadIdx.x; i < 96 * 8; i += kGemmThreads) s_w2t[i] = p.w2t[i];
    if (threadIdx.x < 8) s_b2[threadIdx.x] = p.b2[threadIdx.x];
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], C::STAGE_BYTES);
          tma_load_2d(sA + s * A_STAGE_BYTES, &tmA, &full_bar[s], kb * BK, m_blk * BM);
          tma_load_2d(sB + s * C::B_STAGE_BYTES, &tmW, &full_bar[s], kb * BK, n_blk * BN);
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(BN);
    int s = 0;
    uint32_t ph = 0;
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_ph ^ 1);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + acc * C::ACC_STRIDE;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[s], ph);
        tcgen05_fence_after();
        if (lane == 0) {
          const uint64_t da = umma_desc_sw128(smem_u32(sA + s * A_STAGE_BYTES));
          const uint64_t db = umma_desc_sw128(smem_u32(sB + s * C::B_STAGE_BYTES));
          const int ksteps = min(BK, p.K - kb * BK) / 16;
          for (int k = 0; k < ksteps; ++k) {
            // advancing K by 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in the (addr >> 4) field
            umma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[s]);
          if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
      acc ^= 1;
      if (acc == 0) acc_ph ^= 1;
    }
  } else {
    // ------------------------------------------------------------------ epilogue (4 warps)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      mbar_wait(&tfull_bar[acc], acc_ph);
      tcgen05_fence_after();
      const int row = m_blk * BM + q * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * C::ACC_STRIDE;
      float y[8];
      if (EPI == EPI_ADJ_HEAD) {
#pragma unroll
        for (int c = 0; c < 8; ++c) y[c] = s_b2[c];
      }
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(t_row + c0, r);
        tmem_ld_wait();
        const int col = n_blk * BN + c0;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col + j));
            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
          }
        }
        if (EPI == EPI_GELU_BF16 || EPI == EPI_ADJ_HEAD) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
        }
        if (EPI == EPI_BF16 || EPI == EPI_GELU_BF16) {
          if (row_ok) {
            bf16* o = reinterpret_cast<bf16*>(p.out) + static_cast<size_t>(row) * p.ldo + col;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 w;
              w.x = pack_bf16x2(v[j], v[j + 1]);
              w.y = pack_bf16x2(v[j + 2], v[j + 3]);
              w.z = pack_bf16x2(v[j + 4], v[j + 5]);
              w.w = pack_bf16x2(v[j + 6], v[j + 7]);
              *reinterpret_cast<uint4*>(o + j) = w;
            }
          }
        } else if (EPI == EPI_RES_F32 || EPI == EPI_F32) {
          if (row_ok) {
            float* o = reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + col;
            if (EPI == EPI_RES_F32) {
              const float* rs = p.res + static_cast<size_t>(row) * p.ldo + col;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 t = *reinterpret_cast<const float4*>(rs + j);
                v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
              }
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
        } else {  // EPI_ADJ_HEAD: second linear of the read-out MLP, 96 -> c_e (padded to 8)
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float4 wa = *reinterpret_cast<const float4*>(&s_w2t[(c0 + j) * 8]);
            const float4 wb = *reinterpret_cast<const float4*>(&s_w2t[(c0 + j) * 8 + 4]);
            y[0] = fmaf(wa.x, v[j], y[0]); y[1] = fmaf(wa.y, v[j], y[1]);
            y[2] = fmaf(wa.z, v[j], y[2]); y[3] = fmaf(wa.w, v[j], y[3]);
            y[4] = fmaf(wb.x, v[j], y[4]); y[5] = fmaf(wb.y, v[j], y[5]);
            y[6] = fmaf(wb.z, v[j], y[6]); y[7] = fmaf(wb.w, v[j], y[7]);
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (EPI == EPI_ADJ_HEAD && row_ok) {
        // row = pixel (b, i, j); zero rows/cols of padded nodes (utils/graph_utils.py:5-38), optional EDM
        // output preconditioning D = c_skip x + c_out F (model/precond/precond.py:102-104)
        const int n = p.n_img;
        const int nn = n * n;
        const int b = row / nn;
        const int ij = row - b * nn;
        const int i = ij / n, j = ij - i * n;
        const bool ok = p.flags[b * n + i] != 0 && p.flags[b * n + j] != 0;
        float cs = 0.f, co = 1.f;
        if (p.x_adj != nullptr) { cs = p.c_skip[b]; co = p.c_out[b]; }
        for (int c = 0; c < p.c_e; ++c) {
          const size_t o = (static_cast<size_t>(b) * p.c_e + c) * nn + ij;
          float val = y[c];
          if (p.x_adj != nullptr) val = __fadd_rn(__fmul_rn(cs, p.x_adj[o]), __fmul_rn(co, val));
          reinterpret_cast<float*>(p.out)[o] = ok ? val : 0.f;
        }
      }
      acc ^= 1;
      if (acc == 0) acc_ph ^= 1;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 1) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
  }
  return fn;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BN, int EPI>
int launch_t(const CUtensorMap* tmA, const CUtensorMap* tmW, const GemmParams& p, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    DSG_CUDA_CHECK(cudaFuncSetAttribute(gemm_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg<BN>::SMEM_BYTES));
    configured = true;
  }
  const int tiles = ((p.M + BM - 1) / BM) * (p.N / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_kernel<BN, EPI><<<grid, kGemmThreads, Cfg<BN>::SMEM_BYTES, st>>>(*tmA, *tmW, p);
  DSG_LAUNCH_CHECK();
  return DSG_OK;
}

}  // namespace

int make_tmap_bf16(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_last_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return DSG_ERR_CUDA;
  }
  DSG_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (cols * 2) % 16 == 0 && box_rows > 0 && box_rows <= 256,
              "make_tmap_bf16: unaligned base/pitch or bad box (rows=%lld cols=%lld box=%d)", (long long)rows,
              (long long)cols, box_rows);
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(cols) * 2};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld box_rows=%d)", (int)r,
                   (long long)rows, (long long)cols, box_rows);
    return DSG_ERR_CUDA;
  }
  return DSG_OK;
}

int gemm_block_n(int N) { return (N % 192 == 0) ? 192 : 96; }

int launch_gemm(const CUtensorMap* tmA, const CUtensorMap* tmW, int epi, const GemmParams& p, cudaStream_t st) {
  DSG_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0 && p.N % 96 == 0 && p.K % 16 == 0,
              "gemm: unsupported shape M=%d N=%d K=%d (N %% 96 == 0 and K %% 16 == 0 required)", p.M, p.N, p.K);
  DSG_REQUIRE(p.out != nullptr, "gemm: null output");
  const int bn = gemm_block_n(p.N);
  if (epi == EPI_ADJ_HEAD) {
    DSG_REQUIRE(p.N == 96 && p.c_e >= 1 && p.c_e <= 8 && p.w2t && p.b2 && p.flags && p.n_img > 0,
                "gemm: adj-head epilogue needs N == 96, c_e <= 8 and the head tensors");
    return launch_t<96, EPI_ADJ_HEAD>(tmA, tmW, p, st);
  }
  DSG_REQUIRE(p.ldo % 8 == 0, "gemm: ldo must be a multiple of 8");
  if (epi == EPI_RES_F32) DSG_REQUIRE(p.res != nullptr, "gemm: residual epilogue without residual");
  if (bn == 192) {
    switch (epi) {
      case EPI_BF16: return launch_t<192, EPI_BF16>(tmA, tmW, p, st);
      case EPI_GELU_BF16: return launch_t<192, EPI_GELU_BF16>(tmA, tmW, p, st);
      case EPI_RES_F32: return launch_t<192, EPI_RES_F32>(tmA, tmW, p, st);
      case EPI_F32: return launch_t<192, EPI_F32>(tmA, tmW, p, st);
    }
  } else {
    switch (epi) {
      case EPI_BF16: return launch_t<96, EPI_BF16>(tmA, tmW, p, st);
      case EPI_GELU_BF16: return launch_t<96, EPI_GELU_BF16>(tmA, tmW, p, st);
      case EPI_RES_F32: return launch_t<96, EPI_RES_F32>(tmA, tmW, p, st);
      case EPI_F32: return launch_t<96, EPI_F32>(tmA, tmW, p, st);
    }
  }
  set_last_error("gemm: unknown epilogue %d", epi);
  return DSG_ERR_INVALID;
}

}  // namespace dsg
