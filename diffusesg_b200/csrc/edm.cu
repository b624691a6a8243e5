// Fused elementwise kernels of the EDM stochastic-Heun sampler (runner/mcmc_sampler/edm.py:350-434 of the
// reference; ~113 ATen launches per step there, two here).
//
// State layout: adj [B, C_e, N, N] and node [B, N, C_n], fp32, exactly the reference's tensors.  Pure
// HBM-bound streaming: 128-bit loads/stores over the adjacency tensor (N % 4 == 0), grid-stride with a grid
// that is a multiple of the SM count.  Arithmetic is written with explicit round-to-nearest intrinsics in
// the reference's operation order (no FMA contraction) so that, given the same inputs, the results are
// bit-identical to the fp32 torch expressions:
//
//   pre :  x_hat  = mask(x + c * eps)                      c = sqrt(t_hat^2 - t_cur^2) * S_noise     (:361-366)
//   post:  k      = mask(inv_t * x_hat - inv_t * D1)                                                  (:384-387)
//          x'     = x_hat + h * k                                                                    (:389-390)
//          x_next = mask(x_hat + h * (0.5 k + 0.5 (inv_tp * x' - inv_tp * D2)))     (Heun, :414-422)
//          x_next = mask(x')                                                        (last step, :395-396)
#include <curand_kernel.h>

#include "common.cuh"
#include "kernels.h"

namespace dsg {
namespace {

struct EdmShape {
  int batch, c_e, n, c_n;
};

DSG_DEVICE float pre_one(float x, float e, float c) { return __fadd_rn(x, __fmul_rn(c, e)); }

DSG_DEVICE float post_one(float xh, float d1, float d2, float inv_t, float h, float inv_tp, bool heun) {
  const float k = __fsub_rn(__fmul_rn(inv_t, xh), __fmul_rn(inv_t, d1));
  const float xp = __fadd_rn(xh, __fmul_rn(h, k));
  if (!heun) return xp;
  const float kp = __fsub_rn(__fmul_rn(inv_tp, xp), __fmul_rn(inv_tp, d2));
  return __fadd_rn(xh, __fmul_rn(h, __fadd_rn(__fmul_rn(0.5f, k), __fmul_rn(0.5f, kp))));
}

// MODE 0: pre-step, 1: post-step (heun), 2: post-step (euler / last), 3: x = mask(x * scale)
template <int MODE>
__global__ void __launch_bounds__(256)
edm_kernel(const float* __restrict__ a0, const float* __restrict__ a1, const float* __restrict__ a2,
           float* __restrict__ a_out, const float* __restrict__ n0, const float* __restrict__ n1,
           const float* __restrict__ n2, float* __restrict__ n_out, const uint8_t* __restrict__ flags, float s0,
           float s1, float s2, EdmShape sh) {
  const int n = sh.n, n4 = sh.n >> 2;
  const long long adj_vec = static_cast<long long>(sh.batch) * sh.c_e * n * n4;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (long long v = tid; v < adj_vec; v += stride) {
    const int j4 = static_cast<int>(v % n4);
    const long long r = v / n4;
    const int i = static_cast<int>(r % n);
    const int b = static_cast<int>(r / (static_cast<long long>(n) * sh.c_e));
    const uint8_t* f = flags + static_cast<size_t>(b) * n;
    const bool fi = f[i] != 0;
    const uchar4 fj = *reinterpret_cast<const uchar4*>(f + 4 * j4);
    float4 x = reinterpret_cast<const float4*>(a0)[v];
    float4 o;
    if (MODE == 0) {
      const float4 e = reinterpret_cast<const float4*>(a1)[v];
      o = make_float4(pre_one(x.x, e.x, s0), pre_one(x.y, e.y, s0), pre_one(x.z, e.z, s0), pre_one(x.w, e.w, s0));
    } else if (MODE == 1 || MODE == 2) {
      const float4 d1 = reinterpret_cast<const float4*>(a1)[v];
      float4 d2 = d1;
      if (MODE == 1) d2 = reinterpret_cast<const float4*>(a2)[v];
      o = make_float4(post_one(x.x, d1.x, d2.x, s0, s1, s2, MODE == 1), post_one(x.y, d1.y, d2.y, s0, s1, s2, MODE == 1),
                      post_one(x.z, d1.z, d2.z, s0, s1, s2, MODE == 1), post_one(x.w, d1.w, d2.w, s0, s1, s2, MODE == 1));
    } else {
      o = make_float4(__fmul_rn(x.x, s0), __fmul_rn(x.y, s0), __fmul_rn(x.z, s0), __fmul_rn(x.w, s0));
    }
    o.x = (fi && fj.x) ? o.x : 0.f;
    o.y = (fi && fj.y) ? o.y : 0.f;
    o.z = (fi && fj.z) ? o.z : 0.f;
    o.w = (fi && fj.w) ? o.w : 0.f;
    reinterpret_cast<float4*>(a_out)[v] = o;
  }
  const long long node_el = static_cast<long long>(sh.batch) * n * sh.c_n;
  for (long long v = tid; v < node_el; v += stride) {
    const long long bi = v / sh.c_n;
    const bool ok = flags[bi] != 0;
    const float x = n0[v];
    float o;
    if (MODE == 0) {
      o = pre_one(x, n1[v], s0);
    } else if (MODE == 1) {
      o = post_one(x, n1[v], n2[v], s0, s1, s2, true);
    } else if (MODE == 2) {
      o = post_one(x, n1[v], 0.f, s0, s1, s2, false);
    } else {
      o = __fmul_rn(x, s0);
    }
    n_out[v] = ok ? o : 0.f;
  }
}

// Decode of the final sample (runner/sampler/sampler_node_adj.py:199-285 of the reference, 'bits' encodings):
//   class = clamp(bin2dec(x > 0, MSB first), 0, num_types - 1), masked; self-loops removed; box = x * 0.5 + 0.5, masked.
// (clamp(x, -1, 1) > 0 == x > 0, so the reference's clamp -> sign -> gt(0) chain collapses to one comparison.)
__global__ void __launch_bounds__(256)
decode_kernel(const float* __restrict__ adj, const float* __restrict__ node, const uint8_t* __restrict__ flags,
              int32_t* __restrict__ adj_cls, int32_t* __restrict__ node_cls, float* __restrict__ bbox, int num_adj_type,
              int num_node_type, EdmShape sh) {
  const int n = sh.n, nn = sh.n * sh.n;
  const long long pixels = static_cast<long long>(sh.batch) * nn;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (long long px = tid; px < pixels; px += stride) {
    const int b = static_cast<int>(px / nn);
    const int ij = static_cast<int>(px - static_cast<long long>(b) * nn);
    const int i = ij / n, j = ij - i * n;
    int v = 0;
    if (i != j && flags[b * n + i] != 0 && flags[b * n + j] != 0) {
      for (int c = 0; c < sh.c_e; ++c) v = (v << 1) | (adj[(static_cast<size_t>(b) * sh.c_e + c) * nn + ij] > 0.0f ? 1 : 0);
      v = v < 0 ? 0 : (v > num_adj_type - 1 ? num_adj_type - 1 : v);
    }
    adj_cls[px] = v;
  }
  const int nb = sh.c_n - 4;  // class bits; the last four node channels are the box
  const long long nodes = static_cast<long long>(sh.batch) * n;
  for (long long bi = tid; bi < nodes; bi += stride) {
    const bool ok = flags[bi] != 0;
    const float* x = node + bi * sh.c_n;
    int v = 0;
    if (ok) {
      for (int c = 0; c < nb; ++c) v = (v << 1) | (x[c] > 0.0f ? 1 : 0);
      v = v < 0 ? 0 : (v > num_node_type - 1 ? num_node_type - 1 : v);
    }
    node_cls[bi] = v;
#pragma unroll
    for (int c = 0; c < 4; ++c) bbox[bi * 4 + c] = ok ? __fadd_rn(__fmul_rn(x[nb + c], 0.5f), 0.5f) : 0.f;
  }
}

// Pre-step with the noise drawn IN the kernel, bit-compatible with the two `torch.randn_like` calls of the reference
// (runner/mcmc_sampler/edm.py:358-364): ATen's normal_ on CUDA is a grid-stride kernel over Philox4x32-10,
//   curand_init(seed, thread, offset); per iteration  float4 r = curand_normal4()  ->  elements  li + k T, k < 4,
//   T = 256 * grid threads, grid = min(SMs * (max threads per SM / 256), ceil(numel / 256))
// (ATen/native/cuda/DistributionTemplates.h: calc_execution_policy, distribution_elementwise_grid_stride_kernel).
// Launching the same virtual grid with the generator's (seed, offset) reproduces eps element for element; the caller
// advances the generator by the same counter offset.  One launch per tensor (each has its own offset and grid).
template <bool ADJ>
__global__ void __launch_bounds__(256)
edm_pre_philox_kernel(const float* __restrict__ x, float* __restrict__ out, const uint8_t* __restrict__ flags, float c,
                      unsigned long long seed, unsigned long long offset, long long numel, EdmShape sh) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  curandStatePhilox4_32_10_t state;
  curand_init(seed, idx, offset, &state);
  const long long T = static_cast<long long>(blockDim.x) * gridDim.x;
  const long long rounded = ((numel - 1) / (T * 4) + 1) * T * 4;
  const int n = sh.n;
  for (long long li = idx; li < rounded; li += T * 4) {
    const float4 r = curand_normal4(&state);
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const long long e = li + T * ii;
      if (e < numel) {
        const float eps = ii == 0 ? r.x : (ii == 1 ? r.y : (ii == 2 ? r.z : r.w));
        bool ok;
        const unsigned eu = static_cast<unsigned>(e);  // numel < 2^31 (checked by the launcher): 32-bit divisions
        if (ADJ) {
          const unsigned un = static_cast<unsigned>(n);
          const unsigned rr = eu / un, j = eu - rr * un;
          const unsigned q2 = rr / un, i = rr - q2 * un;       // q2 = b * c_e + c
          const unsigned b = q2 / static_cast<unsigned>(sh.c_e);
          ok = flags[b * un + i] != 0 && flags[b * un + j] != 0;
        } else {
          ok = flags[eu / static_cast<unsigned>(sh.c_n)] != 0;
        }
        out[e] = ok ? pre_one(x[e], eps, c) : 0.f;
      }
    }
  }
}

template <int MODE>
int launch_mode(const float* a0, const float* a1, const float* a2, float* a_out, const float* n0, const float* n1,
                const float* n2, float* n_out, const uint8_t* flags, float s0, float s1, float s2, int batch, int c_e,
                int n, int c_n, cudaStream_t st) {
  DSG_REQUIRE(batch > 0 && c_e > 0 && n > 0 && c_n > 0 && n % 4 == 0, "edm step: bad shape B=%d C_e=%d N=%d C_n=%d",
              batch, c_e, n, c_n);
  DSG_REQUIRE(((reinterpret_cast<uintptr_t>(a0) | reinterpret_cast<uintptr_t>(a_out) | reinterpret_cast<uintptr_t>(a1) |
                reinterpret_cast<uintptr_t>(a2)) & 15) == 0 && (reinterpret_cast<uintptr_t>(flags) & 3) == 0,
              "edm step: adjacency tensors must be 16-byte aligned (flags 4-byte)");
  const long long vec = static_cast<long long>(batch) * c_e * n * (n / 4);
  long long blocks = (vec + 255) / 256;
  const long long cap = 148LL * 8;  // 8 resident CTAs of 256 threads per SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  EdmShape sh{batch, c_e, n, c_n};
  edm_kernel<MODE><<<static_cast<unsigned>(blocks), 256, 0, st>>>(a0, a1, a2, a_out, n0, n1, n2, n_out, flags, s0, s1,
                                                                  s2, sh);
  DSG_LAUNCH_CHECK();
  return DSG_OK;
}

}  // namespace

int launch_edm_pre_step(const float* adj, const float* node, const float* eps_adj, const float* eps_node,
                        const uint8_t* flags, float noise_coef, float* adj_hat, float* node_hat, int batch, int c_e,
                        int n, int c_n, cudaStream_t st) {
  return launch_mode<0>(adj, eps_adj, nullptr, adj_hat, node, eps_node, nullptr, node_hat, flags, noise_coef, 0.f, 0.f,
                        batch, c_e, n, c_n, st);
}

int launch_edm_pre_step_philox(const float* adj, const float* node, const uint8_t* flags, float noise_coef,
                               unsigned long long seed, unsigned long long offset_adj, int grid_adj,
                               unsigned long long offset_node, int grid_node, float* adj_hat, float* node_hat, int batch,
                               int c_e, int n, int c_n, cudaStream_t st) {
  DSG_REQUIRE(batch > 0 && c_e > 0 && n > 0 && c_n > 0 && grid_adj > 0 && grid_node > 0,
              "edm philox pre-step: bad shape / grid B=%d C_e=%d N=%d C_n=%d grids %d %d", batch, c_e, n, c_n, grid_adj,
              grid_node);
  EdmShape sh{batch, c_e, n, c_n};
  const long long na = static_cast<long long>(batch) * c_e * n * n, nn = static_cast<long long>(batch) * n * c_n;
  DSG_REQUIRE(na < 2147483647LL && nn < 2147483647LL, "edm philox pre-step: %lld / %lld elements (32-bit indexing)", na, nn);
  edm_pre_philox_kernel<true><<<static_cast<unsigned>(grid_adj), 256, 0, st>>>(adj, adj_hat, flags, noise_coef, seed,
                                                                               offset_adj, na, sh);
  DSG_LAUNCH_CHECK();
  edm_pre_philox_kernel<false><<<static_cast<unsigned>(grid_node), 256, 0, st>>>(node, node_hat, flags, noise_coef, seed,
                                                                                 offset_node, nn, sh);
  DSG_LAUNCH_CHECK();
  return DSG_OK;
}

int launch_edm_post_step(const float* adj_hat, const float* node_hat, const float* d1_adj, const float* d1_node,
                         const float* d2_adj, const float* d2_node, const uint8_t* flags, float inv_t_hat, float h,
                         float inv_t_prime, float* adj_next, float* node_next, int batch, int c_e, int n, int c_n,
                         cudaStream_t st) {
  if (d2_adj != nullptr)
    return launch_mode<1>(adj_hat, d1_adj, d2_adj, adj_next, node_hat, d1_node, d2_node, node_next, flags, inv_t_hat, h,
                          inv_t_prime, batch, c_e, n, c_n, st);
  return launch_mode<2>(adj_hat, d1_adj, nullptr, adj_next, node_hat, d1_node, nullptr, node_next, flags, inv_t_hat, h,
                        0.f, batch, c_e, n, c_n, st);
}

int launch_decode(const float* adj, const float* node, const uint8_t* flags, int32_t* adj_cls, int32_t* node_cls,
                  float* bbox, int num_adj_type, int num_node_type, int batch, int c_e, int n, int c_n, cudaStream_t st) {
  DSG_REQUIRE(batch > 0 && c_e > 0 && c_e <= 30 && n > 0 && c_n > 4 && c_n - 4 <= 30 && num_adj_type > 0 && num_node_type > 0,
              "decode: bad shape B=%d C_e=%d N=%d C_n=%d", batch, c_e, n, c_n);
  const long long pixels = static_cast<long long>(batch) * n * n;
  long long blocks = (pixels + 255) / 256;
  if (blocks > 148LL * 8) blocks = 148LL * 8;
  EdmShape sh{batch, c_e, n, c_n};
  decode_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(adj, node, flags, adj_cls, node_cls, bbox, num_adj_type,
                                                              num_node_type, sh);
  DSG_LAUNCH_CHECK();
  return DSG_OK;
}

int launch_mask_scale(const float* adj, const float* node, const uint8_t* flags, float scale, float* adj_out,
                      float* node_out, int batch, int c_e, int n, int c_n, cudaStream_t st) {
  return launch_mode<3>(adj, nullptr, nullptr, adj_out, node, nullptr, nullptr, node_out, flags, scale, 0.f, 0.f, batch,
                        c_e, n, c_n, st);
}

}  // namespace dsg
