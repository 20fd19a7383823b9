// Row N1 of SURVEY.md section 8(f): the residual + LayerNorm that follows MultiScaleDeformableAttention (and the FFN)
// in the reference's DeformableTransformerEncoderLayer (transformer_for_adapter.py:901-902, :882-885):
//     src = norm(src + dropout(src2))          (dropout is the identity: p = 0.0 in the ZiRa configuration)
// as ONE pass over the activation in each direction instead of add + LayerNorm (fwd) and LayerNorm-grad + add (bwd).
// 16-bit activations, fp32 statistics.  One warp per row; C <= 1024, C % 8 == 0.
//   forward : z = x + r (stored, 16-bit: it is LayerNorm's saved input), y = (z - mean) * rstd * gamma + beta
//   backward: dz = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma,  xhat = (z - mean) * rstd
//             (dz is the gradient of BOTH x and r)
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/msda_b200.h"
#include "layer_common.cuh"

namespace msda {
extern long long g_launches;
}

namespace {
using namespace msda_layer;

// NVEC = ceil(C / 256): 8-element vectors per lane (compile-time so that C = 256 keeps 8 values, not 32, in registers)
template <int kMaxVec>
__global__ void __launch_bounds__(256)
add_ln_fwd_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ r, const float* __restrict__ gamma,
                  const float* __restrict__ beta, long long R, int C, float eps, int is_half, uint16_t* __restrict__ z,
                  uint16_t* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                  uint16_t* __restrict__ y2, const float* __restrict__ shift2) {
  // r == nullptr: plain LayerNorm(x) (z is not written);  y2 != nullptr: a second output y2 = y + shift2[column], the
  // residual operand of a GEMM whose bias is folded in here (fuse_modules.py)
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= R) return;
  const int lane = threadIdx.x & 31;
  constexpr int nvec = kMaxVec;
  const bool h = is_half != 0;
  float v[kMaxVec][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (i < nvec && c < C) {
      float a[8], b[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(x + row * C + c)), h, a);
      if (r != nullptr) {
        unpack8(__ldg(reinterpret_cast<const uint4*>(r + row * C + c)), h, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = a[j] + b[j];
        // LayerNorm sees the 16-bit rounded sum, exactly as the unfused add -> LayerNorm sequence does
        const uint4 zz = pack8(v[i], h);
        *reinterpret_cast<uint4*>(z + row * C + c) = zz;
        unpack8(zz, h, v[i]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = a[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[i][j];
    }
  }
  const float mean = warp_sum(sum) / C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (i < nvec && c < C) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; sq = fmaf(d, d, sq); }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / C + eps);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (i < nvec && c < C) {
      float o[8];
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf((v[i][j] - mean) * rstd, gg[j], bb[j]);
      const uint4 yy = pack8(o, h);
      *reinterpret_cast<uint4*>(y + row * C + c) = yy;
      if (y2 != nullptr) {          // from the ROUNDED y, as the unfused `y + shift` would
        unpack8(yy, h, o);
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(shift2 + c)), s1 = __ldg(reinterpret_cast<const float4*>(shift2 + c + 4));
        const float ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += ss[j];
        *reinterpret_cast<uint4*>(y2 + row * C + c) = pack8(o, h);
      }
    }
  }
  if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
}

template <int kMaxVec>
__global__ void __launch_bounds__(256)
add_ln_bwd_kernel(const uint16_t* __restrict__ dy, const uint16_t* __restrict__ z, const float* __restrict__ gamma,
                  const float* __restrict__ mean_in, const float* __restrict__ rstd_in, long long R, int C, int is_half,
                  uint16_t* __restrict__ dz) {
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= R) return;
  const int lane = threadIdx.x & 31;
  constexpr int nvec = kMaxVec;
  const bool h = is_half != 0;
  const float mean = mean_in[row], rstd = rstd_in[row];
  float g[kMaxVec][8], xh[kMaxVec][8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (i < nvec && c < C) {
      float a[8], b[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(dy + row * C + c)), h, a);
      unpack8(__ldg(reinterpret_cast<const uint4*>(z + row * C + c)), h, b);
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        g[i][j] = a[j] * gg[j];
        xh[i][j] = (b[j] - mean) * rstd;
        s1 += g[i][j];
        s2 = fmaf(g[i][j], xh[i][j], s2);
      }
    }
  }
  const float m1 = warp_sum(s1) / C, m2 = warp_sum(s2) / C;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (i < nvec && c < C) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = rstd * (g[i][j] - m1 - xh[i][j] * m2);
      *reinterpret_cast<uint4*>(dz + row * C + c) = pack8(o, h);
    }
  }
}
}  // namespace

extern "C" {

int msda_add_layernorm_fwd_16(const void* x, const void* r, const float* gamma, const float* beta, long long R, int C,
                              float eps, void* z, void* y, float* mean, float* rstd, int is_half, void* stream) {
  if (!x || !r || !gamma || !beta || !z || !y || !mean || !rstd) return MSDA_ERR_NULL_POINTER;
  if (R <= 0 || C <= 0 || C % 8 || C > 1024) return MSDA_ERR_BAD_SHAPE;
  ++msda::g_launches;
  const unsigned grid = static_cast<unsigned>((R + 7) / 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define LN_FWD(NV) add_ln_fwd_kernel<NV><<<grid, 256, 0, st>>>(static_cast<const uint16_t*>(x), static_cast<const uint16_t*>(r), \
      gamma, beta, R, C, eps, is_half, static_cast<uint16_t*>(z), static_cast<uint16_t*>(y), mean, rstd, nullptr, nullptr)
  if (C <= 256) LN_FWD(1); else if (C <= 512) LN_FWD(2); else LN_FWD(4);
#undef LN_FWD
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

int msda_layernorm_fwd_16(const void* x, const float* gamma, const float* beta, long long R, int C, float eps, void* y,
                          void* y2, const float* shift2, float* mean, float* rstd, int is_half, void* stream) {
  if (!x || !gamma || !beta || !y || !mean || !rstd || (y2 && !shift2)) return MSDA_ERR_NULL_POINTER;
  if (R <= 0 || C <= 0 || C % 8 || C > 1024) return MSDA_ERR_BAD_SHAPE;
  ++msda::g_launches;
  const unsigned grid = static_cast<unsigned>((R + 7) / 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define LN_FWD(NV) add_ln_fwd_kernel<NV><<<grid, 256, 0, st>>>(static_cast<const uint16_t*>(x), nullptr, gamma, beta, R, C, eps, is_half, \
      nullptr, static_cast<uint16_t*>(y), mean, rstd, static_cast<uint16_t*>(y2), shift2)
  if (C <= 256) LN_FWD(1); else if (C <= 512) LN_FWD(2); else LN_FWD(4);
#undef LN_FWD
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

int msda_add_layernorm_bwd_16(const void* dy, const void* z, const float* gamma, const float* mean, const float* rstd,
                              long long R, int C, void* dz, int is_half, void* stream) {
  if (!dy || !z || !gamma || !mean || !rstd || !dz) return MSDA_ERR_NULL_POINTER;
  if (R <= 0 || C <= 0 || C % 8 || C > 1024) return MSDA_ERR_BAD_SHAPE;
  ++msda::g_launches;
  const unsigned grid = static_cast<unsigned>((R + 7) / 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define LN_BWD(NV) add_ln_bwd_kernel<NV><<<grid, 256, 0, st>>>(static_cast<const uint16_t*>(dy), static_cast<const uint16_t*>(z), \
      gamma, mean, rstd, R, C, is_half, static_cast<uint16_t*>(dz))
  if (C <= 256) LN_BWD(1); else if (C <= 512) LN_BWD(2); else LN_BWD(4);
#undef LN_BWD
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

}  // extern "C"
