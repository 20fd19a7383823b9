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
// ROWS rows per warp (C <= 256 only): the loads of both rows are issued before either reduction, so a warp keeps twice the
// bytes in flight across its load -> shuffle-reduce -> shuffle-reduce -> store chain (one row per warp: 0.73 of the copy rate).
template <int kMaxVec, int ROWS = 1>
__global__ void __launch_bounds__(256)
add_ln_fwd_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ r, const float* __restrict__ gamma,
                  const float* __restrict__ beta, long long R, int C, float eps, int is_half, uint16_t* __restrict__ z,
                  uint16_t* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                  uint16_t* __restrict__ y2, const float* __restrict__ shift2) {
  // r == nullptr: plain LayerNorm(x) (z is not written);  y2 != nullptr: a second output y2 = y + shift2[column], the
  // residual operand of a GEMM whose bias is folded in here (fuse_modules.py)
  const long long row0 = (static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5)) * ROWS;
  if (row0 >= R) return;
  const int lane = threadIdx.x & 31;
  constexpr int nvec = kMaxVec;
  const bool h = is_half != 0;
  float v[ROWS][kMaxVec][8];
  float sum[ROWS];
  uint4 xa[ROWS][kMaxVec], ra[ROWS][kMaxVec];
#pragma unroll
  for (int q = 0; q < ROWS; ++q) {
    const long long row = row0 + q < R ? row0 + q : R - 1;        // a tail warp's second row repeats the last one (not stored)
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (i < nvec && c < C) {
        xa[q][i] = __ldg(reinterpret_cast<const uint4*>(x + row * C + c));
        if (r != nullptr) ra[q][i] = __ldg(reinterpret_cast<const uint4*>(r + row * C + c));
      }
    }
  }
#pragma unroll
  for (int q = 0; q < ROWS; ++q) {
    const long long row = row0 + q;
    const bool live = row < R;
    sum[q] = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (i < nvec && c < C) {
        float a[8], b[8];
        unpack8(xa[q][i], h, a);
        if (r != nullptr) {
          unpack8(ra[q][i], h, b);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[q][i][j] = a[j] + b[j];
          // LayerNorm sees the 16-bit rounded sum, exactly as the unfused add -> LayerNorm sequence does
          const uint4 zz = pack8(v[q][i], h);
          if (live) *reinterpret_cast<uint4*>(z + row * C + c) = zz;
          unpack8(zz, h, v[q][i]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[q][i][j] = a[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) sum[q] += v[q][i][j];
      }
    }
  }
  float mean[ROWS], rstd[ROWS];
#pragma unroll
  for (int q = 0; q < ROWS; ++q) mean[q] = warp_sum(sum[q]) / C;
#pragma unroll
  for (int q = 0; q < ROWS; ++q) {
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (i < nvec && c < C) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[q][i][j] - mean[q]; sq = fmaf(d, d, sq); }
      }
    }
    rstd[q] = sq;
  }
#pragma unroll
  for (int q = 0; q < ROWS; ++q) rstd[q] = rsqrtf(warp_sum(rstd[q]) / C + eps);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (i < nvec && c < C) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int q = 0; q < ROWS; ++q) {
        const long long row = row0 + q;
        if (row >= R) continue;
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf((v[q][i][j] - mean[q]) * rstd[q], gg[j], bb[j]);
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
  }
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < ROWS; ++q)
      if (row0 + q < R) { mean_out[row0 + q] = mean[q]; rstd_out[row0 + q] = rstd[q]; }
  }
}

template <int kMaxVec, int ROWS = 1>
__global__ void __launch_bounds__(256)
add_ln_bwd_kernel(const uint16_t* __restrict__ dy, const uint16_t* __restrict__ z, const float* __restrict__ gamma,
                  const float* __restrict__ mean_in, const float* __restrict__ rstd_in, long long R, int C, int is_half,
                  uint16_t* __restrict__ dz) {
  const long long row0 = (static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5)) * ROWS;
  if (row0 >= R) return;
  const int lane = threadIdx.x & 31;
  constexpr int nvec = kMaxVec;
  const bool h = is_half != 0;
  uint4 da[ROWS][kMaxVec], za[ROWS][kMaxVec];
  float mean[ROWS], rstd[ROWS];
#pragma unroll
  for (int q = 0; q < ROWS; ++q) {
    const long long row = row0 + q < R ? row0 + q : R - 1;
    mean[q] = mean_in[row]; rstd[q] = rstd_in[row];
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (i < nvec && c < C) {
        da[q][i] = __ldg(reinterpret_cast<const uint4*>(dy + row * C + c));
        za[q][i] = __ldg(reinterpret_cast<const uint4*>(z + row * C + c));
      }
    }
  }
  float g[ROWS][kMaxVec][8], xh[ROWS][kMaxVec][8];
  float s1[ROWS], s2[ROWS];
#pragma unroll
  for (int q = 0; q < ROWS; ++q) { s1[q] = 0.f; s2[q] = 0.f; }
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (i < nvec && c < C) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int q = 0; q < ROWS; ++q) {
        float a[8], b[8];
        unpack8(da[q][i], h, a);
        unpack8(za[q][i], h, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          g[q][i][j] = a[j] * gg[j];
          xh[q][i][j] = (b[j] - mean[q]) * rstd[q];
          s1[q] += g[q][i][j];
          s2[q] = fmaf(g[q][i][j], xh[q][i][j], s2[q]);
        }
      }
    }
  }
  float m1[ROWS], m2[ROWS];
#pragma unroll
  for (int q = 0; q < ROWS; ++q) { m1[q] = warp_sum(s1[q]) / C; m2[q] = warp_sum(s2[q]) / C; }
#pragma unroll
  for (int q = 0; q < ROWS; ++q) {
    const long long row = row0 + q;
    if (row >= R) continue;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (i < nvec && c < C) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd[q] * (g[q][i][j] - m1[q] - xh[q][i][j] * m2[q]);
        *reinterpret_cast<uint4*>(dz + row * C + c) = pack8(o, h);
      }
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
  const unsigned grid = static_cast<unsigned>((R + 7) / 8), grid2 = static_cast<unsigned>((R + 15) / 16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define LN_FWD(NV) add_ln_fwd_kernel<NV><<<grid, 256, 0, st>>>(static_cast<const uint16_t*>(x), static_cast<const uint16_t*>(r), \
      gamma, beta, R, C, eps, is_half, static_cast<uint16_t*>(z), static_cast<uint16_t*>(y), mean, rstd, nullptr, nullptr)
  if (C <= 256)
    add_ln_fwd_kernel<1, 2><<<grid2, 256, 0, st>>>(static_cast<const uint16_t*>(x), static_cast<const uint16_t*>(r), gamma, beta, R, C, eps,
                                                   is_half, static_cast<uint16_t*>(z), static_cast<uint16_t*>(y), mean, rstd, nullptr, nullptr);
  else if (C <= 512) LN_FWD(2); else LN_FWD(4);
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
  const unsigned grid = static_cast<unsigned>((R + 7) / 8), grid2 = static_cast<unsigned>((R + 15) / 16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define LN_BWD(NV) add_ln_bwd_kernel<NV><<<grid, 256, 0, st>>>(static_cast<const uint16_t*>(dy), static_cast<const uint16_t*>(z), \
      gamma, mean, rstd, R, C, is_half, static_cast<uint16_t*>(dz))
  // two rows per warp measured no gain here (26.9 vs 27.3 us at 4 images: the backward already has three loads per row in
  // flight), so the backward keeps one row per warp; the forward gains 13 % (38.0 -> 33.0 us)
  (void)grid2;
  if (C <= 256) LN_BWD(1); else if (C <= 512) LN_BWD(2); else LN_BWD(4);
#undef LN_BWD
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

}  // extern "C"
