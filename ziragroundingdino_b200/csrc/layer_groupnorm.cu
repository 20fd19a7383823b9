// Row N3 of SURVEY.md section 8(f): the GroupNorm that closes every level of the reference's input projection,
//     src_l = GroupNorm(32, 256)(conv_0(x_l) + adapter(x_l))       (groundingdino_dual_zero_rep_branch.py:258-277, :492-493)
// on the channels-last layout the projection GEMM produces ([N, H*W, C], i.e. already the flattened token layout the
// encoder consumes, transformer_for_adapter.py:258-262), so no NCHW round trip is needed.  16-bit activations, fp32 math,
// fp64 accumulation of the per-image statistics.
//   forward : two passes -- per-channel (sum, sum of squares) per image, then y = (x - mean_g) * rstd_g * gamma_c + beta_c
//   backward: two passes -- per-channel (sum dy*xhat, sum dy) per image (these ARE d gamma / d beta once summed over the
//             batch), then dx = rstd_g * (dy*gamma_c - xhat * A_g - B_g),  A_g = mean_g(dy*gamma*xhat), B_g = mean_g(dy*gamma)
// Every pointer may address a level slice of a larger [N, S, C] tensor: rows of image n start at n * image_stride.
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

// MODE 0: sums[n][c] = (sum x, sum x^2).  MODE 1: sums[n][c] = (sum dy * xhat, sum dy).
// 256 threads = (256 / (C/8)) rows x (C/8) channel groups of 8; a thread keeps its channel group on every trip.
template <int MODE>
__global__ void __launch_bounds__(256)
gn_channel_sums_kernel(const uint16_t* __restrict__ a, long long a_image_stride, const uint16_t* __restrict__ x,
                       long long x_image_stride, const float* __restrict__ mean_rstd, long long HW, int C, int G, int is_half,
                       double* __restrict__ sums) {
  extern __shared__ float s_acc[];   // [C][2]
  const int n = blockIdx.y;
  const int c8 = C / 8, rows_per_trip = 256 / c8;
  const int cg = threadIdx.x % c8, rslot = threadIdx.x / c8;
  const bool h = is_half != 0;
  const int cpg = C / G;
  float p[8], q[8], mu[8], rs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    p[j] = q[j] = 0.f;
    if (MODE == 1) {
      const int g = (cg * 8 + j) / cpg;
      mu[j] = __ldg(mean_rstd + (static_cast<long long>(n) * G + g) * 2);
      rs[j] = __ldg(mean_rstd + (static_cast<long long>(n) * G + g) * 2 + 1);
    }
  }
  const uint16_t* ap = a + n * a_image_stride + cg * 8;
  const uint16_t* xp = MODE == 1 ? x + n * x_image_stride + cg * 8 : nullptr;
  // four independent row loads in flight per thread (the loop is latency-bound otherwise)
  const long long step = static_cast<long long>(gridDim.x) * rows_per_trip;
  for (long long r0 = static_cast<long long>(blockIdx.x) * rows_per_trip + rslot; r0 < HW; r0 += 4 * step) {
    uint4 raw[4], rawx[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long r = r0 + u * step;
      if (r < HW) {
        raw[u] = __ldg(reinterpret_cast<const uint4*>(ap + r * C));
        if (MODE == 1) rawx[u] = __ldg(reinterpret_cast<const uint4*>(xp + r * C));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (r0 + u * step >= HW) break;
      float v[8];
      unpack8(raw[u], h, v);
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { p[j] += v[j]; q[j] = fmaf(v[j], v[j], q[j]); }
      } else {
        float xv[8];
        unpack8(rawx[u], h, xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) { p[j] = fmaf(v[j], (xv[j] - mu[j]) * rs[j], p[j]); q[j] += v[j]; }
      }
    }
  }
  for (int k = threadIdx.x; k < 2 * C; k += blockDim.x) s_acc[k] = 0.f;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(&s_acc[(cg * 8 + j) * 2], p[j]);
    atomicAdd(&s_acc[(cg * 8 + j) * 2 + 1], q[j]);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < 2 * C; k += blockDim.x)
    atomicAdd(sums + static_cast<long long>(n) * 2 * C + k, static_cast<double>(s_acc[k]));
}

__global__ void __launch_bounds__(256)
gn_fwd_apply_kernel(const uint16_t* __restrict__ x, long long x_image_stride, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const double* __restrict__ sums, long long HW, int C, int G, float eps,
                    int is_half, uint16_t* __restrict__ y, long long y_image_stride, float* __restrict__ mean_rstd_out) {
  extern __shared__ float s_mr[];   // [G][2]
  const int n = blockIdx.y;
  const int cpg = C / G, c8 = C / 8;
  const bool h = is_half != 0;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    double S = 0.0, Q = 0.0;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      S += sums[(static_cast<long long>(n) * C + c) * 2];
      Q += sums[(static_cast<long long>(n) * C + c) * 2 + 1];
    }
    const double cnt = static_cast<double>(HW) * cpg;
    const double mean = S / cnt;
    double var = Q / cnt - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    s_mr[2 * g] = static_cast<float>(mean);
    s_mr[2 * g + 1] = rstd;
    if (blockIdx.x == 0) {
      mean_rstd_out[(static_cast<long long>(n) * G + g) * 2] = static_cast<float>(mean);
      mean_rstd_out[(static_cast<long long>(n) * G + g) * 2 + 1] = rstd;
    }
  }
  __syncthreads();
  // blockDim.x (256) is a multiple of C/8, so a thread keeps its 8 channels on every trip: constants hoisted
  const int c0 = static_cast<int>(threadIdx.x % c8) * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (c0 + j) / cpg;
    sc[j] = s_mr[2 * g + 1] * __ldg(gamma + c0 + j);
    sh[j] = fmaf(-s_mr[2 * g], sc[j], __ldg(beta + c0 + j));
  }
  const int rows_per_trip = 256 / c8;
  const long long step = static_cast<long long>(gridDim.x) * rows_per_trip;
  const uint16_t* xp = x + n * x_image_stride + c0;
  uint16_t* yp = y + n * y_image_stride + c0;
  for (long long r0 = static_cast<long long>(blockIdx.x) * rows_per_trip + threadIdx.x / c8; r0 < HW; r0 += 2 * step) {
    const long long r1 = r0 + step;
    const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(xp + r0 * C));
    uint4 a1 = a0;
    if (r1 < HW) a1 = __ldg(reinterpret_cast<const uint4*>(xp + r1 * C));
    float v[8];
    unpack8(a0, h, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
    *reinterpret_cast<uint4*>(yp + r0 * C) = pack8(v, h);
    if (r1 < HW) {
      unpack8(a1, h, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
      *reinterpret_cast<uint4*>(yp + r1 * C) = pack8(v, h);
    }
  }
}

__global__ void __launch_bounds__(256)
gn_bwd_apply_kernel(const uint16_t* __restrict__ dy, long long dy_image_stride, const uint16_t* __restrict__ x,
                    long long x_image_stride, const float* __restrict__ gamma, const float* __restrict__ mean_rstd,
                    const double* __restrict__ sums, long long HW, int C, int G, int is_half, uint16_t* __restrict__ dx,
                    long long dx_image_stride) {
  extern __shared__ float s_g[];   // [G][4] = mean, rstd, A, B
  const int n = blockIdx.y;
  const int cpg = C / G, c8 = C / 8;
  const bool h = is_half != 0;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    double A = 0.0, B = 0.0;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      const double gm = static_cast<double>(__ldg(gamma + c));
      A += gm * sums[(static_cast<long long>(n) * C + c) * 2];
      B += gm * sums[(static_cast<long long>(n) * C + c) * 2 + 1];
    }
    const double cnt = static_cast<double>(HW) * cpg;
    s_g[4 * g] = __ldg(mean_rstd + (static_cast<long long>(n) * G + g) * 2);
    s_g[4 * g + 1] = __ldg(mean_rstd + (static_cast<long long>(n) * G + g) * 2 + 1);
    s_g[4 * g + 2] = static_cast<float>(A / cnt);
    s_g[4 * g + 3] = static_cast<float>(B / cnt);
  }
  __syncthreads();
  // dx = dy * a - x * b - c with per-channel constants a = rstd*gamma, b = rstd^2*A, c = rstd*(B - mean*rstd*A)
  const int c0 = static_cast<int>(threadIdx.x % c8) * 8;
  float ka[8], kb[8], kc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (c0 + j) / cpg;
    const float mean = s_g[4 * g], rstd = s_g[4 * g + 1], A = s_g[4 * g + 2], B = s_g[4 * g + 3];
    ka[j] = rstd * __ldg(gamma + c0 + j);
    kb[j] = rstd * rstd * A;
    kc[j] = rstd * (B - mean * rstd * A);
  }
  const int rows_per_trip = 256 / c8;
  const long long step = static_cast<long long>(gridDim.x) * rows_per_trip;
  const uint16_t* dyp = dy + n * dy_image_stride + c0;
  const uint16_t* xp = x + n * x_image_stride + c0;
  uint16_t* dxp = dx + n * dx_image_stride + c0;
  for (long long r0 = static_cast<long long>(blockIdx.x) * rows_per_trip + threadIdx.x / c8; r0 < HW; r0 += 2 * step) {
    const long long r1 = r0 + step;
    const uint4 d0 = __ldg(reinterpret_cast<const uint4*>(dyp + r0 * C));
    const uint4 x0 = __ldg(reinterpret_cast<const uint4*>(xp + r0 * C));
    uint4 d1 = d0, x1 = x0;
    if (r1 < HW) {
      d1 = __ldg(reinterpret_cast<const uint4*>(dyp + r1 * C));
      x1 = __ldg(reinterpret_cast<const uint4*>(xp + r1 * C));
    }
    float d[8], xv[8];
    unpack8(d0, h, d); unpack8(x0, h, xv);
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] = fmaf(d[j], ka[j], -fmaf(xv[j], kb[j], kc[j]));
    *reinterpret_cast<uint4*>(dxp + r0 * C) = pack8(d, h);
    if (r1 < HW) {
      unpack8(d1, h, d); unpack8(x1, h, xv);
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = fmaf(d[j], ka[j], -fmaf(xv[j], kb[j], kc[j]));
      *reinterpret_cast<uint4*>(dxp + r1 * C) = pack8(d, h);
    }
  }
}

int check_shape(int N, long long HW, int C, int G) {
  if (N <= 0 || HW <= 0 || C <= 0 || G <= 0 || C % 8 || C % G) return MSDA_ERR_BAD_SHAPE;
  const int c8 = C / 8;
  if (c8 > 256 || 256 % c8) return MSDA_ERR_UNSUPPORTED;   // a thread must keep its channel group on every trip
  return 0;
}

dim3 sums_grid(int N, long long HW, int C) {
  const int rows_per_trip = 256 / (C / 8);
  long long bx = (HW + rows_per_trip * 8 - 1) / (rows_per_trip * 8);   // >= 8 trips per CTA before the flush
  const long long cap = (148 * 8 + N - 1) / N;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  return dim3(static_cast<unsigned>(bx), static_cast<unsigned>(N));
}

dim3 apply_grid(int N, long long HW, int C) {
  const int rows_per_trip = 256 / (C / 8);
  long long bx = (HW + 2 * rows_per_trip - 1) / (2 * rows_per_trip);   // two rows per thread per trip
  const long long cap = (148 * 16 + N - 1) / N;
  if (bx > cap) bx = cap;
  return dim3(static_cast<unsigned>(bx), static_cast<unsigned>(N));
}
}  // namespace

extern "C" {

int msda_group_norm_fwd_16(const void* x, long long x_image_stride, const float* gamma, const float* beta, int N, long long HW,
                           int C, int G, float eps, void* y, long long y_image_stride, float* mean_rstd, double* scratch,
                           int is_half, void* stream) {
  if (!x || !gamma || !beta || !y || !mean_rstd || !scratch) return MSDA_ERR_NULL_POINTER;
  if (int rc = check_shape(N, HW, C, G)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  msda::g_launches += 2;
  gn_channel_sums_kernel<0><<<sums_grid(N, HW, C), 256, 2 * C * sizeof(float), st>>>(
      static_cast<const uint16_t*>(x), x_image_stride, nullptr, 0, nullptr, HW, C, G, is_half, scratch);
  gn_fwd_apply_kernel<<<apply_grid(N, HW, C), 256, 2 * G * sizeof(float), st>>>(
      static_cast<const uint16_t*>(x), x_image_stride, gamma, beta, scratch, HW, C, G, eps, is_half,
      static_cast<uint16_t*>(y), y_image_stride, mean_rstd);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

int msda_group_norm_bwd_16(const void* dy, long long dy_image_stride, const void* x, long long x_image_stride,
                           const float* gamma, const float* mean_rstd, int N, long long HW, int C, int G, void* dx,
                           long long dx_image_stride, double* scratch, int is_half, void* stream) {
  if (!dy || !x || !gamma || !mean_rstd || !dx || !scratch) return MSDA_ERR_NULL_POINTER;
  if (int rc = check_shape(N, HW, C, G)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  msda::g_launches += 2;
  gn_channel_sums_kernel<1><<<sums_grid(N, HW, C), 256, 2 * C * sizeof(float), st>>>(
      static_cast<const uint16_t*>(dy), dy_image_stride, static_cast<const uint16_t*>(x), x_image_stride, mean_rstd, HW, C, G,
      is_half, scratch);
  gn_bwd_apply_kernel<<<apply_grid(N, HW, C), 256, 4 * G * sizeof(float), st>>>(
      static_cast<const uint16_t*>(dy), dy_image_stride, static_cast<const uint16_t*>(x), x_image_stride, gamma, mean_rstd,
      scratch, HW, C, G, is_half, static_cast<uint16_t*>(dx), dx_image_stride);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

}  // extern "C"
