// The step either side of the encoder / decoder stacks on the device.
//
//  * flatten_levels: the reference's per-level `src.flatten(2).transpose(1, 2)`, `mask.flatten(1)`,
//    `pos_embed.flatten(2).transpose(1, 2) + level_embed[lvl]` and the three torch.cat calls
//    (transformer_for_adapter.py:238-262) -- ~6 eager kernels per level -- as ONE launch: a tiled NCHW -> [N, S, C]
//    transposition for all levels at once, the level embedding added on the way (fp32 add, rounded once to the storage
//    type: bit-identical to the eager `pos + level_embed`), the masks copied by the same CTAs.
//  * level_valid_counts: valid_W / valid_H per (image, level) from the flattened mask (utils.py:74-75; the numerators of
//    get_valid_ratio, transformer_for_adapter.py:216-223).
//  * encoder_proposals: gen_encoder_output_proposals (utils.py:56-116) in one pass over the encoder memory: per token the
//    (cx, cy, w, h) proposal, its validity test, the inverse sigmoid, and the masked copy of the memory row.  Same fp32
//    operation sequence as the reference (IEEE division, logf), so results equal torch's on the same device.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/msda_b200.h"

namespace msda {
extern long long g_launches;
}

namespace {

struct LevelsArg {
  const void* src[MSDA_MAX_LEVELS];
  const void* pos[MSDA_MAX_LEVELS];
  const uint8_t* mask[MSDA_MAX_LEVELS];   // [N, H*W] per level (may be null: no mask output)
  int hw[MSDA_MAX_LEVELS], start[MSDA_MAX_LEVELS], tile0[MSDA_MAX_LEVELS + 1];
  int L;
};

template <typename T> __device__ __forceinline__ float ld_f(T v);
template <> __device__ __forceinline__ float ld_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float ld_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float ld_f<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T st_f(float v);
template <> __device__ __forceinline__ float st_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 st_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half st_f<__half>(float v) { return __float2half_rn(v); }

// grid: x = 32-token tiles of all levels, y = 32-channel tiles, z = image.  256 threads = 32 x 8.
template <typename T>
__global__ void __launch_bounds__(256)
flatten_levels_kernel(LevelsArg a, const T* __restrict__ level_embed, int C, int S, T* __restrict__ src_out, T* __restrict__ pos_out,
                      uint8_t* __restrict__ mask_out) {
  __shared__ T ts[32][33], tp[32][33];
  int l = 0;
  while (l + 1 < a.L && static_cast<int>(blockIdx.x) >= a.tile0[l + 1]) ++l;
  const int p0 = (static_cast<int>(blockIdx.x) - a.tile0[l]) * 32, c0 = blockIdx.y * 32, n = blockIdx.z;
  const int hw = a.hw[l];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const T* s = static_cast<const T*>(a.src[l]) + static_cast<size_t>(n) * C * hw;
  const T* p = a.pos[l] ? static_cast<const T*>(a.pos[l]) + static_cast<size_t>(n) * C * hw : nullptr;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k, px = p0 + tx;
    if (c < C && px < hw) {
      if (src_out) ts[ty + 8 * k][tx] = s[static_cast<size_t>(c) * hw + px];
      if (p) tp[ty + 8 * k][tx] = p[static_cast<size_t>(c) * hw + px];
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int px = p0 + ty + 8 * k, c = c0 + tx;
    if (c < C && px < hw) {
      const size_t o = (static_cast<size_t>(n) * S + a.start[l] + px) * C + c;
      if (src_out) src_out[o] = ts[tx][ty + 8 * k];
      if (p) {
        const T v = tp[tx][ty + 8 * k];
        pos_out[o] = level_embed ? st_f<T>(ld_f<T>(v) + ld_f<T>(level_embed[static_cast<size_t>(l) * C + c])) : v;
      }
    }
  }
  if (blockIdx.y == 0 && mask_out && a.mask[l] && threadIdx.x < 32 && p0 + tx < hw)
    mask_out[static_cast<size_t>(n) * S + a.start[l] + p0 + tx] = a.mask[l][static_cast<size_t>(n) * hw + p0 + tx];
}

// one warp per (image, level): counts of valid (mask == 0) pixels along the first row (-> valid_W) and first column (-> valid_H)
__global__ void __launch_bounds__(32)
level_valid_counts_kernel(const uint8_t* __restrict__ mask, const int64_t* __restrict__ shapes, const int64_t* __restrict__ lstart,
                          int S, int L, int* __restrict__ counts) {
  const int l = blockIdx.x % L, n = blockIdx.x / L;
  const int H = static_cast<int>(shapes[2 * l]), W = static_cast<int>(shapes[2 * l + 1]);
  const uint8_t* m = mask + static_cast<size_t>(n) * S + lstart[l];
  int vw = 0, vh = 0;
  for (int x = threadIdx.x; x < W; x += 32) vw += m[x] == 0;
  for (int y = threadIdx.x; y < H; y += 32) vh += m[static_cast<size_t>(y) * W] == 0;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) { vw += __shfl_xor_sync(0xffffffffu, vw, o); vh += __shfl_xor_sync(0xffffffffu, vh, o); }
  if (threadIdx.x == 0) { counts[(n * L + l) * 2] = vw; counts[(n * L + l) * 2 + 1] = vh; }
}

// 32 tokens per CTA: lanes 0-31 of warp 0 compute the proposals and the drop flags, then every warp copies rows.
// mem rows are `row_bytes` long (multiple of 16).
__global__ void __launch_bounds__(256)
encoder_proposals_kernel(const uint4* __restrict__ memory, const uint8_t* __restrict__ mask, const int64_t* __restrict__ shapes,
                         const int64_t* __restrict__ lstart, const int* __restrict__ counts, const float* __restrict__ wh_base,
                         int N, int S, int L, int row16, uint4* __restrict__ mem_out, float4* __restrict__ prop_out) {
  __shared__ int s_drop[32];
  const long long t0 = static_cast<long long>(blockIdx.x) * 32;
  const long long total = static_cast<long long>(N) * S;
  if (threadIdx.x < 32) {
    const long long t = t0 + threadIdx.x;
    if (t < total) {
      const int n = static_cast<int>(t / S), s = static_cast<int>(t % S);
      int l = 0;
      while (l + 1 < L && s >= static_cast<int>(lstart[l + 1])) ++l;
      const int W = static_cast<int>(shapes[2 * l + 1]);
      const int o = s - static_cast<int>(lstart[l]);
      const float gx = static_cast<float>(o % W), gy = static_cast<float>(o / W);
      const float vw = static_cast<float>(counts[(n * L + l) * 2]), vh = static_cast<float>(counts[(n * L + l) * 2 + 1]);
      const float sc = exp2f(static_cast<float>(l));            // 2.0 ** lvl, exact
      float v[4] = {__fdiv_rn(gx + 0.5f, vw), __fdiv_rn(gy + 0.5f, vh), wh_base[0] * sc, wh_base[1] * sc};
      bool ok = true;
#pragma unroll
      for (int i = 0; i < 4; ++i) ok = ok && (v[i] > 0.01f) && (v[i] < 0.99f);
      const bool drop = !ok || mask[t] != 0;
      float4 out;
      float* po = reinterpret_cast<float*>(&out);
#pragma unroll
      for (int i = 0; i < 4; ++i) po[i] = drop ? CUDART_INF_F : logf(__fdiv_rn(v[i], 1.0f - v[i]));
      prop_out[t] = out;
      s_drop[threadIdx.x] = drop ? 1 : 0;
    } else {
      s_drop[threadIdx.x] = 1;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * row16; i += 256) {
    const int r = i / row16;
    const long long t = t0 + r;
    if (t < total) {
      const size_t o = static_cast<size_t>(t) * row16 + (i % row16);
      mem_out[o] = s_drop[r] ? make_uint4(0u, 0u, 0u, 0u) : memory[o];
    }
  }
}

template <typename T>
cudaError_t launch_flatten(const LevelsArg& a, const void* level_embed, int N, int C, int S, void* src_out, void* pos_out,
                           uint8_t* mask_out, cudaStream_t st) {
  const dim3 grid(static_cast<unsigned>(a.tile0[a.L]), static_cast<unsigned>((C + 31) / 32), static_cast<unsigned>(N));
  flatten_levels_kernel<T><<<grid, 256, 0, st>>>(a, static_cast<const T*>(level_embed), C, S, static_cast<T*>(src_out),
                                                  static_cast<T*>(pos_out), mask_out);
  return cudaGetLastError();
}

}  // namespace

extern "C" {

int msda_flatten_levels(const void* const* src_levels, const void* const* pos_levels, const uint8_t* const* mask_levels,
                        const int* level_hw, int L, int N, int C, const void* level_embed, int dtype, void* src_out, void* pos_out,
                        uint8_t* mask_out, void* stream) {
  if (!src_levels || !level_hw || L <= 0 || L > MSDA_MAX_LEVELS || N <= 0 || C <= 0) return MSDA_ERR_BAD_SHAPE;
  if (dtype < 0 || dtype > 2) return MSDA_ERR_UNSUPPORTED;
  LevelsArg a;
  a.L = L;
  int start = 0, tiles = 0;
  for (int l = 0; l < L; ++l) {
    if (!src_levels[l] || level_hw[l] <= 0) return MSDA_ERR_NULL_POINTER;
    a.src[l] = src_levels[l];
    a.pos[l] = pos_levels ? pos_levels[l] : nullptr;
    a.mask[l] = mask_levels ? mask_levels[l] : nullptr;
    a.hw[l] = level_hw[l];
    a.start[l] = start;
    a.tile0[l] = tiles;
    start += level_hw[l];
    tiles += (level_hw[l] + 31) / 32;
  }
  a.tile0[L] = tiles;
  if (N > 65535) return MSDA_ERR_BAD_SHAPE;
  ++msda::g_launches;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = dtype == 2 ? launch_flatten<float>(a, level_embed, N, C, start, src_out, pos_out, mask_out, st)
                  : dtype == 1 ? launch_flatten<__half>(a, level_embed, N, C, start, src_out, pos_out, mask_out, st)
                               : launch_flatten<__nv_bfloat16>(a, level_embed, N, C, start, src_out, pos_out, mask_out, st);
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

int msda_level_valid_counts(const uint8_t* mask_flatten, const int64_t* spatial_shapes, const int64_t* level_start_index, int N,
                            int S, int L, int* counts, void* stream) {
  if (!mask_flatten || !spatial_shapes || !level_start_index || !counts) return MSDA_ERR_NULL_POINTER;
  if (N <= 0 || S <= 0 || L <= 0 || L > MSDA_MAX_LEVELS) return MSDA_ERR_BAD_SHAPE;
  ++msda::g_launches;
  level_valid_counts_kernel<<<N * L, 32, 0, static_cast<cudaStream_t>(stream)>>>(mask_flatten, spatial_shapes, level_start_index, S, L, counts);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

int msda_encoder_proposals(const void* memory, const uint8_t* mask_flatten, const int64_t* spatial_shapes,
                           const int64_t* level_start_index, const int* counts, const float* wh_base, int N, int S, int L,
                           int row_bytes, void* memory_out, float* proposals_out, void* stream) {
  if (!memory || !mask_flatten || !spatial_shapes || !level_start_index || !counts || !wh_base || !memory_out || !proposals_out)
    return MSDA_ERR_NULL_POINTER;
  if (N <= 0 || S <= 0 || L <= 0 || L > MSDA_MAX_LEVELS || row_bytes <= 0 || row_bytes % 16) return MSDA_ERR_BAD_SHAPE;
  if ((reinterpret_cast<uintptr_t>(memory) | reinterpret_cast<uintptr_t>(memory_out) | reinterpret_cast<uintptr_t>(proposals_out)) & 15u)
    return MSDA_ERR_MISALIGNED;
  const long long total = static_cast<long long>(N) * S;
  ++msda::g_launches;
  encoder_proposals_kernel<<<static_cast<unsigned>((total + 31) / 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(memory), mask_flatten, spatial_shapes, level_start_index, counts, wh_base, N, S, L, row_bytes / 16,
      static_cast<uint4*>(memory_out), reinterpret_cast<float4*>(proposals_out));
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

}  // extern "C"
