// Shared helpers of the layer-level kernels (residual + LayerNorm, GroupNorm): 8 x 16-bit <-> fp32 packing, warp sum.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace msda_layer {

__device__ __forceinline__ void unpack8(const uint4& r, bool h, float (&f)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (h) {
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      f[2 * i] = t.x; f[2 * i + 1] = t.y;
    } else {
      f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8], bool h) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (h) { __half2 t = __floats2half2_rn(f[2 * i], f[2 * i + 1]); w[i] = *reinterpret_cast<uint32_t*>(&t); }
    else { __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]); w[i] = *reinterpret_cast<uint32_t*>(&t); }
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace msda_layer
