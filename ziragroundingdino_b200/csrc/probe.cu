// Measurement aid (bench.py / tools only): achievable L2 gather bandwidth on this GPU.
//
// MEASURED_PEAKS.json has no L2 figure, and the gather kernels are bounded by how fast SMs can pull
// random, row-sized segments out of an L2-resident map.  This probe reproduces exactly that access
// shape with nothing else in the way: groups of `seg_bytes/16` lanes read random seg_bytes-aligned
// segments (128 B = one fp32 head row, 64 B = one bf16 head row) of a buffer small enough to live in
// L2, eight independent 128-bit loads in flight per lane, no arithmetic beyond an XOR sink.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/msda_b200.h"

namespace {
__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

template <int LPG>
__global__ void __launch_bounds__(256) gather_probe_kernel(const uint4* __restrict__ buf, uint32_t nseg, int iters,
                                                           uint32_t* __restrict__ sink) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t group = tid / LPG, lig = tid % LPG;
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (int i = 0; i < iters; i += 8) {
    uint4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t seg = mix(group * 2654435761u + static_cast<uint32_t>(i + k) * 40503u) % nseg;
      v[k] = __ldg(buf + static_cast<size_t>(seg) * LPG + lig);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) { acc.x ^= v[k].x; acc.y ^= v[k].y; acc.z ^= v[k].z; acc.w ^= v[k].w; }
  }
  if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) sink[0] = tid;  // keep the loads alive
}

// Scatter twin of the probe above: groups of 8 lanes send one 128-byte fp32 row (red.global.add.v4.f32 per lane, the
// instruction the backward kernel uses for grad_value) to random 128 B-aligned rows of an L2-sized buffer, nothing else
// in the way.  MODE 0 = red.v4.f32, 1 = st.global.v4 (plain store, same shape), 2 = scalar red.f32 (32 lanes per row).
template <int MODE>
__global__ void __launch_bounds__(256) scatter_probe_kernel(float* __restrict__ buf, uint32_t nrow, int iters) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  constexpr int LPG = MODE == 2 ? 32 : 8;
  const uint32_t group = tid / LPG, lig = tid % LPG;
  const float v = 1e-6f * static_cast<float>(lig);
  for (int i = 0; i < iters; ++i) {
    const uint32_t row = mix(group * 2654435761u + static_cast<uint32_t>(i) * 40503u) % nrow;
    float* p = buf + static_cast<size_t>(row) * 32;
    if (MODE == 0)
      asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p + lig * 4), "f"(v), "f"(v), "f"(v), "f"(v) : "memory");
    else if (MODE == 1)
      asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p + lig * 4), "f"(v), "f"(v), "f"(v), "f"(v) : "memory");
    else
      asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" :: "l"(p + lig), "f"(v) : "memory");
  }
}

// Round-2 shapes (is the scatter bound by REQUESTS or by BYTES on the SM->L2 path?)
//   MODE 3: 64-byte fp32 rows (4 lanes x red.v4.f32)            -- half the bytes per row, same row count
//   MODE 4: 64-byte bf16 rows (4 lanes x red.v4.bf16x2 = 32 ch) -- packed 16-bit reduction, half the bytes of a 128 B fp32 row
//   MODE 5: 256-byte contiguous fp32 (16 lanes x red.v4.f32)    -- two neighbouring rows as one warp-contiguous request
//   MODE 6: 128-byte rows staged in shared memory, sent by cp.reduce.async.bulk.global.shared::cta.add.f32 (TMA path,
//           one elected lane per 4-row warp tile) -- moves the reduction off the LSU -> XBAR request path
template <int MODE>
__global__ void __launch_bounds__(256) scatter_probe2_kernel(float* __restrict__ buf, uint32_t nrow128, int iters) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (MODE == 6) {
    __shared__ __align__(128) float stage[8][2][4 * 32];    // per warp: 2 buffers x 4 rows x 128 B
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t wg = tid >> 5;
    for (int i = 0; i < iters; ++i) {
      float* sb = stage[warp][i & 1];
      if (i >= 2 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      __syncwarp();
      *reinterpret_cast<float4*>(sb + lane * 4) = make_float4(1e-6f, 1e-6f, 1e-6f, 1e-6f);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane < 4) {
        const uint32_t row = mix((wg * 4 + lane) * 2654435761u + static_cast<uint32_t>(i) * 40503u) % nrow128;
        float* dst = buf + static_cast<size_t>(row) * 32;
        const uint32_t src = static_cast<uint32_t>(__cvta_generic_to_shared(sb + lane * 32));
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 128;" :: "l"(dst), "r"(src) : "memory");
      }
      if (lane == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    return;
  }
  constexpr int LPG = MODE == 5 ? 16 : 4;
  const uint32_t group = tid / LPG, lig = tid % LPG;
  const float v = 1e-6f * static_cast<float>(lig);
  const uint32_t nseg = MODE == 5 ? nrow128 / 2 : nrow128 * 2;     // 256 B / 64 B segments
  for (int i = 0; i < iters; ++i) {
    const uint32_t seg = mix(group * 2654435761u + static_cast<uint32_t>(i) * 40503u) % nseg;
    float* p = buf + static_cast<size_t>(seg) * (MODE == 5 ? 64 : 16) + lig * 4;
    if (MODE == 4) {
      const uint32_t w = 0x3c003c00u;
      asm volatile("red.relaxed.gpu.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" :: "l"(p), "r"(w), "r"(w), "r"(w), "r"(w) : "memory");
    } else {
      asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(v), "f"(v), "f"(v), "f"(v) : "memory");
    }
  }
}
}  // namespace

extern "C" int msda_b200_probe_scatter(void* buf, long long bytes, int mode, int iters, int blocks, void* stream) {
  if (!buf || bytes < 128 || iters <= 0 || blocks <= 0) return MSDA_ERR_BAD_SHAPE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint32_t nrow = static_cast<uint32_t>(bytes / 128);
  float* b = static_cast<float*>(buf);
  if (mode == 0) scatter_probe_kernel<0><<<blocks, 256, 0, st>>>(b, nrow, iters);
  else if (mode == 1) scatter_probe_kernel<1><<<blocks, 256, 0, st>>>(b, nrow, iters);
  else if (mode == 2) scatter_probe_kernel<2><<<blocks, 256, 0, st>>>(b, nrow, iters);
  else if (mode == 3) scatter_probe2_kernel<3><<<blocks, 256, 0, st>>>(b, nrow, iters);
  else if (mode == 4) scatter_probe2_kernel<4><<<blocks, 256, 0, st>>>(b, nrow, iters);
  else if (mode == 5) scatter_probe2_kernel<5><<<blocks, 256, 0, st>>>(b, nrow, iters);
  else if (mode == 6) scatter_probe2_kernel<6><<<blocks, 256, 0, st>>>(b, nrow, iters);
  else return MSDA_ERR_UNSUPPORTED;
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

extern "C" int msda_b200_probe_gather(const void* buf, long long bytes, int seg_bytes, int iters, int blocks,
                                      void* sink, void* stream) {
  if (!buf || !sink || bytes < seg_bytes || iters <= 0 || blocks <= 0) return MSDA_ERR_BAD_SHAPE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint32_t nseg = static_cast<uint32_t>(bytes / seg_bytes);
  iters = (iters + 7) / 8 * 8;
  if (seg_bytes == 128)
    gather_probe_kernel<8><<<blocks, 256, 0, st>>>(static_cast<const uint4*>(buf), nseg, iters, static_cast<uint32_t*>(sink));
  else if (seg_bytes == 64)
    gather_probe_kernel<4><<<blocks, 256, 0, st>>>(static_cast<const uint4*>(buf), nseg, iters, static_cast<uint32_t*>(sink));
  else if (seg_bytes == 512)
    gather_probe_kernel<32><<<blocks, 256, 0, st>>>(static_cast<const uint4*>(buf), nseg, iters, static_cast<uint32_t*>(sink));
  else
    return MSDA_ERR_UNSUPPORTED;
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}
