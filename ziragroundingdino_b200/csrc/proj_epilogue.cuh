// Epilogue parameter block and per-row epilogue math shared by the tcgen05 projection GEMMs (proj_gemm.cu: 16-bit
// operands; proj_gemm_f32.cu: fp32 operands as 3 x TF32).  The thread that runs these owns ONE accumulator row and a
// chunk of 32 consecutive columns (tcgen05.ld 32x32b.x32).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/msda_b200.h"
#include "tc_common.cuh"

namespace pg {

enum EpiMode { EPI_STORE = 0, EPI_QUERY = 1, EPI_ZIRA = 2 };

struct EpiParams {
  int mode;
  // EPI_STORE: out[row * out_ld + col] = acc + bias[col]  (row zeroed when row_mask[row] != 0)
  void* out;
  int out_ld;
  int out_f32;             // 1: fp32 output, 0: same 16-bit type as the inputs
  int out_half;            // 16-bit output is IEEE half instead of bf16
  const float* bias;       // [Nout] fp32, may be null
  const uint8_t* row_mask; // [R] or null
  int relu;                // max(., 0) after the bias
  const void* gate;        // optional 16-bit [R, out_ld]: out = acc where gate > 0 else 0 (ReLU backward fused in a dgrad)
  // 1-bit form of the same gate: word (c / 32) * R + row holds bits (column c + j > 0), j = 0..31 -- [Nout/32, R]
  // uint32, word-major so that the 32 lanes (rows) of a warp read / write 128 contiguous bytes.
  const void* accum;         // optional 16-bit [R, out_ld] added to the product before the store (may alias `out`: a tile is
                             // read and then written by the same warp) -- gradient accumulation without a separate add pass
  uint32_t* relu_bits;       // written by a RELU launch (may be null)
  const uint32_t* gate_bits; // read by a GATE launch instead of `gate` (may be null)
  // EPI_QUERY: columns [0, n_loc) are sampling offsets laid out (m, l, p, xy); columns [n_loc, n_loc + n_aw)
  // are attention logits laid out (m, l*p).
  float* loc_out;          // [R, n_loc]
  float* aw_out;           // [R, n_aw]
  const float* ref;        // [R, L, ref_dim]
  const int64_t* shapes;   // device [L, 2] (H, W)
  int ref_dim, L, P, n_loc, n_aw;
  // EPI_ZIRA (training-mode ZiRa projection): the stacked weight is [W_0; W_f; W_b] interleaved in runs of 32
  // output features, so columns [96g, 96g+32) / [+32, +64) / [+64, +96) are the base / soft-frozen / branch
  // products of features [32g, 32g+32).  bias = [b_0 | b_f | b_b] (3F floats).
  //   branch  = s * (x W_b^T + b_b)            adapter = branch + x W_f^T + b_f           y = x W_0^T + b_0 + adapter
  //   loss_sums[0] += sum SmoothL1(branch), loss_sums[1] += sum SmoothL1(adapter)   (all rows, also masked ones)
  // LN (see linear_tc_kernel): LayerNorm of the stored sum in the same launch
  const float* ln_gamma;   // [Nout] fp32, null = off
  const float* ln_beta;    // [Nout] fp32
  float ln_eps;
  void* ln_y;              // [R, out_ld] 16-bit: the normalised output
  float* ln_mean;          // [R]
  float* ln_rstd;          // [R]
  const float* scaling;    // device scalar s
  void* pre_out;           // [R, F] 16-bit: x W_b^T + b_b   (saved for backward; may be null)
  void* adapter_out;       // [R, F] 16-bit: adapter          (saved for backward; may be null)
  float* loss_sums;        // 2 floats, pre-zeroed by the caller
  int F;
};

__device__ __forceinline__ uint32_t pack16(float a, float b, bool half_out) {
  if (half_out) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// max(x, 0) fused into the conversion (cvt.rn.relu): first argument in the low half, like pack16
__device__ __forceinline__ uint32_t pack16_relu(float a, float b, bool half_out) {
  uint32_t d;
  if (half_out) asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}
// bit jx of the result = (x[jx] > 0), two instructions per element: the sign of -bits(x) is set exactly for the positive
// floats (bits in [1, 0x7fffffff]); a funnel shift moves it into the mask.  (-0.0f, bits 0x80000000, would count as
// positive; a pre-activation is +0 on exact cancellation, never -0, in round-to-nearest.)
__device__ __forceinline__ uint32_t positive_mask32(const float (&x)[32]) {
  uint32_t m = 0u;
#pragma unroll
  for (int jx = 31; jx >= 0; --jx) m = __funnelshift_l(0u - __float_as_uint(x[jx]), m, 1);
  return m;
}

// ---- epilogue -------------------------------------------------------------------------------------
// tcgen05.ld hands every thread ONE accumulator row (32 consecutive columns per chunk).  Measured in round 1
// (profiles/r1_gemm_ab.txt): the kernel is bound by the latency of this per-chunk instruction chain on the few epilogue
// warps, not by the store pattern (a shared-memory transposition to coalesce the stores changed nothing) and not by
// shared-memory traffic (resident vs streamed W changed nothing).  Hence: EIGHT epilogue warps (two per TMEM lane
// quarter, alternating column chunks) and every mode / dtype decision a template parameter.
// Warp-private transposition buffer (32 rows x N16*16 bytes, +16 bytes pitch against bank conflicts): registers ->
// shared (row per lane), __syncwarp, shared -> global with lanes running along the row, so a store instruction
// covers 32/N16 rows x N16*16 contiguous bytes instead of 32 rows x 16 bytes.  With eight epilogue warps the LSU
// wavefront count of the direct pattern (32 per instruction) is what bounds the kernel.
template <int N16>
__device__ __forceinline__ void staged_store(uint8_t* buf, int lane, const uint4 (&regs)[N16], void* gdst_row0,
                                             long long ld_bytes, int rows_valid) {
  constexpr int PITCH = N16 * 16 + 16, RPI = 32 / N16;
#pragma unroll
  for (int i = 0; i < N16; ++i) *reinterpret_cast<uint4*>(buf + lane * PITCH + i * 16) = regs[i];
  __syncwarp();
  const int c16 = lane % N16, rsub = lane / N16;
  uint8_t* g = static_cast<uint8_t*>(gdst_row0);
#pragma unroll
  for (int it = 0; it < N16; ++it) {
    const int r = it * RPI + rsub;
    const uint4 v = *reinterpret_cast<const uint4*>(buf + r * PITCH + c16 * 16);
    if (r < rows_valid) *reinterpret_cast<uint4*>(g + r * ld_bytes + c16 * 16) = v;
  }
  __syncwarp();
}

__device__ __forceinline__ void pack_f32(const float (&v)[32], uint4 (&o)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
    o[i] = make_uint4(__float_as_uint(v[4 * i]), __float_as_uint(v[4 * i + 1]), __float_as_uint(v[4 * i + 2]), __float_as_uint(v[4 * i + 3]));
}
__device__ __forceinline__ void pack_16(const float (&v)[32], bool half_out, bool zero, uint4 (&o)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    o[i].x = zero ? 0u : pack16(v[8 * i], v[8 * i + 1], half_out);
    o[i].y = zero ? 0u : pack16(v[8 * i + 2], v[8 * i + 3], half_out);
    o[i].z = zero ? 0u : pack16(v[8 * i + 4], v[8 * i + 5], half_out);
    o[i].w = zero ? 0u : pack16(v[8 * i + 6], v[8 * i + 7], half_out);
  }
}

__device__ __forceinline__ void pack_16_relu(const float (&v)[32], bool half_out, uint4 (&o)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    o[i].x = pack16_relu(v[8 * i], v[8 * i + 1], half_out);
    o[i].y = pack16_relu(v[8 * i + 2], v[8 * i + 3], half_out);
    o[i].z = pack16_relu(v[8 * i + 4], v[8 * i + 5], half_out);
    o[i].w = pack16_relu(v[8 * i + 6], v[8 * i + 7], half_out);
  }
}

__device__ __forceinline__ void unpack16x2(uint32_t w, bool half_in, float& a, float& b) {
  if (half_in) {
    const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w));
    a = t.x; b = t.y;
  } else {
    a = __uint_as_float(w << 16); b = __uint_as_float(w & 0xffff0000u);
  }
}

// v[0..31] += 32 16-bit values packed in 16 words (bf16 or IEEE half)
__device__ __forceinline__ void add_packed16(float (&v)[32], const uint32_t* w, bool half_in) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (half_in) {
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
      v[2 * j] += t.x; v[2 * j + 1] += t.y;
    } else {
      v[2 * j] += __uint_as_float(w[j] << 16); v[2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u);
    }
  }
}

// sampling locations (ms_deform_attn.py:306-319): columns are (m, l, p, xy).  Generic shapes.
__device__ __forceinline__ void epi_loc(const EpiParams& ep, const float (&v)[32], long long row, int gc,
                                        const float* s_inv, float (&r)[32]) {
  const float* rp = ep.ref + row * ep.L * ep.ref_dim;
  const float half_over_p = 0.5f / static_cast<float>(ep.P);
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int idx = gc + j, xy = idx & 1, l = (idx / (2 * ep.P)) % ep.L;
    if (ep.ref_dim == 2) r[j] = fmaf(v[j], s_inv[l * 2 + xy], rp[l * 2 + xy]);
    else r[j] = fmaf(v[j] * half_over_p, rp[l * 4 + 2 + xy], rp[l * 4 + xy]);
  }
}

// L = 4, P = 4 (every GroundingDINO configuration): one 32-column chunk is exactly one head, so the level
// and the x/y selector of each column are compile-time and the row's reference points sit in registers.
template <int REF_DIM>
__device__ __forceinline__ void epi_loc_l4p4(const EpiParams& ep, const float (&v)[32], long long row,
                                             const float* s_inv, float (&r)[32]) {
  constexpr int L = 4, P = 4;
  float ref[L * REF_DIM];
  const float4* rp = reinterpret_cast<const float4*>(ep.ref + row * L * REF_DIM);
#pragma unroll
  for (int i = 0; i < L * REF_DIM / 4; ++i) {
    const float4 t = __ldg(rp + i);
    ref[4 * i] = t.x; ref[4 * i + 1] = t.y; ref[4 * i + 2] = t.z; ref[4 * i + 3] = t.w;
  }
  float inv[2 * L];
#pragma unroll
  for (int i = 0; i < 2 * L; ++i) inv[i] = s_inv[i];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    constexpr float half_over_p = 0.5f / P;
    const int l = j / (2 * P), xy = j & 1;
    if (REF_DIM == 2) r[j] = fmaf(v[j], inv[l * 2 + xy], ref[l * 2 + xy]);
    else r[j] = fmaf(v[j] * half_over_p, ref[l * 4 + 2 + xy], ref[l * 4 + xy]);
  }
}

// softmax over each run of L*P logits (ms_deform_attn.py:293-303); 32 % (L*P) == 0 is guaranteed by the host
template <int LP>
__device__ __forceinline__ void softmax_runs(float (&v)[32]) {
#pragma unroll
  for (int g0 = 0; g0 < 32; g0 += LP) {
    float mx = v[g0];
#pragma unroll
    for (int j = 1; j < LP; ++j) mx = fmaxf(mx, v[g0 + j]);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < LP; ++j) { v[g0 + j] = __expf(v[g0 + j] - mx); sum += v[g0 + j]; }
    const float inv = __fdividef(1.f, sum);
#pragma unroll
    for (int j = 0; j < LP; ++j) v[g0 + j] *= inv;
  }
}

__device__ __forceinline__ void epi_softmax(const EpiParams& ep, float (&v)[32]) {
  const int lp = ep.L * ep.P;
  if (32 % lp != 0) return;   // runs straddle chunks (e.g. L*P = 20): logits are stored raw, softmax_rows_kernel finishes
  if (lp == 16) softmax_runs<16>(v);
  else if (lp == 32) softmax_runs<32>(v);
  else if (lp == 8) softmax_runs<8>(v);
  else if (lp == 4) softmax_runs<4>(v);
  else if (lp == 2) softmax_runs<2>(v);
  else softmax_runs<1>(v);
}

__device__ __forceinline__ float smooth_l1(float x) {   // beta = 1 (torch.nn.SmoothL1Loss default)
  const float a = fabsf(x);
  return a < 1.f ? 0.5f * x * x : a - 0.5f;
}

// ---- host side (defined in proj_gemm.cu) ----------------------------------------------------------
extern thread_local char t_err[256];
// row-major [rows, cols] matrix, box = box_rows x box_cols (box_cols * element size == 128 bytes), 128-byte swizzle.
// dtype: 0 bf16, 1 f16, 2 f32.
int make_map(CUtensorMap* map, const void* base, long long rows, long long cols, int box_rows, int box_cols, int dtype);
// [batch, rows, cols] tensor of 16-bit elements (row stride = cols), box = 1 x box_rows x 64 columns, 128-byte swizzle;
// rows beyond `rows` read as zeros and are clipped on store (per batch entry -- the reason for the third dimension).
int make_map3(CUtensorMap* map, const void* base, long long batch, long long rows, long long cols, int box_rows, int dtype);
// [d3, d2, rows, cols] dense tensor of 16-bit elements, box = 1 x 1 x box_rows x 64 columns, 128-byte swizzle
int make_map4(CUtensorMap* map, const void* base, long long d3, long long d2, long long rows, long long cols, int box_rows, int dtype);
// in-place softmax over runs of `lp` consecutive fp32 values (L*P does not divide the 32-column epilogue chunk)
int softmax_rows(float* x, long long runs, int lp, cudaStream_t st);

}  // namespace pg
