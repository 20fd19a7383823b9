// Elementwise companions of the projection GEMMs (backward side of ms_deform_attn.py:286-325).
//
//  * query_bwd_prep: turns the core op's grad_sampling_loc / grad_attn_weight into the gradient of the
//    stacked [sampling_offsets | attention_weights] pre-activations, in the 16-bit type the dgrad GEMM
//    consumes: d_off = grad_loc / (W_l, H_l)            (2-d reference points, :306-311)
//              d_off = grad_loc * ref_wh * 0.5 / P      (4-d reference boxes,  :312-319)
//              d_logit = aw * (grad_aw - sum_j grad_aw_j * aw_j)   over each run of L*P (softmax, :296)
//  * cast_mask: fp32 grad_value -> 16-bit with padded rows zeroed (backward of masked_fill, :287-288).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "msda_common.cuh"

namespace {

__device__ __forceinline__ uint16_t to16(float v, bool is_half) {
  if (is_half) { __half h = __float2half_rn(v); return *reinterpret_cast<uint16_t*>(&h); }
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  return *reinterpret_cast<uint16_t*>(&h);
}

// One thread per 4 consecutive columns of the stacked [R, 3*M*L*P] output row: coalesced 16-byte loads,
// 8-byte stores.  Offset columns are a per-column scale; logit columns need the softmax dot product of
// their run of L*P weights, taken with shuffles across the L*P/4 neighbouring threads.  Needs (L*P) % 4 == 0.
__global__ void __launch_bounds__(256)
query_bwd_prep_kernel(const float* __restrict__ grad_loc, const float* __restrict__ grad_aw, const float* __restrict__ aw,
                      const float* __restrict__ ref, const int64_t* __restrict__ shapes, long long R, int M, int L, int P,
                      int ref_dim, int is_half, uint16_t* __restrict__ out, int ld_out) {
  __shared__ float s_inv[MSDA_MAX_LEVELS * 2];
  if (threadIdx.x < L) {
    s_inv[2 * threadIdx.x] = 1.f / static_cast<float>(shapes[2 * threadIdx.x + 1]);
    s_inv[2 * threadIdx.x + 1] = 1.f / static_cast<float>(shapes[2 * threadIdx.x]);
  }
  __syncthreads();
  // ld_out >= 3*M*L*P: the row may be zero-padded to a multiple of 64 so that it can be the K dimension of the dgrad GEMM
  const int lp = L * P, n_aw = M * lp, n_loc = 2 * n_aw, ld = ld_out, tpr = ld / 4;   // threads per row
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool live = idx < R * tpr;
  const long long row = live ? idx / tpr : 0;
  const int col = live ? static_cast<int>(idx % tpr) * 4 : 0;
  float o[4] = {0.f, 0.f, 0.f, 0.f};
  const bool is_logit = col >= n_loc && col < n_loc + n_aw;
  const bool is_pad = col >= n_loc + n_aw;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), g = a;
  if (live && is_logit) {
    a = __ldg(reinterpret_cast<const float4*>(aw + row * n_aw + (col - n_loc)));
    g = __ldg(reinterpret_cast<const float4*>(grad_aw + row * n_aw + (col - n_loc)));
  }
  // softmax backward: runs of lp logits = lp/4 consecutive threads.  Power-of-two groups reduce with shuffles (a run
  // never straddles a warp); other run lengths (L*P = 20) re-read the run -- 2*lp cached loads per thread.
  float dot = g.x * a.x + g.y * a.y + g.z * a.z + g.w * a.w;
  const int tpg = lp / 4;
  if ((tpg & (tpg - 1)) == 0) {
    for (int o_ = 1; o_ < tpg; o_ <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o_);
  } else if (live && is_logit) {
    const int run0 = ((col - n_loc) / lp) * lp;
    const float* ga = grad_aw + row * n_aw + run0;
    const float* aa = aw + row * n_aw + run0;
    dot = 0.f;
    for (int j = 0; j < lp; ++j) dot = fmaf(__ldg(ga + j), __ldg(aa + j), dot);
  }
  if (!live) return;
  if (is_pad) {
    // zero padding
  } else if (is_logit) {
    o[0] = a.x * (g.x - dot); o[1] = a.y * (g.y - dot); o[2] = a.z * (g.z - dot); o[3] = a.w * (g.w - dot);
  } else {
    const float4 gl = __ldg(reinterpret_cast<const float4*>(grad_loc + row * n_loc + col));
    const float gv[4] = {gl.x, gl.y, gl.z, gl.w};
    const float* rp = ref + row * L * ref_dim;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = col + j, xy = c & 1, l = (c / (2 * P)) % L;
      const float sc = (ref_dim == 2) ? s_inv[2 * l + xy] : rp[l * 4 + 2 + xy] * (0.5f / P);
      o[j] = gv[j] * sc;
    }
  }
  if (is_half == 2) {   // fp32 output (fp32 modules)
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + row * ld + col) = make_float4(o[0], o[1], o[2], o[3]);
    return;
  }
  uint2 w;
  w.x = to16(o[0], is_half != 0) | (static_cast<uint32_t>(to16(o[1], is_half != 0)) << 16);
  w.y = to16(o[2], is_half != 0) | (static_cast<uint32_t>(to16(o[3], is_half != 0)) << 16);
  *reinterpret_cast<uint2*>(out + row * ld + col) = w;
}

// ---- fp32 modules: the elementwise tail of the query projections as two kernels instead of ~10 eager ops -------
// raw [R, 3*M*L*P] = [sampling_offsets | attention logits] pre-activations (bias included), straight from the GEMM.
//   loc = ref + raw_off / (W_l, H_l)                    (2-d reference points, ms_deform_attn.py:306-311)
//   loc = ref_xy + raw_off / P * ref_wh * 0.5           (4-d reference boxes,  :312-319)     -- same operation order
//   aw  = softmax over each run of L*P logits           (:296)
__global__ void __launch_bounds__(256)
query_post_loc_kernel(const float* __restrict__ raw, const float* __restrict__ ref, const int64_t* __restrict__ shapes,
                      long long R, int M, int L, int P, int ref_dim, float* __restrict__ loc) {
  __shared__ float s_norm[MSDA_MAX_LEVELS * 2];
  if (threadIdx.x < L) {
    s_norm[2 * threadIdx.x] = static_cast<float>(shapes[2 * threadIdx.x + 1]);
    s_norm[2 * threadIdx.x + 1] = static_cast<float>(shapes[2 * threadIdx.x]);
  }
  __syncthreads();
  const int n_loc = 2 * M * L * P, ld = 3 * M * L * P, tpr = n_loc / 4;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= R * tpr) return;
  const long long row = idx / tpr;
  const int col = static_cast<int>(idx % tpr) * 4;
  const float4 v = __ldg(reinterpret_cast<const float4*>(raw + row * ld + col));
  const float in[4] = {v.x, v.y, v.z, v.w};
  const float* rp = ref + row * L * ref_dim;
  float o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = col + j, xy = c & 1, l = (c / (2 * P)) % L;
    if (ref_dim == 2) o[j] = rp[l * 2 + xy] + __fdiv_rn(in[j], s_norm[2 * l + xy]);
    else o[j] = rp[l * 4 + xy] + __fdiv_rn(in[j], static_cast<float>(P)) * rp[l * 4 + 2 + xy] * 0.5f;
  }
  *reinterpret_cast<float4*>(loc + row * n_loc + col) = make_float4(o[0], o[1], o[2], o[3]);
}

// one thread per run of lp logits (adjacent threads own adjacent runs: coalesced across the warp)
__global__ void __launch_bounds__(256)
query_post_softmax_kernel(const float* __restrict__ raw, long long R, int M, int lp, float* __restrict__ aw) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= R * M) return;
  const long long row = i / M;
  const int m = static_cast<int>(i % M);
  const float* p = raw + row * (3ll * M * lp) + 2 * M * lp + m * lp;
  float* o = aw + i * lp;
  float mx = p[0];
  for (int j = 1; j < lp; ++j) mx = fmaxf(mx, p[j]);
  float sum = 0.f;
  for (int j = 0; j < lp; ++j) sum += expf(p[j] - mx);
  for (int j = 0; j < lp; ++j) o[j] = expf(p[j] - mx) / sum;
}

// 8 elements per thread; `cols` (row length) must be a multiple of 8
__global__ void __launch_bounds__(256)
cast_mask_kernel(const float* __restrict__ in, const uint8_t* __restrict__ row_mask, long long n8, int cols8, int is_half,
                 uint4* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const bool zero = row_mask != nullptr && row_mask[i / cols8] != 0;
  const float4 a = __ldg(reinterpret_cast<const float4*>(in) + 2 * i), b = __ldg(reinterpret_cast<const float4*>(in) + 2 * i + 1);
  uint4 w;
  const bool h = is_half != 0;
  w.x = to16(a.x, h) | (static_cast<uint32_t>(to16(a.y, h)) << 16);
  w.y = to16(a.z, h) | (static_cast<uint32_t>(to16(a.w, h)) << 16);
  w.z = to16(b.x, h) | (static_cast<uint32_t>(to16(b.y, h)) << 16);
  w.w = to16(b.z, h) | (static_cast<uint32_t>(to16(b.w, h)) << 16);
  out[i] = zero ? make_uint4(0u, 0u, 0u, 0u) : w;
}
// Consumer of the scaled-fp16 grad_value map (msda_backward_fusedq_h16): sum of a pixel's replicas / scale -> 16-bit storage
// (padded rows zeroed) or fp32.  Scale and replica layout are recomputed from the amax word behind the map and the
// device-side shapes, exactly as the scatter kernel did.  One thread per 8 channels.
template <bool OUT_F32>
__global__ void __launch_bounds__(256)
cast_mask_h16_kernel(const uint4* __restrict__ in, const uint32_t* __restrict__ amax, const int64_t* __restrict__ shapes,
                     const int64_t* __restrict__ lstart, int L, int S, long long rows_h, int Lq,
                     const uint8_t* __restrict__ row_mask, long long n8, int cols8, int is_half, void* __restrict__ out) {
  __shared__ int sStart[MSDA_MAX_LEVELS], sHStart[MSDA_MAX_LEVELS], sHW[MSDA_MAX_LEVELS], sRep[MSDA_MAX_LEVELS];
  if (threadIdx.x == 0) {
    long long at = 0;
    for (int l = 0; l < L; ++l) {
      const int hw = static_cast<int>(shapes[2 * l] * shapes[2 * l + 1]), k = msda::f16acc_replicas(Lq, hw);
      sStart[l] = static_cast<int>(lstart[l]); sHStart[l] = static_cast<int>(at); sHW[l] = hw; sRep[l] = k;
      at += static_cast<long long>(k) * hw;
    }
  }
  __syncthreads();
  const float inv = 1.f / msda::f16acc_scale(__ldg(amax), Lq);        // a power of two (NaN if the scatter kernel refused the layout)
  const long long b = blockIdx.y;
  const unsigned total = static_cast<unsigned>(S) * cols8;           // image blockIdx.y: 8-channel groups, grid-stride
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int s = static_cast<int>(idx / static_cast<unsigned>(cols8));
    const int c8 = static_cast<int>(idx - static_cast<unsigned>(s) * cols8);
    const long long r = b * S + s;
    const long long i = r * cols8 + c8;
    int l = 0;
    while (l + 1 < L && s >= sStart[l + 1]) ++l;
    const uint4* src = in + ((b * rows_h + sHStart[l] + (s - sStart[l])) * cols8 + c8);
    const long long rep_stride = static_cast<long long>(sHW[l]) * cols8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int k = 0; k < sRep[l]; ++k) {
      const uint4 v = __ldg(src + k * rep_stride);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
        acc[2 * j] += f.x; acc[2 * j + 1] += f.y;
      }
    }
    const bool zero = row_mask != nullptr && row_mask[r] != 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = zero ? 0.f : acc[j] * inv;
    if constexpr (OUT_F32) {
      float4* o = static_cast<float4*>(out) + 2 * i;
      o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    } else {
      const bool h = is_half != 0;
      uint4 w;
      w.x = to16(acc[0], h) | (static_cast<uint32_t>(to16(acc[1], h)) << 16);
      w.y = to16(acc[2], h) | (static_cast<uint32_t>(to16(acc[3], h)) << 16);
      w.z = to16(acc[4], h) | (static_cast<uint32_t>(to16(acc[5], h)) << 16);
      w.w = to16(acc[6], h) | (static_cast<uint32_t>(to16(acc[7], h)) << 16);
      static_cast<uint4*>(out)[i] = w;
    }
  }
}
__device__ __forceinline__ float from16(uint16_t v, bool is_half) {
  if (is_half) return __half2float(*reinterpret_cast<const __half*>(&v));
  return __uint_as_float(static_cast<uint32_t>(v) << 16);
}

// ZiRa training-mode projection, backward prep (one pass): from the upstream gradient dY, the saved branch
// pre-activation and adapter output, and the upstream gradient of the zero-inter loss, build the K-stacked
// operand [dY_eff | dO | dB] of the dgrad GEMM:
//   dY_eff = dY (0 on masked rows);  dO = dY_eff + gl * SmoothL1'(adapter);  dB = dO + gl * SmoothL1'(s * pre)
// where gl = dLoss / (R * F) and SmoothL1'(x) = clamp(x, -1, 1).  8 features per thread.
__global__ void __launch_bounds__(256)
zira_bwd_prep_kernel(const uint16_t* __restrict__ dy, const uint16_t* __restrict__ pre, const uint16_t* __restrict__ adapter,
                     const uint8_t* __restrict__ row_mask, const float* __restrict__ scaling, const float* __restrict__ dloss,
                     long long R, int F, int is_half, uint16_t* __restrict__ out, float* __restrict__ ds_out,
                     float* __restrict__ colsum_out) {
  // Grid-stride over (row, 8-feature group).  The host only asks for column sums when 256 % (F/8) == 0, so a thread
  // keeps the same feature group on every trip and its 3 x 8 partial column sums stay in registers.
  const int f8 = F / 8;
  const bool h = is_half != 0;
  const float s = __ldg(scaling);
  const float gl = __ldg(dloss) / (static_cast<float>(R) * static_cast<float>(F));
  float ds = 0.f;   // this thread's share of d(loss)/d(scaling) = sum dB * pre, taken before dB is rounded to 16 bit
  float cs[3][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) cs[0][j] = cs[1][j] = cs[2][j] = 0.f;
  const long long total = R * f8, stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long row = i / f8;
    const int col = static_cast<int>(i % f8) * 8;
    const bool masked = row_mask != nullptr && row_mask[row] != 0;
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(dy + row * F + col));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(pre + row * F + col));
    const uint4 c = __ldg(reinterpret_cast<const uint4*>(adapter + row * F + col));
    const uint16_t* pa = reinterpret_cast<const uint16_t*>(&a);
    const uint16_t* pb = reinterpret_cast<const uint16_t*>(&b);
    const uint16_t* pc = reinterpret_cast<const uint16_t*>(&c);
    uint16_t o0[8], o1[8], o2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float dye = masked ? 0.f : from16(pa[j], h);
      const float d_o = dye + gl * fminf(fmaxf(from16(pc[j], h), -1.f), 1.f);
      const float d_b = d_o + gl * fminf(fmaxf(s * from16(pb[j], h), -1.f), 1.f);
      o0[j] = to16(dye, h); o1[j] = to16(d_o, h); o2[j] = to16(d_b, h);
      ds = fmaf(d_b, from16(pb[j], h), ds);
      cs[0][j] += dye; cs[1][j] += d_o; cs[2][j] += d_b;
    }
    uint16_t* orow = out + row * 3 * F + col;
    *reinterpret_cast<uint4*>(orow) = *reinterpret_cast<const uint4*>(o0);
    *reinterpret_cast<uint4*>(orow + F) = *reinterpret_cast<const uint4*>(o1);
    *reinterpret_cast<uint4*>(orow + 2 * F) = *reinterpret_cast<const uint4*>(o2);
  }
  if (colsum_out != nullptr) {   // bias gradients: column sums of dY_eff | dO | dB, combined per CTA in shared memory
    extern __shared__ float s_cols[];   // 3 * F floats
    for (int k = threadIdx.x; k < 3 * F; k += blockDim.x) s_cols[k] = 0.f;
    __syncthreads();
    const int col = static_cast<int>((static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) % f8) * 8;
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&s_cols[q * F + col + j], cs[q][j]);
    __syncthreads();
    for (int k = threadIdx.x; k < 3 * F; k += blockDim.x) atomicAdd(colsum_out + k, s_cols[k]);
  }
  if (ds_out != nullptr) {   // block reduction, one atomic per CTA
    __shared__ float s_part[8];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) ds += __shfl_xor_sync(0xffffffffu, ds, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = ds;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += s_part[w];
      atomicAdd(ds_out, t);
    }
  }
}
}  // namespace

extern "C" {

int msda_zira_bwd_prep_16(const void* dy, const void* pre, const void* adapter, const uint8_t* row_mask, const float* scaling,
                          const float* dloss, long long R, int F, void* out, float* ds_out, float* colsum_out, int is_half,
                          void* stream) {
  if (!dy || !pre || !adapter || !scaling || !dloss || !out) return MSDA_ERR_NULL_POINTER;
  if (R <= 0 || F <= 0 || F % 8) return MSDA_ERR_BAD_SHAPE;
  if (colsum_out != nullptr && (256 % (F / 8) != 0 || 3 * F * sizeof(float) > 48 * 1024)) return MSDA_ERR_UNSUPPORTED;
  const long long n = R * (F / 8);
  const long long want = (n + 255) / 256;
  // with column sums: a few CTAs per SM so the per-CTA flush (3F atomics) stays negligible
  const unsigned blocks = static_cast<unsigned>(colsum_out != nullptr ? (want < 148 * 8 ? want : 148 * 8) : want);
  ++msda::g_launches;
  zira_bwd_prep_kernel<<<blocks, 256, colsum_out != nullptr ? 3 * F * sizeof(float) : 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint16_t*>(dy), static_cast<const uint16_t*>(pre), static_cast<const uint16_t*>(adapter), row_mask,
      scaling, dloss, R, F, is_half, static_cast<uint16_t*>(out), ds_out, colsum_out);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}


int msda_query_post_f32(const float* raw, const float* ref, int ref_dim, const int64_t* spatial_shapes, long long R, int M, int L,
                        int P, float* loc_out, float* aw_out, void* stream) {
  if (!raw || !ref || !spatial_shapes || !loc_out || !aw_out) return MSDA_ERR_NULL_POINTER;
  if (R <= 0 || M <= 0 || L <= 0 || L > MSDA_MAX_LEVELS || P <= 0 || (ref_dim != 2 && ref_dim != 4) || (2 * M * L * P) % 4)
    return MSDA_ERR_BAD_SHAPE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n1 = R * (2ll * M * L * P / 4), n2 = R * M;
  msda::g_launches += 2;
  query_post_loc_kernel<<<static_cast<unsigned>((n1 + 255) / 256), 256, 0, st>>>(raw, ref, spatial_shapes, R, M, L, P, ref_dim, loc_out);
  query_post_softmax_kernel<<<static_cast<unsigned>((n2 + 255) / 256), 256, 0, st>>>(raw, R, M, L * P, aw_out);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

int msda_query_bwd_prep_16(const float* grad_loc, const float* grad_aw, const float* aw, const float* ref, int ref_dim,
                           const int64_t* spatial_shapes, long long R, int M, int L, int P, void* out, int ld_out, int is_half,
                           void* stream) {
  if (!grad_loc || !grad_aw || !aw || !ref || !spatial_shapes || !out) return MSDA_ERR_NULL_POINTER;
  if (R <= 0 || M <= 0 || L <= 0 || L > MSDA_MAX_LEVELS || P <= 0 || (ref_dim != 2 && ref_dim != 4)) return MSDA_ERR_BAD_SHAPE;
  if ((L * P) % 4 || (L * P) > 128 || ld_out < 3 * M * L * P || ld_out % 4) return MSDA_ERR_UNSUPPORTED;
  const int tpg = L * P / 4;
  if ((tpg & (tpg - 1)) == 0 && 32 % tpg) return MSDA_ERR_UNSUPPORTED;
  const long long n = R * (ld_out / 4);
  ++msda::g_launches;
  query_bwd_prep_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      grad_loc, grad_aw, aw, ref, spatial_shapes, R, M, L, P, ref_dim, is_half, static_cast<uint16_t*>(out), ld_out);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

int msda_cast_mask_16(const float* in, const uint8_t* row_mask, long long rows, int cols, void* out, int is_half,
                      void* stream) {
  if (!in || !out) return MSDA_ERR_NULL_POINTER;
  if (rows <= 0 || cols <= 0 || cols % 8) return MSDA_ERR_BAD_SHAPE;
  const long long n8 = rows * (cols / 8);
  ++msda::g_launches;
  cast_mask_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, row_mask, n8, cols / 8, is_half, static_cast<uint4*>(out));
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

int msda_cast_mask_h16(const void* gv_h, const int64_t* shapes, const int64_t* lstart, int L, int N, int S, int cols, int Lq,
                       long long rows_h, const uint8_t* row_mask, void* out, int out_f32, int is_half, void* stream) {
  if (!gv_h || !out || !shapes || !lstart) return MSDA_ERR_NULL_POINTER;
  if (N <= 0 || N > 65535 || S <= 0 || cols <= 0 || cols % 8 || Lq <= 0 || L <= 0 || L > MSDA_MAX_LEVELS || rows_h < S ||
      static_cast<long long>(S) * (cols / 8) >= (1ll << 31))
    return MSDA_ERR_BAD_SHAPE;
  const long long n8 = static_cast<long long>(N) * S * (cols / 8);
  const uint32_t* amax = reinterpret_cast<const uint32_t*>(static_cast<const char*>(gv_h) + 2ll * N * rows_h * cols);
  const long long per_image = (static_cast<long long>(S) * (cols / 8) + 255) / 256;
  const dim3 blocks(static_cast<unsigned>(std::max<long long>(1, std::min<long long>(per_image, (148 * 8 + N - 1) / N))), static_cast<unsigned>(N));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ++msda::g_launches;
  if (out_f32)
    cast_mask_h16_kernel<true><<<blocks, 256, 0, st>>>(static_cast<const uint4*>(gv_h), amax, shapes, lstart, L, S, rows_h, Lq, row_mask, n8,
                                                       cols / 8, 0, out);
  else
    cast_mask_h16_kernel<false><<<blocks, 256, 0, st>>>(static_cast<const uint4*>(gv_h), amax, shapes, lstart, L, S, rows_h, Lq, row_mask,
                                                        n8, cols / 8, is_half, out);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

}  // extern "C"
