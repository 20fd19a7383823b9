// Multi-scale deformable attention, forward (bilinear gather) and backward (scatter), for sm_100a.
//
// Replaces the reference kernels ms_deformable_im2col_gpu_kernel (ms_deform_im2col_cuda.cuh:237-299)
// and ms_deformable_col2im_gpu_kernel_shm_blocksize_aware_reduce_v1 (:301-403); the arithmetic per
// sample (coordinate convention, guards, bilinear weights, gradient formulas) follows :33-159 so
// results agree with the reference op to fp32 round-off.  The design does not:
//
//  * a group of D/CH lanes (CH = channels in one 16-byte load: 4 fp32 / 8 bf16) owns one
//    (image, query, head) unit; every corner row of that head is ONE 128-bit load per lane and the
//    attention-weighted sum lives in CH registers per lane -- 4x (fp32) / 8x (bf16) fewer load
//    instructions than the reference's thread-per-channel mapping, and the coordinate math is done
//    by D/CH lanes instead of D;
//  * spatial_shapes / level_start_index are staged once per CTA in shared memory (the reference
//    re-reads the int64 arrays from global memory per thread per level);
//  * sampling locations and attention weights of a level are fetched with three 128-bit loads;
//  * backward: grad_value goes out as 128-bit vector reductions (REDG.E.ADD.F32x4), the per-sample
//    channel sums for grad_sampling_loc / grad_attn_weight are reduced with a warp-shuffle
//    reduce-scatter (14 shuffles per 4 samples for D=32 fp32) instead of shared memory plus a serial
//    loop on thread 0, and every (q, m, l, p) slot of both is written exactly once, so neither
//    needs a memset.
#include <type_traits>

#include "msda_common.cuh"

namespace msda {

// =================================================================================================
// forward, vectorised: D in {16, 32, 64, 128}, P == 4
// =================================================================================================
template <typename VT, int D, int SB, bool SHARE>
__global__ void __launch_bounds__(kThreads)
msda_fwd_vec_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                    const int64_t* __restrict__ lstart, const float* __restrict__ loc,
                    const float* __restrict__ aw, VT* __restrict__ out, int S, int M, int L, int Lq,
                    long long units, int passes, int q_fast) {
  constexpr int CH = Vec<VT>::CH;
  constexpr int LPG = D / CH;           // lanes per unit
  constexpr int UPW = 32 / LPG;         // units per warp
  constexpr int TILE = UPW * kWarpsPerBlock;
  constexpr int P = 4;
  static_assert(LPG >= 1 && LPG <= 32 && (LPG & (LPG - 1)) == 0, "D / CH must be a power of two <= 32");

  __shared__ int sH[MSDA_MAX_LEVELS], sW[MSDA_MAX_LEVELS], sStart[MSDA_MAX_LEVELS];
  if (threadIdx.x < L) {
    sH[threadIdx.x] = static_cast<int>(shapes[2 * threadIdx.x]);
    sW[threadIdx.x] = static_cast<int>(shapes[2 * threadIdx.x + 1]);
    sStart[threadIdx.x] = static_cast<int>(lstart[threadIdx.x]);
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lig = lane % LPG;                       // lane in group
  const int j = warp * UPW + lane / LPG;            // unit slot inside the pass tile
  const int row = M * D;                            // elements between neighbouring pixels

  constexpr bool SH = SHARE && LPG >= 4;   // taps computed once per lane group (shared_taps): every lane stays in the loop
  for (int it = 0; it < passes; ++it) {
    long long u = unit_of(static_cast<long long>(blockIdx.x) * passes + it, j, TILE, M, q_fast != 0);
    const bool active = u < units;
    if (!active) {
      if (!SH) continue;
      u = 0;
    }
    const int m = static_cast<int>(u % M);
    const long long bq = u / M;
    const long long b = bq / Lq;
    const VT* vb = value + (static_cast<size_t>(b) * S * M + m) * D + lig * CH;
    const float4* lp = reinterpret_cast<const float4*>(loc + static_cast<size_t>(u) * L * P * 2);
    const float4* ap = reinterpret_cast<const float4*>(aw + static_cast<size_t>(u) * L * P);

    float acc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = 0.f;

    for (int l = 0; l < L; ++l) {
      const int H = sH[l], W = sW[l];
      const int wrow = W * row;
      const VT* vl = vb + static_cast<size_t>(sStart[l]) * row;
      const float4 xy01 = __ldg(lp + 2 * l), xy23 = __ldg(lp + 2 * l + 1), a4 = __ldg(ap + l);
      const float xs[4] = {xy01.x, xy01.z, xy23.x, xy23.z};
      const float ys[4] = {xy01.y, xy01.w, xy23.y, xy23.w};
      const float as[4] = {a4.x, a4.y, a4.z, a4.w};
      Tap<float> t4[4];
      if constexpr (SH) shared_taps<LPG>(xs, ys, H, W, lig, t4);
#pragma unroll
      for (int p0 = 0; p0 < P; p0 += SB) {
        float v[SB][4][CH];
        float k[SB][4];
#pragma unroll
        for (int s = 0; s < SB; ++s) {
          Tap<float> t;
          if constexpr (SH) t = t4[p0 + s]; else t = make_tap<float>(xs[p0 + s], ys[p0 + s], H, W);
          // 32-bit element offsets (S*M*D < 2^31 is checked on the host): one IMAD + three adds per sample
          const int e1 = t.o1 * row, e3 = e1 + wrow;
          Vec<VT>::load(vl + e1, t.c1, v[s][0]);
          Vec<VT>::load(vl + e1 + row, t.c2, v[s][1]);
          Vec<VT>::load(vl + e3, t.c3, v[s][2]);
          Vec<VT>::load(vl + e3 + row, t.c4, v[s][3]);
          k[s][0] = t.hh * t.hw; k[s][1] = t.hh * t.lw; k[s][2] = t.lh * t.hw; k[s][3] = t.lh * t.lw;
        }
#pragma unroll
        for (int s = 0; s < SB; ++s) {
          // same association as the reference (im2col.cuh:80-82, :290): bilinear value first, then times the attention
          // weight.  Folding the weight into the four taps saves one FMA per channel but moves the result ~1e-5 away
          // from the reference op at sigma = 1 inputs (measured: profiles/r1_parity_table.txt), i.e. onto the parity bar.
          const float a = as[p0 + s];
          if constexpr (sizeof(VT) == 2 && CH % 2 == 0) {
            // 16-bit storage: the same multiply / fused multiply-add sequence as packed fp32x2 instructions (FMUL2 / FFMA2,
            // sm_100), two channels per issue slot in a kernel bound by issue slots (-9 % measured, profiles/r2_exp_pk2.txt).
            // The packed form is not bit-identical to the scalar one on the device, so the fp32 kernel -- whose forward is
            // bit-identical to the reference CUDA op -- keeps the scalar sequence below.
            const float2 k0 = make_float2(k[s][0], k[s][0]), k1 = make_float2(k[s][1], k[s][1]);
            const float2 k2 = make_float2(k[s][2], k[s][2]), k3 = make_float2(k[s][3], k[s][3]), a2 = make_float2(a, a);
#pragma unroll
            for (int c = 0; c < CH; c += 2) {
              float2 val = __fmul2_rn(k0, make_float2(v[s][0][c], v[s][0][c + 1]));
              val = __ffma2_rn(k1, make_float2(v[s][1][c], v[s][1][c + 1]), val);
              val = __ffma2_rn(k2, make_float2(v[s][2][c], v[s][2][c + 1]), val);
              val = __ffma2_rn(k3, make_float2(v[s][3][c], v[s][3][c + 1]), val);
              const float2 r = __ffma2_rn(val, a2, make_float2(acc[c], acc[c + 1]));
              acc[c] = r.x; acc[c + 1] = r.y;
            }
          } else {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
              const float val = k[s][0] * v[s][0][c] + k[s][1] * v[s][1][c] + k[s][2] * v[s][2][c] + k[s][3] * v[s][3][c];
              acc[c] = fmaf(val, a, acc[c]);
            }
          }
        }
      }
    }
    if (active) Vec<VT>::store(out + static_cast<size_t>(u) * D + lig * CH, acc);
  }
}

// =================================================================================================
// backward, vectorised: D in {16, 32, 64, 128} (LPG <= 16), P == 4
// =================================================================================================
// Sums v[0..15] across the LPG lanes of a group; afterwards lane `lig` holds entries
// [lig*R, lig*R + R) of the 16 sums in v[0..R), R = 16 / LPG.
template <int LPG>
__device__ __forceinline__ void reduce_scatter16(float (&v)[16], int lig) {
  int n = 16;
#pragma unroll
  for (int mask = LPG / 2; mask >= 1; mask >>= 1) {
    n >>= 1;
    const bool up = (lig & mask) != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < n) {
        const float send = up ? v[i] : v[i + n];
        const float keep = up ? v[i + n] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
      }
    }
  }
}

__device__ __forceinline__ float sel4(const float (&a)[4], int p) { return p == 0 ? a[0] : (p == 1 ? a[1] : (p == 2 ? a[2] : a[3])); }

// DOTS formulation of the per-sample sums.  The reference accumulates, per channel, d(bilinear)/dw * top_grad etc.
// (im2col.cuh:123-151); all three sums are linear in the four corner dot products d_i = <grad_out row, value row of
// corner i> (guarded-off corners contribute 0):
//     grad_aw = k1 d1 + k2 d2 + k3 d3 + k4 d4
//     grad_x  = W a [hh (d2 - d1) + lh (d4 - d3)]        grad_y = H a [hw (d3 - d1) + lw (d4 - d2)]
// so a lane only accumulates 4 FMAs per channel instead of 14 operations, the group reduce-scatters the 16 dots of a
// level (4 points x 4 corners: the slot that used to be padding now carries d4), and the lanes that end up holding a
// point's dots finish the three sums.  combine_dots() turns v[] from the dots layout into the (w, h, a, pad) layout the
// rest of the kernel expects.  Same mathematics, different association: fp32 round-off moves by ~1e-6 relative.
template <int LPG>
__device__ __forceinline__ void combine_dots(float (&v)[16], int lig, const float (&lh)[4], const float (&lw)[4],
                                             const float (&as)[4], float Wf, float Hf) {
  constexpr int R = 16 / LPG;
  auto finish = [&](int p, float d1, float d2, float d3, float d4, float& gx, float& gy, float& ga) {
    const float h_l = sel4(lh, p), w_l = sel4(lw, p), a = sel4(as, p);
    const float h_h = 1.f - h_l, w_h = 1.f - w_l;
    ga = (h_h * w_h) * d1 + (h_h * w_l) * d2 + (h_l * w_h) * d3 + (h_l * w_l) * d4;
    gx = Wf * a * (h_h * (d2 - d1) + h_l * (d4 - d3));
    gy = Hf * a * (w_h * (d3 - d1) + w_l * (d4 - d2));
  };
  if constexpr (R >= 4) {
#pragma unroll
    for (int r0 = 0; r0 < R; r0 += 4) {
      float gx, gy, ga;
      finish(((lig * R + r0) >> 2) & 3, v[r0], v[r0 + 1], v[r0 + 2], v[r0 + 3], gx, gy, ga);
      v[r0] = gx; v[r0 + 1] = gy; v[r0 + 2] = ga; v[r0 + 3] = 0.f;
    }
  } else if constexpr (R == 2) {
    const float o0 = __shfl_xor_sync(0xffffffffu, v[0], 1), o1 = __shfl_xor_sync(0xffffffffu, v[1], 1);
    const bool odd = (lig & 1) != 0;
    float gx, gy, ga;
    finish((lig >> 1) & 3, odd ? o0 : v[0], odd ? o1 : v[1], odd ? v[0] : o0, odd ? v[1] : o1, gx, gy, ga);
    v[0] = odd ? ga : gx;
    v[1] = odd ? 0.f : gy;
  } else {
    const int base = (threadIdx.x & 31) & ~3;
    const float d1 = __shfl_sync(0xffffffffu, v[0], base), d2 = __shfl_sync(0xffffffffu, v[0], base + 1);
    const float d3 = __shfl_sync(0xffffffffu, v[0], base + 2), d4 = __shfl_sync(0xffffffffu, v[0], base + 3);
    float gx, gy, ga;
    finish((lig >> 2) & 3, d1, d2, d3, d4, gx, gy, ga);
    const int comp = lig & 3;
    v[0] = comp == 0 ? gx : (comp == 1 ? gy : (comp == 2 ? ga : 0.f));
  }
}

// FUSEQ (L == 4, P == 4, 8 lanes per unit): instead of writing grad_sampling_loc / grad_attn_weight in fp32, finish the
// backward of the query-side epilogue in registers -- d_offset = grad_loc / (W_l, H_l) (2-d reference points) or
// grad_loc * ref_wh * 0.5 / P (4-d), d_logit = aw * (grad_aw - sum grad_aw * aw) -- and write the 16-bit operand
// [d_offsets | d_logits] of the query-projection dgrad GEMM directly (fq.dq, row stride 3*M*L*P).
struct FuseQ {
  const float* ref;     // [N*Lq, L, ref_dim]
  uint16_t* dq;         // [N*Lq, 3*M*L*P]
  int ref_dim;
  int is_half;
  uint32_t* amax;        // F16ACC: bits of max |grad_out| (written by grad_amax_kernel), else unused
  long long rows_h;      // F16ACC: rows per image of the scaled-fp16 map (levels with replicas), as the host sized it
  int h_start[4], h_hw[4], h_mask[4];   // F16ACC: per level, first row of its replica block, H*W, replicas - 1 (host-computed)
};

// 3 resident CTAs per SM (<= 85 registers): the kernel needs its occupancy to keep enough reductions and corner loads in
// flight.  The plain variant fits in 80 without spilling; the hit-mask bookkeeping (HITS) would take ~25
// more and spills 8-36 bytes instead, which is cheaper than dropping to 2 CTAs.
// MODE: 0 = every corner is a reduction (the round-1 kernel, unchanged); 1 = the coarse tail of the level list is owned by
// msda_scatter_mma_kernel; 2 = the last `mma_levels` levels are owned by msda_scatter_mma2_kernel and this kernel also
// writes the per-chunk hit masks it needs.
#ifndef MSDA_BWD_MINB
#define MSDA_BWD_MINB 3      // resident CTAs per SM the scatter kernel is compiled for (A/B: -DMSDA_BWD_MINB=4 -> 64 registers)
#endif
// F16ACC (16-bit storage, 4 channels per lane): grad_value is a SCALED fp16 map -- every contribution is multiplied by the
// power of two f16acc_scale(max |grad_out|, Lq) and leaves as red.global.add.noftz.v2.f16x2 (64 bytes per corner row
// instead of 128: the SM->L2 reduction path is byte-bound).  `grad_value` then points at __half storage.
template <typename VT, int D, typename V, bool FUSEQ, int MODE, bool SHARE, bool DOTS, bool F16ACC = false>
__global__ void __launch_bounds__(kThreads, MSDA_BWD_MINB)
msda_bwd_vec_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                    const int64_t* __restrict__ lstart, const float* __restrict__ loc,
                    const float* __restrict__ aw, const VT* __restrict__ grad_out,
                    float* __restrict__ grad_value, float* __restrict__ grad_loc,
                    float* __restrict__ grad_aw, int S, int M, int L, int Lq, long long units,
                    int passes, int q_fast, int mma_mode, int mma_levels, unsigned long long* __restrict__ hit, FuseQ fq) {
  constexpr int CH = V::CH;
  constexpr int LPG = D / CH;
  constexpr int UPW = 32 / LPG;
  constexpr int TILE = UPW * kWarpsPerBlock;
  constexpr int P = 4;
  constexpr int R = 16 / LPG;
  constexpr bool HITS = MODE == 2;
  static_assert(LPG >= 1 && LPG <= 16 && (LPG & (LPG - 1)) == 0, "D / CH must be a power of two <= 16");

  __shared__ int sH[MSDA_MAX_LEVELS], sW[MSDA_MAX_LEVELS], sStart[MSDA_MAX_LEVELS];
  __shared__ int sRedLevels;
  __shared__ float sInvW[MSDA_MAX_LEVELS], sInvH[MSDA_MAX_LEVELS];     // FUSEQ: 1/W_l, 1/H_l (IEEE division, once per CTA)
  __shared__ RangePlan plan;       // HITS (mma_mode 2) only; the compiler drops it otherwise
  __shared__ float sScale;
  static_assert(!F16ACC || (FUSEQ && MODE == 0 && V::CH == 4 && sizeof(VT) == 2), "F16ACC: fused-query 16-bit kernel only");
  __shared__ int sHStart[MSDA_MAX_LEVELS], sHW[MSDA_MAX_LEVELS], sRepMask[MSDA_MAX_LEVELS];   // F16ACC: replica blocks of the map
  if (F16ACC && threadIdx.x == 32) sScale = f16acc_scale(*fq.amax, Lq);
  if (threadIdx.x < L) {
    sH[threadIdx.x] = static_cast<int>(shapes[2 * threadIdx.x]);
    sW[threadIdx.x] = static_cast<int>(shapes[2 * threadIdx.x + 1]);
    sStart[threadIdx.x] = static_cast<int>(lstart[threadIdx.x]);
    if (FUSEQ) {
      sInvW[threadIdx.x] = 1.f / static_cast<float>(shapes[2 * threadIdx.x + 1]);
      sInvH[threadIdx.x] = 1.f / static_cast<float>(shapes[2 * threadIdx.x]);
    }
  }
  __syncthreads();
  // mma_mode 1: the grad_value contributions of the coarse tail (levels >= red_levels) are accumulated by
  // msda_scatter_mma_kernel; mma_mode 2: those of the last `mma_levels` levels by msda_scatter_mma2_kernel, which needs
  // to know which pixel ranges each 64-query chunk touches -- this kernel has every tap in hand and ORs the bits into `hit`
  if (threadIdx.x == 0) {
    if (HITS) { plan_ranges(plan, sH, sW, sStart, L, S, mma_levels); sRedLevels = plan.first_level; }
    else sRedLevels = mma_mode == 1 ? coarse_first_level(sH, sW, sStart, L, S) : L;
  }
  if (F16ACC && threadIdx.x < 4) {
    // the replica layout comes from the HOST's copy of the shapes (it sized and zeroed the map, and a serial plan in every
    // short-lived CTA costs more than it looks: the first version spent 0.9 warp-stalls per issue at this barrier)
    sHStart[threadIdx.x] = fq.h_start[threadIdx.x]; sHW[threadIdx.x] = fq.h_hw[threadIdx.x]; sRepMask[threadIdx.x] = fq.h_mask[threadIdx.x];
  }
  __syncthreads();
  if constexpr (F16ACC) {
    // refuse to write into a buffer laid out for other shapes than the device-side ones, and poison the amax word so that
    // the consumer produces NaN instead of silently wrong gradients
    if (sH[0] * sW[0] != sHW[0] || sH[1] * sW[1] != sHW[1] || sH[2] * sW[2] != sHW[2] || sH[3] * sW[3] != sHW[3]) {
      if (blockIdx.x == 0 && threadIdx.x == 0) atomicMax(fq.amax, 0xffffffffu);
      return;
    }
  }
  const int red_levels = sRedLevels;
  const int cpq = (Lq + 63) >> 6;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lig = lane % LPG;
  const int j = warp * UPW + lane / LPG;
  const int row = M * D;

  for (int it = 0; it < passes; ++it) {
    long long u = unit_of(static_cast<long long>(blockIdx.x) * passes + it, j, TILE, M, q_fast != 0);
    const bool active = u < units;   // inactive groups still take part in the shuffles
    if (!active) u = 0;
    const int m = static_cast<int>(u % M);
    const long long bq = u / M;
    const long long b = bq / Lq;
    const size_t voff = (static_cast<size_t>(b) * S * M + m) * D + lig * CH;
    const VT* vb = value + voff;
    float* gvb = grad_value + voff;
    __half* gvb_h = reinterpret_cast<__half*>(grad_value);      // F16ACC: image b of the replicated map, head m, this lane's channels
    const int q_of_unit = static_cast<int>(bq - b * Lq);
    if constexpr (F16ACC) gvb_h += (static_cast<size_t>(b) * fq.rows_h * M + m) * D + lig * CH;
    const float4* lp = reinterpret_cast<const float4*>(loc + static_cast<size_t>(u) * L * P * 2);
    const float4* ap = reinterpret_cast<const float4*>(aw + static_cast<size_t>(u) * L * P);
    float* glp = grad_loc + static_cast<size_t>(u) * L * P * 2;
    float* gap = grad_aw + static_cast<size_t>(u) * L * P;

    float g[CH];
    V::load(grad_out + static_cast<size_t>(u) * D + lig * CH, active, g);
    // Channels this lane REDUCES into grad_value.  fp32: the 4 it loads.  16-bit storage: a lane
    // loads 8 contiguous channels (16 B) but the fp32 gradient of those is 32 B, so issuing them as
    // two 16-byte reductions would leave every 32-byte L2 sector half written per instruction.
    // Instead the lane reduces channels [4*lig, +4) and [4*LPG + 4*lig, +4): each instruction of the
    // group then covers a contiguous 16*LPG bytes.  Costs one extra 2 x 8-byte load of grad_out per unit.
    float gr[CH];
    if constexpr (CH == 4) {
#pragma unroll
      for (int c = 0; c < CH; ++c) gr[c] = g[c];
    } else {
      const VT* gp = grad_out + static_cast<size_t>(u) * D;
      const uint2 lo = active ? __ldg(reinterpret_cast<const uint2*>(gp + 4 * lig)) : make_uint2(0u, 0u);
      const uint2 hi = active ? __ldg(reinterpret_cast<const uint2*>(gp + 4 * LPG + 4 * lig)) : make_uint2(0u, 0u);
      const uint32_t w[4] = {lo.x, lo.y, hi.x, hi.y};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if constexpr (sizeof(VT) == 2 && !std::is_same<VT, __half>::value) {
          gr[2 * i] = __uint_as_float(w[i] << 16);
          gr[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        } else {
          const float2 t2 = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
          gr[2 * i] = t2.x;
          gr[2 * i + 1] = t2.y;
        }
      }
    }

    // FUSEQ: after the reduce-scatter lane `lig` holds, per level, either (grad_x, grad_y) of point lig/2 (even lanes)
    // or grad_attn of point lig/2 (odd lanes); they are kept until the softmax dot product over all 16 samples is known.
    // Even lanes write their (d offset_x, d offset_y) pair inside the level loop; odd lanes keep (grad_attn, attn weight) of
    // their point for each of the 4 levels in eight scalars (selected by level, so nothing lives in local memory).
    // F16ACC: the lane's four grad_out channels times the scale, as two half2 (exact for 16-bit grad_out unless subnormal)
    __half2 gh01 = __half2(), gh23 = __half2();
    if constexpr (F16ACC) {
      const float sc = sScale;
      gh01 = __floats2half2_rn(gr[0] * sc, gr[1] * sc);
      gh23 = __floats2half2_rn(gr[2] * sc, gr[3] * sc);
    }
    float fq_g0 = 0.f, fq_g1 = 0.f, fq_g2 = 0.f, fq_g3 = 0.f, fq_w0 = 0.f, fq_w1 = 0.f, fq_w2 = 0.f, fq_w3 = 0.f;
    float fq_dot = 0.f;
    unsigned long long hit_bits = 0ull;      // mma_mode 2: ranges touched by this unit's samples

    // One level.  DO_RED is a COMPILE-TIME property of the call: with a run-time `if (l < red_levels)` around the reductions
    // the compiler fences each group of reductions off in its own branch region and can no longer hoist the next point's
    // corner loads above them -- measured: 959 -> 1527 us per launch at config 2 (profiles/r2_*).  Two loops over the two
    // level classes keep each body straight-line.
    auto level_body = [&](const int l, auto red_tag) {
      constexpr bool do_red = decltype(red_tag)::value;
      const int H = sH[l], W = sW[l];
      const size_t loff = static_cast<size_t>(sStart[l]) * row;
      const VT* vl = vb + loff;
      float* gvl = gvb + loff;
      __half* gvl_h = gvb_h;
      if constexpr (F16ACC) gvl_h += (static_cast<size_t>(sHStart[l]) + static_cast<size_t>((q_of_unit >> 2) & sRepMask[l]) * sHW[l]) * row;
      const float4 xy01 = __ldg(lp + 2 * l), xy23 = __ldg(lp + 2 * l + 1), a4 = __ldg(ap + l);
      const float xs[4] = {xy01.x, xy01.z, xy23.x, xy23.z};
      const float ys[4] = {xy01.y, xy01.w, xy23.y, xy23.w};
      const float as[4] = {a4.x, a4.y, a4.z, a4.w};
      int o_min = 0x7fffffff, o_max = -1;    // clamped corner-offset extent of this level's inside samples (mma_mode 2)
      float red[16];
      constexpr bool SH = SHARE && LPG >= 4;
      Tap<float> t4[4];
      if constexpr (SH) shared_taps<LPG>(xs, ys, H, W, lig, t4);
      float plh[4], plw[4];     // DOTS: fractional parts of the level's four points, for combine_dots()
#pragma unroll
      for (int p = 0; p < P; ++p) {
        Tap<float> t;
        if constexpr (SH) t = t4[p]; else t = make_tap<float>(xs[p], ys[p], H, W);
        t.c1 = t.c1 && active; t.c2 = t.c2 && active; t.c3 = t.c3 && active; t.c4 = t.c4 && active;
        if (HITS && t.ok) { o_min = min(o_min, t.o1); o_max = max(o_max, t.o4); }
        float v1[CH], v2[CH], v3[CH], v4[CH];
        const long long e1 = static_cast<long long>(t.o1) * row, e2 = static_cast<long long>(t.o2) * row;
        const long long e3 = static_cast<long long>(t.o3) * row, e4 = static_cast<long long>(t.o4) * row;
        V::load(vl + e1, t.c1, v1);
        V::load(vl + e2, t.c2, v2);
        V::load(vl + e3, t.c3, v3);
        V::load(vl + e4, t.c4, v4);
        const float k1 = t.hh * t.hw, k2 = t.hh * t.lw, k3 = t.lh * t.hw, k4 = t.lh * t.lw;
        const float a = as[p];
        if constexpr (DOTS) {
          float d1 = 0.f, d2 = 0.f, d3 = 0.f, d4 = 0.f;
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            d1 = fmaf(g[c], v1[c], d1); d2 = fmaf(g[c], v2[c], d2);
            d3 = fmaf(g[c], v3[c], d3); d4 = fmaf(g[c], v4[c], d4);
          }
          red[4 * p + 0] = d1; red[4 * p + 1] = d2; red[4 * p + 2] = d3; red[4 * p + 3] = d4;
          plh[p] = t.lh; plw[p] = t.lw;
        } else {
        float s_w = 0.f, s_h = 0.f, s_a = 0.f;
        float tg[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          tg[c] = g[c] * a;  // top_grad_value (im2col.cuh:115)
          // d(bilinear)/dh and /dw (im2col.cuh:123-151); guarded-off corners loaded as 0
          const float gh = t.hw * (v3[c] - v1[c]) + t.lw * (v4[c] - v2[c]);
          const float gw = t.hh * (v2[c] - v1[c]) + t.lh * (v4[c] - v3[c]);
          const float val = k1 * v1[c] + k2 * v2[c] + k3 * v3[c] + k4 * v4[c];
          s_a = fmaf(g[c], val, s_a);
          s_w = fmaf(gw, tg[c], s_w);
          s_h = fmaf(gh, tg[c], s_h);
        }
        red[4 * p + 0] = s_w * static_cast<float>(W);
        red[4 * p + 1] = s_h * static_cast<float>(H);
        red[4 * p + 2] = s_a;
        red[4 * p + 3] = 0.f;
        }
        if constexpr (F16ACC) {
          // weight (bilinear x attention) in fp32, rounded once to fp16; the two products round once more
          if (do_red && t.c1) { const __half2 kh = __float2half2_rn(k1 * a); red_add_v2h(gvl_h + e1, __hmul2(kh, gh01), __hmul2(kh, gh23)); }
          if (do_red && t.c2) { const __half2 kh = __float2half2_rn(k2 * a); red_add_v2h(gvl_h + e2, __hmul2(kh, gh01), __hmul2(kh, gh23)); }
          if (do_red && t.c3) { const __half2 kh = __float2half2_rn(k3 * a); red_add_v2h(gvl_h + e3, __hmul2(kh, gh01), __hmul2(kh, gh23)); }
          if (do_red && t.c4) { const __half2 kh = __float2half2_rn(k4 * a); red_add_v2h(gvl_h + e4, __hmul2(kh, gh01), __hmul2(kh, gh23)); }
        } else {
#pragma unroll
        for (int c0 = 0; c0 < CH; c0 += 4) {
          // element offset of this 4-channel slice relative to the lane's load slice (see `gr` above)
          const int ro = (CH == 4) ? 0 : ((c0 == 0 ? 4 * lig : 4 * LPG + 4 * lig) - CH * lig);
          const float r0 = gr[c0] * a, r1 = gr[c0 + 1] * a, r2 = gr[c0 + 2] * a, r3 = gr[c0 + 3] * a;
          if (do_red && t.c1) red_add_v4(gvl + e1 + ro, k1 * r0, k1 * r1, k1 * r2, k1 * r3);
          if (do_red && t.c2) red_add_v4(gvl + e2 + ro, k2 * r0, k2 * r1, k2 * r2, k2 * r3);
          if (do_red && t.c3) red_add_v4(gvl + e3 + ro, k3 * r0, k3 * r1, k3 * r2, k3 * r3);
          if (do_red && t.c4) red_add_v4(gvl + e4 + ro, k4 * r0, k4 * r1, k4 * r2, k4 * r3);
        }
        }
      }
      if (HITS && !do_red && o_max >= 0) {
        // every valid corner of the level's samples lies in [max(o_min, 0), min(o_max, H*W - 1)]; ranges are contiguous
        // pixel intervals, so the ranges touched are (a subset of) rid(lo) .. rid(hi): a clear bit is exact, a set one
        // may be a false positive (costs an empty product, never a wrong result)
        const int lo_ = max(o_min, 0), hi_ = min(o_max, H * W - 1);
        const int r_lo = plan.rbase[l] + lo_ / kR2Px, r_hi = plan.rbase[l] + hi_ / kR2Px;
        hit_bits |= (2ull << r_hi) - (1ull << r_lo);
      }
      reduce_scatter16<LPG>(red, lig);
      if constexpr (DOTS) combine_dots<LPG>(red, lig, plh, plw, as, static_cast<float>(W), static_cast<float>(H));
      if constexpr (FUSEQ) {
        // LPG == 8, R == 2: red[0], red[1] = entries 2*lig, 2*lig+1 of (w, h, a, pad) x 4 points
        const int p = lig >> 1;
        const float my_aw = p == 0 ? as[0] : (p == 1 ? as[1] : (p == 2 ? as[2] : as[3]));
        if (lig & 1) {                                           // odd lanes hold grad_attn of point p
          fq_dot = fmaf(red[0], my_aw, fq_dot);
          fq_g0 = l == 0 ? red[0] : fq_g0; fq_g1 = l == 1 ? red[0] : fq_g1; fq_g2 = l == 2 ? red[0] : fq_g2; fq_g3 = l == 3 ? red[0] : fq_g3;
          fq_w0 = l == 0 ? my_aw : fq_w0; fq_w1 = l == 1 ? my_aw : fq_w1; fq_w2 = l == 2 ? my_aw : fq_w2; fq_w3 = l == 3 ? my_aw : fq_w3;
        } else if (active && l < 4) {                            // even lanes hold (grad_x, grad_y) of point p: d(offset) goes out now
          float sx, sy;
          if (fq.ref_dim == 2) { sx = sInvW[l]; sy = sInvH[l]; }
          else {
            const float* rp = fq.ref + static_cast<size_t>(bq) * 4 * fq.ref_dim;
            sx = rp[l * 4 + 2] * 0.125f; sy = rp[l * 4 + 3] * 0.125f;      // * 0.5 / P with P == 4
          }
          const float dx = red[0] * sx, dy = red[1] * sy;
          uint32_t w32;
          if (fq.is_half) { __half2 t2 = __floats2half2_rn(dx, dy); w32 = *reinterpret_cast<uint32_t*>(&t2); }
          else { __nv_bfloat162 t2 = __floats2bfloat162_rn(dx, dy); w32 = *reinterpret_cast<uint32_t*>(&t2); }
          *reinterpret_cast<uint32_t*>(fq.dq + static_cast<size_t>(bq) * (3 * M * 16) + m * 32 + (l * 4 + p) * 2) = w32;
        }
      } else {
        if (active) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int idx = lig * R + r, p = idx >> 2, comp = idx & 3;
            if (comp < 2) glp[(l * P + p) * 2 + comp] = red[r];
            else if (comp == 2) gap[l * P + p] = red[r];
          }
        }
      }
    };
    if constexpr (MODE == 0) {
      for (int l = 0; l < L; ++l) level_body(l, std::true_type{});
    } else {
      int l = 0;
#pragma unroll 1
      for (; l < red_levels; ++l) level_body(l, std::true_type{});
#pragma unroll 1
      for (; l < L; ++l) level_body(l, std::false_type{});
    }
    if (HITS && active && lig == 0 && hit_bits != 0ull)
      atomicOr(hit + ((b * M + m) * cpq + (static_cast<int>(bq % Lq) >> 6)), hit_bits);
    if constexpr (FUSEQ) {
      // softmax dot product over the unit's 16 samples: the 4 odd lanes hold 4 terms each
      fq_dot += __shfl_xor_sync(0xffffffffu, fq_dot, 1);
      fq_dot += __shfl_xor_sync(0xffffffffu, fq_dot, 2);
      fq_dot += __shfl_xor_sync(0xffffffffu, fq_dot, 4);
      if (active && (lig & 1)) {
        const int p = lig >> 1;
        const int n_aw = M * 16;
        uint16_t* orow = fq.dq + static_cast<size_t>(bq) * (3 * n_aw) + 2 * n_aw + m * 16 + p;
        const float gs[4] = {fq_g0, fq_g1, fq_g2, fq_g3}, ws[4] = {fq_w0, fq_w1, fq_w2, fq_w3};
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          const float d = ws[l] * (gs[l] - fq_dot);
          uint16_t w16;
          if (fq.is_half) { __half t2 = __float2half_rn(d); w16 = *reinterpret_cast<uint16_t*>(&t2); }
          else { __nv_bfloat16 t2 = __float2bfloat16_rn(d); w16 = *reinterpret_cast<uint16_t*>(&t2); }
          orow[l * 4] = w16;
        }
      }
    }
  }
}

// =================================================================================================
// generic kernels: any D, any P, fp32 / fp64 / bf16 / f16 storage (accumulate in Acc)
// =================================================================================================
template <typename VT, typename AT>
__global__ void __launch_bounds__(256)
msda_fwd_generic_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                        const int64_t* __restrict__ lstart, const AT* __restrict__ loc,
                        const AT* __restrict__ aw, VT* __restrict__ out, int S, int M, int D, int L,
                        int Lq, int P, long long total) {
  __shared__ int sH[MSDA_MAX_LEVELS], sW[MSDA_MAX_LEVELS], sStart[MSDA_MAX_LEVELS];
  if (threadIdx.x < L) {
    sH[threadIdx.x] = static_cast<int>(shapes[2 * threadIdx.x]);
    sW[threadIdx.x] = static_cast<int>(shapes[2 * threadIdx.x + 1]);
    sStart[threadIdx.x] = static_cast<int>(lstart[threadIdx.x]);
  }
  __syncthreads();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % D);
  const long long u = idx / D;
  const int m = static_cast<int>(u % M);
  const long long b = (u / M) / Lq;
  const size_t row = static_cast<size_t>(M) * D;
  const VT* vb = value + (static_cast<size_t>(b) * S * M + m) * D + c;
  const AT* lp = loc + static_cast<size_t>(u) * L * P * 2;
  const AT* ap = aw + static_cast<size_t>(u) * L * P;
  AT acc = AT(0);
  for (int l = 0; l < L; ++l) {
    const int H = sH[l], W = sW[l];
    const VT* vl = vb + static_cast<size_t>(sStart[l]) * row;
    for (int p = 0; p < P; ++p) {
      const Tap<AT> t = make_tap<AT>(lp[(l * P + p) * 2], lp[(l * P + p) * 2 + 1], H, W);
      const AT v1 = t.c1 ? AT(to_acc(vl[static_cast<long long>(t.o1) * row])) : AT(0);
      const AT v2 = t.c2 ? AT(to_acc(vl[static_cast<long long>(t.o2) * row])) : AT(0);
      const AT v3 = t.c3 ? AT(to_acc(vl[static_cast<long long>(t.o3) * row])) : AT(0);
      const AT v4 = t.c4 ? AT(to_acc(vl[static_cast<long long>(t.o4) * row])) : AT(0);
      const AT val = t.hh * t.hw * v1 + t.hh * t.lw * v2 + t.lh * t.hw * v3 + t.lh * t.lw * v4;
      acc += val * ap[l * P + p];
    }
  }
  out[idx] = from_acc<VT, AT>(acc);
}

// one warp per (b, q, m) unit; lanes stride the channels
template <typename VT, typename AT>
__global__ void __launch_bounds__(256)
msda_bwd_generic_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                        const int64_t* __restrict__ lstart, const AT* __restrict__ loc,
                        const AT* __restrict__ aw, const VT* __restrict__ grad_out,
                        AT* __restrict__ grad_value, AT* __restrict__ grad_loc,
                        AT* __restrict__ grad_aw, int S, int M, int D, int L, int Lq, int P,
                        long long units) {
  __shared__ int sH[MSDA_MAX_LEVELS], sW[MSDA_MAX_LEVELS], sStart[MSDA_MAX_LEVELS];
  if (threadIdx.x < L) {
    sH[threadIdx.x] = static_cast<int>(shapes[2 * threadIdx.x]);
    sW[threadIdx.x] = static_cast<int>(shapes[2 * threadIdx.x + 1]);
    sStart[threadIdx.x] = static_cast<int>(lstart[threadIdx.x]);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long u = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (u >= units) return;  // whole warp leaves together
  const int m = static_cast<int>(u % M);
  const long long b = (u / M) / Lq;
  const size_t row = static_cast<size_t>(M) * D;
  const size_t voff = (static_cast<size_t>(b) * S * M + m) * D;
  const AT* lp = loc + static_cast<size_t>(u) * L * P * 2;
  const AT* ap = aw + static_cast<size_t>(u) * L * P;
  const VT* gp = grad_out + static_cast<size_t>(u) * D;
  for (int l = 0; l < L; ++l) {
    const int H = sH[l], W = sW[l];
    const size_t loff = voff + static_cast<size_t>(sStart[l]) * row;
    for (int p = 0; p < P; ++p) {
      const Tap<AT> t = make_tap<AT>(lp[(l * P + p) * 2], lp[(l * P + p) * 2 + 1], H, W);
      const AT a = ap[l * P + p];
      const AT k1 = t.hh * t.hw, k2 = t.hh * t.lw, k3 = t.lh * t.hw, k4 = t.lh * t.lw;
      AT s_w = AT(0), s_h = AT(0), s_a = AT(0);
      for (int c = lane; c < D; c += 32) {
        const AT g = AT(to_acc(gp[c])), tg = g * a;
        const size_t e1 = loff + static_cast<long long>(t.o1) * row + c, e2 = loff + static_cast<long long>(t.o2) * row + c;
        const size_t e3 = loff + static_cast<long long>(t.o3) * row + c, e4 = loff + static_cast<long long>(t.o4) * row + c;
        const AT v1 = t.c1 ? AT(to_acc(value[e1])) : AT(0), v2 = t.c2 ? AT(to_acc(value[e2])) : AT(0);
        const AT v3 = t.c3 ? AT(to_acc(value[e3])) : AT(0), v4 = t.c4 ? AT(to_acc(value[e4])) : AT(0);
        if (t.c1) atomicAdd(grad_value + e1, k1 * tg);
        if (t.c2) atomicAdd(grad_value + e2, k2 * tg);
        if (t.c3) atomicAdd(grad_value + e3, k3 * tg);
        if (t.c4) atomicAdd(grad_value + e4, k4 * tg);
        s_a += g * (k1 * v1 + k2 * v2 + k3 * v3 + k4 * v4);
        s_w += (t.hh * (v2 - v1) + t.lh * (v4 - v3)) * tg;
        s_h += (t.hw * (v3 - v1) + t.lw * (v4 - v2)) * tg;
      }
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        s_w += __shfl_xor_sync(0xffffffffu, s_w, o);
        s_h += __shfl_xor_sync(0xffffffffu, s_h, o);
        s_a += __shfl_xor_sync(0xffffffffu, s_a, o);
      }
      if (lane == 0) {
        grad_loc[static_cast<size_t>(u) * L * P * 2 + (l * P + p) * 2] = s_w * AT(W);
        grad_loc[static_cast<size_t>(u) * L * P * 2 + (l * P + p) * 2 + 1] = s_h * AT(H);
        grad_aw[static_cast<size_t>(u) * L * P + l * P + p] = s_a;
      }
    }
  }
}

// =================================================================================================
// host-side launchers
// =================================================================================================
Tuning g_tuning;
long long g_launches = 0;

template <typename VT, int D>
static cudaError_t launch_fwd_vec(const VT* value, const int64_t* shapes, const int64_t* lstart,
                                  const float* loc, const float* aw, VT* out, int N, int S, int M, int L,
                                  int Lq, cudaStream_t st) {
  constexpr int TILE = (32 / (D / Vec<VT>::CH)) * kWarpsPerBlock;
  const long long units = static_cast<long long>(N) * Lq * M;
  const int q_fast = (g_tuning.fwd_q_fast && TILE % M == 0) ? 1 : 0;
  const int passes = g_tuning.fwd_passes;
  const long long blocks = (units + static_cast<long long>(TILE) * passes - 1) / (static_cast<long long>(TILE) * passes);
  const dim3 grid(static_cast<unsigned>(blocks));
  ++g_launches;
  if (g_tuning.tap_share && g_tuning.fwd_sample_batch == 4) {
    msda_fwd_vec_kernel<VT, D, 4, true><<<grid, kThreads, 0, st>>>(value, shapes, lstart, loc, aw, out, S, M, L, Lq, units, passes, q_fast);
    return cudaGetLastError();
  }
  switch (g_tuning.fwd_sample_batch) {
    case 1: msda_fwd_vec_kernel<VT, D, 1, false><<<grid, kThreads, 0, st>>>(value, shapes, lstart, loc, aw, out, S, M, L, Lq, units, passes, q_fast); break;
    case 4: msda_fwd_vec_kernel<VT, D, 4, false><<<grid, kThreads, 0, st>>>(value, shapes, lstart, loc, aw, out, S, M, L, Lq, units, passes, q_fast); break;
    default: msda_fwd_vec_kernel<VT, D, 2, false><<<grid, kThreads, 0, st>>>(value, shapes, lstart, loc, aw, out, S, M, L, Lq, units, passes, q_fast); break;
  }
  return cudaGetLastError();
}

template <typename VT>
cudaError_t launch_scatter_mma(const int64_t* shapes, const int64_t* lstart, const float* loc, const float* aw,
                               const VT* grad_out, float* gv, int N, int S, int M, int L, int Lq, cudaStream_t st);

template <typename VT>
cudaError_t launch_scatter_mma2(const int64_t* shapes, const int64_t* lstart, const float* loc, const float* aw,
                                const VT* grad_out, float* gv, const unsigned long long* hit, int N, int S, int M, int L,
                                int Lq, int max_levels, cudaStream_t st);

// Caller-provided scratch of the current backward call (msda_backward_16_ws); null for the entry points without one.
thread_local void* t_workspace = nullptr;
thread_local long long t_workspace_bytes = 0;

long long backward_workspace_bytes(int N, int M, int Lq) { return static_cast<long long>(N) * M * ((Lq + 63) / 64) * 8; }

template <typename VT, int D, typename V, bool FUSEQ>
static cudaError_t launch_bwd_vec_t(const VT* value, const int64_t* shapes, const int64_t* lstart,
                                    const float* loc, const float* aw, const VT* grad_out, float* gv,
                                    float* gl, float* ga, int N, int S, int M, int L, int Lq, cudaStream_t st, FuseQ fq) {
  constexpr int TILE = (32 / (D / V::CH)) * kWarpsPerBlock;
  const long long units = static_cast<long long>(N) * Lq * M;
  const int q_fast = (g_tuning.bwd_q_fast && TILE % M == 0) ? 1 : 0;
  const int passes = bwd_passes_auto((units + TILE - 1) / TILE, false);
  const long long blocks = (units + static_cast<long long>(TILE) * passes - 1) / (static_cast<long long>(TILE) * passes);
  // 16-bit storage, D = 32: the coarse tail of the level list accumulates in tensor memory (msda_scatter_mma.cu); this
  // kernel then skips those reductions.  Both kernels derive the same split from the device-side shapes.
  int mma_mode = 0;
  unsigned long long* hit = nullptr;
  if constexpr (sizeof(VT) == 2 && D == 32) {
    if (g_tuning.bwd_mma && units >= g_tuning.bwd_mma_min_units) {
      mma_mode = 1;
      if (g_tuning.bwd_mma_levels > 0 && t_workspace != nullptr && t_workspace_bytes >= backward_workspace_bytes(N, M, Lq) &&
          (reinterpret_cast<uintptr_t>(t_workspace) & 7u) == 0) {
        mma_mode = 2;
        hit = static_cast<unsigned long long*>(t_workspace);
        cudaError_t ez = cudaMemsetAsync(hit, 0, static_cast<size_t>(backward_workspace_bytes(N, M, Lq)), st);
        if (ez != cudaSuccess) return ez;
      }
    }
  }
  ++g_launches;
  bool launched = false;
  const dim3 grid(static_cast<unsigned>(blocks));
#define MSDA_BWD_LAUNCH(MODE_, SHARE_)                                                                               \
  do {                                                                                                               \
    if (g_tuning.bwd_dots)                                                                                           \
      msda_bwd_vec_kernel<VT, D, V, FUSEQ, MODE_, SHARE_, true><<<grid, kThreads, 0, st>>>(                          \
          value, shapes, lstart, loc, aw, grad_out, gv, gl, ga, S, M, L, Lq, units, passes, q_fast, mma_mode,        \
          g_tuning.bwd_mma_levels, hit, fq);                                                                         \
    else                                                                                                             \
      msda_bwd_vec_kernel<VT, D, V, FUSEQ, MODE_, SHARE_, false><<<grid, kThreads, 0, st>>>(                         \
          value, shapes, lstart, loc, aw, grad_out, gv, gl, ga, S, M, L, Lq, units, passes, q_fast, mma_mode,        \
          g_tuning.bwd_mma_levels, hit, fq);                                                                         \
  } while (0)
  if constexpr (sizeof(VT) == 2 && D == 32) {
    if (mma_mode == 2) {
      if (g_tuning.tap_share) MSDA_BWD_LAUNCH(2, true); else MSDA_BWD_LAUNCH(2, false);
      launched = true;
    } else if (mma_mode == 1) {
      if (g_tuning.tap_share) MSDA_BWD_LAUNCH(1, true); else MSDA_BWD_LAUNCH(1, false);
      launched = true;
    }
  }
  if (!launched) {
    if (g_tuning.tap_share) MSDA_BWD_LAUNCH(0, true); else MSDA_BWD_LAUNCH(0, false);
  }
#undef MSDA_BWD_LAUNCH
  cudaError_t e = cudaGetLastError();
  if constexpr (sizeof(VT) == 2 && D == 32) {
    if (e == cudaSuccess && mma_mode == 1) e = launch_scatter_mma<VT>(shapes, lstart, loc, aw, grad_out, gv, N, S, M, L, Lq, st);
    if (e == cudaSuccess && mma_mode == 2)
      e = launch_scatter_mma2<VT>(shapes, lstart, loc, aw, grad_out, gv, hit, N, S, M, L, Lq, g_tuning.bwd_mma_levels, st);
  }
  return e;
}

template <typename VT, int D>
static cudaError_t launch_bwd_vec(const VT* value, const int64_t* shapes, const int64_t* lstart,
                                  const float* loc, const float* aw, const VT* grad_out, float* gv,
                                  float* gl, float* ga, int N, int S, int M, int L, int Lq, cudaStream_t st) {
  const FuseQ none{nullptr, nullptr, 0, 0, nullptr, 0, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
  // narrow layout needs D/4 <= 16 lanes per unit
  if constexpr (Vec<VT>::CH == 8 && D <= 64) {
    if (g_tuning.bwd_narrow)
      return launch_bwd_vec_t<VT, D, VecQ<VT>, false>(value, shapes, lstart, loc, aw, grad_out, gv, gl, ga, N, S, M, L, Lq, st, none);
  }
  return launch_bwd_vec_t<VT, D, Vec<VT>, false>(value, shapes, lstart, loc, aw, grad_out, gv, gl, ga, N, S, M, L, Lq, st, none);
}

// 16-bit storage, D == 32, L == 4, P == 4: backward with the query-side epilogue backward fused in (see FuseQ)
template <typename VT>
cudaError_t backward_fused_q(const VT* value, const int64_t* shapes, const int64_t* lstart, const float* loc, const float* aw,
                             const VT* grad_out, float* gv, const float* ref, int ref_dim, void* dq, int is_half, int N, int S,
                             int M, int Lq, cudaStream_t st) {
  const FuseQ fq{ref, static_cast<uint16_t*>(dq), ref_dim, is_half, nullptr, 0, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
  return launch_bwd_vec_t<VT, 32, VecQ<VT>, true>(value, shapes, lstart, loc, aw, grad_out, gv, nullptr, nullptr, N, S, M, 4, Lq, st, fq);
}
template cudaError_t backward_fused_q<__nv_bfloat16>(const __nv_bfloat16*, const int64_t*, const int64_t*, const float*, const float*, const __nv_bfloat16*, float*, const float*, int, void*, int, int, int, int, int, cudaStream_t);
template cudaError_t backward_fused_q<__half>(const __half*, const int64_t*, const int64_t*, const float*, const float*, const __half*, float*, const float*, int, void*, int, int, int, int, int, cudaStream_t);

// max |grad_out| over a 16-bit tensor, as float bits (monotone for non-negative floats, NaN above inf above finite), by
// atomicMax into a zero-initialised word: the input of f16acc_scale().
template <bool IS_HALF>
__global__ void __launch_bounds__(256) grad_amax_kernel(const uint4* __restrict__ g, long long n8, uint32_t* __restrict__ amax) {
  uint32_t m = 0u;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 v = __ldg(g + i);
    m = __vmaxu2(m, __vmaxu2(__vmaxu2(v.x & 0x7fff7fffu, v.y & 0x7fff7fffu), __vmaxu2(v.z & 0x7fff7fffu, v.w & 0x7fff7fffu)));
  }
  uint32_t m16 = max(m >> 16, m & 0xffffu);
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) m16 = max(m16, __shfl_xor_sync(0xffffffffu, m16, o));
  __shared__ uint32_t sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m16;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) m16 = max(m16, sm[w]);
    const uint32_t bits = IS_HALF ? __float_as_uint(__half2float(__ushort_as_half(static_cast<unsigned short>(m16)))) : (m16 << 16);
    if (bits != 0u) atomicMax(amax, bits);
  }
}

// 16-bit storage, D == 32, L == 4, P == 4: the fused-query backward with grad_value accumulated in scaled fp16
// (F16ACC).  gv_h [N*S*M*32] halves and the amax word must arrive zeroed.
template <typename VT>
cudaError_t backward_fused_q_h16(const VT* value, const int64_t* shapes, const int64_t* lstart, const float* loc, const float* aw,
                                 const VT* grad_out, void* gv_h, const int64_t* shapes_host, uint32_t* amax, const float* ref,
                                 int ref_dim, void* dq, int is_half, int N, int S, int M, int Lq, cudaStream_t st) {
  using V = VecQ<VT>;
  constexpr int D = 32;
  constexpr int TILE = (32 / (D / V::CH)) * kWarpsPerBlock;
  const long long units = static_cast<long long>(N) * Lq * M;
  const long long n8 = units * D / 8;
  const unsigned ablocks = static_cast<unsigned>(std::min<long long>((n8 + 255) / 256, 148 * 8));
  g_launches += 2;
  grad_amax_kernel<std::is_same<VT, __half>::value><<<ablocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(grad_out), n8, amax);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  FuseQ fq{ref, static_cast<uint16_t*>(dq), ref_dim, is_half, amax, 0, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
  for (int l = 0; l < 4; ++l) {
    const long long hw = shapes_host[2 * l] * shapes_host[2 * l + 1];
    const int k = f16acc_replicas(Lq, hw);
    fq.h_start[l] = static_cast<int>(fq.rows_h); fq.h_hw[l] = static_cast<int>(hw); fq.h_mask[l] = k - 1;
    fq.rows_h += hw * k;
  }
  const int q_fast = (g_tuning.bwd_q_fast && TILE % M == 0) ? 1 : 0;
  const int passes = bwd_passes_auto((units + TILE - 1) / TILE, true);
  const long long blocks = (units + static_cast<long long>(TILE) * passes - 1) / (static_cast<long long>(TILE) * passes);
  msda_bwd_vec_kernel<VT, D, V, true, 0, false, true, true><<<dim3(static_cast<unsigned>(blocks)), kThreads, 0, st>>>(
      value, shapes, lstart, loc, aw, grad_out, static_cast<float*>(gv_h), nullptr, nullptr, S, M, 4, Lq, units, passes, q_fast, 0, 0,
      nullptr, fq);
  return cudaGetLastError();
}
template cudaError_t backward_fused_q_h16<__nv_bfloat16>(const __nv_bfloat16*, const int64_t*, const int64_t*, const float*, const float*, const __nv_bfloat16*, void*, const int64_t*, uint32_t*, const float*, int, void*, int, int, int, int, int, cudaStream_t);
template cudaError_t backward_fused_q_h16<__half>(const __half*, const int64_t*, const int64_t*, const float*, const float*, const __half*, void*, const int64_t*, uint32_t*, const float*, int, void*, int, int, int, int, int, cudaStream_t);

// Storage types whose locations / weights / gradients are fp32 (float, bf16, half).
template <typename VT>
cudaError_t forward_f32acc(const VT* value, const int64_t* shapes, const int64_t* lstart, const float* loc,
                           const float* aw, VT* out, int N, int S, int M, int D, int L, int Lq, int P,
                           cudaStream_t st) {
  constexpr int CH = Vec<VT>::CH;
  if (P == 4) {
    if (D == 32) return launch_fwd_vec<VT, 32>(value, shapes, lstart, loc, aw, out, N, S, M, L, Lq, st);
    if (D == 64) return launch_fwd_vec<VT, 64>(value, shapes, lstart, loc, aw, out, N, S, M, L, Lq, st);
    if (D == 16) return launch_fwd_vec<VT, 16>(value, shapes, lstart, loc, aw, out, N, S, M, L, Lq, st);
    if constexpr (CH == 8) if (D == 128) return launch_fwd_vec<VT, 128>(value, shapes, lstart, loc, aw, out, N, S, M, L, Lq, st);
  }
  const long long total = static_cast<long long>(N) * Lq * M * D;
  ++g_launches;
  msda_fwd_generic_kernel<VT, float><<<dim3(static_cast<unsigned>((total + 255) / 256)), 256, 0, st>>>(
      value, shapes, lstart, loc, aw, out, S, M, D, L, Lq, P, total);
  return cudaGetLastError();
}

template <typename VT>
cudaError_t backward_f32acc(const VT* value, const int64_t* shapes, const int64_t* lstart, const float* loc,
                            const float* aw, const VT* grad_out, float* gv, float* gl, float* ga, int N,
                            int S, int M, int D, int L, int Lq, int P, cudaStream_t st) {
  constexpr int CH = Vec<VT>::CH;
  if (P == 4) {
    if (D == 32) return launch_bwd_vec<VT, 32>(value, shapes, lstart, loc, aw, grad_out, gv, gl, ga, N, S, M, L, Lq, st);
    if (D == 64) return launch_bwd_vec<VT, 64>(value, shapes, lstart, loc, aw, grad_out, gv, gl, ga, N, S, M, L, Lq, st);
    if (D == 16) return launch_bwd_vec<VT, 16>(value, shapes, lstart, loc, aw, grad_out, gv, gl, ga, N, S, M, L, Lq, st);
    if constexpr (CH == 8) if (D == 128) return launch_bwd_vec<VT, 128>(value, shapes, lstart, loc, aw, grad_out, gv, gl, ga, N, S, M, L, Lq, st);
  }
  const long long units = static_cast<long long>(N) * Lq * M;
  ++g_launches;
  msda_bwd_generic_kernel<VT, float><<<dim3(static_cast<unsigned>((units + 7) / 8)), 256, 0, st>>>(
      value, shapes, lstart, loc, aw, grad_out, gv, gl, ga, S, M, D, L, Lq, P, units);
  return cudaGetLastError();
}

cudaError_t forward_f64(const double* value, const int64_t* shapes, const int64_t* lstart, const double* loc,
                        const double* aw, double* out, int N, int S, int M, int D, int L, int Lq, int P,
                        cudaStream_t st) {
  const long long total = static_cast<long long>(N) * Lq * M * D;
  ++g_launches;
  msda_fwd_generic_kernel<double, double><<<dim3(static_cast<unsigned>((total + 255) / 256)), 256, 0, st>>>(
      value, shapes, lstart, loc, aw, out, S, M, D, L, Lq, P, total);
  return cudaGetLastError();
}

cudaError_t backward_f64(const double* value, const int64_t* shapes, const int64_t* lstart, const double* loc,
                         const double* aw, const double* grad_out, double* gv, double* gl, double* ga, int N,
                         int S, int M, int D, int L, int Lq, int P, cudaStream_t st) {
  const long long units = static_cast<long long>(N) * Lq * M;
  ++g_launches;
  msda_bwd_generic_kernel<double, double><<<dim3(static_cast<unsigned>((units + 7) / 8)), 256, 0, st>>>(
      value, shapes, lstart, loc, aw, grad_out, gv, gl, ga, S, M, D, L, Lq, P, units);
  return cudaGetLastError();
}

template cudaError_t forward_f32acc<float>(const float*, const int64_t*, const int64_t*, const float*, const float*, float*, int, int, int, int, int, int, int, cudaStream_t);
template cudaError_t forward_f32acc<__nv_bfloat16>(const __nv_bfloat16*, const int64_t*, const int64_t*, const float*, const float*, __nv_bfloat16*, int, int, int, int, int, int, int, cudaStream_t);
template cudaError_t forward_f32acc<__half>(const __half*, const int64_t*, const int64_t*, const float*, const float*, __half*, int, int, int, int, int, int, int, cudaStream_t);
template cudaError_t backward_f32acc<float>(const float*, const int64_t*, const int64_t*, const float*, const float*, const float*, float*, float*, float*, int, int, int, int, int, int, int, cudaStream_t);
template cudaError_t backward_f32acc<__nv_bfloat16>(const __nv_bfloat16*, const int64_t*, const int64_t*, const float*, const float*, const __nv_bfloat16*, float*, float*, float*, int, int, int, int, int, int, int, cudaStream_t);
template cudaError_t backward_f32acc<__half>(const __half*, const int64_t*, const int64_t*, const float*, const float*, const __half*, float*, float*, float*, int, int, int, int, int, int, int, cudaStream_t);

}  // namespace msda
