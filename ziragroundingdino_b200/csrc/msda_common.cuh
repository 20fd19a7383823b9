// Shared device helpers for the sm_100a multi-scale deformable attention kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstring>

#include "../../include/msda_b200.h"

namespace msda {

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;

// Kernel-variant knobs (msda_b200_set_tuning); the defaults are the shipped configuration.
struct Tuning {
  int fwd_sample_batch = 4;  // samples whose 4 corner loads are issued back to back
  int fwd_q_fast = 1;        // lane groups of a warp span consecutive queries of one head
  int fwd_passes = 1;        // consecutive unit tiles handled by one CTA
  int bwd_q_fast = 1;
  int bwd_passes = 0;        // 0 = by launch size (bwd_passes_auto()); 1..64 = fixed
  int bwd_narrow = 1;        // 16-bit storage: 4 channels per lane in the backward kernel (full-line reductions)
  int bwd_dots = 1;          // backward: per-sample sums from the four corner dot products (combine_dots) instead of per channel
                             // (measured -13 % on the scatter kernel, same parity: profiles/r2_scatter_variants.jsonl)
  int tap_share = 0;         // taps of a level computed once per lane group and exchanged by shuffles (shared_taps())
  int bwd_mma = 0;           // 16-bit storage, D = 32, P = 4: coarse levels accumulate in tensor memory (msda_scatter_mma.cu).
                             // Opt-in: correct (tests), but the scatter kernel is issue-bound, not reduction-bound, so dropping
                             // half of its reductions saves less (~125 us) than the accumulation kernel costs (~160-190 us)
  int bwd_mma_levels = 0;    // > 0 and a workspace given: the last `bwd_mma_levels` levels go through the range-planned
                             // second-generation kernel (msda_scatter_mma2.cu) instead; 0 = first-generation tail only
  int bwd_mma_min_units = 131072;   // ... when N*Lq*M is at least this (below it the reductions it saves do not pay for a launch)
};
extern Tuning g_tuning;
// Unit tiles per CTA of the scatter kernel when the tuning key is 0.  A CTA's prologue (shapes, 1/W, 1/H, two barriers) is
// paid per CTA, and a CTA with one tile lives ~17 us: two tiles per CTA measure -2 % on the fp32-accumulating kernel at
// config 2's launch (937 -> 917 us; 3, 4: the same; 8: 932; 16: 969), four tiles -7 % on the scaled-fp16 one, which is
// latency- rather than bandwidth-bound (893 -> 828 us; profiles/r2_bwd_passes.jsonl).  Small launches (the decoder's 1 800
// tiles at 8 images) keep one tile per CTA so that the grid still covers the 444 resident CTAs several times.
inline int bwd_passes_auto(long long tiles, bool f16acc) {
  if (g_tuning.bwd_passes > 0) return g_tuning.bwd_passes;
  if (tiles >= 8192) return f16acc ? 4 : 2;
  return tiles >= 4096 ? 2 : 1;
}
extern long long g_launches;
// scratch of the backward call in flight on this thread (set by msda_backward_16_ws around the launch; msda_core.cu)
extern thread_local void* t_workspace;
extern thread_local long long t_workspace_bytes;
long long backward_workspace_bytes(int N, int M, int Lq);

// ---- storage-type traits: one lane always moves 16 bytes of channels -----------------------------
template <typename VT> struct Vec;

template <> struct Vec<float> {
  static constexpr int CH = 4;
  using Acc = float;
  __device__ static __forceinline__ void load(const float* p, bool ok, float (&f)[4]) {
    float4 r = ok ? __ldg(reinterpret_cast<const float4*>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
    f[0] = r.x; f[1] = r.y; f[2] = r.z; f[3] = r.w;
  }
  __device__ static __forceinline__ void store(float* p, const float (&f)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
};

template <> struct Vec<__nv_bfloat16> {
  static constexpr int CH = 8;
  using Acc = float;
  __device__ static __forceinline__ void load(const __nv_bfloat16* p, bool ok, float (&f)[8]) {
    uint4 r = ok ? __ldg(reinterpret_cast<const uint4*>(p)) : make_uint4(0u, 0u, 0u, 0u);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static __forceinline__ void store(__nv_bfloat16* p, const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

template <> struct Vec<__half> {
  static constexpr int CH = 8;
  using Acc = float;
  __device__ static __forceinline__ void load(const __half* p, bool ok, float (&f)[8]) {
    uint4 r = ok ? __ldg(reinterpret_cast<const uint4*>(p)) : make_uint4(0u, 0u, 0u, 0u);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  __device__ static __forceinline__ void store(__half* p, const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// Narrow variants for the backward kernel: 4 channels per lane whatever the storage type (fp32: the same 16-byte
// load; bf16/f16: an 8-byte load).  With 4 channels per lane a unit is D/4 lanes and ONE 128-bit reduction
// instruction covers the unit's whole D*4-byte fp32 gradient row, i.e. one request per 128-byte line instead of two
// half-line requests -- the backward kernel is bound by L1TEX->XBAR reduction requests.
template <typename VT> struct VecQ;
template <> struct VecQ<float> : Vec<float> {};
template <> struct VecQ<__nv_bfloat16> {
  static constexpr int CH = 4;
  __device__ static __forceinline__ void load(const __nv_bfloat16* p, bool ok, float (&f)[4]) {
    const uint2 r = ok ? __ldg(reinterpret_cast<const uint2*>(p)) : make_uint2(0u, 0u);
    f[0] = __uint_as_float(r.x << 16); f[1] = __uint_as_float(r.x & 0xffff0000u);
    f[2] = __uint_as_float(r.y << 16); f[3] = __uint_as_float(r.y & 0xffff0000u);
  }
};
template <> struct VecQ<__half> {
  static constexpr int CH = 4;
  __device__ static __forceinline__ void load(const __half* p, bool ok, float (&f)[4]) {
    const uint2 r = ok ? __ldg(reinterpret_cast<const uint2*>(p)) : make_uint2(0u, 0u);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
  }
};

// 128-bit reduction into fp32 global memory (REDG.E.ADD.F32x4 on sm_100a).
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// 64-bit packed-half reduction (4 channels) into an fp16 map (REDG.E.ADD.F16x2.RN on sm_100a).
__device__ __forceinline__ void red_add_v2h(__half* p, __half2 a, __half2 b) {
  asm volatile("red.relaxed.gpu.global.add.noftz.v2.f16x2 [%0], {%1, %2};"
               :: "l"(p), "r"(*reinterpret_cast<const uint32_t*>(&a)), "r"(*reinterpret_cast<const uint32_t*>(&b)) : "memory");
}

// Scaled-fp16 accumulation of grad_value: the power of two s with s * amax < 2^floor(log2(60000 / Lq)).  For one
// (pixel, head) row the contributions of a query sum to at most |grad_out| (its attention weights sum to 1 and a sample
// touches a pixel with at most one corner of weight <= 1), so every partial sum stays below Lq * s * amax < 60000 < the
// fp16 maximum: the accumulation cannot overflow whatever the sampling pattern.  Producer and consumers call this with
// the same (amax bits, Lq) and get the same scale.
__host__ __device__ inline uint32_t f16acc_bits(float x) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(x);
#else
  uint32_t b; memcpy(&b, &x, 4); return b;
#endif
}
__host__ __device__ inline float f16acc_scale(uint32_t amax_bits, int Lq) {
  if (amax_bits == 0u || amax_bits >= 0x7f800000u || Lq <= 0) return 1.f;   // zero, NaN or inf gradient: any scale will do
  // integer form of frexp(): x = f * 2^e, f in [0.5, 1)  <=>  e = biased exponent - 126 (a subnormal maximum is treated as
  // 2^-126: a smaller scale than necessary, never a larger one)
  const int e = static_cast<int>(amax_bits >> 23) - 126;                                   // amax < 2^e
  const int be = static_cast<int>(f16acc_bits(60000.f / static_cast<float>(Lq)) >> 23) - 126;   // 2^(be-1) <= 60000 / Lq
  int k = be - 1 - e;
  k = k < -126 ? -126 : (k > 126 ? 126 : k);       // a normal float; only max * Lq > 60000 * 2^126 (~5e42) meets the lower clamp
  const uint32_t sb = static_cast<uint32_t>(127 + k) << 23;
#ifdef __CUDA_ARCH__
  return __uint_as_float(sb);
#else
  float r; memcpy(&r, &sb, 4); return r;
#endif
}

// fp16 has 11 significand bits: a row that receives n contributions of similar size loses ~log2(n) of them, and the two
// coarse levels of an encoder launch receive hundreds to thousands per (pixel, head) row (16 * Lq / (H*W) if the samples
// spread evenly).  The scaled-fp16 map therefore holds K_l REPLICAS of level l's rows -- query q accumulates into replica
// (q / 4) mod K_l (the four consecutive queries a warp holds keep sharing rows, so their reductions still leave as one
// request per line), the consumer sums the replicas in fp32 -- with K_l the power of two that brings the expected count per row
// to <= 96 (Swin-T 800x1333, Lq = 22 223: K = 1, 1, 4, 16; 29 468 map rows per image instead of 22 223).
__host__ __device__ inline int f16acc_replicas(int Lq, long long hw) {
  const long long n = 16ll * Lq / (hw > 0 ? hw : 1);
  int k = 1;
  while (k < 64 && n > 96ll * k) k <<= 1;
  return k;
}

// ---- scalar element access for the generic kernels -----------------------------------------------
template <typename VT> struct Scalar { using Acc = float; };
template <> struct Scalar<double> { using Acc = double; };
__device__ __forceinline__ float to_acc(float v) { return v; }
__device__ __forceinline__ double to_acc(double v) { return v; }
__device__ __forceinline__ float to_acc(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_acc(__half v) { return __half2float(v); }
template <typename VT, typename A> __device__ __forceinline__ VT from_acc(A v);
template <> __device__ __forceinline__ float from_acc<float, float>(float v) { return v; }
template <> __device__ __forceinline__ double from_acc<double, double>(double v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_acc<__nv_bfloat16, float>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_acc<__half, float>(float v) { return __float2half_rn(v); }

// ---- the reference's sampling-coordinate arithmetic ----------------------------------------------
// ms_deform_im2col_cuda.cuh:285-286 writes `loc * size - 0.5`.  As nvcc compiles the reference op (default -fmad=true)
// this is ONE fused multiply-add -- `FFMA R, loc, size, -0.5` / `DFMA ...` in the SASS of oracle/_ref, forward and
// backward kernels alike -- so it is a single rounding here too: with two roundings the result sits 1.3e-5 from the
// reference op at config 1 (sigma = 1 values), i.e. over the 1e-5 bar; with the FMA it agrees to ~2e-7.
__device__ __forceinline__ float im_coord(float loc, int size) { return __fmaf_rn(loc, static_cast<float>(size), -0.5f); }
__device__ __forceinline__ double im_coord(double loc, int size) { return __fma_rn(loc, static_cast<double>(size), -0.5); }
__device__ __forceinline__ int floor_int(float x) { return __float2int_rd(x); }
__device__ __forceinline__ int floor_int(double x) { return __double2int_rd(x); }

// One bilinear sample: corner offsets (in pixels, relative to the level origin), guards and weights.
template <typename A> struct Tap {
  bool ok;          // sample inside (-1, size) on both axes (im2col.cuh:288)
  bool c1, c2, c3, c4;  // corner guards (im2col.cuh:56-78)
  int o1, o2, o3, o4;   // pixel offsets h*W + w of the four corners
  A lh, lw, hh, hw;
};

template <typename A>
__device__ __forceinline__ Tap<A> make_tap(A loc_x, A loc_y, int H, int W) {
  Tap<A> t;
  const A h_im = im_coord(loc_y, H), w_im = im_coord(loc_x, W);
  t.ok = (h_im > A(-1)) && (w_im > A(-1)) && (h_im < A(H)) && (w_im < A(W));
  const int h0 = floor_int(h_im), w0 = floor_int(w_im);
  const int h1 = h0 + 1, w1 = w0 + 1;
  t.lh = h_im - A(h0);
  t.lw = w_im - A(w0);
  t.hh = A(1) - t.lh;
  t.hw = A(1) - t.lw;
  const bool top = t.ok && (h0 >= 0), bot = t.ok && (h1 <= H - 1);
  const bool lef = (w0 >= 0), rig = (w1 <= W - 1);
  t.c1 = top && lef; t.c2 = top && rig; t.c3 = bot && lef; t.c4 = bot && rig;
  t.o1 = h0 * W + w0;
  t.o2 = t.o1 + 1;
  t.o3 = t.o1 + W;
  t.o4 = t.o3 + 1;
  return t;
}

// ---- levels owned by the tensor-memory scatter (msda_scatter_mma.cu) ------------------------------
constexpr int kMmaBlocks = 12;                 // 128-pixel fp32 accumulator blocks (32 TMEM columns each)
constexpr int kMmaMaxPixels = kMmaBlocks * 128;
constexpr int kMmaMaxLevels = 4;               // two levels per builder thread
// First level of the tail [first, L) whose pixels form the contiguous end [level_start[first], S) of the flattened value
// map and fit the accumulators; L when none does (or level_start is not the cumulative layout the reference builds,
// transformer_for_adapter.py:254-256).  Evaluated identically by msda_bwd_vec_kernel (which skips those reductions) and
// msda_scatter_mma_kernel (which owns them).
__host__ __device__ __forceinline__ int coarse_first_level(const int* sH, const int* sW, const int* sStart, int L, int S) {
  int first = L, end = S;
  for (int l = L - 1; l >= 0 && L - l <= kMmaMaxLevels; --l) {
    if (sStart[l] + sH[l] * sW[l] != end || S - sStart[l] > kMmaMaxPixels) break;
    first = l;
    end = sStart[l];
  }
  return first;
}

// ---- range plan of the second-generation tensor-memory scatter (msda_scatter_mma2.cu) ---------------
// The levels [first_level, L) are cut into RANGES: contiguous pixel intervals of the flattened value map of at most
// kR2Px pixels (kR2Blocks accumulator blocks) -- a level larger than that is split into ceil(px / kR2Px) ranges of kR2Px
// pixels, consecutive small levels at the coarse end are merged into one range.  Range ids count up from the coarse end.
// msda_bwd_vec_kernel marks, per (image, head, 64-query chunk), which ranges the chunk's samples touch (one bit each);
// msda_scatter_mma2_kernel walks (image, head, range) work units and accumulates only the chunks that hit its range.
constexpr int kR2Blocks = 6;
constexpr int kR2Px = kR2Blocks * 128;
constexpr int kR2MaxRanges = 64;
constexpr int kR2SegChunks = 16;               // chunks (of 64 queries) per work unit
struct RangePlan {
  int first_level, nranges;
  int rbase[MSDA_MAX_LEVELS];                  // id of the first range of level l (l >= first_level)
  int lo[kR2MaxRanges], hi[kR2MaxRanges];      // pixel interval [lo, hi) of range r
  int lv0[kR2MaxRanges], lv1[kR2MaxRanges];    // levels [lv0, lv1) that intersect range r
};
__host__ __device__ inline void plan_ranges(RangePlan& p, const int* sH, const int* sW, const int* sStart, int L, int S, int max_levels) {
  p.first_level = L;
  p.nranges = 0;
  int end = S, l = L - 1;
  while (l >= 0 && L - l <= max_levels) {
    const int px = sH[l] * sW[l];
    if (px <= 0 || sStart[l] + px != end) break;                 // not the cumulative layout: stop here
    if (px > kR2Px) {
      const int nr = (px + kR2Px - 1) / kR2Px;
      if (p.nranges + nr > kR2MaxRanges) break;
      p.rbase[l] = p.nranges;
      for (int k = 0; k < nr; ++k) {
        const int r = p.nranges + k;
        p.lo[r] = sStart[l] + k * kR2Px;
        p.hi[r] = (p.lo[r] + kR2Px < end) ? p.lo[r] + kR2Px : end;
        p.lv0[r] = l;
        p.lv1[r] = l + 1;
      }
      p.nranges += nr;
      p.first_level = l;
      end = sStart[l];
      --l;
    } else {
      if (p.nranges + 1 > kR2MaxRanges) break;
      // merge group: levels l, l-1, ... while they stay contiguous, small, at most 4 and within the level budget
      const int r = p.nranges, g_hi = end, g_lv1 = l + 1;
      int total = 0, cnt = 0;
      while (l >= 0 && L - l <= max_levels && cnt < 4) {
        const int q = sH[l] * sW[l];
        if (q <= 0 || sStart[l] + q != end || total + q > kR2Px) break;
        p.rbase[l] = r;
        total += q;
        end = sStart[l];
        ++cnt;
        --l;
      }
      if (cnt == 0) break;
      p.lo[r] = end; p.hi[r] = g_hi; p.lv0[r] = l + 1; p.lv1[r] = g_lv1;
      p.first_level = l + 1;
      ++p.nranges;
    }
  }
}

// The four taps of a level, computed ONCE per lane group instead of once per lane: lane `lig` of the group evaluates the
// tap of point lig & 3 and the group exchanges (corner offset, guards, lh, lw) with 16 shuffles; every lane then rebuilds
// o2..o4 / hh / hw with the same two subtractions make_tap() does, so the values are bit-identical to the unshared form.
// Needs LPG >= 4 lanes per group (all 32 lanes of the warp must call it).
template <int LPG>
__device__ __forceinline__ void shared_taps(const float (&xs)[4], const float (&ys)[4], int H, int W, int lig, Tap<float> (&t)[4]) {
  static_assert(LPG >= 4, "tap sharing needs at least four lanes per unit");
  const int pm = lig & 3;
  const float x = pm == 0 ? xs[0] : (pm == 1 ? xs[1] : (pm == 2 ? xs[2] : xs[3]));
  const float y = pm == 0 ? ys[0] : (pm == 1 ? ys[1] : (pm == 2 ? ys[2] : ys[3]));
  const Tap<float> mine = make_tap<float>(x, y, H, W);
  const int flags = (mine.c1 ? 1 : 0) | (mine.c2 ? 2 : 0) | (mine.c3 ? 4 : 0) | (mine.c4 ? 8 : 0) | (mine.ok ? 16 : 0);
  const int base = (threadIdx.x & 31) & ~(LPG - 1);     // first lane of this group
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int o1 = __shfl_sync(0xffffffffu, mine.o1, base + p);
    const int f = __shfl_sync(0xffffffffu, flags, base + p);
    t[p].lh = __shfl_sync(0xffffffffu, mine.lh, base + p);
    t[p].lw = __shfl_sync(0xffffffffu, mine.lw, base + p);
    t[p].hh = 1.f - t[p].lh;
    t[p].hw = 1.f - t[p].lw;
    t[p].o1 = o1; t[p].o2 = o1 + 1; t[p].o3 = o1 + W; t[p].o4 = o1 + W + 1;
    t[p].c1 = (f & 1) != 0; t[p].c2 = (f & 2) != 0; t[p].c3 = (f & 4) != 0; t[p].c4 = (f & 8) != 0; t[p].ok = (f & 16) != 0;
  }
}

// Unit (b, q, m) handled by lane-group `j` of pass `pass`.  A pass covers `tile` consecutive units in
// (b, q, m)-major order; with q_fast the tile is walked query-fastest so that the lane groups of a
// warp hold consecutive queries of ONE head (their corner rows often coincide on coarse levels and
// coalesce into a single 128-byte wavefront); otherwise head-fastest.
__device__ __forceinline__ long long unit_of(long long pass, int j, int tile, int M, bool q_fast) {
  if (q_fast) {
    const int qpp = tile / M;  // caller guarantees tile % M == 0 when q_fast
    j = (j % qpp) * M + (j / qpp);
  }
  return pass * tile + j;
}

}  // namespace msda
