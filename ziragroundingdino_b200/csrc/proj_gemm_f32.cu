// fp32 projections of the MultiScaleDeformableAttention module on tcgen05 -- three TF32 products per tile.
//
// The reference trains with AMP off (groundingdino/config/configs/common/train.py:11), so its four nn.Linear calls
// (ms_deform_attn.py:286-350) are fp32 GEMMs and north_star's bars for that mode are 1e-5 forward / 1e-4 backward.  One
// kind::tf32 product (10-bit mantissas) cannot meet them; three can:
//
//     x = x_hi + x_lo,  w = w_hi + w_lo      (hi = the upper 19 bits of the fp32 pattern, lo = the exact remainder)
//     x w^T  ~=  x_hi w_hi^T + x_lo w_hi^T + x_hi w_lo^T          (dropped: x_lo w_lo^T, 2^-22 relative)
//
// accumulated in fp32 in tensor memory: ~1e-6 relative, i.e. fp32-GEMM accuracy at half the bf16 tensor rate / 3 -- still
// ~3x the FFMA GEMM the library runs for fp32 with TF32 disabled (torch's default).  The weight split is done once on the
// host side (W is passed as [w_hi; w_lo], 2*Nout rows); the activation split happens in shared memory: warps 2-5 turn each
// TMA-loaded fp32 tile into its hi tile (in place) and a lo tile, element-wise at identical offsets, so the 128-byte
// swizzle never has to be decoded.
//
//   warp 0      TMA producer (A 128 x 32 fp32, B_hi and B_lo 64 x 32 fp32 per stage; 4 stages of 48 KiB)
//   warp 1      MMA issuer: 4 k-steps x 3 tcgen05.mma.kind::tf32 (M = 128, N = 64, K = 8) per stage; owns TMEM (2 x 64 columns)
//   warps 2-5   splitter (see above)
//   warps 6-9   epilogue: tcgen05.ld -> bias / row mask / accumulate, or sampling locations + softmax (the query
//               projection), -> 128B-swizzled tile -> TMA store
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>

#include "proj_epilogue.cuh"

namespace msda {
extern long long g_launches;
}

namespace pg {
namespace f32 {

constexpr int BM = 128, BK = 32, BN = 64, STAGES = 4;
constexpr int A_TILE = BM * 128, B_TILE = BN * 128;
constexpr int STAGE_BYTES = 2 * A_TILE + 2 * B_TILE;
constexpr int SPLIT_WARPS = 4, EPI_WARPS = 4;
constexpr int THREADS = 64 + 32 * (SPLIT_WARPS + EPI_WARPS);
constexpr int MAX_NOUT = 768;
constexpr int AUX_BYTES = 256 + MAX_NOUT * 4 + MSDA_MAX_LEVELS * 8;
constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + EPI_WARPS * 4096 + AUX_BYTES;

__host__ __device__ inline uint32_t idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1)
linear_tf32x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2, int R, int Nout, int K,
                     EpiParams ep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* tiles = smem + STAGES * STAGE_BYTES;            // per epilogue warp: one 32 x 128-byte store tile
  uint8_t* aux = tiles + EPI_WARPS * 4096;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
  uint64_t* split_bar = full_bar + STAGES;
  uint64_t* empty_bar = split_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* s_bias = reinterpret_cast<float*>(aux + 256);
  float* s_norm = s_bias + MAX_NOUT;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k = K / BK, num_n = Nout / BN, num_m = (R + BM - 1) / BM;
  const int t_end = num_m * num_n;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    for (int i = 0; i < STAGES; ++i) { mbar_init(full_bar + i, 1); mbar_init(split_bar + i, SPLIT_WARPS * 32); mbar_init(empty_bar + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull_bar + i, 1); mbar_init(tempty_bar + i, EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < Nout; i += THREADS) s_bias[i] = ep.bias ? ep.bias[i] : 0.f;
  if (MODE == EPI_QUERY && threadIdx.x < ep.L) {
    s_norm[2 * threadIdx.x] = 1.f / static_cast<float>(ep.shapes[2 * threadIdx.x + 1]);      // 1/W
    s_norm[2 * threadIdx.x + 1] = 1.f / static_cast<float>(ep.shapes[2 * threadIdx.x]);      // 1/H
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < t_end; t += gridDim.x) {
        const int m_idx = (t / num_n) * BM, n_idx = (t % num_n) * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(empty_bar + stage, phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          mbar_expect_tx(full_bar + stage, static_cast<uint32_t>(A_TILE + 2 * B_TILE));
          tma_load_2d(&tmA, full_bar + stage, sa, kb * BK, m_idx);
          tma_load_2d(&tmB, full_bar + stage, sa + 2 * A_TILE, kb * BK, n_idx);                  // w_hi rows
          tma_load_2d(&tmB, full_bar + stage, sa + 2 * A_TILE + B_TILE, kb * BK, Nout + n_idx);   // w_lo rows
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = idesc_tf32(BM, BN);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int t = blockIdx.x; t < t_end; t += gridDim.x) {
      mbar_wait(tempty_bar + acc, acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(split_bar + stage, phase);
        tc_fence_after();
        if (elect_one()) {
          const uint8_t* sa = smem + stage * STAGE_BYTES;
          const uint8_t* sb = sa + 2 * A_TILE;
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t a_hi = umma_desc_sw128(sa, k * 32), a_lo = umma_desc_sw128(sa + A_TILE, k * 32);
            const uint64_t b_hi = umma_desc_sw128(sb, k * 32), b_lo = umma_desc_sw128(sb + B_TILE, k * 32);
            umma_tf32(tmem_d, a_lo, b_hi, idesc, (kb | k) != 0 ? 1u : 0u);     // small terms first
            umma_tf32(tmem_d, a_hi, b_lo, idesc, 1u);
            umma_tf32(tmem_d, a_hi, b_hi, idesc, 1u);
          }
          umma_commit(empty_bar + stage);
          if (kb == num_k - 1) umma_commit(tfull_bar + acc);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp < 2 + SPLIT_WARPS) {
    // ===== splitter: fp32 tile -> hi (in place) + lo, same offsets =====
    const int t128 = threadIdx.x - 64;
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < t_end; t += gridDim.x) {
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(full_bar + stage, phase);
        uint4* hi = reinterpret_cast<uint4*>(smem + stage * STAGE_BYTES);
        uint4* lo = reinterpret_cast<uint4*>(smem + stage * STAGE_BYTES + A_TILE);
#pragma unroll
        for (int i = 0; i < A_TILE / 16 / (SPLIT_WARPS * 32); ++i) {
          const int idx = t128 + i * SPLIT_WARPS * 32;
          const uint4 v = hi[idx];
          const uint4 h = make_uint4(v.x & 0xffffe000u, v.y & 0xffffe000u, v.z & 0xffffe000u, v.w & 0xffffe000u);
          hi[idx] = h;
          lo[idx] = make_uint4(__float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x)), __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y)),
                               __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z)), __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w)));
        }
        fence_async_smem();
        mbar_arrive(split_bar + stage);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===== epilogue: one warp per TMEM lane quarter, two 32-column chunks per tile =====
    const int ew = warp - 2 - SPLIT_WARPS;
    const int quarter = warp & 3;
    uint8_t* tile = tiles + ew * 4096;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < t_end; t += gridDim.x) {
      const int m_idx = (t / num_n) * BM, n_idx = (t % num_n) * BN;
      mbar_wait(tfull_bar + acc, acc_phase);
      tc_fence_after();
      const long long row0 = static_cast<long long>(m_idx) + quarter * 32;
      const long long row = row0 + lane;
      const bool live = row < R;
      const bool zero = live && ep.row_mask != nullptr && ep.row_mask[row] != 0;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * BN);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        const int gc = n_idx + c0;
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        float v[32], o32[32];
        const float4* bp = reinterpret_cast<const float4*>(s_bias + gc);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 bb = bp[i];
          v[4 * i] = __uint_as_float(r[4 * i]) + bb.x; v[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + bb.y;
          v[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + bb.z; v[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + bb.w;
        }
        bool is_loc = false;
        if (MODE == EPI_STORE) {
          if (ep.accum != nullptr && live) {
            const float4* ap = reinterpret_cast<const float4*>(static_cast<const float*>(ep.accum) + row * ep.out_ld + gc);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 a = ap[i];
              v[4 * i] += a.x; v[4 * i + 1] += a.y; v[4 * i + 2] += a.z; v[4 * i + 3] += a.w;
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) o32[j] = zero ? 0.f : v[j];
        } else {
          is_loc = gc < ep.n_loc;
          if (is_loc) {
            const long long rr = live ? row : 0;
            if (ep.L == 4 && ep.P == 4) {
              if (ep.ref_dim == 2) epi_loc_l4p4<2>(ep, v, rr, s_norm, o32);
              else epi_loc_l4p4<4>(ep, v, rr, s_norm, o32);
            } else {
              epi_loc(ep, v, rr, gc, s_norm, o32);
            }
          } else {
            epi_softmax(ep, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) o32[j] = v[j];
          }
        }
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<float4*>(swz(tile, lane, i)) = make_float4(o32[4 * i], o32[4 * i + 1], o32[4 * i + 2], o32[4 * i + 3]);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (MODE == EPI_STORE || is_loc) tma_store_2d(&tmC, tile, gc, static_cast<int>(row0));
          else tma_store_2d(&tmC2, tile, gc - ep.n_loc, static_cast<int>(row0));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

static int launch(const float* x, const float* w_split, long long R, int K, int Nout, const EpiParams& ep, cudaStream_t st) {
  if (!x || !w_split) { snprintf(t_err, sizeof(t_err), "null operand"); return MSDA_ERR_NULL_POINTER; }
  if (R <= 0 || R >= (1ll << 31) || K <= 0 || K % BK || Nout <= 0 || Nout % BN || Nout > MAX_NOUT) {
    snprintf(t_err, sizeof(t_err), "unsupported fp32 GEMM shape R=%lld K=%d Nout=%d (K %% 32, Nout %% 64, Nout <= %d)", R, K, Nout, MAX_NOUT);
    return MSDA_ERR_UNSUPPORTED;
  }
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_split)) & 15u) {
    snprintf(t_err, sizeof(t_err), "operands must be 16-byte aligned");
    return MSDA_ERR_MISALIGNED;
  }
  CUtensorMap tmA, tmB, tmC, tmC2;
  int rc = make_map(&tmA, x, R, K, BM, BK, 2);
  if (rc) return rc;
  rc = make_map(&tmB, w_split, 2ll * Nout, K, BN, BK, 2);
  if (rc) return rc;
  if (ep.mode == EPI_QUERY) {
    rc = make_map(&tmC, ep.loc_out, R, ep.n_loc, 32, 32, 2);
    if (rc) return rc;
    rc = make_map(&tmC2, ep.aw_out, R, ep.n_aw, 32, 32, 2);
  } else {
    if ((reinterpret_cast<uintptr_t>(ep.out) & 15u) || ep.out_ld % 4) { snprintf(t_err, sizeof(t_err), "fp32 output must be 16-byte aligned with ld %% 4 == 0"); return MSDA_ERR_MISALIGNED; }
    rc = make_map(&tmC, ep.out, R, ep.out_ld, 32, 32, 2);
    tmC2 = tmC;
  }
  if (rc) return rc;
  int dev_id = 0;
  cudaGetDevice(&dev_id);
  static int sms_of[64] = {};
  if (!sms_of[dev_id & 63]) cudaDeviceGetAttribute(&sms_of[dev_id & 63], cudaDevAttrMultiProcessorCount, dev_id);
  const long long tiles = ((R + BM - 1) / BM) * (Nout / BN);
  const int grid = static_cast<int>(tiles < sms_of[dev_id & 63] ? tiles : sms_of[dev_id & 63]);
  const int Ri = static_cast<int>(R);
  cudaError_t cfg = cudaSuccess;
  ++msda::g_launches;
#define F32_LAUNCH(MODE)                                                                                              \
  do {                                                                                                                 \
    static bool configured[64] = {};                                                                                   \
    if (!configured[dev_id & 63]) {                                                                                    \
      cfg = cudaFuncSetAttribute(linear_tf32x3_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES); \
      configured[dev_id & 63] = cfg == cudaSuccess;                                                                    \
    }                                                                                                                  \
    if (cfg == cudaSuccess) linear_tf32x3_kernel<MODE><<<grid, THREADS, SMEM_BYTES, st>>>(tmA, tmB, tmC, tmC2, Ri, Nout, K, ep); \
  } while (0)
  if (ep.mode == EPI_QUERY) F32_LAUNCH(EPI_QUERY); else F32_LAUNCH(EPI_STORE);
#undef F32_LAUNCH
  if (cfg != cudaSuccess) { snprintf(t_err, sizeof(t_err), "cudaFuncSetAttribute: %s", cudaGetErrorString(cfg)); return static_cast<int>(cfg); }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(t_err, sizeof(t_err), "linear_tf32x3_kernel launch: %s", cudaGetErrorString(e)); return static_cast<int>(e); }
  return 0;
}

}  // namespace f32
}  // namespace pg

extern "C" {

int msda_linear_f32(const float* x, const float* w_split, const float* bias, long long R, int K, int Nout, const float* accum,
                    float* out, int out_ld, const uint8_t* row_mask, void* stream) {
  pg::t_err[0] = 0;
  if (!out) { snprintf(pg::t_err, sizeof(pg::t_err), "null output"); return MSDA_ERR_NULL_POINTER; }
  pg::EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.mode = pg::EPI_STORE;
  ep.out = out; ep.out_ld = out_ld; ep.out_f32 = 1; ep.bias = bias; ep.row_mask = row_mask; ep.accum = accum;
  return pg::f32::launch(x, w_split, R, K, Nout, ep, static_cast<cudaStream_t>(stream));
}

int msda_query_proj_f32(const float* query, const float* w_cat_split, const float* bias_cat, const float* ref, int ref_dim,
                        const int64_t* spatial_shapes, long long R, int K, int M, int L, int P, float* loc_out, float* aw_out,
                        void* stream) {
  pg::t_err[0] = 0;
  if (!ref || !spatial_shapes || !loc_out || !aw_out) { snprintf(pg::t_err, sizeof(pg::t_err), "null pointer"); return MSDA_ERR_NULL_POINTER; }
  const int n_aw = M * L * P, n_loc = 2 * n_aw, lp = L * P;
  if ((ref_dim != 2 && ref_dim != 4) || L > MSDA_MAX_LEVELS || (3 * n_aw) % pg::f32::BN || 3 * n_aw > pg::f32::MAX_NOUT || n_aw % 32) {
    snprintf(pg::t_err, sizeof(pg::t_err), "fused fp32 query projection needs M*L*P %% 64 == 0 and 3*M*L*P <= %d (L=%d P=%d M=%d)", pg::f32::MAX_NOUT, L, P, M);
    return MSDA_ERR_UNSUPPORTED;
  }
  pg::EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.mode = pg::EPI_QUERY;
  ep.bias = bias_cat; ep.loc_out = loc_out; ep.aw_out = aw_out; ep.ref = ref; ep.shapes = spatial_shapes;
  ep.ref_dim = ref_dim; ep.L = L; ep.P = P; ep.n_loc = n_loc; ep.n_aw = n_aw;
  int rc = pg::f32::launch(query, w_cat_split, R, K, n_loc + n_aw, ep, static_cast<cudaStream_t>(stream));
  if (rc == 0 && 32 % lp != 0) rc = pg::softmax_rows(aw_out, R * M, lp, static_cast<cudaStream_t>(stream));
  return rc;
}

}  // extern "C"
