// Tokens-side gradient of the image <-> text attention from the STORED logits gradient:  d k = scale * dS^T . q.
//
// The rows-orientation logits-gradient kernel (layer_biattn_bwd.cu) has every dS tile in shared memory anyway; with
// `store_terms` it also sends it to global memory as a 16-bit [B, H, S, Tpad] tensor (pass 0 stores its term, pass 1 adds
// its own with a TMA reduction).  The tokens-side gradient is then a plain product over the image axis instead of a second
// recomputation of logits and dP (6 logits-sized products): out[b, t, h*256 + d] = scale * sum_s dS[b, h, s, t] *
// q[b, s, h*256 + d], 1 product.  BOTH operands are consumed exactly as they lie in memory -- [s, t] and [s, d] tiles with
// the contraction index s as the row -- i.e. as MN-major A and MN-major B operands (instruction-descriptor bits 15 and 16).
// HBM-bound: dS and q are each read once (a work item covers up to 256 tokens with two accumulators fed by one q tile).
//
//   work item  (b, h, up to two 128-token tiles, one of nsplit ranges of the image axis) -> fp32 partials (or the result)
//   stage      64 image rows: dS tile [64, up to 256 tokens] (slabs of 8 KiB) + q tile [64, 256] (4 slabs)
//   warp 0 TMA producer, warp 1 MMA issuer (4 instructions of M = 128, N = 256, K = 16 per token tile and stage),
//   warps 2-5 final stage; 512 TMEM columns
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>

#include "proj_epilogue.cuh"

namespace msda {
extern long long g_launches;
}

namespace pg {
namespace bit {

constexpr int HD = 256, BM = 128, BK = 64;
constexpr int SLAB = 8192;                       // 64 rows x 128 bytes
constexpr int STAGE = 8 * SLAB, NSTAGE = 3;      // up to 4 token slabs (two 128-token tiles) + 4 head-dim slabs
constexpr int THREADS = 64 + 128;
constexpr int SMEM_BYTES = 1024 + NSTAGE * STAGE + 256;

struct TnParams {
  int B, H, S, T;
  int mtiles, mgroups;     // 128-token tiles; work items take them two at a time (one q tile feeds both accumulators)
  int ktiles, nsplit, tiles_per_split;
  float scale;
  void* out16;             // nsplit == 1: [B, T, H*256]
  float* part_o;           // nsplit > 1: [(b, h, mtile, split), 128, 256]
  int half_in;
};

template <bool HALF>
__global__ void __launch_bounds__(THREADS, 1)
biattn_tn_kernel(const __grid_constant__ CUtensorMap tmT, const __grid_constant__ CUtensorMap tmQ, TnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE);
  uint64_t* full = bars;               // [NSTAGE]
  uint64_t* empty = full + NSTAGE;     // [NSTAGE]
  uint64_t* acc_full = empty + NSTAGE;
  uint64_t* acc_free = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int items = p.B * p.H * p.mgroups * p.nsplit;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmT)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    mbar_init(acc_full, 1); mbar_init(acc_free, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t_acc = *tmem_slot;

  // work item -> (b, h, first token tile, number of token tiles (1 or 2), image-row tile range)
  auto decode = [&](int item, int& b, int& h, int& mt, int& nm, int& k0, int& n) {
    const int c = item % p.nsplit;
    int r = item / p.nsplit;
    mt = (r % p.mgroups) * 2;
    r /= p.mgroups;
    h = r % p.H;
    b = r / p.H;
    nm = p.mtiles - mt >= 2 ? 2 : 1;
    k0 = c * p.tiles_per_split;
    int k1 = k0 + p.tiles_per_split;
    if (k1 > p.ktiles) k1 = p.ktiles;
    n = k1 > k0 ? k1 - k0 : 0;
  };

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        int b, h, mt, nm, k0, n;
        decode(item, b, h, mt, nm, k0, n);
        for (int kk = 0; kk < n; ++kk) {
          const int row = (k0 + kk) * BK;
          uint8_t* st = smem + stage * STAGE;
          mbar_wait(empty + stage, ph ^ 1);
          mbar_expect_tx(full + stage, (2 * nm + 4) * SLAB);
          for (int sl = 0; sl < 2 * nm; ++sl) tma_load_4d(&tmT, full + stage, st + sl * SLAB, mt * BM + sl * 64, row, h, b);
          for (int sl = 0; sl < 4; ++sl) tma_load_3d(&tmQ, full + stage, st + (4 + sl) * SLAB, h * HD + sl * 64, row, b);
          if (++stage == NSTAGE) { stage = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc(BM, 256, HALF) | kIdescAMn | kIdescBMn;
    int stage = 0;
    uint32_t ph = 0, ph_free = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int b_, h_, mt_, nm, k0_, n;
      decode(item, b_, h_, mt_, nm, k0_, n);
      mbar_wait(acc_free, ph_free ^ 1);          // the final stage has read the previous item's accumulators
      ph_free ^= 1;
      tc_fence_after();
      for (int kk = 0; kk < n; ++kk) {
        mbar_wait(full + stage, ph);
        tc_fence_after();
        if (elect_one()) {
          const uint8_t* st = smem + stage * STAGE;
          for (int mi = 0; mi < nm; ++mi) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {          // 64 image rows, 16 per instruction
              const uint64_t da = umma_desc_mn_sw128(st + mi * 2 * SLAB, k * 2048, SLAB);
              const uint64_t db = umma_desc_mn_sw128(st + 4 * SLAB, k * 2048, SLAB);
              umma_f16(t_acc + static_cast<uint32_t>(mi * 256), da, db, idesc, (kk | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(empty + stage);
          if (kk == n - 1) umma_commit(acc_full);
        }
        __syncwarp();
        if (++stage == NSTAGE) { stage = 0; ph ^= 1; }
      }
    }
  } else {
    const int quarter = warp & 3;
    const int trow = quarter * 32 + lane;
    const uint32_t lane_bits = static_cast<uint32_t>(quarter * 32) << 16;
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int b, h, mt, nm, k0, n;
      decode(item, b, h, mt, nm, k0, n);
      const int c = item % p.nsplit;
      if (n > 0) {
        mbar_wait(acc_full, ph);
        ph ^= 1;
        tc_fence_after();
      }
      for (int mi = 0; mi < nm; ++mi) {
        const int tok = (mt + mi) * BM + trow;
        const bool live = tok < p.T;
        const size_t pitem = ((static_cast<size_t>(b) * p.H + h) * p.mtiles + (mt + mi)) * p.nsplit + c;
#pragma unroll 1
        for (int q8 = 0; q8 < 8; ++q8) {
          uint32_t r[32];
          float v[32];
          if (n > 0) tmem_ld32(t_acc + lane_bits + static_cast<uint32_t>(mi * 256 + q8 * 32), r);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = n > 0 ? __uint_as_float(r[i]) * p.scale : 0.f;
          if (p.nsplit == 1) {
            uint4 pk[4];
            pack_16(v, HALF, false, pk);
            if (live) {
              uint4* dst = reinterpret_cast<uint4*>(static_cast<uint8_t*>(p.out16) +
                                                    ((static_cast<size_t>(b) * p.T + tok) * (static_cast<size_t>(p.H) * HD) + h * HD + q8 * 32) * 2);
#pragma unroll
              for (int i = 0; i < 4; ++i) dst[i] = pk[i];
            }
          } else {
            float4* dst = reinterpret_cast<float4*>(p.part_o + (pitem * BM + trow) * HD + q8 * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_free);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(t_acc), "r"(512u) : "memory");
  }
}

}  // namespace bit
}  // namespace pg

extern "C" {

int msda_biattn_tn_splits(int S, int nsplit) {
  const int ktiles = (S + pg::bit::BK - 1) / pg::bit::BK;
  if (nsplit < 1) nsplit = 1;
  if (nsplit > ktiles) nsplit = ktiles;
  const int tps = (ktiles + nsplit - 1) / nsplit;
  return (ktiles + tps - 1) / tps;
}

// ds16: the [B, H, S, tpad] 16-bit logits gradient msda_biattn_ds_terms_16 wrote (tpad = ceil(T/64)*64); q: [B, S, H*256].
// out16 [B, T, H*256] when the image axis is not split, else part_o [B*H*ceil(T/128)*splits, 128, 256] fp32 to be summed by
// msda_biattn_combine_16(given = 1); splits = msda_biattn_tn_splits(S, nsplit).
int msda_biattn_tn_16(const void* ds16, const void* q, int B, int H, int S, int T, float scale, void* out16, float* part_o,
                      int nsplit, int is_half, void* stream) {
  using namespace pg;
  using namespace pg::bit;
  t_err[0] = 0;
  if (!ds16 || !q) { snprintf(t_err, sizeof(t_err), "null pointer"); return MSDA_ERR_NULL_POINTER; }
  if (B <= 0 || H <= 0 || S <= 0 || T <= 0) { snprintf(t_err, sizeof(t_err), "bad shape"); return MSDA_ERR_BAD_SHAPE; }
  TnParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.H = H; p.S = S; p.T = T;
  p.mtiles = (T + BM - 1) / BM;
  p.mgroups = (p.mtiles + 1) / 2;
  p.ktiles = (S + BK - 1) / BK;
  p.nsplit = msda_biattn_tn_splits(S, nsplit);
  p.tiles_per_split = (p.ktiles + p.nsplit - 1) / p.nsplit;
  p.scale = scale; p.out16 = out16; p.part_o = part_o; p.half_in = is_half;
  if (p.nsplit == 1 && !out16) { snprintf(t_err, sizeof(t_err), "null output"); return MSDA_ERR_NULL_POINTER; }
  if (p.nsplit > 1 && !part_o) { snprintf(t_err, sizeof(t_err), "null partial buffer"); return MSDA_ERR_NULL_POINTER; }
  const int dt = is_half ? 1 : 0;
  const long long tpad = static_cast<long long>((T + 63) / 64) * 64;
  CUtensorMap tmT, tmQ;
  int rc;
  if ((rc = make_map4(&tmT, ds16, B, H, S, tpad, BK, dt))) return rc;
  if ((rc = make_map3(&tmQ, q, B, S, static_cast<long long>(H) * HD, BK, dt))) return rc;
  int dev_id = 0;
  cudaGetDevice(&dev_id);
  static int sms_of[64] = {};
  if (!sms_of[dev_id & 63]) cudaDeviceGetAttribute(&sms_of[dev_id & 63], cudaDevAttrMultiProcessorCount, dev_id);
  const long long items = static_cast<long long>(B) * H * p.mgroups * p.nsplit;
  if (items >= (1ll << 31)) return MSDA_ERR_BAD_SHAPE;
  const int grid = items < sms_of[dev_id & 63] ? static_cast<int>(items) : sms_of[dev_id & 63];
  static bool configured[64] = {};
  if (!configured[dev_id & 63]) {
    cudaError_t cfg = cudaFuncSetAttribute(biattn_tn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (cfg == cudaSuccess) cfg = cudaFuncSetAttribute(biattn_tn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (cfg != cudaSuccess) { snprintf(t_err, sizeof(t_err), "cudaFuncSetAttribute: %s", cudaGetErrorString(cfg)); return static_cast<int>(cfg); }
    configured[dev_id & 63] = true;
  }
  ++msda::g_launches;
  if (is_half) biattn_tn_kernel<true><<<grid, THREADS, SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(tmT, tmQ, p);
  else biattn_tn_kernel<false><<<grid, THREADS, SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(tmT, tmQ, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(t_err, sizeof(t_err), "biattn_tn_kernel launch: %s", cudaGetErrorString(e)); return static_cast<int>(e); }
  return 0;
}

}  // extern "C"
