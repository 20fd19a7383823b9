// Image <-> text bidirectional attention on tcgen05.
//
// The reference's BiMultiHeadAttention (fuse_modules.py:146-250; called once per encoder layer,
// transformer_for_adapter.py:578-593) builds the [B*heads, n_img, n_text] logits in fp32, transposes them, and runs two
// softmaxes and two batched products over them (728 MB of attention matrices at 4 images).  Both directions are softmax
// attention over the SAME logits S = q k^T (head_dim 256):
//
//     image <- text :  O_v[s, :] = sum_t softmax_t(S[s, t]) val_l[t, :]
//     text <- image :  O_l[t, :] = sum_s softmax_s(S[s, t]) val_v[s, :]
//
// so one kernel serves both, run in two ORIENTATIONS: a CTA keeps a 128-row "stationary" tile A (TMEM lanes = its rows)
// and streams 128-row tiles of B and X past it,
//
//   GEMM1_j  acc1[j & 1] (TMEM, 128 columns) = A[128, 256] . B_j[128, 256]^T                (logits, fp32)
//   mid      registers <- acc1: scale, mask, p = exp2(s - reference); P_j (16 bit) -> shared memory, K-major, swizzled
//   GEMM2_j  acc2 (TMEM, 256 columns)       += P_j[128, 128] . X_j[128, 256]
//
// rows orientation:   A = q tile, B = k, X = val_l   (2 column tiles; result normalised and stored directly)
// tokens orientation: A = k half, B = q tiles, X = val_v (174 column tiles, split across CTAs; partial results + combine)
//
// X_j is consumed straight from the row-major [rows, head_dim] tile TMA delivers: it is the B operand of GEMM2 in
// MN-major form (instruction-descriptor bit 16), so no transposed copy of the values exists anywhere.  Heads are
// addressed inside the [B, L, heads*256] projections by the tensor map's column coordinate: no head-major copies either.
//
// Two softmax modes.  ONLINE (forward): the softmax runs over the streamed columns; each lane keeps a running reference
// maximum that is only raised (with the accumulator rescaled in TMEM) when a tile exceeds it by more than 2^8, and the
// log2-domain log-sum-exp per lane is saved.  GIVEN (backward, dVal = P^T dO): the statistics of the OTHER direction are
// supplied per column, p = exp2(s - stat[column]) exactly reproduces that direction's probabilities, transposed.
//
//   warp 0      TMA producer: the stationary tile once per work item, then B / X k-blocks through a ring of 16 KiB slots
//   warp 1      MMA issuer (GEMM1_{j+1} is issued before GEMM2_j); owns the 512 TMEM columns (2 x 128 + 256)
//   warps 2-9   mid / final stages: two warps per TMEM lane quarter, each 64 of a tile's 128 columns
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>

#include "proj_epilogue.cuh"

namespace msda {
extern long long g_launches;
}

namespace pg {
namespace bia {

constexpr int HD = 256;                  // head dimension: K of GEMM1, N of GEMM2
constexpr int BM = 128, BN = 128;        // stationary rows per work item; streamed rows per column tile
constexpr int A_BYTES = 4 * 16384;       // 4 k-blocks of 128 rows x 128 bytes
constexpr int SLOT = 16384;
constexpr int P_BYTES = 2 * 16384;       // one probability tile: 2 k-blocks of 128 rows x 128 bytes
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr float kRaise = 8.f;            // raise the running reference only when a tile exceeds it by 2^8

struct PvParams {
  int B, H, LA, LB;            // batch, heads, stationary length, streamed length
  int mtiles, ctiles;          // ceil(LA / 128), ceil(LB / 128)
  int nsplit, tiles_per_split; // column tiles are split over nsplit work items per (b, h, mtile)
  int given;                   // 0: online softmax over the streamed columns; 1: per-column statistics supplied
  float scale_log2;            // logits scale x log2(e)
  const uint8_t* mask;         // online: [B, ctiles*128] 1 = column masked (padding columns are 1)
                               // given : [B, mtiles*128] 1 = lane masked (padding lanes are 1)
  const float* col_stat;       // given: [B, H, ctiles*128] log2-domain log-sum-exp of each column's softmax (over the
                               // lanes); +inf in the padding (and for columns that attend to nothing)
  float* lane_stat;            // online, nsplit == 1: [B, H, mtiles*128] log2-domain log-sum-exp per lane (may be null)
  float* part_o;               // nsplit > 1: [items, 128, 256] fp32 un-normalised partial results
  float* part_m;               // nsplit > 1, online: [items, 128] running reference;  part_l: [items, 128] partial sums
  float* part_l;
  int half_in;
};

template <int NSLOT>
struct PvSmem {
  static constexpr int kBytes = 1024 + A_BYTES + NSLOT * SLOT + 2 * P_BYTES + 2 * 2 * BM * 4 + 2 * BM * 4 + 512;
};

template <int NSLOT>
__global__ void __launch_bounds__(THREADS, 1)
biattn_pv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmOut, PvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sRing = sA + A_BYTES;
  uint8_t* sP = sRing + NSLOT * SLOT;
  float* sMax = reinterpret_cast<float*>(sP + 2 * P_BYTES);      // [2 buffers][2 halves][128]
  float* sSum = sMax + 2 * 2 * BM;                                 // [2 halves][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sSum + 2 * BM);
  uint64_t* full = bars;              // [NSLOT]
  uint64_t* empty = full + NSLOT;     // [NSLOT]
  uint64_t* x_full = empty + NSLOT;   // stationary tile landed
  uint64_t* x_free = x_full + 1;      // last GEMM1 of the work item retired
  uint64_t* a1_full = x_free + 1;     // [2] logits tile complete
  uint64_t* a1_free = a1_full + 2;    // [2] mid stage has read it
  uint64_t* h_full = a1_free + 2;     // [2] probability tile written to shared memory
  uint64_t* h_free = h_full + 2;      // [2] GEMM2 has read it (== GEMM2 of that tile complete)
  uint64_t* a2_full = h_free + 2;     // result complete
  uint64_t* a2_free = a2_full + 1;    // final stage has read it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a2_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int items = p.B * p.H * p.mtiles * p.nsplit;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
    for (int i = 0; i < NSLOT; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    mbar_init(x_full, 1); mbar_init(x_free, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(a1_full + i, 1); mbar_init(a1_free + i, EPI_WARPS); mbar_init(h_full + i, EPI_WARPS); mbar_init(h_free + i, 1); }
    mbar_init(a2_full, 1); mbar_init(a2_free, EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_acc1 = tmem_base, t_acc2 = tmem_base + 256u;

  // work item -> (b, h, stationary tile, column-tile range)
  auto decode = [&](int item, int& b, int& h, int& mt, int& j0, int& n) {
    const int c = item % p.nsplit;
    int r = item / p.nsplit;
    mt = r % p.mtiles;
    r /= p.mtiles;
    h = r % p.H;
    b = r / p.H;
    j0 = c * p.tiles_per_split;
    int j1 = j0 + p.tiles_per_split;
    if (j1 > p.ctiles) j1 = p.ctiles;
    n = j1 > j0 ? j1 - j0 : 0;
  };

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      int slot = 0;
      uint32_t ph_slot = 0, ph_x = 0;
      auto next_slot = [&]() { if (++slot == NSLOT) { slot = 0; ph_slot ^= 1; } };
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        int b, h, mt, j0, n;
        decode(item, b, h, mt, j0, n);
        mbar_wait(x_free, ph_x ^ 1);
        ph_x ^= 1;
        mbar_expect_tx(x_full, A_BYTES);
        for (int kb = 0; kb < 4; ++kb) tma_load_3d(&tmA, x_full, sA + kb * 16384, h * HD + kb * 64, mt * BM, b);
        for (int jj = 0; jj <= n; ++jj) {
          if (jj < n) {
            for (int kb = 0; kb < 4; ++kb) {
              mbar_wait(empty + slot, ph_slot ^ 1);
              mbar_expect_tx(full + slot, SLOT);
              tma_load_3d(&tmB, full + slot, sRing + slot * SLOT, h * HD + kb * 64, (j0 + jj) * BN, b);
              next_slot();
            }
          }
          if (jj >= 1) {
            for (int sl = 0; sl < 4; ++sl) {
              mbar_wait(empty + slot, ph_slot ^ 1);
              mbar_expect_tx(full + slot, SLOT);
              tma_load_3d(&tmX, full + slot, sRing + slot * SLOT, h * HD + sl * 64, (j0 + jj - 1) * BN, b);
              next_slot();
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc1 = umma_idesc(BM, BN, p.half_in != 0);
    const uint32_t idesc2 = umma_idesc(BM, 64, p.half_in != 0) | kIdescBMn;
    int slot = 0;
    uint32_t ph_slot = 0, ph_x = 0, ph_a2 = 0, g1 = 0, g2 = 0;     // g1 / g2: running tile counters of GEMM1 / GEMM2
    uint32_t ph_a1free[2] = {0, 0}, ph_hfull[2] = {0, 0};
    auto next_slot = [&]() { if (++slot == NSLOT) { slot = 0; ph_slot ^= 1; } };
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int b_, h_, mt_, j0_, n;
      decode(item, b_, h_, mt_, j0_, n);
      mbar_wait(x_full, ph_x);
      ph_x ^= 1;
      tc_fence_after();
      for (int jj = 0; jj <= n; ++jj) {
        if (jj < n) {
          const int b = g1 & 1;
          ++g1;
          mbar_wait(a1_free + b, ph_a1free[b] ^ 1);
          ph_a1free[b] ^= 1;
          tc_fence_after();
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(full + slot, ph_slot);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t da = umma_desc_sw128(sA + kb * 16384, k * 32);
                const uint64_t db = umma_desc_sw128(sRing + slot * SLOT, k * 32);
                umma_f16(t_acc1 + static_cast<uint32_t>(b * BN), da, db, idesc1, (kb | k) != 0 ? 1u : 0u);
              }
              umma_commit(empty + slot);
              if (kb == 3) {
                umma_commit(a1_full + b);
                if (jj == n - 1) umma_commit(x_free);
              }
            }
            __syncwarp();
            next_slot();
          }
        }
        if (jj >= 1) {
          const int t = jj - 1, b = g2 & 1;
          ++g2;
          if (t == 0) {                        // first product into acc2: the previous item's result must have been read out
            mbar_wait(a2_free, ph_a2 ^ 1);
            ph_a2 ^= 1;
          }
          mbar_wait(h_full + b, ph_hfull[b]);
          ph_hfull[b] ^= 1;
          tc_fence_after();
          for (int sl = 0; sl < 4; ++sl) {
            mbar_wait(full + slot, ph_slot);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {    // K = 128 streamed rows, 16 per instruction
                const uint64_t da = umma_desc_sw128(sP + b * P_BYTES + (k >> 2) * 16384, (k & 3) * 32);
                const uint64_t db = umma_desc_mn_sw128(sRing + slot * SLOT, k * 2048, 16384u);
                umma_f16(t_acc2 + static_cast<uint32_t>(sl * 64), da, db, idesc2, (t | k) != 0 ? 1u : 0u);
              }
              umma_commit(empty + slot);
              if (sl == 3) {
                umma_commit(h_free + b);
                if (t == n - 1) umma_commit(a2_full);
              }
            }
            __syncwarp();
            next_slot();
          }
        }
      }
    }
  } else {
    // ===== mid / final stages =====
    const int quarter = warp & 3, half = (warp - 2) >> 2;     // TMEM lane quarter; which 64 of a tile's 128 columns
    const int trow = quarter * 32 + lane;                     // row inside the stationary tile == TMEM lane
    const uint32_t lane_bits = static_cast<uint32_t>(quarter * 32) << 16;
    uint32_t ph_a1full[2] = {0, 0}, ph_hfree[2] = {0, 0}, ph_a2 = 0, g = 0;
    uint8_t* stile = sP + (half * 4 + quarter) * 4096;        // the 4 KiB of P buffer 0 only this warp ever writes
    const int lb_pad = p.ctiles * BN, la_pad = p.mtiles * BM;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int b, h, mt, j0, n;
      decode(item, b, h, mt, j0, n);
      const int arow = mt * BM + trow;                        // row of the stationary operand
      const bool live = arow < p.LA;
      bool lane_ok = live;
      if (p.given && p.mask != nullptr) lane_ok = lane_ok && p.mask[static_cast<size_t>(b) * la_pad + arow] == 0;
      float m_ref = -CUDART_INF_F, l_part = 0.f;
      for (int jj = 0; jj < n; ++jj, ++g) {
        const int bb = g & 1;
        mbar_wait(a1_full + bb, ph_a1full[bb]);
        ph_a1full[bb] ^= 1;
        tc_fence_after();
        uint32_t r0[32], r1[32];
        const uint32_t t_s = t_acc1 + lane_bits + static_cast<uint32_t>(bb * BN + half * 64);
        tmem_ld32(t_s, r0);
        tmem_ld32(t_s + 32u, r1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a1_free + bb);             // logits are in registers: GEMM1 of tile g + 2 may overwrite them
        const int col0 = (j0 + jj) * BN + half * 64;          // first streamed row (= logits column) this thread handles
        float s[64];
        float ref_use;
        if (!p.given) {
          const uint4* mp = reinterpret_cast<const uint4*>(p.mask + static_cast<size_t>(b) * lb_pad + col0);
          float tmax = -CUDART_INF_F;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const uint4 mk = __ldg(mp + q4);
            const uint32_t w[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int c = q4 * 16 + i;
              const float raw = __uint_as_float(c < 32 ? r0[c & 31] : r1[c & 31]);
              const bool masked = ((w[i >> 2] >> ((i & 3) * 8)) & 0xffu) != 0u;
              s[c] = masked ? -CUDART_INF_F : raw * p.scale_log2;
              tmax = fmaxf(tmax, s[c]);
            }
          }
          float* mx = sMax + bb * 2 * BM;
          mx[half * BM + trow] = tmax;
          named_bar(1 + quarter, 64);
          tmax = fmaxf(tmax, mx[(half ^ 1) * BM + trow]);
          const bool need = tmax > m_ref + kRaise;            // also true for the first finite tile (m_ref = -inf)
          float sc = 1.f;
          if (need) {
            sc = ex2(m_ref - tmax);                            // 0 when m_ref = -inf
            l_part *= sc;
            m_ref = tmax;
          }
          if (jj > 0 && __any_sync(0xffffffffu, need)) {
            // raise the reference: rescale this warp's half of the accumulator once GEMM2 of the previous tile has retired
            const int pb = (g - 1) & 1;
            mbar_wait(h_free + pb, ph_hfree[pb] ^ 1);          // peek: the toggle happens at tile g + 1 as usual
            tc_fence_after();
#pragma unroll 1
            for (int q4 = 0; q4 < 4; ++q4) {
              uint32_t o[32];
              const uint32_t t_o = t_acc2 + lane_bits + static_cast<uint32_t>(half * 128 + q4 * 32);
              tmem_ld32(t_o, o);
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * sc);
              tmem_st32(t_o, o);
            }
            tmem_st_wait();
          }
          ref_use = m_ref == -CUDART_INF_F ? 0.f : m_ref;
        } else {
          const float4* cs = reinterpret_cast<const float4*>(p.col_stat + (static_cast<size_t>(b) * p.H + h) * lb_pad + col0);
#pragma unroll
          for (int c4 = 0; c4 < 16; ++c4) {
            const float4 st = __ldg(cs + c4);
            const float stv[4] = {st.x, st.y, st.z, st.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int c = c4 * 4 + i;
              const float raw = __uint_as_float(c < 32 ? r0[c & 31] : r1[c & 31]);
              s[c] = lane_ok ? raw * p.scale_log2 - stv[i] : -CUDART_INF_F;
            }
          }
          ref_use = 0.f;
        }
        mbar_wait(h_free + bb, ph_hfree[bb] ^ 1);             // GEMM2 of tile g - 2 has finished reading this P buffer
        ph_hfree[bb] ^= 1;
        uint8_t* pk_base = sP + bb * P_BYTES + half * 16384;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v[i] = ex2(s[hf * 32 + i] - ref_use);
            l_part += v[i];
          }
          uint4 pk[4];
          pack_16(v, p.half_in != 0, false, pk);
#pragma unroll
          for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(swz(pk_base, trow, 4 * hf + i)) = pk[i];
        }
        tc_fence_before();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(h_full + bb);
      }
      // ---- final stage ----
      if (n > 0) {
        mbar_wait(a2_full, ph_a2);
        ph_a2 ^= 1;
        tc_fence_after();
      }
      sSum[half * BM + trow] = l_part;
      named_bar(1 + quarter, 64);
      const float l_tot = l_part + sSum[(half ^ 1) * BM + trow];
      named_bar(1 + quarter, 64);
      if (p.nsplit == 1) {
        const float inv = p.given ? 1.f : (l_tot > 0.f ? 1.f / l_tot : 0.f);
        if (!p.given && p.lane_stat != nullptr && half == 0 && live)
          p.lane_stat[(static_cast<size_t>(b) * p.H + h) * la_pad + arow] = l_tot > 0.f ? m_ref + log2f(l_tot) : CUDART_INF_F;
#pragma unroll 1
        for (int gq = 0; gq < 2; ++gq) {
          const int gc = half * 128 + gq * 64;
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t r[32];
            float v[32];
            if (n > 0) tmem_ld32(t_acc2 + lane_bits + static_cast<uint32_t>(gc + 32 * hf), r);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = n > 0 ? __uint_as_float(r[i]) * inv : 0.f;
            uint4 pk[4];
            pack_16(v, p.half_in != 0, false, pk);
#pragma unroll
            for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(swz(stile, lane, 4 * hf + i)) = pk[i];
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) tma_store_3d(&tmOut, stile, h * HD + gc, mt * BM + quarter * 32, b);
        }
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
      } else {
        float* po = p.part_o + (static_cast<size_t>(item) * BM + trow) * HD + half * 128;
#pragma unroll 1
        for (int q4 = 0; q4 < 4; ++q4) {
          uint32_t r[32];
          if (n > 0) tmem_ld32(t_acc2 + lane_bits + static_cast<uint32_t>(half * 128 + q4 * 32), r);
          float4* dst = reinterpret_cast<float4*>(po + q4 * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            dst[i] = n > 0 ? make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (!p.given && half == 0) {
          p.part_m[static_cast<size_t>(item) * BM + trow] = m_ref;
          p.part_l[static_cast<size_t>(item) * BM + trow] = l_tot;
        }
      }
      if (n > 0) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a2_free);
      }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// Merge the per-split partial results of one (b, h, stationary tile): flash-decoding style for the online mode
// (out = sum_c o_c 2^(m_c - m*) / sum_c l_c 2^(m_c - m*)), a plain sum for the given-statistics mode.
__global__ void biattn_combine_kernel(const float* __restrict__ part_o, const float* __restrict__ part_m,
                                      const float* __restrict__ part_l, int H, int LA, int mtiles, int nsplit, int given,
                                      void* __restrict__ out16, float* __restrict__ lane_stat, int half_out) {
  const int row = blockIdx.x;                 // over B * H * LA
  const int a = row % LA, bh = row / LA, h = bh % H, b = bh / H;
  const int mt = a / BM, tr = a % BM;
  const size_t item0 = (static_cast<size_t>(bh) * mtiles + mt) * nsplit;
  const int d = threadIdx.x;                  // 256 threads: one per head-dim column
  float acc = 0.f, inv = 1.f;
  if (!given) {
    float mstar = -CUDART_INF_F;
    for (int c = 0; c < nsplit; ++c) mstar = fmaxf(mstar, part_m[(item0 + c) * BM + tr]);
    float lstar = 0.f;
    for (int c = 0; c < nsplit; ++c) {
      const float mc = part_m[(item0 + c) * BM + tr];
      const float w = mc == -CUDART_INF_F ? 0.f : exp2f(mc - mstar);
      lstar += part_l[(item0 + c) * BM + tr] * w;
      acc += part_o[((item0 + c) * BM + tr) * HD + d] * w;
    }
    inv = lstar > 0.f ? 1.f / lstar : 0.f;
    if (d == 0 && lane_stat != nullptr) lane_stat[static_cast<size_t>(bh) * mtiles * BM + a] = lstar > 0.f ? mstar + log2f(lstar) : CUDART_INF_F;
  } else {
    for (int c = 0; c < nsplit; ++c) acc += part_o[((item0 + c) * BM + tr) * HD + d];
  }
  const float v = acc * inv;
  const size_t o = (static_cast<size_t>(b) * LA + a) * (static_cast<size_t>(H) * HD) + h * HD + d;
  if (half_out) static_cast<__half*>(out16)[o] = __float2half_rn(v);
  else static_cast<__nv_bfloat16*>(out16)[o] = __float2bfloat16_rn(v);
}

constexpr int kSlots = 5;

static int launch_pv(const void* a, const void* b, const void* x, void* out16, PvParams p, cudaStream_t st) {
  if (!a || !b || !x) { snprintf(t_err, sizeof(t_err), "null operand"); return MSDA_ERR_NULL_POINTER; }
  if (p.B <= 0 || p.H <= 0 || p.LA <= 0 || p.LB <= 0) { snprintf(t_err, sizeof(t_err), "bad shape"); return MSDA_ERR_BAD_SHAPE; }
  p.mtiles = (p.LA + BM - 1) / BM;
  p.ctiles = (p.LB + BN - 1) / BN;
  if (p.nsplit < 1) p.nsplit = 1;
  if (p.nsplit > p.ctiles) p.nsplit = p.ctiles;
  p.tiles_per_split = (p.ctiles + p.nsplit - 1) / p.nsplit;
  p.nsplit = (p.ctiles + p.tiles_per_split - 1) / p.tiles_per_split;     // no empty splits
  if (p.nsplit == 1 && !out16) { snprintf(t_err, sizeof(t_err), "null output"); return MSDA_ERR_NULL_POINTER; }
  if (p.nsplit > 1 && (!p.part_o || (!p.given && (!p.part_m || !p.part_l)))) { snprintf(t_err, sizeof(t_err), "null partial buffers"); return MSDA_ERR_NULL_POINTER; }
  if (!p.given && !p.mask) { snprintf(t_err, sizeof(t_err), "online mode needs the padded column mask"); return MSDA_ERR_NULL_POINTER; }
  if (p.given && !p.col_stat) { snprintf(t_err, sizeof(t_err), "given mode needs column statistics"); return MSDA_ERR_NULL_POINTER; }
  const int dt = p.half_in ? 1 : 0;
  const long long E = static_cast<long long>(p.H) * HD;
  CUtensorMap tmA, tmB, tmX, tmOut;
  int rc = make_map3(&tmA, a, p.B, p.LA, E, BM, dt);
  if (rc) return rc;
  rc = make_map3(&tmB, b, p.B, p.LB, E, BN, dt);
  if (rc) return rc;
  rc = make_map3(&tmX, x, p.B, p.LB, E, BN, dt);
  if (rc) return rc;
  rc = make_map3(&tmOut, out16 ? out16 : a, p.B, p.LA, E, 32, dt);
  if (rc) return rc;
  int dev_id = 0;
  cudaGetDevice(&dev_id);
  static int sms_of[64] = {};
  if (!sms_of[dev_id & 63]) cudaDeviceGetAttribute(&sms_of[dev_id & 63], cudaDevAttrMultiProcessorCount, dev_id);
  const long long items = static_cast<long long>(p.B) * p.H * p.mtiles * p.nsplit;
  if (items >= (1ll << 31)) return MSDA_ERR_BAD_SHAPE;
  const int grid = items < sms_of[dev_id & 63] ? static_cast<int>(items) : sms_of[dev_id & 63];
  constexpr int smem = PvSmem<kSlots>::kBytes;
  static bool configured[64] = {};
  if (!configured[dev_id & 63]) {
    cudaError_t cfg = cudaFuncSetAttribute(biattn_pv_kernel<kSlots>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (cfg != cudaSuccess) { snprintf(t_err, sizeof(t_err), "cudaFuncSetAttribute: %s", cudaGetErrorString(cfg)); return static_cast<int>(cfg); }
    configured[dev_id & 63] = true;
  }
  ++msda::g_launches;
  biattn_pv_kernel<kSlots><<<grid, THREADS, smem, st>>>(tmA, tmB, tmX, tmOut, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(t_err, sizeof(t_err), "biattn_pv_kernel launch: %s", cudaGetErrorString(e)); return static_cast<int>(e); }
  return p.nsplit;
}

}  // namespace bia
}  // namespace pg

extern "C" {

// Number of column splits msda_biattn_pv_16 will use for a requested `nsplit` (it never leaves a split empty).
int msda_biattn_splits(int LB, int nsplit) {
  const int ctiles = (LB + pg::bia::BN - 1) / pg::bia::BN;
  if (nsplit < 1) nsplit = 1;
  if (nsplit > ctiles) nsplit = ctiles;
  const int tps = (ctiles + nsplit - 1) / nsplit;
  return (ctiles + tps - 1) / tps;
}

int msda_biattn_pv_16(const void* a, const void* b, const void* x, int B, int H, int LA, int LB, float scale,
                      const uint8_t* mask_padded, const float* col_stat, void* out16, float* lane_stat, float* part_o,
                      float* part_m, float* part_l, int nsplit, int is_half, void* stream) {
  pg::t_err[0] = 0;
  pg::bia::PvParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.H = H; p.LA = LA; p.LB = LB;
  p.nsplit = nsplit;
  p.given = col_stat != nullptr;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.mask = mask_padded; p.col_stat = col_stat; p.lane_stat = lane_stat;
  p.part_o = part_o; p.part_m = part_m; p.part_l = part_l;
  p.half_in = is_half;
  const int rc = pg::bia::launch_pv(a, b, x, out16, p, static_cast<cudaStream_t>(stream));
  return rc > 0 ? 0 : (rc == 0 ? MSDA_ERR_BAD_SHAPE : rc);
}

int msda_biattn_combine_16(const float* part_o, const float* part_m, const float* part_l, int B, int H, int LA, int nsplit,
                           int given, void* out16, float* lane_stat, int is_half, void* stream) {
  pg::t_err[0] = 0;
  if (!part_o || !out16 || (!given && (!part_m || !part_l))) { snprintf(pg::t_err, sizeof(pg::t_err), "null pointer"); return MSDA_ERR_NULL_POINTER; }
  const long long rows = static_cast<long long>(B) * H * LA;
  if (rows <= 0 || rows >= (1ll << 31)) return MSDA_ERR_BAD_SHAPE;
  const int mtiles = (LA + pg::bia::BM - 1) / pg::bia::BM;
  ++msda::g_launches;
  pg::bia::biattn_combine_kernel<<<static_cast<unsigned>(rows), pg::bia::HD, 0, static_cast<cudaStream_t>(stream)>>>(
      part_o, part_m, part_l, H, LA, mtiles, nsplit, given, out16, lane_stat, is_half);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(pg::t_err, sizeof(pg::t_err), "biattn_combine_kernel launch: %s", cudaGetErrorString(e)); return static_cast<int>(e); }
  return 0;
}

}  // extern "C"
