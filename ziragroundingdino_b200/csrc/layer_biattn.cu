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
//   warp 0      TMA producer 1: the stationary tile once per work item, then B k-blocks through ring 1 (16 KiB slots)
//   warp 10     TMA producer 2: X slabs through ring 2 -- separate rings and producers, so that slabs waiting for their
//               probability tile never hold back the prefetch of the next logits operands
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
  void* out16;                 // nsplit == 1: [B, LA, H*256] 16-bit result
  float* lane_stat;            // online, nsplit == 1: [B, H, mtiles*128] log2-domain log-sum-exp per lane (may be null)
  float* part_o;               // nsplit > 1: [items, 128, 256] fp32 un-normalised partial results
  float* part_m;               // nsplit > 1, online: [items, 128] running reference;  part_l: [items, 128] partial sums
  float* part_l;
  int half_in;
  long long* trace;            // debug: per-CTA cycle counters of the pipeline waits (null = off)
};

constexpr int NS1 = 4, NS2 = 2;         // ring depths: logits operands (B k-blocks, 16 KiB) / value half-tiles (X, 32 KiB)
constexpr int SLOT2 = 2 * SLOT;         // two 64-column slabs of X, contiguous: GEMM2 runs as N = 128 instructions
constexpr int PV_THREADS = THREADS + 32; // + the second TMA producer warp
constexpr int PV_SMEM = 1024 + A_BYTES + NS1 * SLOT + NS2 * SLOT2 + P_BYTES + 2 * BM * 4 + 256;     // 231 680 of 232 448 bytes

// -DBIA_TRACE (MSDA_NVCC_EXTRA=-DBIA_TRACE csrc/build.sh) builds the pipeline-wait counters in for tools/trace_biattn.py;
// otherwise the waits are plain and the counters fold away
#ifdef BIA_TRACE
#define TWAIT(bar, par, ctr) do { const long long t0_ = clock64(); mbar_wait(bar, par); (ctr) += clock64() - t0_; } while (0)
#define TCLK() clock64()
#else
#define TWAIT(bar, par, ctr) mbar_wait(bar, par)
#define TCLK() 0ll
#endif

template <bool HALF, bool GIVEN>
__global__ void __launch_bounds__(PV_THREADS, 1)
biattn_pv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmX, PvParams p) {
  // every byte of the 227 KiB is spoken for: alignment pad, stationary tile, rings, P tile, one exchange buffer, barriers
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sA = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sRing1 = sA + A_BYTES;
  uint8_t* sRing2 = sRing1 + NS1 * SLOT;
  uint8_t* sP = sRing2 + NS2 * SLOT2;                              // ONE probability tile: the next tile's exponentials are
  float* sMax = reinterpret_cast<float*>(sP + P_BYTES);            // computed while GEMM2 still reads it; [2 halves][128]
  float* sSum = sMax;                                              // final stage only, barriers either side
  uint64_t* bars = reinterpret_cast<uint64_t*>(sMax + 2 * BM);
  uint64_t* full1 = bars;             // [NS1]
  uint64_t* empty1 = full1 + NS1;     // [NS1]
  uint64_t* full2 = empty1 + NS1;     // [NS2]
  uint64_t* empty2 = full2 + NS2;     // [NS2]
  uint64_t* x_full = empty2 + NS2;    // stationary tile landed
  uint64_t* x_free = x_full + 1;      // last GEMM1 of the work item retired
  uint64_t* a1_full = x_free + 1;     // [2] logits tile complete
  uint64_t* a1_free = a1_full + 2;    // [2] mid stage has read it
  uint64_t* h_full = a1_free + 2;     // probability tile written to shared memory
  uint64_t* h_free = h_full + 1;      // GEMM2 has read it (== GEMM2 of that tile complete)
  uint64_t* a2_full = h_free + 1;     // result complete
  uint64_t* a2_free = a2_full + 1;    // final stage has read it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a2_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int items = p.B * p.H * p.mtiles * p.nsplit;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
    for (int i = 0; i < NS1; ++i) { mbar_init(full1 + i, 1); mbar_init(empty1 + i, 1); }
    for (int i = 0; i < NS2; ++i) { mbar_init(full2 + i, 1); mbar_init(empty2 + i, 1); }
    mbar_init(x_full, 1); mbar_init(x_free, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(a1_full + i, 1); mbar_init(a1_free + i, EPI_WARPS); }
    mbar_init(h_full, EPI_WARPS); mbar_init(h_free, 1);
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
    // ===== TMA producer 1: the stationary tile, then the logits operands (ring 1) =====
    if (elect_one()) {
      int slot = 0;
      uint32_t ph_slot = 0, ph_x = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        int b, h, mt, j0, n;
        decode(item, b, h, mt, j0, n);
        mbar_wait(x_free, ph_x ^ 1);
        ph_x ^= 1;
        mbar_expect_tx(x_full, A_BYTES);
        for (int kb = 0; kb < 4; ++kb) tma_load_3d(&tmA, x_full, sA + kb * 16384, h * HD + kb * 64, mt * BM, b);
        for (int jj = 0; jj < n; ++jj) {
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(empty1 + slot, ph_slot ^ 1);
            mbar_expect_tx(full1 + slot, SLOT);
            tma_load_3d(&tmB, full1 + slot, sRing1 + slot * SLOT, h * HD + kb * 64, (j0 + jj) * BN, b);
            if (++slot == NS1) { slot = 0; ph_slot ^= 1; }
          }
        }
      }
    }
  } else if (warp == 2 + EPI_WARPS) {
    // ===== TMA producer 2: the value slabs (ring 2) -- its own warp, so waiting for GEMM2 never holds back ring 1 =====
    if (elect_one()) {
      int slot = 0;
      uint32_t ph_slot = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        int b, h, mt, j0, n;
        decode(item, b, h, mt, j0, n);
        for (int jj = 0; jj < n; ++jj) {
          for (int sl = 0; sl < 2; ++sl) {     // half a value tile: head-dim columns [sl*128, +128) as two 64-column slabs
            mbar_wait(empty2 + slot, ph_slot ^ 1);
            mbar_expect_tx(full2 + slot, SLOT2);
            tma_load_3d(&tmX, full2 + slot, sRing2 + slot * SLOT2, h * HD + sl * 128, (j0 + jj) * BN, b);
            tma_load_3d(&tmX, full2 + slot, sRing2 + slot * SLOT2 + SLOT, h * HD + sl * 128 + 64, (j0 + jj) * BN, b);
            if (++slot == NS2) { slot = 0; ph_slot ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc1 = umma_idesc(BM, BN, HALF);
    const uint32_t idesc2 = umma_idesc(BM, 128, HALF) | kIdescBMn;
    int slot1 = 0, slot2 = 0;
    uint32_t ph1 = 0, ph2 = 0, ph_x = 0, ph_a2 = 0, ph_hfull = 0, g1 = 0;     // g1: running tile counter of GEMM1
    uint32_t ph_a1free[2] = {0, 0};
    long long tc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_begin = TCLK();
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int b_, h_, mt_, j0_, n;
      decode(item, b_, h_, mt_, j0_, n);
      TWAIT(x_full, ph_x, tc[0]);
      ph_x ^= 1;
      tc_fence_after();
      for (int jj = 0; jj <= n; ++jj) {
        if (jj < n) {
          const int b = g1 & 1;
          ++g1;
          TWAIT(a1_free + b, ph_a1free[b] ^ 1, tc[1]);
          ph_a1free[b] ^= 1;
          tc_fence_after();
          for (int kb = 0; kb < 4; ++kb) {
            TWAIT(full1 + slot1, ph1, tc[2]);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t da = umma_desc_sw128(sA + kb * 16384, k * 32);
                const uint64_t db = umma_desc_sw128(sRing1 + slot1 * SLOT, k * 32);
                umma_f16(t_acc1 + static_cast<uint32_t>(b * BN), da, db, idesc1, (kb | k) != 0 ? 1u : 0u);
              }
              umma_commit(empty1 + slot1);
              if (kb == 3) {
                umma_commit(a1_full + b);
                if (jj == n - 1) umma_commit(x_free);
              }
            }
            __syncwarp();
            if (++slot1 == NS1) { slot1 = 0; ph1 ^= 1; }
          }
        }
        if (jj >= 1) {
          const int t = jj - 1;
          if (t == 0) {                        // first product into acc2: the previous item's result must have been read out
            TWAIT(a2_free, ph_a2 ^ 1, tc[3]);
            ph_a2 ^= 1;
          }
          TWAIT(h_full, ph_hfull, tc[4]);
          ph_hfull ^= 1;
          tc_fence_after();
          for (int sl = 0; sl < 2; ++sl) {
            TWAIT(full2 + slot2, ph2, tc[5]);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {    // K = 128 streamed rows, 16 per instruction; N = 128: two slabs 16 KiB apart
                const uint64_t da = umma_desc_sw128(sP + (k >> 2) * 16384, (k & 3) * 32);
                const uint64_t db = umma_desc_mn_sw128(sRing2 + slot2 * SLOT2, k * 2048, 16384u);
                umma_f16(t_acc2 + static_cast<uint32_t>(sl * 128), da, db, idesc2, (t | k) != 0 ? 1u : 0u);
              }
              umma_commit(empty2 + slot2);
              if (sl == 1) {
                umma_commit(h_free);
                if (t == n - 1) umma_commit(a2_full);
              }
            }
            __syncwarp();
            if (++slot2 == NS2) { slot2 = 0; ph2 ^= 1; }
          }
        }
      }
    }
    if (p.trace != nullptr && lane == 0) {
      tc[6] = TCLK() - t_begin;
      for (int i = 0; i < 8; ++i) p.trace[blockIdx.x * 16 + i] = tc[i];
    }
  } else {
    // ===== mid / final stages =====
    const int quarter = warp & 3, half = (warp - 2) >> 2;     // TMEM lane quarter; which 64 of a tile's 128 columns
    const int trow = quarter * 32 + lane;                     // row inside the stationary tile == TMEM lane
    const uint32_t lane_bits = static_cast<uint32_t>(quarter * 32) << 16;
    uint32_t ph_a1full[2] = {0, 0}, ph_hfree = 0, ph_a2 = 0, g = 0;
    long long te[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_begin = TCLK();
    const int lb_pad = p.ctiles * BN, la_pad = p.mtiles * BM;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int b, h, mt, j0, n;
      decode(item, b, h, mt, j0, n);
      const int arow = mt * BM + trow;                        // row of the stationary operand
      const bool live = arow < p.LA;
      bool lane_ok = live;
      if (GIVEN && p.mask != nullptr) lane_ok = lane_ok && p.mask[static_cast<size_t>(b) * la_pad + arow] == 0;
      float m_ref = -CUDART_INF_F, l_part = 0.f;
      for (int jj = 0; jj < n; ++jj, ++g) {
        const int bb = g & 1;
        const int col0 = (j0 + jj) * BN + half * 64;          // first streamed row (= logits column) this thread handles
        // the tile's mask bytes / column statistics are fetched before waiting for its logits (an L2 round trip otherwise
        // sits between the TMEM load and the first use)
        uint4 mk[4];
        const float4* cs = nullptr;
        if (!GIVEN) {
          const uint4* mp = reinterpret_cast<const uint4*>(p.mask + static_cast<size_t>(b) * lb_pad + col0);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) mk[q4] = __ldg(mp + q4);
        } else {                                              // 256 bytes of statistics: pull the two lines into L1
          cs = reinterpret_cast<const float4*>(p.col_stat + (static_cast<size_t>(b) * p.H + h) * lb_pad + col0);
          asm volatile("prefetch.global.L1 [%0];" ::"l"(cs));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(cs + 8));
        }
        TWAIT(a1_full + bb, ph_a1full[bb], te[0]);
        ph_a1full[bb] ^= 1;
        tc_fence_after();
        uint32_t r0[32], r1[32];
        const uint32_t t_s = t_acc1 + lane_bits + static_cast<uint32_t>(bb * BN + half * 64);
        { const long long t0_ = TCLK();
        tmem_ld32_nowait(t_s, r0);
        tmem_ld32_nowait(t_s + 32u, r1);
        tmem_ld_wait();
        te[5] += TCLK() - t0_; }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a1_free + bb);             // logits are in registers: GEMM1 of tile g + 2 may overwrite them
        float s[64];
        float sc = 1.f;
        bool need = false;
        if (!GIVEN) {
          uint32_t any_mask = 0u;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) any_mask |= mk[q4].x | mk[q4].y | mk[q4].z | mk[q4].w;
          // raw logits (masked ones -> -inf) stay in s[]; the scale is positive, so max(raw) * scale is the scaled maximum
          float rm[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};     // independent chains
          if (any_mask == 0u) {                               // uniform across the warp: the mask is per column
#pragma unroll
            for (int c = 0; c < 64; ++c) {
              s[c] = __uint_as_float(c < 32 ? r0[c & 31] : r1[c & 31]);
              rm[c & 3] = fmaxf(rm[c & 3], s[c]);
            }
          } else {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const uint32_t w[4] = {mk[q4].x, mk[q4].y, mk[q4].z, mk[q4].w};
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int c = q4 * 16 + i;
                const float raw = __uint_as_float(c < 32 ? r0[c & 31] : r1[c & 31]);
                const bool masked = ((w[i >> 2] >> ((i & 3) * 8)) & 0xffu) != 0u;
                s[c] = masked ? -CUDART_INF_F : raw;
                rm[c & 3] = fmaxf(rm[c & 3], s[c]);
              }
            }
          }
          float tmax = fmaxf(fmaxf(rm[0], rm[1]), fmaxf(rm[2], rm[3])) * p.scale_log2;
          sMax[half * BM + trow] = tmax;
          { const long long t0_ = TCLK(); named_bar(1 + quarter, 64); te[1] += TCLK() - t0_; }
          tmax = fmaxf(tmax, sMax[(half ^ 1) * BM + trow]);
          named_bar(1 + quarter, 64);                         // one exchange buffer: both have read before the next tile writes
          need = tmax > m_ref + kRaise;                       // also true for the first finite tile (m_ref = -inf)
          if (need) {
            sc = ex2(m_ref - tmax);                            // 0 when m_ref = -inf
            l_part *= sc;
            m_ref = tmax;
          }
          const float ref_use = m_ref == -CUDART_INF_F ? 0.f : m_ref;
#pragma unroll
          for (int c = 0; c < 64; ++c) s[c] = fmaf(s[c], p.scale_log2, -ref_use);
        } else {
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
        }
        // exponentials and packing overlap GEMM2 of the previous tile, which still reads the P tile
        const long long t_exp = TCLK();
        uint4 pk[8];
        float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v[i] = ex2(s[hf * 32 + i]);
            ls[i & 3] += v[i];
          }
          uint4 o4[4];
          pack_16(v, HALF, false, o4);
#pragma unroll
          for (int i = 0; i < 4; ++i) pk[hf * 4 + i] = o4[i];
        }
        l_part += (ls[0] + ls[1]) + (ls[2] + ls[3]);
        te[7] += TCLK() - t_exp;
        TWAIT(h_free, ph_hfree ^ 1, te[2]);                      // GEMM2 of tile g - 1 retired: P tile and accumulator are ours
        ph_hfree ^= 1;
        if (jj > 0 && __any_sync(0xffffffffu, need)) {
          // raise the reference: rescale this warp's half of the accumulator before the next product is added to it
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
        uint8_t* pk_base = sP + half * 16384;
#pragma unroll
        for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(swz(pk_base, trow, i)) = pk[i];
        tc_fence_before();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(h_full);
      }
      // ---- final stage ----
      if (n > 0) {
        TWAIT(a2_full, ph_a2, te[3]);
        ph_a2 ^= 1;
        tc_fence_after();
      }
      const long long t_fin = TCLK();
      sSum[half * BM + trow] = l_part;
      named_bar(1 + quarter, 64);
      const float l_tot = l_part + sSum[(half ^ 1) * BM + trow];
      named_bar(1 + quarter, 64);
      if (p.nsplit == 1) {
        const float inv = GIVEN ? 1.f : (l_tot > 0.f ? 1.f / l_tot : 0.f);
        if (!GIVEN && p.lane_stat != nullptr && half == 0 && live)
          p.lane_stat[(static_cast<size_t>(b) * p.H + h) * la_pad + arow] = l_tot > 0.f ? m_ref + log2f(l_tot) : CUDART_INF_F;
        // 16-bit result: TMEM -> registers -> this warp's 4 KiB of the (idle) P tile -> coalesced global stores
        uint8_t* stage = sP + (half * 4 + quarter) * 4096;    // rows only this warp writes in the mid stage
        const int row0 = mt * BM + quarter * 32;
        const int rows_valid = p.LA - row0;
        uint8_t* gbase = static_cast<uint8_t*>(p.out16) +
                         ((static_cast<size_t>(b) * p.LA + row0) * (static_cast<size_t>(p.H) * HD) + h * HD + half * 128) * 2;
#pragma unroll 1
        for (int gq = 0; gq < 2; ++gq) {
          uint4 pk8[8];
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t r[32];
            float v[32];
            tmem_ld32(t_acc2 + lane_bits + static_cast<uint32_t>(half * 128 + gq * 64 + hf * 32), r);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * inv;
            uint4 o4[4];
            pack_16(v, HALF, false, o4);
#pragma unroll
            for (int i = 0; i < 4; ++i) pk8[hf * 4 + i] = o4[i];
          }
          store_rows_128B(stage, pk8, gbase + gq * 128, static_cast<size_t>(p.H) * HD * 2, lane, rows_valid);
        }
      } else {
        // fp32 partial result of this warp's 32 rows x 128 columns, 32 columns (128 bytes per row) per staging pass
        uint8_t* stage = sP + (half * 4 + quarter) * 4096;
        uint8_t* gbase = reinterpret_cast<uint8_t*>(p.part_o + (static_cast<size_t>(item) * BM + quarter * 32) * HD + half * 128);
#pragma unroll 1
        for (int q4 = 0; q4 < 4; ++q4) {
          uint32_t r[32];
          if (n > 0) tmem_ld32(t_acc2 + lane_bits + static_cast<uint32_t>(half * 128 + q4 * 32), r);
          uint4 pk8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            pk8[i] = n > 0 ? make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]) : make_uint4(0u, 0u, 0u, 0u);
          store_rows_128B(stage, pk8, gbase + q4 * 128, HD * 4, lane, 32);
        }
        if (!GIVEN && half == 0) {
          p.part_m[static_cast<size_t>(item) * BM + trow] = m_ref;
          p.part_l[static_cast<size_t>(item) * BM + trow] = l_tot;
        }
      }
      if (n > 0) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a2_free);
      }
      te[4] += TCLK() - t_fin;
    }
    if (p.trace != nullptr && warp == 2 && lane == 0) {
      te[6] = TCLK() - t_begin;
      for (int i = 0; i < 8; ++i) p.trace[blockIdx.x * 16 + 8 + i] = te[i];
    }
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

long long* g_pv_trace = nullptr;

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
  CUtensorMap tmA, tmB, tmX;
  int rc = make_map3(&tmA, a, p.B, p.LA, E, BM, dt);
  if (rc) return rc;
  rc = make_map3(&tmB, b, p.B, p.LB, E, BN, dt);
  if (rc) return rc;
  rc = make_map3(&tmX, x, p.B, p.LB, E, BN, dt);
  if (rc) return rc;
  int dev_id = 0;
  cudaGetDevice(&dev_id);
  static int sms_of[64] = {};
  if (!sms_of[dev_id & 63]) cudaDeviceGetAttribute(&sms_of[dev_id & 63], cudaDevAttrMultiProcessorCount, dev_id);
  const long long items = static_cast<long long>(p.B) * p.H * p.mtiles * p.nsplit;
  if (items >= (1ll << 31)) return MSDA_ERR_BAD_SHAPE;
  const int grid = items < sms_of[dev_id & 63] ? static_cast<int>(items) : sms_of[dev_id & 63];
  constexpr int smem = PV_SMEM;
  static bool configured[64] = {};
  if (!configured[dev_id & 63]) {
    cudaError_t cfg = cudaFuncSetAttribute(biattn_pv_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (cfg == cudaSuccess) cfg = cudaFuncSetAttribute(biattn_pv_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (cfg == cudaSuccess) cfg = cudaFuncSetAttribute(biattn_pv_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (cfg == cudaSuccess) cfg = cudaFuncSetAttribute(biattn_pv_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (cfg != cudaSuccess) { snprintf(t_err, sizeof(t_err), "cudaFuncSetAttribute: %s", cudaGetErrorString(cfg)); return static_cast<int>(cfg); }
    configured[dev_id & 63] = true;
  }
  p.trace = g_pv_trace;
  p.out16 = out16;
  ++msda::g_launches;
  if (p.half_in) {
    if (p.given) biattn_pv_kernel<true, true><<<grid, PV_THREADS, smem, st>>>(tmA, tmB, tmX, p);
    else biattn_pv_kernel<true, false><<<grid, PV_THREADS, smem, st>>>(tmA, tmB, tmX, p);
  } else {
    if (p.given) biattn_pv_kernel<false, true><<<grid, PV_THREADS, smem, st>>>(tmA, tmB, tmX, p);
    else biattn_pv_kernel<false, false><<<grid, PV_THREADS, smem, st>>>(tmA, tmB, tmX, p);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(t_err, sizeof(t_err), "biattn_pv_kernel launch: %s", cudaGetErrorString(e)); return static_cast<int>(e); }
  return p.nsplit;
}

}  // namespace bia
}  // namespace pg

extern "C" {

// Debug: device buffer of 16 int64 per CTA receiving the pipeline wait counters of the next launches (null = off).
void msda_biattn_set_trace(long long* buf) { pg::bia::g_pv_trace = buf; }

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
