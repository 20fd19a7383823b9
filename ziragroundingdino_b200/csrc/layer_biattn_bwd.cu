// Backward of the bidirectional image <-> text attention: the logits gradient and its product, on tcgen05.
//
// With S = scale * A B^T (A = stationary side, B = streamed side; q/k in the rows orientation, k/q in the tokens
// orientation) both softmax directions read the same logits, so
//
//     dS[i, j] = P1[i, j] (dP1[i, j] - D1[i])  +  P2[i, j] (dP2[i, j] - D2[j])
//     P1 = exp2(s2 - Llane[i])   the direction whose softmax runs over the streamed axis   dP1 = dOA . XB^T
//     P2 = exp2(s2 - Lcol[j])    the direction whose softmax runs over the stationary axis dP2 = XA  . dOB^T
//     dA = scale * dS . B
//
// (reference: autograd through fuse_modules.py:172-227).  Rows orientation gives dq, tokens orientation gives dk.
// TMEM cannot hold three logits-sized accumulators beside a 256-column result for head_dim 256, and shared memory
// cannot hold three stationary tiles, so a work item makes TWO passes over its column tiles -- pass 0 the lane-statistics
// term (stationary pair A, dOA; streamed B, XB), pass 1 the column-statistics term (A, XA; B, dOB) -- and both accumulate
// into the same TMEM result, which is rounded once.
//
//   GEMM1_t  S_t  (64 TMEM columns) = A . B_t^T ;  dP_t (64 columns) = A2 . X2_t^T         (double-buffered pairs)
//   mid      dS_t = exp2(s2 - stat) * (dP - delta), masked, 16 bit -> ONE shared-memory tile (K-major, 128-byte rows;
//            the arithmetic of tile t+1 overlaps GEMM2_t, only the store waits for it)
//   GEMM2_t  acc (256 columns) += dS_t[128, 64] . B_t[64, 256]      (B_t again, read MN-major as TMA delivers it)
//
//   warp 0 TMA producer 1 (stationary operands + GEMM1 operands, ring 1), warp 10 TMA producer 2 (GEMM2 operand, ring 2),
//   warp 1 MMA issuer, warps 2-9 mid / final stages (two per TMEM lane quarter, 32 columns each)
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math_constants.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>
#include <type_traits>

#include "proj_epilogue.cuh"

namespace msda {
extern long long g_launches;
}

namespace pg {
namespace bid {

constexpr int HD = 256, BM = 128, BN = 64;
constexpr int A_BYTES = 4 * 16384;       // one stationary operand: 4 k-blocks of 128 rows x 128 bytes
constexpr int SLOT = 8192;               // streamed k-block / slab: 64 rows x 128 bytes
constexpr int NS1 = 6, NS2 = 2;          // ring depths: GEMM1 operands (B, X2 k-blocks, 8 KiB) / GEMM2 operand (B half-tiles, 16 KiB)
constexpr int SLOT2 = 2 * SLOT;          // two 64-column slabs, contiguous: GEMM2 runs as N = 128 instructions
constexpr int P_BYTES = 16384;           // the dS tile: 128 rows x 64 columns, 16 bit
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 96 + 32 * EPI_WARPS;     // producer 1, MMA issuer, 8 mid/final warps, producer 2
constexpr int SMEM_BYTES = 1024 + 2 * A_BYTES + NS1 * SLOT + NS2 * SLOT2 + P_BYTES + 512;

struct DsParams {
  int B, H, LA, LB;
  int mtiles, ctiles;            // ceil(LA / 128); ceil(LB / 64)
  int nsplit, tiles_per_split;
  float scale, scale_log2;
  const uint8_t* mask_a;         // [B, la_pad] 1 = stationary row masked in the softmax over the stationary axis (pass 1)
  const uint8_t* mask_b;         // [B, lb_pad] 1 = streamed row masked in the softmax over the streamed axis (pass 0)
  const float* lane_stat;        // [B, H, la_pad] log2-domain log-sum-exp, +inf padding   (pass 0)
  const float* lane_delta;       // [B, H, la_pad] rowsum(dO * O) of that direction        (pass 0)
  const float* col_stat;         // [B, H, lb_pad]                                          (pass 1)
  const float* col_delta;        // [B, H, lb_pad]                                          (pass 1)
  void* out16;                   // nsplit == 1: [B, LA, H*256] 16-bit result
  int store_terms;               // also write dS (16 bit, [B, H, LA, ctiles*64]) through tmT0: pass 0 stores its term,
                                 // pass 1 adds its own with a TMA reduction
  float* part_o;                 // nsplit > 1: [items, 128, 256] fp32 partial results (already scaled)
  int half_in;
};

template <bool HALF>
__global__ void __launch_bounds__(THREADS, 1)
biattn_ds_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2a,
                 const __grid_constant__ CUtensorMap tmA2b, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmX2a, const __grid_constant__ CUtensorMap tmX2b,
                 const __grid_constant__ CUtensorMap tmT0, DsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sA2 = sA + A_BYTES;
  uint8_t* sRing1 = sA2 + A_BYTES;
  uint8_t* sRing2 = sRing1 + NS1 * SLOT;
  uint8_t* sP = sRing2 + NS2 * SLOT2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BYTES);
  uint64_t* full1 = bars;             // [NS1]
  uint64_t* empty1 = full1 + NS1;     // [NS1]
  uint64_t* full2 = empty1 + NS1;     // [NS2]
  uint64_t* empty2 = full2 + NS2;     // [NS2]
  uint64_t* x_full = empty2 + NS2;    // A landed (once per item)
  uint64_t* x_free = x_full + 1;      // last GEMM1 of the item retired
  uint64_t* x2_full = x_free + 1;     // second stationary operand landed (twice per item)
  uint64_t* x2_free = x2_full + 1;    // last GEMM1 of a pass retired
  uint64_t* a1_full = x2_free + 1;    // [2]
  uint64_t* a1_free = a1_full + 2;    // [2]
  uint64_t* h_full = a1_free + 2;     // dS tile written
  uint64_t* h_free = h_full + 1;      // GEMM2 of that tile retired
  uint64_t* st_free = h_free + 1;     // the dS tile has been read by its TMA store (store_terms only)
  uint64_t* a2_full = st_free + 1;
  uint64_t* a2_free = a2_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a2_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int items = p.B * p.H * p.mtiles * p.nsplit;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    for (int i = 0; i < NS1; ++i) { mbar_init(full1 + i, 1); mbar_init(empty1 + i, 1); }
    for (int i = 0; i < NS2; ++i) { mbar_init(full2 + i, 1); mbar_init(empty2 + i, 1); }
    mbar_init(x_full, 1); mbar_init(x_free, 1); mbar_init(x2_full, 1); mbar_init(x2_free, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(a1_full + i, 1); mbar_init(a1_free + i, EPI_WARPS); }
    mbar_init(h_full, EPI_WARPS); mbar_init(h_free, 1); mbar_init(st_free, 1);
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
  const uint32_t t_acc1 = tmem_base, t_acc2 = tmem_base + 256u;     // acc1: [buffer][S | dP] x 64 columns

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
    // ===== TMA producer 1: stationary operands, then the GEMM1 operands (ring 1) =====
    if (elect_one()) {
      int slot = 0;
      uint32_t ph_slot = 0, ph_x = 0, ph_x2 = 0;
      auto stream = [&](const CUtensorMap* tm, int c0, int r0, int b) {
        mbar_wait(empty1 + slot, ph_slot ^ 1);
        mbar_expect_tx(full1 + slot, SLOT);
        tma_load_3d(tm, full1 + slot, sRing1 + slot * SLOT, c0, r0, b);
        if (++slot == NS1) { slot = 0; ph_slot ^= 1; }
      };
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        int b, h, mt, j0, n;
        decode(item, b, h, mt, j0, n);
        mbar_wait(x_free, ph_x ^ 1);
        ph_x ^= 1;
        mbar_expect_tx(x_full, A_BYTES);
        for (int kb = 0; kb < 4; ++kb) tma_load_3d(&tmA, x_full, sA + kb * 16384, h * HD + kb * 64, mt * BM, b);
        const int total = 2 * n;
        for (int tt = 0; tt < total; ++tt) {
          const int pass = tt >= n ? 1 : 0, j = j0 + (pass ? tt - n : tt);
          if (tt == 0 || tt == n) {
            mbar_wait(x2_free, ph_x2 ^ 1);
            ph_x2 ^= 1;
            mbar_expect_tx(x2_full, A_BYTES);
            for (int kb = 0; kb < 4; ++kb) tma_load_3d(pass ? &tmA2b : &tmA2a, x2_full, sA2 + kb * 16384, h * HD + kb * 64, mt * BM, b);
          }
          for (int kb = 0; kb < 4; ++kb) stream(&tmB, h * HD + kb * 64, j * BN, b);
          for (int kb = 0; kb < 4; ++kb) stream(pass ? &tmX2b : &tmX2a, h * HD + kb * 64, j * BN, b);
        }
      }
    }
  } else if (warp == 2 + EPI_WARPS) {
    // ===== TMA producer 2: the GEMM2 operand (B_t again, as half-tiles) through ring 2; with store_terms it also sends each
    // finished dS tile to global memory (the tokens-orientation product reads it back instead of recomputing it) =====
    if (elect_one()) {
      int slot = 0;
      uint32_t ph_slot = 0, ph_hfull = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        int b, h, mt, j0, n;
        decode(item, b, h, mt, j0, n);
        const int total = 2 * n;
        for (int t = 0; t < total; ++t) {
          const int pass = t >= n ? 1 : 0, j = j0 + (pass ? t - n : t);
          for (int sl = 0; sl < 2; ++sl) {
            mbar_wait(empty2 + slot, ph_slot ^ 1);
            mbar_expect_tx(full2 + slot, SLOT2);
            tma_load_3d(&tmB, full2 + slot, sRing2 + slot * SLOT2, h * HD + sl * 128, j * BN, b);
            tma_load_3d(&tmB, full2 + slot, sRing2 + slot * SLOT2 + SLOT, h * HD + sl * 128 + 64, j * BN, b);
            if (++slot == NS2) { slot = 0; ph_slot ^= 1; }
          }
          if (p.store_terms) {
            mbar_wait(h_full, ph_hfull);          // the mid stage has written (and proxy-fenced) the tile
            ph_hfull ^= 1;
            if (pass == 0) {
              tma_store_4d(&tmT0, sP, j * BN, mt * BM, h, b);
            } else {
              if (t == n) tma_store_wait_all();     // every pass-0 store of this item has landed before anything is added to it
              tma_reduce_add_4d(&tmT0, sP, j * BN, mt * BM, h, b);
            }
            tma_store_wait_read();
            mbar_arrive(st_free);
          }
        }
      }
      if (p.store_terms) tma_store_wait_all();
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc1 = umma_idesc(BM, BN, HALF);
    const uint32_t idesc2 = umma_idesc(BM, 128, HALF) | kIdescBMn;
    int slot1 = 0, slot2 = 0;
    uint32_t ph1 = 0, ph2 = 0, ph_x = 0, ph_x2 = 0, ph_a2 = 0, ph_hfull = 0, g1 = 0;
    uint32_t ph_a1free[2] = {0, 0};
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int b_, h_, mt_, j0_, n;
      decode(item, b_, h_, mt_, j0_, n);
      mbar_wait(x_full, ph_x);
      ph_x ^= 1;
      tc_fence_after();
      const int total = 2 * n;
      for (int tt = 0; tt <= total; ++tt) {
        if (tt < total) {
          const int b = g1 & 1;
          ++g1;
          if (tt == 0 || tt == n) {
            mbar_wait(x2_full, ph_x2);
            ph_x2 ^= 1;
          }
          mbar_wait(a1_free + b, ph_a1free[b] ^ 1);
          ph_a1free[b] ^= 1;
          tc_fence_after();
          for (int prod = 0; prod < 2; ++prod) {            // 0: logits, 1: dP
            const uint8_t* stat_op = prod ? sA2 : sA;
            const uint32_t t_d = t_acc1 + static_cast<uint32_t>(b * 128 + prod * 64);
            for (int kb = 0; kb < 4; ++kb) {
              mbar_wait(full1 + slot1, ph1);
              tc_fence_after();
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint64_t da = umma_desc_sw128(stat_op + kb * 16384, k * 32);
                  const uint64_t db = umma_desc_sw128(sRing1 + slot1 * SLOT, k * 32);
                  umma_f16(t_d, da, db, idesc1, (kb | k) != 0 ? 1u : 0u);
                }
                umma_commit(empty1 + slot1);
                if (prod == 1 && kb == 3) {
                  umma_commit(a1_full + b);
                  if (tt == n - 1 || tt == total - 1) umma_commit(x2_free);
                  if (tt == total - 1) umma_commit(x_free);
                }
              }
              __syncwarp();
              if (++slot1 == NS1) { slot1 = 0; ph1 ^= 1; }
            }
          }
        }
        if (tt >= 1) {
          const int t = tt - 1;
          if (t == 0) {
            mbar_wait(a2_free, ph_a2 ^ 1);
            ph_a2 ^= 1;
          }
          mbar_wait(h_full, ph_hfull);
          ph_hfull ^= 1;
          tc_fence_after();
          for (int sl = 0; sl < 2; ++sl) {
            mbar_wait(full2 + slot2, ph2);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {                 // K = 64 streamed rows, 16 per instruction; N = 128: two slabs 8 KiB apart
                const uint64_t da = umma_desc_sw128(sP, k * 32);
                const uint64_t db = umma_desc_mn_sw128(sRing2 + slot2 * SLOT2, k * 2048, 8192u);
                umma_f16(t_acc2 + static_cast<uint32_t>(sl * 128), da, db, idesc2, (t | k) != 0 ? 1u : 0u);
              }
              umma_commit(empty2 + slot2);
              if (sl == 1) {
                umma_commit(h_free);
                if (t == total - 1) umma_commit(a2_full);
              }
            }
            __syncwarp();
            if (++slot2 == NS2) { slot2 = 0; ph2 ^= 1; }
          }
        }
      }
    }
  } else {
    // ===== mid / final stages =====
    const int quarter = warp & 3, half = (warp - 2) >> 2;     // TMEM lane quarter; which 32 of a tile's 64 columns
    const int trow = quarter * 32 + lane;
    const uint32_t lane_bits = static_cast<uint32_t>(quarter * 32) << 16;
    uint32_t ph_a1full[2] = {0, 0}, ph_hfree = 0, ph_a2 = 0, g = 0;
    const int la_pad = p.mtiles * BM, lb_pad = (p.ctiles * BN + 127) / 128 * 128;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int b, h, mt, j0, n;
      decode(item, b, h, mt, j0, n);
      const int arow = mt * BM + trow;
      const size_t sa = (static_cast<size_t>(b) * p.H + h) * la_pad + arow;
      const float l_stat = __ldg(p.lane_stat + sa), l_delta = __ldg(p.lane_delta + sa);
      const bool lane_ok = p.mask_a[static_cast<size_t>(b) * la_pad + arow] == 0;
      const int total = 2 * n;
      for (int tt = 0; tt < total; ++tt, ++g) {
        const int bb = g & 1;
        const int pass = tt >= n ? 1 : 0, j = j0 + (pass ? tt - n : tt);
        const int col0 = j * BN + half * 32;
        // per-column inputs of this tile are fetched before waiting for its accumulators
        uint4 mk[2];
        float4 cst[8], cdl[8];
        if (pass == 0) {
          const uint4* mp = reinterpret_cast<const uint4*>(p.mask_b + static_cast<size_t>(b) * lb_pad + col0);
          mk[0] = __ldg(mp); mk[1] = __ldg(mp + 1);
        } else {
          const size_t sb = (static_cast<size_t>(b) * p.H + h) * lb_pad + col0;
          const float4* cs = reinterpret_cast<const float4*>(p.col_stat + sb);
          const float4* cd = reinterpret_cast<const float4*>(p.col_delta + sb);
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) { cst[c4] = __ldg(cs + c4); cdl[c4] = __ldg(cd + c4); }
        }
        mbar_wait(a1_full + bb, ph_a1full[bb]);
        ph_a1full[bb] ^= 1;
        tc_fence_after();
        uint32_t rs[32], rd[32];
        tmem_ld32_nowait(t_acc1 + lane_bits + static_cast<uint32_t>(bb * 128 + half * 32), rs);
        tmem_ld32_nowait(t_acc1 + lane_bits + static_cast<uint32_t>(bb * 128 + 64 + half * 32), rd);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a1_free + bb);
        float v[32];
        if (pass == 0) {
#pragma unroll
          for (int q4 = 0; q4 < 2; ++q4) {
            const uint32_t w[4] = {mk[q4].x, mk[q4].y, mk[q4].z, mk[q4].w};
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int c = q4 * 16 + i;
              const bool masked = ((w[i >> 2] >> ((i & 3) * 8)) & 0xffu) != 0u;
              const float pr = masked ? 0.f : ex2(__uint_as_float(rs[c]) * p.scale_log2 - l_stat);
              v[c] = pr * (__uint_as_float(rd[c]) - l_delta);
            }
          }
        } else {
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 st = cst[c4], dl = cdl[c4];
            const float stv[4] = {st.x, st.y, st.z, st.w}, dlv[4] = {dl.x, dl.y, dl.z, dl.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int c = c4 * 4 + i;
              const float pr = lane_ok ? ex2(__uint_as_float(rs[c]) * p.scale_log2 - stv[i]) : 0.f;
              v[c] = pr * (__uint_as_float(rd[c]) - dlv[i]);
            }
          }
        }
        uint4 pk[4];
        pack_16(v, HALF, false, pk);
        mbar_wait(h_free, ph_hfree ^ 1);                      // GEMM2 of the previous tile has finished reading the dS tile
        ph_hfree ^= 1;
        if (p.store_terms) mbar_wait(st_free, ph_hfree);      // ... and so has its store (same phase sequence as h_free)
#pragma unroll
        for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(swz(sP, trow, 4 * half + i)) = pk[i];
        tc_fence_before();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(h_full);
      }
      // ---- final stage: acc2 * scale -> 16 bit (direct) or fp32 partial ----
      mbar_wait(a2_full, ph_a2);
      ph_a2 ^= 1;
      tc_fence_after();
      if (p.store_terms) mbar_wait(st_free, ph_hfree ^ 1);  // the last tile's store has read the dS tile: it becomes staging space
      if (p.nsplit == 1) {
        // 16-bit result straight from registers: 64 contiguous bytes per thread and 32-column group (whole sectors)
        // 16-bit result: TMEM -> registers -> 2 KiB of the (idle) dS tile per warp -> coalesced global stores
        uint8_t* stage = sP + (warp - 2) * 2048;
        const int row0 = mt * BM + quarter * 32;
        const int rows_valid = p.LA - row0;
        uint8_t* gbase = static_cast<uint8_t*>(p.out16) +
                         ((static_cast<size_t>(b) * p.LA + row0) * (static_cast<size_t>(p.H) * HD) + h * HD + half * 128) * 2;
#pragma unroll 1
        for (int q4 = 0; q4 < 4; ++q4) {
          uint32_t r[32];
          float o[32];
          tmem_ld32(t_acc2 + lane_bits + static_cast<uint32_t>(half * 128 + q4 * 32), r);
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(r[i]) * p.scale;
          uint4 pk[4];
          pack_16(o, HALF, false, pk);
          store_rows_64B(stage, pk, gbase + q4 * 64, static_cast<size_t>(p.H) * HD * 2, lane, rows_valid);
        }
      } else {
        // fp32 partial result (already scaled): 16 columns (64 bytes per row) per staging pass
        uint8_t* stage = sP + (warp - 2) * 2048;
        uint8_t* gbase = reinterpret_cast<uint8_t*>(p.part_o + (static_cast<size_t>(item) * BM + quarter * 32) * HD + half * 128);
#pragma unroll 1
        for (int q4 = 0; q4 < 4; ++q4) {
          uint32_t r[32];
          tmem_ld32(t_acc2 + lane_bits + static_cast<uint32_t>(half * 128 + q4 * 32), r);
#pragma unroll
          for (int hq = 0; hq < 2; ++hq) {
            uint4 pk[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
              pk[i] = make_uint4(__float_as_uint(__uint_as_float(r[hq * 16 + 4 * i]) * p.scale), __float_as_uint(__uint_as_float(r[hq * 16 + 4 * i + 1]) * p.scale),
                                 __float_as_uint(__uint_as_float(r[hq * 16 + 4 * i + 2]) * p.scale), __float_as_uint(__uint_as_float(r[hq * 16 + 4 * i + 3]) * p.scale));
            store_rows_64B(stage, pk, gbase + q4 * 128 + hq * 64, HD * 4, lane, 32);
          }
        }
      }
      tc_fence_before();
      named_bar(1, 32 * EPI_WARPS);        // the staging pieces lie in rows other warps write in the next mid stage
      if (lane == 0) mbar_arrive(a2_free);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// delta[b, h, l] = sum_d dO[b, l, h*256 + d] * O[b, l, h*256 + d]  (one warp per (b, l, h)); `lpad` = row stride of delta
template <typename T>
__global__ void biattn_rowdot_kernel(const T* __restrict__ d_o, const T* __restrict__ o, long long rows, int H, int L, int lpad,
                                     float* __restrict__ delta) {
  const long long w = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;     // over B * L * H
  if (w >= rows * H) return;
  const int lane = threadIdx.x & 31;
  const int h = static_cast<int>(w % H);
  const long long bl = w / H;
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(d_o + w * HD) + lane);
  const uint4 c = __ldg(reinterpret_cast<const uint4*>(o + w * HD) + lane);
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, cw[4] = {c.x, c.y, c.z, c.w};
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 x, y;
    if (sizeof(T) == 2 && std::is_same<T, __half>::value) {
      x = __half22float2(*reinterpret_cast<const __half2*>(&aw[i]));
      y = __half22float2(*reinterpret_cast<const __half2*>(&cw[i]));
    } else {
      x = make_float2(__uint_as_float(aw[i] << 16), __uint_as_float(aw[i] & 0xffff0000u));
      y = make_float2(__uint_as_float(cw[i] << 16), __uint_as_float(cw[i] & 0xffff0000u));
    }
    acc = fmaf(x.x, y.x, acc);
    acc = fmaf(x.y, y.y, acc);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) {
    const long long b = bl / L, l = bl % L;
    delta[(b * H + h) * lpad + l] = acc;
  }
}

}  // namespace bid
}  // namespace pg

extern "C" {

int msda_biattn_ds_splits(int LB, int nsplit) {
  const int ctiles = (LB + pg::bid::BN - 1) / pg::bid::BN;
  if (nsplit < 1) nsplit = 1;
  if (nsplit > ctiles) nsplit = ctiles;
  const int tps = (ctiles + nsplit - 1) / nsplit;
  return (ctiles + tps - 1) / tps;
}

static int biattn_ds_impl(const void* a, const void* d_oa, const void* xa, const void* b, const void* xb, const void* d_ob, int B,
                          int H, int LA, int LB, float scale, const uint8_t* mask_a_padded, const uint8_t* mask_b_padded,
                          const float* lane_stat, const float* lane_delta, const float* col_stat, const float* col_delta,
                          void* out16, float* part_o, void* terms, int nsplit, int is_half, void* stream) {
  using namespace pg;
  using namespace pg::bid;
  t_err[0] = 0;
  if (!a || !d_oa || !xa || !b || !xb || !d_ob || !mask_a_padded || !mask_b_padded || !lane_stat || !lane_delta || !col_stat || !col_delta) {
    snprintf(t_err, sizeof(t_err), "null pointer");
    return MSDA_ERR_NULL_POINTER;
  }
  if (B <= 0 || H <= 0 || LA <= 0 || LB <= 0) { snprintf(t_err, sizeof(t_err), "bad shape"); return MSDA_ERR_BAD_SHAPE; }
  DsParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.H = H; p.LA = LA; p.LB = LB;
  p.mtiles = (LA + BM - 1) / BM;
  p.ctiles = (LB + BN - 1) / BN;
  p.nsplit = msda_biattn_ds_splits(LB, nsplit);
  p.tiles_per_split = (p.ctiles + p.nsplit - 1) / p.nsplit;
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  p.mask_a = mask_a_padded; p.mask_b = mask_b_padded;
  p.lane_stat = lane_stat; p.lane_delta = lane_delta; p.col_stat = col_stat; p.col_delta = col_delta;
  p.out16 = out16; p.part_o = part_o; p.half_in = is_half;
  if (p.nsplit == 1 && !out16) { snprintf(t_err, sizeof(t_err), "null output"); return MSDA_ERR_NULL_POINTER; }
  if (p.nsplit > 1 && !part_o) { snprintf(t_err, sizeof(t_err), "null partial buffer"); return MSDA_ERR_NULL_POINTER; }
  const int dt = is_half ? 1 : 0;
  const long long E = static_cast<long long>(H) * HD;
  CUtensorMap tmA, tmA2a, tmA2b, tmB, tmX2a, tmX2b, tmT0;
  int rc;
  if ((rc = make_map3(&tmA, a, B, LA, E, BM, dt))) return rc;
  if ((rc = make_map3(&tmA2a, d_oa, B, LA, E, BM, dt))) return rc;
  if ((rc = make_map3(&tmA2b, xa, B, LA, E, BM, dt))) return rc;
  if ((rc = make_map3(&tmB, b, B, LB, E, BN, dt))) return rc;
  if ((rc = make_map3(&tmX2a, xb, B, LB, E, BN, dt))) return rc;
  if ((rc = make_map3(&tmX2b, d_ob, B, LB, E, BN, dt))) return rc;
  p.store_terms = terms != nullptr;
  if (p.store_terms && p.nsplit != 1) { snprintf(t_err, sizeof(t_err), "dS terms are only written by unsplit launches"); return MSDA_ERR_UNSUPPORTED; }
  {
    const long long tpad = static_cast<long long>(p.ctiles) * BN;
    const void* t0 = terms ? terms : a;        // placeholders keep the maps valid when nothing is stored
    if ((rc = make_map4(&tmT0, t0, terms ? B : 1, terms ? H : 1, terms ? LA : BM, terms ? tpad : 64, BM, dt))) return rc;
  }
  int dev_id = 0;
  cudaGetDevice(&dev_id);
  static int sms_of[64] = {};
  if (!sms_of[dev_id & 63]) cudaDeviceGetAttribute(&sms_of[dev_id & 63], cudaDevAttrMultiProcessorCount, dev_id);
  const long long items = static_cast<long long>(B) * H * p.mtiles * p.nsplit;
  if (items >= (1ll << 31)) return MSDA_ERR_BAD_SHAPE;
  const int grid = items < sms_of[dev_id & 63] ? static_cast<int>(items) : sms_of[dev_id & 63];
  static bool configured[64] = {};
  if (!configured[dev_id & 63]) {
    cudaError_t cfg = cudaFuncSetAttribute(biattn_ds_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (cfg == cudaSuccess) cfg = cudaFuncSetAttribute(biattn_ds_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (cfg != cudaSuccess) { snprintf(t_err, sizeof(t_err), "cudaFuncSetAttribute: %s", cudaGetErrorString(cfg)); return static_cast<int>(cfg); }
    configured[dev_id & 63] = true;
  }
  ++msda::g_launches;
  if (is_half) biattn_ds_kernel<true><<<grid, THREADS, SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(tmA, tmA2a, tmA2b, tmB, tmX2a, tmX2b, tmT0, p);
  else biattn_ds_kernel<false><<<grid, THREADS, SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(tmA, tmA2a, tmA2b, tmB, tmX2a, tmX2b, tmT0, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(t_err, sizeof(t_err), "biattn_ds_kernel launch: %s", cudaGetErrorString(e)); return static_cast<int>(e); }
  return 0;
}

int msda_biattn_ds_16(const void* a, const void* d_oa, const void* xa, const void* b, const void* xb, const void* d_ob, int B,
                      int H, int LA, int LB, float scale, const uint8_t* mask_a_padded, const uint8_t* mask_b_padded,
                      const float* lane_stat, const float* lane_delta, const float* col_stat, const float* col_delta,
                      void* out16, float* part_o, int nsplit, int is_half, void* stream) {
  return biattn_ds_impl(a, d_oa, xa, b, xb, d_ob, B, H, LA, LB, scale, mask_a_padded, mask_b_padded, lane_stat, lane_delta, col_stat,
                        col_delta, out16, part_o, nullptr, nsplit, is_half, stream);
}

int msda_biattn_ds_terms_16(const void* a, const void* d_oa, const void* xa, const void* b, const void* xb, const void* d_ob, int B,
                            int H, int LA, int LB, float scale, const uint8_t* mask_a_padded, const uint8_t* mask_b_padded,
                            const float* lane_stat, const float* lane_delta, const float* col_stat, const float* col_delta,
                            void* out16, void* ds16, int is_half, void* stream) {
  if (!ds16) { snprintf(pg::t_err, sizeof(pg::t_err), "null dS buffer"); return MSDA_ERR_NULL_POINTER; }
  return biattn_ds_impl(a, d_oa, xa, b, xb, d_ob, B, H, LA, LB, scale, mask_a_padded, mask_b_padded, lane_stat, lane_delta, col_stat,
                        col_delta, out16, nullptr, ds16, 1, is_half, stream);
}

int msda_biattn_rowdot_16(const void* d_o, const void* o, int B, int L, int H, int lpad, float* delta, int is_half, void* stream) {
  using namespace pg;
  t_err[0] = 0;
  if (!d_o || !o || !delta) { snprintf(t_err, sizeof(t_err), "null pointer"); return MSDA_ERR_NULL_POINTER; }
  const long long rows = static_cast<long long>(B) * L;
  if (rows <= 0 || H <= 0 || lpad < L) return MSDA_ERR_BAD_SHAPE;
  const long long warps = rows * H;
  const unsigned blocks = static_cast<unsigned>((warps + 7) / 8);
  ++msda::g_launches;
  if (is_half)
    bid::biattn_rowdot_kernel<__half><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __half*>(d_o), static_cast<const __half*>(o), rows, H, L, lpad, delta);
  else
    bid::biattn_rowdot_kernel<__nv_bfloat16><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(d_o), static_cast<const __nv_bfloat16*>(o), rows, H, L, lpad, delta);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(t_err, sizeof(t_err), "biattn_rowdot_kernel launch: %s", cudaGetErrorString(e)); return static_cast<int>(e); }
  return 0;
}

}  // extern "C"
