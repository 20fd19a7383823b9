// tcgen05 / TMA projection GEMMs for the MultiScaleDeformableAttention module (sm_100a only).
//
// Replaces the four nn.Linear calls of the reference module and the elementwise passes around them
// (ms_deform_attn.py:286-325): value_proj + key-padding mask (:286-289), sampling_offsets ->
// sampling locations (:290-292, :306-319), attention_weights -> softmax (:293-303) and output_proj
// (:350).  Every one of them is Y[R, Nout] = X[R, K] * W[Nout, K]^T with K = 256: left of the B200
// ridge, i.e. bound by reading X and writing Y, so the design goal is ONE pass over the activation
// with everything else folded into the epilogue, not peak MMA rate.
//
// Structure (one persistent CTA per SM, 320 threads, warp-specialised):
//   warp 0     TMA producer: cp.async.bulk.tensor 128x64 (A) and block_n x 64 (B) bf16 boxes, 128B swizzle,
//              into a 4-stage shared-memory ring; completion on mbarriers (expect_tx).
//   warp 1     MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=block_n,
//              K=16) four times per stage; tcgen05.commit releases the stage and, after the last K
//              block, publishes the accumulator.  Also owns the TMEM allocation (2 x block_n columns).
//   warps 2-9  epilogue (EIGHT warps: two per TMEM lane quarter, alternating column chunks): tcgen05.ld 32 lanes x 32
//              columns -> registers (thread = one output row, 32 consecutive columns), bias + {mask | sampling-location |
//              softmax | ReLU / gate | ZiRa fold} math, then a per-warp 128B-swizzled shared-memory tile written back with
//              ONE TMA store (cp.async.bulk.tensor) per 32 x 128-byte tile, or staged 16-byte stores for the fp32 /
//              ZiRa outputs.  Double-buffered accumulators let the epilogue of tile i overlap the MMAs of tile i+1.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>

#include "../../include/msda_b200.h"
#include "proj_epilogue.cuh"
#include "tc_common.cuh"

namespace msda {
extern long long g_launches;
}

namespace pg {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;              // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;                // maximum ring depth; the host may pick 3 to make room for the store tiles
constexpr int SMEM_A = BLOCK_M * BLOCK_K * 2;   // 16 KiB
constexpr int MAX_N = 2048;              // widest (stacked) weight handled by one launch
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + 32 * EPI_WARPS;   // TMA warp + MMA warp + epilogue warps
constexpr int AUX_BASE = 256 + MAX_N * 4 + MSDA_MAX_LEVELS * 8;
constexpr int AUX_BYTES_16 = AUX_BASE + EPI_WARPS * 32 * 80;    // 16-bit outputs
constexpr int AUX_BYTES_32 = AUX_BASE + EPI_WARPS * 32 * 144;   // fp32 outputs
constexpr int TMA_TILE_BYTES = 32 * 128;                        // one epilogue warp's 128B-swizzled store tile
constexpr int AUX_BYTES_TMA = AUX_BASE + EPI_WARPS * TMA_TILE_BYTES;
constexpr int kMaxSmem = 227 * 1024;
bool g_allow_resident = true;
int g_store_bufs = 1;            // store tiles per epilogue warp; 2 measured slower (profiles/r1_gemm_ab.txt), kept as an A/B knob
bool g_prefer_resident = true;   // true: never give up a resident W slice for the second store tile
bool g_tma_store = true;         // msda_b200_gemm_set_staged(0) turns the TMA-store epilogue off (A/B)
bool g_staged_store = true;      // msda_b200_gemm_set_staged(0|1)   // msda_b200_set_tuning("gemm_resident", 0|1)

// ---- the kernel -----------------------------------------------------------------------------------
// LN (EPI_STORE through TMA, accumulate form, block_n == Nout <= 256 so that a tile holds whole rows): the launch also
// performs the LayerNorm that follows the projection in the reference layer -- `src = norm1(src + output_proj(...))`,
// transformer_for_adapter.py:901-902 -- on the rounded sum z = accum + x W^T + b it stores: row statistics are collected
// while z is written, exchanged between the two warps of a lane quarter, and a second pass over the accumulator (still in
// tensor memory) writes y = (z - mean) * rstd * gamma + beta through the second output map; mean / rstd go to ep.ln_mean /
// ep.ln_rstd for the backward.
template <int MODE, bool OUT_F32, bool HALF_OUT, bool RELU, bool GATE, bool TMA_OUT, bool LN = false>
__global__ void __launch_bounds__(THREADS, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
                 const __grid_constant__ CUtensorMap tmC2, int R, int Nout, int K, int k_split, int k_b, int block_n,
                 int b_resident, int stages, int store_bufs, int half_in, EpiParams ep) {
  // Two activation operands, K-concatenated: k-blocks [0, k_split) come from tmA, the rest from tmA2 (same rows), so
  // [x1 | x2] W^T needs no concatenated copy of the activations.  W has k_b k-blocks and k-block kb multiplies W's block
  // kb % k_b: k_b == K / 64 is the plain product, k_b == k_split == K / 128 is (x1 + x2) W^T = x1 W^T + x2 W^T with ONE
  // resident copy of W (`query = src + pos`, transformer_for_adapter.py:867-869, folded into the query projection).
  // b_resident: the CTA owns ONE n-block for its whole life and keeps that slice of W (block_n x K) in shared
  // memory, loaded once; only the activation tiles stream through the ring (K = 256, N <= 256: 128 KiB of W).
  // Otherwise A and B tiles stream together (any shape).
  extern __shared__ uint8_t smem_raw[];
  // align by OFFSET, not by casting through an integer: the compiler then still knows these are shared-memory
  // addresses and emits LDS/STS instead of generic LD/ST (which go through L1TEX address translation)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int b_tile_bytes = block_n * BLOCK_K * 2;
  const int stage_bytes = b_resident ? SMEM_A : SMEM_A + b_tile_bytes;
  uint8_t* smem_bres = smem + stages * stage_bytes;                       // resident W slice: (K/64) tiles
  uint8_t* tma_tiles = smem_bres + (b_resident ? k_b * b_tile_bytes : 0);   // 1024-aligned (all sizes above are)
  uint8_t* aux = tma_tiles + (TMA_OUT ? EPI_WARPS * store_bufs * TMA_TILE_BYTES : 0);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* bres_bar = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres_bar + 1);
  float* s_bias = reinterpret_cast<float*>(aux + 256);
  float* s_norm = s_bias + MAX_N;   // (1/W_l, 1/H_l)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k = K / BLOCK_K;
  const int num_n = Nout / block_n;
  const int num_m = (R + BLOCK_M - 1) / BLOCK_M;
  // tile enumeration: streaming mode walks (m, n) tiles n-fastest; resident mode fixes n = blockIdx.x % num_n and walks m
  const int t_first = b_resident ? blockIdx.x / num_n : blockIdx.x;
  const int t_step = b_resident ? gridDim.x / num_n : gridDim.x;
  const int t_end = b_resident ? num_m : num_m * num_n;
  const int n_fixed = (blockIdx.x % num_n) * block_n;
  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(2 * block_n)) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    if (k_split < num_k) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA2)) : "memory");
    for (int i = 0; i < STAGES; ++i) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull_bar + i, 1); mbar_init(tempty_bar + i, EPI_WARPS); }
    mbar_init(bres_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < Nout; i += THREADS) s_bias[i] = ep.bias ? ep.bias[i] : 0.f;
  if constexpr (LN) {   // s_bias[256..511] = gamma, [512..767] = beta, [768..1791] = the statistics exchange, two buffers (MAX_N = 2048 floats)
    for (int i = threadIdx.x; i < Nout; i += THREADS) { s_bias[256 + i] = ep.ln_gamma[i]; s_bias[512 + i] = ep.ln_beta[i]; }
  }
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
      if (b_resident && t_first < t_end) {
        mbar_expect_tx(bres_bar, static_cast<uint32_t>(k_b * b_tile_bytes));
        for (int kb = 0; kb < k_b; ++kb) tma_load_2d(&tmB, bres_bar, smem_bres + kb * b_tile_bytes, kb * BLOCK_K, n_fixed);
      }
      for (int t = t_first; t < t_end; t += t_step) {
        const int m_idx = (b_resident ? t : t / num_n) * BLOCK_M, n_idx = b_resident ? n_fixed : (t % num_n) * block_n;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(empty_bar + stage, phase ^ 1);
          uint8_t* sa = smem + stage * stage_bytes;
          mbar_expect_tx(full_bar + stage, static_cast<uint32_t>(stage_bytes));
          if (kb < k_split) tma_load_2d(&tmA, full_bar + stage, sa, kb * BLOCK_K, m_idx);
          else tma_load_2d(&tmA2, full_bar + stage, sa, (kb - k_split) * BLOCK_K, m_idx);
          if (!b_resident) tma_load_2d(&tmB, full_bar + stage, sa + SMEM_A, (kb % k_b) * BLOCK_K, n_idx);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = umma_idesc(BLOCK_M, block_n, half_in != 0);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    if (b_resident && t_first < t_end) mbar_wait(bres_bar, 0);
    for (int t = t_first; t < t_end; t += t_step) {
      mbar_wait(tempty_bar + acc, acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * block_n);
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(full_bar + stage, phase);
        tc_fence_after();
        if (elect_one()) {
          const uint8_t* sa = smem + stage * stage_bytes;
          const uint8_t* sb = b_resident ? smem_bres + (kb % k_b) * b_tile_bytes : sa + SMEM_A;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = umma_desc_sw128(sa, k * UMMA_K * 2);
            const uint64_t db = umma_desc_sw128(sb, k * UMMA_K * 2);
            umma_f16(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar + stage);                      // smem stage free once these MMAs retire
          if (kb == num_k - 1) umma_commit(tfull_bar + acc);    // accumulator complete
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===== epilogue warps: TMEM lane quarter = warp % 4; the two warps of a quarter alternate column chunks =====
    const int quarter = warp & 3;
    const int chunk_par = (warp - 2) >> 2;     // 0 or 1
    constexpr int kStagePitch = (OUT_F32 || MODE == EPI_QUERY) ? 144 : 80;
    uint8_t* stage = reinterpret_cast<uint8_t*>(s_norm + 2 * MSDA_MAX_LEVELS) + (warp - 2) * 32 * kStagePitch;
    int acc = 0;
    uint32_t acc_phase = 0;
    float zsum_b = 0.f, zsum_o = 0.f;   // EPI_ZIRA: this thread's share of the two SmoothL1 sums
    for (int t = t_first; t < t_end; t += t_step) {
      const int m_idx = (b_resident ? t : t / num_n) * BLOCK_M, n_idx = b_resident ? n_fixed : (t % num_n) * block_n;
      // LN: this row's residual columns (both 64-column groups of this warp) are fetched BEFORE the accumulator wait and kept
      // for both passes -- the epilogue is otherwise a chain of exposed global-load latencies (measured: 10.8 us per tile)
      uint4 ln_res0[8], ln_res1[8];
      if constexpr (LN) {
        const long long rr = static_cast<long long>(m_idx) + quarter * 32 + lane;
        const uint4* ap = reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(ep.accum) + rr * ep.out_ld + n_idx + chunk_par * 64);
        const bool two = chunk_par * 64 + 128 < block_n;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          ln_res0[i] = rr < R ? __ldg(ap + i) : make_uint4(0u, 0u, 0u, 0u);
          ln_res1[i] = (rr < R && two) ? __ldg(ap + 16 + i) : make_uint4(0u, 0u, 0u, 0u);
        }
      }
      mbar_wait(tfull_bar + acc, acc_phase);
      tc_fence_after();
      const long long row = static_cast<long long>(m_idx) + quarter * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * block_n);
      const bool live = row < R;
      const bool zero = live && ep.row_mask != nullptr && ep.row_mask[row] != 0;
      const long long row0 = static_cast<long long>(m_idx) + quarter * 32;
      const int rows_valid = static_cast<int>(R - row0 < 32 ? (R - row0 < 0 ? 0 : R - row0) : 32);
      if (MODE == EPI_ZIRA) {
        const float s = __ldg(ep.scaling);
        for (int c0 = chunk_par * 96; c0 < block_n; c0 += 192) {
          uint32_t r0[32], r1[32], r2[32];
          tmem_ld32(taddr + c0, r0);
          tmem_ld32(taddr + c0 + 32, r1);
          tmem_ld32(taddr + c0 + 64, r2);
          const int f0 = (n_idx + c0) / 3;     // first of the 32 features of this triple
          float y[32], pre[32], ad[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            pre[j] = __uint_as_float(r2[j]) + s_bias[2 * ep.F + f0 + j];
            const float br = s * pre[j];
            ad[j] = br + __uint_as_float(r1[j]) + s_bias[ep.F + f0 + j];
            y[j] = __uint_as_float(r0[j]) + s_bias[f0 + j] + ad[j];
            if (live) { zsum_b += smooth_l1(br); zsum_o += smooth_l1(ad[j]); }
          }
          uint4 pk[4];
          pack_16(y, HALF_OUT, zero, pk);
          staged_store<4>(stage, lane, pk, static_cast<uint16_t*>(ep.out) + row0 * ep.out_ld + f0, 2ll * ep.out_ld, rows_valid);
          if (ep.pre_out) {
            pack_16(pre, HALF_OUT, false, pk);
            staged_store<4>(stage, lane, pk, static_cast<uint16_t*>(ep.pre_out) + row0 * ep.F + f0, 2ll * ep.F, rows_valid);
          }
          if (ep.adapter_out) {
            pack_16(ad, HALF_OUT, false, pk);
            staged_store<4>(stage, lane, pk, static_cast<uint16_t*>(ep.adapter_out) + row0 * ep.F + f0, 2ll * ep.F, rows_valid);
          }
        }
      } else if (TMA_OUT && MODE == EPI_STORE) {
        // ---- 16-bit output through TMA: 64 columns (128 bytes per row) per step, written into this warp's
        // 128B-swizzled 32 x 128-byte tile (conflict-free STS), then ONE cp.async.bulk.tensor store.  Rows >= R are
        // clipped by the tensor map.  The two warps of a lane quarter alternate 64-column groups.
        // one or two store tiles per warp: with two, a step only waits for the store issued two steps ago
        uint8_t* tile0 = tma_tiles + (warp - 2) * store_bufs * TMA_TILE_BYTES;
        int buf = 0;
        float ln_s1 = 0.f, ln_s2 = 0.f;
        uint4 gnext[8];
        const int g_j = lane & 7, g_rsub = lane >> 3;           // gate fetch: 8 lanes along a 128-byte row, 4 rows per instr
        auto gate_fetch = [&](int c) {
          const uint8_t* g = reinterpret_cast<const uint8_t*>(static_cast<const uint16_t*>(ep.gate) + row0 * ep.out_ld + n_idx + c);
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + g_rsub;
            gnext[it] = (rr < rows_valid) ? __ldg(reinterpret_cast<const uint4*>(g + rr * 2ll * ep.out_ld + g_j * 16)) : make_uint4(0u, 0u, 0u, 0u);
          }
        };
        const bool bit_gate = GATE && ep.gate_bits != nullptr;
        uint32_t gb_next[2] = {0u, 0u};
        auto bits_fetch = [&](int c) {   // this row's two gate words of a 64-column step, one step ahead
          if (live) {
            gb_next[0] = __ldg(ep.gate_bits + static_cast<size_t>((n_idx + c) >> 5) * R + row);
            gb_next[1] = __ldg(ep.gate_bits + static_cast<size_t>(((n_idx + c) >> 5) + 1) * R + row);
          }
        };
        if (GATE && chunk_par * 64 < block_n) { if (bit_gate) bits_fetch(chunk_par * 64); else gate_fetch(chunk_par * 64); }
        for (int c0 = chunk_par * 64; c0 < block_n; c0 += 128) {
          const int gc = n_idx + c0;
          uint8_t* tile = tile0 + buf * TMA_TILE_BYTES;
          uint4 gt[8];
          uint32_t gb[2] = {gb_next[0], gb_next[1]};
          uint4 ain[8];
          if constexpr (LN) {
            const bool first = c0 == chunk_par * 64;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              ain[i].x = first ? ln_res0[i].x : ln_res1[i].x; ain[i].y = first ? ln_res0[i].y : ln_res1[i].y;
              ain[i].z = first ? ln_res0[i].z : ln_res1[i].z; ain[i].w = first ? ln_res0[i].w : ln_res1[i].w;
            }
          } else if (!RELU && !GATE && ep.accum != nullptr) {   // this row's 64 columns of the tensor to accumulate onto
            const uint4* ap = reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(ep.accum) + row * ep.out_ld + gc);
#pragma unroll
            for (int i = 0; i < 8; ++i) ain[i] = live ? ap[i] : make_uint4(0u, 0u, 0u, 0u);
          }
          if (GATE && bit_gate) {
            if (c0 + 128 < block_n) bits_fetch(c0 + 128);
          } else if (GATE) {
            // the gate transposition reuses the tile: the store that last used it must have read it
            if (lane == 0) { if (store_bufs == 2) tma_store_wait_read1(); else tma_store_wait_read(); }
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 8; ++it) *reinterpret_cast<uint4*>(swz(tile, it * 4 + g_rsub, g_j)) = gnext[it];
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) gt[j] = *reinterpret_cast<const uint4*>(swz(tile, lane, j));
            __syncwarp();
            if (c0 + 128 < block_n) gate_fetch(c0 + 128);
          }
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t r[32];
            tmem_ld32(taddr + c0 + 32 * hf, r);
            float v[32];
            const float4* bp = reinterpret_cast<const float4*>(s_bias + gc + 32 * hf);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 bb = bp[i];
              v[4 * i] = __uint_as_float(r[4 * i]) + bb.x; v[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + bb.y;
              v[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + bb.z; v[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + bb.w;
            }
            if (!RELU && !GATE && ep.accum != nullptr) add_packed16(v, reinterpret_cast<const uint32_t*>(ain) + 16 * hf, HALF_OUT);
            if (RELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (RELU && ep.relu_bits != nullptr) {   // 1 bit per activation for the backward's gate (word-major: coalesced)
              uint32_t m = 0u;
#pragma unroll
              for (int j = 0; j < 32; ++j) m |= (v[j] > 0.f ? 1u : 0u) << j;
              if (live) ep.relu_bits[static_cast<size_t>((gc >> 5) + hf) * R + row] = m;
            }
            if (GATE && bit_gate) {
              const uint32_t m = gb[hf];
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = (m >> j) & 1u ? v[j] : 0.f;
            } else if (GATE) {
              const uint32_t* gw = reinterpret_cast<const uint32_t*>(gt) + 16 * hf;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const uint32_t w2 = gw[j], lo = w2 & 0xffffu, hi = w2 >> 16;
                if (!((lo & 0x7fffu) != 0 && (lo & 0x8000u) == 0)) v[2 * j] = 0.f;
                if (!((hi & 0x7fffu) != 0 && (hi & 0x8000u) == 0)) v[2 * j + 1] = 0.f;
              }
            }
            uint4 pk[4];
            pack_16(v, HALF_OUT, zero, pk);
            if constexpr (LN) {      // statistics of the ROUNDED values, as LayerNorm on the stored z would see them
              const uint32_t* pw = reinterpret_cast<const uint32_t*>(pk);
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                float a, b;
                unpack16x2(pw[j], HALF_OUT, a, b);
                ln_s1 += a + b;
                ln_s2 = fmaf(a, a, fmaf(b, b, ln_s2));
              }
            }
            if ((!GATE || bit_gate) && hf == 0) {   // wait as late as possible: the previous store's smem read overlaps this step's TMEM load + math
              if (lane == 0) { if (store_bufs == 2) tma_store_wait_read1(); else tma_store_wait_read(); }
              __syncwarp();
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(swz(tile, lane, 4 * hf + i)) = pk[i];
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) tma_store_2d(&tmC, tile, gc, static_cast<int>(row0));
          if (store_bufs == 2) buf ^= 1;
        }
        if constexpr (LN) {
          // the other half of this row's columns belongs to the partner warp of the lane quarter (warps w and w ^ 4 of the
          // eight epilogue warps): exchange the partial sums through shared memory under a 64-thread named barrier
          // (slots alternate with the accumulator buffer: a slot written for tile i is rewritten for tile i + 2, which its
          // writer reaches only after the barrier of tile i + 1, i.e. after the partner has read tile i's slot)
          float* xch = s_bias + 768 + acc * 512;
          const int me = (warp - 2) * 32 + lane, other = ((warp - 2) ^ 4) * 32 + lane;
          xch[2 * me] = ln_s1; xch[2 * me + 1] = ln_s2;
          named_bar(1 + quarter, 64);
          const float t1 = ln_s1 + xch[2 * other], t2 = ln_s2 + xch[2 * other + 1];
          const float inv_n = 1.f / static_cast<float>(Nout);
          const float mean = t1 * inv_n;
          const float rstd = rsqrtf(fmaxf(t2 * inv_n - mean * mean, 0.f) + ep.ln_eps);
          if (chunk_par == 0 && live) { ep.ln_mean[row] = mean; ep.ln_rstd[row] = rstd; }
          for (int c0 = chunk_par * 64; c0 < block_n; c0 += 128) {
            const int gc = n_idx + c0;
            uint8_t* tile = tile0 + buf * TMA_TILE_BYTES;
            uint4 ain[8];
            const bool first = c0 == chunk_par * 64;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              ain[i].x = first ? ln_res0[i].x : ln_res1[i].x; ain[i].y = first ? ln_res0[i].y : ln_res1[i].y;
              ain[i].z = first ? ln_res0[i].z : ln_res1[i].z; ain[i].w = first ? ln_res0[i].w : ln_res1[i].w;
            }
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              uint32_t r[32];
              tmem_ld32(taddr + c0 + 32 * hf, r);
              float v[32];
              const float4* bp = reinterpret_cast<const float4*>(s_bias + gc + 32 * hf);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 bb = bp[i];
                v[4 * i] = __uint_as_float(r[4 * i]) + bb.x; v[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + bb.y;
                v[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + bb.z; v[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + bb.w;
              }
              add_packed16(v, reinterpret_cast<const uint32_t*>(ain) + 16 * hf, HALF_OUT);
              uint4 pk[4];
              pack_16(v, HALF_OUT, false, pk);                      // z exactly as stored by the first pass
              const uint32_t* pw = reinterpret_cast<const uint32_t*>(pk);
              const float* gp = s_bias + 256 + gc + 32 * hf;
              const float* btp = s_bias + 512 + gc + 32 * hf;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                float a, b;
                unpack16x2(pw[j], HALF_OUT, a, b);
                v[2 * j] = fmaf((a - mean) * rstd, gp[2 * j], btp[2 * j]);
                v[2 * j + 1] = fmaf((b - mean) * rstd, gp[2 * j + 1], btp[2 * j + 1]);
              }
              pack_16(v, HALF_OUT, false, pk);
              if (hf == 0) {
                if (lane == 0) { if (store_bufs == 2) tma_store_wait_read1(); else tma_store_wait_read(); }
                __syncwarp();
              }
#pragma unroll
              for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(swz(tile, lane, 4 * hf + i)) = pk[i];
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) tma_store_2d(&tmC2, tile, gc, static_cast<int>(row0));
            if (store_bufs == 2) buf ^= 1;
          }
        }
      } else if (TMA_OUT && MODE == EPI_QUERY) {
        // ---- fp32 sampling locations / softmax weights through TMA: 32 columns (128 bytes per row) per step
        uint8_t* tile0 = tma_tiles + (warp - 2) * store_bufs * TMA_TILE_BYTES;
        int buf = 0;
        for (int c0 = chunk_par * 32; c0 < block_n; c0 += 64) {
          const int gc = n_idx + c0;
          uint8_t* tile = tile0 + buf * TMA_TILE_BYTES;
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
          const bool is_loc = gc < ep.n_loc;
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
          if (lane == 0) { if (store_bufs == 2) tma_store_wait_read1(); else tma_store_wait_read(); }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(swz(tile, lane, i)) = make_float4(o32[4 * i], o32[4 * i + 1], o32[4 * i + 2], o32[4 * i + 3]);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (is_loc) tma_store_2d(&tmC, tile, gc, static_cast<int>(row0));
            else tma_store_2d(&tmC2, tile, gc - ep.n_loc, static_cast<int>(row0));
          }
          if (store_bufs == 2) buf ^= 1;
        }
      } else {
        // GATE: the gate rows of a chunk are fetched one chunk AHEAD with coalesced loads (lanes along the row: 8 rows x 64
        // bytes per instruction) and transposed through the staging buffer when the chunk is processed.
        uint4 gnext[4];
        const int g_c16 = lane & 3, g_rsub = lane >> 2;
        auto gate_fetch = [&](int c) {
          const uint8_t* g = reinterpret_cast<const uint8_t*>(static_cast<const uint16_t*>(ep.gate) + row0 * ep.out_ld + n_idx + c);
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int rr = it * 8 + g_rsub;
            gnext[it] = (rr < rows_valid) ? __ldg(reinterpret_cast<const uint4*>(g + rr * 2ll * ep.out_ld + g_c16 * 16)) : make_uint4(0u, 0u, 0u, 0u);
          }
        };
        const bool bit_gate = GATE && ep.gate_bits != nullptr;
        if (GATE && !bit_gate && chunk_par * 32 < block_n) gate_fetch(chunk_par * 32);
        for (int c0 = chunk_par * 32; c0 < block_n; c0 += 64) {
          const int gc = n_idx + c0;
          uint4 gt[4];
          uint32_t gbw = 0u;
          if (GATE && bit_gate) {
            if (live) gbw = __ldg(ep.gate_bits + static_cast<size_t>(gc >> 5) * R + row);
          } else if (GATE) {
#pragma unroll
            for (int it = 0; it < 4; ++it) *reinterpret_cast<uint4*>(stage + (it * 8 + g_rsub) * 80 + g_c16 * 16) = gnext[it];
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) gt[i] = *reinterpret_cast<const uint4*>(stage + lane * 80 + i * 16);
            __syncwarp();
            if (c0 + 64 < block_n) gate_fetch(c0 + 64);
          }
          uint4 ain[4];
          const bool do_acc = MODE == EPI_STORE && !OUT_F32 && !RELU && !GATE && ep.accum != nullptr;
          if (do_acc) {
            const uint4* ap = reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(ep.accum) + row * ep.out_ld + gc);
#pragma unroll
            for (int i = 0; i < 4; ++i) ain[i] = live ? ap[i] : make_uint4(0u, 0u, 0u, 0u);
          }
          uint32_t r[32];
          tmem_ld32(taddr + c0, r);
          float v[32];
          const float4* bp = reinterpret_cast<const float4*>(s_bias + gc);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = bp[i];
            v[4 * i] = __uint_as_float(r[4 * i]) + bb.x; v[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + bb.y;
            v[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + bb.z; v[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + bb.w;
          }
          if (MODE == EPI_STORE) {
            if (do_acc) add_packed16(v, reinterpret_cast<const uint32_t*>(ain), HALF_OUT);
            if (RELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (RELU && ep.relu_bits != nullptr) {
              uint32_t m = 0u;
#pragma unroll
              for (int j = 0; j < 32; ++j) m |= (v[j] > 0.f ? 1u : 0u) << j;
              if (live) ep.relu_bits[static_cast<size_t>(gc >> 5) * R + row] = m;
            }
            if (GATE && bit_gate) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = (gbw >> j) & 1u ? v[j] : 0.f;
            } else if (GATE) {   // keep where the gate activation is positive: ReLU backward fused into the dgrad
              const uint32_t* gw = reinterpret_cast<const uint32_t*>(gt);
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const uint32_t w2 = gw[j], lo = w2 & 0xffffu, hi = w2 >> 16;
                if (!((lo & 0x7fffu) != 0 && (lo & 0x8000u) == 0)) v[2 * j] = 0.f;
                if (!((hi & 0x7fffu) != 0 && (hi & 0x8000u) == 0)) v[2 * j + 1] = 0.f;
              }
            }
            if (OUT_F32) {
              if (zero) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0.f;
              }
              uint4 pk[8];
              pack_f32(v, pk);
              staged_store<8>(stage, lane, pk, static_cast<float*>(ep.out) + row0 * ep.out_ld + gc, 4ll * ep.out_ld, rows_valid);
            } else {
              uint4 pk[4];
              pack_16(v, HALF_OUT, zero, pk);
              staged_store<4>(stage, lane, pk, static_cast<uint16_t*>(ep.out) + row0 * ep.out_ld + gc, 2ll * ep.out_ld, rows_valid);
            }
          } else if (gc < ep.n_loc) {
            float o32[32];
            const long long rr = live ? row : 0;
            if (ep.L == 4 && ep.P == 4) {
              if (ep.ref_dim == 2) epi_loc_l4p4<2>(ep, v, rr, s_norm, o32);
              else epi_loc_l4p4<4>(ep, v, rr, s_norm, o32);
            } else {
              epi_loc(ep, v, rr, gc, s_norm, o32);
            }
            uint4 pk[8];
            pack_f32(o32, pk);
            staged_store<8>(stage, lane, pk, ep.loc_out + row0 * ep.n_loc + gc, 4ll * ep.n_loc, rows_valid);
          } else {
            epi_softmax(ep, v);
            uint4 pk[8];
            pack_f32(v, pk);
            staged_store<8>(stage, lane, pk, ep.aw_out + row0 * ep.n_aw + (gc - ep.n_loc), 4ll * ep.n_aw, rows_valid);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (TMA_OUT && lane == 0) tma_store_wait_all();
    if (MODE == EPI_ZIRA) {
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        zsum_b += __shfl_xor_sync(0xffffffffu, zsum_b, o);
        zsum_o += __shfl_xor_sync(0xffffffffu, zsum_o, o);
      }
      if (lane == 0) { atomicAdd(ep.loss_sums, zsum_b); atomicAdd(ep.loss_sums + 1, zsum_o); }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// In-place softmax over runs of `lp` consecutive fp32 values (one thread per run) -- used when L*P does not divide the
// 32-column epilogue chunk (L = 5, P = 4).
__global__ void __launch_bounds__(256) softmax_rows_kernel_impl(float* __restrict__ x, long long runs, int lp) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= runs) return;
  float* p = x + i * lp;
  float mx = p[0];
  for (int j = 1; j < lp; ++j) mx = fmaxf(mx, p[j]);
  float sum = 0.f;
  for (int j = 0; j < lp; ++j) sum += __expf(p[j] - mx);
  const float inv = __fdividef(1.f, sum);
  for (int j = 0; j < lp; ++j) p[j] = __expf(p[j] - mx) * inv;
}

thread_local char t_err[256] = "";   // declared in proj_epilogue.cuh (shared with proj_gemm_f32.cu)

int softmax_rows(float* x, long long runs, int lp, cudaStream_t st) {
  ++msda::g_launches;
  softmax_rows_kernel_impl<<<static_cast<unsigned>((runs + 255) / 256), 256, 0, st>>>(x, runs, lp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(t_err, sizeof(t_err), "softmax_rows_kernel: %s", cudaGetErrorString(e)); return static_cast<int>(e); }
  return 0;
}

// ---- host side ------------------------------------------------------------------------------------

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// row-major [rows, cols] matrix of `esize`-byte elements, box = box_rows x box_cols (box_cols * esize == 128),
// 128-byte swizzle.  dtype: 0 bf16, 1 f16, 2 f32.
int make_map(CUtensorMap* map, const void* base, long long rows, long long cols, int box_rows, int box_cols, int dtype) {
  // cuTensorMapEncodeTiled is a DRIVER call: it needs the primary context bound to the calling thread.  A thread that
  // has only ever gone through torch's cudaSetDevice (e.g. an autograd worker) may not have it bound yet -> error 201.
  thread_local bool ctx_bound = false;
  if (!ctx_bound) { cudaFree(nullptr); ctx_bound = true; }
  auto fn = encode_fn();
  if (!fn) { snprintf(t_err, sizeof(t_err), "cuTensorMapEncodeTiled unavailable"); return MSDA_ERR_NO_DEVICE; }
  const int esize = dtype == 2 ? 4 : 2;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(cols) * esize};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = dtype == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                            : (dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  CUresult r = fn(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { snprintf(t_err, sizeof(t_err), "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r)); return MSDA_ERR_BAD_SHAPE; }
  return 0;
}

int make_map3(CUtensorMap* map, const void* base, long long batch, long long rows, long long cols, int box_rows, int dtype) {
  thread_local bool ctx_bound = false;
  if (!ctx_bound) { cudaFree(nullptr); ctx_bound = true; }
  auto fn = encode_fn();
  if (!fn) { snprintf(t_err, sizeof(t_err), "cuTensorMapEncodeTiled unavailable"); return MSDA_ERR_NO_DEVICE; }
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(batch)};
  cuuint64_t gstride[2] = {static_cast<cuuint64_t>(cols) * 2, static_cast<cuuint64_t>(cols) * 2 * static_cast<cuuint64_t>(rows)};
  cuuint32_t box[3] = {64u, static_cast<cuuint32_t>(box_rows), 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapDataType dt = dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = fn(map, dt, 3, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { snprintf(t_err, sizeof(t_err), "cuTensorMapEncodeTiled (3-d) failed (%d)", static_cast<int>(r)); return MSDA_ERR_BAD_SHAPE; }
  return 0;
}

int make_map4(CUtensorMap* map, const void* base, long long d3, long long d2, long long rows, long long cols, int box_rows, int dtype) {
  thread_local bool ctx_bound = false;
  if (!ctx_bound) { cudaFree(nullptr); ctx_bound = true; }
  auto fn = encode_fn();
  if (!fn) { snprintf(t_err, sizeof(t_err), "cuTensorMapEncodeTiled unavailable"); return MSDA_ERR_NO_DEVICE; }
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(d2), static_cast<cuuint64_t>(d3)};
  const cuuint64_t rb = static_cast<cuuint64_t>(cols) * 2;
  cuuint64_t gstride[3] = {rb, rb * static_cast<cuuint64_t>(rows), rb * static_cast<cuuint64_t>(rows) * static_cast<cuuint64_t>(d2)};
  cuuint32_t box[4] = {64u, static_cast<cuuint32_t>(box_rows), 1u, 1u};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapDataType dt = dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = fn(map, dt, 4, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { snprintf(t_err, sizeof(t_err), "cuTensorMapEncodeTiled (4-d) failed (%d)", static_cast<int>(r)); return MSDA_ERR_BAD_SHAPE; }
  return 0;
}

// Widest n-block (multiple of `unit`, dividing Nout, <= 256 columns so two accumulators fit TMEM) whose slice of W
// (block_n x K 16-bit) still fits in shared memory beside the activation ring; the smallest legal block otherwise.
static int pick_block_n(int Nout, int K, int unit, bool f32_out) {
  // tail = everything after the operand buffers; the TMA-store tiles are the largest variant
  const int tail = (f32_out ? AUX_BYTES_32 : AUX_BYTES_16) > AUX_BYTES_TMA ? (f32_out ? AUX_BYTES_32 : AUX_BYTES_16) : AUX_BYTES_TMA;
  int best = 0, smallest = 0;
  for (int bn = unit; bn <= 256 && bn <= Nout; bn += unit) {
    if (Nout % bn) continue;
    if (!smallest) smallest = bn;
    if (bn * K * 2 + 3 * SMEM_A + tail + 1024 <= kMaxSmem) best = bn;     // resident W with at least a 3-deep A ring
  }
  if (best) return best;
  for (int bn = 256; bn >= unit; bn -= unit)
    if (Nout % bn == 0 && bn % unit == 0 && 3 * (SMEM_A + bn * BLOCK_K * 2) + tail + 1024 <= kMaxSmem) return bn;
  return smallest ? smallest : unit;
}

// x2 != nullptr: the activation operand is [x | x2] (K1 and K - K1 columns, both multiples of 64); KB = row length of W
// (K, or K1 == K / 2 for the shared-weight sum, see the kernel).
static int launch(const void* x, const void* w, long long R, int K, int Nout, int block_n, bool half_in,
                  const EpiParams& ep, cudaStream_t st, const void* x2 = nullptr, int K1 = 0, int KB = 0) {
  if (!x || !w) { snprintf(t_err, sizeof(t_err), "null operand"); return MSDA_ERR_NULL_POINTER; }
  if (!x2) { K1 = K; KB = K; }
  if (K1 <= 0 || K1 > K || K1 % BLOCK_K || (x2 && K1 == K) || (KB != K && !(x2 && KB == K1 && 2 * K1 == K)) ||
      (x2 && (reinterpret_cast<uintptr_t>(x2) & 15u))) {
    snprintf(t_err, sizeof(t_err), "unsupported operand split K=%d K1=%d KB=%d", K, K1, KB);
    return MSDA_ERR_UNSUPPORTED;
  }
  if (R <= 0 || R >= (1ll << 31) || K <= 0 || K % BLOCK_K || Nout <= 0 || Nout > MAX_N || block_n % 32 || block_n > 256 ||
      block_n < 32 || Nout % block_n) {
    snprintf(t_err, sizeof(t_err), "unsupported GEMM shape R=%lld K=%d Nout=%d block_n=%d", R, K, Nout, block_n);
    return MSDA_ERR_UNSUPPORTED;
  }
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w)) & 15u) {
    snprintf(t_err, sizeof(t_err), "operands must be 16-byte aligned");
    return MSDA_ERR_MISALIGNED;
  }
  const int in_dt = half_in ? 1 : 0;
  CUtensorMap tmA, tmA2, tmB, tmC, tmC2;
  int rc = make_map(&tmA, x, R, K1, BLOCK_M, BLOCK_K, in_dt);
  if (rc) return rc;
  tmA2 = tmA;
  if (x2) { rc = make_map(&tmA2, x2, R, K - K1, BLOCK_M, BLOCK_K, in_dt); if (rc) return rc; }
  rc = make_map(&tmB, w, Nout, KB, block_n, BLOCK_K, in_dt);
  if (rc) return rc;
  tmC = tmA; tmC2 = tmA;   // placeholders when the epilogue does not store through TMA
  // TMA-store epilogue: plain 16-bit stores in 64-column steps, and the fp32 query outputs in 32-column steps
  bool tma_out = false;
  if (g_tma_store) {
    if (ep.mode == EPI_STORE && !ep.out_f32 && block_n % 64 == 0 && (reinterpret_cast<uintptr_t>(ep.out) & 15u) == 0 &&
        (ep.out_ld * 2) % 16 == 0) {
      rc = make_map(&tmC, ep.out, R, ep.out_ld, 32, 64, ep.out_half ? 1 : 0);
      if (rc) return rc;
      tma_out = true;
      if (ep.ln_gamma != nullptr) {
        rc = make_map(&tmC2, ep.ln_y, R, ep.out_ld, 32, 64, ep.out_half ? 1 : 0);
        if (rc) return rc;
      }
    } else if (ep.mode == EPI_QUERY) {
      rc = make_map(&tmC, ep.loc_out, R, ep.n_loc, 32, 32, 2);
      if (rc) return rc;
      rc = make_map(&tmC2, ep.aw_out, R, ep.n_aw, 32, 32, 2);
      if (rc) return rc;
      tma_out = true;
    }
  }
  int dev_id = 0;
  cudaGetDevice(&dev_id);
  static int sms_of[64] = {};
  if (!sms_of[dev_id & 63]) cudaDeviceGetAttribute(&sms_of[dev_id & 63], cudaDevAttrMultiProcessorCount, dev_id);
  const int sms = sms_of[dev_id & 63];
  const int num_n = Nout / block_n;
  const long long num_m = (R + BLOCK_M - 1) / BLOCK_M;
  const int bres_bytes = block_n * KB * 2;
  const int tail1 = tma_out ? AUX_BYTES_TMA : ((ep.mode == EPI_QUERY || ep.out_f32) ? AUX_BYTES_32 : AUX_BYTES_16);
  const int tail2 = tail1 + EPI_WARPS * TMA_TILE_BYTES;                    // second store tile per epilogue warp
  const int stream_stage = SMEM_A + block_n * BLOCK_K * 2;
  auto fits_res = [&](int tail) { return g_allow_resident && num_n <= sms && bres_bytes + 3 * SMEM_A + tail + 1024 <= kMaxSmem; };
  auto fits_str = [&](int tail) { return 3 * stream_stage + tail + 1024 <= kMaxSmem; };
  bool b_res;
  int store_bufs = 1, tail = tail1;
  if (tma_out && g_store_bufs == 2 && fits_res(tail2)) { b_res = true; store_bufs = 2; tail = tail2; }
  else if (tma_out && g_store_bufs == 2 && !g_prefer_resident && fits_str(tail2)) { b_res = false; store_bufs = 2; tail = tail2; }
  else b_res = fits_res(tail1);
  const int stage_bytes = b_res ? SMEM_A : stream_stage;
  const int fixed = (b_res ? bres_bytes : 0) + tail + 1024;
  int stages = STAGES;
  while (stages > 2 && stages * stage_bytes + fixed > kMaxSmem) --stages;
  if (stages * stage_bytes + fixed > kMaxSmem) { snprintf(t_err, sizeof(t_err), "GEMM tile does not fit shared memory"); return MSDA_ERR_UNSUPPORTED; }
  const int smem = stages * stage_bytes + fixed;
  int grid;
  if (b_res) {
    long long per_n = sms / num_n;
    if (per_n > num_m) per_n = num_m;
    grid = static_cast<int>(per_n) * num_n;
  } else {
    const long long tiles = num_m * num_n;
    grid = static_cast<int>(tiles < sms ? tiles : sms);
  }
  ++msda::g_launches;
  const int Ri = static_cast<int>(R), br = b_res ? 1 : 0, hi = half_in ? 1 : 0;
  cudaError_t cfg = cudaSuccess;
#define PG_LAUNCH(MODE, F32, HALF, RELU, GATE, TMA)                                                                        \
  do {                                                                                                                     \
    static bool configured[64] = {};   /* the attribute is per device */                                                 \
    if (!configured[dev_id & 63]) {                                                                                        \
      cfg = cudaFuncSetAttribute(linear_tc_kernel<MODE, F32, HALF, RELU, GATE, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem); \
      configured[dev_id & 63] = cfg == cudaSuccess;                                                                        \
    }                                                                                                                      \
    if (cfg == cudaSuccess)                                                                                                \
      linear_tc_kernel<MODE, F32, HALF, RELU, GATE, TMA><<<grid, THREADS, smem, st>>>(tmA, tmA2, tmB, tmC, tmC2, Ri, Nout, K, K1 / BLOCK_K, KB / BLOCK_K, block_n, br, stages, store_bufs, hi, ep); \
  } while (0)
#define PG_LAUNCH_LN(HALF)                                                                                                 \
  do {                                                                                                                     \
    static bool configured[64] = {};                                                                                       \
    if (!configured[dev_id & 63]) {                                                                                        \
      cfg = cudaFuncSetAttribute(linear_tc_kernel<EPI_STORE, false, HALF, false, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem); \
      configured[dev_id & 63] = cfg == cudaSuccess;                                                                        \
    }                                                                                                                      \
    if (cfg == cudaSuccess)                                                                                                \
      linear_tc_kernel<EPI_STORE, false, HALF, false, false, true, true><<<grid, THREADS, smem, st>>>(tmA, tmA2, tmB, tmC, tmC2, Ri, Nout, K, K1 / BLOCK_K, KB / BLOCK_K, block_n, br, stages, store_bufs, hi, ep); \
  } while (0)
#define PG_STORE16(RELU, GATE)                                                                                             \
  do {                                                                                                                     \
    if (tma_out) { if (half_out) PG_LAUNCH(EPI_STORE, false, true, RELU, GATE, true); else PG_LAUNCH(EPI_STORE, false, false, RELU, GATE, true); } \
    else { if (half_out) PG_LAUNCH(EPI_STORE, false, true, RELU, GATE, false); else PG_LAUNCH(EPI_STORE, false, false, RELU, GATE, false); }       \
  } while (0)
  const bool half_out = ep.out_half != 0;
  if (ep.ln_gamma != nullptr) {
    // LayerNorm in the epilogue: whole rows in one tile, accumulate form, TMA stores
    if (!(tma_out && ep.mode == EPI_STORE && block_n == Nout && Nout <= 256 && Nout % 128 == 0 && ep.accum && ep.ln_beta && ep.ln_y &&
          ep.ln_mean && ep.ln_rstd && !ep.relu && !ep.gate && !ep.gate_bits && !ep.row_mask && ep.out_ld == Nout &&
          (reinterpret_cast<uintptr_t>(ep.ln_y) & 15u) == 0)) {
      snprintf(t_err, sizeof(t_err), "LayerNorm epilogue needs Nout in {128, 256} as one tile, an accumulate operand and TMA stores");
      return MSDA_ERR_UNSUPPORTED;
    }
    if (half_out) PG_LAUNCH_LN(true); else PG_LAUNCH_LN(false);
  } else if (ep.mode == EPI_QUERY) { if (tma_out) PG_LAUNCH(EPI_QUERY, true, false, false, false, true); else PG_LAUNCH(EPI_QUERY, true, false, false, false, false); }
  else if (ep.mode == EPI_ZIRA) { if (half_out) PG_LAUNCH(EPI_ZIRA, false, true, false, false, false); else PG_LAUNCH(EPI_ZIRA, false, false, false, false, false); }
  else if (ep.out_f32) PG_LAUNCH(EPI_STORE, true, false, false, false, false);
  else if (ep.gate || ep.gate_bits) PG_STORE16(false, true);
  else if (ep.relu) PG_STORE16(true, false);
  else PG_STORE16(false, false);
#undef PG_STORE16
#undef PG_LAUNCH_LN
#undef PG_LAUNCH
  if (cfg != cudaSuccess) { snprintf(t_err, sizeof(t_err), "cudaFuncSetAttribute: %s", cudaGetErrorString(cfg)); return static_cast<int>(cfg); }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(t_err, sizeof(t_err), "linear_tc_kernel launch: %s", cudaGetErrorString(e)); return static_cast<int>(e); }
  return 0;
}

}  // namespace pg

extern "C" int msda_b200_gemm_set_resident(int on) { pg::g_allow_resident = on != 0; return 0; }
extern "C" int msda_b200_gemm_set_staged(int on) { pg::g_tma_store = on != 0; return 0; }
extern "C" int msda_b200_gemm_set_store_bufs(int bufs, int prefer_resident) {
  pg::g_store_bufs = bufs == 2 ? 2 : 1;
  pg::g_prefer_resident = prefer_resident != 0;
  return 0;
}

extern "C" {

const char* msda_b200_gemm_last_error(void) { return pg::t_err; }

int msda_linear_16(const void* x, const void* w, const float* bias, long long R, int K, int Nout, void* out,
                   int out_ld, int out_f32, const uint8_t* row_mask, int is_half, void* stream) {
  pg::t_err[0] = 0;
  if (!out) { snprintf(pg::t_err, sizeof(pg::t_err), "null output"); return MSDA_ERR_NULL_POINTER; }
  pg::EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.mode = pg::EPI_STORE;
  ep.out = out; ep.out_ld = out_ld; ep.out_f32 = out_f32; ep.out_half = is_half; ep.bias = bias; ep.row_mask = row_mask;
  return pg::launch(x, w, R, K, Nout, pg::pick_block_n(Nout, K, 32, out_f32 != 0), is_half != 0, ep, static_cast<cudaStream_t>(stream));
}

int msda_linear_accum_16(const void* x, const void* w, const float* bias, long long R, int K, int Nout, const void* accum,
                         void* out, int is_half, void* stream) {
  pg::t_err[0] = 0;
  if (!out || !accum) { snprintf(pg::t_err, sizeof(pg::t_err), "null pointer"); return MSDA_ERR_NULL_POINTER; }
  if (Nout % 8) { snprintf(pg::t_err, sizeof(pg::t_err), "accumulating store needs Nout %% 8 == 0"); return MSDA_ERR_UNSUPPORTED; }
  pg::EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.mode = pg::EPI_STORE;
  ep.out = out; ep.out_ld = Nout; ep.out_f32 = 0; ep.out_half = is_half; ep.bias = bias; ep.accum = accum;
  return pg::launch(x, w, R, K, Nout, pg::pick_block_n(Nout, K, 32, false), is_half != 0, ep, static_cast<cudaStream_t>(stream));
}

int msda_linear_accum2_16(const void* x1, int K1, const void* x2, int K2, const void* w, const float* bias, long long R, int Nout,
                          const void* accum, void* out, int is_half, void* stream) {
  pg::t_err[0] = 0;
  if (!out || !accum || !x2) { snprintf(pg::t_err, sizeof(pg::t_err), "null pointer"); return MSDA_ERR_NULL_POINTER; }
  if (Nout % 8 || K2 <= 0 || K2 % pg::BLOCK_K) { snprintf(pg::t_err, sizeof(pg::t_err), "accumulating store needs Nout %% 8 == 0, K2 %% 64 == 0"); return MSDA_ERR_UNSUPPORTED; }
  pg::EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.mode = pg::EPI_STORE;
  ep.out = out; ep.out_ld = Nout; ep.out_f32 = 0; ep.out_half = is_half; ep.bias = bias; ep.accum = accum;
  return pg::launch(x1, w, R, K1 + K2, Nout, pg::pick_block_n(Nout, K1 + K2, 32, false), is_half != 0, ep, static_cast<cudaStream_t>(stream),
                    x2, K1, K1 + K2);
}

int msda_linear_add_layernorm_16(const void* x, const void* w, const float* bias, long long R, int K, int Nout, const void* residual,
                                 const float* gamma, const float* beta, float eps, void* z, void* y, float* mean, float* rstd,
                                 int is_half, void* stream) {
  pg::t_err[0] = 0;
  if (!residual || !gamma || !beta || !z || !y || !mean || !rstd) { snprintf(pg::t_err, sizeof(pg::t_err), "null pointer"); return MSDA_ERR_NULL_POINTER; }
  pg::EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.mode = pg::EPI_STORE;
  ep.out = z; ep.out_ld = Nout; ep.out_f32 = 0; ep.out_half = is_half; ep.bias = bias; ep.accum = residual;
  ep.ln_gamma = gamma; ep.ln_beta = beta; ep.ln_eps = eps; ep.ln_y = y; ep.ln_mean = mean; ep.ln_rstd = rstd;
  return pg::launch(x, w, R, K, Nout, Nout, is_half != 0, ep, static_cast<cudaStream_t>(stream));
}

int msda_linear_act_16(const void* x, const void* w, const float* bias, long long R, int K, int Nout, void* out, int relu,
                       const void* gate, int is_half, void* stream) {
  pg::t_err[0] = 0;
  if (!out) { snprintf(pg::t_err, sizeof(pg::t_err), "null output"); return MSDA_ERR_NULL_POINTER; }
  pg::EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.mode = pg::EPI_STORE;
  ep.out = out; ep.out_ld = Nout; ep.out_f32 = 0; ep.out_half = is_half; ep.bias = bias; ep.relu = relu; ep.gate = gate;
  return pg::launch(x, w, R, K, Nout, pg::pick_block_n(Nout, K, 32, false), is_half != 0, ep, static_cast<cudaStream_t>(stream));
}

int msda_linear_act_bits_16(const void* x, const void* w, const float* bias, long long R, int K, int Nout, void* out,
                            uint32_t* relu_bits_out, const uint32_t* gate_bits, int is_half, void* stream) {
  pg::t_err[0] = 0;
  if (!out) { snprintf(pg::t_err, sizeof(pg::t_err), "null output"); return MSDA_ERR_NULL_POINTER; }
  if ((relu_bits_out != nullptr) == (gate_bits != nullptr)) {
    snprintf(pg::t_err, sizeof(pg::t_err), "exactly one of relu_bits_out / gate_bits must be given");
    return MSDA_ERR_BAD_SHAPE;
  }
  if (Nout % 32) { snprintf(pg::t_err, sizeof(pg::t_err), "bit gates need Nout %% 32 == 0"); return MSDA_ERR_UNSUPPORTED; }
  pg::EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.mode = pg::EPI_STORE;
  ep.out = out; ep.out_ld = Nout; ep.out_f32 = 0; ep.out_half = is_half; ep.bias = bias;
  ep.relu = relu_bits_out != nullptr; ep.relu_bits = relu_bits_out; ep.gate_bits = gate_bits;
  return pg::launch(x, w, R, K, Nout, pg::pick_block_n(Nout, K, 32, false), is_half != 0, ep, static_cast<cudaStream_t>(stream));
}

int msda_query_proj2_16(const void* query, const void* query_add, const void* w_cat, const float* bias_cat, const float* ref,
                        int ref_dim, const int64_t* spatial_shapes, long long R, int K, int M, int L, int P, float* loc_out,
                        float* aw_out, int is_half, void* stream);

int msda_query_proj_16(const void* query, const void* w_cat, const float* bias_cat, const float* ref, int ref_dim,
                       const int64_t* spatial_shapes, long long R, int K, int M, int L, int P, float* loc_out,
                       float* aw_out, int is_half, void* stream) {
  return msda_query_proj2_16(query, nullptr, w_cat, bias_cat, ref, ref_dim, spatial_shapes, R, K, M, L, P, loc_out, aw_out, is_half, stream);
}

int msda_query_proj2_16(const void* query, const void* query_add, const void* w_cat, const float* bias_cat, const float* ref,
                        int ref_dim, const int64_t* spatial_shapes, long long R, int K, int M, int L, int P, float* loc_out,
                        float* aw_out, int is_half, void* stream) {
  pg::t_err[0] = 0;
  if (!ref || !spatial_shapes || !loc_out || !aw_out) { snprintf(pg::t_err, sizeof(pg::t_err), "null pointer"); return MSDA_ERR_NULL_POINTER; }
  const int n_aw = M * L * P, n_loc = 2 * n_aw, lp = L * P;
  if ((ref_dim != 2 && ref_dim != 4) || L > MSDA_MAX_LEVELS || n_aw % 32 || 3 * n_aw > pg::MAX_N) {
    snprintf(pg::t_err, sizeof(pg::t_err), "fused query projection needs M*L*P %% 32 == 0 and 3*M*L*P <= %d (L=%d P=%d M=%d)", pg::MAX_N, L, P, M);
    return MSDA_ERR_UNSUPPORTED;
  }
  pg::EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.mode = pg::EPI_QUERY;
  ep.bias = bias_cat; ep.loc_out = loc_out; ep.aw_out = aw_out; ep.ref = ref; ep.shapes = spatial_shapes;
  ep.ref_dim = ref_dim; ep.L = L; ep.P = P; ep.n_loc = n_loc; ep.n_aw = n_aw;
  int rc = query_add ? pg::launch(query, w_cat, R, 2 * K, n_loc + n_aw, pg::pick_block_n(n_loc + n_aw, K, 32, true), is_half != 0, ep,
                                  static_cast<cudaStream_t>(stream), query_add, K, K)
                     : pg::launch(query, w_cat, R, K, n_loc + n_aw, pg::pick_block_n(n_loc + n_aw, K, 32, true), is_half != 0, ep,
                                  static_cast<cudaStream_t>(stream));
  if (rc == 0 && 32 % lp != 0) {   // softmax runs straddle the epilogue chunks: normalise the stored logits in place
    rc = pg::softmax_rows(aw_out, R * M, lp, static_cast<cudaStream_t>(stream));
  }
  return rc;
}

int msda_zira_linear_16(const void* x, const void* w_stack, const float* bias3, const float* scaling, long long R, int K,
                        int F, void* out, const uint8_t* row_mask, void* pre_out, void* adapter_out, float* loss_sums,
                        int is_half, void* stream) {
  pg::t_err[0] = 0;
  if (!out || !bias3 || !scaling || !loss_sums) { snprintf(pg::t_err, sizeof(pg::t_err), "null pointer"); return MSDA_ERR_NULL_POINTER; }
  if (F % 64 || 3 * F > pg::MAX_N) { snprintf(pg::t_err, sizeof(pg::t_err), "ZiRa projection needs F %% 64 == 0 and F <= %d", pg::MAX_N / 3); return MSDA_ERR_UNSUPPORTED; }
  pg::EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.mode = pg::EPI_ZIRA;
  ep.out = out; ep.out_ld = F; ep.out_half = is_half; ep.bias = bias3; ep.row_mask = row_mask;
  ep.scaling = scaling; ep.pre_out = pre_out; ep.adapter_out = adapter_out; ep.loss_sums = loss_sums; ep.F = F;
  return pg::launch(x, w_stack, R, K, 3 * F, pg::pick_block_n(3 * F, K, 96, false), is_half != 0, ep, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
