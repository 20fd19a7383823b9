// Chained feed-forward block on tcgen05: linear1 -> ReLU -> linear2 with the hidden activation kept on chip.
//
// The encoder layer's FFN (reference transformer_for_adapter.py:876-885, d_model 256 -> d_ffn 2048 -> 256) as two
// launches per direction -- linear1 + ReLU writing the [R, 2048] hidden activation (364 MB at 4 images), linear2 reading
// it back -- is bound by that round trip through HBM.  Here one CTA owns a 128-row tile and walks the
// hidden dimension in chunks of 128:
//
//   GEMM1_j   acc1[j & 1] (TMEM, 128 columns)  = X_tile[128, 256] . Wa[j*128 : (j+1)*128, :]^T        (K = 256)
//   mid       registers <- acc1: + bias, ReLU, ONE BIT per activation to HBM (the backward's gate), round to 16 bit,
//             written into shared memory in the K-major 128B-swizzled layout the next product reads   (H_j, 32 KiB)
//   GEMM2_j   acc2 (TMEM, 256 columns)        += H_j[128, 128] . Wb[:, j*128 : (j+1)*128]^T           (K = 128)
//
// and only the [128, 256] result leaves the SM.  The backward is the same chain with the roles swapped
// (dh_j = gate_bits(dz . W2[:, chunk]), dx = dz + dh . W1[chunk, :]): Wa = W2^T, Wb = W1^T, the mid stage applies the bit
// gate instead of bias + ReLU, and the final stage accumulates onto dz -- so the 2048-wide tensor never exists in HBM in
// either direction (2 x 364 MB written + 2 x 364 MB read per layer disappear; what remains is the 23 MB bit mask).
//
//   warp 0      TMA producer: the X tile once per row tile (4 boxes of 128 x 64), then the weight stream through a ring
//               of three 32 KiB slots in the order the products consume it (Wa chunk = 2 slots of two 128 x 64 boxes,
//               Wb chunk = 2 slots of one 256 x 64 box)
//   warp 1      MMA issuer, software-pipelined: GEMM1_{j+1} is issued before GEMM2_j so the tensor core never waits for
//               the mid stage; owns the 512 TMEM columns (2 x 128 + 256)
//   warps 2-9   mid / final stages: two warps per TMEM lane quarter, each 64 of a chunk's 128 columns
//
// Tensor-bound by design (2 x 2 x R x 256 x 2048 FLOP per call); the weights (2 MB per row tile) stream from L2.
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
namespace ffn {

constexpr int C = 256;                   // d_model: K of the first product, N of the second
constexpr int BM = 128, HC = 128;        // rows per tile, hidden columns per chunk
constexpr int X_BYTES = 4 * 16384;       // 4 k-blocks of 128 rows x 128 bytes
constexpr int SLOT = 32768, NSLOT = 3;
constexpr int H_BYTES = 2 * 16384;       // one hidden chunk: 2 k-blocks of 128 rows x 128 bytes
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr int SMEM_BYTES = 1024 + X_BYTES + NSLOT * SLOT + 2 * H_BYTES + 512 + 1024;      // pad, operands, barriers, LayerNorm exchange

struct Params {
  int R, F;                  // rows, hidden width (multiple of 128)
  int backward;              // 0: bias + ReLU + write bits / + bias2;  1: read bits as gate / accumulate onto `accum`
  const float* bias1;        // [F] (forward) or null
  const float* bias2;        // [C] (forward) or null
  uint32_t* bits;            // [F/32, R] word-major: written (forward) / read (backward)
  const void* accum;         // backward: 16-bit [R, C] added to the result (may alias the output)
  int half_in;
  // forward with the layer's residual + LayerNorm fused into the final stage (LN = true): z = x + ffn(x) and y = LN(z)
  const float* ln_gamma;     // [C]
  const float* ln_beta;      // [C]
  float ln_eps;
  float* ln_mean;            // [R]
  float* ln_rstd;            // [R]
};

template <bool BWD, bool HALF, bool LN>
__global__ void __launch_bounds__(THREADS, 1)
ffn_chain_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmA,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
                 const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmXr, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sX = smem;
  uint8_t* sRing = sX + X_BYTES;
  uint8_t* sH = sRing + NSLOT * SLOT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sH + 2 * H_BYTES);
  uint64_t* full = bars;              // [NSLOT]
  uint64_t* empty = full + NSLOT;     // [NSLOT]
  uint64_t* x_full = empty + NSLOT;   // X tile landed
  uint64_t* x_free = x_full + 1;      // last GEMM1 of the tile retired
  uint64_t* a1_full = x_free + 1;     // [2] GEMM1 chunk complete
  uint64_t* a1_free = a1_full + 2;    // [2] mid stage has read it
  uint64_t* h_full = a1_free + 2;     // [2] hidden chunk written to shared memory
  uint64_t* h_free = h_full + 2;      // [2] GEMM2 has read it
  uint64_t* a2_full = h_free + 2;     // result tile complete
  uint64_t* a2_free = a2_full + 1;    // final stage has read it
  uint64_t* res_full = a2_free + 1;   // [EPI_WARPS] LN: a warp's residual pieces have landed in its staging tiles
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + EPI_WARPS);
  float2* sStat = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(bars) + 512);     // [128 rows] LN partial sums / statistics

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunk = p.F / HC;
  const int num_tiles = (p.R + BM - 1) / BM;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    for (int i = 0; i < NSLOT; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    mbar_init(x_full, 1); mbar_init(x_free, 1);
    for (int i = 0; i < EPI_WARPS; ++i) mbar_init(res_full + i, 1);
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

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      int slot = 0;
      uint32_t ph_slot = 0, ph_x = 0;
      auto next_slot = [&]() { if (++slot == NSLOT) { slot = 0; ph_slot ^= 1; } };
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(x_free, ph_x ^ 1);
        ph_x ^= 1;
        mbar_expect_tx(x_full, X_BYTES);
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(&tmX, x_full, sX + kb * 16384, kb * 64, tile * BM);
        for (int j = 0; j <= nchunk; ++j) {
          if (j < nchunk) {               // Wa chunk j: rows [j*128, +128), K = 256 -> two slots of two 128 x 64 boxes
            for (int s2 = 0; s2 < 2; ++s2) {
              mbar_wait(empty + slot, ph_slot ^ 1);
              mbar_expect_tx(full + slot, SLOT);
              tma_load_2d(&tmA, full + slot, sRing + slot * SLOT, (2 * s2) * 64, j * HC);
              tma_load_2d(&tmA, full + slot, sRing + slot * SLOT + 16384, (2 * s2 + 1) * 64, j * HC);
              next_slot();
            }
          }
          if (j >= 1) {                   // Wb chunk j-1: all 256 rows, K columns [(j-1)*128, +128) -> two slots of 256 x 64
            for (int kb2 = 0; kb2 < 2; ++kb2) {
              mbar_wait(empty + slot, ph_slot ^ 1);
              mbar_expect_tx(full + slot, SLOT);
              tma_load_2d(&tmB, full + slot, sRing + slot * SLOT, (j - 1) * HC + kb2 * 64, 0);
              next_slot();
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc1 = umma_idesc(BM, HC, HALF), idesc2 = umma_idesc(BM, C, HALF);
    int slot = 0;
    uint32_t ph_slot = 0, ph_x = 0, ph_a2 = 0;
    uint32_t ph_a1free[2] = {0, 0}, ph_hfull[2] = {0, 0};
    auto next_slot = [&]() { if (++slot == NSLOT) { slot = 0; ph_slot ^= 1; } };
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(x_full, ph_x);
      ph_x ^= 1;
      tc_fence_after();
      for (int j = 0; j <= nchunk; ++j) {
        if (j < nchunk) {
          const int b = j & 1;
          mbar_wait(a1_free + b, ph_a1free[b] ^ 1);
          ph_a1free[b] ^= 1;
          tc_fence_after();
          for (int s2 = 0; s2 < 2; ++s2) {
            mbar_wait(full + slot, ph_slot);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int kb = 2 * s2 + h;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint64_t da = umma_desc_sw128(sX + kb * 16384, k * 32);
                  const uint64_t db = umma_desc_sw128(sRing + slot * SLOT + h * 16384, k * 32);
                  umma_f16(t_acc1 + static_cast<uint32_t>(b * HC), da, db, idesc1, (kb | k) != 0 ? 1u : 0u);
                }
              }
              umma_commit(empty + slot);
              if (s2 == 1) {
                umma_commit(a1_full + b);
                if (j == nchunk - 1) umma_commit(x_free);      // X tile no longer needed
              }
            }
            __syncwarp();
            next_slot();
          }
        }
        if (j >= 1) {
          const int jj = j - 1, b = jj & 1;
          if (jj == 0) {                       // first product into acc2: the previous tile's result must have been read out.
            mbar_wait(a2_free, ph_a2 ^ 1);     // Waiting HERE (not at the top of the tile) lets GEMM1_0 / GEMM1_1 of this tile
            ph_a2 ^= 1;                        // overlap the previous tile's final stage.
          }
          mbar_wait(h_full + b, ph_hfull[b]);
          ph_hfull[b] ^= 1;
          tc_fence_after();
          for (int kb2 = 0; kb2 < 2; ++kb2) {
            mbar_wait(full + slot, ph_slot);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t da = umma_desc_sw128(sH + b * H_BYTES + kb2 * 16384, k * 32);
                const uint64_t db = umma_desc_sw128(sRing + slot * SLOT, k * 32);
                umma_f16(t_acc2, da, db, idesc2, (jj | kb2 | k) != 0 ? 1u : 0u);
              }
              umma_commit(empty + slot);
              if (kb2 == 1) {
                umma_commit(h_free + b);
                if (jj == nchunk - 1) umma_commit(a2_full);
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
    const int quarter = warp & 3, half = (warp - 2) >> 2;     // TMEM lane quarter; which 64 of a chunk's 128 columns
    uint32_t ph_a1full[2] = {0, 0}, ph_hfree[2] = {0, 0}, ph_a2 = 0, ph_res = 0;
    // final stage: per-warp 32 x 128-byte store tile inside H[0] (idle by then) -- the SAME 4 KiB this warp writes in the mid
    // stage (k-block `half`, rows of `quarter`), so only the warp itself ever reuses it, after its own store has been read
    uint8_t* stile = sH + (half * 4 + quarter) * 4096;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const long long row0 = static_cast<long long>(tile) * BM + quarter * 32;
      const long long row = row0 + lane;
      const bool live = row < p.R;
      for (int j = 0; j < nchunk; ++j) {
        const int b = j & 1;
        mbar_wait(a1_full + b, ph_a1full[b]);
        ph_a1full[b] ^= 1;
        tc_fence_after();
        mbar_wait(h_free + b, ph_hfree[b] ^ 1);     // GEMM2 of chunk j-2 has finished reading this H buffer
        ph_hfree[b] ^= 1;
        uint8_t* hk = sH + b * H_BYTES + half * 16384;        // this warp's k-block of the hidden chunk
        const int col0 = j * HC + half * 64;                  // first hidden column this thread handles
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t r[32];
          tmem_ld32(t_acc1 + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(b * HC + half * 64 + 32 * hf), r);
          float v[32];
          const int gc = col0 + 32 * hf;
          if (!BWD) {
            const float4* bp = reinterpret_cast<const float4*>(p.bias1 + gc);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 bb = __ldg(bp + i);
              v[4 * i] = __uint_as_float(r[4 * i]) + bb.x; v[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + bb.y;
              v[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + bb.z; v[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + bb.w;
            }
            const uint32_t m = positive_mask32(v);          // the ReLU itself is fused into the 16-bit conversion below
            if (live && p.bits) p.bits[static_cast<size_t>(gc >> 5) * p.R + row] = m;
          } else {
            const uint32_t m = live ? __ldg(p.bits + static_cast<size_t>(gc >> 5) * p.R + row) : 0u;
#pragma unroll
            for (int jx = 0; jx < 32; ++jx) v[jx] = (m >> jx) & 1u ? __uint_as_float(r[jx]) : 0.f;
          }
          uint4 pk[4];
          if (!BWD) pack_16_relu(v, HALF, pk); else pack_16(v, HALF, false, pk);
#pragma unroll
          for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(swz(hk, quarter * 32 + lane, 4 * hf + i)) = pk[i];
        }
        tc_fence_before();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) { mbar_arrive(h_full + b); mbar_arrive(a1_free + b); }
      }
      // ---- final stage: acc2 (+ bias2 | + accum) -> 16 bit -> TMA store; this warp's 128 of the 256 columns ----
      mbar_wait(a2_full, ph_a2);
      ph_a2 ^= 1;
      tc_fence_after();
      if (!LN) {
#pragma unroll 1
      for (int g = 0; g < 2; ++g) {
        const int gc = half * 128 + g * 64;
        uint4 ain[8];
        if (BWD && p.accum != nullptr) {
          const uint4* ap = reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(p.accum) + row * C + gc);
#pragma unroll
          for (int i = 0; i < 8; ++i) ain[i] = live ? ap[i] : make_uint4(0u, 0u, 0u, 0u);
        }
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t r[32];
          tmem_ld32(t_acc2 + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(gc + 32 * hf), r);
          float v[32];
#pragma unroll
          for (int jx = 0; jx < 32; ++jx) v[jx] = __uint_as_float(r[jx]);
          if (!BWD && p.bias2 != nullptr) {
            const float4* bp = reinterpret_cast<const float4*>(p.bias2 + gc + 32 * hf);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 bb = __ldg(bp + i);
              v[4 * i] += bb.x; v[4 * i + 1] += bb.y; v[4 * i + 2] += bb.z; v[4 * i + 3] += bb.w;
            }
          }
          if (BWD && p.accum != nullptr) add_packed16(v, reinterpret_cast<const uint32_t*>(ain) + 16 * hf, HALF);
          uint4 pk[4];
          pack_16(v, HALF, false, pk);
#pragma unroll
          for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(swz(stile, lane, 4 * hf + i)) = pk[i];
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) tma_store_2d(&tmOut, stile, gc, static_cast<int>(row0));
      }
      } else {
        // ---- residual + LayerNorm in the final stage (reference transformer_for_adapter.py:882-885): z = x + ffn(x) rounded to
        // 16 bit (what LayerNorm's backward keeps), y = LN(z).  Each warp fetches its 32 x 128 piece of the residual x by TMA
        // straight into its two staging tiles (the X operand tile is NOT held back: the next tile's products overlap this
        // stage), turns it into z in place, stores z, and -- after the two warps of a lane quarter have exchanged partial sums
        // of the ROUNDED values -- normalises the packed z row it kept in registers and stores y.
        uint8_t* stile1 = stile + H_BYTES;              // this warp's 4 KiB of H[1]
        if (lane == 0) {
          tma_store_wait_read();                        // the previous tile's stores have read both staging tiles
          mbar_expect_tx(res_full + (warp - 2), 2 * 4096);
          tma_load_2d(&tmXr, res_full + (warp - 2), stile, half * 128, static_cast<int>(row0));
          tma_load_2d(&tmXr, res_full + (warp - 2), stile1, half * 128 + 64, static_cast<int>(row0));
        }
        __syncwarp();
        mbar_wait(res_full + (warp - 2), ph_res);
        ph_res ^= 1;
        uint4 zk[2][2][4];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int gc = half * 128 + g * 64;
          uint8_t* st = g ? stile1 : stile;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t r[32];
            float v[32];
            tmem_ld32(t_acc2 + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(gc + 32 * hf), r);
            const float4* bp = reinterpret_cast<const float4*>(p.bias2 + gc + 32 * hf);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 bb = __ldg(bp + i);
              v[4 * i] = __uint_as_float(r[4 * i]) + bb.x; v[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + bb.y;
              v[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + bb.z; v[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + bb.w;
            }
            uint4 xin[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) xin[i] = *reinterpret_cast<const uint4*>(swz(st, lane, 4 * hf + i));
            add_packed16(v, reinterpret_cast<const uint32_t*>(xin), HALF);
            pack_16(v, HALF, false, zk[g][hf]);
            float zr[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) zr[i] = 0.f;
            add_packed16(zr, reinterpret_cast<const uint32_t*>(zk[g][hf]), HALF);
#pragma unroll
            for (int i = 0; i < 32; ++i) { s1 += zr[i]; s2 = fmaf(zr[i], zr[i], s2); }
#pragma unroll
            for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(swz(st, lane, 4 * hf + i)) = zk[g][hf][i];
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) tma_store_2d(&tmZ, st, gc, static_cast<int>(row0));
        }
        // half 1 hands its partial sums to half 0, which returns mean / rstd through the same slots
        const int trow = quarter * 32 + lane;
        if (half == 1) sStat[trow] = make_float2(s1, s2);
        named_bar(1 + quarter, 64);
        float mean = 0.f, rstd = 0.f;
        if (half == 0) {
          const float2 o = sStat[trow];
          mean = (s1 + o.x) * (1.f / C);
          rstd = rsqrtf(fmaxf((s2 + o.y) * (1.f / C) - mean * mean, 0.f) + p.ln_eps);
          sStat[trow] = make_float2(mean, rstd);
          if (live) { p.ln_mean[row] = mean; p.ln_rstd[row] = rstd; }
        }
        named_bar(1 + quarter, 64);
        if (half == 1) { const float2 o = sStat[trow]; mean = o.x; rstd = o.y; }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int gc = half * 128 + g * 64;
          uint8_t* st = g ? stile1 : stile;
          if (lane == 0) tma_store_wait_read1();        // both z stores are older than the one still allowed in flight
          __syncwarp();
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            float zr[32], v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) zr[i] = 0.f;
            add_packed16(zr, reinterpret_cast<const uint32_t*>(zk[g][hf]), HALF);
            const float4* gp = reinterpret_cast<const float4*>(p.ln_gamma + gc + 32 * hf);
            const float4* bp = reinterpret_cast<const float4*>(p.ln_beta + gc + 32 * hf);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 gg = __ldg(gp + i), bb = __ldg(bp + i);
              v[4 * i] = fmaf((zr[4 * i] - mean) * rstd, gg.x, bb.x); v[4 * i + 1] = fmaf((zr[4 * i + 1] - mean) * rstd, gg.y, bb.y);
              v[4 * i + 2] = fmaf((zr[4 * i + 2] - mean) * rstd, gg.z, bb.z); v[4 * i + 3] = fmaf((zr[4 * i + 3] - mean) * rstd, gg.w, bb.w);
            }
            uint4 pk[4];
            pack_16(v, HALF, false, pk);
#pragma unroll
            for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(swz(st, lane, 4 * hf + i)) = pk[i];
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) tma_store_2d(&tmOut, st, gc, static_cast<int>(row0));
        }
        named_bar(1 + quarter, 64);                     // the exchange slots are free for the next tile
      }
      if (lane == 0) tma_store_wait_read();      // the store tiles live in the H buffers: drained before the next tile's mid stage
      __syncwarp();
      tc_fence_before();
      if (lane == 0) mbar_arrive(a2_free);
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

static int launch(const void* x, const void* wa, const void* wb, void* out, void* z_out, const Params& p, cudaStream_t st) {
  if (!x || !wa || !wb || !out) { snprintf(t_err, sizeof(t_err), "null pointer"); return MSDA_ERR_NULL_POINTER; }
  if (p.R <= 0 || p.F <= 0 || p.F % HC || p.F > 8192) { snprintf(t_err, sizeof(t_err), "chained FFN needs d_ffn %% 128 == 0 (R=%d F=%d)", p.R, p.F); return MSDA_ERR_UNSUPPORTED; }
  const int dt = p.half_in ? 1 : 0;
  CUtensorMap tmX, tmA, tmB, tmOut, tmZ, tmXr;
  int rc = make_map(&tmX, x, p.R, C, BM, 64, dt);
  if (rc) return rc;
  rc = make_map(&tmA, wa, p.F, C, HC, 64, dt);
  if (rc) return rc;
  rc = make_map(&tmB, wb, C, p.F, C, 64, dt);
  if (rc) return rc;
  rc = make_map(&tmOut, out, p.R, C, 32, 64, dt);
  if (rc) return rc;
  rc = make_map(&tmZ, z_out ? z_out : out, p.R, C, 32, 64, dt);
  if (rc) return rc;
  rc = make_map(&tmXr, x, p.R, C, 32, 64, dt);      // the residual, in the final stage's 32-row pieces
  if (rc) return rc;
  int dev_id = 0;
  cudaGetDevice(&dev_id);
  static int sms_of[64] = {};
  if (!sms_of[dev_id & 63]) cudaDeviceGetAttribute(&sms_of[dev_id & 63], cudaDevAttrMultiProcessorCount, dev_id);
  const int tiles = (p.R + BM - 1) / BM;
  const int grid = tiles < sms_of[dev_id & 63] ? tiles : sms_of[dev_id & 63];
  cudaError_t cfg = cudaSuccess;
  ++msda::g_launches;
#define FFN_LAUNCH(BWD, HALF, LN)                                                                                          \
  do {                                                                                                                      \
    static bool configured[64] = {};                                                                                        \
    if (!configured[dev_id & 63]) {                                                                                         \
      cfg = cudaFuncSetAttribute(ffn_chain_kernel<BWD, HALF, LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);  \
      configured[dev_id & 63] = cfg == cudaSuccess;                                                                         \
    }                                                                                                                       \
    if (cfg == cudaSuccess) ffn_chain_kernel<BWD, HALF, LN><<<grid, THREADS, SMEM_BYTES, st>>>(tmX, tmA, tmB, tmOut, tmZ, tmXr, p); \
  } while (0)
  if (p.backward) { if (p.half_in) FFN_LAUNCH(true, true, false); else FFN_LAUNCH(true, false, false); }
  else if (z_out) { if (p.half_in) FFN_LAUNCH(false, true, true); else FFN_LAUNCH(false, false, true); }
  else { if (p.half_in) FFN_LAUNCH(false, true, false); else FFN_LAUNCH(false, false, false); }
#undef FFN_LAUNCH
  if (cfg != cudaSuccess) { snprintf(t_err, sizeof(t_err), "cudaFuncSetAttribute: %s", cudaGetErrorString(cfg)); return static_cast<int>(cfg); }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(t_err, sizeof(t_err), "ffn_chain_kernel launch: %s", cudaGetErrorString(e)); return static_cast<int>(e); }
  return 0;
}

}  // namespace ffn
}  // namespace pg

extern "C" {

int msda_ffn_chain_fwd_16(const void* x, const void* w1, const float* b1, const void* w2, const float* b2, long long R, int C,
                          int F, void* out, uint32_t* relu_bits_out, int is_half, void* stream) {
  pg::t_err[0] = 0;
  if (C != pg::ffn::C) { snprintf(pg::t_err, sizeof(pg::t_err), "chained FFN is built for d_model = %d (got %d)", pg::ffn::C, C); return MSDA_ERR_UNSUPPORTED; }
  if (!b1) { snprintf(pg::t_err, sizeof(pg::t_err), "null bias"); return MSDA_ERR_NULL_POINTER; }
  if (R >= (1ll << 31)) return MSDA_ERR_BAD_SHAPE;
  pg::ffn::Params p;
  memset(&p, 0, sizeof(p));
  p.R = static_cast<int>(R); p.F = F; p.backward = 0; p.bias1 = b1; p.bias2 = b2; p.bits = relu_bits_out; p.half_in = is_half;
  return pg::ffn::launch(x, w1, w2, out, nullptr, p, static_cast<cudaStream_t>(stream));
}

// z = x + (relu(x W1^T + b1) W2^T + b2) rounded to 16 bit, y = LayerNorm(z) * gamma + beta, mean / rstd [R] fp32: the FFN half of an
// encoder layer including `src = norm2(src + src2)` (transformer_for_adapter.py:877-885) in one launch.
int msda_ffn_chain_ln_fwd_16(const void* x, const void* w1, const float* b1, const void* w2, const float* b2, long long R, int C,
                             int F, const float* gamma, const float* beta, float eps, void* z, void* y, float* mean, float* rstd,
                             uint32_t* relu_bits_out, int is_half, void* stream) {
  pg::t_err[0] = 0;
  if (C != pg::ffn::C) { snprintf(pg::t_err, sizeof(pg::t_err), "chained FFN is built for d_model = %d (got %d)", pg::ffn::C, C); return MSDA_ERR_UNSUPPORTED; }
  if (!b1 || !b2 || !gamma || !beta || !z || !mean || !rstd) { snprintf(pg::t_err, sizeof(pg::t_err), "null pointer"); return MSDA_ERR_NULL_POINTER; }
  if (R >= (1ll << 31)) return MSDA_ERR_BAD_SHAPE;
  pg::ffn::Params p;
  memset(&p, 0, sizeof(p));
  p.R = static_cast<int>(R); p.F = F; p.backward = 0; p.bias1 = b1; p.bias2 = b2; p.bits = relu_bits_out; p.half_in = is_half;
  p.ln_gamma = gamma; p.ln_beta = beta; p.ln_eps = eps; p.ln_mean = mean; p.ln_rstd = rstd;
  return pg::ffn::launch(x, w1, w2, y, z, p, static_cast<cudaStream_t>(stream));
}

int msda_ffn_chain_bwd_16(const void* dy, const void* w2_t, const void* w1_t, const uint32_t* gate_bits, const void* accum,
                          long long R, int C, int F, void* dx, int is_half, void* stream) {
  pg::t_err[0] = 0;
  if (C != pg::ffn::C) { snprintf(pg::t_err, sizeof(pg::t_err), "chained FFN is built for d_model = %d (got %d)", pg::ffn::C, C); return MSDA_ERR_UNSUPPORTED; }
  if (!gate_bits) { snprintf(pg::t_err, sizeof(pg::t_err), "null gate bits"); return MSDA_ERR_NULL_POINTER; }
  if (R >= (1ll << 31)) return MSDA_ERR_BAD_SHAPE;
  pg::ffn::Params p;
  memset(&p, 0, sizeof(p));
  p.R = static_cast<int>(R); p.F = F; p.backward = 1; p.bits = const_cast<uint32_t*>(gate_bits); p.accum = accum; p.half_in = is_half;
  return pg::ffn::launch(dy, w2_t, w1_t, dx, nullptr, p, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
