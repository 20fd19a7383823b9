// Chained feed-forward block on CTA PAIRS (tcgen05 cta_group::2): the kernel of layer_ffn_chain.cu with M = 256.
//
// layer_ffn_chain.cu is bound by streaming the weights from L2: every 128-row tile re-reads all of W1 and W2 (2 MB), which
// asks each SM for ~64 B/clk when the chip delivers ~42.  Here two CTAs of a cluster (one TPC) work as one:
// each keeps ITS OWN 128-row X tile and hidden chunk (the A operands, split by rows), loads only HALF of every weight
// tile (the B operands, split by N: 64 of a W1 chunk's 128 rows, 128 of W2's 256), and the leader issues
// tcgen05.mma.cta_group::2 instructions of M = 256 that read both CTAs' shared memory and write both CTAs' tensor memory.
// Weight traffic per SM halves; everything else (mid stage, 1-bit ReLU mask, final stage, fused residual + LayerNorm)
// is the single-CTA kernel's code, run by each CTA on its own rows.
//
//   both CTAs   TMA producer (own X tile, own half of each weight tile; the transaction bytes are counted on the LEADER's
//               barriers: cp.async.bulk.tensor...cta_group::2 with the peer bit of the barrier address cleared)
//   leader      MMA issuer; tcgen05.commit...multicast::cluster releases ring slots / accumulators in BOTH CTAs
//   both CTAs   8 mid / final warps; "consumed" arrivals (a1_free, h_full, a2_free) go to the leader's barriers (count 16)
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
namespace ffn2 {

constexpr int C = 256;                   // d_model: K of the first product, N of the second
constexpr int BM = 128, HC = 128;        // rows per CTA tile, hidden columns per chunk
constexpr int X_BYTES = 4 * 16384;       // 4 k-blocks of 128 rows x 128 bytes
constexpr int SLOT = 32768, NSLOT = 3;   // one slot = this CTA's half of a weight chunk (W1: 64 x 256, W2: 128 x 128)
constexpr int H_BYTES = 2 * 16384;       // one hidden chunk: 2 k-blocks of 128 rows x 128 bytes
constexpr int EW = 4;                    // mid / final warps per TMEM lane quarter
constexpr int EPI_WARPS = 4 * EW;
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr int SMEM_BYTES = 1024 + X_BYTES + NSLOT * SLOT + 2 * H_BYTES + 512 + 1024;

struct Params {
  int R, F;
  int backward;
  const float* bias1;
  const float* bias2;
  uint32_t* bits;
  const void* accum;
  int half_in;
  const float* ln_gamma;
  const float* ln_beta;
  float ln_eps;
  float* ln_mean;
  float* ln_rstd;
  long long* trace;          // -DFFN2_TRACE: 8 cycle counters per cluster (MMA issuer waits)
};

#ifdef FFN2_TRACE
#define TW(bar, par, ctr) do { const long long t0_ = clock64(); mbar_wait(bar, par); (ctr) += clock64() - t0_; } while (0)
#else
#define TW(bar, par, ctr) mbar_wait(bar, par)
#endif

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the LEADER's copy of a barrier (rank 0 of the pair), from either CTA.  RELAXED on purpose: `.release.cluster`
// compiles to MEMBAR.ALL.GPU in front of the arrive (~2 200 cycles per hidden chunk, measured: more than the products take).
// What the leader's products read after this signal is this CTA's own shared memory through the async proxy, and those
// writes have been ordered by the fence.proxy.async every caller issues first; TMEM reads are ordered by tcgen05.fence.
__device__ __forceinline__ void arrive_leader(uint64_t* bar) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(ra) : "r"(smem_u32(bar)));
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
// TMA load whose transaction bytes are counted on the leader's barrier (same offset; the peer bit of the address cleared)
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// completion of all prior MMAs arrives on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3)) : "memory");
}

template <bool BWD, bool HALF, bool LN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
ffn_chain2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmA,
                  const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
                  const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmXr, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sX = smem;
  uint8_t* sRing = sX + X_BYTES;
  uint8_t* sH = sRing + NSLOT * SLOT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sH + 2 * H_BYTES);
  uint64_t* full = bars;              // [NSLOT]  leader: both CTAs' halves of the slot have landed
  uint64_t* empty = full + NSLOT;     // [NSLOT]  each CTA: the products reading the slot have retired
  uint64_t* x_full = empty + NSLOT;   // leader: both X tiles landed
  uint64_t* x_free = x_full + 1;      // each CTA: last GEMM1 of the tile pair retired
  uint64_t* a1_full = x_free + 1;     // [2] each CTA: GEMM1 chunk complete
  uint64_t* a1_free = a1_full + 2;    // [2] leader: both CTAs' mid stages have read it
  uint64_t* h_full = a1_free + 2;     // [2] leader: both CTAs' hidden chunks written
  uint64_t* h_free = h_full + 2;      // [2] each CTA: GEMM2 has read it
  uint64_t* a2_full = h_free + 2;     // each CTA: result complete
  uint64_t* a2_free = a2_full + 1;    // leader: both CTAs' final stages have read it
  uint64_t* res_full = a2_free + 1;   // [EPI_WARPS] LN: a warp's residual piece has landed in its staging tile (CTA-local)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + EPI_WARPS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_rank());
  const int nchunk = p.F / HC;
  const int num_tiles = (p.R + BM - 1) / BM;
  const int num_pairs = (num_tiles + 1) / 2;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    for (int i = 0; i < NSLOT; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    mbar_init(x_full, 1); mbar_init(x_free, 1);
    for (int i = 0; i < EPI_WARPS; ++i) mbar_init(res_full + i, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(a1_full + i, 1); mbar_init(a1_free + i, 2 * EPI_WARPS); mbar_init(h_full + i, 2 * EPI_WARPS); mbar_init(h_free + i, 1); }
    mbar_init(a2_full, 1); mbar_init(a2_free, 2 * EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {      // the same warp of BOTH CTAs, same destination offset
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();      // barriers of the peer are initialised before anything signals them; includes the CTA-wide sync
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_acc1 = tmem_base, t_acc2 = tmem_base + 256u;

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====
    if (elect_one()) {
      int slot = 0;
      uint32_t ph_slot = 0, ph_x = 0;
      auto next_slot = [&]() { if (++slot == NSLOT) { slot = 0; ph_slot ^= 1; } };
      for (int pair = blockIdx.x >> 1; pair < num_pairs; pair += gridDim.x >> 1) {
        const int tile = 2 * pair + rank;
        mbar_wait(x_free, ph_x ^ 1);
        ph_x ^= 1;
        if (rank == 0) mbar_expect_tx(x_full, 2 * X_BYTES);
        for (int kb = 0; kb < 4; ++kb) tma_load_2d_pair(&tmX, x_full, sX + kb * 16384, kb * 64, tile * BM);
        for (int j = 0; j <= nchunk; ++j) {
          if (j < nchunk) {               // this CTA's 64 of the 128 rows of Wa chunk j, K = 256: four 64 x 64 boxes
            mbar_wait(empty + slot, ph_slot ^ 1);
            if (rank == 0) mbar_expect_tx(full + slot, 2 * SLOT);
            for (int kb = 0; kb < 4; ++kb) tma_load_2d_pair(&tmA, full + slot, sRing + slot * SLOT + kb * 8192, kb * 64, j * HC + rank * 64);
            next_slot();
          }
          if (j >= 1) {                   // this CTA's 128 of the 256 rows of Wb, K columns of chunk j-1: two 128 x 64 boxes
            mbar_wait(empty + slot, ph_slot ^ 1);
            if (rank == 0) mbar_expect_tx(full + slot, 2 * SLOT);
            for (int kb2 = 0; kb2 < 2; ++kb2) tma_load_2d_pair(&tmB, full + slot, sRing + slot * SLOT + kb2 * 16384, (j - 1) * HC + kb2 * 64, rank * 128);
            next_slot();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only): M = 256 over the pair =====
    if (rank == 0) {
    const uint32_t idesc1 = umma_idesc(2 * BM, HC, HALF), idesc2 = umma_idesc(2 * BM, C, HALF);
    int slot = 0;
    uint32_t ph_slot = 0, ph_x = 0, ph_a2 = 0;
    uint32_t ph_a1free[2] = {0, 0}, ph_hfull[2] = {0, 0};
    long long tc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#ifdef FFN2_TRACE
    const long long t_begin = clock64();
#endif
    auto next_slot = [&]() { if (++slot == NSLOT) { slot = 0; ph_slot ^= 1; } };
    for (int pair = blockIdx.x >> 1; pair < num_pairs; pair += gridDim.x >> 1) {
      TW(x_full, ph_x, tc[0]);
      ph_x ^= 1;
      tc_fence_after();
      for (int j = 0; j <= nchunk; ++j) {
        if (j < nchunk) {
          const int b = j & 1;
          TW(a1_free + b, ph_a1free[b] ^ 1, tc[1]);
          ph_a1free[b] ^= 1;
          tc_fence_after();
          TW(full + slot, ph_slot, tc[2]);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int kb = 0; kb < 4; ++kb) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t da = umma_desc_sw128(sX + kb * 16384, k * 32);
                const uint64_t db = umma_desc_sw128(sRing + slot * SLOT + kb * 8192, k * 32);
                umma2_f16(t_acc1 + static_cast<uint32_t>(b * HC), da, db, idesc1, (kb | k) != 0 ? 1u : 0u);
              }
            }
            umma2_commit(empty + slot);
            umma2_commit(a1_full + b);
            if (j == nchunk - 1) umma2_commit(x_free);      // X tiles no longer needed
          }
          __syncwarp();
          next_slot();
        }
        if (j >= 1) {
          const int jj = j - 1, b = jj & 1;
          if (jj == 0) {
            TW(a2_free, ph_a2 ^ 1, tc[3]);
            ph_a2 ^= 1;
          }
          TW(h_full + b, ph_hfull[b], tc[4]);
          ph_hfull[b] ^= 1;
          tc_fence_after();
          TW(full + slot, ph_slot, tc[5]);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int kb2 = 0; kb2 < 2; ++kb2) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t da = umma_desc_sw128(sH + b * H_BYTES + kb2 * 16384, k * 32);
                const uint64_t db = umma_desc_sw128(sRing + slot * SLOT + kb2 * 16384, k * 32);
                umma2_f16(t_acc2, da, db, idesc2, (jj | kb2 | k) != 0 ? 1u : 0u);
              }
            }
            umma2_commit(empty + slot);
            umma2_commit(h_free + b);
            if (jj == nchunk - 1) umma2_commit(a2_full);
          }
          __syncwarp();
          next_slot();
        }
      }
    }
#ifdef FFN2_TRACE
    if (p.trace != nullptr && lane == 0) {
      tc[6] = clock64() - t_begin;
      for (int i = 0; i < 8; ++i) p.trace[(blockIdx.x >> 1) * 8 + i] = tc[i];
    }
#endif
    (void)tc;
    }
  } else {
    // ===== mid / final stages: EW warps per TMEM lane quarter, each HC/EW hidden columns (mid) and C/EW output columns (final).
    // The single-CTA kernel turned out to be bound by the LATENCY of this stage (the issuer waits for h_full more than half of
    // the time once the weights arrive fast enough), so the pair kernel runs it on twice as many warps. =====
    const int quarter = warp & 3, slice = (warp - 2) >> 2;
    uint32_t ph_a1full[2] = {0, 0}, ph_hfree[2] = {0, 0}, ph_a2 = 0, ph_res = 0;
    const uint32_t lane_bits = static_cast<uint32_t>(quarter * 32) << 16;
    long long te[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#ifdef FFN2_TRACE
    const long long te_begin = clock64();
#define TSEG(i) do { const long long t1_ = clock64(); te[i] += t1_ - tseg_; tseg_ = t1_; } while (0)
#else
#define TSEG(i) do {} while (0)
#endif
    uint8_t* stile = sH + (slice * 4 + quarter) * 4096;       // final-stage staging: 4 KiB per warp, all of H[0] and H[1]
    for (int pair = blockIdx.x >> 1; pair < num_pairs; pair += gridDim.x >> 1) {
      const int tile = 2 * pair + rank;
      const long long row0 = static_cast<long long>(tile) * BM + quarter * 32;
      const long long row = row0 + lane;
      const bool live = row < p.R;
      for (int j = 0; j < nchunk; ++j) {
        const int b = j & 1;
#ifdef FFN2_TRACE
        long long tseg_ = clock64();
#endif
        // backward: this chunk's gate word is fetched before waiting for its accumulator (an L2 round trip otherwise)
        uint32_t gate = 0u;
        if (BWD && live) gate = __ldg(p.bits + static_cast<size_t>((j * HC + slice * (HC / EW)) >> 5) * p.R + row);
        mbar_wait(a1_full + b, ph_a1full[b]);
        ph_a1full[b] ^= 1;
        tc_fence_after();
        TSEG(0);
        mbar_wait(h_free + b, ph_hfree[b] ^ 1);     // GEMM2 of chunk j-2 has finished reading this H buffer
        ph_hfree[b] ^= 1;
        TSEG(1);
        constexpr int MC = HC / EW;                  // 32 hidden columns per warp
        static_assert(MC == 32, "one TMEM load per warp and chunk");
        uint8_t* hk = sH + b * H_BYTES + ((slice * MC) >> 6) * 16384;
        const int chunk0 = ((slice * MC) & 63) >> 3;          // first 16-byte piece inside the k-block's 128-byte row
        const int gc = j * HC + slice * MC;
        uint32_t r[32];
        tmem_ld32(t_acc1 + lane_bits + static_cast<uint32_t>(b * HC + slice * MC), r);
        TSEG(2);
        float v[32];
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
          const uint32_t m = gate;
#pragma unroll
          for (int jx = 0; jx < 32; ++jx) v[jx] = (m >> jx) & 1u ? __uint_as_float(r[jx]) : 0.f;
        }
        uint4 pk[4];
        if (!BWD) pack_16_relu(v, HALF, pk); else pack_16(v, HALF, false, pk);
#pragma unroll
        for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(swz(hk, quarter * 32 + lane, chunk0 + i)) = pk[i];
        TSEG(3);
        tc_fence_before();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) { arrive_leader(h_full + b); arrive_leader(a1_free + b); }
        TSEG(4);
      }
      // ---- final stage: acc2 (+ bias2 | + accum) -> 16 bit -> TMA store; this warp's 64 of the 256 columns ----
#ifdef FFN2_TRACE
      long long tseg_ = clock64();
#endif
      mbar_wait(a2_full, ph_a2);
      ph_a2 ^= 1;
      tc_fence_after();
      constexpr int FC = C / EW;                     // 64 output columns per warp
      static_assert(FC == 64, "one staging tile per warp");
      const int gc = slice * FC;
      if (!LN) {
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
          tmem_ld32(t_acc2 + lane_bits + static_cast<uint32_t>(gc + 32 * hf), r);
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
      } else {
        // ---- residual + LayerNorm (transformer_for_adapter.py:882-885): z = x + ffn(x) rounded to 16 bit, y = LN(z).  The warp
        // fetches its 32 x 64 piece of the residual by TMA into its staging tile, forms z in registers, exchanges the partial
        // sums of the ROUNDED values with the three other warps of its lane quarter THROUGH the staging tiles (idle until the
        // z store), stores z, then normalises the packed row it kept and stores y.
        if (lane == 0) {
          tma_store_wait_read();
          mbar_expect_tx(res_full + (warp - 2), 4096);
          tma_load_2d(&tmXr, res_full + (warp - 2), stile, gc, static_cast<int>(row0));
        }
        __syncwarp();
        mbar_wait(res_full + (warp - 2), ph_res);
        ph_res ^= 1;
        uint4 zk[2][4];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t r[32];
          float v[32];
          tmem_ld32(t_acc2 + lane_bits + static_cast<uint32_t>(gc + 32 * hf), r);
          const float4* bp = reinterpret_cast<const float4*>(p.bias2 + gc + 32 * hf);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = __ldg(bp + i);
            v[4 * i] = __uint_as_float(r[4 * i]) + bb.x; v[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + bb.y;
            v[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + bb.z; v[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + bb.w;
          }
          uint4 xin[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) xin[i] = *reinterpret_cast<const uint4*>(swz(stile, lane, 4 * hf + i));
          add_packed16(v, reinterpret_cast<const uint32_t*>(xin), HALF);
          pack_16(v, HALF, false, zk[hf]);
          float zr[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) zr[i] = 0.f;
          add_packed16(zr, reinterpret_cast<const uint32_t*>(zk[hf]), HALF);
#pragma unroll
          for (int i = 0; i < 32; ++i) { s1 += zr[i]; s2 = fmaf(zr[i], zr[i], s2); }
        }
        __syncwarp();                                   // every lane has taken its residual out of the staging tile
        reinterpret_cast<float2*>(stile)[lane] = make_float2(s1, s2);
        named_bar(2 + quarter, 32 * EW);
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int sl = 0; sl < EW; ++sl) {
          const float2 o = reinterpret_cast<const float2*>(sH + (sl * 4 + quarter) * 4096)[lane];
          t1 += o.x; t2 += o.y;
        }
        named_bar(2 + quarter, 32 * EW);               // all four have read: the staging tiles may take z
        const float mean = t1 * (1.f / C);
        const float rstd = rsqrtf(fmaxf(t2 * (1.f / C) - mean * mean, 0.f) + p.ln_eps);
        if (slice == 0 && live) { p.ln_mean[row] = mean; p.ln_rstd[row] = rstd; }
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
#pragma unroll
          for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(swz(stile, lane, 4 * hf + i)) = zk[hf][i];
        fence_async_smem();
        __syncwarp();
        if (lane == 0) { tma_store_2d(&tmZ, stile, gc, static_cast<int>(row0)); tma_store_wait_read(); }
        __syncwarp();
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          float zr[32], v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) zr[i] = 0.f;
          add_packed16(zr, reinterpret_cast<const uint32_t*>(zk[hf]), HALF);
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
          for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(swz(stile, lane, 4 * hf + i)) = pk[i];
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) tma_store_2d(&tmOut, stile, gc, static_cast<int>(row0));
      }
      if (lane == 0) tma_store_wait_read();      // the staging tiles lie in rows other warps write in the next mid stage
      __syncwarp();
      tc_fence_before();
      named_bar(1, 32 * EPI_WARPS);
      if (lane == 0) arrive_leader(a2_free);
      TSEG(5);
    }
#ifdef FFN2_TRACE
    if (p.trace != nullptr && rank == 0 && warp == 2 && lane == 0) {
      te[6] = clock64() - te_begin;
      for (int i = 0; i < 8; ++i) p.trace[592 + (blockIdx.x >> 1) * 8 + i] = te[i];
    }
#endif
    (void)te;
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();      // both CTAs are done with each other's shared and tensor memory
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

long long* g_trace = nullptr;

static int launch(const void* x, const void* wa, const void* wb, void* out, void* z_out, const Params& p_in, cudaStream_t st) {
  Params p = p_in;
  p.trace = g_trace;
  if (!x || !wa || !wb || !out) { snprintf(t_err, sizeof(t_err), "null pointer"); return MSDA_ERR_NULL_POINTER; }
  if (p.R <= 0 || p.F <= 0 || p.F % HC || p.F > 8192) { snprintf(t_err, sizeof(t_err), "chained FFN needs d_ffn %% 128 == 0 (R=%d F=%d)", p.R, p.F); return MSDA_ERR_UNSUPPORTED; }
  const int dt = p.half_in ? 1 : 0;
  CUtensorMap tmX, tmA, tmB, tmOut, tmZ, tmXr;
  int rc = make_map(&tmX, x, p.R, C, BM, 64, dt);
  if (rc) return rc;
  rc = make_map(&tmA, wa, p.F, C, 64, 64, dt);           // half of a chunk's 128 rows per CTA
  if (rc) return rc;
  rc = make_map(&tmB, wb, C, p.F, 128, 64, dt);          // half of the 256 rows per CTA
  if (rc) return rc;
  rc = make_map(&tmOut, out, p.R, C, 32, 64, dt);
  if (rc) return rc;
  rc = make_map(&tmZ, z_out ? z_out : out, p.R, C, 32, 64, dt);
  if (rc) return rc;
  rc = make_map(&tmXr, x, p.R, C, 32, 64, dt);      // the residual, in the final stage's 32 x 64 pieces
  if (rc) return rc;
  int dev_id = 0;
  cudaGetDevice(&dev_id);
  static int sms_of[64] = {};
  if (!sms_of[dev_id & 63]) cudaDeviceGetAttribute(&sms_of[dev_id & 63], cudaDevAttrMultiProcessorCount, dev_id);
  const int tiles = (p.R + BM - 1) / BM, pairs = (tiles + 1) / 2;
  const int max_pairs = sms_of[dev_id & 63] / 2;
  const int grid = 2 * (pairs < max_pairs ? pairs : max_pairs);
  cudaError_t cfg = cudaSuccess;
  ++msda::g_launches;
#define FFN_LAUNCH(BWD, HALF, LN)                                                                                           \
  do {                                                                                                                       \
    static bool configured[64] = {};                                                                                         \
    if (!configured[dev_id & 63]) {                                                                                          \
      cfg = cudaFuncSetAttribute(ffn_chain2_kernel<BWD, HALF, LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);  \
      configured[dev_id & 63] = cfg == cudaSuccess;                                                                          \
    }                                                                                                                        \
    if (cfg == cudaSuccess) ffn_chain2_kernel<BWD, HALF, LN><<<grid, THREADS, SMEM_BYTES, st>>>(tmX, tmA, tmB, tmOut, tmZ, tmXr, p); \
  } while (0)
  if (p.backward) { if (p.half_in) FFN_LAUNCH(true, true, false); else FFN_LAUNCH(true, false, false); }
  else if (z_out) { if (p.half_in) FFN_LAUNCH(false, true, true); else FFN_LAUNCH(false, false, true); }
  else { if (p.half_in) FFN_LAUNCH(false, true, false); else FFN_LAUNCH(false, false, false); }
#undef FFN_LAUNCH
  if (cfg != cudaSuccess) { snprintf(t_err, sizeof(t_err), "cudaFuncSetAttribute: %s", cudaGetErrorString(cfg)); return static_cast<int>(cfg); }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(t_err, sizeof(t_err), "ffn_chain2_kernel launch: %s", cudaGetErrorString(e)); return static_cast<int>(e); }
  return 0;
}

}  // namespace ffn2
}  // namespace pg

extern "C" {

void msda_ffn_chain2_set_trace(long long* buf) { pg::ffn2::g_trace = buf; }

// Same contracts as msda_ffn_chain_fwd_16 / msda_ffn_chain_ln_fwd_16 / msda_ffn_chain_bwd_16, on CTA pairs.
int msda_ffn_chain2_fwd_16(const void* x, const void* w1, const float* b1, const void* w2, const float* b2, long long R, int C,
                           int F, void* out, uint32_t* relu_bits_out, int is_half, void* stream) {
  pg::t_err[0] = 0;
  if (C != pg::ffn2::C) { snprintf(pg::t_err, sizeof(pg::t_err), "chained FFN is built for d_model = %d (got %d)", pg::ffn2::C, C); return MSDA_ERR_UNSUPPORTED; }
  if (!b1) { snprintf(pg::t_err, sizeof(pg::t_err), "null bias"); return MSDA_ERR_NULL_POINTER; }
  if (R >= (1ll << 31)) return MSDA_ERR_BAD_SHAPE;
  pg::ffn2::Params p;
  memset(&p, 0, sizeof(p));
  p.R = static_cast<int>(R); p.F = F; p.backward = 0; p.bias1 = b1; p.bias2 = b2; p.bits = relu_bits_out; p.half_in = is_half;
  return pg::ffn2::launch(x, w1, w2, out, nullptr, p, static_cast<cudaStream_t>(stream));
}

int msda_ffn_chain2_ln_fwd_16(const void* x, const void* w1, const float* b1, const void* w2, const float* b2, long long R, int C,
                              int F, const float* gamma, const float* beta, float eps, void* z, void* y, float* mean, float* rstd,
                              uint32_t* relu_bits_out, int is_half, void* stream) {
  pg::t_err[0] = 0;
  if (C != pg::ffn2::C) { snprintf(pg::t_err, sizeof(pg::t_err), "chained FFN is built for d_model = %d (got %d)", pg::ffn2::C, C); return MSDA_ERR_UNSUPPORTED; }
  if (!b1 || !b2 || !gamma || !beta || !z || !mean || !rstd) { snprintf(pg::t_err, sizeof(pg::t_err), "null pointer"); return MSDA_ERR_NULL_POINTER; }
  if (R >= (1ll << 31)) return MSDA_ERR_BAD_SHAPE;
  pg::ffn2::Params p;
  memset(&p, 0, sizeof(p));
  p.R = static_cast<int>(R); p.F = F; p.backward = 0; p.bias1 = b1; p.bias2 = b2; p.bits = relu_bits_out; p.half_in = is_half;
  p.ln_gamma = gamma; p.ln_beta = beta; p.ln_eps = eps; p.ln_mean = mean; p.ln_rstd = rstd;
  return pg::ffn2::launch(x, w1, w2, y, z, p, static_cast<cudaStream_t>(stream));
}

int msda_ffn_chain2_bwd_16(const void* dy, const void* w2_t, const void* w1_t, const uint32_t* gate_bits, const void* accum,
                           long long R, int C, int F, void* dx, int is_half, void* stream) {
  pg::t_err[0] = 0;
  if (C != pg::ffn2::C) { snprintf(pg::t_err, sizeof(pg::t_err), "chained FFN is built for d_model = %d (got %d)", pg::ffn2::C, C); return MSDA_ERR_UNSUPPORTED; }
  if (!gate_bits) { snprintf(pg::t_err, sizeof(pg::t_err), "null gate bits"); return MSDA_ERR_NULL_POINTER; }
  if (R >= (1ll << 31)) return MSDA_ERR_BAD_SHAPE;
  pg::ffn2::Params p;
  memset(&p, 0, sizeof(p));
  p.R = static_cast<int>(R); p.F = F; p.backward = 1; p.bits = const_cast<uint32_t*>(gate_bits); p.accum = accum; p.half_in = is_half;
  return pg::ffn2::launch(dy, w2_t, w1_t, dx, nullptr, p, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
