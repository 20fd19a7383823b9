// Tensor-memory scatter, second generation: fine levels too, by RANGES of the value map and per-chunk hit masks.
//
// msda_scatter_mma.cu owns only the coarse tail of the level list because one CTA's accumulators hold 1536 pixels.  Here
// the owned levels are cut into ranges of <= 768 pixels (plan_ranges(), msda_common.cuh) and a work unit is
// (image, head, range, segment of 16 query chunks).  A unit's CTA accumulates, in tensor memory, the contributions of
// those chunks to ITS range only:  grad_value[range pixels, 0:32] += Wt[pixels, 64 queries] * G[64 queries, 0:32]  per
// chunk (same operand construction as the first generation), and flushes the range once per unit.  What makes this cheap
// is that deformable attention is local -- a chunk of 64 consecutive queries samples a few pixel rows of each level --
// so most (chunk, range) pairs are empty: msda_bwd_vec_kernel, which computes every sampling tap anyway, ORs one bit per
// touched range into hit[(image, head), chunk] (64-bit words in the caller's workspace), and a unit skips the chunks
// whose bit for its range is clear without recomputing anything.  With uniformly random locations every pair hits and
// the kernel degrades to ~the reduction path's cost; with trained-model statistics the reductions that leave the SM drop
// from 64 per (query, head) to the flushes: one 128-byte row per touched range pixel per unit.
//
// Two CTAs per SM (101 KiB of shared memory, 256 TMEM columns each): while one CTA's tensor-core products run, the
// other builds its operand tiles.  16-bit storage, D = 32, P = 4 (see msda_scatter_mma.cu for the precision argument).
#include <type_traits>

#include "msda_common.cuh"
#include "tc_common.cuh"

namespace msda {

using namespace pg;

namespace {
constexpr int KQ = 64;
constexpr int A_BYTES = kR2Blocks * 128 * 128;
constexpr int B_BYTES = 32 * 128;
constexpr int BUILD_WARPS = 4;
constexpr int THREADS = (BUILD_WARPS + 1) * 32;
constexpr int SMEM_BYTES = 1024 + A_BYTES + B_BYTES + 64;
constexpr uint32_t TMEM_COLS = 256;

__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(static_cast<uint16_t>(v)) : "memory");
}
template <typename VT> __device__ __forceinline__ float w_to_f32(uint32_t h);
template <> __device__ __forceinline__ float w_to_f32<__nv_bfloat16>(uint32_t h) { return __uint_as_float(h << 16); }
template <> __device__ __forceinline__ float w_to_f32<__half>(uint32_t h) { return __half2float(__ushort_as_half(static_cast<unsigned short>(h))); }
template <typename VT> __device__ __forceinline__ uint32_t f32_to_w(float f);
template <> __device__ __forceinline__ uint32_t f32_to_w<__nv_bfloat16>(float f) { return __bfloat16_as_ushort(__float2bfloat16_rn(f)); }
template <> __device__ __forceinline__ uint32_t f32_to_w<__half>(float f) { return __half_as_ushort(__float2half_rn(f)); }
__device__ __forceinline__ uint32_t tile_off(uint32_t row, uint32_t k) {
  return row * 128u + ((((k >> 3) ^ row) & 7u) << 4) + (k & 7u) * 2u;
}
}  // namespace

template <typename VT>
__global__ void __launch_bounds__(THREADS, 2)
msda_scatter_mma2_kernel(const int64_t* __restrict__ shapes, const int64_t* __restrict__ lstart,
                         const float* __restrict__ loc, const float* __restrict__ aw, const VT* __restrict__ grad_out,
                         float* __restrict__ grad_value, const unsigned long long* __restrict__ hit, int N, int S, int M,
                         int L, int Lq, int max_levels) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + A_BYTES;
  uint64_t* bar_built = reinterpret_cast<uint64_t*>(sB + B_BYTES);
  uint64_t* bar_done = bar_built + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_done + 1);
  __shared__ int sH[MSDA_MAX_LEVELS], sW[MSDA_MAX_LEVELS], sStart[MSDA_MAX_LEVELS];
  __shared__ RangePlan plan;
  __shared__ uint32_t sUnitHits;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < L) {
    sH[threadIdx.x] = static_cast<int>(shapes[2 * threadIdx.x]);
    sW[threadIdx.x] = static_cast<int>(shapes[2 * threadIdx.x + 1]);
    sStart[threadIdx.x] = static_cast<int>(lstart[threadIdx.x]);
  }
  __syncthreads();
  if (threadIdx.x == 0) plan_ranges(plan, sH, sW, sStart, L, S, max_levels);
  __syncthreads();
  const int R = plan.nranges;
  if (R == 0) return;

  const int cpq = (Lq + KQ - 1) / KQ;
  const int segs = (cpq + kR2SegChunks - 1) / kR2SegChunks;
  const long long units = static_cast<long long>(N) * M * R * segs;

  for (int i = threadIdx.x; i < (A_BYTES + B_BYTES) / 16; i += THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (threadIdx.x == 0) {
    mbar_init(bar_built, BUILD_WARPS * 32);
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == BUILD_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // builder state (unused by the MMA warp)
  const int q_in = threadIdx.x & (KQ - 1), slot = (threadIdx.x >> 6) & 1;
  const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
  const uint32_t kswz = static_cast<uint32_t>(q_in >> 3), klo = static_cast<uint32_t>(q_in & 7) * 2u;
  auto a_addr = [&](int pix) {
    const uint32_t r = static_cast<uint32_t>(pix) & 127u;
    return a_base + (static_cast<uint32_t>(pix) >> 7) * 16384u + r * 128u + (((kswz ^ r) & 7u) << 4) + klo;
  };
  uint32_t saved[2][8];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) saved[i][j] = 0xffffffffu;
  uint32_t ph_built = 0, ph_done = 0;     // mbarrier phases: one completion of each per accumulated chunk
  const uint32_t idesc = umma_idesc(128, 32, std::is_same<VT, __half>::value);

  for (long long u = blockIdx.x; u < units; u += gridDim.x) {
    const int seg = static_cast<int>(u % segs);
    const int r = static_cast<int>((u / segs) % R);
    const long long bm = u / (static_cast<long long>(segs) * R);
    const int c0 = seg * kR2SegChunks;
    // which of this unit's chunks touch range r (bits written by msda_bwd_vec_kernel)
    if (warp == 0) {
      const int c = c0 + lane;
      const bool h = lane < kR2SegChunks && c < cpq && ((hit[bm * cpq + c] >> r) & 1ull) != 0ull;
      const uint32_t m = __ballot_sync(0xffffffffu, h);
      if (lane == 0) sUnitHits = m;
    }
    __syncthreads();
    const uint32_t hits = sUnitHits;
    __syncthreads();                       // everyone has read it before the next unit overwrites it
    if (hits == 0u) continue;

    const int lo = plan.lo[r], npx = plan.hi[r] - lo, nblk = (npx + 127) >> 7;
    const int lv0 = plan.lv0[r], lv1 = plan.lv1[r];
    const long long b = bm / M;
    const int m = static_cast<int>(bm % M);

    if (warp == BUILD_WARPS) {
      // ===== MMA issuer =====
      bool fresh = true;
      for (uint32_t rest = hits; rest != 0u; rest &= rest - 1u) {
        mbar_wait(bar_built, ph_built);
        ph_built ^= 1;
        tc_fence_after();
        if (elect_one()) {
          for (int blk = 0; blk < nblk; ++blk) {
#pragma unroll
            for (int k = 0; k < KQ / 16; ++k) {
              const uint64_t da = umma_desc_sw128(sA + blk * 16384, k * 32);
              const uint64_t db = umma_desc_sw128(sB, k * 32);
              umma_f16(tmem_base + static_cast<uint32_t>(blk * 32), da, db, idesc, (!fresh || k != 0) ? 1u : 0u);
            }
          }
          umma_commit(bar_done);
        }
        __syncwarp();
        fresh = false;
      }
    } else {
      // ===== builders =====
      bool pending = false;
      auto restore = [&]() {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t w = saved[i][j];
            if ((w & 0xffffu) != 0xffffu) sts_u16(a_addr(static_cast<int>(w & 0xffffu)), 0u);
            if ((w >> 16) != 0xffffu) sts_u16(a_addr(static_cast<int>(w >> 16)), 0u);
            saved[i][j] = 0xffffffffu;
          }
      };
      for (uint32_t rest = hits; rest != 0u; rest &= rest - 1u) {
        const int c = c0 + __ffs(rest) - 1;
        const int q = c * KQ + q_in;
        const bool valid = q < Lq;
        const size_t un = (static_cast<size_t>(b) * Lq + (valid ? q : 0)) * M + m;
        float4 xy01[2], xy23[2], a4[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int l = lv0 + slot + 2 * i;
          xy01[i] = xy23[i] = a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (l < lv1 && valid) {
            const float4* lp = reinterpret_cast<const float4*>(loc + un * L * 8 + l * 8);
            xy01[i] = __ldg(lp); xy23[i] = __ldg(lp + 1);
            a4[i] = __ldg(reinterpret_cast<const float4*>(aw + un * L * 4 + l * 4));
          }
        }
        uint4 g0 = make_uint4(0u, 0u, 0u, 0u), g1 = g0;
        if (valid) {
          const uint4* gp = reinterpret_cast<const uint4*>(grad_out + un * 32 + 16 * slot);
          g0 = __ldg(gp); g1 = __ldg(gp + 1);
        }
        // ---- taps of this chunk -> registers (pixel ids + weights), BEFORE waiting for the previous product: the coordinate
        // math overlaps the tensor core; only the shared-memory traffic below sits on the chunk-to-chunk critical path ----
        uint32_t npix[2][8];
        float nwgt[2][16];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
#pragma unroll
          for (int j = 0; j < 8; ++j) npix[i][j] = 0xffffffffu;
          const int l = lv0 + slot + 2 * i;
          if (l < lv1 && valid) {
            const int H = sH[l], W = sW[l], base = sStart[l] - lo;
            const float xs[4] = {xy01[i].x, xy01[i].z, xy23[i].x, xy23[i].z};
            const float ys[4] = {xy01[i].y, xy01[i].w, xy23[i].y, xy23[i].w};
            const float as[4] = {a4[i].x, a4[i].y, a4[i].z, a4[i].w};
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              const Tap<float> t = make_tap<float>(xs[p], ys[p], H, W);
              const float a = as[p];
              const int px[4] = {base + t.o1, base + t.o2, base + t.o3, base + t.o4};
              const bool cg[4] = {t.c1 && px[0] >= 0 && px[0] < npx, t.c2 && px[1] >= 0 && px[1] < npx,
                                      t.c3 && px[2] >= 0 && px[2] < npx, t.c4 && px[3] >= 0 && px[3] < npx};
              nwgt[i][4 * p] = t.hh * t.hw * a; nwgt[i][4 * p + 1] = t.hh * t.lw * a;
              nwgt[i][4 * p + 2] = t.lh * t.hw * a; nwgt[i][4 * p + 3] = t.lh * t.lw * a;
              npix[i][2 * p] = (cg[0] ? static_cast<uint32_t>(px[0]) : 0xffffu) | ((cg[1] ? static_cast<uint32_t>(px[1]) : 0xffffu) << 16);
              npix[i][2 * p + 1] = (cg[2] ? static_cast<uint32_t>(px[2]) : 0xffffu) | ((cg[3] ? static_cast<uint32_t>(px[3]) : 0xffffu) << 16);
            }
          }
        }
        if (pending) {
          mbar_wait(bar_done, ph_done);
          ph_done ^= 1;
          tc_fence_after();
          restore();
        }
        // ---- A: read-modify-write the new entries (a point's four corners are distinct pixels; points are serialised because
        // two points of one query may share a pixel) ----
#pragma unroll
        for (int i = 0; i < 2; ++i) {
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const uint32_t w0 = npix[i][2 * p], w1 = npix[i][2 * p + 1];
            const uint32_t pid[4] = {w0 & 0xffffu, w0 >> 16, w1 & 0xffffu, w1 >> 16};
            uint32_t old[4], addr[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              addr[k] = a_addr(pid[k] != 0xffffu ? static_cast<int>(pid[k]) : 0);
              old[k] = pid[k] != 0xffffu ? lds_u16(addr[k]) : 0u;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (pid[k] != 0xffffu) sts_u16(addr[k], f32_to_w<VT>(w_to_f32<VT>(old[k]) + nwgt[i][4 * p + k]));
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) saved[i][j] = npix[i][j];
        }
        {
          const uint32_t w[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t d0 = static_cast<uint32_t>(16 * slot + 2 * j);
            sts_u16(b_base + tile_off(d0, static_cast<uint32_t>(q_in)), w[j] & 0xffffu);
            sts_u16(b_base + tile_off(d0 + 1, static_cast<uint32_t>(q_in)), w[j] >> 16);
          }
        }
        fence_async_smem();
        mbar_arrive(bar_built);
        pending = true;
      }
      // unit end: last product done -> clean the tile, flush the range's accumulators
      mbar_wait(bar_done, ph_done);
      ph_done ^= 1;
      tc_fence_after();
      restore();
      for (int blk = 0; blk < nblk; ++blk) {
        uint32_t rr[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(blk * 32), rr);
        const int pix = blk * 128 + warp * 32 + lane;
        uint32_t nz = 0u;
#pragma unroll
        for (int j = 0; j < 32; ++j) nz |= rr[j] & 0x7fffffffu;
        if (pix < npx && nz != 0u) {
          float* dst = grad_value + ((static_cast<size_t>(b) * S + lo + pix) * M + m) * 32;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            red_add_v4(dst + 4 * j, __uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]), __uint_as_float(rr[4 * j + 2]),
                       __uint_as_float(rr[4 * j + 3]));
        }
      }
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == BUILD_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <typename VT>
cudaError_t launch_scatter_mma2(const int64_t* shapes, const int64_t* lstart, const float* loc, const float* aw,
                                const VT* grad_out, float* gv, const unsigned long long* hit, int N, int S, int M, int L,
                                int Lq, int max_levels, cudaStream_t st) {
  int dev = 0;
  cudaGetDevice(&dev);
  static int sms_of[64] = {};
  static bool configured[64] = {};
  if (!sms_of[dev & 63]) cudaDeviceGetAttribute(&sms_of[dev & 63], cudaDevAttrMultiProcessorCount, dev);
  if (!configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(msda_scatter_mma2_kernel<VT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured[dev & 63] = true;
  }
  ++g_launches;
  msda_scatter_mma2_kernel<VT><<<2 * sms_of[dev & 63], THREADS, SMEM_BYTES, st>>>(shapes, lstart, loc, aw, grad_out, gv, hit, N, S,
                                                                                    M, L, Lq, max_levels);
  return cudaGetLastError();
}

template cudaError_t launch_scatter_mma2<__nv_bfloat16>(const int64_t*, const int64_t*, const float*, const float*, const __nv_bfloat16*, float*, const unsigned long long*, int, int, int, int, int, int, cudaStream_t);
template cudaError_t launch_scatter_mma2<__half>(const int64_t*, const int64_t*, const float*, const float*, const __half*, float*, const unsigned long long*, int, int, int, int, int, int, cudaStream_t);

}  // namespace msda
