// Backward of multi-scale deformable attention, coarse levels: grad_value accumulated in TENSOR MEMORY instead of being
// scattered to L2 with one 128-byte reduction per bilinear corner.
//
// Replaces, for the levels it owns, the grad_value half of ms_deformable_col2im_gpu_kernel_shm_blocksize_aware_reduce_v1
// (ms_deform_im2col_cuda.cuh:301-403, atomicAdd at :115-157): the reference -- and msda_bwd_vec_kernel -- send
// N*Lq*M*P*4 reductions per level whatever the level's size, so a 13x21 level receives ~1300 reductions per pixel and
// head.  Round 1 measured that path at 0.95 of what the SM -> L2 reduction path can carry (profiles/r1_probe_scatter.jsonl);
// only sending fewer rows helps, and shared-memory float atomics are too slow to combine them (B300_MICROARCH.md: 2 clk/lane).
//
// The combine is a matrix product.  For one (image, head) and a chunk of 64 queries,
//
//     grad_value[pixel, 0:32] += sum_q  Wt[pixel, q] * G[q, 0:32],     Wt[pixel, q] = sum over the samples of query q that
//                                                                       touch `pixel` of (attention weight x bilinear weight)
//
// so each chunk is ONE small GEMM: A = Wt (pixels x 64 queries, 16-bit, built in shared memory by 2-byte scatter writes
// -- at most 16 non-zeros per query and level), B = G^T (32 channels x 64 queries, the incoming gradient rows transposed
// on the way into shared memory), D in TMEM (fp32, 128 pixels x 32 channels per block, up to 12 blocks = 1536 pixels).
// The accumulators persist over every chunk a CTA walks for the same (image, head); they leave the SM once, as 128-byte
// fp32 reductions, when the (image, head) changes: ~150 x 2 x 1323 rows per launch instead of 22.7 M at the encoder shape.
//
// The levels owned are the tail [first, L) of the level list whose pixels fit the accumulators (coarse_first_level();
// Swin-T 800x1333: 25x42 + 13x21 = 1323 pixels); msda_bwd_vec_kernel skips exactly those reductions and still produces
// grad_sampling_loc / grad_attn_weight for every level.  16-bit storage only: the products W*G are exact in the fp32
// accumulator but W is rounded to the storage type (2^-9 relative for bf16), inside the 1e-2 bar of north_star and
// outside the fp32 path's 1e-4 -- fp32 storage keeps the reduction path.
#include <type_traits>

#include "msda_common.cuh"
#include "tc_common.cuh"

namespace msda {

using namespace pg;

namespace {
constexpr int KQ = 64;                          // queries per chunk = K extent of an operand tile row (128 bytes)
constexpr int A_BYTES = kMmaBlocks * 128 * 128;   // 128 pixel rows x 128 bytes per block
constexpr int B_BYTES = 32 * 128;               // 32 channel rows x 128 bytes
constexpr int BUILD_WARPS = 4;
constexpr int THREADS = (BUILD_WARPS + 1) * 32;
constexpr int SMEM_BYTES = 1024 + A_BYTES + B_BYTES + 64;

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

// byte offset of element (row, k) inside a K-major, 128-byte-swizzled operand tile (rows of 64 16-bit elements, 8-row
// groups 1024 bytes apart -- the layout umma_desc_sw128 describes and TMA's SWIZZLE_128B produces)
__device__ __forceinline__ uint32_t tile_off(uint32_t row, uint32_t k) {
  return row * 128u + ((((k >> 3) ^ row) & 7u) << 4) + (k & 7u) * 2u;
}
}  // namespace

template <typename VT>
__global__ void __launch_bounds__(THREADS, 1)
msda_scatter_mma_kernel(const int64_t* __restrict__ shapes, const int64_t* __restrict__ lstart,
                        const float* __restrict__ loc, const float* __restrict__ aw,
                        const VT* __restrict__ grad_out, float* __restrict__ grad_value, int N, int S, int M, int L, int Lq) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + A_BYTES;
  uint64_t* bar_built = reinterpret_cast<uint64_t*>(sB + B_BYTES);
  uint64_t* bar_done = bar_built + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_done + 1);
  __shared__ int sH[MSDA_MAX_LEVELS], sW[MSDA_MAX_LEVELS], sStart[MSDA_MAX_LEVELS];
  __shared__ int sFirst;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < L) {
    sH[threadIdx.x] = static_cast<int>(shapes[2 * threadIdx.x]);
    sW[threadIdx.x] = static_cast<int>(shapes[2 * threadIdx.x + 1]);
    sStart[threadIdx.x] = static_cast<int>(lstart[threadIdx.x]);
  }
  __syncthreads();
  if (threadIdx.x == 0) sFirst = coarse_first_level(sH, sW, sStart, L, S);
  __syncthreads();
  const int first = sFirst;
  if (first >= L) return;   // no level fits the accumulators: msda_bwd_vec_kernel sent every reduction itself

  const int px0 = sStart[first];
  const int npx = S - px0;
  const int nblk = (npx + 127) >> 7;
  const int cpq = (Lq + KQ - 1) / KQ;
  const long long total = static_cast<long long>(N) * M * cpq;
  const long long c_begin = total * blockIdx.x / gridDim.x, c_end = total * (blockIdx.x + 1) / gridDim.x;

  // operand tiles start out zero; the builders restore the entries they touched after every product
  for (int i = threadIdx.x; i < (A_BYTES + B_BYTES) / 16; i += THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (threadIdx.x == 0) {
    mbar_init(bar_built, BUILD_WARPS * 32);
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == BUILD_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == BUILD_WARPS) {
    // ===== MMA issuer: per chunk, nblk x 4 products (M = 128 pixels, N = 32 channels, K = 16 queries) =====
    const uint32_t idesc = umma_idesc(128, 32, std::is_same<VT, __half>::value);
    uint32_t phase = 0;
    long long cur_bm = -1;
    for (long long c = c_begin; c < c_end; ++c) {
      const long long bm = c / cpq;
      const bool fresh = bm != cur_bm;       // first chunk of an (image, head): overwrite the accumulators
      cur_bm = bm;
      mbar_wait(bar_built, phase);
      phase ^= 1;
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
    }
  } else {
    // ===== builders: thread = (query of the chunk, level slot); 2-byte read-modify-writes into the A tile =====
    const int q_in = threadIdx.x & (KQ - 1), slot = threadIdx.x >> 6;
    const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
    const uint32_t kswz = static_cast<uint32_t>(q_in >> 3), klo = static_cast<uint32_t>(q_in & 7) * 2u;
    auto a_addr = [&](int pix) {
      const uint32_t r = static_cast<uint32_t>(pix) & 127u;
      return a_base + (static_cast<uint32_t>(pix) >> 7) * 16384u + r * 128u + (((kswz ^ r) & 7u) << 4) + klo;
    };
    // pixels written for the previous chunk (two levels per thread at most), packed two per word; 0xffff = none
    uint32_t saved[2][8];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) saved[i][j] = 0xffffffffu;
    uint32_t phase = 0;
    long long cur_bm = -1;

    auto flush = [&](long long bm) {
      // accumulators -> grad_value: thread = one pixel row of 32 channels (TMEM lane = warp * 32 + lane), 8 x 128-bit reductions
      const long long b = bm / M;
      const int m = static_cast<int>(bm % M);
      for (int blk = 0; blk < nblk; ++blk) {
        uint32_t r[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(blk * 32), r);
        const int pix = blk * 128 + warp * 32 + lane;
        uint32_t nz = 0u;
#pragma unroll
        for (int j = 0; j < 32; ++j) nz |= r[j] & 0x7fffffffu;
        if (pix < npx && nz != 0u) {
          float* dst = grad_value + ((static_cast<size_t>(b) * S + px0 + pix) * M + m) * 32;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            red_add_v4(dst + 4 * j, __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                       __uint_as_float(r[4 * j + 3]));
        }
      }
      tc_fence_before();
    };

    for (long long c = c_begin; c < c_end; ++c) {
      const long long bm = c / cpq;
      const int qc = static_cast<int>(c % cpq);
      const long long b = bm / M;
      const int m = static_cast<int>(bm % M);
      const int q = qc * KQ + q_in;
      const bool valid = q < Lq;
      const size_t u = (static_cast<size_t>(b) * Lq + (valid ? q : 0)) * M + m;
      // global loads of this chunk go out before the wait on the previous product
      float4 xy01[2], xy23[2], a4[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int l = first + slot + 2 * i;
        xy01[i] = xy23[i] = a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (l < L && valid) {
          const float4* lp = reinterpret_cast<const float4*>(loc + u * L * 8 + l * 8);
          xy01[i] = __ldg(lp); xy23[i] = __ldg(lp + 1);
          a4[i] = __ldg(reinterpret_cast<const float4*>(aw + u * L * 4 + l * 4));
        }
      }
      uint4 g0 = make_uint4(0u, 0u, 0u, 0u), g1 = g0;
      if (valid) {
        const uint4* gp = reinterpret_cast<const uint4*>(grad_out + u * 32 + 16 * slot);
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
        const int l = first + slot + 2 * i;
        if (l < L && valid) {
          const int H = sH[l], W = sW[l], base = sStart[l] - px0;
          const float xs[4] = {xy01[i].x, xy01[i].z, xy23[i].x, xy23[i].z};
          const float ys[4] = {xy01[i].y, xy01[i].w, xy23[i].y, xy23[i].w};
          const float as[4] = {a4[i].x, a4[i].y, a4[i].z, a4[i].w};
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const Tap<float> t = make_tap<float>(xs[p], ys[p], H, W);
            const float a = as[p];
            const int px[4] = {base + t.o1, base + t.o2, base + t.o3, base + t.o4};
            const bool cg[4] = {t.c1, t.c2, t.c3, t.c4};
            nwgt[i][4 * p] = t.hh * t.hw * a; nwgt[i][4 * p + 1] = t.hh * t.lw * a;
            nwgt[i][4 * p + 2] = t.lh * t.hw * a; nwgt[i][4 * p + 3] = t.lh * t.lw * a;
            npix[i][2 * p] = (cg[0] ? static_cast<uint32_t>(px[0]) : 0xffffu) | ((cg[1] ? static_cast<uint32_t>(px[1]) : 0xffffu) << 16);
            npix[i][2 * p + 1] = (cg[2] ? static_cast<uint32_t>(px[2]) : 0xffffu) | ((cg[3] ? static_cast<uint32_t>(px[3]) : 0xffffu) << 16);
          }
        }
      }
      if (c > c_begin) {
        mbar_wait(bar_done, phase);      // the previous chunk's products have read the tiles and landed in TMEM
        phase ^= 1;
        tc_fence_after();
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t w = saved[i][j];
            if ((w & 0xffffu) != 0xffffu) sts_u16(a_addr(static_cast<int>(w & 0xffffu)), 0u);
            if ((w >> 16) != 0xffffu) sts_u16(a_addr(static_cast<int>(w >> 16)), 0u);
            saved[i][j] = 0xffffffffu;
          }
        if (bm != cur_bm) flush(cur_bm);
      }
      cur_bm = bm;
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
      // ---- B: 16 channels of this query's incoming gradient row, transposed (row = channel, k = query) ----
      {
        const uint32_t w[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t d0 = static_cast<uint32_t>(16 * slot + 2 * j);
          sts_u16(b_base + tile_off(d0, static_cast<uint32_t>(q_in)), w[j] & 0xffffu);
          sts_u16(b_base + tile_off(d0 + 1, static_cast<uint32_t>(q_in)), w[j] >> 16);
        }
      }
      fence_async_smem();          // generic-proxy writes -> visible to the tensor core's async-proxy reads
      mbar_arrive(bar_built);
    }
    if (c_end > c_begin) {
      mbar_wait(bar_done, phase);
      tc_fence_after();
      flush(cur_bm);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == BUILD_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <typename VT>
cudaError_t launch_scatter_mma(const int64_t* shapes, const int64_t* lstart, const float* loc, const float* aw,
                               const VT* grad_out, float* gv, int N, int S, int M, int L, int Lq, cudaStream_t st) {
  int dev = 0;
  cudaGetDevice(&dev);
  static int sms_of[64] = {};
  static bool configured[64] = {};
  if (!sms_of[dev & 63]) cudaDeviceGetAttribute(&sms_of[dev & 63], cudaDevAttrMultiProcessorCount, dev);
  if (!configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(msda_scatter_mma_kernel<VT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured[dev & 63] = true;
  }
  const long long total = static_cast<long long>(N) * M * ((Lq + KQ - 1) / KQ);
  const int grid = static_cast<int>(total < sms_of[dev & 63] ? total : sms_of[dev & 63]);
  ++g_launches;
  msda_scatter_mma_kernel<VT><<<grid, THREADS, SMEM_BYTES, st>>>(shapes, lstart, loc, aw, grad_out, gv, N, S, M, L, Lq);
  return cudaGetLastError();
}

template cudaError_t launch_scatter_mma<__nv_bfloat16>(const int64_t*, const int64_t*, const float*, const float*, const __nv_bfloat16*, float*, int, int, int, int, int, cudaStream_t);
template cudaError_t launch_scatter_mma<__half>(const int64_t*, const int64_t*, const float*, const float*, const __half*, float*, int, int, int, int, int, cudaStream_t);

}  // namespace msda
