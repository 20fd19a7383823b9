// Backward scatter with in-register run merging (sm_100a).
//
// Why: msda_bwd_vec_kernel is bound by the number of 32-byte reduction sectors an SM can hand to the crossbar
// (ncu: l1tex__m_l1tex2xbar_req_cycles_active 85 %, DRAM 6 %; profiles/r1_core_kernels_ncu_v1.txt), and
// shared-memory float atomics are CAS loops on sm_100a, so the only way to go faster is to COMBINE
// contributions before they leave the SM.  Neighbouring queries of the same head sample the same cell of the
// coarser levels (a step of one level-0 pixel is 1/2, 1/4, 1/8 of a pixel on levels 1..3), so here a lane group
// walks a RUN of T consecutive queries of one (image, head) and keeps, per level, one open accumulator per
// sampling point (4 points x 4 corners x 4 channels in registers).  While consecutive queries land in the same
// cell the contributions are added in registers; the 128-bit reductions are issued only when the cell changes
// or the run ends.  No shuffles, no shared memory, no assumption about the offsets: a sample that lands
// elsewhere simply flushes.  grad_sampling_loc / grad_attn_weight are produced exactly as in the plain kernel.
//
// One lane always owns 4 channels (fp32: one 16-byte load; bf16/f16: one 8-byte load), so a group is D/4 lanes
// and every reduction instruction of a group covers D*4 contiguous bytes.
#include <type_traits>

#include "msda_common.cuh"

namespace msda {

template <typename VT> struct Quad;   // 4 channels per lane
template <> struct Quad<float> {
  __device__ static __forceinline__ void load(const float* p, bool ok, float (&f)[4]) {
    const float4 r = ok ? __ldg(reinterpret_cast<const float4*>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
    f[0] = r.x; f[1] = r.y; f[2] = r.z; f[3] = r.w;
  }
};
template <> struct Quad<__nv_bfloat16> {
  __device__ static __forceinline__ void load(const __nv_bfloat16* p, bool ok, float (&f)[4]) {
    const uint2 r = ok ? __ldg(reinterpret_cast<const uint2*>(p)) : make_uint2(0u, 0u);
    f[0] = __uint_as_float(r.x << 16); f[1] = __uint_as_float(r.x & 0xffff0000u);
    f[2] = __uint_as_float(r.y << 16); f[3] = __uint_as_float(r.y & 0xffff0000u);
  }
};
template <> struct Quad<__half> {
  __device__ static __forceinline__ void load(const __half* p, bool ok, float (&f)[4]) {
    const uint2 r = ok ? __ldg(reinterpret_cast<const uint2*>(p)) : make_uint2(0u, 0u);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
  }
};

template <int LPG>
__device__ __forceinline__ void reduce_scatter16_run(float (&v)[16], int lig) {
  int n = 16;
#pragma unroll
  for (int mask = LPG / 2; mask >= 1; mask >>= 1) {
    n >>= 1;
    const bool up = (lig & mask) != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < n) {
        const float send = up ? v[i] : v[i + n];
        const float keep = up ? v[i + n] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
      }
    }
  }
}

// Open accumulator of one sampling point: the cell (pixel offset of its top-left corner, corner guards) and
// the four corner gradients of this lane's 4 channels.
struct OpenCell {
  int o1;        // h0 * W + w0 of the cell; INT_MIN when nothing is open
  int guards;    // bit k set = corner k inside the map
  float a[4][4];
};

template <typename VT, int D, int T>
__global__ void __launch_bounds__(kThreads, 2)
msda_bwd_run_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                    const int64_t* __restrict__ lstart, const float* __restrict__ loc,
                    const float* __restrict__ aw, const VT* __restrict__ grad_out,
                    float* __restrict__ grad_value, float* __restrict__ grad_loc,
                    float* __restrict__ grad_aw, int S, int M, int L, int Lq, int runs_per_image_head,
                    long long total_runs) {
  constexpr int CH = 4;
  constexpr int LPG = D / CH;
  constexpr int UPW = 32 / LPG;
  constexpr int P = 4;
  constexpr int R = 16 / LPG;
  constexpr int NONE = -2147483647 - 1;
  static_assert(LPG >= 2 && LPG <= 16 && (LPG & (LPG - 1)) == 0, "D / 4 must be a power of two in [2, 16]");

  __shared__ int sH[MSDA_MAX_LEVELS], sW[MSDA_MAX_LEVELS], sStart[MSDA_MAX_LEVELS];
  if (threadIdx.x < L) {
    sH[threadIdx.x] = static_cast<int>(shapes[2 * threadIdx.x]);
    sW[threadIdx.x] = static_cast<int>(shapes[2 * threadIdx.x + 1]);
    sStart[threadIdx.x] = static_cast<int>(lstart[threadIdx.x]);
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lig = lane % LPG;
  // consecutive groups / warps take consecutive runs of the SAME (image, head): run index is
  // ((b * M + m) * runs_per_image_head + chunk), chunk-fastest.
  long long run = (static_cast<long long>(blockIdx.x) * kWarpsPerBlock + warp) * UPW + lane / LPG;
  const bool run_ok = run < total_runs;
  if (!run_ok) run = 0;
  const int chunk = static_cast<int>(run % runs_per_image_head);
  const long long bm = run / runs_per_image_head;
  const int m = static_cast<int>(bm % M);
  const long long b = bm / M;
  const int q0 = chunk * T;
  const int row = M * D;
  const size_t voff = (static_cast<size_t>(b) * S * M + m) * D + lig * CH;
  const VT* vb = value + voff;
  float* gvb = grad_value + voff;

  for (int l = 0; l < L; ++l) {
    const int H = sH[l], W = sW[l];
    const size_t loff = static_cast<size_t>(sStart[l]) * row;
    const VT* vl = vb + loff;
    float* gvl = gvb + loff;
    OpenCell cell[P];
#pragma unroll
    for (int p = 0; p < P; ++p) cell[p].o1 = NONE;

    auto flush = [&](OpenCell& c) {
      if (c.o1 != NONE) {
        const long long e1 = static_cast<long long>(c.o1) * row;
        const long long e3 = e1 + static_cast<long long>(W) * row;
        if (c.guards & 1) red_add_v4(gvl + e1, c.a[0][0], c.a[0][1], c.a[0][2], c.a[0][3]);
        if (c.guards & 2) red_add_v4(gvl + e1 + row, c.a[1][0], c.a[1][1], c.a[1][2], c.a[1][3]);
        if (c.guards & 4) red_add_v4(gvl + e3, c.a[2][0], c.a[2][1], c.a[2][2], c.a[2][3]);
        if (c.guards & 8) red_add_v4(gvl + e3 + row, c.a[3][0], c.a[3][1], c.a[3][2], c.a[3][3]);
      }
    };

#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      const int q = q0 + t;
      const bool active = run_ok && q < Lq;
      const size_t u = ((static_cast<size_t>(b) * Lq + (active ? q : 0)) * M + m);
      const float4* lp = reinterpret_cast<const float4*>(loc + u * L * P * 2);
      const float4* ap = reinterpret_cast<const float4*>(aw + u * L * P);
      float g[CH];
      Quad<VT>::load(grad_out + u * D + lig * CH, active, g);
      const float4 xy01 = __ldg(lp + 2 * l), xy23 = __ldg(lp + 2 * l + 1), a4 = __ldg(ap + l);
      const float xs[4] = {xy01.x, xy01.z, xy23.x, xy23.z};
      const float ys[4] = {xy01.y, xy01.w, xy23.y, xy23.w};
      const float as[4] = {a4.x, a4.y, a4.z, a4.w};
      float red[16];
#pragma unroll
      for (int p = 0; p < P; ++p) {
        Tap<float> tp = make_tap<float>(xs[p], ys[p], H, W);
        const bool ok = tp.ok && active;
        tp.c1 = tp.c1 && active; tp.c2 = tp.c2 && active; tp.c3 = tp.c3 && active; tp.c4 = tp.c4 && active;
        float v1[CH], v2[CH], v3[CH], v4[CH];
        const long long e1 = static_cast<long long>(tp.o1) * row;
        const long long e3 = e1 + static_cast<long long>(W) * row;
        Quad<VT>::load(vl + e1, tp.c1, v1);
        Quad<VT>::load(vl + e1 + row, tp.c2, v2);
        Quad<VT>::load(vl + e3, tp.c3, v3);
        Quad<VT>::load(vl + e3 + row, tp.c4, v4);
        const float k1 = tp.hh * tp.hw, k2 = tp.hh * tp.lw, k3 = tp.lh * tp.hw, k4 = tp.lh * tp.lw;
        const float a = as[p];
        float s_w = 0.f, s_h = 0.f, s_a = 0.f, tg[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          tg[c] = g[c] * a;
          const float gh = tp.hw * (v3[c] - v1[c]) + tp.lw * (v4[c] - v2[c]);
          const float gw = tp.hh * (v2[c] - v1[c]) + tp.lh * (v4[c] - v3[c]);
          const float val = k1 * v1[c] + k2 * v2[c] + k3 * v3[c] + k4 * v4[c];
          s_a = fmaf(g[c], val, s_a);
          s_w = fmaf(gw, tg[c], s_w);
          s_h = fmaf(gh, tg[c], s_h);
        }
        red[4 * p + 0] = s_w * static_cast<float>(W);
        red[4 * p + 1] = s_h * static_cast<float>(H);
        red[4 * p + 2] = s_a;
        red[4 * p + 3] = 0.f;

        // ---- run merging: same cell as the open accumulator of this point -> add in registers ----
        const int key = ok ? tp.o1 : NONE;
        OpenCell& c = cell[p];
        if (key != c.o1) {
          flush(c);
          c.o1 = key;
          c.guards = (tp.c1 ? 1 : 0) | (tp.c2 ? 2 : 0) | (tp.c3 ? 4 : 0) | (tp.c4 ? 8 : 0);
#pragma unroll
          for (int cc = 0; cc < CH; ++cc) {
            c.a[0][cc] = k1 * tg[cc]; c.a[1][cc] = k2 * tg[cc]; c.a[2][cc] = k3 * tg[cc]; c.a[3][cc] = k4 * tg[cc];
          }
        } else if (ok) {
#pragma unroll
          for (int cc = 0; cc < CH; ++cc) {
            c.a[0][cc] = fmaf(k1, tg[cc], c.a[0][cc]); c.a[1][cc] = fmaf(k2, tg[cc], c.a[1][cc]);
            c.a[2][cc] = fmaf(k3, tg[cc], c.a[2][cc]); c.a[3][cc] = fmaf(k4, tg[cc], c.a[3][cc]);
          }
        }
      }
      reduce_scatter16_run<LPG>(red, lig);
      if (active) {
        float* glp = grad_loc + u * L * P * 2;
        float* gap = grad_aw + u * L * P;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int idx = lig * R + r, p = idx >> 2, comp = idx & 3;
          if (comp < 2) glp[(l * P + p) * 2 + comp] = red[r];
          else if (comp == 2) gap[l * P + p] = red[r];
        }
      }
    }
#pragma unroll
    for (int p = 0; p < P; ++p) flush(cell[p]);
  }
}

template <typename VT, int D>
static cudaError_t launch_run(const VT* value, const int64_t* shapes, const int64_t* lstart, const float* loc,
                              const float* aw, const VT* grad_out, float* gv, float* gl, float* ga, int N, int S,
                              int M, int L, int Lq, int T, cudaStream_t st) {
  constexpr int UPW = 32 / (D / 4);
  const int rpih = (Lq + T - 1) / T;
  const long long total = static_cast<long long>(N) * M * rpih;
  const long long per_cta = static_cast<long long>(UPW) * kWarpsPerBlock;
  const dim3 grid(static_cast<unsigned>((total + per_cta - 1) / per_cta));
  ++g_launches;
  if (T == 4) msda_bwd_run_kernel<VT, D, 4><<<grid, kThreads, 0, st>>>(value, shapes, lstart, loc, aw, grad_out, gv, gl, ga, S, M, L, Lq, rpih, total);
  else if (T == 16) msda_bwd_run_kernel<VT, D, 16><<<grid, kThreads, 0, st>>>(value, shapes, lstart, loc, aw, grad_out, gv, gl, ga, S, M, L, Lq, rpih, total);
  else msda_bwd_run_kernel<VT, D, 8><<<grid, kThreads, 0, st>>>(value, shapes, lstart, loc, aw, grad_out, gv, gl, ga, S, M, L, Lq, rpih, total);
  return cudaGetLastError();
}

// D in {16, 32, 64}, P == 4.  Returns cudaErrorNotSupported for anything else (caller falls back).
template <typename VT>
cudaError_t backward_run(const VT* value, const int64_t* shapes, const int64_t* lstart, const float* loc,
                         const float* aw, const VT* grad_out, float* gv, float* gl, float* ga, int N, int S, int M,
                         int D, int L, int Lq, int P, int T, cudaStream_t st) {
  if (P != 4) return cudaErrorNotSupported;
  if (D == 32) return launch_run<VT, 32>(value, shapes, lstart, loc, aw, grad_out, gv, gl, ga, N, S, M, L, Lq, T, st);
  if (D == 64) return launch_run<VT, 64>(value, shapes, lstart, loc, aw, grad_out, gv, gl, ga, N, S, M, L, Lq, T, st);
  if (D == 16) return launch_run<VT, 16>(value, shapes, lstart, loc, aw, grad_out, gv, gl, ga, N, S, M, L, Lq, T, st);
  return cudaErrorNotSupported;
}

template cudaError_t backward_run<float>(const float*, const int64_t*, const int64_t*, const float*, const float*, const float*, float*, float*, float*, int, int, int, int, int, int, int, int, cudaStream_t);
template cudaError_t backward_run<__nv_bfloat16>(const __nv_bfloat16*, const int64_t*, const int64_t*, const float*, const float*, const __nv_bfloat16*, float*, float*, float*, int, int, int, int, int, int, int, int, cudaStream_t);
template cudaError_t backward_run<__half>(const __half*, const int64_t*, const int64_t*, const float*, const float*, const __half*, float*, float*, float*, int, int, int, int, int, int, int, int, cudaStream_t);

}  // namespace msda
