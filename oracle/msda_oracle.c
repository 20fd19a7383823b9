/*
 * TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's multi-scale deformable attention
 * arithmetic.  Nothing under ziragroundingdino_b200/ may link, import or call this file; it is the
 * checker used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 *
 * What it restates (paths relative to /root/reference/groundingdino/models/GroundingDINO/):
 *   forward   csrc/MsDeformAttn/ms_deform_im2col_cuda.cuh:237-299 (per-output loop over levels and
 *             points, h_im = loc_h*H - 0.5, sample guard at :288) and :33-84 (bilinear taps with the
 *             four corner guards at :56-78);
 *   backward  ms_deform_im2col_cuda.cuh:301-403 (per (b,q,m) loop, per-sample reduction over the
 *             channels of grad_sampling_loc / grad_attn_weight) and :87-159 (corner scatter into
 *             grad_value, gradient of the bilinear weights, scaling by width / height at :157-158).
 *
 * Pinning: the reference ships no golden vectors for this path, so this
 * oracle is pinned against fixtures generated here by importing the reference's own Python
 * implementation (ms_deform_attn.py:90-130, multi_scale_deformable_attn_pytorch, plus autograd)
 * -- see tests/golden/make_golden.py and tests/test_oracle_golden.py.
 *
 * Layouts (all contiguous, row-major), as the reference op receives them (ms_deform_attn_cuda.cu:41-49):
 *   value  [N][S][M][D]      loc [N][Lq][M][L][P][2] (x then y, normalised to [0,1])
 *   aw     [N][Lq][M][L][P]  out [N][Lq][M*D]        shapes [L][2] = (H, W) int64, start [L] int64
 *
 * Coordinate arithmetic: the source line `loc * size - 0.5` (im2col.cuh:285-286, :363-364) is compiled by nvcc into a
 * single fused multiply-add (FFMA / DFMA with a -0.5 immediate in the SASS of the reference op built under oracle/_ref),
 * so the restatement uses fmaf()/fma() -- one rounding.  Everything else is compiled with -ffp-contract=off.
 *
 * The CUDA kernels accumulate grad_value with atomics in an undefined order; here accumulation is
 * sequential in (q, l, p, corner) order inside each (b, m) slice, and slices are independent, so
 * the OpenMP version is deterministic.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define FMA_f32(a, b, c) fmaf((a), (b), (c))
#define FMA_f64(a, b, c) fma((a), (b), (c))

#define ORACLE_DEFINE(T, SUFFIX)                                                                   \
  void msda_oracle_forward_##SUFFIX(const T *value, const int64_t *shapes, const int64_t *start,   \
                                    const T *loc, const T *aw, int N, int S, int M, int D, int L,  \
                                    int Lq, int P, T *out) {                                       \
    const int64_t row = (int64_t)M * D;                                                            \
    _Pragma("omp parallel for collapse(2) schedule(static)")                                       \
    for (int b = 0; b < N; ++b) {                                                                  \
      for (int q = 0; q < Lq; ++q) {                                                               \
        for (int m = 0; m < M; ++m) {                                                              \
          const int64_t qm = ((int64_t)b * Lq + q) * M + m;                                        \
          const T *lp = loc + qm * L * P * 2;                                                      \
          const T *ap = aw + qm * L * P;                                                           \
          T *op = out + qm * D;                                                                    \
          for (int c = 0; c < D; ++c) op[c] = (T)0;                                                \
          for (int l = 0; l < L; ++l) {                                                            \
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];                          \
            const T *vbase = value + ((int64_t)b * S + start[l]) * row + (int64_t)m * D;           \
            for (int p = 0; p < P; ++p) {                                                          \
              const T lw_ = lp[(l * P + p) * 2], lh_ = lp[(l * P + p) * 2 + 1];                    \
              const T a = ap[l * P + p];                                                           \
              /* im2col.cuh:285-286 as nvcc compiles it: one fused multiply-add (see header) */    \
              const T h = FMA_##SUFFIX(lh_, (T)H, (T)-0.5), w = FMA_##SUFFIX(lw_, (T)W, (T)-0.5);  \
              if (!(h > -1 && w > -1 && h < H && w < W)) continue;                                 \
              const int h0 = (int)floor((double)h), w0 = (int)floor((double)w);                    \
              const int h1 = h0 + 1, w1 = w0 + 1;                                                  \
              const T lh = h - h0, lw = w - w0, hh = 1 - lh, hw = 1 - lw;                          \
              const T k1 = hh * hw, k2 = hh * lw, k3 = lh * hw, k4 = lh * lw;                      \
              const T *p1 = (h0 >= 0 && w0 >= 0) ? vbase + ((int64_t)h0 * W + w0) * row : 0;       \
              const T *p2 = (h0 >= 0 && w1 <= W - 1) ? vbase + ((int64_t)h0 * W + w1) * row : 0;   \
              const T *p3 = (h1 <= H - 1 && w0 >= 0) ? vbase + ((int64_t)h1 * W + w0) * row : 0;   \
              const T *p4 = (h1 <= H - 1 && w1 <= W - 1) ? vbase + ((int64_t)h1 * W + w1) * row : 0; \
              for (int c = 0; c < D; ++c) {                                                        \
                const T v1 = p1 ? p1[c] : (T)0, v2 = p2 ? p2[c] : (T)0;                            \
                const T v3 = p3 ? p3[c] : (T)0, v4 = p4 ? p4[c] : (T)0;                            \
                op[c] += (k1 * v1 + k2 * v2 + k3 * v3 + k4 * v4) * a;                              \
              }                                                                                    \
            }                                                                                      \
          }                                                                                        \
        }                                                                                          \
      }                                                                                            \
    }                                                                                              \
  }                                                                                                \
                                                                                                   \
  void msda_oracle_backward_##SUFFIX(const T *value, const int64_t *shapes, const int64_t *start,  \
                                     const T *loc, const T *aw, const T *grad_out, int N, int S,   \
                                     int M, int D, int L, int Lq, int P, T *grad_value,            \
                                     T *grad_loc, T *grad_aw) {                                    \
    const int64_t row = (int64_t)M * D;                                                            \
    memset(grad_value, 0, sizeof(T) * (size_t)N * S * M * D);                                      \
    memset(grad_loc, 0, sizeof(T) * (size_t)N * Lq * M * L * P * 2);                               \
    memset(grad_aw, 0, sizeof(T) * (size_t)N * Lq * M * L * P);                                    \
    /* (b, m) slices touch disjoint parts of every output: race-free and deterministic */         \
    _Pragma("omp parallel for collapse(2) schedule(dynamic, 1)")                                   \
    for (int b = 0; b < N; ++b) {                                                                  \
      for (int m = 0; m < M; ++m) {                                                                \
        for (int q = 0; q < Lq; ++q) {                                                             \
          const int64_t qm = ((int64_t)b * Lq + q) * M + m;                                        \
          const T *lp = loc + qm * L * P * 2;                                                      \
          const T *ap = aw + qm * L * P;                                                           \
          const T *gp = grad_out + qm * D;                                                         \
          T *glp = grad_loc + qm * L * P * 2;                                                      \
          T *gap = grad_aw + qm * L * P;                                                           \
          for (int l = 0; l < L; ++l) {                                                            \
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];                          \
            const int64_t off = ((int64_t)b * S + start[l]) * row + (int64_t)m * D;                \
            const T *vbase = value + off;                                                          \
            T *gvbase = grad_value + off;                                                          \
            for (int p = 0; p < P; ++p) {                                                          \
              const T lw_ = lp[(l * P + p) * 2], lh_ = lp[(l * P + p) * 2 + 1];                    \
              const T a = ap[l * P + p];                                                           \
              const T h = FMA_##SUFFIX(lh_, (T)H, (T)-0.5), w = FMA_##SUFFIX(lw_, (T)W, (T)-0.5);  \
              if (!(h > -1 && w > -1 && h < H && w < W)) continue;                                 \
              const int h0 = (int)floor((double)h), w0 = (int)floor((double)w);                    \
              const int h1 = h0 + 1, w1 = w0 + 1;                                                  \
              const T lh = h - h0, lw = w - w0, hh = 1 - lh, hw = 1 - lw;                          \
              const T k1 = hh * hw, k2 = hh * lw, k3 = lh * hw, k4 = lh * lw;                      \
              const int64_t o1 = (h0 >= 0 && w0 >= 0) ? ((int64_t)h0 * W + w0) * row : -1;         \
              const int64_t o2 = (h0 >= 0 && w1 <= W - 1) ? ((int64_t)h0 * W + w1) * row : -1;     \
              const int64_t o3 = (h1 <= H - 1 && w0 >= 0) ? ((int64_t)h1 * W + w0) * row : -1;     \
              const int64_t o4 = (h1 <= H - 1 && w1 <= W - 1) ? ((int64_t)h1 * W + w1) * row : -1; \
              T s_w = 0, s_h = 0, s_a = 0;                                                         \
              for (int c = 0; c < D; ++c) {                                                        \
                const T g = gp[c], ga = g * a; /* top_grad_value, im2col.cuh:115 */                \
                T gh = 0, gw = 0, v1 = 0, v2 = 0, v3 = 0, v4 = 0;                                  \
                if (o1 >= 0) { v1 = vbase[o1 + c]; gh -= hw * v1; gw -= hh * v1; gvbase[o1 + c] += k1 * ga; } \
                if (o2 >= 0) { v2 = vbase[o2 + c]; gh -= lw * v2; gw += hh * v2; gvbase[o2 + c] += k2 * ga; } \
                if (o3 >= 0) { v3 = vbase[o3 + c]; gh += hw * v3; gw -= lh * v3; gvbase[o3 + c] += k3 * ga; } \
                if (o4 >= 0) { v4 = vbase[o4 + c]; gh += lw * v4; gw += lh * v4; gvbase[o4 + c] += k4 * ga; } \
                s_a += g * (k1 * v1 + k2 * v2 + k3 * v3 + k4 * v4);                                \
                s_w += (T)W * gw * ga;                                                             \
                s_h += (T)H * gh * ga;                                                             \
              }                                                                                    \
              glp[(l * P + p) * 2] = s_w;                                                          \
              glp[(l * P + p) * 2 + 1] = s_h;                                                      \
              gap[l * P + p] = s_a;                                                                \
            }                                                                                      \
          }                                                                                        \
        }                                                                                          \
      }                                                                                            \
    }                                                                                              \
  }

ORACLE_DEFINE(float, f32)
ORACLE_DEFINE(double, f64)

int msda_oracle_abi_version(void) { return 1; }
