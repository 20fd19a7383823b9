/*
 * msda_b200.h -- C ABI of the B200-native multi-scale deformable attention path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every entry point names
 * the reference interface it replaces (paths relative to
 * /root/reference/groundingdino/models/GroundingDINO/).  INTEGRATION.md shows the reference-side
 * binding (a ctypes stub standing in for `groundingdino._C`).
 *
 * Conventions
 *   - All data pointers are DEVICE pointers on the current CUDA device, contiguous row-major,
 *     16-byte aligned.  `spatial_shapes` / `level_start_index` are DEVICE int64 arrays, exactly as
 *     the reference op receives them (csrc/MsDeformAttn/ms_deform_attn_cuda.cu:36-37); kernels stage
 *     them in shared memory, so no host copy and no host synchronisation is needed.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls only
 *     enqueue work: no allocation, no synchronisation, no global state => re-entrant per stream,
 *     like the reference (ms_deform_attn_cuda.cu:66, at::cuda::getCurrentCUDAStream()).
 *   - Return value: 0 on success; MSDA_ERR_* (< 0) for rejected arguments; a positive cudaError_t
 *     when a launch failed.  Unlike the reference, launch errors are returned, not printf'd
 *     (ms_deform_im2col_cuda.cuh:948-952, :1321-1325).  msda_b200_last_error() gives the text.
 *   - Layouts: value [N][S][M][D]; loc [N][Lq][M][L][P][2] (x, y in [0,1]); aw [N][Lq][M][L][P];
 *     out / grad_out [N][Lq][M*D]; shapes [L][2] = (H, W); level_start [L].
 *   - Dtype suffix = storage type of value / out / grad_out.  For bf16 and f16 the sampling
 *     locations, attention weights and ALL gradients are fp32 and accumulation is fp32 (the
 *     reference up-casts fp16 to fp32 around the op, ms_deform_attn.py:326-344; bf16 is new).
 */
#ifndef MSDA_B200_H_
#define MSDA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSDA_B200_ABI_VERSION 1
#define MSDA_MAX_LEVELS 16

enum {
  MSDA_OK = 0,
  MSDA_ERR_NULL_POINTER = -1,
  MSDA_ERR_BAD_SHAPE = -2,     /* non-positive dim, L > MSDA_MAX_LEVELS, index range > int32 */
  MSDA_ERR_MISALIGNED = -3,    /* a pointer is not 16-byte aligned */
  MSDA_ERR_UNSUPPORTED = -4,   /* dtype / shape combination not built */
  MSDA_ERR_NO_DEVICE = -5      /* no sm_100 device current */
};

int msda_b200_abi_version(void);
/* Text of the last error raised on the calling thread ("" if none). */
const char *msda_b200_last_error(void);
/* 10*major+minor of the current device (100 on B200), or MSDA_ERR_NO_DEVICE. */
int msda_b200_device_arch(void);

/* ---- forward: replaces groundingdino._C.ms_deform_attn_forward --------------------------------
 * csrc/vision.cpp:54, csrc/MsDeformAttn/ms_deform_attn.h:21-40, ms_deform_attn_cuda.cu:21-81,
 * kernel ms_deform_im2col_cuda.cuh:237-299.  `out` is fully overwritten (no pre-zeroing needed;
 * the reference memsets it at ms_deform_attn_cuda.cu:55).  im2col_step is not part of this ABI:
 * it only chunks the batch (ms_deform_attn_cuda.cu:51-76); the host mirror keeps its observable
 * `batch % min(batch, step) == 0` check. */
int msda_forward_f32(const float *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                     const float *loc, const float *aw, int N, int S, int M, int D, int L, int Lq, int P,
                     float *out, void *stream);
int msda_forward_f64(const double *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                     const double *loc, const double *aw, int N, int S, int M, int D, int L, int Lq, int P,
                     double *out, void *stream);
/* value/out are bf16 (resp. IEEE half) bit patterns; loc/aw fp32 */
int msda_forward_bf16(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                      const float *loc, const float *aw, int N, int S, int M, int D, int L, int Lq, int P,
                      void *out, void *stream);
int msda_forward_f16(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                     const float *loc, const float *aw, int N, int S, int M, int D, int L, int Lq, int P,
                     void *out, void *stream);

/* ---- backward: replaces groundingdino._C.ms_deform_attn_backward ------------------------------
 * csrc/vision.cpp:55, ms_deform_attn.h:42-62, ms_deform_attn_cuda.cu:84-154, kernel
 * ms_deform_im2col_cuda.cuh:301-403 (+ :87-159).  grad_loc and grad_aw are fully overwritten.
 * grad_value is ACCUMULATED into with atomics (summation order is not deterministic, as in the
 * reference); pass zero_grad_value != 0 to have it memset on `stream` first (the reference always
 * does, ms_deform_attn_cuda.cu:122). */
int msda_backward_f32(const float *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                      const float *loc, const float *aw, const float *grad_out, int N, int S, int M, int D,
                      int L, int Lq, int P, float *grad_value, float *grad_loc, float *grad_aw,
                      int zero_grad_value, void *stream);
int msda_backward_f64(const double *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                      const double *loc, const double *aw, const double *grad_out, int N, int S, int M, int D,
                      int L, int Lq, int P, double *grad_value, double *grad_loc, double *grad_aw,
                      int zero_grad_value, void *stream);
/* value/grad_out bf16 (resp. half); grad_value, grad_loc, grad_aw fp32 */
int msda_backward_bf16(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                       const float *loc, const float *aw, const void *grad_out, int N, int S, int M, int D,
                       int L, int Lq, int P, float *grad_value, float *grad_loc, float *grad_aw,
                       int zero_grad_value, void *stream);
int msda_backward_f16(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                      const float *loc, const float *aw, const void *grad_out, int N, int S, int M, int D,
                      int L, int Lq, int P, float *grad_value, float *grad_loc, float *grad_aw,
                      int zero_grad_value, void *stream);

/* Backward with the query-side epilogue backward fused in (16-bit storage, D = 32, L = 4, P = 4): same scatter into
 * grad_value, but instead of grad_sampling_loc / grad_attn_weight it writes dq [N*Lq, 3*M*L*P] (16-bit) =
 * [d sampling_offsets | d attention logits], i.e. the backward of ms_deform_attn.py:296 (softmax) and :306-319
 * (offsets -> locations; ref is [N*Lq, L, ref_dim] fp32) finished in registers -- the operand of the
 * query-projection dgrad GEMM, with no fp32 round trip through HBM. */
int msda_backward_fusedq_16(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                            const float *loc, const float *aw, const void *grad_out, const float *ref, int ref_dim,
                            int N, int S, int M, int D, int L, int Lq, int P, float *grad_value, void *dq,
                            int zero_grad_value, int is_half, void *stream);

/* msda_backward_fusedq_16 with grad_value accumulated in SCALED fp16 instead of fp32 (same reference lines; the
 * reductions of ms_deform_im2col_cuda.cuh:87-159 leave the SM as 64-byte packed-half rows instead of 128-byte fp32 rows --
 * the SM->L2 reduction path is byte-bound on this part).
 *   grad_value_h : device buffer of 2 * N * rows_h * M * D + 128 bytes, rows_h = msda_grad_value_h16_rows(host copy of
 *                  spatial_shapes, L, Lq): per image, level l is stored as K_l replicas of its H_l*W_l rows (query q adds
 *                  into replica q mod K_l; K_l = f16acc_replicas(Lq, H_l*W_l) keeps the expected number of contributions
 *                  per fp16 row <= 96), followed by one 128-byte tail whose first word receives the bits of max |grad_out|.
 *                  The call zeroes all of it on `stream`, measures the maximum and accumulates scale * contribution with
 *                  scale = the power of two f16acc_scale(max, Lq) (msda_common.cuh), chosen so that no partial sum can
 *                  overflow fp16 whatever the sampling pattern.  spatial_shapes_host = the caller's HOST copy of spatial_shapes
 *                  (it sized the buffer from it; the replica layout is derived from it once per call instead of once per
 *                  CTA); if it disagrees with the device-side shapes the kernel writes nothing and poisons the tail (the
 *                  consumer then yields NaN).
 * Error vs fp32 accumulation: ~1e-3 of max |grad_value| (fp16 has 11 significand bits; the consumer rounds to bf16's 8
 * anyway) -- inside the 1e-2 bar of 16-bit storage, outside fp32's 1e-4, so fp32 storage never takes this path.
 * msda_cast_mask_h16 is the consumer: sum of replicas / scale -> [N*S, cols = M*D] 16-bit rows (out_f32 = 0; padded rows
 * zeroed: backward of ms_deform_attn.py:287-288) or fp32 rows (out_f32 = 1). */
long long msda_grad_value_h16_rows(const int64_t *spatial_shapes_host, int L, int Lq);
/* The scale the scatter kernel and the consumer derive from the bits of max |grad_out| (host evaluation of the same inline
 * function, for tests and diagnostics): a power of two s with s * max * Lq < 60000. */
float msda_f16acc_scale(unsigned int amax_bits, int Lq);
int msda_backward_fusedq_h16(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                             const float *loc, const float *aw, const void *grad_out, const float *ref, int ref_dim,
                             int N, int S, int M, int D, int L, int Lq, int P, void *grad_value_h,
                             const int64_t *spatial_shapes_host, void *dq, int is_half, void *stream);
int msda_cast_mask_h16(const void *grad_value_h, const int64_t *spatial_shapes, const int64_t *level_start_index, int L,
                       int N, int S, int cols, int Lq, long long rows_h, const uint8_t *row_mask, void *out, int out_f32,
                       int is_half, void *stream);

/* The 16-bit backward with a caller-provided scratch buffer.  Same results as msda_backward_{bf16,f16} (dq == NULL:
 * grad_loc / grad_aw written, ref ignored) or msda_backward_fusedq_16 (dq != NULL: grad_loc / grad_aw ignored); the
 * scratch lets the tensor-memory scatter own fine levels as well (tuning key "bwd_mma_levels"): the scatter kernel marks
 * there, per (image, head, 64-query chunk), which pixel ranges of the value map the chunk's samples touch, and the
 * accumulation kernel skips every (chunk, range) pair that is empty.  `workspace` = msda_backward_workspace_bytes(N, M, Lq)
 * bytes of device memory, 8-byte aligned, contents ignored and clobbered; NULL = behave as the entry points above. */
long long msda_backward_workspace_bytes(int N, int M, int Lq);
int msda_backward_16_ws(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index, const float *loc,
                        const float *aw, const void *grad_out, const float *ref, int ref_dim, int N, int S, int M, int D,
                        int L, int Lq, int P, float *grad_value, float *grad_loc, float *grad_aw, void *dq,
                        int zero_grad_value, int is_half, void *workspace, long long workspace_bytes, void *stream);

/* ---- tuning / introspection (not part of the reference interface) ------------------------------
 * Kernel variant knobs used by bench.py sweeps; defaults are the shipped configuration.
 *   key "fwd_sample_batch"  : samples whose corner loads are issued together (1, 2 or 4)
 *   key "fwd_q_fast"        : 1 = lanes of a warp span consecutive queries of one head, 0 = heads
 *   key "bwd_q_fast"        : same for the backward kernel
 *   key "fwd_passes" / "bwd_passes" : consecutive unit tiles per CTA (1..64; bwd_passes 0, the default, = by launch size)
 *   key "bwd_narrow"        : 16-bit storage, 1 = 4 channels per lane in the backward kernel (one full-line
 *                             reduction per corner), 0 = 8 channels per lane
 *   key "bwd_dots"          : 1 = the backward forms grad_sampling_loc / grad_attn_weight from the four corner dot products
 *                             <grad_out row, value row> of a sample (4 FMAs per channel) instead of per-channel bilinear
 *                             derivatives (14 operations per channel); same mathematics, fp32 round-off moves by ~1e-6
 *   key "tap_share"         : 1 = the four sampling taps of a level are computed once per lane group and exchanged with
 *                             shuffles (bit-identical values) instead of once per lane; vector kernels with >= 4 lanes per unit
 *   key "bwd_mma"           : 16-bit storage, D = 32, P = 4: 1 = the coarse tail of the level list (<= 1536 pixels, <= 4
 *                             levels) accumulates grad_value in tensor memory (tcgen05) and leaves the SM once per
 *                             (image, head) instead of once per bilinear corner; 0 = every corner is a reduction
 *   key "bwd_mma_min_units" : smallest N*Lq*M for which bwd_mma applies (default 131072; 0 = always)
 *   key "bwd_mma_levels"    : with a workspace (msda_backward_16_ws): how many levels, counted from the coarse end, the
 *                             range-planned tensor-memory scatter owns (0 = only the coarse tail, first-generation kernel)
 * Returns 0, or MSDA_ERR_UNSUPPORTED for an unknown key / value. */
int msda_b200_set_tuning(const char *key, int value);
int msda_b200_get_tuning(const char *key);
/* Number of kernel launches issued through this library by the calling process so far. */
long long msda_b200_launch_count(void);

/* ---- projections: replace the four nn.Linear calls of MultiScaleDeformableAttention.forward ------
 * (ms_deform_attn.py:184-187 parameters; :286 value_proj, :290 sampling_offsets, :293 attention_weights,
 * :350 output_proj) and the elementwise passes between them, as tcgen05/TMA GEMMs over 16-bit (bf16, or
 * IEEE half when is_half != 0) activations with fp32 accumulation.  X is [R, K] row-major, W is the
 * nn.Linear weight [Nout, K] row-major, bias fp32.  K % 64 == 0; Nout % 32 == 0; Nout <= 2048.
 *
 * msda_linear_16: out[r, :] = X[r, :] W^T + bias, rows with row_mask[r] != 0 written as zero (the
 *   masked_fill of :287-288; pass NULL for none).  out is 16-bit (out_f32 == 0) or fp32, leading
 *   dimension out_ld elements.  Also used for the dgrad products (pass W^T as `w`).
 * msda_query_proj_16: the sampling_offsets and attention_weights linears share their input, so they run as
 *   ONE GEMM against w_cat = [W_offsets; W_attention] ([3*M*L*P, K]); the epilogue turns offsets into
 *   sampling locations (:306-319; ref is [R, L, ref_dim], ref_dim 2 or 4) and logits into softmax
 *   weights (:296), writing loc_out [R, M, L, P, 2] and aw_out [R, M, L, P] in fp32 -- exactly the two
 *   tensors msda_forward_* consumes.  Needs M*L*P % 32 == 0; when L*P does not divide 32 (L = 5) the logits are stored
 *   raw and a second small kernel applies the softmax in place.  msda_query_bwd_prep_16 writes rows of ld_out >= 3*M*L*P
 *   elements (zero padded: the row is the K dimension of the dgrad GEMM, which needs a multiple of 64).
 * msda_query_bwd_prep_16 / msda_cast_mask_16: elementwise backward companions (see proj_elementwise.cu). */
int msda_linear_16(const void *x, const void *w, const float *bias, long long R, int K, int Nout, void *out,
                   int out_ld, int out_f32, const uint8_t *row_mask, int is_half, void *stream);
/* out = accum + x W^T + bias, all 16-bit [R, Nout]: the dgrad of a projection added straight onto the gradient that is
 * already there (residual path, other consumers of the same activation) instead of a separate elementwise add pass.
 * `out` may alias `accum`.  Nout % 8 == 0. */
int msda_linear_accum_16(const void *x, const void *w, const float *bias, long long R, int K, int Nout, const void *accum,
                         void *out, int is_half, void *stream);
/* The same accumulating product over TWO activation operands, K-concatenated without a copy: out = accum + [x1 | x2] w^T,
 * x1 [R, K1], x2 [R, K2] (K1, K2 multiples of 64), w [Nout, K1 + K2].  The encoder block's two input dgrads
 * (d value_proj input + d query, ms_deform_attn.py:286 and :290-293 backward) as one launch. */
int msda_linear_accum2_16(const void *x1, int K1, const void *x2, int K2, const void *w, const float *bias, long long R,
                          int Nout, const void *accum, void *out, int is_half, void *stream);
/* Projection + residual + LayerNorm in ONE launch (`src = norm1(src + output_proj(x))`, transformer_for_adapter.py:901-902 with
 * ms_deform_attn.py:350): z = residual + x w^T + bias (stored, 16-bit: LayerNorm's saved input), y = LayerNorm(z) * gamma +
 * beta with the statistics of the ROUNDED z (as msda_add_layernorm_fwd_16 computes them), mean / rstd [R] fp32 for
 * msda_add_layernorm_bwd_16.  Nout in {128, 256} (a GEMM tile holds whole rows), K % 64 == 0. */
int msda_linear_add_layernorm_16(const void *x, const void *w, const float *bias, long long R, int K, int Nout,
                                 const void *residual, const float *gamma, const float *beta, float eps, void *z, void *y,
                                 float *mean, float *rstd, int is_half, void *stream);
/* FFN companions (row N1; reference transformer_for_adapter.py:876-885): out = relu(x W^T + bias) when relu != 0, and
 * out = (x W^T + bias) where gate > 0 else 0 when gate != NULL (gate: 16-bit [R, Nout]) -- the ReLU backward fused into
 * the dgrad GEMM of linear2.  16-bit output, leading dimension Nout; Nout <= 2048. */
int msda_linear_act_16(const void *x, const void *w, const float *bias, long long R, int K, int Nout, void *out, int relu,
                       const void *gate, int is_half, void *stream);
/* The same two fusions with the ReLU mask kept as ONE BIT per activation instead of re-reading the 16-bit activation in
 * the backward: a forward launch (relu_bits_out != NULL) applies bias + ReLU and also writes bits[(c/32) * R + row]
 * (bit j = out[row, c + j] > 0; [Nout/32, R] uint32, word-major); the backward launch (gate_bits != NULL) keeps acc where
 * the bit is set.  Exactly one of the two pointers must be non-NULL.  Nout % 32 == 0. */
int msda_linear_act_bits_16(const void *x, const void *w, const float *bias, long long R, int K, int Nout, void *out,
                            uint32_t *relu_bits_out, const uint32_t *gate_bits, int is_half, void *stream);
int msda_query_proj_16(const void *query, const void *w_cat, const float *bias_cat, const float *ref, int ref_dim,
                       const int64_t *spatial_shapes, long long R, int K, int M, int L, int P, float *loc_out,
                       float *aw_out, int is_half, void *stream);
/* msda_query_proj_16 of (query + query_add) without forming the sum: both operands stream against ONE resident copy of
 * w_cat and the two products accumulate in fp32 (`query = src + pos`, transformer_for_adapter.py:867-869 / :893, folded
 * into the projection).  query_add == NULL is msda_query_proj_16. */
int msda_query_proj2_16(const void *query, const void *query_add, const void *w_cat, const float *bias_cat, const float *ref,
                        int ref_dim, const int64_t *spatial_shapes, long long R, int K, int M, int L, int P, float *loc_out,
                        float *aw_out, int is_half, void *stream);
int msda_query_bwd_prep_16(const float *grad_loc, const float *grad_aw, const float *aw, const float *ref, int ref_dim,
                           const int64_t *spatial_shapes, long long R, int M, int L, int P, void *out, int ld_out,
                           int is_half, void *stream);
/* ---- fp32 modules (the reference's default precision: AMP off, groundingdino/config/configs/common/train.py:11) -----
 * The same projections for fp32 activations, on tcgen05 as THREE kind::tf32 products per tile (x_hi w_hi + x_lo w_hi +
 * x_hi w_lo, fp32 accumulation in tensor memory; hi = upper 19 bits of the fp32 pattern, lo = the exact remainder):
 * ~1e-6 relative, inside north_star's 1e-5 / 1e-4 bars where a single TF32 product is not.  X [R, K] fp32; the weight is
 * passed pre-split as w_split = [w_hi; w_lo] ([2*Nout, K] fp32: w_hi = W with the 13 low mantissa bits cleared,
 * w_lo = W - w_hi).  K % 32 == 0, Nout % 64 == 0, Nout <= 768.  The activation split happens in shared memory.
 * msda_linear_f32: out[r, :] = (accum ? accum[r, :] : 0) + X[r, :] W^T + bias; rows with row_mask[r] != 0 are written as
 *   zero.  out / accum fp32 with leading dimension out_ld (out may alias accum).
 * msda_query_proj_f32: the fused [sampling_offsets; attention_weights] projection with the sampling-location + softmax
 *   epilogue of msda_query_proj_16 (needs 3*M*L*P % 64 == 0).  Non-finite inputs give NaN where the reference gives Inf. */
int msda_linear_f32(const float *x, const float *w_split, const float *bias, long long R, int K, int Nout, const float *accum,
                    float *out, int out_ld, const uint8_t *row_mask, void *stream);
int msda_query_proj_f32(const float *query, const float *w_cat_split, const float *bias_cat, const float *ref, int ref_dim,
                        const int64_t *spatial_shapes, long long R, int K, int M, int L, int P, float *loc_out,
                        float *aw_out, void *stream);
/* Elementwise tail of ms_deform_attn.py:290-319 on its own, for fp32 shapes the fused projection does not take: raw =
 * [offsets | logits] pre-activations [R, 3*M*L*P] -> loc_out [R, M, L, P, 2], aw_out [R, M, L, P].  Its backward is
 * msda_query_bwd_prep_16 with is_half = 2 (fp32 output, ld_out >= 3*M*L*P). */
int msda_query_post_f32(const float *raw, const float *ref, int ref_dim, const int64_t *spatial_shapes, long long R, int M,
                        int L, int P, float *loc_out, float *aw_out, void *stream);
int msda_cast_mask_16(const float *in, const uint8_t *row_mask, long long rows, int cols, void *out, int is_half,
                      void *stream);
/* ZiRa training-mode projection (semantics of RepZeroLinear, groundingdino_dual_zero_rep_branch.py:119-125, beside a
 * frozen nn.Linear): ONE GEMM over the stacked weight [W_0; W_f; W_b] -- rows interleaved in runs of 32 output
 * features (base 0-31, soft-frozen 0-31, branch 0-31, base 32-63, ...) -- whose epilogue forms
 *   branch = s*(x W_b^T + b_b), adapter = branch + x W_f^T + b_f, out = x W_0^T + b_0 + adapter  (masked rows -> 0)
 * and reduces the zero-inter loss terms: loss_sums[0] += sum SmoothL1(branch), loss_sums[1] += sum SmoothL1(adapter)
 * (caller zeroes loss_sums and divides by R*F).  bias3 = [b_0 | b_f | b_b] fp32, scaling = device scalar.  pre_out
 * (x W_b^T + b_b) and adapter_out, both [R, F] 16-bit, are what the backward needs; either may be NULL.
 * msda_zira_bwd_prep_16 builds the K-stacked operand [dY_eff | dO | dB] ([R, 3F]) of the dgrad GEMM
 * (msda_linear_16 against [W_0^T | W_f^T | s W_b^T]) and adds d(loss)/d(scaling) = sum dB * pre to *ds_out
 * (pre-zeroed fp32 scalar, may be NULL).  colsum_out (pre-zeroed fp32 [3F], may be NULL; needs 256 % (F/8) == 0)
 * receives the column sums of the three blocks before rounding = the bias gradients of W_0 | W_f | W_b / s. */
int msda_zira_linear_16(const void *x, const void *w_stack, const float *bias3, const float *scaling, long long R, int K,
                        int F, void *out, const uint8_t *row_mask, void *pre_out, void *adapter_out, float *loss_sums,
                        int is_half, void *stream);
int msda_zira_bwd_prep_16(const void *dy, const void *pre, const void *adapter, const uint8_t *row_mask,
                          const float *scaling, const float *dloss, long long R, int F, void *out, float *ds_out,
                          float *colsum_out, int is_half, void *stream);
const char *msda_b200_gemm_last_error(void);
/* A/B switch for benchmarks: 1 (default) keeps each CTA's slice of W resident in shared memory when it fits. */
int msda_b200_gemm_set_resident(int on);
/* A/B switch: 1 (default) stores the epilogue through TMA from 128B-swizzled tiles; 0 = shared-memory transposition +
 * coalesced st.global. */
int msda_b200_gemm_set_staged(int on);
/* A/B switch: store tiles per epilogue warp (1 or 2) and whether a resident W slice is kept in preference to the second tile. */
int msda_b200_gemm_set_store_bufs(int bufs, int prefer_resident);

/* ---- the op's immediate caller: residual + LayerNorm ------------------------------
 * `src = norm(src + src2)` of the reference's DeformableTransformerEncoderLayer (transformer_for_adapter.py:901-902 after
 * the attention, :882-885 after the FFN; dropout p = 0), one pass per direction over 16-bit activations [R, C] with fp32
 * statistics.  fwd writes z = x + r (LayerNorm's saved input), y, mean[R], rstd[R]; bwd returns dz (gradient of both x
 * and r).  C % 8 == 0, C <= 1024.  gamma / beta fp32. */
int msda_add_layernorm_fwd_16(const void *x, const void *r, const float *gamma, const float *beta, long long R, int C,
                              float eps, void *z, void *y, float *mean, float *rstd, int is_half, void *stream);
/* Plain LayerNorm with the same kernel (no residual operand): y = LN(x); optionally y2 = y + shift2[column] as a second output
 * (the residual operand of an output projection whose bias has been folded into it, fuse_modules.py).  Its backward is
 * msda_add_layernorm_bwd_16 with z = x. */
int msda_layernorm_fwd_16(const void *x, const float *gamma, const float *beta, long long R, int C, float eps, void *y,
                          void *y2, const float *shift2, float *mean, float *rstd, int is_half, void *stream);
int msda_add_layernorm_bwd_16(const void *dy, const void *z, const float *gamma, const float *mean, const float *rstd,
                              long long R, int C, void *dz, int is_half, void *stream);

/* ---- the step in front of the encoder: GroupNorm of the input projection ----------
 * `GroupNorm(32, hidden_dim)` that closes each level of the reference's input_proj (groundingdino_dual_zero_rep_branch.py:
 * 258-277; with the ZiRa conv adapter :492-493) on channels-last rows: image n's HW rows of C 16-bit channels start at
 * x + n * x_image_stride (elements), so x / y / dy / dx may be level slices of a flattened [N, S, C] tensor (the layout
 * the projection GEMM writes and the encoder reads).  gamma / beta fp32 [C].  scratch = N*C*2 doubles, zeroed by the
 * caller.  fwd writes y and mean_rstd [N, G, 2] fp32.  bwd writes dx; on return scratch[n][c] = (sum dy*xhat, sum dy),
 * whose sums over n are d gamma / d beta.  C % 8 == 0, C % G == 0, C/8 a power of two <= 256. */
int msda_group_norm_fwd_16(const void *x, long long x_image_stride, const float *gamma, const float *beta, int N,
                           long long HW, int C, int G, float eps, void *y, long long y_image_stride, float *mean_rstd,
                           double *scratch, int is_half, void *stream);
int msda_group_norm_bwd_16(const void *dy, long long dy_image_stride, const void *x, long long x_image_stride,
                           const float *gamma, const float *mean_rstd, int N, long long HW, int C, int G, void *dx,
                           long long dx_image_stride, double *scratch, int is_half, void *stream);

/* ---- the whole FFN of the layer as ONE launch per direction (row N1; transformer_for_adapter.py:876-885) ------------
 * msda_ffn_chain_fwd_16: out[R, C] = relu(x W1^T + b1) W2^T + b2 with the [R, F] hidden activation kept on chip (128 rows
 *   x 128 hidden columns at a time: TMEM -> registers -> shared memory -> second tcgen05 product); relu_bits_out [F/32, R]
 *   receives one bit per hidden activation (same layout as msda_linear_act_bits_16).  x, W1 [F, C], W2 [C, F], out 16-bit;
 *   b1 [F], b2 [C] fp32.  C == 256, F % 128 == 0.
 * msda_ffn_chain_bwd_16: dx[R, C] = accum + gate(dy W2) W1 with the same chaining: w2_t = W2^T [F, C], w1_t = W1^T [C, F],
 *   gate_bits from the forward; accum (16-bit [R, C], may be NULL) and dx may alias dy. */
int msda_ffn_chain_fwd_16(const void *x, const void *w1, const float *b1, const void *w2, const float *b2, long long R, int C,
                          int F, void *out, uint32_t *relu_bits_out, int is_half, void *stream);
/* msda_ffn_chain_ln_fwd_16: the same launch with the layer's `src = norm2(src + src2)` (transformer_for_adapter.py:882-885) in its
 * final stage: z = x + ffn(x) (16 bit; LayerNorm's saved input), y = LayerNorm(z) * gamma + beta, mean / rstd [R] fp32. */
int msda_ffn_chain_ln_fwd_16(const void *x, const void *w1, const float *b1, const void *w2, const float *b2, long long R, int C,
                             int F, const float *gamma, const float *beta, float eps, void *z, void *y, float *mean,
                             float *rstd, uint32_t *relu_bits_out, int is_half, void *stream);
/* The same three entry points on CTA pairs (tcgen05 cta_group::2, clusters of two: each SM loads half of every weight tile). */
int msda_ffn_chain2_fwd_16(const void *x, const void *w1, const float *b1, const void *w2, const float *b2, long long R, int C,
                           int F, void *out, uint32_t *relu_bits_out, int is_half, void *stream);
int msda_ffn_chain2_ln_fwd_16(const void *x, const void *w1, const float *b1, const void *w2, const float *b2, long long R,
                              int C, int F, const float *gamma, const float *beta, float eps, void *z, void *y, float *mean,
                              float *rstd, uint32_t *relu_bits_out, int is_half, void *stream);
int msda_ffn_chain2_bwd_16(const void *dy, const void *w2_t, const void *w1_t, const uint32_t *gate_bits, const void *accum,
                           long long R, int C, int F, void *dx, int is_half, void *stream);
int msda_ffn_chain_bwd_16(const void *dy, const void *w2_t, const void *w1_t, const uint32_t *gate_bits, const void *accum,
                          long long R, int C, int F, void *dx, int is_half, void *stream);

/* ---- the steps either side of the stacks --------------------------------------------
 * msda_flatten_levels: transformer_for_adapter.py:238-262 as one launch.  src_levels / pos_levels / mask_levels are HOST
 *   arrays of L device pointers: per level an NCHW map [N, C, H*W] (dtype: 0 bf16, 1 f16, 2 f32), its position embedding
 *   (same shape; pos_levels may be NULL) and its padding mask [N, H*W] bytes (mask_levels may be NULL); level_hw[l] = H*W.
 *   Writes src_out [N, S, C], pos_out [N, S, C] = pos + level_embed[l] (level_embed [L, C] device, may be NULL) and
 *   mask_out [N, S].  Any of the three outputs may be NULL.
 * msda_level_valid_counts: counts[n][l] = (valid_W, valid_H) = unmasked pixels of the first row / first column of level l
 *   (utils.py:74-75; divided by (W, H) they are get_valid_ratio, transformer_for_adapter.py:216-223).
 * msda_encoder_proposals: gen_encoder_output_proposals (utils.py:56-116): proposals_out [N, S, 4] fp32 = inverse sigmoid of
 *   ((x + 0.5) / valid_W, (y + 0.5) / valid_H, wh_base * 2^level), +inf where the token is padded or any component leaves
 *   (0.01, 0.99); memory_out = memory with exactly those rows zeroed.  wh_base: 2 device floats (0.05, 0.05 or
 *   sigmoid(learnedwh)); rows are row_bytes long (multiple of 16). */
int msda_flatten_levels(const void *const *src_levels, const void *const *pos_levels, const uint8_t *const *mask_levels,
                        const int *level_hw, int L, int N, int C, const void *level_embed, int dtype, void *src_out,
                        void *pos_out, uint8_t *mask_out, void *stream);
int msda_level_valid_counts(const uint8_t *mask_flatten, const int64_t *spatial_shapes, const int64_t *level_start_index, int N,
                            int S, int L, int *counts, void *stream);
int msda_encoder_proposals(const void *memory, const uint8_t *mask_flatten, const int64_t *spatial_shapes,
                           const int64_t *level_start_index, const int *counts, const float *wh_base, int N, int S, int L,
                           int row_bytes, void *memory_out, float *proposals_out, void *stream);

/* ---- image <-> text fusion attention -------------------------------------------------
 * Replaces the attention core of BiMultiHeadAttention.forward (reference fuse_modules.py:172-227: fp32 logits
 * [B*heads, n_img, n_text], their transpose, two softmaxes, two batched products) and autograd through it.  Heads are 256
 * wide; operands are the projections as they leave their GEMMs, [B, L, H*256] 16-bit, heads addressed by column offset.
 * Both directions share the logits s = scale * a . b^T, so every entry point exists in two ORIENTATIONS: `a` is the
 * stationary side (LA rows, 128 per CTA), `b` / `x` the streamed side (LB rows).  Statistics are log2-domain
 * log-sum-exp values, fp32 [B, H, pad128(L)] with +inf in the padding; masks are bytes [B, pad128(L)], 1 = masked, padding 1.
 *
 * msda_biattn_pv_16: out[B, LA, H*256] = P . x.
 *   col_stat == NULL ("online"): P = softmax over the LB axis of s (mask_padded masks streamed rows); lane_stat receives
 *     the statistics.  out_v = pv(q, k, val_l, mask_l), out_l = pv(k, q, val_v, mask_v).
 *   col_stat != NULL ("given"):  P[i, j] = exp2(s2[i, j] - col_stat[j]), zero for masked stationary rows i (mask_padded
 *     masks stationary rows): the other direction's probabilities, transposed -- d val_l = pv(k, q, d out_v, stat_v),
 *     d val_v = pv(q, k, d out_l, stat_l).
 *   nsplit > 1 splits the streamed axis over CTAs: partial results go to part_o [items, 128, 256] fp32 (+ part_m, part_l
 *   [items, 128] online), items = B * H * ceil(LA/128) * msda_biattn_splits(LB, nsplit); msda_biattn_combine_16 merges
 *   them (softmax-weighted when given == 0, plain sum otherwise) into out16 / lane_stat.
 * msda_biattn_rowdot_16: delta[B, H, lpad] = per-head dot product of d_o and o ([B, L, H*256]).
 * msda_biattn_ds_16: d a[B, LA, H*256] = scale * dS . b with dS[i, j] = P1 (d_oa[i].xb[j] - lane_delta[i]) +
 *   P2 (xa[i].d_ob[j] - col_delta[j]), P1 = exp2(s2 - lane_stat[i]) masked by mask_b (streamed rows), P2 = exp2(s2 -
 *   col_stat[j]) masked by mask_a (stationary rows).  d q = ds(q, d out_v, val_v, k, val_l, d out_l, ...),
 *   d k = ds(k, d out_l, val_l, q, val_v, d out_v, ...).  nsplit as above (64-row column tiles; msda_biattn_ds_splits),
 *   partials merged by msda_biattn_combine_16(given = 1).
 * msda_biattn_ds_terms_16: the unsplit msda_biattn_ds_16 that also stores dS (16 bit, [B, H, LA, ceil(LB/64)*64]; pass 0
 *   stores its term, pass 1 adds its own with a TMA reduction); msda_biattn_tn_16: d b = scale * dS^T . a from it, i.e.
 *   d k = tn(dS of the d q launch, q) without a second recomputation (image axis split as above, msda_biattn_tn_splits).
 * msda_biattn_set_trace: debug -- 16 int64 cycle counters per CTA of the next msda_biattn_pv_16 launches (NULL = off). */
int msda_biattn_splits(int LB, int nsplit);
int msda_biattn_pv_16(const void *a, const void *b, const void *x, int B, int H, int LA, int LB, float scale,
                      const uint8_t *mask_padded, const float *col_stat, void *out16, float *lane_stat, float *part_o,
                      float *part_m, float *part_l, int nsplit, int is_half, void *stream);
int msda_biattn_combine_16(const float *part_o, const float *part_m, const float *part_l, int B, int H, int LA, int nsplit,
                           int given, void *out16, float *lane_stat, int is_half, void *stream);
int msda_biattn_rowdot_16(const void *d_o, const void *o, int B, int L, int H, int lpad, float *delta, int is_half,
                          void *stream);
int msda_biattn_ds_splits(int LB, int nsplit);
int msda_biattn_ds_16(const void *a, const void *d_oa, const void *xa, const void *b, const void *xb, const void *d_ob, int B,
                      int H, int LA, int LB, float scale, const uint8_t *mask_a_padded, const uint8_t *mask_b_padded,
                      const float *lane_stat, const float *lane_delta, const float *col_stat, const float *col_delta,
                      void *out16, float *part_o, int nsplit, int is_half, void *stream);
int msda_biattn_ds_terms_16(const void *a, const void *d_oa, const void *xa, const void *b, const void *xb, const void *d_ob,
                            int B, int H, int LA, int LB, float scale, const uint8_t *mask_a_padded,
                            const uint8_t *mask_b_padded, const float *lane_stat, const float *lane_delta, const float *col_stat,
                            const float *col_delta, void *out16, void *ds16, int is_half, void *stream);
int msda_biattn_tn_splits(int S, int nsplit);
int msda_biattn_tn_16(const void *ds16, const void *q, int B, int H, int S, int T, float scale, void *out16, float *part_o,
                      int nsplit, int is_half, void *stream);
void msda_biattn_set_trace(long long *buf);

/* Measurement aid: random seg_bytes-aligned (64, 128 or 512) segment reads from `buf` (bytes long,
 * keep it L2-sized), `iters` segments per lane group, `blocks` CTAs of 256 threads.  Bytes moved =
 * blocks * 256 * 16 * iters (iters rounded up to a multiple of 8).  `sink` is 4 writable bytes. */
int msda_b200_probe_gather(const void *buf, long long bytes, int seg_bytes, int iters, int blocks, void *sink,
                           void *stream);
/* Scatter twin: groups of lanes send 128-byte fp32 rows to random 128 B-aligned rows of `buf` (bytes long, fp32,
 * contents are modified).  mode 0 = red.global.add.v4.f32 x 8 lanes (the backward kernel's instruction), 1 = st.global.v4
 * x 8 lanes, 2 = scalar red.global.add.f32 x 32 lanes.  Row payload moved = blocks * 256 * iters * (mode == 2 ? 4 : 16) bytes. */
int msda_b200_probe_scatter(void *buf, long long bytes, int mode, int iters, int blocks, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MSDA_B200_H_ */
