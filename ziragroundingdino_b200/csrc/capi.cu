// extern "C" surface of libmsda_b200.so (declared in include/msda_b200.h).
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "msda_common.cuh"

namespace msda {
template <typename VT>
cudaError_t forward_f32acc(const VT*, const int64_t*, const int64_t*, const float*, const float*, VT*, int, int, int, int, int, int, int, cudaStream_t);
template <typename VT>
cudaError_t backward_f32acc(const VT*, const int64_t*, const int64_t*, const float*, const float*, const VT*, float*, float*, float*, int, int, int, int, int, int, int, cudaStream_t);
template <typename VT>
cudaError_t backward_fused_q(const VT*, const int64_t*, const int64_t*, const float*, const float*, const VT*, float*, const float*, int, void*, int, int, int, int, int, cudaStream_t);
template <typename VT>
cudaError_t backward_fused_q_h16(const VT*, const int64_t*, const int64_t*, const float*, const float*, const VT*, void*, const int64_t*, uint32_t*, const float*, int, void*, int, int, int, int, int, cudaStream_t);
cudaError_t forward_f64(const double*, const int64_t*, const int64_t*, const double*, const double*, double*, int, int, int, int, int, int, int, cudaStream_t);
cudaError_t backward_f64(const double*, const int64_t*, const int64_t*, const double*, const double*, const double*, double*, double*, double*, int, int, int, int, int, int, int, cudaStream_t);
}  // namespace msda

namespace {
thread_local char t_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_status(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return MSDA_OK;
  snprintf(t_err, sizeof(t_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return static_cast<int>(e);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int check_common(const void* value, const void* shapes, const void* lstart, const void* loc, const void* aw,
                 int N, int S, int M, int D, int L, int Lq, int P, size_t elem) {
  if (!value || !shapes || !lstart || !loc || !aw) return fail(MSDA_ERR_NULL_POINTER, "null input pointer");
  if (N < 0 || S <= 0 || M <= 0 || D <= 0 || L <= 0 || Lq < 0 || P <= 0)
    return fail(MSDA_ERR_BAD_SHAPE, "non-positive dimension (N=%d S=%d M=%d D=%d L=%d Lq=%d P=%d)", N, S, M, D, L, Lq, P);
  if (L > MSDA_MAX_LEVELS) return fail(MSDA_ERR_BAD_SHAPE, "num_levels %d > MSDA_MAX_LEVELS %d", L, MSDA_MAX_LEVELS);
  // per-image value offsets are 32-bit inside the kernels (batch offsets are 64-bit)
  if (static_cast<long long>(S) * M * D >= (1ll << 31))
    return fail(MSDA_ERR_BAD_SHAPE, "S*M*D = %lld does not fit 32-bit per-image indexing", static_cast<long long>(S) * M * D);
  if (static_cast<long long>(N) * Lq * M / 32 >= (1ll << 31))
    return fail(MSDA_ERR_BAD_SHAPE, "N*Lq*M too large for one launch");
  if (!aligned16(value) || !aligned16(loc) || !aligned16(aw) || (reinterpret_cast<uintptr_t>(shapes) & 7u) ||
      (reinterpret_cast<uintptr_t>(lstart) & 7u))
    return fail(MSDA_ERR_MISALIGNED, "input pointers must be 16-byte aligned (int64 arrays 8-byte)");
  (void)elem;
  return MSDA_OK;
}

template <typename VT, typename Fn>
int run_forward(const VT* value, const int64_t* shapes, const int64_t* lstart, const void* loc, const void* aw,
                int N, int S, int M, int D, int L, int Lq, int P, VT* out, void* stream, Fn fn) {
  t_err[0] = 0;
  int rc = check_common(value, shapes, lstart, loc, aw, N, S, M, D, L, Lq, P, sizeof(VT));
  if (rc) return rc;
  if (!out) return fail(MSDA_ERR_NULL_POINTER, "null output pointer");
  if (!aligned16(out)) return fail(MSDA_ERR_MISALIGNED, "out must be 16-byte aligned");
  if (N == 0 || Lq == 0) return MSDA_OK;  // empty batch / no queries: nothing to write
  return cuda_status(fn(static_cast<cudaStream_t>(stream)), "msda_forward launch");
}

template <typename VT, typename GT, typename Fn>
int run_backward(const VT* value, const int64_t* shapes, const int64_t* lstart, const void* loc, const void* aw,
                 const VT* grad_out, int N, int S, int M, int D, int L, int Lq, int P, GT* gv, GT* gl, GT* ga,
                 int zero_gv, void* stream, Fn fn) {
  t_err[0] = 0;
  int rc = check_common(value, shapes, lstart, loc, aw, N, S, M, D, L, Lq, P, sizeof(VT));
  if (rc) return rc;
  if (!grad_out || !gv || !gl || !ga) return fail(MSDA_ERR_NULL_POINTER, "null gradient pointer");
  if (!aligned16(grad_out) || !aligned16(gv) || !aligned16(gl) || !aligned16(ga))
    return fail(MSDA_ERR_MISALIGNED, "gradient pointers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (zero_gv && N > 0) {
    rc = cuda_status(cudaMemsetAsync(gv, 0, sizeof(GT) * static_cast<size_t>(N) * S * M * D, st), "grad_value memset");
    if (rc) return rc;
  }
  if (N == 0 || Lq == 0) return MSDA_OK;
  return cuda_status(fn(st), "msda_backward launch");
}
}  // namespace

extern "C" {

int msda_b200_abi_version(void) { return MSDA_B200_ABI_VERSION; }
const char* msda_b200_last_error(void) { return t_err; }
long long msda_b200_launch_count(void) { return msda::g_launches; }

int msda_b200_device_arch(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return fail(MSDA_ERR_NO_DEVICE, "no CUDA device"); }
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  return major * 10 + minor;
}

static int* tuning_slot(const char* key) {
  if (!key) return nullptr;
  if (!strcmp(key, "fwd_sample_batch")) return &msda::g_tuning.fwd_sample_batch;
  if (!strcmp(key, "fwd_q_fast")) return &msda::g_tuning.fwd_q_fast;
  if (!strcmp(key, "fwd_passes")) return &msda::g_tuning.fwd_passes;
  if (!strcmp(key, "bwd_q_fast")) return &msda::g_tuning.bwd_q_fast;
  if (!strcmp(key, "bwd_passes")) return &msda::g_tuning.bwd_passes;
  if (!strcmp(key, "bwd_narrow")) return &msda::g_tuning.bwd_narrow;
  if (!strcmp(key, "bwd_dots")) return &msda::g_tuning.bwd_dots;
  if (!strcmp(key, "tap_share")) return &msda::g_tuning.tap_share;
  if (!strcmp(key, "bwd_mma")) return &msda::g_tuning.bwd_mma;
  if (!strcmp(key, "bwd_mma_min_units")) return &msda::g_tuning.bwd_mma_min_units;
  if (!strcmp(key, "bwd_mma_levels")) return &msda::g_tuning.bwd_mma_levels;
  return nullptr;
}

int msda_b200_set_tuning(const char* key, int value) {
  int* slot = tuning_slot(key);
  if (!slot) return fail(MSDA_ERR_UNSUPPORTED, "unknown tuning key '%s'", key ? key : "(null)");
  const bool is_batch = slot == &msda::g_tuning.fwd_sample_batch;
  const bool is_pass = slot == &msda::g_tuning.fwd_passes || slot == &msda::g_tuning.bwd_passes;
  if (slot == &msda::g_tuning.bwd_mma_levels) {
    if (value < 0 || value > MSDA_MAX_LEVELS) return fail(MSDA_ERR_UNSUPPORTED, "bad value %d for tuning key '%s'", value, key);
    *slot = value;
    return MSDA_OK;
  }
  if (slot == &msda::g_tuning.bwd_mma_min_units) {
    if (value < 0) return fail(MSDA_ERR_UNSUPPORTED, "bad value %d for tuning key '%s'", value, key);
    *slot = value;
    return MSDA_OK;
  }
  const int pass_min = slot == &msda::g_tuning.bwd_passes ? 0 : 1;      // bwd_passes 0 = chosen by launch size
  if ((is_batch && value != 1 && value != 2 && value != 4) || (is_pass && (value < pass_min || value > 64)) ||
      (!is_batch && !is_pass && value != 0 && value != 1))
    return fail(MSDA_ERR_UNSUPPORTED, "bad value %d for tuning key '%s'", value, key);
  *slot = value;
  return MSDA_OK;
}

int msda_b200_get_tuning(const char* key) {
  int* slot = tuning_slot(key);
  return slot ? *slot : fail(MSDA_ERR_UNSUPPORTED, "unknown tuning key '%s'", key ? key : "(null)");
}

#define MSDA_DEFINE_F32ACC(SUFFIX, CTYPE, VT)                                                                      \
  int msda_forward_##SUFFIX(const CTYPE* value, const int64_t* shapes, const int64_t* lstart, const float* loc,  \
                            const float* aw, int N, int S, int M, int D, int L, int Lq, int P, CTYPE* out,        \
                            void* stream) {                                                                        \
    const VT* v = reinterpret_cast<const VT*>(value);                                                              \
    VT* o = reinterpret_cast<VT*>(out);                                                                            \
    return run_forward<VT>(v, shapes, lstart, loc, aw, N, S, M, D, L, Lq, P, o, stream, [&](cudaStream_t st) {     \
      return msda::forward_f32acc<VT>(v, shapes, lstart, loc, aw, o, N, S, M, D, L, Lq, P, st);                    \
    });                                                                                                            \
  }                                                                                                                \
  int msda_backward_##SUFFIX(const CTYPE* value, const int64_t* shapes, const int64_t* lstart, const float* loc, \
                             const float* aw, const CTYPE* grad_out, int N, int S, int M, int D, int L, int Lq,   \
                             int P, float* gv, float* gl, float* ga, int zero_gv, void* stream) {                  \
    const VT* v = reinterpret_cast<const VT*>(value);                                                              \
    const VT* g = reinterpret_cast<const VT*>(grad_out);                                                           \
    return run_backward<VT, float>(v, shapes, lstart, loc, aw, g, N, S, M, D, L, Lq, P, gv, gl, ga, zero_gv,       \
                                   stream, [&](cudaStream_t st) {                                                  \
      return msda::backward_f32acc<VT>(v, shapes, lstart, loc, aw, g, gv, gl, ga, N, S, M, D, L, Lq, P, st);       \
    });                                                                                                            \
  }

MSDA_DEFINE_F32ACC(f32, float, float)
MSDA_DEFINE_F32ACC(bf16, void, __nv_bfloat16)
MSDA_DEFINE_F32ACC(f16, void, __half)

int msda_backward_fusedq_16(const void* value, const int64_t* shapes, const int64_t* lstart, const float* loc, const float* aw,
                            const void* grad_out, const float* ref, int ref_dim, int N, int S, int M, int D, int L, int Lq, int P,
                            float* gv, void* dq, int zero_gv, int is_half, void* stream) {
  t_err[0] = 0;
  int rc = check_common(value, shapes, lstart, loc, aw, N, S, M, D, L, Lq, P, 2);
  if (rc) return rc;
  if (!grad_out || !gv || !dq || !ref) return fail(MSDA_ERR_NULL_POINTER, "null pointer");
  if (D != 32 || L != 4 || P != 4 || (ref_dim != 2 && ref_dim != 4))
    return fail(MSDA_ERR_UNSUPPORTED, "fused query backward needs D=32, L=4, P=4 (got D=%d L=%d P=%d)", D, L, P);
  if (!aligned16(grad_out) || !aligned16(gv) || !aligned16(dq) || !aligned16(ref))
    return fail(MSDA_ERR_MISALIGNED, "pointers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (zero_gv && N > 0) {
    rc = cuda_status(cudaMemsetAsync(gv, 0, sizeof(float) * static_cast<size_t>(N) * S * M * D, st), "grad_value memset");
    if (rc) return rc;
  }
  if (N == 0 || Lq == 0) return MSDA_OK;
  cudaError_t e;
  if (is_half)
    e = msda::backward_fused_q<__half>(static_cast<const __half*>(value), shapes, lstart, loc, aw, static_cast<const __half*>(grad_out),
                                       gv, ref, ref_dim, dq, 1, N, S, M, Lq, st);
  else
    e = msda::backward_fused_q<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(value), shapes, lstart, loc, aw,
                                              static_cast<const __nv_bfloat16*>(grad_out), gv, ref, ref_dim, dq, 0, N, S, M, Lq, st);
  return cuda_status(e, "msda_backward_fusedq launch");
}

long long msda_grad_value_h16_rows(const int64_t* shapes_host, int L, int Lq) {
  if (!shapes_host || L <= 0 || L > MSDA_MAX_LEVELS || Lq <= 0) return 0;
  long long rows = 0;
  for (int l = 0; l < L; ++l) {
    const long long hw = shapes_host[2 * l] * shapes_host[2 * l + 1];
    if (hw <= 0) return 0;
    rows += hw * msda::f16acc_replicas(Lq, hw);
  }
  return rows;
}

float msda_f16acc_scale(unsigned int amax_bits, int Lq) { return msda::f16acc_scale(amax_bits, Lq); }

int msda_backward_fusedq_h16(const void* value, const int64_t* shapes, const int64_t* lstart, const float* loc, const float* aw,
                             const void* grad_out, const float* ref, int ref_dim, int N, int S, int M, int D, int L, int Lq, int P,
                             void* gv_h, const int64_t* shapes_host, void* dq, int is_half, void* stream) {
  t_err[0] = 0;
  int rc = check_common(value, shapes, lstart, loc, aw, N, S, M, D, L, Lq, P, 2);
  if (rc) return rc;
  if (!grad_out || !gv_h || !dq || !ref || !shapes_host) return fail(MSDA_ERR_NULL_POINTER, "null pointer");
  if (D != 32 || L != 4 || P != 4 || (ref_dim != 2 && ref_dim != 4))
    return fail(MSDA_ERR_UNSUPPORTED, "fused query backward needs D=32, L=4, P=4 (got D=%d L=%d P=%d)", D, L, P);
  const long long rows_h = msda_grad_value_h16_rows(shapes_host, L, Lq);
  long long s_host = 0;
  for (int l = 0; l < L; ++l) s_host += shapes_host[2 * l] * shapes_host[2 * l + 1];
  if (rows_h <= 0 || s_host != S) return fail(MSDA_ERR_BAD_SHAPE, "host copy of spatial_shapes sums to %lld rows, S = %d", s_host, S);
  if (rows_h * M * D >= (1ll << 31)) return fail(MSDA_ERR_BAD_SHAPE, "scaled-fp16 map of %lld rows does not fit 32-bit per-image indexing", rows_h);
  if (!aligned16(grad_out) || !aligned16(gv_h) || !aligned16(dq) || !aligned16(ref))
    return fail(MSDA_ERR_MISALIGNED, "pointers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t map_bytes = 2 * static_cast<size_t>(N) * rows_h * M * D;
  rc = cuda_status(cudaMemsetAsync(gv_h, 0, map_bytes + 128, st), "grad_value memset");      // the map and the amax word behind it
  if (rc) return rc;
  if (N == 0 || Lq == 0) return MSDA_OK;
  uint32_t* amax = reinterpret_cast<uint32_t*>(static_cast<char*>(gv_h) + map_bytes);
  cudaError_t e;
  if (is_half)
    e = msda::backward_fused_q_h16<__half>(static_cast<const __half*>(value), shapes, lstart, loc, aw, static_cast<const __half*>(grad_out),
                                           gv_h, shapes_host, amax, ref, ref_dim, dq, 1, N, S, M, Lq, st);
  else
    e = msda::backward_fused_q_h16<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(value), shapes, lstart, loc, aw,
                                                  static_cast<const __nv_bfloat16*>(grad_out), gv_h, shapes_host, amax, ref, ref_dim, dq, 0, N, S,
                                                  M, Lq, st);
  return cuda_status(e, "msda_backward_fusedq_h16 launch");
}

long long msda_backward_workspace_bytes(int N, int M, int Lq) {
  if (N <= 0 || M <= 0 || Lq <= 0) return 0;
  return msda::backward_workspace_bytes(N, M, Lq);
}

int msda_backward_16_ws(const void* value, const int64_t* shapes, const int64_t* lstart, const float* loc, const float* aw,
                        const void* grad_out, const float* ref, int ref_dim, int N, int S, int M, int D, int L, int Lq, int P,
                        float* gv, float* gl, float* ga, void* dq, int zero_gv, int is_half, void* workspace,
                        long long workspace_bytes, void* stream) {
  msda::t_workspace = workspace;
  msda::t_workspace_bytes = workspace ? workspace_bytes : 0;
  int rc;
  if (dq != nullptr)
    rc = msda_backward_fusedq_16(value, shapes, lstart, loc, aw, grad_out, ref, ref_dim, N, S, M, D, L, Lq, P, gv, dq, zero_gv, is_half, stream);
  else if (is_half)
    rc = msda_backward_f16(value, shapes, lstart, loc, aw, grad_out, N, S, M, D, L, Lq, P, gv, gl, ga, zero_gv, stream);
  else
    rc = msda_backward_bf16(value, shapes, lstart, loc, aw, grad_out, N, S, M, D, L, Lq, P, gv, gl, ga, zero_gv, stream);
  msda::t_workspace = nullptr;
  msda::t_workspace_bytes = 0;
  return rc;
}

int msda_forward_f64(const double* value, const int64_t* shapes, const int64_t* lstart, const double* loc,
                     const double* aw, int N, int S, int M, int D, int L, int Lq, int P, double* out, void* stream) {
  return run_forward<double>(value, shapes, lstart, loc, aw, N, S, M, D, L, Lq, P, out, stream, [&](cudaStream_t st) {
    return msda::forward_f64(value, shapes, lstart, loc, aw, out, N, S, M, D, L, Lq, P, st);
  });
}

int msda_backward_f64(const double* value, const int64_t* shapes, const int64_t* lstart, const double* loc,
                      const double* aw, const double* grad_out, int N, int S, int M, int D, int L, int Lq, int P,
                      double* gv, double* gl, double* ga, int zero_gv, void* stream) {
  return run_backward<double, double>(value, shapes, lstart, loc, aw, grad_out, N, S, M, D, L, Lq, P, gv, gl, ga,
                                      zero_gv, stream, [&](cudaStream_t st) {
    return msda::backward_f64(value, shapes, lstart, loc, aw, grad_out, gv, gl, ga, N, S, M, D, L, Lq, P, st);
  });
}

}  // extern "C"
