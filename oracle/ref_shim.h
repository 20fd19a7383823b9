// TEST INFRASTRUCTURE ONLY -- force-included (-include) when oracle/build_ref.py compiles the
// UNMODIFIED reference sources where they lie under /root/reference.
//
// The reference calls AT_DISPATCH_FLOATING_TYPES(value.type(), ...) at
// csrc/MsDeformAttn/ms_deform_attn_cuda.cu:65 and :135.  torch >= 2.x removed the
// DeprecatedTypeProperties overload of ::detail::scalar_type that this relies on, which is the only
// reason the reference does not build on torch 2.11.  Re-adding that single overload here lets the
// reference compile without copying or patching a byte of it.
#pragma once
#ifdef __cplusplus
#include <ATen/ATen.h>
#include <ATen/Dispatch.h>
#include <ATen/core/DeprecatedTypeProperties.h>

namespace detail {
inline at::ScalarType scalar_type(const at::DeprecatedTypeProperties& t) {
  return t.scalarType();
}
}  // namespace detail
#endif
