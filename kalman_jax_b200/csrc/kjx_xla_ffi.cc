// kjx_xla_ffi.cc — thin XLA-FFI custom-call handlers over the C ABI of include/kjx.h.
//
// BASELINE.json's north_star asks for the host code to stay Python/JAX and reach CUDA "through a thin XLA-FFI
// custom call".  Neither JAX nor the XLA FFI headers exist in this image (SURVEY.md TL;DR 3), so this translation
// unit is not part of libkjx.so; it is the exact shim a JAX build would register, kept compiling by the stub test.
// With jaxlib's headers on the include path:
//
//     g++ -shared -fPIC -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") kjx_xla_ffi.cc -L.. -lkjx
//     jax.ffi.register_ffi_target("kjx_run_f64", jax.ffi.pycapsule(lib.KjxRunF64), platform="CUDA")
//
// The model description travels as attributes: kjx_model_t (a plain-old-data struct) as one opaque byte string with its
// two pointer members NULLED, and the small matrices they point to (Pinf [d,d], H [1,d]) as f64 array attributes that
// the handler re-attaches — so nothing process-local is ever baked into the compiled XLA module.
// tests/test_xla_ffi_stub.py compiles this file against a minimal stand-in for the header (tests/xla_stub/), which also
// type-checks each handler against its binding.
#if __has_include("xla/ffi/api/ffi.h")
#include <cstring>
#include <string>
#include <string_view>

#include <cuda_runtime_api.h>

#include "../../include/kjx.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error status_to_error(int rc, const char* what) {
  if (rc == KJX_OK) return ffi::Error::Success();
  return ffi::Error(rc == KJX_ERR_INVALID_ARG ? ffi::ErrorCode::kInvalidArgument
                    : rc == KJX_ERR_UNSUPPORTED ? ffi::ErrorCode::kUnimplemented
                                                : ffi::ErrorCode::kInternal,
                    std::string(what) + ": " + kjx_status_string(rc));
}

// kjx_model_t from its attribute form (see the header comment)
static bool model_from_attrs(std::string_view bytes, ffi::Span<const double> pinf, ffi::Span<const double> h, kjx_model_t* m) {
  if (bytes.size() != sizeof(kjx_model_t)) return false;
  std::memcpy(m, bytes.data(), sizeof(*m));
  if (m->state_dim <= 0 || pinf.size() != static_cast<size_t>(m->state_dim) * m->state_dim) return false;
  if (h.size() != static_cast<size_t>(m->func_dim) * m->state_dim) return false;
  m->Pinf = pinf.data(); m->H = h.data(); m->H_steps = nullptr;
  return true;
}

// One SDEGP.run() iteration (sde_gp.py:179-187 minus the gradient): see kjx_run_f64.
static ffi::Error KjxRunF64Impl(cudaStream_t stream, std::string_view model_bytes, ffi::Span<const double> pinf,
                                ffi::Span<const double> h, ffi::Buffer<ffi::F64> dt,
                                ffi::Buffer<ffi::F64> y, ffi::Buffer<ffi::F64> site_mean, ffi::Buffer<ffi::F64> site_cov,
                                ffi::Buffer<ffi::U8> workspace, ffi::ResultBuffer<ffi::F64> new_site_mean,
                                ffi::ResultBuffer<ffi::F64> new_site_cov, ffi::ResultBuffer<ffi::F64> post_mean,
                                ffi::ResultBuffer<ffi::F64> post_var, ffi::ResultBuffer<ffi::F64> nlml) {
  kjx_model_t m;
  if (!model_from_attrs(model_bytes, pinf, h, &m)) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "bad kjx_model_t attributes");
  const int64_t N = static_cast<int64_t>(dt.element_count());
  const int rc = kjx_run_f64(&m, N, dt.typed_data(), y.typed_data(), site_mean.typed_data(), site_cov.typed_data(),
                             /*filt_mean=*/nullptr, /*filt_cov=*/nullptr, new_site_mean->typed_data(),
                             new_site_cov->typed_data(), post_mean->typed_data(), post_var->typed_data(),
                             nlml->typed_data(), workspace.untyped_data(), workspace.size_bytes(), 0,
                             reinterpret_cast<kjx_stream_t>(stream));
  return status_to_error(rc, "kjx_run_f64");
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(KjxRunF64, KjxRunF64Impl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<std::string_view>("model")
                                  .Attr<ffi::Span<const double>>("pinf")
                                  .Attr<ffi::Span<const double>>("h")
                                  .Arg<ffi::Buffer<ffi::F64>>()   // dt        [N]
                                  .Arg<ffi::Buffer<ffi::F64>>()   // y         [N, 1]
                                  .Arg<ffi::Buffer<ffi::F64>>()   // site_mean [N, 1, 1]
                                  .Arg<ffi::Buffer<ffi::F64>>()   // site_cov  [N, 1, 1]
                                  .Arg<ffi::Buffer<ffi::U8>>()    // workspace (kjx_workspace_bytes)
                                  .Ret<ffi::Buffer<ffi::F64>>()   // new site_mean
                                  .Ret<ffi::Buffer<ffi::F64>>()   // new site_cov
                                  .Ret<ffi::Buffer<ffi::F64>>()   // posterior marginal mean
                                  .Ret<ffi::Buffer<ffi::F64>>()   // posterior marginal variance
                                  .Ret<ffi::Buffer<ffi::F64>>()); // nlml [1]

// Per-step site update / NLPD term (approximate_inference.py:36, sde_gp.py:139-142): see kjx_site_update_f64.
static ffi::Error KjxSiteUpdateF64Impl(cudaStream_t stream, std::string_view model_bytes, ffi::Span<const double> pinf,
                                       ffi::Span<const double> h, ffi::Buffer<ffi::F64> y,
                                       ffi::Buffer<ffi::F64> post_mean, ffi::Buffer<ffi::F64> post_var,
                                       ffi::Buffer<ffi::F64> site_mean_prev, ffi::Buffer<ffi::F64> site_var_prev,
                                       ffi::ResultBuffer<ffi::F64> log_lik, ffi::ResultBuffer<ffi::F64> site_mean,
                                       ffi::ResultBuffer<ffi::F64> site_var) {
  kjx_model_t m;
  if (!model_from_attrs(model_bytes, pinf, h, &m)) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "bad kjx_model_t attributes");
  const int rc = kjx_site_update_f64(&m, static_cast<int64_t>(y.element_count()), y.typed_data(), post_mean.typed_data(),
                                     post_var.typed_data(), site_mean_prev.typed_data(), site_var_prev.typed_data(),
                                     log_lik->typed_data(), site_mean->typed_data(), site_var->typed_data(),
                                     reinterpret_cast<kjx_stream_t>(stream));
  return status_to_error(rc, "kjx_site_update_f64");
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(KjxSiteUpdateF64, KjxSiteUpdateF64Impl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<std::string_view>("model")
                                  .Attr<ffi::Span<const double>>("pinf")
                                  .Attr<ffi::Span<const double>>("h")
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>());
#endif  // __has_include("xla/ffi/api/ffi.h")
