// XLA-FFI handlers over the C ABI (include/dlux_b200.h).  Built ONLY where the XLA FFI
// headers exist (jax.ffi.include_dir()); JAX is not installable in the build image, so this
// file is untested there and is excluded from dlux_b200/build.py.  See INTEGRATION.md.
//
//   g++ -shared -fPIC -std=c++17 -I$(python -c "import jax; print(jax.ffi.include_dir())") \
//       -I include integration/jax_ffi/dlux_ffi.cc -L dlux_b200/lib -ldlux_b200 -o libdlux_b200_ffi.so
#if __has_include("xla/ffi/api/ffi.h")
#include <cuda_runtime_api.h>
#include "xla/ffi/api/ffi.h"
#include "dlux_b200.h"

namespace ffi = xla::ffi;

static ffi::Error to_error(int rc, const char* what) {
  if (rc == DLUX_OK) return ffi::Error::Success();
  return ffi::Error(ffi::ErrorCode::kInternal, std::string(what) + ": " + dlux_error_string(rc));
}

// dlu.MFT (src/dLux/utils/propagation.py:178-256) and its conjugate transpose.
static ffi::Error MftImpl(cudaStream_t stream, ffi::Buffer<ffi::C64> in, ffi::Buffer<ffi::F32> scale_out,
                          ffi::Buffer<ffi::F32> shift_xy, ffi::Buffer<ffi::F32> norm,
                          ffi::ResultBuffer<ffi::C64> out, ffi::ResultBuffer<ffi::U8> scratch,
                          int32_t n_in, int32_t n_out, int32_t inverse, int32_t adjoint, int32_t precision) {
  auto dims = in.dimensions();
  int64_t batch = 1;
  for (size_t i = 0; i + 2 < dims.size(); ++i) batch *= dims[i];
  dlux_mft_desc d{n_in, n_out, (int32_t)batch, inverse, adjoint, precision};
  return to_error(dlux_mft_c64(&d, in.untyped_data(), scale_out.typed_data(), shift_xy.typed_data(), nullptr,
                               norm.typed_data(), out->untyped_data(), scratch->untyped_data(),
                               scratch->element_count(), stream), "dlux_mft_c64");
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(dlux_mft_ffi, MftImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::C64>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::C64>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int32_t>("n_in")
                                  .Attr<int32_t>("n_out")
                                  .Attr<int32_t>("inverse")
                                  .Attr<int32_t>("adjoint")
                                  .Attr<int32_t>("precision"),
                              {ffi::Traits::kCmdBufferCompatible});

// A zero-element buffer stands for an absent optional operand / result (NULL in the C ABI).
template <typename B>
static auto* opt(B& b) { return b.element_count() == 0 ? nullptr : b.typed_data(); }
template <typename B>
static auto* opt_res(B& b) { return b->element_count() == 0 ? nullptr : b->typed_data(); }

// OpticalSystem.propagate over a pupil-only stack under PointSource(s).model
// (optical_systems.py:147-223, sources.py:316-327, 392-411): the fused forward.
static ffi::Error PolyFwdImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> transmission, ffi::Buffer<ffi::F32> opd,
                              ffi::Buffer<ffi::F32> phase, ffi::Buffer<ffi::F32> wavenumber,
                              ffi::Buffer<ffi::F32> scale_out, ffi::Buffer<ffi::F32> norm,
                              ffi::Buffer<ffi::F32> weights, ffi::Buffer<ffi::F32> delta_xy,
                              ffi::ResultBuffer<ffi::F32> psf, ffi::ResultBuffer<ffi::C64> field,
                              ffi::ResultBuffer<ffi::U8> scratch, int32_t n_pupil, int32_t n_psf,
                              int32_t normalise, int32_t precision) {
  const int32_t L = (int32_t)wavenumber.element_count();
  const int32_t S = (int32_t)(weights.element_count() / (L > 0 ? L : 1));
  const int32_t save = field->element_count() != 0;
  dlux_polypsf_desc d{n_pupil, n_psf, L, S, normalise, precision, save, 0};
  return to_error(dlux_polypsf_fwd(&d, opt(transmission), opt(opd), opt(phase), wavenumber.typed_data(),
                                   scale_out.typed_data(), norm.typed_data(), weights.typed_data(),
                                   opt(delta_xy), psf->typed_data(), save ? field->untyped_data() : nullptr,
                                   scratch->untyped_data(), scratch->element_count(), stream),
                  "dlux_polypsf_fwd");
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(dlux_polypsf_fwd_ffi, PolyFwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()   // transmission (or empty)
                                  .Arg<ffi::Buffer<ffi::F32>>()   // opd (or empty)
                                  .Arg<ffi::Buffer<ffi::F32>>()   // phase (or empty)
                                  .Arg<ffi::Buffer<ffi::F32>>()   // wavenumber [L]
                                  .Arg<ffi::Buffer<ffi::F32>>()   // scale_out [L]
                                  .Arg<ffi::Buffer<ffi::F32>>()   // norm [L]
                                  .Arg<ffi::Buffer<ffi::F32>>()   // weights [S, L]
                                  .Arg<ffi::Buffer<ffi::F32>>()   // delta_xy [S, L, 2] (or empty)
                                  .Ret<ffi::Buffer<ffi::F32>>()   // psf [M, M]
                                  .Ret<ffi::Buffer<ffi::C64>>()   // field [S*L, M, M] (or empty)
                                  .Ret<ffi::Buffer<ffi::U8>>()    // scratch
                                  .Attr<int32_t>("n_pupil")
                                  .Attr<int32_t>("n_psf")
                                  .Attr<int32_t>("normalise")
                                  .Attr<int32_t>("precision"),
                              {ffi::Traits::kCmdBufferCompatible});

// jax.grad through the same lines: every cotangent is optional (empty result = not computed).
static ffi::Error PolyBwdImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> transmission, ffi::Buffer<ffi::F32> opd,
                              ffi::Buffer<ffi::F32> phase, ffi::Buffer<ffi::F32> wavenumber,
                              ffi::Buffer<ffi::F32> scale_out, ffi::Buffer<ffi::F32> norm,
                              ffi::Buffer<ffi::F32> weights, ffi::Buffer<ffi::F32> delta_xy,
                              ffi::Buffer<ffi::C64> field, ffi::Buffer<ffi::F32> psf_bar,
                              ffi::ResultBuffer<ffi::F32> opd_bar, ffi::ResultBuffer<ffi::F32> phase_bar,
                              ffi::ResultBuffer<ffi::F32> weights_bar, ffi::ResultBuffer<ffi::F32> delta_bar,
                              ffi::ResultBuffer<ffi::F32> transmission_bar, ffi::ResultBuffer<ffi::F32> scale_bar,
                              ffi::ResultBuffer<ffi::F32> wavenumber_bar, ffi::ResultBuffer<ffi::U8> scratch,
                              int32_t n_pupil, int32_t n_psf, int32_t normalise, int32_t precision) {
  const int32_t L = (int32_t)wavenumber.element_count();
  const int32_t S = (int32_t)(weights.element_count() / (L > 0 ? L : 1));
  dlux_polypsf_desc d{n_pupil, n_psf, L, S, normalise, precision, 1, 0};
  return to_error(dlux_polypsf_bwd(&d, opt(transmission), opt(opd), opt(phase), wavenumber.typed_data(),
                                   scale_out.typed_data(), norm.typed_data(), weights.typed_data(),
                                   opt(delta_xy), field.untyped_data(), psf_bar.typed_data(), opt_res(opd_bar),
                                   opt_res(phase_bar), opt_res(weights_bar), opt_res(delta_bar),
                                   opt_res(transmission_bar), opt_res(scale_bar), opt_res(wavenumber_bar),
                                   scratch->untyped_data(), scratch->element_count(), stream),
                  "dlux_polypsf_bwd");
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(dlux_polypsf_bwd_ffi, PolyBwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::C64>>()   // field saved by the forward
                                  .Arg<ffi::Buffer<ffi::F32>>()   // psf_bar [M, M]
                                  .Ret<ffi::Buffer<ffi::F32>>()   // opd_bar
                                  .Ret<ffi::Buffer<ffi::F32>>()   // phase_bar
                                  .Ret<ffi::Buffer<ffi::F32>>()   // weights_bar
                                  .Ret<ffi::Buffer<ffi::F32>>()   // delta_bar
                                  .Ret<ffi::Buffer<ffi::F32>>()   // transmission_bar
                                  .Ret<ffi::Buffer<ffi::F32>>()   // scale_bar
                                  .Ret<ffi::Buffer<ffi::F32>>()   // wavenumber_bar
                                  .Ret<ffi::Buffer<ffi::U8>>()    // scratch
                                  .Attr<int32_t>("n_pupil")
                                  .Attr<int32_t>("n_psf")
                                  .Attr<int32_t>("normalise")
                                  .Attr<int32_t>("precision"),
                              {ffi::Traits::kCmdBufferCompatible});
#endif
