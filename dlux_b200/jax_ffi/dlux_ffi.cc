// XLA-FFI handlers over the C ABI (include/dlux_b200.h).  Built ONLY where the XLA FFI
// headers exist (jax.ffi.include_dir()); JAX is not installable in the build image, so this
// file is untested there and is excluded from dlux_b200/build.py.  See INTEGRATION.md.
//
//   g++ -shared -fPIC -std=c++17 -I$(python -c "import jax; print(jax.ffi.include_dir())") \
//       -I include dlux_b200/jax_ffi/dlux_ffi.cc -L dlux_b200/lib -ldlux_b200 -o libdlux_b200_ffi.so
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
#endif
