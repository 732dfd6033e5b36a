/*
 * dlux_b200 -- C ABI of the B200-native dLux diffraction hot path.
 *
 * The reference (LouisDesdoigts/dLux v0.16.0) is pure Python/JAX and has no FFI of
 * its own; each entry point below states the reference lines it replaces (paths
 * relative to /root/reference/).  This is the surface an XLA-FFI custom call
 * (jax.ffi) or a ctypes binder binds; see INTEGRATION.md.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host; arrays are
 *    row-major; complex64 is interleaved (re, im) float32 pairs;
 *  - the caller owns all buffers including scratch; the library never allocates or
 *    frees device memory and fully overwrites its outputs;
 *  - work is enqueued on the given CUDA stream only (no device synchronisation, no
 *    host reads of device scalars: wavelengths / scales / shifts are device operands
 *    because under JAX they are tracers);
 *  - return value 0 on success, a negative DLUX_ERR_* otherwise; never aborts;
 *  - re-entrant; the only global state is an immutable-after-init kernel attribute
 *    cache.
 */
#ifndef DLUX_B200_H
#define DLUX_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define DLUX_API __attribute__((visibility("default")))
#else
#define DLUX_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define DLUX_B200_ABI_VERSION 2

enum {
  DLUX_OK = 0,
  DLUX_ERR_ARG = -1,          /* null pointer / bad enum */
  DLUX_ERR_SHAPE = -2,        /* unsupported size */
  DLUX_ERR_SCRATCH = -3,      /* scratch buffer too small */
  DLUX_ERR_CUDA = -4,         /* a CUDA call failed (see dlux_last_cuda_error) */
  DLUX_ERR_UNSUPPORTED = -5,  /* e.g. tcgen05 path on a non-sm_100 device */
  DLUX_ERR_ALIGN = -6         /* pointer not 16-byte aligned */
};

/* Arithmetic of the two dense contractions. */
enum {
  DLUX_PREC_3XTF32 = 0, /* default: tcgen05.mma kind::tf32, hi/lo split operands, fp32 TMEM accumulators */
  DLUX_PREC_FP32 = 1    /* CUDA-core FFMA; validation path for the tensor kernel */
};

DLUX_API int dlux_abi_version(void);
DLUX_API const char* dlux_error_string(int code);
/* cudaError_t of the last failing CUDA call on this thread (0 if none). */
DLUX_API int dlux_last_cuda_error(void);
/* Number of kernels this library has launched since load (process-wide counter). */
DLUX_API uint64_t dlux_launch_count(void);

/* Optional per-kernel timing of the phasor-GEMM launches (the dominant kernel) with CUDA
 * events recorded on the launching stream; used by bench.py for the roofline line.
 * dlux_profile_read synchronises on the recorded events, returns the summed kernel
 * time [ms], the number of GEMM launches and their algorithmic FLOPs (8 per complex
 * MAC), and clears the record. */
DLUX_API int dlux_profile_enable(int on);
DLUX_API int dlux_profile_read(double* gemm_ms, uint64_t* gemm_launches, double* gemm_flops);

/* Measurement aid (no reference counterpart): an MMA-only tcgen05 microbenchmark of the instruction
 * shapes the phasor GEMM issues (cta_group::2, M256 x N128, A in tensor memory, B in shared memory,
 * operands resident).  kind 0: kind::tf32 (K=8), 1: kind::f16 bf16 (K=16), 2: the GEMM's own mix.
 * Each leader CTA issues n_batches x 32 MMAs; *flops_host receives the real FLOPs of the launch.
 * bench.py times it with CUDA events to measure the tensor-pipe peak the roofline is quoted against. */
DLUX_API int dlux_tc_peak_probe(int32_t kind, int32_t n_batches, float* sink /* 1 float, or NULL */,
                       double* flops_host, void* cuda_stream);

/* ------------------------------------------------------------------------------
 * dlux_mft_c64: batched matrix Fourier transform.
 * Replaces dLux.utils.propagation.MFT (src/dLux/utils/propagation.py:178-256)
 * = transfer_matrix (:67-127) x2, (y_mat.T @ phasor) @ x_mat (:243), normalise (:254).
 *
 *   out[b] = norm[b] * A_y(b)^T . in[b] . A_x(b)
 *   A_x(b)[j,c] = exp(-/+ 2 pi i x_j u_c),  x = nd_coords(n_in, 1/n_in, shift_x/n_in),
 *                                          u = nd_coords(n_out, s_b, shift_x*s_b) - delta_x
 * with s_b = scale_out[b] (= pixel_scale_out / fringe_size [/ focal_length], :110-120),
 * norm[b] = exp(log nfringes - log n_in - log n_out) (:246-254), both formed by the
 * caller in float32 exactly as the reference does.  adjoint != 0 applies the
 * conjugate-transpose operator (n_out x n_out -> n_in x n_in), which is what the
 * custom_vjp / transpose rule of the JAX primitive calls.
 * -------------------------------------------------------------------------- */
typedef struct {
  int32_t n_in;      /* pupil side N (phasor.shape[-1]) */
  int32_t n_out;     /* npixels_out M */
  int32_t batch;     /* leading batch (vmap) size */
  int32_t inverse;   /* propagation.py:125-126: sign flip of the exponent */
  int32_t adjoint;   /* 0: forward operator, 1: its conjugate transpose */
  int32_t precision; /* DLUX_PREC_* */
  int32_t dft_period; /* 0: MFT.  P > 0: exact DFT of period P (dlu.FFT, propagation.py:8-64, evaluated as a
                         centred, zero-padded DFT): phasors exp(-/+ 2 pi i ((j - j0)(b - b0) mod P) / P) with the
                         integer origins j0 = shift_xy, b0 = delta_xy (scale_out is ignored) */
  int32_t reserved;
} dlux_mft_desc;

DLUX_API size_t dlux_mft_scratch_bytes(const dlux_mft_desc* desc);

DLUX_API int dlux_mft_c64(const dlux_mft_desc* desc,
                 const void* in,          /* c64 [batch, n, n], n = adjoint ? n_out : n_in */
                 const float* scale_out,  /* [batch] */
                 const float* shift_xy,   /* [batch, 2] pixels (x, y); may be NULL = 0 */
                 const float* delta_xy,   /* [batch, 2] extra output-coordinate offset in fringes; may be NULL */
                 const float* norm,       /* [batch]; may be NULL = 1 */
                 void* out,               /* c64 [batch, m, m], m = adjoint ? n_in : n_out */
                 void* scratch, size_t scratch_bytes, void* cuda_stream);

/* The two coordinate vectors of transfer_matrix (propagation.py:113-121 through
 * utils/coordinates.py:329-332, jnp.linspace lerp form), bit-exact in float32.
 * xin: [batch, 2, n_in], uout: [batch, 2, n_out] (axis 0 = x, 1 = y). */
DLUX_API int dlux_mft_coords(int32_t n_in, int32_t n_out, int32_t batch, const float* scale_out,
                    const float* shift_xy, const float* delta_xy, float* xin, float* uout,
                    void* cuda_stream);

/* ------------------------------------------------------------------------------
 * Fused polychromatic PSF.  Replaces, for a pupil-only layer stack followed by an
 * MFT to focus, OpticalSystem.propagate (src/dLux/optical_systems.py:147-223) under
 * PointSource(s).model (src/dLux/sources.py:316-327, 392-411):
 *   P_l   = amp * T * exp(i (k_l * opd + phase))            (wavefronts.py:349,368;
 *                                                            layers/optics.py:91-96)
 *   E_sl  = norm_l * A_y^T P_l A_x    with the source tilt (wavefronts.py:370-395)
 *           folded into the output coordinates: delta = theta * D / lambda fringes
 *   psf   = sum_{s,l} w[s,l] |E_sl|^2                       (optical_systems.py:213-223,
 *                                                            sources.py:409-411)
 * amp = 1/N^2 (wavefronts.py:111) times, if normalise, 1/sqrt(sum |T/N^2|^2)
 * (wavefronts.py:418-424), computed on the device.
 * -------------------------------------------------------------------------- */
typedef struct {
  int32_t n_pupil;    /* N = wf_npixels */
  int32_t n_psf;      /* M = psf_npixels * oversample */
  int32_t n_wavels;   /* L */
  int32_t n_sources;  /* S ([S, L] operands are source-major: index s * L + l; work is ordered wavelength-major) */
  int32_t normalise;  /* Optic(normalise=True) */
  int32_t precision;  /* DLUX_PREC_* */
  int32_t save_field; /* fwd: also write E [S*L, M, M] c64 (the VJP residual) */
  int32_t sparse;     /* 1: skip blocks where transmission == 0 (exact; assumes finite opd / phase there) */
} dlux_polypsf_desc;

DLUX_API size_t dlux_polypsf_scratch_bytes(const dlux_polypsf_desc* desc);

DLUX_API int dlux_polypsf_fwd(const dlux_polypsf_desc* desc,
                     const float* transmission, /* [N, N] or NULL (=1) */
                     const float* opd,          /* [N, N] metres, or NULL */
                     const float* phase,        /* [N, N] radians, or NULL */
                     const float* wavenumber,   /* [L]  2 pi / lambda */
                     const float* scale_out,    /* [L] */
                     const float* norm,         /* [L] */
                     const float* weights,      /* [S, L] flux * spectral weight */
                     const float* delta_xy,     /* [S, L, 2] fringes, or NULL */
                     float* psf,                /* [M, M] (overwritten) */
                     void* field,               /* c64 [S*L, M, M] if save_field else NULL: opaque VJP residual
                                                   (stored in the library's processing order) */
                     void* scratch, size_t scratch_bytes, void* cuda_stream);

/* VJP of dlux_polypsf_fwd w.r.t. opd (-> Zernike coefficients through
 * dlux_basis_reduce), phase, transmission, weights (-> flux, spectrum) and the source offsets delta_xy
 * (-> PointSources.position through delta = theta * D / lambda), given psf_bar = dL/dpsf and
 * the saved field.  Stands in for jax.grad through the same lines
 * (docs/phase_retrieval.md:269-287). */
DLUX_API int dlux_polypsf_bwd(const dlux_polypsf_desc* desc,
                     const float* transmission, const float* opd, const float* phase,
                     const float* wavenumber, const float* scale_out, const float* norm,
                     const float* weights, const float* delta_xy,
                     const void* field,        /* c64 [S*L, M, M] from fwd */
                     const float* psf_bar,     /* [M, M] */
                     float* opd_bar,           /* [N, N] or NULL */
                     float* phase_bar,         /* [N, N] or NULL */
                     float* weights_bar,       /* [S, L] or NULL */
                     float* delta_bar,         /* [S, L, 2] or NULL */
                     float* transmission_bar,  /* [N, N] or NULL (incl. the power-normalisation term) */
                     float* scale_bar,         /* [S, L] or NULL: d/d scale_out (two extra adjoint MFTs) */
                     float* wavenumber_bar,    /* [S, L] or NULL: d/d wavenumber through exp(i k opd) only */
                     void* scratch, size_t scratch_bytes, void* cuda_stream);

/* Second order of the fused PSF w.r.t. the OPD, forward-over-reverse (what jax.hessian / zdx.hessian of a
 * loss through OpticalSystem.propagate needs, docs/mask_design.md:480-488).  With L(opd) = <psf_bar, psf(opd)>,
 * opd_bar(opd, psf_bar) its gradient (dlux_polypsf_bwd) and V a pupil-plane direction [N, N]:
 *   psf_tan = (d psf / d opd) V                  = sum_{s,l} 2 w Re(conj(E) dE),   dE = MFT(i k V P)
 *   opd_hv  = d/d opd <V, opd_bar(opd, psf_bar)> = sum_{s,l} k Im(conj(P) A^H(2 w psf_bar dE))
 *                                                            - k^2 V Re(conj(P) A^H(2 w psf_bar E))
 * i.e. the two cotangents of the map (opd, psf_bar) -> opd_bar: one more forward MFT and two adjoint MFTs
 * per (source, wavelength), on the same kernels. */
DLUX_API int dlux_polypsf_hvp(const dlux_polypsf_desc* desc,
                     const float* transmission, const float* opd, const float* phase,
                     const float* wavenumber, const float* scale_out, const float* norm,
                     const float* weights, const float* delta_xy,
                     const void* field,          /* c64 [S*L, M, M] from fwd */
                     const float* psf_bar,       /* [M, M] */
                     const float* opd_tangent,   /* V [N, N] */
                     float* psf_tan,             /* [M, M] or NULL */
                     float* opd_hv,              /* [N, N] or NULL */
                     void* scratch, size_t scratch_bytes, void* cuda_stream);

/* ------------------------------------------------------------------------------
 * Parameter-batched fused PSF: B coefficient vectors of one OPD basis, each giving its own
 * polychromatic PSF (and, backward, its own coefficient gradient).  This is what the reference gets
 * from vmapping OpticalSystem.propagate / jax.grad over a batch of parameter sets
 * (docs/mask_design.md:454-488, optical_systems.py:213-216; BASELINE config 5):
 *   opd_b  = base_opd + sum_z coeffs[b, z] * basis[z]          (utils/math.py:177-196, fused in)
 *   psf[b] = sum_l w[l] |norm_l A_y^T (amp T exp(i (k_l opd_b + phase))) A_x|^2
 * Outputs are per batch element: nothing is summed over b, so a batch sharded over GPUs needs no
 * collective.  Items are batch-major (item = b * L + l) and are processed in chunks of whole batch
 * elements inside the caller's scratch buffer.
 * -------------------------------------------------------------------------- */
typedef struct {
  int32_t n_pupil;    /* N */
  int32_t n_psf;      /* M */
  int32_t n_wavels;   /* L */
  int32_t n_batch;    /* B parameter sets */
  int32_t n_basis;    /* nz */
  int32_t normalise;
  int32_t precision;  /* DLUX_PREC_* */
  int32_t save_field; /* fwd: also write E [B*L, M, M] c64 (the VJP residual) */
} dlux_polypsf_batch_desc;

DLUX_API size_t dlux_polypsf_batch_scratch_bytes(const dlux_polypsf_batch_desc* desc);

DLUX_API int dlux_polypsf_batch_fwd(const dlux_polypsf_batch_desc* desc,
                     const float* transmission, /* [N, N] or NULL */
                     const float* base_opd,     /* [N, N] or NULL: OPD of the other layers */
                     const float* phase,        /* [N, N] or NULL */
                     const float* basis,        /* [nz, N, N] */
                     const float* coeffs,       /* [B, nz] */
                     const float* wavenumber,   /* [L] */
                     const float* scale_out,    /* [L] */
                     const float* norm,         /* [L] */
                     const float* weights,      /* [L] */
                     const float* delta_xy,     /* [L, 2] or NULL */
                     float* psf,                /* [B, M, M] */
                     void* field,               /* c64 [B*L, M, M] if save_field else NULL */
                     void* scratch, size_t scratch_bytes, void* cuda_stream);

DLUX_API int dlux_polypsf_batch_bwd(const dlux_polypsf_batch_desc* desc,
                     const float* transmission, const float* base_opd, const float* phase,
                     const float* basis, const float* coeffs, const float* wavenumber,
                     const float* scale_out, const float* norm, const float* weights, const float* delta_xy,
                     const void* field,         /* c64 [B*L, M, M] from fwd */
                     const float* psf_bar,      /* [B, M, M] */
                     float* coeff_bar,          /* [B, nz] (overwritten) */
                     void* scratch, size_t scratch_bytes, void* cuda_stream);

/* ------------------------------------------------------------------------------
 * dlu.eval_basis (src/dLux/utils/math.py:177-196): out = base + sum_z c_z basis_z,
 * and its transpose (the coefficient gradient) coeff_bar_z = <basis_z, out_bar>.
 * -------------------------------------------------------------------------- */
DLUX_API int dlux_basis_eval(int32_t n_basis, int64_t n_pix, const float* basis /*[nz, n_pix]*/,
                    const float* coeffs /*[nz]*/, const float* base /*[n_pix] or NULL*/,
                    float* out /*[n_pix]*/, void* cuda_stream);
DLUX_API int dlux_basis_reduce(int32_t n_basis, int64_t n_pix, const float* basis,
                      const float* out_bar /*[n_pix]*/, float* coeff_bar /*[nz], overwritten*/,
                      void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* DLUX_B200_H */
