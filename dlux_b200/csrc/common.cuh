// Shared declarations of the dlux_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/dlux_b200.h"

namespace dlux {

// ---------------------------------------------------------------------------
// One "phasor GEMM" stage.  Every stage of the forward and adjoint MFT is
//
//   Out[item][n][m] = scale[item] * sum_k Data[d(item)][m][k] * exp(i * sign2pi * fl(kvec[k] * nvec[n]))
//
// i.e. a complex contraction of a data matrix (planar, hi/lo-split fp32 planes)
// with a DFT phasor matrix that is generated on the fly from two float32
// coordinate vectors, written TRANSPOSED (m fastest) so that two stages chain.
// The phase argument is formed exactly as the reference does
// (/root/reference/src/dLux/utils/propagation.py:124): fl32(fl32(-2pi) * fl32(x*u)).
// ---------------------------------------------------------------------------
enum EpilogueMode : int {
  EPI_PLANES = 0,  // out_planes[p][item][n][m], p = re_hi, re_lo, im_hi, im_lo  (feeds the next stage)
  EPI_C64 = 1      // out_c64[item][n][m]
};

struct GemmParams {
  // data operand: 4 planes, each [n_data][rows][K] float32
  const float* a_planes[4];
  int a_pitch;  // floats per data row (>= K, multiple of 4 so TMA can stride it)
  int n_data;   // number of data matrices in a_planes
  int rows;    // M dimension of the data matrix (rows m)
  int K;       // contraction length
  int n_out;   // number of generated output coordinates (n)
  int n_items;
  const int* item_data;  // [n_items] -> data matrix index, or nullptr (identity)
  const float* kvec;     // [n_items][kvec_stride], coordinate along k
  const float* nvec;     // [n_items][nvec_stride], coordinate along n
  int kvec_stride, nvec_stride;
  float sign2pi;         // float32(-2*pi) forward, float32(+2*pi) inverse / adjoint-of-forward
  const float* scale;    // [n_items] or nullptr
  int mode;
  float* out_planes[4];  // EPI_PLANES: [n_items][n_out][out_pitch]
  int out_pitch;
  float2* out_c64;       // EPI_C64: [n_items][n_out][rows]
};

__device__ __forceinline__ float tf32_hi(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// Reference phase argument: two float32 multiplies, no fused contraction.
__device__ __forceinline__ float phase_arg(float sign2pi, float x, float u) {
  return __fmul_rn(sign2pi, __fmul_rn(x, u));
}

// sin/cos of a float32 argument, ~1.5 ulp: Cody-Waite reduction by pi/2 in three parts
// followed by minimax polynomials on [-pi/4, pi/4] (the classic single-precision
// sincosf fast path).  Valid for |a| < 1e5; larger arguments (never reached by optical
// MFT geometries, where |a| ~ pi * nfringes / 2) take the libdevice slow path.
__device__ __forceinline__ void fast_sincos(float a, float* sn, float* cs) {
  if (fabsf(a) > 1.0e5f) {
    sincosf(a, sn, cs);
    return;
  }
  float j = fmaf(a, 0.636619747f, 12582912.0f);
  const int q = __float_as_int(j);
  j -= 12582912.0f;
  float r = fmaf(j, -1.57079601e+00f, a);
  r = fmaf(j, -3.13916473e-07f, r);
  r = fmaf(j, -5.39030253e-15f, r);
  const float s2 = r * r;
  float c = 2.44677067e-5f;
  c = fmaf(c, s2, -1.38877297e-3f);
  c = fmaf(c, s2, 4.16666567e-2f);
  c = fmaf(c, s2, -5.00000000e-1f);
  c = fmaf(c, s2, 1.0f);
  float s = 2.86567956e-6f;
  s = fmaf(s, s2, -1.98559923e-4f);
  s = fmaf(s, s2, 8.33338592e-3f);
  s = fmaf(s, s2, -1.66666672e-1f);
  const float t = r * s2;
  s = fmaf(s, t, r);
  const float so = (q & 1) ? c : s;
  const float co = (q & 1) ? s : c;
  *sn = (q & 2) ? -so : so;
  *cs = ((q + 1) & 2) ? -co : co;
}

// Shared epilogue for one output element D[m][n] = (re, im) of `item`.
__device__ __forceinline__ void epilogue_store(const GemmParams& p, int item, int m, int n,
                                               float re, float im) {
  const float sc = p.scale ? __ldg(p.scale + item) : 1.0f;
  re *= sc;
  im *= sc;
  const size_t idx = ((size_t)item * p.n_out + n) * p.rows + m;
  if (p.mode == EPI_PLANES) {
    const size_t po = ((size_t)item * p.n_out + n) * p.out_pitch + m;
    const float rh = tf32_hi(re), ih = tf32_hi(im);
    p.out_planes[0][po] = rh;
    p.out_planes[1][po] = re - rh;
    p.out_planes[2][po] = ih;
    p.out_planes[3][po] = im - ih;
  } else {  // EPI_C64
    p.out_c64[idx] = make_float2(re, im);
  }
}

// launch counters / error plumbing (api.cu)
void note_launch(int n = 1);
int check_launch(const char* what);

// kernels' host launchers
int launch_gemm_simt(const GemmParams& p, cudaStream_t st);
int launch_gemm_tc(const GemmParams& p, cudaStream_t st);
size_t gemm_tc_workspace_bytes();

int launch_coords(int n_in, int n_out, int batch, const float* scale_out, const float* shift_xy,
                  const float* delta_xy, int delta_stride_items, float* xin, float* uout,
                  cudaStream_t st);
__host__ __device__ inline int pitch4(int k) { return (k + 3) & ~3; }
// in: [n_mat][rows][cols] c64 -> planes [n_mat][rows][pitch4(cols)]
int launch_split_c64(const float2* in, size_t n_mat_rows, int cols, float* p0, float* p1, float* p2,
                     float* p3, cudaStream_t st);
int launch_pupil(int N, int L, const float* T, const float* opd, const float* phase,
                 const float* wavenumber, const float* amp_scale /*device scalar*/,
                 float* p0, float* p1, float* p2, float* p3, cudaStream_t st);
int launch_power(int N, const float* T, int normalise, float* amp_scale, cudaStream_t st);
int launch_cotangent(int M, int n_items, const float2* field, const float* psf_bar,
                     const float* w, float* p0, float* p1, float* p2, float* p3, float* w_bar,
                     cudaStream_t st);
int launch_basis_eval(int nz, int64_t npix, const float* basis, const float* coeffs,
                      const float* base, float* out, cudaStream_t st);
int launch_basis_reduce(int nz, int64_t npix, const float* basis, const float* out_bar,
                        float* coeff_bar, cudaStream_t st);
int launch_zero(float* p, size_t n, cudaStream_t st);
// delta_bar[item][axis] = 2 pi sum_ij x_axis(i or j) * Im(conj(P_ij) Q_ij[item])  (source-offset VJP)
int launch_pos_grad(int N, int n_items, const float2* q, const float* k, const float* T, const float* opd,
                    const float* phase, const float* amp_scale, float a0, float* delta_bar, cudaStream_t st);
// psf[i] (+)= sum_item w[item] |field[item][i]|^2
int launch_psf_reduce(size_t npix, int n_items, const float2* field, const float* w, float* psf,
                      int accumulate, cudaStream_t st);
// Gradient of the pupil phase from the adjoint field Q = MFT^H(Ebar), summed over items:
//   g = Im(conj(P) Q),  P = a0 * amp_scale * T * exp(i (k * opd + phase))
//   opd_bar[i] (+)= sum_item k[item] g[item][i];  phase_bar[i] (+)= sum_item g[item][i]
int launch_grad_reduce(size_t npix, int n_items, const float2* q, const float* k, const float* T,
                       const float* opd, const float* phase, const float* amp_scale, float a0,
                       float* opd_bar, float* phase_bar, int accumulate, cudaStream_t st);

}  // namespace dlux
