// Shared declarations of the dlux_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/dlux_b200.h"

namespace dlux {

// ---------------------------------------------------------------------------
// One "phasor GEMM" stage.  Every stage of the forward and adjoint MFT is
//
//   Out[item][n][m] = scale[item] * sum_k Data[d(item)][m][k] * exp(i * sign2pi * fl(kvec[k] * nvec[n]))
//
// i.e. a complex contraction of a data matrix (planar, split operand planes, see PlaneSet)
// with a DFT phasor matrix that is generated on the fly from two float32
// coordinate vectors, written TRANSPOSED (m fastest) so that two stages chain.
// The phase argument is formed exactly as the reference does
// (/root/reference/src/dLux/utils/propagation.py:124): fl32(fl32(-2pi) * fl32(x*u)).
// ---------------------------------------------------------------------------
enum EpilogueMode : int {
  EPI_PLANES = 0,  // out planes [item][n][m]  (feeds the next stage)
  EPI_C64 = 1,     // out_c64[item][n][m]
  EPI_PSF = 2      // out_psf[n][m] += item_w[item] * |scale[item] * D|^2   (forward-only image: no field is written)
};

// Planar representation of a complex matrix Z[n][rows][K], 8 bytes per element: two float32 planes
//   hi[0] = Re Z, hi[1] = Im Z      [n][rows][pitch4(K)]
// (row pitches are multiples of 16 bytes, as TMA requires).  The values are UNSPLIT float32.
// The tensor kernel splits them on chip: after a tile-chunk lands in shared memory, converter warps
// round each value to tf32 ("hi", written back in place for the kind::tf32 MMA) and derive
// bf16(hi) and bf16(z - hi) planes for the two bf16 correction MMAs,
//   z*g = hi*g_hi (tf32) + hi*g_lo + lo*g_hi (bf16):
// the correction terms are 2^-11 of the product, so bf16's 2^-9 leaves ~2^-20 (measured 5.6e-7
// relative per contraction) at 2/3 of the tensor work of 3xTF32.  Round 1 kept the six split planes
// in HBM (16 bytes per element); splitting on chip halves every operand's traffic and the
// intermediate's epilogue.
struct PlaneSet {
  float* hi[2];
};
__host__ __device__ inline int pitch4(int k) { return (k + 3) & ~3; }

struct GemmParams {
  PlaneSet a;   // data operand, [n_data][rows][K]
  int n_data;   // number of data matrices in `a`
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
  PlaneSet out;          // EPI_PLANES: [n_items][n_out][rows] (pitch4 of rows)
  float2* out_c64;       // EPI_C64: [n_items][n_out][rows]
  float* out_psf;        // EPI_PSF: [n_out][rows], accumulated over the items (zeroed by the caller)
  const float* item_w;   // EPI_PSF: [n_items] spectral weight x flux
  // Exact zero-block skipping (opt-in, tensor kernel only; all nullptr = dense):
  //  chunk_cnt[tiles_mp], chunk_idx[tiles_mp][k_chunks]: for every pair of 128-row data tiles, the 16-k
  //  chunks in which those 256 rows are not all zero (device arrays; the same for every item);
  //  unit_list[*unit_count]: the (n-tile pair, m-tile pair) output blocks that are needed at all, as
  //  t = np * tiles_mp + mp (the count is read on the device too: nothing is copied back to the host).
  const int* chunk_cnt;
  const int* chunk_idx;
  const int* unit_list;
  const int* unit_count;
  // Exact-DFT mode (dlu.FFT): kvec / nvec hold integer index offsets and the phase is +-2 pi ((k n) mod P) / P
  float dft_period;      // P = padded size, 0 = the MFT's float32 phase form
};

constexpr int GEMM_TC_BM2 = 256;   // data rows per unit (two 128-row tiles)
constexpr int GEMM_TC_BN2 = 128;   // output coordinates per unit (64 per CTA of the pair)
constexpr int GEMM_TC_BK = 16;     // k per chunk

// round-to-nearest (ties away from zero) to tf32's 10 mantissa bits.  Equivalent to
// cvt.rna.tf32.f32 for finite inputs; spelled with integer ops because sm_100a expands that
// cvt into a ~5-instruction sequence and the hot loops do it twice per phasor / output.
__device__ __forceinline__ float tf32_hi(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

// Reference phase argument: two float32 multiplies, no fused contraction.
__device__ __forceinline__ float phase_arg(float sign2pi, float x, float u) {
  return __fmul_rn(sign2pi, __fmul_rn(x, u));
}
// Exact-DFT phase argument (dlu.FFT through these kernels): x and u are integer index offsets, the phase is
// +-2 pi ((x u) mod P) / P.  x u is an exact integer in float32 (|x u| < 2^24 is checked on the host), the
// reduction is exact, and only the final scaling rounds -- the float32 MFT form fl(2 pi fl(x u / P)) would carry
// ~1e-4 rad of rounding at N_pad = 2048, which the reference's FFT (an FFT, not a phasor matrix) does not have.
__device__ __forceinline__ float dft_arg(float sign2pi, float x, float u, float period, float inv_period) {
  const float pf = x * u;
  const float q = rintf(pf * inv_period);
  const float r = fmaf(-q, period, pf);          // exact: |r| <= period
  return sign2pi * (r * inv_period);
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

// Lean variant for the in-kernel phasor generators: the same exact Cody-Waite reduction,
// then MUFU.SIN / MUFU.COS on the reduced argument |r| <= pi/4 (abs. error ~3e-7, about 4x
// the polynomial's; the parity tests bound the end-to-end effect) and branch-free quadrant
// logic on the sign bits.  `qshift` rotates the result by quarter turns exactly: returns
// sincos(a + qshift * pi/2).  No slow path: the 3-term reduction is accurate for
// |a| < ~1e5 rad (MFT phases are ~ pi * nfringes / 2) and degrades gracefully beyond.
__device__ __forceinline__ void fast_sincos_mufu(float a, int qshift, float* sn, float* cs) {
  float j = fmaf(a, 0.636619747f, 12582912.0f);
  const uint32_t q = (uint32_t)(__float_as_int(j) + qshift);
  j -= 12582912.0f;
  float r = fmaf(j, -1.57079601e+00f, a);
  r = fmaf(j, -3.13916473e-07f, r);
  r = fmaf(j, -5.39030253e-15f, r);
  const float s = __sinf(r), c = __cosf(r);
  const bool swap = (q & 1u) != 0;
  const uint32_t so = __float_as_uint(swap ? c : s);
  const uint32_t co = __float_as_uint(swap ? s : c);
  *sn = __uint_as_float(so ^ ((q & 2u) << 30));          // negate in quadrants 2, 3
  *cs = __uint_as_float(co ^ (((q + 1u) & 2u) << 30));   // negate in quadrants 1, 2
}

// Leanest variant, for the tensor-core kernel's generator warps (their instruction count bounds the kernel):
// ONE reduction modulo 2 pi (the same Cody-Waite constants x 4; the third term is < 4e-10 rad for |a| < 1e5 and
// is dropped), then MUFU.SIN / MUFU.COS on |r| <= pi, where the unit's absolute error is the same ~3e-7 as on
// the first quadrant -- no quadrant swap / sign logic at all (9 instructions instead of 19 per phasor).
__device__ __forceinline__ void fast_sincos_turn(float a, float* sn, float* cs) {
  float j = fmaf(a, 0.159154937f, 12582912.0f);
  j -= 12582912.0f;
  float r = fmaf(j, -4.0f * 1.57079601e+00f, a);
  r = fmaf(j, -4.0f * 3.13916473e-07f, r);
  *sn = __sinf(r);
  *cs = __cosf(r);
}

// (re, im) -> the two operand planes at element offset o4
__device__ __forceinline__ void plane_store(const PlaneSet& ps, size_t o4, float re, float im) {
  ps.hi[0][o4] = re;
  ps.hi[1][o4] = im;
}

// Shared epilogue for one output element D[m][n] = (re, im) of `item`.
__device__ __forceinline__ void epilogue_store(const GemmParams& p, int item, int m, int n,
                                               float re, float im) {
  const float sc = p.scale ? __ldg(p.scale + item) : 1.0f;
  re *= sc;
  im *= sc;
  const size_t idx = ((size_t)item * p.n_out + n) * p.rows + m;
  if (p.mode == EPI_PLANES) {
    const size_t row = (size_t)item * p.n_out + n;
    plane_store(p.out, row * pitch4(p.rows) + m, re, im);
  } else if (p.mode == EPI_PSF) {
    atomicAdd(p.out_psf + (size_t)n * p.rows + m, __ldg(p.item_w + item) * (re * re + im * im));
  } else {  // EPI_C64
    p.out_c64[idx] = make_float2(re, im);
  }
}

// launch counters / error plumbing (api.cu)
void note_launch(int n = 1);
int check_launch(const char* what);
void note_cuda_error(int code);   // remembered for dlux_last_cuda_error()

// kernels' host launchers
int launch_gemm_simt(const GemmParams& p, cudaStream_t st);
int launch_gemm_tc(const GemmParams& p, cudaStream_t st);
// fused two-stage launch (see gemm_tc.cu): ring slots it would use (0 = not applicable) and the launch itself
int gemm_tc_fused_ring(const GemmParams& g1, const GemmParams& g2, int max_ring, int* lag_out);
int launch_gemm_tc_fused(const GemmParams& g1, const GemmParams& g2, int ring, int lag, int* sync_ws, cudaStream_t st);
size_t gemm_tc_workspace_bytes();
int launch_tc_peak_probe(int kind, int n_batches, float* sink, double* flops, cudaStream_t st);

int launch_coords(int n_in, int n_out, int batch, const float* scale_out, const float* shift_xy,
                  const float* delta_xy, int delta_stride_items, float* xin, float* uout,
                  cudaStream_t st, int dft = 0);
// in: [n_mat][rows][cols] c64 -> split planes
int launch_split_c64(const float2* in, size_t n_mat_rows, int cols, const PlaneSet& out, cudaStream_t st);
int launch_pupil(int N, int L, const float* T, const float* opd, const float* phase,
                 const float* wavenumber, const float* amp_scale /*device scalar*/,
                 const PlaneSet& out, cudaStream_t st, int n_batch = 1, const float* tangent = nullptr);
int launch_power(int N, const float* T, int normalise, float* amp_scale, cudaStream_t st);
// `weight_axis`: -1 none; 0 / 1: multiply by the output-pixel index (col / row) minus (M-1)/2,
// the derivative of the output coordinate w.r.t. scale_out
int launch_cotangent(int M, int n_items, const float2* field, const float* psf_bar,
                     const float* w, const PlaneSet& out, float* w_bar, int weight_axis, cudaStream_t st,
                     int items_per_bar = 0);
int launch_basis_eval(int nz, int64_t npix, const float* basis, const float* coeffs,
                      const float* base, float* out, cudaStream_t st, int n_batch = 1);
int launch_basis_reduce(int nz, int64_t npix, const float* basis, const float* out_bar,
                        float* coeff_bar, cudaStream_t st, int n_batch = 1);
int launch_zero(float* p, size_t n, cudaStream_t st);
// flags / cnt / idx of the non-zero (bh x bw) blocks of T [N, N]: per block row, or (rows_as_one) one list of
// block indices br * nbc + bc
int launch_block_lists(int N, const float* T, int bh, int bw, int rows_as_one, int* flags, int* cnt, int* idx,
                       cudaStream_t st);
// delta_bar[item][axis] = 2 pi sum_ij x_axis(i or j) * Im(conj(P_ij) Q_ij[item])  (source-offset VJP)
// sel 0: delta_bar[item][2] += 2 pi (sum x g, sum y g); sel 1 / 2: out[item] -= 2 pi sum x g / sum y g;
// sel 3: out[item] += sum opd g   (g = Im(conj(P) Q))
int launch_pos_grad(int N, int n_items, const float2* q, const float* k, const float* T, const float* opd,
                    const float* phase, const float* amp_scale, float a0, float* out, int sel, cudaStream_t st);
// psf[i] (+)= sum_item w[item] |field[item][i]|^2
int launch_psf_reduce(size_t npix, int n_items, const float2* field, const float* w, float* psf,
                      int accumulate, cudaStream_t st, int n_batch = 1);
// Gradient of the pupil phase from the adjoint field Q = MFT^H(Ebar), summed over items:
//   g = Im(conj(P) Q),  P = a0 * amp_scale * T * exp(i (k * opd + phase))
//   opd_bar[i] (+)= sum_item k[item] g[item][i];  phase_bar[i] (+)= sum_item g[item][i]
int launch_grad_reduce(size_t npix, int n_items, const float2* q, const float* k, const float* T,
                       const float* opd, const float* phase, const float* amp_scale, float a0,
                       float* opd_bar, float* phase_bar, float* t_bar, int accumulate, cudaStream_t st,
                       int n_batch = 1);
int launch_psf_tangent(size_t npix, int n_items, const float2* field, const float2* dfield, const float* w,
                       float* out, int accumulate, cudaStream_t st);
int launch_hv_reduce(size_t npix, int n_items, const float2* q, const float* k, const float* T, const float* opd,
                     const float* phase, const float* amp_scale, float a0, const float* V, float* out, int mode,
                     int accumulate, cudaStream_t st);
int launch_q_reduce(int N, int n_items, const float2* q, const float* k, const float* T, const float* opd,
                    const float* phase, const float* amp_scale, float a0, float* opd_bar, float* phase_bar,
                    float* t_bar, int accumulate, float* dbar_item, float* kbar_item, cudaStream_t st);
// T_bar -= (sum_i T_i T_bar_i) * amp^2 * a0^2 * T   (the power-normalisation term; `work` = 256 doubles)
int launch_tbar_finalize(size_t npix, const float* T, const float* amp_scale, float a0, float* t_bar,
                         double* work, cudaStream_t st);

}  // namespace dlux
