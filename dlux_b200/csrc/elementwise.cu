// HBM-bound helpers around the two phasor GEMM stages: coordinate vectors, operand
// splitting, pupil phasor, power normalisation, cotangent, basis eval / reduce.
#include "common.cuh"

namespace dlux {

static inline int grid_for(size_t n, int block, int max_blocks = 148 * 16) {
  size_t g = (n + block - 1) / block;
  if (g > (size_t)max_blocks) g = max_blocks;
  if (g < 1) g = 1;
  return (int)g;
}

// ---------------------------------------------------------------------------
// jnp.linspace lerp form, element i of n (see oracle/mft_oracle.py:jnp_linspace)
__device__ __forceinline__ float lerp_linspace(float start, float stop, int i, int n) {
  if (n == 1) return start;
  if (i == n - 1) return stop;
  const float t = __fdiv_rn((float)i, (float)(n - 1));
  return __fadd_rn(__fmul_rn(start, __fsub_rn(1.0f, t)), __fmul_rn(stop, t));
}

// transfer_matrix's two coordinate vectors (/root/reference/src/dLux/utils/
// propagation.py:113-121 via utils/coordinates.py:329-332), float32 bit-exact.
__global__ void coords_kernel(int n_in, int n_out, int batch, const float* __restrict__ scale_out,
                              const float* __restrict__ shift_xy, const float* __restrict__ delta_xy,
                              float* __restrict__ xin, float* __restrict__ uout, int dft) {
  const int item = blockIdx.y, axis = blockIdx.z;
  if (dft) {   // exact-DFT mode: integer index offsets from the origins shift_xy (input) and delta_xy (output)
    const float j0 = shift_xy ? shift_xy[item * 2 + axis] : 0.0f, b0 = delta_xy ? delta_xy[item * 2 + axis] : 0.0f;
    float* xi = xin + ((size_t)item * 2 + axis) * n_in;
    float* uo = uout + ((size_t)item * 2 + axis) * n_out;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_in; i += gridDim.x * blockDim.x) xi[i] = (float)i - j0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += gridDim.x * blockDim.x) uo[i] = (float)i - b0;
    return;
  }
  const float shift = shift_xy ? shift_xy[item * 2 + axis] : 0.0f;
  const float delta = delta_xy ? delta_xy[item * 2 + axis] : 0.0f;
  const float s = scale_out[item];
  // in_vec: scale_in = 1.0 / n_in is a Python float in the reference -> -(n-1)/2*scale
  // is evaluated in float64 and rounded once; the offset shift*scale_in is float32.
  const double scale_in = 1.0 / (double)n_in;
  const double h_in = (double)(n_in - 1) / 2.0;
  const float off_in = __fmul_rn(shift, (float)scale_in);
  const float start_in = __fsub_rn((float)(-h_in * scale_in), off_in);
  const float stop_in = __fsub_rn((float)(h_in * scale_in), off_in);
  const float h_out = (float)(n_out - 1) * 0.5f;
  const float off_out = __fmul_rn(shift, s);
  const float start_out = __fsub_rn(__fmul_rn(-h_out, s), off_out);
  const float stop_out = __fsub_rn(__fmul_rn(h_out, s), off_out);
  float* xi = xin + ((size_t)item * 2 + axis) * n_in;
  float* uo = uout + ((size_t)item * 2 + axis) * n_out;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_in; i += gridDim.x * blockDim.x)
    xi[i] = lerp_linspace(start_in, stop_in, i, n_in);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += gridDim.x * blockDim.x) {
    float u = lerp_linspace(start_out, stop_out, i, n_out);
    if (delta_xy) u = __fsub_rn(u, delta);
    uo[i] = u;
  }
}

int launch_coords(int n_in, int n_out, int batch, const float* scale_out, const float* shift_xy,
                  const float* delta_xy, int, float* xin, float* uout, cudaStream_t st, int dft) {
  if (batch <= 0) return DLUX_OK;
  int mx = n_in > n_out ? n_in : n_out;
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    int nb = batch - b0 < 65535 ? batch - b0 : 65535;
    dim3 grid((mx + 255) / 256 > 8 ? 8 : (mx + 255) / 256, nb, 2);
    coords_kernel<<<grid, 256, 0, st>>>(n_in, n_out, nb, scale_out + b0,
                                         shift_xy ? shift_xy + 2 * (size_t)b0 : nullptr,
                                         delta_xy ? delta_xy + 2 * (size_t)b0 : nullptr,
                                         xin + (size_t)b0 * 2 * n_in, uout + (size_t)b0 * 2 * n_out, dft);
    note_launch();
  }
  return check_launch("coords");
}

// ---------------------------------------------------------------------------
// complex64 interleaved -> planar operand (PlaneSet)
__global__ void split_c64_kernel(const float2* __restrict__ in, size_t n_rows, int cols, PlaneSet out) {
  const size_t n = n_rows * (size_t)cols;
  const int p4 = pitch4(cols);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / cols;
    const size_t c = i - r * cols;
    const float2 v = in[i];
    plane_store(out, r * p4 + c, v.x, v.y);
  }
}

int launch_split_c64(const float2* in, size_t n_rows, int cols, const PlaneSet& out, cudaStream_t st) {
  if (n_rows == 0) return DLUX_OK;
  split_c64_kernel<<<grid_for(n_rows * cols, 256), 256, 0, st>>>(in, n_rows, cols, out);
  note_launch();
  return check_launch("split_c64");
}

// ---------------------------------------------------------------------------
// power normalisation (wavefronts.py:418-424): amp_scale = sqrt(1 / sum (T/N^2)^2)
// (|exp(i phi)| = 1, so the power depends on the transmission only).  Deterministic
// two-pass sum in float64.  scratch layout: double partial[256] then float amp_scale.
__global__ void power_partial_kernel(int N, const float* __restrict__ T, double* __restrict__ partial) {
  __shared__ double sm[256];
  const size_t n = (size_t)N * N;
  const float a0 = 1.0f / (float)((long long)N * N);
  double acc = 0.0;
  if (T && (n & 3) == 0) {      // 16-byte loads; the float32 products are summed in float64 in a fixed order
    const float4* T4 = reinterpret_cast<const float4*>(T);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n / 4; i += (size_t)gridDim.x * blockDim.x) {
      const float4 t = T4[i];
      const float v0 = a0 * t.x, v1 = a0 * t.y, v2 = a0 * t.z, v3 = a0 * t.w;
      acc += ((double)(v0 * v0) + (double)(v1 * v1)) + ((double)(v2 * v2) + (double)(v3 * v3));
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
      const float v = T ? a0 * T[i] : a0;
      acc += (double)(v * v);
    }
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}

__global__ void power_final_kernel(const double* __restrict__ partial, int n, int normalise,
                                   float* __restrict__ amp_scale) {
  // fixed-order tree over 256 partials: deterministic
  __shared__ double sm[256];
  sm[threadIdx.x] = threadIdx.x < n ? partial[threadIdx.x] : 0.0;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float power = (float)sm[0];
    amp_scale[0] = normalise ? sqrtf(1.0f / power) : 1.0f;
  }
}

int launch_power(int N, const float* T, int normalise, float* amp_scale, cudaStream_t st) {
  // amp_scale points at: [float amp_scale][pad to 8][double partial[256]]
  double* partial = reinterpret_cast<double*>(reinterpret_cast<char*>(amp_scale) + 16);
  power_partial_kernel<<<256, 256, 0, st>>>(N, T, partial);
  power_final_kernel<<<1, 256, 0, st>>>(partial, 256, normalise, amp_scale);
  note_launch(2);
  return check_launch("power");
}

// ---------------------------------------------------------------------------
// pupil phasor P_l = (1/N^2) T exp(i k_l opd) exp(i phase) * amp_scale as planar float32
// (wavefronts.py:111-113, 349, 368; layers/optics.py:91-96).  One thread = 4 consecutive pixels of one
// row for ALL wavelengths: T / opd / phase are read once and the L planes leave as 16-byte stores.
__global__ void pupil_kernel(int N, int L, const float* __restrict__ T, const float* __restrict__ opd,
                             const float* __restrict__ phase, const float* __restrict__ wavenumber,
                             const float* __restrict__ amp_scale, PlaneSet out, const float* __restrict__ tangent) {
  // blockIdx.z = element of a parameter batch: its own OPD map [N, N] and its own L planes
  if (opd) opd += (size_t)blockIdx.z * N * N;
  out.hi[0] += (size_t)blockIdx.z * L * N * pitch4(N);
  out.hi[1] += (size_t)blockIdx.z * L * N * pitch4(N);
  const int p4 = pitch4(N);
  const int groups_per_row = p4 / 4;
  const size_t n_groups = (size_t)N * groups_per_row;
  const float a0 = 1.0f / (float)((long long)N * N);
  const float sc = amp_scale[0];
  const int l0 = blockIdx.y * (int)blockDim.y + threadIdx.y;       // wavelength slice of this thread row
  const int lstep = gridDim.y * blockDim.y;
  for (size_t gidx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; gidx < n_groups;
       gidx += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(gidx / groups_per_row);
    const int c = (int)(gidx - (size_t)r * groups_per_row) * 4;
    float a[4], o[4], pc[4], ps[4], tv[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      a[e] = 0.0f; o[e] = 0.0f; pc[e] = 1.0f; ps[e] = 0.0f; tv[e] = 0.0f;
      if (c + e < N) {
        const size_t i = (size_t)r * N + c + e;
        a[e] = T ? a0 * __ldg(T + i) : a0;
        if (opd) o[e] = __ldg(opd + i);
        if (phase) fast_sincos(__ldg(phase + i), &ps[e], &pc[e]);
        if (tangent) tv[e] = __ldg(tangent + i);
      }
    }
    for (int l = l0; l < L; l += lstep) {
      const float k = __ldg(wavenumber + l);
      float re[4], im[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        re[e] = a[e];
        im[e] = 0.0f;
        if (opd) {
          float sn, cs;
          // exact Cody-Waite reduction + MUFU on |r| <= pi/4 (abs. error ~3e-7): this kernel evaluates L x N^2
          // phasors and was half compute-bound with the polynomial (-40 us per C3 step)
          fast_sincos_mufu(__fmul_rn(k, o[e]), 0, &sn, &cs);
          re[e] = a[e] * cs;
          im[e] = a[e] * sn;
        }
        if (phase) {
          const float r2 = __fsub_rn(__fmul_rn(re[e], pc[e]), __fmul_rn(im[e], ps[e]));
          const float i2 = __fadd_rn(__fmul_rn(re[e], ps[e]), __fmul_rn(im[e], pc[e]));
          re[e] = r2;
          im[e] = i2;
        }
        re[e] *= sc;
        im[e] *= sc;
        if (tangent) {                    // dP = i k V P: the pupil tangent along an OPD direction V
          const float kv = k * tv[e], r2 = -kv * im[e];
          im[e] = kv * re[e];
          re[e] = r2;
        }
      }
      const size_t o4 = ((size_t)l * N + r) * p4 + c;
      *reinterpret_cast<float4*>(out.hi[0] + o4) = make_float4(re[0], re[1], re[2], re[3]);
      *reinterpret_cast<float4*>(out.hi[1] + o4) = make_float4(im[0], im[1], im[2], im[3]);
    }
  }
}

int launch_pupil(int N, int L, const float* T, const float* opd, const float* phase,
                 const float* wavenumber, const float* amp_scale, const PlaneSet& out, cudaStream_t st,
                 int n_batch, const float* tangent) {
  // 64 x 4 threads: x runs along the pixel groups, y splits the wavelengths four ways
  const int ly = L < 4 ? L : 4;
  dim3 block(64, ly);
  dim3 grid(grid_for((size_t)N * (pitch4(N) / 4), 64, n_batch > 1 ? 148 * 2 : 148 * 8), 1, n_batch);
  pupil_kernel<<<grid, block, 0, st>>>(N, L, T, opd, phase, wavenumber, amp_scale, out, tangent);
  note_launch();
  return check_launch("pupil");
}

// ---------------------------------------------------------------------------
// cotangent of the field: Ebar = 2 w psf_bar .* E (planes), w_bar[item] = sum psf_bar |E|^2
__global__ void cotangent_kernel(int M, const float2* __restrict__ field,
                                 const float* __restrict__ psf_bar, const float* __restrict__ w,
                                 PlaneSet out, float* __restrict__ w_bar, int weight_axis, int items_per_bar) {
  __shared__ float sm[256];
  const size_t n = (size_t)M * M;
  const int item = blockIdx.y;
  // parameter batch: item = b * L + l has its own cotangent image psf_bar[b] and the shared weight w[l]
  if (items_per_bar > 0) psf_bar += (size_t)(item / items_per_bar) * n;
  const float w2 = 2.0f * w[items_per_bar > 0 ? item % items_per_bar : item];
  const float half = 0.5f * (float)(M - 1);
  const int p4 = pitch4(M);
  float acc = 0.0f;
  if ((M & 3) == 0) {
    // four consecutive pixels of one row per thread: 2 x 16-byte loads of the field, 16-byte stores to both planes
    const size_t n4 = n / 4;
    const int m4 = M / 4;
    const float4* f4 = reinterpret_cast<const float4*>(field + (size_t)item * n);
    const float4* g4 = reinterpret_cast<const float4*>(psf_bar);
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (size_t)gridDim.x * blockDim.x) {
      const float4 e0 = f4[2 * q], e1 = f4[2 * q + 1];
      const float4 g = g4[q];
      const int r = (int)(q / m4), c0 = (int)(q - (size_t)r * m4) * 4;
      acc += g.x * (e0.x * e0.x + e0.y * e0.y) + g.y * (e0.z * e0.z + e0.w * e0.w) +
             g.z * (e1.x * e1.x + e1.y * e1.y) + g.w * (e1.z * e1.z + e1.w * e1.w);
      float wg[4] = {w2 * g.x, w2 * g.y, w2 * g.z, w2 * g.w};
      if (weight_axis == 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) wg[e] *= (float)(c0 + e) - half;
      } else if (weight_axis == 1) {
#pragma unroll
        for (int e = 0; e < 4; ++e) wg[e] *= (float)r - half;
      }
      const size_t o = ((size_t)item * M + r) * p4 + c0;
      *reinterpret_cast<float4*>(out.hi[0] + o) = make_float4(wg[0] * e0.x, wg[1] * e0.z, wg[2] * e1.x, wg[3] * e1.z);
      *reinterpret_cast<float4*>(out.hi[1] + o) = make_float4(wg[0] * e0.y, wg[1] * e0.w, wg[2] * e1.y, wg[3] * e1.w);
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
      const float2 e = field[(size_t)item * n + i];
      const float g = psf_bar[i];
      acc += g * (e.x * e.x + e.y * e.y);
      const size_t r = i / M;
      float wg = w2 * g;
      if (weight_axis == 0) wg *= (float)(i - r * M) - half;
      else if (weight_axis == 1) wg *= (float)r - half;
      const size_t row = (size_t)item * M + r;
      plane_store(out, row * p4 + (i - r * M), wg * e.x, wg * e.y);
    }
  }
  if (w_bar) {
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(w_bar + item, sm[0]);
  }
}

int launch_cotangent(int M, int n_items, const float2* field, const float* psf_bar, const float* w,
                     const PlaneSet& out, float* w_bar, int weight_axis, cudaStream_t st, int items_per_bar) {
  if (items_per_bar > 0 && n_items > 65535) return DLUX_ERR_SHAPE;   // (chunks of a batch are far smaller)
  for (int b0 = 0; b0 < n_items; b0 += 65535) {
    int nb = n_items - b0 < 65535 ? n_items - b0 : 65535;
    const size_t off = (size_t)b0 * M * M;
    PlaneSet o = out;
    for (int i = 0; i < 2; ++i) o.hi[i] += (size_t)b0 * M * pitch4(M);
    dim3 grid(grid_for((size_t)M * M, 256, 64), nb);
    cotangent_kernel<<<grid, 256, 0, st>>>(M, field + off, psf_bar, items_per_bar > 0 ? w : w + b0, o,
                                            w_bar ? w_bar + b0 : nullptr, weight_axis, items_per_bar);
    note_launch();
  }
  return check_launch("cotangent");
}

// ---------------------------------------------------------------------------
// eval_basis (utils/math.py:177-196) and its transpose
__global__ void basis_eval_kernel(int nz, size_t npix, const float* __restrict__ basis,
                                  const float* __restrict__ coeffs, const float* __restrict__ base,
                                  float* __restrict__ out) {
  extern __shared__ float c_sm[];
  coeffs += (size_t)blockIdx.y * nz;          // blockIdx.y = element of a parameter batch
  out += (size_t)blockIdx.y * npix;
  // (the buffer is padded to a multiple of 8 coefficients: the unrolled loop below may read them in 16-byte pieces)
  for (int z = threadIdx.x; z < ((nz + 7) & ~7); z += blockDim.x) c_sm[z] = z < nz ? coeffs[z] : 0.0f;
  __syncthreads();
  if ((npix & 3) == 0 && ((reinterpret_cast<uintptr_t>(basis) | reinterpret_cast<uintptr_t>(out) |
                            reinterpret_cast<uintptr_t>(base)) & 15) == 0) {
    // four pixels per thread: 16-byte loads, 8 modes = 128 bytes in flight per thread
    const size_t n4 = npix / 4;
    const float4* b4 = reinterpret_cast<const float4*>(basis);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
      float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll 8
      for (int z = 0; z < nz; ++z) {
        const float4 b = __ldg(b4 + (size_t)z * n4 + i);
        const float c = c_sm[z];
        acc.x = fmaf(c, b.x, acc.x); acc.y = fmaf(c, b.y, acc.y); acc.z = fmaf(c, b.z, acc.z); acc.w = fmaf(c, b.w, acc.w);
      }
      if (base) {
        const float4 v = reinterpret_cast<const float4*>(base)[i];
        acc.x = v.x + acc.x; acc.y = v.y + acc.y; acc.z = v.z + acc.z; acc.w = v.w + acc.w;
      }
      reinterpret_cast<float4*>(out)[i] = acc;
    }
    return;
  }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix;
       i += (size_t)gridDim.x * blockDim.x) {
    float acc = 0.0f;
#pragma unroll 8
    for (int z = 0; z < nz; ++z) acc = fmaf(c_sm[z], __ldg(basis + (size_t)z * npix + i), acc);
    out[i] = base ? base[i] + acc : acc;
  }
}

int launch_basis_eval(int nz, int64_t npix, const float* basis, const float* coeffs,
                      const float* base, float* out, cudaStream_t st, int n_batch) {
  dim3 grid(grid_for((npix & 3) == 0 ? (size_t)npix / 4 : (size_t)npix, 256, n_batch > 1 ? 148 * 4 : 148 * 16), n_batch);
  basis_eval_kernel<<<grid, 256, ((nz + 7) & ~7) * sizeof(float), st>>>(nz, (size_t)npix, basis, coeffs, base, out);
  note_launch();
  return check_launch("basis_eval");
}

__global__ void zero_kernel(float* p, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    p[i] = 0.0f;
}

__global__ void basis_reduce_kernel(int nz, size_t npix, const float* __restrict__ basis,
                                    const float* __restrict__ out_bar, float* __restrict__ coeff_bar) {
  __shared__ float sm[4][8];
  constexpr int PER = 4;
  float g[PER];
  out_bar += (size_t)blockIdx.y * npix;       // blockIdx.y = element of a parameter batch
  coeff_bar += (size_t)blockIdx.y * nz;
  // each block owns a contiguous chunk of PER*blockDim pixels (grid sized to cover npix); a thread takes four
  // CONSECUTIVE pixels with 16-byte loads when the arrays allow it, else four pixels blockDim apart
  const bool vec = (npix & 3) == 0 &&
                   ((reinterpret_cast<uintptr_t>(basis) | reinterpret_cast<uintptr_t>(out_bar)) & 15) == 0;
  const size_t base_i = ((size_t)blockIdx.x * blockDim.x) * PER + (vec ? (size_t)threadIdx.x * PER : threadIdx.x);
  const size_t step = vec ? 1 : blockDim.x;
  if (vec) {
    const float4 v = base_i < npix ? *reinterpret_cast<const float4*>(out_bar + base_i) : make_float4(0.f, 0.f, 0.f, 0.f);
    g[0] = v.x; g[1] = v.y; g[2] = v.z; g[3] = v.w;
  } else {
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const size_t i = base_i + (size_t)j * step;
      g[j] = i < npix ? out_bar[i] : 0.0f;
    }
  }
  for (int z0 = 0; z0 < nz; z0 += 4) {        // four modes per round: 16 values (64 bytes) in flight per thread
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int zz = 0; zz < 4; ++zz) {
      if (z0 + zz < nz) {
        const float* bz = basis + (size_t)(z0 + zz) * npix;
        if (vec) {
          if (base_i < npix) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(bz + base_i));
            acc[zz] = fmaf(g[3], b.w, fmaf(g[2], b.z, fmaf(g[1], b.y, g[0] * b.x)));
          }
        } else {
#pragma unroll
          for (int j = 0; j < PER; ++j) {
            const size_t i = base_i + (size_t)j * step;
            if (i < npix) acc[zz] = fmaf(g[j], __ldg(bz + i), acc[zz]);
          }
        }
      }
    }
#pragma unroll
    for (int zz = 0; zz < 4; ++zz) {
      for (int o = 16; o > 0; o >>= 1) acc[zz] += __shfl_xor_sync(0xffffffffu, acc[zz], o);
      if ((threadIdx.x & 31) == 0) sm[zz][threadIdx.x >> 5] = acc[zz];
    }
    __syncthreads();
    if (threadIdx.x < 4 && z0 + threadIdx.x < nz) {
      float s = 0.0f;
      for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) s += sm[threadIdx.x][wv];
      atomicAdd(coeff_bar + z0 + threadIdx.x, s);
    }
    __syncthreads();
  }
}

int launch_basis_reduce(int nz, int64_t npix, const float* basis, const float* out_bar,
                        float* coeff_bar, cudaStream_t st, int n_batch) {
  zero_kernel<<<grid_for((size_t)nz * n_batch, 256), 256, 0, st>>>(coeff_bar, (size_t)nz * n_batch);
  const size_t per_block = 256 * 4;           // PER of basis_reduce_kernel
  dim3 grid((unsigned)(((size_t)npix + per_block - 1) / per_block), n_batch);
  basis_reduce_kernel<<<grid, 256, 0, st>>>(nz, (size_t)npix, basis, out_bar, coeff_bar);
  note_launch(2);
  return check_launch("basis_reduce");
}

// Sums over items replace per-element atomics in the GEMM epilogues: each (source,
// wavelength) writes its contribution with plain coalesced stores and these HBM-bound
// kernels reduce them (optical_systems.py:222-223 `psf.sum(0)`, sources.py:409-411).
__global__ void psf_reduce_kernel(size_t npix, int n_items, const float2* __restrict__ field,
                                  const float* __restrict__ w, float* __restrict__ psf, int accumulate) {
  // blockIdx.y = element of a parameter batch: its own n_items fields and its own image, shared weights
  field += (size_t)blockIdx.y * n_items * npix;
  psf += (size_t)blockIdx.y * npix;
  if ((npix & 1) == 0) {  // two pixels (one float4 of field) per thread; item stride stays 16-B aligned
    const size_t npair = npix / 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npair;
         i += (size_t)gridDim.x * blockDim.x) {
      float a0 = accumulate ? psf[2 * i] : 0.0f, a1 = accumulate ? psf[2 * i + 1] : 0.0f;
      const float4* f = reinterpret_cast<const float4*>(field) + i;
#pragma unroll 8
      for (int it = 0; it < n_items; ++it) {
        const float4 e = f[(size_t)it * npair];
        const float wt = __ldg(w + it);
        a0 = fmaf(wt, e.x * e.x + e.y * e.y, a0);
        a1 = fmaf(wt, e.z * e.z + e.w * e.w, a1);
      }
      psf[2 * i] = a0;
      psf[2 * i + 1] = a1;
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix;
         i += (size_t)gridDim.x * blockDim.x) {
      float acc = accumulate ? psf[i] : 0.0f;
#pragma unroll 8
      for (int it = 0; it < n_items; ++it) {
        const float2 e = field[(size_t)it * npix + i];
        acc = fmaf(__ldg(w + it), e.x * e.x + e.y * e.y, acc);
      }
      psf[i] = acc;
    }
  }
}

int launch_psf_reduce(size_t npix, int n_items, const float2* field, const float* w, float* psf,
                      int accumulate, cudaStream_t st, int n_batch) {
  dim3 grid(grid_for(npix / 2 + 1, 128, n_batch > 1 ? 148 * 4 : 148 * 16), n_batch);
  psf_reduce_kernel<<<grid, 128, 0, st>>>(npix, n_items, field, w, psf, accumulate);
  note_launch();
  return check_launch("psf_reduce");
}

// VJP of the pupil phasor (wavefronts.py:349,368): the adjoint stage leaves Q = MFT^H(Ebar)
// per (source, wavelength); dL/d(phase) = Im(conj(P) Q) with P re-evaluated here from the
// L2-resident pupil arrays.  Pixels outside the aperture (T == 0) cost nothing.
__global__ void grad_reduce_kernel(size_t npix, int n_items, const float2* __restrict__ q,
                                   const float* __restrict__ k, const float* __restrict__ T,
                                   const float* __restrict__ opd, const float* __restrict__ phase,
                                   const float* __restrict__ amp_scale, float a0,
                                   float* __restrict__ opd_bar, float* __restrict__ phase_bar,
                                   float* __restrict__ t_bar, int accumulate) {
  // blockIdx.y = element of a parameter batch (opd_bar only): its own OPD map, adjoint fields and output
  q += (size_t)blockIdx.y * n_items * npix;
  if (opd) opd += (size_t)blockIdx.y * npix;
  if (opd_bar) opd_bar += (size_t)blockIdx.y * npix;
  const float amp = a0 * amp_scale[0];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix;
       i += (size_t)gridDim.x * blockDim.x) {
    float ao = (accumulate && opd_bar) ? opd_bar[i] : 0.0f;
    float ap = (accumulate && phase_bar) ? phase_bar[i] : 0.0f;
    float at = (accumulate && t_bar) ? t_bar[i] : 0.0f;
    const float t = T ? T[i] : 1.0f;
    if (t != 0.0f || t_bar) {  // blocked pixels only matter for the transmission gradient
      const float a = amp * t;
      const float o = opd ? opd[i] : 0.0f;
      const float ph = phase ? phase[i] : 0.0f;
#pragma unroll 4
      for (int it = 0; it < n_items; ++it) {
        const float2 v = q[(size_t)it * npix + i];
        const float kw = __ldg(k + it);
        float sn, cs;
        fast_sincos(__fmul_rn(kw, o) + ph, &sn, &cs);
        const float g = a * (cs * v.y - sn * v.x);   // Im(conj(P) Q)
        ao = fmaf(kw, g, ao);
        ap += g;
        at = fmaf(amp, cs * v.x + sn * v.y, at);     // Re(conj(Q) dP/dT), direct term
      }
    }
    if (opd_bar) opd_bar[i] = ao;
    if (phase_bar) phase_bar[i] = ap;
    if (t_bar) t_bar[i] = at;
  }
}

int launch_grad_reduce(size_t npix, int n_items, const float2* q, const float* k, const float* T,
                       const float* opd, const float* phase, const float* amp_scale, float a0,
                       float* opd_bar, float* phase_bar, float* t_bar, int accumulate, cudaStream_t st,
                       int n_batch) {
  if (n_batch > 1 && (phase_bar || t_bar)) return DLUX_ERR_ARG;
  dim3 grid(grid_for(npix, 128, n_batch > 1 ? 148 * 4 : 148 * 16), n_batch);
  grad_reduce_kernel<<<grid, 128, 0, st>>>(npix, n_items, q, k, T, opd, phase,
                                                                    amp_scale, a0, opd_bar, phase_bar, t_bar,
                                                                    accumulate);
  note_launch();
  return check_launch("grad_reduce");
}


// ---------------------------------------------------------------------------
// Second order of the fused PSF w.r.t. the OPD (see dlux_polypsf_hvp).
// psf_tan[i] (+)= sum_item 2 w[item] Re(conj(E[item][i]) dE[item][i])
__global__ void psf_tangent_kernel(size_t npix, int n_items, const float2* __restrict__ field,
                                   const float2* __restrict__ dfield, const float* __restrict__ w,
                                   float* __restrict__ out, int accumulate) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
    float acc = accumulate ? out[i] : 0.0f;
#pragma unroll 4
    for (int it = 0; it < n_items; ++it) {
      const float2 e = field[(size_t)it * npix + i], d = dfield[(size_t)it * npix + i];
      acc = fmaf(2.0f * __ldg(w + it), e.x * d.x + e.y * d.y, acc);
    }
    out[i] = acc;
  }
}

int launch_psf_tangent(size_t npix, int n_items, const float2* field, const float2* dfield, const float* w,
                       float* out, int accumulate, cudaStream_t st) {
  psf_tangent_kernel<<<grid_for(npix, 128, 148 * 16), 128, 0, st>>>(npix, n_items, field, dfield, w, out, accumulate);
  note_launch();
  return check_launch("psf_tangent");
}

// mode 0: out[i] (+)= sum_item k Im(conj(P) Q[item])          (Q = adjoint of the tangent field's cotangent)
// mode 1: out[i] (+)= -V[i] sum_item k^2 Re(conj(P) Q[item])  (Q = the first-order adjoint field)
__global__ void hv_reduce_kernel(size_t npix, int n_items, const float2* __restrict__ q, const float* __restrict__ k,
                                 const float* __restrict__ T, const float* __restrict__ opd,
                                 const float* __restrict__ phase, const float* __restrict__ amp_scale, float a0,
                                 const float* __restrict__ V, float* __restrict__ out, int mode, int accumulate) {
  const float amp = a0 * amp_scale[0];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
    float acc = 0.0f;
    const float t = T ? T[i] : 1.0f;
    if (t != 0.0f) {
      const float a = amp * t, o = opd ? opd[i] : 0.0f, ph = phase ? phase[i] : 0.0f;
#pragma unroll 4
      for (int it = 0; it < n_items; ++it) {
        const float2 v = q[(size_t)it * npix + i];
        const float kw = __ldg(k + it);
        float sn, cs;
        fast_sincos(__fmul_rn(kw, o) + ph, &sn, &cs);
        if (mode == 0) acc = fmaf(kw, a * (cs * v.y - sn * v.x), acc);
        else acc = fmaf(kw * kw, a * (cs * v.x + sn * v.y), acc);
      }
      if (mode == 1) acc = -V[i] * acc;
    }
    out[i] = accumulate ? out[i] + acc : acc;
  }
}

int launch_hv_reduce(size_t npix, int n_items, const float2* q, const float* k, const float* T, const float* opd,
                     const float* phase, const float* amp_scale, float a0, const float* V, float* out, int mode,
                     int accumulate, cudaStream_t st) {
  hv_reduce_kernel<<<grid_for(npix, 128, 148 * 16), 128, 0, st>>>(npix, n_items, q, k, T, opd, phase, amp_scale, a0, V,
                                                                  out, mode, accumulate);
  note_launch();
  return check_launch("hv_reduce");
}


// ---------------------------------------------------------------------------
// One pass over the adjoint fields Q for every pupil-plane cotangent of the fused backward: the phase-like
// gradients summed over items (grad_reduce_kernel's job) AND the per-item inner products that give the
// source-offset and wavenumber gradients (pos_grad_kernel's job).  Q is the largest array of the backward
// pass (8 N^2 bytes per (source, wavelength): 137 GB per step at 2048 px, 64 stars x 64 wavelengths), so
// reading it once instead of two or three times is what matters.  Items arrive wavelength-major, so the
// pupil phasor's sincos is re-evaluated only when the wavenumber changes.
__global__ void q_reduce_kernel(int N, int n_items, const float2* __restrict__ q, const float* __restrict__ k,
                                const float* __restrict__ T, const float* __restrict__ opd,
                                const float* __restrict__ phase, const float* __restrict__ amp_scale, float a0,
                                float* __restrict__ opd_bar, float* __restrict__ phase_bar, float* __restrict__ t_bar,
                                int accumulate, float* __restrict__ dbar_item, float* __restrict__ kbar_item) {
  extern __shared__ float sm[];          // [n_items][3]: sum x g, sum y g, sum opd g of this block
  const bool per_item = dbar_item != nullptr || kbar_item != nullptr;
  if (per_item) {
    for (int i = threadIdx.x; i < 3 * n_items; i += blockDim.x) sm[i] = 0.0f;
    __syncthreads();
  }
  const size_t npix = (size_t)N * N;
  const float amp = a0 * amp_scale[0];
  const float half = 0.5f * (float)(N - 1), inv = 1.0f / (float)N;
  const int lane = threadIdx.x & 31;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t rounds = (npix + stride - 1) / stride;
  for (size_t rd = 0; rd < rounds; ++rd) {
    const size_t i = rd * stride + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < npix;
    const float t = in ? (T ? T[i] : 1.0f) : 0.0f;
    const bool live = in && (t != 0.0f || t_bar != nullptr);   // blocked pixels only matter for the transmission gradient
    if (__ballot_sync(0xffffffffu, live) == 0u) {
      if (in && !accumulate) {
        if (opd_bar) opd_bar[i] = 0.0f;
        if (phase_bar) phase_bar[i] = 0.0f;
      }
      continue;
    }
    float ao = 0.0f, ap = 0.0f, at = 0.0f;
    if (in && accumulate) {
      if (opd_bar) ao = opd_bar[i];
      if (phase_bar) ap = phase_bar[i];
      if (t_bar) at = t_bar[i];
    }
    const float a = amp * t;
    const float o = (in && opd) ? opd[i] : 0.0f;
    const float ph = (in && phase) ? phase[i] : 0.0f;
    const int r = (int)(i / N), c = (int)(i - (size_t)r * N);
    const float xc = ((float)c - half) * inv, yc = ((float)r - half) * inv;
    float kprev = 0.0f, sn = 0.0f, cs = 1.0f;
    bool have = false;
    constexpr int UN = 8;                     // eight independent loads of Q in flight per thread
    for (int it0 = 0; it0 < n_items; it0 += UN) {
      float2 vv[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        vv[u] = make_float2(0.0f, 0.0f);
        if (live && it0 + u < n_items) vv[u] = q[(size_t)(it0 + u) * npix + i];
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int it = it0 + u;
        if (it >= n_items) break;
        const float kw = __ldg(k + it);
        if (!have || kw != kprev) {
          fast_sincos(__fmul_rn(kw, o) + ph, &sn, &cs);
          kprev = kw;
          have = true;
        }
        const float2 v = vv[u];
        const float g = a * (cs * v.y - sn * v.x);          // Im(conj(P) Q)
        ao = fmaf(kw, g, ao);
        ap += g;
        at = fmaf(amp, cs * v.x + sn * v.y, at);            // Re(conj(Q) dP/dT), direct term
        if (per_item) {
          float sx = xc * g, sy = yc * g, so = o * g;
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) {
            sx += __shfl_xor_sync(0xffffffffu, sx, d);
            sy += __shfl_xor_sync(0xffffffffu, sy, d);
            so += __shfl_xor_sync(0xffffffffu, so, d);
          }
          if (lane == 0) {
            atomicAdd(sm + 3 * it, sx);
            atomicAdd(sm + 3 * it + 1, sy);
            atomicAdd(sm + 3 * it + 2, so);
          }
        }
      }
    }
    if (in) {
      if (opd_bar) opd_bar[i] = ao;
      if (phase_bar) phase_bar[i] = ap;
      if (t_bar) t_bar[i] = at;
    }
  }
  if (per_item) {
    __syncthreads();
    const float two_pi = 6.283185307179586f;
    for (int it = threadIdx.x; it < n_items; it += blockDim.x) {
      if (dbar_item) {
        atomicAdd(dbar_item + 2 * it, two_pi * sm[3 * it]);
        atomicAdd(dbar_item + 2 * it + 1, two_pi * sm[3 * it + 1]);
      }
      if (kbar_item) atomicAdd(kbar_item + it, sm[3 * it + 2]);
    }
  }
}

int launch_q_reduce(int N, int n_items, const float2* q, const float* k, const float* T, const float* opd,
                    const float* phase, const float* amp_scale, float a0, float* opd_bar, float* phase_bar,
                    float* t_bar, int accumulate, float* dbar_item, float* kbar_item, cudaStream_t st) {
  const size_t npix = (size_t)N * N;
  constexpr int MAX_ITEMS = 2048;        // 24 KiB of per-item partial sums per block
  for (int b0 = 0; b0 < n_items; b0 += MAX_ITEMS) {
    const int nb = n_items - b0 < MAX_ITEMS ? n_items - b0 : MAX_ITEMS;
    const bool per_item = dbar_item || kbar_item;
    // per-item sums want few, fat blocks (one flush of the shared partials each); the pixel-only case wants many
    const int threads = per_item ? 256 : 128;
    q_reduce_kernel<<<grid_for(npix, threads, per_item ? 148 * 4 : 148 * 16), threads,
                      per_item ? 3 * nb * sizeof(float) : 0, st>>>(
        N, nb, q + (size_t)b0 * npix, k + b0, T, opd, phase, amp_scale, a0, opd_bar, phase_bar, t_bar,
        (accumulate || b0 > 0) ? 1 : 0, dbar_item ? dbar_item + 2 * (size_t)b0 : nullptr,
        kbar_item ? kbar_item + b0 : nullptr);
    note_launch();
  }
  return check_launch("q_reduce");
}

// Power normalisation (wavefronts.py:418-424) makes amp = (sum_j (a0 T_j)^2)^(-1/2) depend on
// T: dL/dT_j gets  -(sum_i T_i Tbar_i) * amp^2 * a0^2 * T_j  on top of the direct term.
__global__ void tbar_partial_kernel(size_t npix, const float* __restrict__ T, const float* __restrict__ t_bar,
                                    double* __restrict__ partial) {
  __shared__ double sm[256];
  double acc = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix;
       i += (size_t)gridDim.x * blockDim.x)
    acc += (double)T[i] * (double)t_bar[i];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}

__global__ void tbar_apply_kernel(size_t npix, const float* __restrict__ T, const float* __restrict__ amp_scale,
                                  float a0, const double* __restrict__ partial, float* __restrict__ t_bar) {
  __shared__ double sm[256];
  sm[threadIdx.x] = partial[threadIdx.x];   // every block re-sums the 256 partials in the same order
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  const float amp = amp_scale[0];
  const float coef = (float)sm[0] * amp * amp * a0 * a0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix;
       i += (size_t)gridDim.x * blockDim.x)
    t_bar[i] -= coef * T[i];
}

int launch_tbar_finalize(size_t npix, const float* T, const float* amp_scale, float a0, float* t_bar,
                         double* work, cudaStream_t st) {
  tbar_partial_kernel<<<256, 256, 0, st>>>(npix, T, t_bar, work);
  tbar_apply_kernel<<<grid_for(npix, 256, 148 * 8), 256, 0, st>>>(npix, T, amp_scale, a0, work, t_bar);
  note_launch(2);
  return check_launch("tbar_finalize");
}

// VJP w.r.t. the source offset: the offset delta (fringes) shifts the output coordinates,
// d(phasor)/d(delta_x) = 2 pi i x_j * phasor, hence dL/d(delta_x) = 2 pi sum_ij x_j g_ij with
// the same g = Im(conj(P) Q) as the phase gradient (x = the MFT input coordinate
// (j - (N-1)/2) / N; rows for the y offset).  This is how PointSources.position gets its
// gradient (wavefronts.py:370-395 tilt, folded into the coordinates).
__global__ void pos_grad_kernel(int N, const float2* __restrict__ q, const float* __restrict__ k,
                                const float* __restrict__ T, const float* __restrict__ opd,
                                const float* __restrict__ phase, const float* __restrict__ amp_scale,
                                float a0, float* __restrict__ out, int sel) {
  __shared__ float smx[8], smy[8];
  const int item = blockIdx.y;
  const size_t npix = (size_t)N * N;
  const float amp = a0 * amp_scale[0];
  const float kw = k[item];
  const float half = 0.5f * (float)(N - 1), inv = 1.0f / (float)N;
  float ax = 0.0f, ay = 0.0f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix;
       i += (size_t)gridDim.x * blockDim.x) {
    const float t = T ? T[i] : 1.0f;
    if (t != 0.0f) {
      const float2 v = q[(size_t)item * npix + i];
      float sn, cs;
      fast_sincos(__fmul_rn(kw, opd ? opd[i] : 0.0f) + (phase ? phase[i] : 0.0f), &sn, &cs);
      const float g = amp * t * (cs * v.y - sn * v.x);
      const int r = (int)(i / N), c = (int)(i - (size_t)r * N);
      if (sel == 3) {
        ax = fmaf(opd ? opd[i] : 0.0f, g, ax);   // d phase / d k = opd
      } else {
        ax = fmaf(((float)c - half) * inv, g, ax);
        ay = fmaf(((float)r - half) * inv, g, ay);
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    ax += __shfl_xor_sync(0xffffffffu, ax, o);
    ay += __shfl_xor_sync(0xffffffffu, ay, o);
  }
  if ((threadIdx.x & 31) == 0) { smx[threadIdx.x >> 5] = ax; smy[threadIdx.x >> 5] = ay; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float sx = 0.0f, sy = 0.0f;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) { sx += smx[wv]; sy += smy[wv]; }
    const float two_pi = 6.283185307179586f;
    if (sel == 0) {
      atomicAdd(out + 2 * item, two_pi * sx);
      atomicAdd(out + 2 * item + 1, two_pi * sy);
    } else if (sel == 3) {
      atomicAdd(out + item, sx);
    } else {
      atomicAdd(out + item, -two_pi * (sel == 1 ? sx : sy));
    }
  }
}

int launch_pos_grad(int N, int n_items, const float2* q, const float* k, const float* T, const float* opd,
                    const float* phase, const float* amp_scale, float a0, float* out, int sel, cudaStream_t st) {
  for (int b0 = 0; b0 < n_items; b0 += 65535) {
    const int nb = n_items - b0 < 65535 ? n_items - b0 : 65535;
    dim3 grid(grid_for((size_t)N * N, 256, 32), nb);
    pos_grad_kernel<<<grid, 256, 0, st>>>(N, q + (size_t)b0 * N * N, k + b0, T, opd, phase, amp_scale, a0,
                                          out + (sel == 0 ? 2 : 1) * (size_t)b0, sel);
    note_launch();
  }
  return check_launch("pos_grad");
}


// ---------------------------------------------------------------------------
// Exact zero-block skipping (opt-in): which blocks of the pupil hold anything but T == 0.
// flags[br][bc] = any(T[br*bh .. +bh][bc*bw .. +bw] != 0)
__global__ void block_mask_kernel(int N, const float* __restrict__ T, int bh, int bw, int nbc, int* __restrict__ flags) {
  const int br = blockIdx.x / nbc, bc = blockIdx.x % nbc;
  int any = 0;
  for (int e = threadIdx.x; e < bh * bw; e += blockDim.x) {
    const int r = br * bh + e / bw, c = bc * bw + e % bw;
    if (r < N && c < N && T[(size_t)r * N + c] != 0.0f) any = 1;
  }
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) flags[blockIdx.x] = any;
}
// row-wise compaction of the flags into index lists: idx[row][0 .. cnt[row]) = the set columns, ascending
__global__ void compact_rows_kernel(int cols, const int* __restrict__ flags, int* __restrict__ cnt, int* __restrict__ idx) {
  if (threadIdx.x != 0) return;
  const int row = blockIdx.x;
  int n = 0;
  for (int c = 0; c < cols; ++c)
    if (flags[row * cols + c]) idx[row * cols + n++] = c;
  cnt[row] = n;
}

int launch_block_lists(int N, const float* T, int bh, int bw, int rows_as_one, int* flags, int* cnt, int* idx,
                       cudaStream_t st) {
  const int nbr = (N + bh - 1) / bh, nbc = (N + bw - 1) / bw;
  block_mask_kernel<<<nbr * nbc, 256, 0, st>>>(N, T, bh, bw, nbc, flags);
  if (rows_as_one) compact_rows_kernel<<<1, 32, 0, st>>>(nbr * nbc, flags, cnt, idx);
  else compact_rows_kernel<<<nbr, 32, 0, st>>>(nbc, flags, cnt, idx);
  note_launch(2);
  return check_launch("block_lists");
}

int launch_zero(float* p, size_t n, cudaStream_t st) {
  if (n == 0) return DLUX_OK;
  zero_kernel<<<grid_for(n, 256), 256, 0, st>>>(p, n);
  note_launch();
  return check_launch("zero");
}

}  // namespace dlux
