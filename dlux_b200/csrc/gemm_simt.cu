// DLUX_PREC_FP32: CUDA-core FFMA version of the phasor GEMM stage.  Same operands,
// same on-the-fly phasor generation and the same epilogues as the tcgen05 kernel
// (gemm_tc.cu); used to validate it and for exact-fp32 accumulation on request.
#include "common.cuh"

namespace dlux {

namespace {
constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

__global__ void __launch_bounds__(NT) gemm_simt_kernel(const GemmParams p, int tiles_m, int tiles_n) {
  __shared__ float As_re[BK][BM + 4], As_im[BK][BM + 4];
  __shared__ float Gs_re[BK][BN + 4], Gs_im[BK][BN + 4];
  const int tiles = tiles_m * tiles_n;
  const int item = blockIdx.x / tiles;
  const int t = blockIdx.x % tiles;
  const int m0 = (t % tiles_m) * BM, n0 = (t / tiles_m) * BN;
  const int d = p.item_data ? __ldg(p.item_data + item) : item;
  const int a_pitch = pitch4(p.K);
  const size_t a_off = (size_t)d * p.rows * a_pitch;
  const float* kv = p.kvec + (size_t)item * p.kvec_stride;
  const float* nv = p.nvec + (size_t)item * p.nvec_stride;
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;

  float acc_re[4][4], acc_im[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc_re[i][j] = acc_im[i][j] = 0.0f;

  for (int k0 = 0; k0 < p.K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < (BM * BK) / NT; ++i) {
      const int idx = tid + i * NT;
      const int ml = idx / BK, kl = idx % BK;
      const int m = m0 + ml, k = k0 + kl;
      float re = 0.0f, im = 0.0f;
      if (m < p.rows && k < p.K) {
        const size_t o = a_off + (size_t)m * a_pitch + k;
        re = __ldg(p.a.hi[0] + o);
        im = __ldg(p.a.hi[1] + o);
      }
      As_re[kl][ml] = re;
      As_im[kl][ml] = im;
    }
#pragma unroll
    for (int i = 0; i < (BN * BK) / NT; ++i) {
      const int idx = tid + i * NT;
      const int kl = idx / BN, nl = idx % BN;
      const int n = n0 + nl, k = k0 + kl;
      float s = 0.0f, c = 0.0f;
      if (n < p.n_out && k < p.K) {
        const float arg = p.dft_period > 0.0f
                              ? dft_arg(p.sign2pi, __ldg(kv + k), __ldg(nv + n), p.dft_period, 1.0f / p.dft_period)
                              : phase_arg(p.sign2pi, __ldg(kv + k), __ldg(nv + n));
        sincosf(arg, &s, &c);
      }
      Gs_re[kl][nl] = c;
      Gs_im[kl][nl] = s;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float ar[4], ai[4], gr[4], gi[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        ar[i] = As_re[kk][tx * 4 + i];
        ai[i] = As_im[kk][tx * 4 + i];
        gr[i] = Gs_re[kk][ty * 4 + i];
        gi[i] = Gs_im[kk][ty * 4 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc_re[i][j] = fmaf(ar[i], gr[j], acc_re[i][j]);
          acc_re[i][j] = fmaf(-ai[i], gi[j], acc_re[i][j]);
          acc_im[i][j] = fmaf(ar[i], gi[j], acc_im[i][j]);
          acc_im[i][j] = fmaf(ai[i], gr[j], acc_im[i][j]);
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + ty * 4 + j;
    if (n >= p.n_out) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + tx * 4 + i;
      if (m < p.rows) epilogue_store(p, item, m, n, acc_re[i][j], acc_im[i][j]);
    }
  }
}
}  // namespace

int launch_gemm_simt(const GemmParams& p, cudaStream_t st) {
  if (p.n_items <= 0) return DLUX_OK;
  const int tiles_m = (p.rows + BM - 1) / BM, tiles_n = (p.n_out + BN - 1) / BN;
  const long long total = (long long)tiles_m * tiles_n * p.n_items;
  if (total > 2147483647LL) return DLUX_ERR_SHAPE;
  gemm_simt_kernel<<<(unsigned)total, NT, 0, st>>>(p, tiles_m, tiles_n);
  note_launch();
  return check_launch("gemm_simt");
}

}  // namespace dlux
