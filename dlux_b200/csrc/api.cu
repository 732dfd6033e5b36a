// extern "C" surface of libdlux_b200.so (see include/dlux_b200.h).
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>
#include "common.cuh"

namespace dlux {

static std::atomic<uint64_t> g_launches{0};
static thread_local int g_last_cuda_error = 0;

void note_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
void note_cuda_error(int code) { g_last_cuda_error = code; }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    g_last_cuda_error = (int)e;
    fprintf(stderr, "[dlux_b200] CUDA error in %s: %s\n", what, cudaGetErrorString(e));
    return DLUX_ERR_CUDA;
  }
  return DLUX_OK;
}

int launch_zero(float* p, size_t n, cudaStream_t st);

// Per-item parameter expansion for the fused poly-PSF path.  The caller's arrays are source-major
// ([S, L]: index s * L + l); items are PROCESSED wavelength-major (i = l * S + s) so that consecutive items
// -- the stars of one wavelength -- contract the same pupil phasor P(l), which then stays L2-resident
// instead of being re-read from HBM for every star (BASELINE config 4: 1000 stars x 64 wavelengths).
__global__ void expand_items_kernel(int n_items, int L, int S, const float* __restrict__ scale_out,
                                    const float* __restrict__ norm, const float* __restrict__ wavenumber,
                                    const float* __restrict__ weights, const float* __restrict__ delta_xy,
                                    int* __restrict__ item_l, float* __restrict__ s_item,
                                    float* __restrict__ norm_item, float* __restrict__ k_item,
                                    float* __restrict__ w_item, float* __restrict__ delta_item) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += gridDim.x * blockDim.x) {
    const int l = i / S, sidx = i - l * S;
    const int src = sidx * L + l;
    item_l[i] = l;
    s_item[i] = scale_out[l];
    norm_item[i] = norm ? norm[l] : 1.0f;
    k_item[i] = wavenumber[l];
    if (w_item) w_item[i] = weights[src];
    if (delta_item && delta_xy) {
      delta_item[2 * i] = delta_xy[2 * src];
      delta_item[2 * i + 1] = delta_xy[2 * src + 1];
    }
  }
}

// processing order -> the caller's [S, L] layout, `width` values per item
__global__ void scatter_items_kernel(int n_items, int L, int S, int width, const float* __restrict__ in,
                                     float* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items * width; i += gridDim.x * blockDim.x) {
    const int it = i / width, e = i - it * width;
    const int l = it / S, sidx = it - l * S;
    out[(sidx * L + l) * width + e] = in[i];
  }
}

struct Bump {
  char* base;
  size_t off, cap;
  bool ok;
  Bump(void* b, size_t c) : base((char*)b), off(0), cap(c), ok(true) {}
  template <class T>
  T* take(size_t count) {
    off = (off + 1023) & ~(size_t)1023;
    const size_t bytes = count * sizeof(T);
    if (base && off + bytes > cap) ok = false;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += bytes;
    return p;
  }
  size_t used() const { return (off + 1023) & ~(size_t)1023; }
};

struct ProfRec { cudaEvent_t a, b; double flops; };
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
static std::atomic<int> g_prof_on{0};

// one split complex matrix stack of `count` rows x cols elements
static PlaneSet take_planes(Bump& b, size_t n_rows, int cols) {
  PlaneSet ps;
  for (int i = 0; i < 2; ++i) ps.hi[i] = b.take<float>(n_rows * pitch4(cols));
  return ps;
}
static size_t plane_bytes(size_t n_rows, int cols) {
  return n_rows * 8 * (size_t)pitch4(cols);
}
// the stage-1 output of the forward ([M][N]) and of the adjoint ([N][M]) share one buffer,
// sized for the larger, and are indexed with their own pitches
static PlaneSet take_mid_planes(Bump& b, size_t c, int N, int M) {
  PlaneSet ps;
  const size_t f4 = std::max((size_t)M * pitch4(N), (size_t)N * pitch4(M));
  for (int i = 0; i < 2; ++i) ps.hi[i] = b.take<float>(c * f4);
  return ps;
}
static size_t mid_bytes(int N, int M) {
  return 8 * std::max((size_t)M * pitch4(N), (size_t)N * pitch4(M));
}

static int run_gemm(const GemmParams& p, int precision, cudaStream_t st) {
  ProfRec r{};
  const bool prof = g_prof_on.load(std::memory_order_relaxed) != 0;
  if (prof) {
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    r.flops = 8.0 * (double)p.rows * p.K * (double)p.n_out * p.n_items;
    cudaEventRecord(r.a, st);
  }
  int rc;
  if (precision == DLUX_PREC_FP32) rc = launch_gemm_simt(p, st);
  else if (precision == DLUX_PREC_3XTF32) rc = launch_gemm_tc(p, st);
  else rc = DLUX_ERR_ARG;
  if (prof) {
    cudaEventRecord(r.b, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(r);
  }
  return rc;
}

// Stage 1 + stage 2 of the same items: ONE persistent launch with the intermediate in an L2-resident ring
// (gemm_tc.cu, FUSED) when the tensor path applies, else two launches through the full intermediate.
static int run_gemm_pair(const GemmParams& g1, const GemmParams& g2, int* sync_ws, int precision, cudaStream_t st) {
  static const bool no_fuse = getenv("DLUX_B200_NO_FUSE") != nullptr;
  if (precision == DLUX_PREC_3XTF32 && sync_ws && !no_fuse && g1.dft_period == 0.0f && !g1.chunk_cnt && !g1.unit_list &&
      !g2.chunk_cnt && !g2.unit_list) {
    int lag = 0;
    const int ring = gemm_tc_fused_ring(g1, g2, g1.n_items, &lag);
    if (ring > 0) {
      ProfRec r{};
      const bool prof = g_prof_on.load(std::memory_order_relaxed) != 0;
      if (prof) {
        cudaEventCreate(&r.a);
        cudaEventCreate(&r.b);
        r.flops = 8.0 * (double)g1.rows * g1.K * (double)g1.n_out * g1.n_items +
                  8.0 * (double)g2.rows * g2.K * (double)g2.n_out * g2.n_items;
        cudaEventRecord(r.a, st);
      }
      const int rc = launch_gemm_tc_fused(g1, g2, ring, lag, sync_ws, st);
      if (prof) {
        cudaEventRecord(r.b, st);
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_prof.push_back(r);
      }
      return rc;
    }
  }
  int rc = run_gemm(g1, precision, st);
  if (rc) return rc;
  return run_gemm(g2, precision, st);
}

// per-chunk intermediates, bytes: 6 GiB of a 180 GB board (DLUX_B200_CHUNK_MB overrides; the tests use it
// to force the multi-chunk paths at small sizes).  Larger chunks = fewer, longer persistent launches
static size_t chunk_budget() {
  static const size_t v = [] {
    const char* e = getenv("DLUX_B200_CHUNK_MB");
    const long mb = e ? atol(e) : 0;
    return mb > 0 ? (size_t)mb << 20 : (size_t)6 << 30;
  }();
  return v;
}
#define kChunkBudget chunk_budget()

// largest chunk <= cmax that splits `n` items into equal-sized chunks (no small tail launch)
static size_t balanced_chunk(size_t n, size_t cmax) {
  if (cmax < 1) cmax = 1;
  if (n <= cmax) return n < 1 ? 1 : n;
  const size_t parts = (n + cmax - 1) / cmax;
  return (n + parts - 1) / parts;
}

static int mft_chunk(const dlux_mft_desc* d) {
  const size_t n_src = d->adjoint ? d->n_out : d->n_in;
  const size_t per_item = plane_bytes(n_src, (int)n_src) + mid_bytes(d->n_in, d->n_out) +
                          8 * (size_t)(d->n_in + d->n_out) + 8192;
  return (int)balanced_chunk((size_t)d->batch, kChunkBudget / per_item);
}

struct MftScratch {
  float *xin, *uout;
  PlaneSet in_pl, mid_pl;
  int* sync_ws;     // 2 ints per item of a chunk: the fused launch's ready / consumed counters
};

static size_t carve_mft(const dlux_mft_desc* d, void* scratch, size_t cap, MftScratch* s, bool* ok) {
  Bump b(scratch, cap);
  const size_t c = mft_chunk(d);
  const size_t n_src = d->adjoint ? d->n_out : d->n_in;
  s->xin = b.take<float>(c * 2 * d->n_in);
  s->uout = b.take<float>(c * 2 * d->n_out);
  s->in_pl = take_planes(b, c * n_src, (int)n_src);
  s->mid_pl = take_mid_planes(b, c, d->n_in, d->n_out);
  s->sync_ws = b.take<int>(2 * c);
  b.take<float>(gemm_tc_workspace_bytes() / sizeof(float) + 1);
  if (ok) *ok = b.ok;
  return b.used();
}

static int check_mft_desc(const dlux_mft_desc* d) {
  if (!d) return DLUX_ERR_ARG;
  if (d->n_in < 1 || d->n_out < 1 || d->batch < 0) return DLUX_ERR_SHAPE;
  if (d->n_in > 32768 || d->n_out > 32768) return DLUX_ERR_SHAPE;
  if (d->precision != DLUX_PREC_3XTF32 && d->precision != DLUX_PREC_FP32) return DLUX_ERR_ARG;
  if (d->dft_period < 0) return DLUX_ERR_ARG;
  // exact-DFT mode: the integer products (j - j0)(b - b0) must be exact in float32
  if (d->dft_period > 0 && (double)(d->n_in + d->dft_period) * (double)(d->n_out + d->dft_period) >= 16777216.0)
    return DLUX_ERR_SHAPE;
  return DLUX_OK;
}

// The two stages of one (chunk of) MFT(s); fwd: data [c][N][N] -> out [c][M][M];
// adjoint: data [c][M][M] -> out [c][N][N].
static void fill_stage(GemmParams& g, bool adjoint, int stage, int N, int M, int c,
                       const float* xin, const float* uout, float sign2pi) {
  // axis 0 = x (contracts the column index j / b), axis 1 = y (row index i / a)
  const int axis = stage == 0 ? 0 : 1;
  g.n_items = c;
  g.n_data = c;
  g.item_data = nullptr;
  g.sign2pi = sign2pi;
  if (!adjoint) {
    g.K = N;
    g.n_out = M;
    g.rows = stage == 0 ? N : M;
    g.kvec = xin + (size_t)axis * N;
    g.kvec_stride = 2 * N;
    g.nvec = uout + (size_t)axis * M;
    g.nvec_stride = 2 * M;
  } else {
    g.K = M;
    g.n_out = N;
    g.rows = stage == 0 ? M : N;
    g.kvec = uout + (size_t)axis * M;
    g.kvec_stride = 2 * M;
    g.nvec = xin + (size_t)axis * N;
    g.nvec_stride = 2 * N;
  }
}


// ------------------------------------------------------------------ parameter-batched poly-PSF
// item = b * L + l; one chunk = a whole number of batch elements
struct BatchScratch {
  float* amp_scale;
  float *s_item, *norm_item, *k_item, *delta_item;   // per chunk [cb * L]
  int* item_l;
  int* sync_ws;       // fused launch counters, 2 per item of a chunk
  float *xin, *uout;
  float* opd_c;       // [cb][N*N]
  float* opdbar_c;    // [cb][N*N]
  PlaneSet p_pl;      // [cb * L][N][N]
  PlaneSet mid_pl;    // [cb * L]
  PlaneSet ebar_pl;   // [cb * L][M][M]
  float2* qbuf;       // [cb * L][N*N]
  float2* fbuf;       // [cb * L][M*M]
  int cb;
};

static int batch_chunk(const dlux_polypsf_batch_desc* d) {
  const size_t N = d->n_pupil, M = d->n_psf, L = d->n_wavels;
  const size_t per_item = plane_bytes(N, (int)N) + mid_bytes((int)N, (int)M) + plane_bytes(M, (int)M) +
                          8 * (N + M) + 8 * N * N + 8 * M * M + 64;
  const size_t per_b = L * per_item + 8 * N * N + 4096;
  size_t cb = kChunkBudget / per_b;
  if (cb < 1) cb = 1;
  if (cb * L > 32768) cb = 32768 / L > 0 ? 32768 / L : 1;     // grid.y limits of the per-item helper kernels
  return (int)balanced_chunk((size_t)d->n_batch, cb);
}

static size_t carve_batch(const dlux_polypsf_batch_desc* d, void* scratch, size_t cap, BatchScratch* s, bool* ok) {
  Bump b(scratch, cap);
  const size_t N = d->n_pupil, M = d->n_psf, L = d->n_wavels;
  const size_t cb = batch_chunk(d), c = cb * L;
  s->cb = (int)cb;
  s->amp_scale = b.take<float>(4 + 512);
  s->s_item = b.take<float>(c);
  s->norm_item = b.take<float>(c);
  s->k_item = b.take<float>(c);
  s->delta_item = b.take<float>(2 * c);
  s->item_l = b.take<int>(c);
  s->sync_ws = b.take<int>(2 * c);
  s->xin = b.take<float>(c * 2 * N);
  s->uout = b.take<float>(c * 2 * M);
  s->opd_c = b.take<float>(cb * N * N);
  s->opdbar_c = b.take<float>(cb * N * N);
  s->p_pl = take_planes(b, c * N, (int)N);
  s->mid_pl = take_mid_planes(b, c, (int)N, (int)M);
  s->ebar_pl = take_planes(b, c * M, (int)M);
  s->qbuf = b.take<float2>(c * N * N);
  s->fbuf = b.take<float2>(c * M * M);
  if (ok) *ok = b.ok;
  return b.used();
}

static int check_batch_desc(const dlux_polypsf_batch_desc* d) {
  if (!d) return DLUX_ERR_ARG;
  if (d->n_pupil < 1 || d->n_psf < 1 || d->n_wavels < 1 || d->n_batch < 1 || d->n_basis < 1) return DLUX_ERR_SHAPE;
  if (d->n_pupil > 32768 || d->n_psf > 32768 || d->n_wavels > 32768 || d->n_basis > 8192) return DLUX_ERR_SHAPE;
  if (d->precision != DLUX_PREC_3XTF32 && d->precision != DLUX_PREC_FP32) return DLUX_ERR_ARG;
  return DLUX_OK;
}

__global__ void expand_batch_kernel(int n_items, int L, const float* __restrict__ scale_out,
                                    const float* __restrict__ norm, const float* __restrict__ wavenumber,
                                    float* __restrict__ s_item, float* __restrict__ norm_item,
                                    float* __restrict__ k_item) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += gridDim.x * blockDim.x) {
    const int l = i % L;     // batch-major: item = b * L + l
    s_item[i] = scale_out[l];
    norm_item[i] = norm ? norm[l] : 1.0f;
    k_item[i] = wavenumber[l];
  }
}

__global__ void expand_delta_kernel(int n_items, int L, const float* __restrict__ delta_l, float* __restrict__ delta_item) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n_items; i += gridDim.x * blockDim.x)
    delta_item[i] = delta_l[2 * ((i / 2) % L) + (i & 1)];
}

// per-chunk geometry operands (the same for every chunk: items repeat the L wavelengths)
static int batch_prologue(const dlux_polypsf_batch_desc* d, const BatchScratch& s, const float* T,
                          const float* wavenumber, const float* scale_out, const float* norm, const float* delta_xy,
                          cudaStream_t st) {
  const int N = d->n_pupil, M = d->n_psf, L = d->n_wavels;
  const int c = s.cb * L;
  int rc = launch_power(N, T, d->normalise, s.amp_scale, st);
  if (rc) return rc;
  expand_batch_kernel<<<(c + 255) / 256, 256, 0, st>>>(c, L, scale_out, norm, wavenumber, s.s_item, s.norm_item,
                                                        s.k_item);
  note_launch();
  if (delta_xy) {
    expand_delta_kernel<<<(2 * c + 255) / 256, 256, 0, st>>>(c, L, delta_xy, s.delta_item);
    note_launch();
  }
  rc = check_launch("batch expand");
  if (rc) return rc;
  return launch_coords(N, M, c, s.s_item, nullptr, delta_xy ? s.delta_item : nullptr, 1, s.xin, s.uout, st);
}

}  // namespace dlux

using namespace dlux;

extern "C" {

int dlux_abi_version(void) { return DLUX_B200_ABI_VERSION; }

const char* dlux_error_string(int code) {
  switch (code) {
    case DLUX_OK: return "ok";
    case DLUX_ERR_ARG: return "invalid argument";
    case DLUX_ERR_SHAPE: return "unsupported shape";
    case DLUX_ERR_SCRATCH: return "scratch buffer too small";
    case DLUX_ERR_CUDA: return "CUDA error";
    case DLUX_ERR_UNSUPPORTED: return "unsupported on this device";
    case DLUX_ERR_ALIGN: return "pointer not 16-byte aligned";
    default: return "unknown error";
  }
}

int dlux_last_cuda_error(void) { return g_last_cuda_error; }
uint64_t dlux_launch_count(void) { return g_launches.load(); }

int dlux_profile_enable(int on) {
  g_prof_on.store(on ? 1 : 0);
  return DLUX_OK;
}

int dlux_profile_read(double* gemm_ms, uint64_t* gemm_launches, double* gemm_flops) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  double ms = 0.0, fl = 0.0;
  for (auto& r : g_prof) {
    float t = 0.f;
    cudaEventSynchronize(r.b);
    cudaEventElapsedTime(&t, r.a, r.b);
    ms += t;
    fl += r.flops;
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  if (gemm_ms) *gemm_ms = ms;
  if (gemm_launches) *gemm_launches = g_prof.size();
  if (gemm_flops) *gemm_flops = fl;
  g_prof.clear();
  return DLUX_OK;
}

int dlux_tc_peak_probe(int32_t kind, int32_t n_batches, float* sink, double* flops_host, void* cuda_stream) {
  return launch_tc_peak_probe(kind, n_batches, sink, flops_host, (cudaStream_t)cuda_stream);
}

size_t dlux_mft_scratch_bytes(const dlux_mft_desc* desc) {
  if (check_mft_desc(desc) != DLUX_OK) return 0;
  MftScratch s;
  return carve_mft(desc, nullptr, 0, &s, nullptr);
}

int dlux_mft_coords(int32_t n_in, int32_t n_out, int32_t batch, const float* scale_out,
                    const float* shift_xy, const float* delta_xy, float* xin, float* uout,
                    void* cuda_stream) {
  if (!scale_out || !xin || !uout) return DLUX_ERR_ARG;
  if (n_in < 1 || n_out < 1 || batch < 0) return DLUX_ERR_SHAPE;
  return launch_coords(n_in, n_out, batch, scale_out, shift_xy, delta_xy, 1, xin, uout,
                       (cudaStream_t)cuda_stream);
}

int dlux_mft_c64(const dlux_mft_desc* d, const void* in, const float* scale_out,
                 const float* shift_xy, const float* delta_xy, const float* norm, void* out,
                 void* scratch, size_t scratch_bytes, void* cuda_stream) {
  int rc = check_mft_desc(d);
  if (rc != DLUX_OK) return rc;
  if (!in || !out || !scale_out || !scratch) return DLUX_ERR_ARG;
  // `in` is read as float2 (an odd-N slice of a batched phasor is only 8-byte aligned); `out`
  // and `scratch` are TMA targets
  if (((uintptr_t)out | (uintptr_t)scratch) & 15) return DLUX_ERR_ALIGN;
  if ((uintptr_t)in & 7) return DLUX_ERR_ALIGN;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  MftScratch s;
  bool ok = true;
  carve_mft(d, scratch, scratch_bytes, &s, &ok);
  if (!ok) return DLUX_ERR_SCRATCH;
  const int N = d->n_in, M = d->n_out;
  const bool adj = d->adjoint != 0;
  const size_t n_src = adj ? M : N, n_dst = adj ? N : M;
  // forward exponent sign is -2 pi (propagation.py:124), flipped by inverse (:125-126);
  // the adjoint conjugates the phasors.
  float sign2pi = (float)(-2.0 * 3.14159265358979323846);
  if (d->inverse) sign2pi = -sign2pi;
  if (adj) sign2pi = -sign2pi;
  const int chunk = mft_chunk(d);
  for (int b0 = 0; b0 < d->batch; b0 += chunk) {
    const int c = d->batch - b0 < chunk ? d->batch - b0 : chunk;
    rc = launch_coords(N, M, c, scale_out + b0, shift_xy ? shift_xy + 2 * (size_t)b0 : nullptr,
                       delta_xy ? delta_xy + 2 * (size_t)b0 : nullptr, 1, s.xin, s.uout, st, d->dft_period > 0);
    if (rc) return rc;
    rc = launch_split_c64((const float2*)in + (size_t)b0 * n_src * n_src, (size_t)c * n_src, (int)n_src,
                          s.in_pl, st);
    if (rc) return rc;
    GemmParams g{};
    fill_stage(g, adj, 0, N, M, c, s.xin, s.uout, sign2pi);
    g.a = s.in_pl;
    g.out = s.mid_pl;
    g.mode = EPI_PLANES;
    g.scale = nullptr;
    g.dft_period = (float)d->dft_period;
    GemmParams h{};
    fill_stage(h, adj, 1, N, M, c, s.xin, s.uout, sign2pi);
    h.a = s.mid_pl;
    h.mode = EPI_C64;
    h.scale = norm ? norm + b0 : nullptr;
    h.dft_period = (float)d->dft_period;
    h.out_c64 = (float2*)out + (size_t)b0 * n_dst * n_dst;
    rc = run_gemm_pair(g, h, s.sync_ws, d->precision, st);
    if (rc) return rc;
  }
  return DLUX_OK;
}

// ------------------------------------------------------------------ fused poly-PSF
struct PolyScratch {
  float* amp_scale;  // 16 B + 256 doubles
  PlaneSet p_pl;     // [L][N][N]
  int* item_l;
  float *s_item, *norm_item, *k_item;
  float *w_item, *delta_item;                          // caller's [S, L] arrays in processing order
  float *wbar_item, *dbar_item, *sbar_item, *kbar_item; // bwd outputs in processing order
  int *sp_flags, *sp_cnt, *sp_idx;                      // zero-block lists (sparse option)
  int* sync_ws;                                         // fused launch counters, 2 per item of a chunk
  float *xin, *uout;  // per chunk
  PlaneSet mid_pl;    // per chunk [c][M*N]
  PlaneSet ebar_pl;   // per chunk [c][M*M]  (bwd only; sized always for simplicity)
  float2* qbuf;       // per chunk [c][N*N]: adjoint field Q per item (bwd)
  float2* fbuf;       // per chunk [c][M*M]: field, when the caller does not keep it (fwd)
  int chunk;
};

static int poly_chunk(const dlux_polypsf_desc* d) {
  const size_t N = d->n_pupil, M = d->n_psf;
  const size_t per_item = mid_bytes((int)N, (int)M) + plane_bytes(M, (int)M) + 8 * (N + M) +
                          8 * N * N + 8 * M * M + 16384;
  const size_t items = (size_t)d->n_sources * d->n_wavels;
  return (int)balanced_chunk(items, kChunkBudget / per_item);
}

static size_t carve_poly(const dlux_polypsf_desc* d, void* scratch, size_t cap, PolyScratch* s, bool* ok) {
  Bump b(scratch, cap);
  const size_t N = d->n_pupil, M = d->n_psf, L = d->n_wavels;
  const size_t items = (size_t)d->n_sources * L;
  const size_t c = poly_chunk(d);
  s->chunk = (int)c;
  s->amp_scale = b.take<float>(4 + 512);
  s->p_pl = take_planes(b, L * N, (int)N);
  s->item_l = b.take<int>(items);
  s->s_item = b.take<float>(items);
  s->norm_item = b.take<float>(items);
  s->k_item = b.take<float>(items);
  s->w_item = b.take<float>(items);
  s->delta_item = b.take<float>(2 * items);
  s->wbar_item = b.take<float>(items);
  s->dbar_item = b.take<float>(2 * items);
  s->sbar_item = b.take<float>(items);
  s->kbar_item = b.take<float>(items);
  {
    const size_t nblk = ((N + 127) / 128) * ((N + 15) / 16) + 64;   // >= both block grids used
    s->sp_flags = b.take<int>(nblk);
    s->sp_cnt = b.take<int>((N + 127) / 128 + 8);
    s->sp_idx = b.take<int>(nblk);
  }
  s->xin = b.take<float>(c * 2 * N);
  s->uout = b.take<float>(c * 2 * M);
  s->sync_ws = b.take<int>(2 * c);
  s->mid_pl = take_mid_planes(b, c, (int)N, (int)M);
  s->ebar_pl = take_planes(b, c * M, (int)M);
  s->qbuf = b.take<float2>(c * N * N);
  s->fbuf = b.take<float2>(c * M * M);
  b.take<float>(gemm_tc_workspace_bytes() / sizeof(float) + 1);
  if (ok) *ok = b.ok;
  return b.used();
}

static int check_poly_desc(const dlux_polypsf_desc* d) {
  if (!d) return DLUX_ERR_ARG;
  if (d->n_pupil < 1 || d->n_psf < 1 || d->n_wavels < 1 || d->n_sources < 1) return DLUX_ERR_SHAPE;
  if (d->n_pupil > 32768 || d->n_psf > 32768) return DLUX_ERR_SHAPE;
  if ((long long)d->n_wavels * d->n_sources > (1LL << 30)) return DLUX_ERR_SHAPE;
  if (d->precision != DLUX_PREC_3XTF32 && d->precision != DLUX_PREC_FP32) return DLUX_ERR_ARG;
  return DLUX_OK;
}

size_t dlux_polypsf_scratch_bytes(const dlux_polypsf_desc* desc) {
  if (check_poly_desc(desc) != DLUX_OK) return 0;
  PolyScratch s;
  return carve_poly(desc, nullptr, 0, &s, nullptr);
}

static int poly_prologue(const dlux_polypsf_desc* d, const PolyScratch& s, const float* T,
                         const float* opd, const float* phase, const float* wavenumber,
                         const float* scale_out, const float* norm, const float* weights, const float* delta_xy,
                         bool need_planes, cudaStream_t st) {
  const int N = d->n_pupil, L = d->n_wavels;
  const int items = d->n_sources * L;
  int rc = launch_power(N, T, d->normalise, s.amp_scale, st);
  if (rc) return rc;
  if (need_planes) {
    rc = launch_pupil(N, L, T, opd, phase, wavenumber, s.amp_scale, s.p_pl, st);
    if (rc) return rc;
  }
  expand_items_kernel<<<(items + 255) / 256 > 1024 ? 1024 : (items + 255) / 256, 256, 0, st>>>(
      items, L, d->n_sources, scale_out, norm, wavenumber, weights, delta_xy, s.item_l, s.s_item, s.norm_item,
      s.k_item, s.w_item, s.delta_item);
  note_launch();
  return check_launch("expand_items");
}

int dlux_polypsf_fwd(const dlux_polypsf_desc* d, const float* T, const float* opd,
                     const float* phase, const float* wavenumber, const float* scale_out,
                     const float* norm, const float* weights, const float* delta_xy, float* psf,
                     void* field, void* scratch, size_t scratch_bytes, void* cuda_stream) {
  int rc = check_poly_desc(d);
  if (rc != DLUX_OK) return rc;
  if (!wavenumber || !scale_out || !weights || !psf || !scratch) return DLUX_ERR_ARG;
  if (d->save_field && !field) return DLUX_ERR_ARG;
  if (((uintptr_t)psf | (uintptr_t)scratch | (uintptr_t)field) & 15) return DLUX_ERR_ALIGN;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  PolyScratch s;
  bool ok = true;
  carve_poly(d, scratch, scratch_bytes, &s, &ok);
  if (!ok) return DLUX_ERR_SCRATCH;
  const int N = d->n_pupil, M = d->n_psf, L = d->n_wavels;
  const int items = d->n_sources * L;
  rc = poly_prologue(d, s, T, opd, phase, wavenumber, scale_out, norm, weights, delta_xy, true, st);
  if (rc) return rc;
  const float sign2pi = (float)(-2.0 * 3.14159265358979323846);
  // Forward-only calls (no VJP residual wanted) fuse |E|^2, the spectral weights and the sum over sources
  // and wavelengths into the last contraction's epilogue (EPI_PSF): no complex field is written or re-read.
  // (The tensor kernel's reduce-add stores need 16-byte image rows.)
  const bool fuse_psf = !d->save_field && (M % 4) == 0 && getenv("DLUX_B200_NO_EPI_PSF") == nullptr;
  if (fuse_psf && (rc = launch_zero(psf, (size_t)M * M, st))) return rc;
  // opt-in exact zero-block skipping: stage 1 contracts P[i][j] over j; a (256 rows x 16 k) block of P is
  // all zeros wherever T is, and contributes exact zeros to every partial sum
  const bool sparse = d->sparse && T && d->precision == DLUX_PREC_3XTF32;
  if (sparse && (rc = launch_block_lists(N, T, GEMM_TC_BM2, GEMM_TC_BK, 0, s.sp_flags, s.sp_cnt, s.sp_idx, st)))
    return rc;
  for (int b0 = 0; b0 < items; b0 += s.chunk) {
    const int c = items - b0 < s.chunk ? items - b0 : s.chunk;
    rc = launch_coords(N, M, c, s.s_item + b0, nullptr, delta_xy ? s.delta_item + 2 * (size_t)b0 : nullptr,
                       1, s.xin, s.uout, st);
    if (rc) return rc;
    GemmParams g{};
    fill_stage(g, false, 0, N, M, c, s.xin, s.uout, sign2pi);
    g.item_data = s.item_l + b0;
    g.n_data = L;
    g.a = s.p_pl;
    g.out = s.mid_pl;
    g.mode = EPI_PLANES;
    if (sparse) {
      g.chunk_cnt = s.sp_cnt;
      g.chunk_idx = s.sp_idx;
    }
    GemmParams h{};
    fill_stage(h, false, 1, N, M, c, s.xin, s.uout, sign2pi);
    h.a = s.mid_pl;
    h.scale = s.norm_item + b0;
    if (fuse_psf) {
      h.mode = EPI_PSF;
      h.out_psf = psf;
      h.item_w = s.w_item + b0;
      rc = run_gemm_pair(g, h, s.sync_ws, d->precision, st);
      if (rc) return rc;
      continue;
    }
    h.mode = EPI_C64;
    h.out_c64 = d->save_field ? (float2*)field + (size_t)b0 * M * M : s.fbuf;
    rc = run_gemm_pair(g, h, s.sync_ws, d->precision, st);
    if (rc) return rc;
    // psf (+)= sum over this chunk's (source, wavelength) items of w |E|^2
    rc = launch_psf_reduce((size_t)M * M, c, h.out_c64, s.w_item + b0, psf, b0 > 0, st);
    if (rc) return rc;
  }
  return DLUX_OK;
}

int dlux_polypsf_bwd(const dlux_polypsf_desc* d, const float* T, const float* opd,
                     const float* phase, const float* wavenumber, const float* scale_out,
                     const float* norm, const float* weights, const float* delta_xy,
                     const void* field, const float* psf_bar, float* opd_bar, float* phase_bar,
                     float* weights_bar, float* delta_bar, float* transmission_bar, float* scale_bar,
                     float* wavenumber_bar, void* scratch, size_t scratch_bytes, void* cuda_stream) {
  int rc = check_poly_desc(d);
  if (rc != DLUX_OK) return rc;
  if (!wavenumber || !scale_out || !weights || !field || !psf_bar || !scratch) return DLUX_ERR_ARG;
  if (((uintptr_t)field | (uintptr_t)scratch) & 15) return DLUX_ERR_ALIGN;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  PolyScratch s;
  bool ok = true;
  carve_poly(d, scratch, scratch_bytes, &s, &ok);
  if (!ok) return DLUX_ERR_SCRATCH;
  const int N = d->n_pupil, M = d->n_psf, L = d->n_wavels;
  const int items = d->n_sources * L;
  // the gradient epilogue re-evaluates the pupil phasor itself: no operand planes needed
  rc = poly_prologue(d, s, T, opd, phase, wavenumber, scale_out, norm, weights, delta_xy, false, st);
  if (rc) return rc;
  // per-item outputs are accumulated in processing order and scattered to the caller's [S, L] layout at the end
  float* const wbar_c = weights_bar, * const dbar_c = delta_bar, * const sbar_c = scale_bar, * const kbar_c = wavenumber_bar;
  if (weights_bar) weights_bar = s.wbar_item;
  if (delta_bar) delta_bar = s.dbar_item;
  if (scale_bar) scale_bar = s.sbar_item;
  if (wavenumber_bar) wavenumber_bar = s.kbar_item;
  if (weights_bar && (rc = launch_zero(weights_bar, (size_t)items, st))) return rc;
  if (delta_bar && (rc = launch_zero(delta_bar, (size_t)items * 2, st))) return rc;
  if (scale_bar && (rc = launch_zero(scale_bar, (size_t)items, st))) return rc;
  if (wavenumber_bar && (rc = launch_zero(wavenumber_bar, (size_t)items, st))) return rc;
  if (transmission_bar && !T) return DLUX_ERR_ARG;
  const bool need_pupil_grad =
      opd_bar || phase_bar || delta_bar || transmission_bar || scale_bar || wavenumber_bar;
  // opt-in exact zero-block skipping: the last adjoint stage writes Q[i][j]; every pupil-plane cotangent
  // except the transmission's ignores the pixels with T == 0, so (128 x 256) output blocks that are entirely
  // blocked are never computed
  const bool sparse = d->sparse && T && !transmission_bar && d->precision == DLUX_PREC_3XTF32 && need_pupil_grad;
  if (sparse && (rc = launch_block_lists(N, T, GEMM_TC_BN2, GEMM_TC_BM2, 1, s.sp_flags, s.sp_cnt, s.sp_idx, st)))
    return rc;
  const float sign2pi = (float)(2.0 * 3.14159265358979323846);  // conj of the forward phasors
  for (int b0 = 0; b0 < items; b0 += s.chunk) {
    const int c = items - b0 < s.chunk ? items - b0 : s.chunk;
    // pass 0: Q = adjoint(Ebar) -> opd / phase / transmission / offset gradients.  Passes 1, 2
    // (scale_bar only): the output coordinate is u_a = scale_out (a - (M-1)/2) - delta, so
    // d/d scale_out weighs Ebar with the pixel index along x, then along y, and reduces the
    // adjoint like the offset gradient (opposite sign).
    const int n_pass = scale_bar ? 3 : 1;
    for (int pass = 0; pass < n_pass; ++pass) {
      rc = launch_cotangent(M, c, (const float2*)field + (size_t)b0 * M * M, psf_bar, s.w_item + b0,
                            s.ebar_pl, (pass == 0 && weights_bar) ? weights_bar + b0 : nullptr,
                            pass - 1, st);
      if (rc) return rc;
      if (!need_pupil_grad) continue;
      if (pass == 0) {
        rc = launch_coords(N, M, c, s.s_item + b0, nullptr, delta_xy ? s.delta_item + 2 * (size_t)b0 : nullptr,
                           1, s.xin, s.uout, st);
        if (rc) return rc;
      }
      GemmParams g{};
      fill_stage(g, true, 0, N, M, c, s.xin, s.uout, sign2pi);
      g.a = s.ebar_pl;
      g.out = s.mid_pl;
      g.mode = EPI_PLANES;
      GemmParams h{};
      fill_stage(h, true, 1, N, M, c, s.xin, s.uout, sign2pi);
      h.a = s.mid_pl;
      h.mode = EPI_C64;
      h.scale = s.norm_item + b0;
      h.out_c64 = s.qbuf;
      if (sparse) {
        h.unit_list = s.sp_idx;
        h.unit_count = s.sp_cnt;
      }
      rc = run_gemm_pair(g, h, s.sync_ws, d->precision, st);
      if (rc) return rc;
      const float a0 = 1.0f / (float)((long long)N * N);
      if (pass > 0) {
        rc = launch_pos_grad(N, c, s.qbuf, s.k_item + b0, T, opd, phase, s.amp_scale, a0, scale_bar + b0,
                             pass, st);
        if (rc) return rc;
        continue;
      }
      // every pupil-plane cotangent of pass 0 from ONE read of Q: opd / phase / transmission gradients summed
      // over the items, source-offset and wavenumber gradients per item
      if (opd_bar || phase_bar || transmission_bar || delta_bar || (wavenumber_bar && opd)) {
        rc = launch_q_reduce(N, c, s.qbuf, s.k_item + b0, T, opd, phase, s.amp_scale, a0, opd_bar, phase_bar,
                             transmission_bar, b0 > 0, delta_bar ? delta_bar + 2 * (size_t)b0 : nullptr,
                             (wavenumber_bar && opd) ? wavenumber_bar + b0 : nullptr, st);
        if (rc) return rc;
      }
    }
  }
  {
    const int S = d->n_sources, g = (items * 2 + 255) / 256 > 1024 ? 1024 : (items * 2 + 255) / 256;
    if (wbar_c) { scatter_items_kernel<<<g, 256, 0, st>>>(items, L, S, 1, s.wbar_item, wbar_c); note_launch(); }
    if (dbar_c) { scatter_items_kernel<<<g, 256, 0, st>>>(items, L, S, 2, s.dbar_item, dbar_c); note_launch(); }
    if (sbar_c) { scatter_items_kernel<<<g, 256, 0, st>>>(items, L, S, 1, s.sbar_item, sbar_c); note_launch(); }
    if (kbar_c) { scatter_items_kernel<<<g, 256, 0, st>>>(items, L, S, 1, s.kbar_item, kbar_c); note_launch(); }
    if ((rc = check_launch("scatter_items"))) return rc;
  }
  if (transmission_bar && d->normalise) {
    // amp_scale's scratch slot is followed by the 256-double work area of the power reduction
    double* work = reinterpret_cast<double*>(reinterpret_cast<char*>(s.amp_scale) + 16);
    rc = launch_tbar_finalize((size_t)N * N, T, s.amp_scale, 1.0f / (float)((long long)N * N),
                              transmission_bar, work, st);
    if (rc) return rc;
  }
  return DLUX_OK;
}



int dlux_polypsf_hvp(const dlux_polypsf_desc* d, const float* T, const float* opd, const float* phase,
                     const float* wavenumber, const float* scale_out, const float* norm, const float* weights,
                     const float* delta_xy, const void* field, const float* psf_bar, const float* opd_tangent,
                     float* psf_tan, float* opd_hv, void* scratch, size_t scratch_bytes, void* cuda_stream) {
  int rc = check_poly_desc(d);
  if (rc != DLUX_OK) return rc;
  if (!wavenumber || !scale_out || !weights || !field || !psf_bar || !opd_tangent || !scratch) return DLUX_ERR_ARG;
  if (!psf_tan && !opd_hv) return DLUX_ERR_ARG;
  if (((uintptr_t)field | (uintptr_t)scratch) & 15) return DLUX_ERR_ALIGN;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  PolyScratch s;
  bool ok = true;
  carve_poly(d, scratch, scratch_bytes, &s, &ok);
  if (!ok) return DLUX_ERR_SCRATCH;
  const int N = d->n_pupil, M = d->n_psf, L = d->n_wavels;
  const int items = d->n_sources * L;
  const size_t npix = (size_t)N * N, mpix = (size_t)M * M;
  rc = poly_prologue(d, s, T, opd, phase, wavenumber, scale_out, norm, weights, delta_xy, false, st);
  if (rc) return rc;
  // tangent pupil planes dP_l = i k_l V P_l (they take the place of the P planes)
  rc = launch_pupil(N, L, T, opd, phase, wavenumber, s.amp_scale, s.p_pl, st, 1, opd_tangent);
  if (rc) return rc;
  const float fwd2pi = (float)(-2.0 * 3.14159265358979323846), adj2pi = -fwd2pi;
  const float a0 = 1.0f / (float)((long long)N * N);
  for (int b0 = 0; b0 < items; b0 += s.chunk) {
    const int c = items - b0 < s.chunk ? items - b0 : s.chunk;
    const float2* E = (const float2*)field + (size_t)b0 * mpix;
    rc = launch_coords(N, M, c, s.s_item + b0, nullptr, delta_xy ? s.delta_item + 2 * (size_t)b0 : nullptr, 1, s.xin,
                       s.uout, st);
    if (rc) return rc;
    // dE = MFT(dP) per item
    GemmParams g{};
    fill_stage(g, false, 0, N, M, c, s.xin, s.uout, fwd2pi);
    g.item_data = s.item_l + b0;
    g.n_data = L;
    g.a = s.p_pl;
    g.out = s.mid_pl;
    g.mode = EPI_PLANES;
    GemmParams h{};
    fill_stage(h, false, 1, N, M, c, s.xin, s.uout, fwd2pi);
    h.a = s.mid_pl;
    h.mode = EPI_C64;
    h.scale = s.norm_item + b0;
    h.out_c64 = s.fbuf;
    rc = run_gemm_pair(g, h, s.sync_ws, d->precision, st);
    if (rc) return rc;
    if (psf_tan) {
      rc = launch_psf_tangent(mpix, c, E, s.fbuf, s.w_item + b0, psf_tan, b0 > 0, st);
      if (rc) return rc;
    }
    if (!opd_hv) continue;
    // two adjoint passes: Q' = A^H(2 w G dE) and Q = A^H(2 w G E)
    for (int pass = 0; pass < 2; ++pass) {
      rc = launch_cotangent(M, c, pass == 0 ? s.fbuf : E, psf_bar, s.w_item + b0, s.ebar_pl, nullptr, -1, st);
      if (rc) return rc;
      GemmParams ga{};
      fill_stage(ga, true, 0, N, M, c, s.xin, s.uout, adj2pi);
      ga.a = s.ebar_pl;
      ga.out = s.mid_pl;
      ga.mode = EPI_PLANES;
      GemmParams ha{};
      fill_stage(ha, true, 1, N, M, c, s.xin, s.uout, adj2pi);
      ha.a = s.mid_pl;
      ha.mode = EPI_C64;
      ha.scale = s.norm_item + b0;
      ha.out_c64 = s.qbuf;
      rc = run_gemm_pair(ga, ha, s.sync_ws, d->precision, st);
      if (rc) return rc;
      rc = launch_hv_reduce(npix, c, s.qbuf, s.k_item + b0, T, opd, phase, s.amp_scale, a0, opd_tangent, opd_hv, pass,
                            (b0 > 0 || pass > 0) ? 1 : 0, st);
      if (rc) return rc;
    }
  }
  return DLUX_OK;
}

size_t dlux_polypsf_batch_scratch_bytes(const dlux_polypsf_batch_desc* desc) {
  if (check_batch_desc(desc) != DLUX_OK) return 0;
  BatchScratch s;
  return carve_batch(desc, nullptr, 0, &s, nullptr);
}

int dlux_polypsf_batch_fwd(const dlux_polypsf_batch_desc* d, const float* T, const float* base_opd,
                           const float* phase, const float* basis, const float* coeffs, const float* wavenumber,
                           const float* scale_out, const float* norm, const float* weights, const float* delta_xy,
                           float* psf, void* field, void* scratch, size_t scratch_bytes, void* cuda_stream) {
  int rc = check_batch_desc(d);
  if (rc != DLUX_OK) return rc;
  if (!basis || !coeffs || !wavenumber || !scale_out || !weights || !psf || !scratch) return DLUX_ERR_ARG;
  if (d->save_field && !field) return DLUX_ERR_ARG;
  if (((uintptr_t)psf | (uintptr_t)scratch | (uintptr_t)field) & 15) return DLUX_ERR_ALIGN;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  BatchScratch s;
  bool ok = true;
  carve_batch(d, scratch, scratch_bytes, &s, &ok);
  if (!ok) return DLUX_ERR_SCRATCH;
  const int N = d->n_pupil, M = d->n_psf, L = d->n_wavels, B = d->n_batch, nz = d->n_basis;
  const size_t npix = (size_t)N * N, mpix = (size_t)M * M;
  rc = batch_prologue(d, s, T, wavenumber, scale_out, norm, delta_xy, st);
  if (rc) return rc;
  const float sign2pi = (float)(-2.0 * 3.14159265358979323846);
  for (int b0 = 0; b0 < B; b0 += s.cb) {
    const int cb = B - b0 < s.cb ? B - b0 : s.cb;
    const int c = cb * L;
    // OPD of every batch element of the chunk (utils/math.py:177-196), then its L pupil phasors
    rc = launch_basis_eval(nz, (int64_t)npix, basis, coeffs + (size_t)b0 * nz, base_opd, s.opd_c, st, cb);
    if (rc) return rc;
    rc = launch_pupil(N, L, T, s.opd_c, phase, wavenumber, s.amp_scale, s.p_pl, st, cb);
    if (rc) return rc;
    GemmParams g{};
    fill_stage(g, false, 0, N, M, c, s.xin, s.uout, sign2pi);
    g.a = s.p_pl;
    g.out = s.mid_pl;
    g.mode = EPI_PLANES;
    GemmParams h{};
    fill_stage(h, false, 1, N, M, c, s.xin, s.uout, sign2pi);
    h.a = s.mid_pl;
    h.mode = EPI_C64;
    h.scale = s.norm_item;
    h.out_c64 = d->save_field ? (float2*)field + (size_t)b0 * L * mpix : s.fbuf;
    rc = run_gemm_pair(g, h, s.sync_ws, d->precision, st);
    if (rc) return rc;
    // psf[b] = sum_l w_l |E_bl|^2: per-item images, no collective, nothing summed over the batch
    rc = launch_psf_reduce(mpix, L, h.out_c64, weights, psf + (size_t)b0 * mpix, 0, st, cb);
    if (rc) return rc;
  }
  return DLUX_OK;
}

int dlux_polypsf_batch_bwd(const dlux_polypsf_batch_desc* d, const float* T, const float* base_opd,
                           const float* phase, const float* basis, const float* coeffs, const float* wavenumber,
                           const float* scale_out, const float* norm, const float* weights, const float* delta_xy,
                           const void* field, const float* psf_bar, float* coeff_bar, void* scratch,
                           size_t scratch_bytes, void* cuda_stream) {
  int rc = check_batch_desc(d);
  if (rc != DLUX_OK) return rc;
  if (!basis || !coeffs || !wavenumber || !scale_out || !weights || !field || !psf_bar || !coeff_bar || !scratch)
    return DLUX_ERR_ARG;
  if (((uintptr_t)field | (uintptr_t)scratch) & 15) return DLUX_ERR_ALIGN;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  BatchScratch s;
  bool ok = true;
  carve_batch(d, scratch, scratch_bytes, &s, &ok);
  if (!ok) return DLUX_ERR_SCRATCH;
  const int N = d->n_pupil, M = d->n_psf, L = d->n_wavels, B = d->n_batch, nz = d->n_basis;
  const size_t npix = (size_t)N * N, mpix = (size_t)M * M;
  rc = batch_prologue(d, s, T, wavenumber, scale_out, norm, delta_xy, st);
  if (rc) return rc;
  const float sign2pi = (float)(2.0 * 3.14159265358979323846);  // conj of the forward phasors
  const float a0 = 1.0f / (float)((long long)N * N);
  for (int b0 = 0; b0 < B; b0 += s.cb) {
    const int cb = B - b0 < s.cb ? B - b0 : s.cb;
    const int c = cb * L;
    rc = launch_basis_eval(nz, (int64_t)npix, basis, coeffs + (size_t)b0 * nz, base_opd, s.opd_c, st, cb);
    if (rc) return rc;
    // Ebar_bl = 2 w_l psf_bar[b] .* E_bl
    rc = launch_cotangent(M, c, (const float2*)field + (size_t)b0 * L * mpix, psf_bar + (size_t)b0 * mpix, weights,
                          s.ebar_pl, nullptr, -1, st, L);
    if (rc) return rc;
    GemmParams g{};
    fill_stage(g, true, 0, N, M, c, s.xin, s.uout, sign2pi);
    g.a = s.ebar_pl;
    g.out = s.mid_pl;
    g.mode = EPI_PLANES;
    GemmParams h{};
    fill_stage(h, true, 1, N, M, c, s.xin, s.uout, sign2pi);
    h.a = s.mid_pl;
    h.mode = EPI_C64;
    h.scale = s.norm_item;
    h.out_c64 = s.qbuf;
    rc = run_gemm_pair(g, h, s.sync_ws, d->precision, st);
    if (rc) return rc;
    // opd_bar[b] = sum_l k_l Im(conj(P_bl) Q_bl), then its projection on the basis: coeff_bar[b]
    rc = launch_grad_reduce(npix, L, s.qbuf, s.k_item, T, s.opd_c, phase, s.amp_scale, a0, s.opdbar_c, nullptr,
                            nullptr, 0, st, cb);
    if (rc) return rc;
    rc = launch_basis_reduce(nz, (int64_t)npix, basis, s.opdbar_c, coeff_bar + (size_t)b0 * nz, st, cb);
    if (rc) return rc;
  }
  return DLUX_OK;
}

int dlux_basis_eval(int32_t nz, int64_t npix, const float* basis, const float* coeffs,
                    const float* base, float* out, void* cuda_stream) {
  if (!basis || !coeffs || !out) return DLUX_ERR_ARG;
  if (nz < 1 || nz > 8192 || npix < 1) return DLUX_ERR_SHAPE;
  return launch_basis_eval(nz, npix, basis, coeffs, base, out, (cudaStream_t)cuda_stream);
}

int dlux_basis_reduce(int32_t nz, int64_t npix, const float* basis, const float* out_bar,
                      float* coeff_bar, void* cuda_stream) {
  if (!basis || !out_bar || !coeff_bar) return DLUX_ERR_ARG;
  if (nz < 1 || npix < 1) return DLUX_ERR_SHAPE;
  return launch_basis_reduce(nz, npix, basis, out_bar, coeff_bar, (cudaStream_t)cuda_stream);
}

}  // extern "C"
