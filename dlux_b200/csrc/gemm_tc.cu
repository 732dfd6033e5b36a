// DLUX_PREC_3XTF32: the phasor GEMM stage on Blackwell tensor cores (sm_100a).
//
//   D[m][n] = sum_k Data[m][k] * G[k][n],   G[k][n] = exp(i * fl(sign2pi * fl(kvec[k]*nvec[n])))
//
// * A operand (data): four planar fp32 planes (re_hi, re_lo, im_hi, im_lo; hi = rna-tf32,
//   lo = exact residual) streamed by TMA (cp.async.bulk.tensor, SWIZZLE_64B, K-major).
// * B operand (DFT phasors): never materialised in HBM.  Generator warps evaluate the
//   reference's float32 phase argument, sincosf, split into tf32 hi/lo and write K-major
//   SWIZZLE_64B tiles straight into shared memory (generic proxy -> fence.proxy.async ->
//   mbarrier).  Two stacked arrangements of the same 64 phasor columns are written,
//   B1 = [cos; sin] and B2 = [-sin; cos] (128 rows each), so that ONE N=128 MMA produces
//   the real and the imaginary halves of the complex product:
//       [Re | Im] += A_re * B1 + A_im * B2.
// * 3xTF32: 6 tcgen05.mma kind::tf32 (M128 x N128 x K8) per k-step -- lo*hi, hi*lo and
//   hi*hi for each of the two products; lo*lo is dropped (2^-22 relative).
// * Tensor-core fp32 accumulation truncates (measured: ~2e-8 relative per accumulate,
//   systematic), so long K chains are NOT kept in TMEM: every FLUSH_CHUNKS k-chunks the
//   partial accumulator (one of four 128-column TMEM buffers) is drained by the epilogue
//   warps with tcgen05.ld and added, round-to-nearest, into fp32 registers (the scheme of
//   Ootomo & Yokota for error-corrected TF32 GEMM).  Draining overlaps the MMAs of the
//   next partial.
// * Persistent CTAs, one per SM, 16 warps in 4 warpgroups: WG0 = drain + fused epilogue
//   (one warp per TMEM lane quarter; setmaxnreg.inc, they hold the 128 running totals),
//   WG1 = TMA producer + MMA issuer (setmaxnreg.dec), WG2/WG3 = phasor generators.
//   The data (TMA) and phasor (generated) operands have separate shared-memory rings
//   (4 x 32 KiB and 3 x 32 KiB) so that HBM/L2 latency gets the deeper prefetch.
#include <cuda.h>
#include <cstdio>
#include <mutex>
#include "common.cuh"

namespace dlux {

namespace {

constexpr int BM = 128;            // data rows per tile (UMMA M)
constexpr int NB = 64;             // generated output coordinates per tile
constexpr int BN = 2 * NB;         // UMMA N: [Re | Im] halves
constexpr int BK = 16;             // k per pipeline stage = one 64-byte swizzle row of fp32
constexpr int UMMA_K = 8;          // kind::tf32
constexpr int A_STAGES = 4;        // data ring (TMA)
constexpr int B_STAGES = 3;        // phasor ring (generated)
constexpr int PLANE_BYTES = BM * BK * 4;            // 8 KiB
constexpr int A_BYTES = 4 * PLANE_BYTES;            // 32 KiB: re_hi, re_lo, im_hi, im_lo
constexpr int B_BYTES = 4 * BN * BK * 4;            // 32 KiB: B1_hi, B1_lo, B2_hi, B2_lo
constexpr int B_BASE = A_STAGES * A_BYTES;
constexpr int RING_BYTES = A_STAGES * A_BYTES + B_STAGES * B_BYTES;  // 224 KiB
constexpr int FLUSH_CHUNKS = 4;    // k-chunks accumulated in TMEM before draining to registers
constexpr int NUM_ACC = 4;         // TMEM partial-accumulator ring
constexpr int NUM_EPI_WARPS = 4;   // warps 0..3   (WG0)
constexpr int WARP_TMA = 4;        // WG1
constexpr int WARP_MMA = 5;
constexpr int FIRST_GEN_WARP = 8;  // WG2, WG3
constexpr int NUM_GEN_WARPS = 8;
constexpr int NUM_GEN_THREADS = 32 * NUM_GEN_WARPS;
constexpr int PH_PER_THREAD = NB * BK / NUM_GEN_THREADS;  // phasors per generator thread per chunk (4)
constexpr int NUM_THREADS = 32 * (FIRST_GEN_WARP + NUM_GEN_WARPS);  // 512
constexpr int REGS_EPI = 232, REGS_CTRL = 40, REGS_GEN = 112;  // setmaxnreg budgets (launch: 128)
constexpr int TMEM_COLS = NUM_ACC * BN;  // 512
static_assert(TMEM_COLS == 512, "TMEM ring must be a power of two <= 512 columns");
static_assert(PH_PER_THREAD % 4 == 0, "each generator thread writes whole 16-byte chunks");
static_assert(128 * (REGS_EPI - 128) <= 128 * (128 - REGS_CTRL) + 256 * (128 - REGS_GEN), "register budget");
constexpr int SMEM_BYTES = RING_BYTES + 1024 /*align*/ + 256 /*barriers*/;
static_assert(SMEM_BYTES <= 232448, "exceeds 227 KiB of shared memory per CTA");

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major SWIZZLE_64B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) (=1, unused) | SBO>>4 [32,46) (8 rows * 64 B = 512)
// | version=1 [46,48) | layout_type=4 (SWIZZLE_64B) [61,64)
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}

// kind::tf32 instruction descriptor: D=f32, A=B=tf32, K-major both, N=128, M=128.
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                           ((uint32_t)(BM >> 4) << 24);

// Fused epilogue of one tile row: thread = data row m, tot[0..NB) = Re, tot[NB..2NB) = Im
// of the NB output coordinates n0.. .  Stores are transposed (m fastest), so a warp writes
// 128 (fp32) or 256 (c64) contiguous bytes per output coordinate.  Global loads of the
// gradient epilogue are issued in batches of 8 so that their latency overlaps.
__device__ __forceinline__ void tile_epilogue(const GemmParams& p, int item, int m, int n0,
                                              float (&tot)[BN]) {
  const float sc = p.scale ? __ldg(p.scale + item) : 1.0f;
  const int nmax = p.n_out - n0;  // columns j < nmax are valid
  if (p.mode == EPI_PLANES) {
    const size_t base = ((size_t)item * p.n_out + n0) * p.out_pitch + m;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      if (j < nmax) {
        const float re = tot[j] * sc, im = tot[NB + j] * sc;
        const float rh = tf32_hi(re), ih = tf32_hi(im);
        const size_t o = base + (size_t)j * p.out_pitch;
        p.out_planes[0][o] = rh;
        p.out_planes[1][o] = re - rh;
        p.out_planes[2][o] = ih;
        p.out_planes[3][o] = im - ih;
      }
    }
  } else if (p.mode == EPI_C64) {
    const size_t base = ((size_t)item * p.n_out + n0) * p.rows + m;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      if (j < nmax) p.out_c64[base + (size_t)j * p.rows] = make_float2(tot[j] * sc, tot[NB + j] * sc);
    }
  } else {  // EPI_GRAD
    const float kw = __ldg(p.w + item);
    const float amp = p.a0 * __ldg(p.amp_scale);
    const size_t base = (size_t)n0 * p.rows + m;
    float* outg = p.out_g + (size_t)item * p.n_out * p.rows + base;
#pragma unroll
    for (int j0 = 0; j0 < NB; j0 += 16) {
      float tv[16], ov[16], pv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const bool ok = (j0 + j) < nmax;
        const size_t o = base + (size_t)(j0 + j) * p.rows;
        tv[j] = (ok && p.pup_T) ? __ldg(p.pup_T + o) : 1.0f;
        ov[j] = (ok && p.pup_opd) ? __ldg(p.pup_opd + o) : 0.0f;
        pv[j] = (ok && p.pup_phase) ? __ldg(p.pup_phase + o) : 0.0f;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if ((j0 + j) < nmax) {
          const float re = tot[j0 + j] * sc, im = tot[NB + j0 + j] * sc;
          float sn, cs;
          fast_sincos(__fmul_rn(kw, ov[j]) + pv[j], &sn, &cs);
          outg[(size_t)(j0 + j) * p.rows] = amp * tv[j] * (cs * im - sn * re);  // Im(conj(P) * v)
        }
      }
    }
  }
}

struct TcParams {
  GemmParams g;
  int tiles_m, tiles_n, n_tiles, k_chunks;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
               const __grid_constant__ CUtensorMap map2, const __grid_constant__ CUtensorMap map3,
               const TcParams tp) {
  extern __shared__ uint8_t smem_raw[];
  const GemmParams& p = tp.g;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));  // generic view of the aligned base
  const uint32_t bar_base = smem_base + RING_BYTES;
  auto fullA_bar = [&](int s) { return bar_base + 8u * s; };
  auto emptyA_bar = [&](int s) { return bar_base + 8u * (A_STAGES + s); };
  auto fullB_bar = [&](int s) { return bar_base + 8u * (2 * A_STAGES + s); };
  auto emptyB_bar = [&](int s) { return bar_base + 8u * (2 * A_STAGES + B_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * A_STAGES + 2 * B_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * A_STAGES + 2 * B_STAGES + NUM_ACC + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(
      smem_gen + RING_BYTES + 8 * (2 * A_STAGES + 2 * B_STAGES + 2 * NUM_ACC));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == WARP_TMA && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map0));
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map1));
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map2));
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map3));
  }
  if (warp == WARP_MMA && lane == 0) {
    for (int s = 0; s < A_STAGES; ++s) {
      mbar_init(fullA_bar(s), 1);   // TMA producer's arrive.expect_tx
      mbar_init(emptyA_bar(s), 1);  // tcgen05.commit
    }
    for (int s = 0; s < B_STAGES; ++s) {
      mbar_init(fullB_bar(s), NUM_GEN_WARPS);  // one arrive per generator warp
      mbar_init(emptyB_bar(s), 1);             // tcgen05.commit
    }
    for (int a = 0; a < NUM_ACC; ++a) {
      mbar_init(tfull_bar(a), 1);                    // tcgen05.commit closing a partial
      mbar_init(tempty_bar(a), NUM_EPI_WARPS * 32);  // every drain thread
    }
    fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Register re-balancing between warpgroups: setmaxnreg is the first instruction of each
  // warpgroup's branch (all four warps of a warpgroup execute it).
  const int tiles_per_item = tp.tiles_m * tp.tiles_n;

  if (warp >= NUM_EPI_WARPS && warp < FIRST_GEN_WARP) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_CTRL));
    if (warp == WARP_TMA) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < tp.n_tiles; tile += gridDim.x) {
        const int item = tile / tiles_per_item;
        const int t = tile % tiles_per_item;
        const int m0 = (t % tp.tiles_m) * BM;
        const int d = p.item_data ? __ldg(p.item_data + item) : item;
        for (int kc = 0; kc < tp.k_chunks; ++kc) {
          mbar_wait(emptyA_bar(stage), phase ^ 1);
          const uint32_t a_dst = smem_base + stage * A_BYTES;
          mbar_arrive_expect_tx(fullA_bar(stage), A_BYTES);
          tma_load_3d(a_dst + 0 * PLANE_BYTES, &map0, fullA_bar(stage), kc * BK, m0, d);
          tma_load_3d(a_dst + 1 * PLANE_BYTES, &map1, fullA_bar(stage), kc * BK, m0, d);
          tma_load_3d(a_dst + 2 * PLANE_BYTES, &map2, fullA_bar(stage), kc * BK, m0, d);
          tma_load_3d(a_dst + 3 * PLANE_BYTES, &map3, fullA_bar(stage), kc * BK, m0, d);
          if (++stage == A_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    } else if (warp == WARP_MMA) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < tp.n_tiles; tile += gridDim.x) {
        for (int kc = 0; kc < tp.k_chunks; ++kc) {
          const int in_partial = kc % FLUSH_CHUNKS;
          if (in_partial == 0) {
            mbar_wait(tempty_bar(acc), acc_phase ^ 1);  // drain warps are done with this buffer
            tc_fence_after();
          }
          const uint32_t d = tmem_base + (uint32_t)(acc * BN);
          mbar_wait(fullB_bar(sb), pb);
          mbar_wait(fullA_bar(sa), pa);
          tc_fence_after();
          const uint32_t a0 = smem_base + sa * A_BYTES;
          const uint32_t b0 = smem_base + B_BASE + sb * B_BYTES;
#pragma unroll
          for (int ks = 0; ks < BK / UMMA_K; ++ks) {
            const uint32_t koff = ks * UMMA_K * 4;  // bytes inside the 64-byte swizzle row
            const uint64_t a_rh = make_desc_sw64(a0 + 0 * PLANE_BYTES + koff);
            const uint64_t a_rl = make_desc_sw64(a0 + 1 * PLANE_BYTES + koff);
            const uint64_t a_ih = make_desc_sw64(a0 + 2 * PLANE_BYTES + koff);
            const uint64_t a_il = make_desc_sw64(a0 + 3 * PLANE_BYTES + koff);
            const uint64_t b1h = make_desc_sw64(b0 + 0 * PLANE_BYTES + koff);
            const uint64_t b1l = make_desc_sw64(b0 + 1 * PLANE_BYTES + koff);
            const uint64_t b2h = make_desc_sw64(b0 + 2 * PLANE_BYTES + koff);
            const uint64_t b2l = make_desc_sw64(b0 + 3 * PLANE_BYTES + koff);
            const uint32_t accum = (in_partial | ks) ? 1u : 0u;
            // [Re | Im] += A_re * [cos; sin] + A_im * [-sin; cos]; small terms first
            umma_tf32(d, a_rl, b1h, IDESC, accum);
            umma_tf32(d, a_rh, b1l, IDESC, 1u);
            umma_tf32(d, a_il, b2h, IDESC, 1u);
            umma_tf32(d, a_ih, b2l, IDESC, 1u);
            umma_tf32(d, a_rh, b1h, IDESC, 1u);
            umma_tf32(d, a_ih, b2h, IDESC, 1u);
          }
          umma_commit(emptyA_bar(sa));  // free both smem slots when these MMAs retire
          umma_commit(emptyB_bar(sb));
          if (in_partial == FLUSH_CHUNKS - 1 || kc == tp.k_chunks - 1) {
            umma_commit(tfull_bar(acc));  // partial complete -> drain warps
            if (++acc == NUM_ACC) { acc = 0; acc_phase ^= 1; }
          }
          if (++sa == A_STAGES) { sa = 0; pa ^= 1; }
          if (++sb == B_STAGES) { sb = 0; pb ^= 1; }
        }
      }
    }
    }
  } else if (warp < NUM_EPI_WARPS) {
    // ===================== drain + epilogue =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_EPI));
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    const int n_partials = (tp.k_chunks + FLUSH_CHUNKS - 1) / FLUSH_CHUNKS;
    for (int tile = blockIdx.x; tile < tp.n_tiles; tile += gridDim.x) {
      const int item = tile / tiles_per_item;
      const int t = tile % tiles_per_item;
      const int m = (t % tp.tiles_m) * BM + q * 32 + lane;
      const int n0 = (t / tp.tiles_m) * NB;
      float tot[BN];
#pragma unroll
      for (int j = 0; j < BN; ++j) tot[j] = 0.0f;
      for (int part = 0; part < n_partials; ++part) {
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll
        for (int c = 0; c < BN; c += 32) {
          uint32_t v0[16], v1[16];
          tmem_ld16(t0 + c, v0);
          tmem_ld16(t0 + c + 16, v1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            tot[c + j] += __uint_as_float(v0[j]);
            tot[c + 16 + j] += __uint_as_float(v1[j]);
          }
        }
        tc_fence_before();
        mbar_arrive(tempty_bar(acc));
        if (++acc == NUM_ACC) { acc = 0; acc_phase ^= 1; }
      }
      if (m < p.rows) tile_epilogue(p, item, m, n0, tot);
    }
  } else {
    // ===================== phasor generators =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_GEN));
    const int gt = threadIdx.x - 32 * FIRST_GEN_WARP;
    const int nl = gt & (NB - 1);   // output coordinate within the tile
    const int kg = gt / NB;         // which PH_PER_THREAD-wide k group of the chunk
    int stage = 0;
    uint32_t phase = 0;
    // SWIZZLE_64B: 16-byte chunk c of row r lives at chunk c ^ ((r >> 1) & 3).  Rows nl
    // and nl + 64 share the same swizzle term (64 is a multiple of 8).
    const uint32_t row_lo = (uint32_t)nl * 64u;          // rows [0, 64): cos (B1) / -sin (B2)
    const uint32_t row_hi = (uint32_t)(nl + NB) * 64u;   // rows [64, 128): sin (B1) / cos (B2)
    const uint32_t sw = ((uint32_t)nl >> 1) & 3u;
    for (int tile = blockIdx.x; tile < tp.n_tiles; tile += gridDim.x) {
      const int item = tile / tiles_per_item;
      const int t = tile % tiles_per_item;
      const int n = (t / tp.tiles_m) * NB + nl;
      const float* kv = p.kvec + (size_t)item * p.kvec_stride;
      const float u = (n < p.n_out) ? __ldg(p.nvec + (size_t)item * p.nvec_stride + n) : 0.0f;
      float xk[PH_PER_THREAD];  // this chunk's k coordinates, prefetched one chunk ahead
#pragma unroll
      for (int j = 0; j < PH_PER_THREAD; ++j) {
        const int k = kg * PH_PER_THREAD + j;
        xk[j] = (k < p.K) ? __ldg(kv + k) : 0.0f;
      }
      for (int kc = 0; kc < tp.k_chunks; ++kc) {
        float xn[PH_PER_THREAD];
#pragma unroll
        for (int j = 0; j < PH_PER_THREAD; ++j) {
          const int k = (kc + 1) * BK + kg * PH_PER_THREAD + j;
          xn[j] = (k < p.K) ? __ldg(kv + k) : 0.0f;
        }
        float c_hi[PH_PER_THREAD], c_lo[PH_PER_THREAD], s_hi[PH_PER_THREAD], s_lo[PH_PER_THREAD];
#pragma unroll
        for (int j = 0; j < PH_PER_THREAD; ++j) {
          float sn, cs;
          fast_sincos(phase_arg(p.sign2pi, xk[j], u), &sn, &cs);
          xk[j] = xn[j];
          c_hi[j] = tf32_hi(cs);
          c_lo[j] = cs - c_hi[j];
          s_hi[j] = tf32_hi(sn);
          s_lo[j] = sn - s_hi[j];
        }
        mbar_wait(emptyB_bar(stage), phase ^ 1);
        uint8_t* bb = smem_gen + B_BASE + stage * B_BYTES;
#pragma unroll
        for (int h = 0; h < PH_PER_THREAD / 4; ++h) {
          const uint32_t chunk = ((uint32_t)(kg * (PH_PER_THREAD / 4) + h) ^ sw) * 16u;
          const float4 ch = make_float4(c_hi[4 * h], c_hi[4 * h + 1], c_hi[4 * h + 2], c_hi[4 * h + 3]);
          const float4 cl = make_float4(c_lo[4 * h], c_lo[4 * h + 1], c_lo[4 * h + 2], c_lo[4 * h + 3]);
          const float4 sh = make_float4(s_hi[4 * h], s_hi[4 * h + 1], s_hi[4 * h + 2], s_hi[4 * h + 3]);
          const float4 sl = make_float4(s_lo[4 * h], s_lo[4 * h + 1], s_lo[4 * h + 2], s_lo[4 * h + 3]);
          const float4 nsh = make_float4(-sh.x, -sh.y, -sh.z, -sh.w);
          const float4 nsl = make_float4(-sl.x, -sl.y, -sl.z, -sl.w);
          *reinterpret_cast<float4*>(bb + 0 * PLANE_BYTES + row_lo + chunk) = ch;   // B1_hi: cos
          *reinterpret_cast<float4*>(bb + 0 * PLANE_BYTES + row_hi + chunk) = sh;   //        sin
          *reinterpret_cast<float4*>(bb + 1 * PLANE_BYTES + row_lo + chunk) = cl;   // B1_lo
          *reinterpret_cast<float4*>(bb + 1 * PLANE_BYTES + row_hi + chunk) = sl;
          *reinterpret_cast<float4*>(bb + 2 * PLANE_BYTES + row_lo + chunk) = nsh;  // B2_hi: -sin
          *reinterpret_cast<float4*>(bb + 2 * PLANE_BYTES + row_hi + chunk) = ch;   //        cos
          *reinterpret_cast<float4*>(bb + 3 * PLANE_BYTES + row_lo + chunk) = nsl;  // B2_lo
          *reinterpret_cast<float4*>(bb + 3 * PLANE_BYTES + row_hi + chunk) = cl;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(fullB_bar(stage));
        if (++stage == B_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

struct TcState {
  EncodeTiledFn encode = nullptr;
  int num_sms = 0;
  int cc_major = 0;
  int rc = DLUX_OK;
};

TcState& tc_state() {
  static TcState st;
  static std::once_flag once;
  std::call_once(once, [] {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { st.rc = DLUX_ERR_CUDA; return; }
    cudaDeviceGetAttribute(&st.num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&st.cc_major, cudaDevAttrComputeCapabilityMajor, dev);
    if (st.cc_major != 10) { st.rc = DLUX_ERR_UNSUPPORTED; return; }
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || !fn) {
      st.rc = DLUX_ERR_CUDA;
      return;
    }
    st.encode = (EncodeTiledFn)fn;
    if (cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) !=
        cudaSuccess) {
      st.rc = DLUX_ERR_CUDA;
      return;
    }
  });
  return st;
}

}  // namespace

size_t gemm_tc_workspace_bytes() { return 0; }

int launch_gemm_tc(const GemmParams& p, cudaStream_t st) {
  if (p.n_items <= 0) return DLUX_OK;
  TcState& s = tc_state();
  if (s.rc != DLUX_OK) return s.rc;
  if (p.a_pitch % 4 != 0) return DLUX_ERR_SHAPE;  // TMA global strides are multiples of 16 bytes

  const cuuint64_t n_data = (cuuint64_t)(p.n_data > 0 ? p.n_data : p.n_items);
  CUtensorMap maps[4];
  for (int i = 0; i < 4; ++i) {
    cuuint64_t dims[3] = {(cuuint64_t)p.K, (cuuint64_t)p.rows, n_data};
    cuuint64_t strides[2] = {(cuuint64_t)p.a_pitch * 4, (cuuint64_t)p.a_pitch * 4 * (cuuint64_t)p.rows};
    cuuint32_t box[3] = {BK, BM, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = s.encode(&maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)p.a_planes[i], dims, strides,
                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      fprintf(stderr, "[dlux_b200] cuTensorMapEncodeTiled failed: %d\n", (int)r);
      return DLUX_ERR_CUDA;
    }
  }
  TcParams tp;
  tp.g = p;
  tp.tiles_m = (p.rows + BM - 1) / BM;
  tp.tiles_n = (p.n_out + NB - 1) / NB;
  const long long total = (long long)tp.tiles_m * tp.tiles_n * p.n_items;
  if (total > 2147483647LL) return DLUX_ERR_SHAPE;
  tp.n_tiles = (int)total;
  tp.k_chunks = (p.K + BK - 1) / BK;
  const int grid = tp.n_tiles < s.num_sms ? tp.n_tiles : s.num_sms;
  gemm_tc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(maps[0], maps[1], maps[2], maps[3], tp);
  note_launch();
  return check_launch("gemm_tc");
}

}  // namespace dlux
