// DLUX_PREC_3XTF32: the phasor GEMM stage on Blackwell tensor cores (sm_100a).
//
//   Out[n][m] = sum_k G[n][k] * Data[m][k],   G[n][k] = exp(i * fl(sign2pi * fl(kvec[k]*nvec[n])))
//
// A CTA pair (cluster of 2, tcgen05.mma.cta_group::2) works on one unit = 128 output coordinates
// n (64 per CTA) x 256 data rows m, computed as TWO tiles (a, b) of 128 data rows that share
// every generated phasor -- generation, not the tensor pipe, is the scarce resource, so each
// phasor is used for 256 data rows.  The MMA is issued in its "TS" form, M = 256 over the pair:
//
// * A operand = the DFT phasors, in TENSOR MEMORY (128 lanes in each CTA).  They are never
//   materialised in HBM nor in shared memory: generator warps evaluate the reference's float32
//   phase argument, an exact Cody-Waite reduction + MUFU sincos, split hi/lo (tf32) and write
//   the operand tiles straight from registers with tcgen05.st.  TMEM lane 2j holds the "real"
//   row of phasor column j and lane 2j+1 its "imaginary" row, in two arrangements
//       G1 = (cos | sin),   G2 = (-sin | cos)        (row 2j | row 2j+1)
//   so that  D = G1 * Re(Data)^T + G2 * Im(Data)^T  has Re(Out[j][:]) in lane 2j and
//   Im(Out[j][:]) in lane 2j+1: one MMA yields both complex parts.
// * B operand = the data (PlaneSet, common.cuh): two planar float32 matrices (Re, Im), 8 bytes
//   per complex element in HBM.  TMA streams them K-major into SWIZZLE_64B boxes (64 rows x
//   16 k per CTA and tile); two CONVERTER warps per CTA then split every value on chip: hi =
//   rna_tf32(z) is written back in place (the kind::tf32 operand) and bf16(hi), bf16(z - hi)
//   go to SWIZZLE_32B planes next to it (the kind::f16 operands).  Each CTA holds 64 of the
//   128 rows of a tile: 32 KiB stages (both tiles, float32 + bf16 planes), a 5-deep ring.
//   Round 1 streamed six pre-split planes (16 bytes per element) from HBM instead.
// * Split precision ("TF32 + 2xBF16"): per 16-k chunk and tile, 4 tcgen05.mma kind::tf32
//   (hi*hi, K=8) + 4 kind::f16 bf16 MMAs (hi*lo and lo*hi, K=16) accumulate into the same
//   fp32 TMEM accumulator -- 8 MMA slots instead of the 12 of 3xTF32; lo*lo is dropped.
//   Measured error of the scheme: 5.6e-7 relative per contraction (3xTF32: 6.6e-8).
// * Tensor-core fp32 accumulation truncates (measured ~2e-8 relative systematic loss per
//   accumulate), so K chains are cut every FLUSH_CHUNKS k-chunks: the partial accumulators
//   (three 128-column TMEM buffers used round-robin by the two tiles) are drained by the
//   tile's epilogue warpgroup with tcgen05.ld and added, round-to-nearest, into fp32
//   registers (Ootomo & Yokota's scheme for error-corrected TF32 GEMM).  Draining overlaps
//   the MMAs of the other tile / next partial.
// * Barriers: rawfull (local: this CTA's TMA bytes have landed, wakes the converters); the
//   "full" side of every ring lives in the LEADER CTA (rank 0), which alone
//   issues MMAs: fullA (both CTAs' converter warps), fullG (both CTAs' generator warps), tempty
//   (both CTAs' drain warpgroups) -- remote arrivals are mbarrier.arrive.relaxed.cluster.  The
//   "empty" side is local to each CTA and signalled by tcgen05.commit ... multicast::cluster:
//   emptyA, emptyG, tfull.
// * Epilogues: EPI_PLANES (the intermediate of a two-stage transform) goes out through TMA
//   stores -- every TMEM lane is one output row, so each drain thread drops its own
//   16-column slice into the swizzled box layout of its plane (Re lanes -> plane 0, Im lanes
//   -> plane 1) and the TMA writes the two boxes of the warp asynchronously, four slices in
//   flight, clipping at the matrix edge.  EPI_C64
//   (final complex64 result) does the same with one SWIZZLE_128B box of 16 rows x 16 complex
//   per warp and slice (real lane -> even words, imaginary lane -> odd words); only an odd
//   row length (global pitch not a multiple of 16 bytes) takes the load/store epilogue, which
//   transposes through shared memory into 128-byte row segments.
// * Persistent CTAs, one per SM, 20 warps in 5 warpgroups: WG0 / WG1 = drain + fused
//   epilogue of tile a / b (setmaxnreg.inc: 128 running totals per thread), WG2 = TMA
//   producer + one MMA issuer warp per tile (setmaxnreg.dec), WG3 / WG4 = phasor generators
//   (two warps per TMEM lane quarter, one per k-step of the chunk).
// TMEM map (512 columns): [0,384) three partial accumulators, [384,512) two phasor stages
// of 64 columns: G1_hi, G2_hi as tf32 (16 k -> 16 columns each) and G1_hi, G1_lo, G2_hi,
// G2_lo as packed bf16 (16 k -> 8 columns each).
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include "common.cuh"

namespace dlux {

namespace {

constexpr int BM = 128;            // data rows per tile  (UMMA N)
constexpr int NB = 64;             // output coordinates per tile; 2*NB TMEM lanes (UMMA M = 128)
constexpr int BK = 16;             // k per pipeline stage = one 64-byte swizzle row of fp32
constexpr int UMMA_K = 8;          // kind::tf32
constexpr int ROWS_CTA = BM / 2;   // rows of a data tile resident in this CTA's shared memory (the peer holds the rest)
#ifndef DLUX_A_STAGES
#define DLUX_A_STAGES 5            // 6 = deeper ring, single-buffered epilogue staging (0.8 % slower)
#endif
constexpr int A_STAGES = DLUX_A_STAGES;   // data ring (TMA), each stage holds this CTA's half of both tiles
constexpr int G_STAGES = 2;        // phasor ring (TMEM)
constexpr int PLANE_BYTES = ROWS_CTA * BK * 4;      // one fp32 (tf32) plane of a tile-chunk
constexpr int BPLANE_BYTES = ROWS_CTA * BK * 2;     // one bf16 plane
constexpr int BPL_BASE = 2 * PLANE_BYTES;           // bf16 planes follow the two fp32 planes
constexpr int TILE_BYTES = 2 * PLANE_BYTES + 4 * BPLANE_BYTES;  // 16 KiB: my half of a tile-chunk
constexpr int A_BYTES = 2 * TILE_BYTES;             // 32 KiB: tiles a and b
constexpr int RING_BYTES = A_STAGES * A_BYTES;      // 160 KiB
#ifndef DLUX_FLUSH_CHUNKS
#define DLUX_FLUSH_CHUNKS 4
#endif
constexpr int FLUSH_CHUNKS = DLUX_FLUSH_CHUNKS;    // k-chunks accumulated in TMEM before draining to registers
constexpr int NUM_ACC = 3;         // TMEM partial-accumulator buffers, used round-robin (a, b, a, b, ...)
constexpr int ACC_COLS = BM;       // 128 fp32 columns per partial
constexpr int G_COLS = 4 * BK;     // 64 columns per phasor stage
constexpr int GB_BASE = 2 * BK;    // packed-bf16 planes start after the two tf32 planes
constexpr int GB_COLS = BK / 2;    // 8 columns per bf16 plane
constexpr int G_BASE_COL = NUM_ACC * ACC_COLS;      // 384
constexpr int TMEM_COLS = 512;
static_assert(G_BASE_COL + G_STAGES * G_COLS <= TMEM_COLS, "TMEM budget");
constexpr int NUM_EPI_WARPS = 8;   // warps 0..3 drain tile a (WG0), warps 4..7 tile b (WG1)
constexpr int WARP_TMA = 8;        // WG2
constexpr int WARP_MMA = 9;
constexpr int RAW_BYTES = 4 * (ROWS_CTA * BK * 4);   // TMA bytes per stage and CTA: Re, Im of tiles a and b
constexpr int FIRST_GEN_WARP = 12; // WG3 (k-step 0 of each chunk), WG4 (k-step 1)
constexpr int NUM_GEN_WARPS = 8;
constexpr int FIRST_CONV_WARP = FIRST_GEN_WARP + NUM_GEN_WARPS;      // WG5: operand converters, warp w -> (tile, plane)
constexpr int NUM_CONV_WARPS = 4;
constexpr int NUM_THREADS = 32 * (FIRST_CONV_WARP + NUM_CONV_WARPS);  // 768
constexpr int REGS_LAUNCH = 80;    // 65536 / 768 rounded down to a multiple of 8
#ifndef DLUX_REGS_EPI
#define DLUX_REGS_EPI 152
#define DLUX_REGS_CTRL 32
#define DLUX_REGS_GEN 56
#define DLUX_REGS_CONV 32
#endif
constexpr int REGS_EPI = DLUX_REGS_EPI, REGS_CTRL = DLUX_REGS_CTRL, REGS_GEN = DLUX_REGS_GEN,
              REGS_CONV = DLUX_REGS_CONV;   // setmaxnreg budgets
// setmaxnreg.inc draws from the registers the CTA was LAUNCHED with (NUM_THREADS * REGS_LAUNCH), fed by
// the .dec of the other warpgroups -- not from unallocated registers of the SM
static_assert(256 * (REGS_EPI - REGS_LAUNCH) <= 128 * (REGS_LAUNCH - REGS_CTRL) + 256 * (REGS_LAUNCH - REGS_GEN) +
                                                    128 * (REGS_LAUNCH - REGS_CONV),
              "register budget");
constexpr int CLUSTER = 2;         // the CTA pair of a cta_group::2 MMA
constexpr int BAR_BYTES = 1024;     // barriers + TMEM slot (keeps the staging 1 KiB aligned)
constexpr int STG_SLICE_BYTES = 2048;                      // one 16-column slice of the two planes of a warp
constexpr int STG_BUFS = A_STAGES >= 6 ? 2 : 4;            // slices in flight per warp
constexpr int STG_WARP_BYTES = STG_BUFS * STG_SLICE_BYTES; // epilogue staging per drain warp
constexpr int SMEM_BYTES = RING_BYTES + 1024 /*align*/ + BAR_BYTES + NUM_EPI_WARPS * STG_WARP_BYTES;
static_assert(SMEM_BYTES <= 232448, "exceeds 227 KiB of shared memory per CTA");

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or the
// limit expires) instead of re-issuing the poll every ~18 cycles -- ncu showed 20 % of all executed
// instructions were YIELD / TRYWAIT / BRA of waiting warps, taking issue slots from the generators
#ifndef DLUX_WAIT_HINT_NS
#define DLUX_WAIT_HINT_NS 0x989680
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
#if DLUX_WAIT_HINT_NS > 0
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
#else
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
#endif
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n" : "=r"(done) : "r"(bar), "r"(parity), "r"((uint32_t)DLUX_WAIT_HINT_NS) : "memory");
  } while (!done);
}
// ---- cluster-scope forms (pair mode): barriers that live in the leader CTA of the pair
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // relaxed: what the arrival publishes lives in tensor memory and is ordered by the
  // tcgen05 fences around it, not by the generic-proxy release
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
#if DLUX_WAIT_HINT_NS > 0
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
#else
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
#endif
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n" : "=r"(done) : "r"(bar), "r"(parity), "r"((uint32_t)DLUX_WAIT_HINT_NS) : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// destination and completion barrier in my own shared memory (the converters of THIS CTA wait on it)
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// the same with an L2 cache policy (createpolicy): the fused launch keeps its ring of intermediates resident
__device__ __forceinline__ void tma_load_3d_hint(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                                 int c0, int c1, int c2, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// The converters' arrival on the leader's fullA barrier (their generic-proxy writes are fenced to the async proxy
// first).  Default semantics, .release at CTA scope: the converted operand never leaves this CTA's shared memory
// -- each SM's tensor core reads its own half -- only the signal crosses to the leader (the form CUTLASS'
// 2-SM transform pipelines use after fence.proxy.async; a .release.cluster arrival cost ~700 cycles per chunk)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void umma_commit_mc2(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask) : "memory");
}
// bulk tensor store shared -> global (rows / columns beyond the tensor are clipped)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_3d_hint(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2,
                                                  uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// D[tmem] (+)= A[tmem] * B[smem] over the CTA pair: M = 256 (128 TMEM lanes in each CTA; A: 8
// columns of tf32 or of packed bf16 per lane, K-major), B = 64 rows from each CTA's shared memory
__device__ __forceinline__ void umma_tf32_ts2(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_bf16_ts2(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
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
__device__ __forceinline__ void tmem_st4u(uint32_t taddr, const uint32_t (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
// two floats -> packed bf16x2, `lo` in bits [0,16) (the lower k index), `hi` in bits [16,32)
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// this thread's lane, 8 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
      ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// One lane of the (converged) warp is elected; the tcgen05/TMA issue blocks are guarded by
// this instead of `lane == 0` so that the compiler keeps them on the uniform datapath.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, 0xFFFFFFFF;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t"
      "}\n" : "=r"(pred));
  return pred != 0;
}

// K-major SWIZZLE_64B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) (=1, unused) | SBO>>4 [32,46) (8 rows * 64 B = 512)
// | version=1 [46,48) | layout_type=4 (SWIZZLE_64B) [61,64)
constexpr uint64_t DESC_SW64_HI = ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
                                  ((uint64_t)4 << 61);
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
  return DESC_SW64_HI | (uint64_t)((saddr & 0x3FFFFu) >> 4);  // CTA-local offset (< 256 KiB): 14 bits after >> 4
}

// K-major SWIZZLE_32B descriptor for the bf16 planes (rows of 16 bf16 = 32 bytes; 8 rows = 256 B)
constexpr uint64_t DESC_SW32_HI = ((uint64_t)1 << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46) |
                                  ((uint64_t)6 << 61);
__device__ __forceinline__ uint64_t make_desc_sw32(uint32_t saddr) {
  return DESC_SW32_HI | (uint64_t)((saddr & 0x3FFFFu) >> 4);
}

// instruction descriptors: D=f32, K-major both, M=128 (lanes), N=128 (data rows);
// A=B=tf32 (format 2, kind::tf32) or A=B=bf16 (format 1, kind::f16)
constexpr int UMMA_M = 4 * NB;     // 256: 128 lanes in each of the two CTAs
constexpr uint32_t IDESC_SHAPE = (1u << 4) | ((uint32_t)(BM >> 3) << 17) | ((uint32_t)(UMMA_M >> 4) << 24);
constexpr uint32_t IDESC = IDESC_SHAPE | (2u << 7) | (2u << 10);
constexpr uint32_t IDESC_BF16 = IDESC_SHAPE | (1u << 7) | (1u << 10);

// Load/store epilogue of EPI_C64 for an ODD row length (the global row pitch is then not a
// multiple of 16 bytes and the TMA cannot be used), staged through shared memory so that the
// global accesses are coalesced.  Thread = TMEM lane: within warp q, lane 2j' carries
// Re(Out[n][m0 + c]) and lane 2j'+1 Im(...) of output row n = nq0 + j' for the 128 data rows c.
// In slices of 16 columns each lane drops its values (already scaled) into a [32 rows][16 cols]
// fp32 staging tile, and the warp reads it back with lanes running along m: real and imaginary
// parts of one element meet in one thread and leave as 8-byte complex stores.
constexpr int STG_COLS = 16;
constexpr int STG_PITCH = 20;                       // floats; conflict-free 128-bit row writes
constexpr int STG_BYTES = 32 * STG_PITCH * 4;       // 2560 B per warp
static_assert(STG_BYTES <= STG_WARP_BYTES, "staging");

__device__ __forceinline__ void tile_epilogue_c64_lsu(const GemmParams& p, int item, int nq0, int m0, int lane,
                                                      float (&tot)[BM], float* stg) {
  const float sc = p.scale ? __ldg(p.scale + item) : 1.0f;
  const int mmax = p.rows - m0;  // tile columns c < mmax are valid
  // lane -> output row j = lane/8 + 4*it (it = 0..3), complex pair at columns 2*(lane%8)
  const int cc = 2 * (lane & 7);
  const int j0 = lane >> 3;
  float2* out0 = p.out_c64 + ((size_t)item * p.n_out + nq0 + j0) * p.rows + m0 + cc;
#pragma unroll
  for (int s = 0; s < BM / STG_COLS; ++s) {
    const int c0 = s * STG_COLS;
#pragma unroll
    for (int v = 0; v < STG_COLS / 4; ++v)
      *reinterpret_cast<float4*>(stg + lane * STG_PITCH + 4 * v) =
          make_float4(tot[c0 + 4 * v] * sc, tot[c0 + 4 * v + 1] * sc, tot[c0 + 4 * v + 2] * sc,
                      tot[c0 + 4 * v + 3] * sc);
    __syncwarp();
    if (c0 < mmax) {  // warp-uniform
      float2 re[4], im[4];
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int j = j0 + 4 * it;
        re[it] = *reinterpret_cast<const float2*>(stg + (2 * j) * STG_PITCH + cc);
        im[it] = *reinterpret_cast<const float2*>(stg + (2 * j + 1) * STG_PITCH + cc);
      }
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        if (nq0 + j0 + 4 * it < p.n_out) {
          float2* out = out0 + (size_t)(4 * it) * p.rows + c0;
          if (c0 + cc < mmax) out[0] = make_float2(re[it].x, im[it].x);
          if (c0 + cc + 1 < mmax) out[1] = make_float2(re[it].y, im[it].y);
        }
      }
    }
    __syncwarp();
  }
}

// EPI_PLANES through TMA stores.  Thread = TMEM lane = one output row (coordinate j = lane / 2,
// real or imaginary part = lane % 2) holding that row's 128 columns: no transposition is needed
// -- each lane scales its own 16-column slice straight into the box layout of its plane
// (16 rows x 16 columns of float32, SWIZZLE_64B), and one lane hands the two boxes of the warp to
// the TMA, which writes them asynchronously and clips what lies beyond the matrix.  Staging per
// warp and slice (2 KiB): [Re 1 KiB | Im 1 KiB]; STG_BUFS slices in flight.
__device__ __forceinline__ void bulk_wait_read_n(int n) {
  if (n <= 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  else if (n == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
  else if (n == 2) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
  else asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
}
__device__ __forceinline__ void tile_epilogue_tma(const GemmParams& p, int item, int out_idx, int nq0, int m0, int lane,
                                                  float (&tot)[BM], uint32_t stg0,
                                                  const CUtensorMap* o_re, const CUtensorMap* o_im,
                                                  bool keep_in_l2 = false) {
  const float sc = p.scale ? __ldg(p.scale + item) : 1.0f;
  const int mmax = p.rows - m0;
  const int j = lane >> 1, part = lane & 1;
  const uint32_t sw64 = (uint32_t)((j >> 1) & 3);
  const uint64_t pol = keep_in_l2 ? l2_policy_evict_last() : 0;
#pragma unroll
  for (int s = 0; s < BM / 16; ++s) {
    const int c0 = s * 16;
    if (c0 >= mmax) break;  // warp-uniform: the rest of the tile lies beyond the matrix
    const uint32_t stg = stg0 + (s % STG_BUFS) * STG_SLICE_BYTES;
    const uint32_t row = stg + part * 1024 + j * 64;               // my row of my plane's box
    if (s >= STG_BUFS) {    // the TMA has read this buffer's previous slice
      if (lane == 0) bulk_wait_read_n(STG_BUFS - 1);
      __syncwarp();
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)   // (the re / im lanes of a pair hit the same banks: 2-way, cheap)
      st_shared_v4(row + (((uint32_t)c ^ sw64) << 4), __float_as_uint(tot[c0 + 4 * c] * sc),
                   __float_as_uint(tot[c0 + 4 * c + 1] * sc), __float_as_uint(tot[c0 + 4 * c + 2] * sc),
                   __float_as_uint(tot[c0 + 4 * c + 3] * sc));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // my writes, before the TMA reads them
    __syncwarp();
    if (lane == 0) {
      const int col = m0 + c0;
      if (keep_in_l2) {
        tma_store_3d_hint(o_re, stg, col, nq0, out_idx, pol);
        tma_store_3d_hint(o_im, stg + 1024, col, nq0, out_idx, pol);
      } else {
        tma_store_3d(o_re, stg, col, nq0, out_idx);
        tma_store_3d(o_im, stg + 1024, col, nq0, out_idx);
      }
      bulk_commit();
    }
  }
  if (lane == 0) bulk_wait_read();  // the staging buffers are free again (the next user may be the C64 path)
  __syncwarp();
}

// EPI_PSF: the image itself, psf[n][m] += w |scale D|^2, through TMA REDUCE-add stores -- no complex field
// leaves the chip (forward-only calls need none: optical_systems.py:213-223 / wavefronts.py:279 fused into
// the last contraction).  The Re lanes stage w re^2 in one 16 x 16 float32 box, the Im lanes w im^2 in a
// second one, and both are added onto the same tile of the image.  The additions of different items reach
// memory in no fixed order (float32 sums differ in the last bits from run to run).
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tile_epilogue_psf(const GemmParams& p, int item, int nq0, int m0, int lane,
                                                  float (&tot)[BM], uint32_t stg0, const CUtensorMap* o_psf) {
  const float sc = p.scale ? __ldg(p.scale + item) : 1.0f;
  const float w = __ldg(p.item_w + item);
  const int mmax = p.rows - m0;
  const int j = lane >> 1, part = lane & 1;
  const uint32_t sw64 = (uint32_t)((j >> 1) & 3);
#pragma unroll
  for (int s = 0; s < BM / 16; ++s) {
    const int c0 = s * 16;
    if (c0 >= mmax) break;
    const uint32_t stg = stg0 + (s % STG_BUFS) * STG_SLICE_BYTES;
    const uint32_t row = stg + part * 1024 + j * 64;
    if (s >= STG_BUFS) {
      if (lane == 0) bulk_wait_read_n(STG_BUFS - 1);
      __syncwarp();
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a = tot[c0 + 4 * c + e] * sc;
        v[e] = w * (a * a);
      }
      st_shared_v4(row + (((uint32_t)c ^ sw64) << 4), __float_as_uint(v[0]), __float_as_uint(v[1]),
                   __float_as_uint(v[2]), __float_as_uint(v[3]));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      tma_reduce_add_2d(o_psf, stg, m0 + c0, nq0);
      tma_reduce_add_2d(o_psf, stg + 1024, m0 + c0, nq0);
      bulk_commit();
    }
  }
  if (lane == 0) bulk_wait_read();
  __syncwarp();
}

// EPI_C64 through TMA stores (needs an even row length: the global row pitch must be a multiple
// of 16 bytes).  The box of a warp and slice is 16 rows x 16 complex = 128-byte rows, SWIZZLE_128B;
// the real lane of a row writes the even words, the imaginary lane the odd ones.  Four 2 KiB
// buffers per warp: up to three slices in flight.
constexpr int C64_SLICE_BYTES = 2048;
constexpr int C64_BUFS = STG_WARP_BYTES / C64_SLICE_BYTES;   // 4 (2 with the 6-stage ring)
__device__ __forceinline__ void tile_epilogue_c64_tma(const GemmParams& p, int item, int nq0, int m0, int lane,
                                                      float (&tot)[BM], uint32_t stg0, const CUtensorMap* o_c) {
  const float sc = p.scale ? __ldg(p.scale + item) : 1.0f;
  const int mmax = p.rows - m0;
  const int j = lane >> 1, part = lane & 1;
  const uint32_t row = (uint32_t)j * 128 + (uint32_t)part * 4;
  const uint32_t sw = (uint32_t)(j & 7);
#pragma unroll
  for (int s = 0; s < BM / 16; ++s) {
    const int c0 = s * 16;
    if (c0 >= mmax) break;  // warp-uniform
    const uint32_t stg = stg0 + (s % C64_BUFS) * C64_SLICE_BYTES;
    if (s >= C64_BUFS) {    // the TMA has read this buffer's previous slice
      if (lane == 0) bulk_wait_read_n(C64_BUFS - 1);
      __syncwarp();
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const uint32_t a = stg + row + ((((uint32_t)(c >> 1)) ^ sw) << 4) + (uint32_t)(c & 1) * 8;
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(__float_as_uint(tot[c0 + c] * sc)) : "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      tma_store_3d(o_c, stg, 2 * (m0 + c0), nq0, item);
      bulk_commit();
    }
  }
  if (lane == 0) bulk_wait_read();
  __syncwarp();
}

// On-chip operand split.  One converter warp owns one plane (Re or Im) of one tile: this CTA's 64 rows
// x 16 k of it, as the TMA left them (float32, SWIZZLE_64B).  Work item = (row, k-half): 8 values -> hi =
// rna_tf32 written back in place, bf16(hi) and bf16(v - hi) as one 16-byte chunk each of the SWIZZLE_32B
// planes.  128 items per plane = 4 per lane, loaded two at a time; every quarter-warp touches 8 distinct
// 16-byte bank groups.
__device__ __forceinline__ void convert_plane(uint8_t* fplane, uint8_t* bplane_hi, int lane) {
  const uint32_t h = (uint32_t)lane & 1u;
  const int rl = lane >> 1;
#pragma unroll
  for (int it2 = 0; it2 < 2; ++it2) {
    float4 v[2][2];
    float4* src[2][2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int r = (it2 * 2 + u) * 16 + rl;
      const uint32_t sw64 = (uint32_t)((r >> 1) & 3);
      uint8_t* frow = fplane + r * 64;
      src[u][0] = reinterpret_cast<float4*>(frow + (((2u * h) ^ sw64) << 4));
      src[u][1] = reinterpret_cast<float4*>(frow + (((2u * h + 1u) ^ sw64) << 4));
      v[u][0] = *src[u][0];
      v[u][1] = *src[u][1];
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int r = (it2 * 2 + u) * 16 + rl;
      const uint32_t sw32 = (uint32_t)((r >> 2) & 1);
      uint4 ph, plo;
      float4 h0, h1;
      h0.x = tf32_hi(v[u][0].x); h0.y = tf32_hi(v[u][0].y); h0.z = tf32_hi(v[u][0].z); h0.w = tf32_hi(v[u][0].w);
      h1.x = tf32_hi(v[u][1].x); h1.y = tf32_hi(v[u][1].y); h1.z = tf32_hi(v[u][1].z); h1.w = tf32_hi(v[u][1].w);
      ph.x = pack_bf16(h0.x, h0.y); ph.y = pack_bf16(h0.z, h0.w);
      ph.z = pack_bf16(h1.x, h1.y); ph.w = pack_bf16(h1.z, h1.w);
      plo.x = pack_bf16(v[u][0].x - h0.x, v[u][0].y - h0.y); plo.y = pack_bf16(v[u][0].z - h0.z, v[u][0].w - h0.w);
      plo.z = pack_bf16(v[u][1].x - h1.x, v[u][1].y - h1.y); plo.w = pack_bf16(v[u][1].z - h1.z, v[u][1].w - h1.w);
      *src[u][0] = h0;
      *src[u][1] = h1;
      uint8_t* brow = bplane_hi + r * 32 + ((h ^ sw32) << 4);
      *reinterpret_cast<uint4*>(brow) = ph;
      *reinterpret_cast<uint4*>(brow + BPLANE_BYTES) = plo;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // my writes, before the MMA reads them
}

// The 8 MMAs of one tile and one 16-k chunk:
//   D (+)= G1_hi * Re_hi^T + G2_hi * Im_hi^T                                  (tf32, two k-steps)
//        + G1_hi * Re_lo^T + G1_lo * Re_hi^T + G2_hi * Im_lo^T + G2_lo * Im_hi^T   (bf16, K = 16)
// `tile_smem`: shared address of the tile-chunk's planes; `g0`: TMEM address of the phasor stage.
__device__ __forceinline__ void issue_tile_chunk(uint32_t d, uint32_t g0, uint32_t tile_smem, bool fresh) {
  constexpr uint64_t KS = (UMMA_K * 4) >> 4;  // descriptor step per k-step inside the 64-byte swizzle row
  const uint64_t d_rh = make_desc_sw64(tile_smem), d_ih = make_desc_sw64(tile_smem + PLANE_BYTES);
  const uint64_t b_rh = make_desc_sw32(tile_smem + BPL_BASE + 0 * BPLANE_BYTES);
  const uint64_t b_rl = make_desc_sw32(tile_smem + BPL_BASE + 1 * BPLANE_BYTES);
  const uint64_t b_ih = make_desc_sw32(tile_smem + BPL_BASE + 2 * BPLANE_BYTES);
  const uint64_t b_il = make_desc_sw32(tile_smem + BPL_BASE + 3 * BPLANE_BYTES);
  const uint32_t gb = g0 + GB_BASE;  // packed bf16: G1_hi, G1_lo, G2_hi, G2_lo
  // small terms first
  umma_bf16_ts2(d, gb + 1 * GB_COLS, b_rh, IDESC_BF16, fresh ? 0u : 1u);  // G1_lo * Re_hi
  umma_bf16_ts2(d, gb + 0 * GB_COLS, b_rl, IDESC_BF16, 1u);                // G1_hi * Re_lo
  umma_bf16_ts2(d, gb + 3 * GB_COLS, b_ih, IDESC_BF16, 1u);                // G2_lo * Im_hi
  umma_bf16_ts2(d, gb + 2 * GB_COLS, b_il, IDESC_BF16, 1u);                // G2_hi * Im_lo
#pragma unroll
  for (int ks = 0; ks < BK / UMMA_K; ++ks) {
    umma_tf32_ts2(d, g0 + ks * UMMA_K, d_rh + ks * KS, IDESC, 1u);        // G1_hi * Re_hi
    umma_tf32_ts2(d, g0 + BK + ks * UMMA_K, d_ih + ks * KS, IDESC, 1u);   // G2_hi * Im_hi
  }
}

// Where the TMEM partial accumulators of the two tiles open and close along the k-chunks of
// a unit.  The FIRST partial of a unit is two (three for K >= 768) flush periods long: while
// the drain warpgroups are still writing the previous unit's epilogue (8-12 chunk-times,
// bounded by the SM's store port) the issuers must not need a buffer that only those
// warpgroups can release.
//   tile a: [0,12) [12,16) [16,20) ...      tile b: [0,10) [10,14) [14,18) ...
// (steady state staggered by half a partial).  Both close at the last chunk.  Acquisition
// order inside a chunk: a, then b.
struct PartialSchedule {
  bool a_open, a_close, b_open, b_close;
};
#ifndef DLUX_LONG_FIRST_MIN
#define DLUX_LONG_FIRST_MIN (12 * FLUSH_CHUNKS)
#endif
#ifndef DLUX_LONG_FIRST_FACTOR
#define DLUX_LONG_FIRST_FACTOR 3
#endif
// partial (= TMEM accumulation group) that k-chunk `kc` of a K loop of `k_chunks` chunks belongs to
__device__ __forceinline__ int partial_group(int kc, int first) {
  return kc < first ? 0 : 1 + (kc - first) / FLUSH_CHUNKS;
}
// Dense K loop: closed form (the MMA issuer's loop is latency-critical; this is the cheapest statement).
//   tile a: [0,A0) [A0,A0+4) ...      tile b: [0,B0) [B0,B0+4) ...   (B0 = A0 - 2); both close at the last chunk.
__device__ __forceinline__ PartialSchedule partial_schedule_dense(int kc, int k_chunks) {
  const int A0 = (k_chunks >= DLUX_LONG_FIRST_MIN ? DLUX_LONG_FIRST_FACTOR : 2) * FLUSH_CHUNKS,
            B0 = A0 - FLUSH_CHUNKS / 2;
  PartialSchedule ps;
  const bool last = kc == k_chunks - 1;
  const int ra = (kc - A0) % FLUSH_CHUNKS, rb = (kc - B0) % FLUSH_CHUNKS;  // used only for kc >= A0 / B0
  ps.a_open = (kc == 0) || (kc >= A0 && ra == 0);
  ps.a_close = last || (kc == A0 - 1) || (kc >= A0 && ra == FLUSH_CHUNKS - 1);
  ps.b_open = (kc == 0) || (kc >= B0 && rb == 0);
  ps.b_close = last || (kc == B0 - 1) || (kc >= B0 && rb == FLUSH_CHUNKS - 1);
  return ps;
}
// `prev` / `next`: the k-chunks processed before / after `kc` in this unit (-1: none).  Dense K loops pass
// kc - 1 / kc + 1; with zero-block skipping the list has holes, and because the groups are defined on the
// ORIGINAL chunk index every partial sums exactly the chunks the dense kernel would have summed into it
// (minus exact zeros): the result is bit-identical.
__device__ __forceinline__ PartialSchedule partial_schedule(int prev, int kc, int next, int k_chunks) {
  // first partial: 3 (long K) or 2 flush periods for tile a, half a period less for tile b
  const int A0 = (k_chunks >= DLUX_LONG_FIRST_MIN ? DLUX_LONG_FIRST_FACTOR : 2) * FLUSH_CHUNKS,
            B0 = A0 - FLUSH_CHUNKS / 2;
  PartialSchedule ps;
  const int ga = partial_group(kc, A0), gb = partial_group(kc, B0);
  ps.a_open = prev < 0 || partial_group(prev, A0) != ga;
  ps.a_close = next < 0 || partial_group(next, A0) != ga;
  ps.b_open = prev < 0 || partial_group(prev, B0) != gb;
  ps.b_close = next < 0 || partial_group(next, B0) != gb;
  return ps;
}

struct TcParams {
  GemmParams g;
  int tiles_mp, tiles_np, n_units, k_chunks;  // pairs of 128-row data tiles, pairs of 64-column n-tiles
  int c64_tma;                                // EPI_C64 output goes through TMA stores (even row length)
  int units_per_item;                         // tiles_mp * tiles_np, or the length of unit_list
  int kvec_vec4;                              // kvec rows can be read four coordinates (16 bytes) at a time
};

// The K loop of one unit: dense (0 .. k_chunks-1) or, with exact zero-block skipping, the list of k-chunks
// in which the unit's 256 data rows hold anything but zeros (GemmParams::chunk_cnt / chunk_idx).
// (SPARSE is a compile-time switch: the dense instantiation carries none of the list logic)
template <bool SPARSE>
struct ChunkWalk {
  const int* idx;
  int n;
  __device__ __forceinline__ ChunkWalk(const TcParams& tp, int mt) {
    const GemmParams& p = tp.g;
    idx = (SPARSE && p.chunk_idx) ? p.chunk_idx + (size_t)mt * tp.k_chunks : nullptr;
    n = (SPARSE && p.chunk_cnt) ? __ldg(p.chunk_cnt + mt) : tp.k_chunks;
  }
  __device__ __forceinline__ int at(int ci) const { return (SPARSE && idx) ? __ldg(idx + ci) : ci; }
  __device__ __forceinline__ int next_of(int ci) const { return ci + 1 < n ? at(ci + 1) : -1; }
};
// unit -> (item, tile index t): dense, or through the list of units whose output block is needed
template <bool SPARSE>
__device__ __forceinline__ void decode_unit(const TcParams& tp, int units_per_item, int unit, int& item, int& t) {
  item = unit / units_per_item;
  t = unit - item * units_per_item;
  if (SPARSE && tp.g.unit_list) t = __ldg(tp.g.unit_list + t);
}

// Two chained stages in ONE persistent launch (FUSED): the stage-1 -> stage-2 intermediate of an item lives in a
// ring of `ring` slots that are rewritten every `ring` items, so it stays L2-resident and never needs HBM.
// Work list (every cluster walks it in order, unit = cluster + i * #clusters):
//   S1(0) .. S1(lag-1) | S1(lag) S2(0) | S1(lag+1) S2(1) | ... | S2(n-lag) .. S2(n-1)        (S = all units of an item)
// Dependencies are per item and point BACKWARDS in the list only (lag < ring), so in-order persistent clusters
// cannot deadlock:  S2(i) loads after ready[i] == all S1(i) stores completed;  S1(i) stores after
// consumed[i - ring] == all S2(i - ring) units finished.
struct FusedTc {
  TcParams s[2];
  int fused;                 // 0: a single stage, s[0]
  int n_items, lag, ring, n_units_total;
  int* ready;                // [n_items], zeroed before the launch
  int* consumed;             // [n_items]
  int ready_target, consumed_target;
};
__device__ __forceinline__ void decode_fused(const FusedTc& f, int unit, int& stage, int& item, int& t) {
  const int n1 = f.s[0].units_per_item, n2 = f.s[1].units_per_item;
  const int head = f.lag * n1;
  if (unit < head) {
    stage = 0; item = unit / n1; t = unit - item * n1;
    return;
  }
  int u = unit - head;
  const int grp = n1 + n2, n_mid = f.n_items - f.lag, mid = n_mid * grp;
  if (u < mid) {
    const int g = u / grp, r = u - g * grp;
    if (r < n1) { stage = 0; item = f.lag + g; t = r; }
    else { stage = 1; item = g; t = r - n1; }
    return;
  }
  u -= mid;
  stage = 1; item = n_mid + u / n2; t = u % n2;
}
__device__ __forceinline__ void wait_counter(const int* ctr, int target) {
  int v;
  do {
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    if (v < target) __nanosleep(100);
  } while (v < target);
}
__device__ __forceinline__ void publish_counter(int* ctr) {
  __threadfence();
  asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(ctr) : "memory");
}

// one generator warp's share of a phasor stage: k-step KS of the two tf32 planes and of the four packed-bf16 planes
template <int KS>
__device__ __forceinline__ void store_phasors(uint32_t g0, const float (&g1h)[8], const float (&g2h)[8],
                                              const uint32_t (&pk)[4][4]) {
  tmem_st8(g0 + KS * UMMA_K, g1h);        // tf32 G1_hi: columns [0,16)
  tmem_st8(g0 + BK + KS * UMMA_K, g2h);   // tf32 G2_hi: columns [16,32)
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) tmem_st4u(g0 + GB_BASE + q4 * GB_COLS + KS * (UMMA_K / 2), pk[q4]);
}

// four consecutive coordinates kv[k .. k+3] (zero past K)
__device__ __forceinline__ void load_k4(const float* __restrict__ kv, int k, int K, bool vec4, float (&x)[4]) {
  if (vec4) {
    const float4 v = (k < K) ? __ldg(reinterpret_cast<const float4*>(kv + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = (k + j < K) ? __ldg(kv + k + j) : 0.0f;
  }
}

template <bool SPARSE, bool DFT, bool FUSED>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
               const __grid_constant__ CUtensorMap omap0, const __grid_constant__ CUtensorMap omap1,
               const __grid_constant__ CUtensorMap omapc, const __grid_constant__ CUtensorMap map2,
               const __grid_constant__ CUtensorMap map3, const FusedTc ftp) {
  extern __shared__ uint8_t smem_raw[];
  const TcParams& tp = ftp.s[0];     // (FUSED: every unit loop re-binds tp / p to its unit's stage)
  int units_per_item = tp.units_per_item, n_units = FUSED ? ftp.n_units_total : tp.n_units;
  if (SPARSE && tp.g.unit_list) {   // only the needed output blocks: the list length lives on the device
    units_per_item = __ldg(tp.g.unit_count);
    n_units = units_per_item * tp.g.n_items;
  }
  const GemmParams& p = tp.g;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));  // generic view of the aligned base
  const uint32_t bar_base = smem_base + RING_BYTES;
  auto fullA_bar = [&](int s) { return bar_base + 8u * s; };
  auto emptyA_bar = [&](int s) { return bar_base + 8u * (A_STAGES + s); };
  auto fullG_bar = [&](int s) { return bar_base + 8u * (2 * A_STAGES + s); };
  auto emptyG_bar = [&](int s) { return bar_base + 8u * (2 * A_STAGES + G_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * A_STAGES + 2 * G_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * A_STAGES + 2 * G_STAGES + NUM_ACC + a); };
  auto rawfull_bar = [&](int s) { return bar_base + 8u * (2 * A_STAGES + 2 * G_STAGES + 2 * NUM_ACC + s); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(
      smem_gen + RING_BYTES + 8 * (3 * A_STAGES + 2 * G_STAGES + 2 * NUM_ACC));

  // (the broadcast tells ptxas that `warp` is warp-uniform: role dispatch, tensor-memory addresses and barrier
  // addresses derived from it then live in uniform registers)
  const int warp = __shfl_sync(0xFFFFFFFFu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  if (warp == WARP_TMA && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map0));
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map1));
    if (FUSED) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map2));
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map3));
    }
    if (tp.g.mode == EPI_PLANES) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&omap0));
      asm volatile("prefetch.tensormap [%0];" ::"l"(&omap1));
    }
    if (FUSED || tp.c64_tma || tp.g.mode == EPI_PSF) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&omapc));
    }
  }
  if (warp == WARP_MMA && lane == 0) {
    for (int s = 0; s < A_STAGES; ++s) {
      mbar_init(fullA_bar(s), 2 * NUM_CONV_WARPS);   // [leader's] the converter warps of both CTAs
      mbar_init(emptyA_bar(s), 2);  // multicast tcgen05.commit of the two issuer warps
      mbar_init(rawfull_bar(s), 1); // my TMA producer's arrive.expect_tx (+ the bytes of my four loads)
    }
    for (int s = 0; s < G_STAGES; ++s) {
      mbar_init(fullG_bar(s), 2 * NUM_GEN_WARPS);  // [leader's] one arrive per generator warp of both CTAs
      mbar_init(emptyG_bar(s), 2);                 // multicast tcgen05.commit of the two issuer warps
    }
    for (int a = 0; a < NUM_ACC; ++a) {
      mbar_init(tfull_bar(a), 1);     // multicast tcgen05.commit closing a partial
      mbar_init(tempty_bar(a), 256);  // [leader's] every thread of the draining warpgroup of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 0) {
    // the same warp of both CTAs: one allocation spanning the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t crank = cluster_ctarank();
  const int cl_id = blockIdx.x / CLUSTER, n_cl = gridDim.x / CLUSTER;
  // fullA / fullG / tempty are the LEADER's (cluster rank 0); everyone addresses them through
  // the cluster window, the leader's own threads included
  const uint32_t lead_delta = mapa_rank(bar_base, 0) - bar_base;

  // cluster work unit = (item, pair of n-tiles, pair of m-tiles); CTA `crank` of the cluster
  // takes n-tile 2 * np + crank (an n-tile beyond the matrix computes on zeros and stores
  // nothing)

  // Register re-balancing between warpgroups: setmaxnreg is the first instruction of each
  // warpgroup's branch (all four warps of a warpgroup execute it).
  if (warp >= NUM_EPI_WARPS && warp < FIRST_GEN_WARP) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_CTRL));
    if (warp == WARP_TMA) {
      // ===================== TMA producer =====================
      // Four loads per chunk (Re / Im of tiles a and b, my 64 rows of each) onto my own rawfull
      // barrier, which wakes this CTA's converter warps.
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t pol_ring = FUSED ? l2_policy_evict_last() : 0;
      for (int unit = cl_id; unit < n_units; unit += n_cl) {
        int ustage = 0, item, t;
        if (FUSED) decode_fused(ftp, unit, ustage, item, t);
        else decode_unit<SPARSE>(tp, units_per_item, unit, item, t);
        const TcParams& tp = ftp.s[FUSED ? ustage : 0];
        const GemmParams& p = tp.g;
        const CUtensorMap* m_re = (FUSED && ustage) ? &map2 : &map0;
        const CUtensorMap* m_im = (FUSED && ustage) ? &map3 : &map1;
        const int m0 = (t % tp.tiles_mp) * (2 * BM);
        int d = p.item_data ? __ldg(p.item_data + item) : item;
        if (FUSED && ustage) {
          d = item % ftp.ring;                          // my data = the ring slot stage 1 filled for this item
          wait_counter(ftp.ready + item, ftp.ready_target);
          asm volatile("fence.proxy.async;" ::: "memory");   // the acquire above, before my async-proxy (TMA) reads
        }
        const ChunkWalk<SPARSE> cw(tp, t % tp.tiles_mp);
        for (int ci = 0; ci < cw.n; ++ci) {
          const int kc = cw.at(ci);
          mbar_wait(emptyA_bar(stage), phase ^ 1);
          if (elect_one()) {
            const uint32_t dst = smem_base + stage * A_BYTES;
            const uint32_t bar = rawfull_bar(stage);
            mbar_arrive_expect_tx(bar, RAW_BYTES);
#pragma unroll
            for (int tb2 = 0; tb2 < 2; ++tb2) {   // (rows beyond the matrix are zero-filled)
              const uint32_t t0 = dst + tb2 * TILE_BYTES;
              const int mr = m0 + tb2 * BM + (int)crank * ROWS_CTA;
              if (FUSED && ustage) {   // the ring: keep it in L2 until its slot is rewritten
                tma_load_3d_hint(t0 + 0 * PLANE_BYTES, m_re, bar, kc * BK, mr, d, pol_ring);
                tma_load_3d_hint(t0 + 1 * PLANE_BYTES, m_im, bar, kc * BK, mr, d, pol_ring);
              } else {
                tma_load_3d(t0 + 0 * PLANE_BYTES, m_re, bar, kc * BK, mr, d);   // Re
                tma_load_3d(t0 + 1 * PLANE_BYTES, m_im, bar, kc * BK, mr, d);   // Im
              }
            }
          }
          __syncwarp();
          if (++stage == A_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == WARP_MMA || warp == WARP_MMA + 1) {
      // ===================== MMA issuers: warp WARP_MMA -> tile a, WARP_MMA + 1 -> tile b =====================
      // One issuing warp per tile halves the (scalar, latency-bound) instruction stream that
      // sits between consecutive tcgen05.mma's.  The whole warp walks the loop (uniform control
      // flow); one elected lane issues.  Both warps mirror the same acquisition sequence of the
      // three TMEM partial buffers (order inside a chunk: a, then b).
      const int which = warp - WARP_MMA;
      int sa = 0, sg = 0;
      uint32_t pa = 0, pg = 0;
      uint32_t nbuf = 0, nphase = 0;      // next buffer to acquire and the parity of its use count
      uint32_t mybuf = 0, myphase = 0;    // my open partial
#ifdef DLUX_DEBUG_TIMING
      long long dbg_g = 0, dbg_a = 0, dbg_t = 0;
      const long long dbg_start = clock64();
#endif
      // pair mode: only the leader CTA issues; its MMAs drive the tensor cores of both SMs
      for (int unit = (crank != 0) ? n_units : cl_id; unit < n_units; unit += n_cl) {
        int n_walk = tp.k_chunks, prev = -1, kc = 0;
        const int* widx = nullptr;
        int k_chunks_u = tp.k_chunks;
        if (FUSED) {
          int ustage, item_, t_;
          decode_fused(ftp, unit, ustage, item_, t_);
          n_walk = k_chunks_u = ftp.s[ustage].k_chunks;
        }
        if (SPARSE) {
          int item_, t_;
          decode_unit<SPARSE>(tp, units_per_item, unit, item_, t_);
          const ChunkWalk<SPARSE> cw(tp, t_ % tp.tiles_mp);
          n_walk = cw.n;
          widx = cw.idx;
          kc = cw.n > 0 ? cw.at(0) : -1;
        }
        for (int ci = 0; ci < n_walk; ++ci) {
          // Partials are FLUSH_CHUNKS long; tile b's boundaries are staggered by half a partial so
          // that the three TMEM buffers are re-acquired >= 2 chunks after they were handed to a
          // drain warpgroup (see partial_schedule()).
          PartialSchedule ps;
          if (SPARSE) {
            const int next = ci + 1 < n_walk ? (widx ? __ldg(widx + ci + 1) : ci + 1) : -1;
            ps = partial_schedule(prev, kc, next, tp.k_chunks);
            prev = kc;
            kc = next;
          } else {
            ps = partial_schedule_dense(ci, k_chunks_u);
          }
          bool opened = false;
          if (ps.a_open) {
            if (which == 0) { mybuf = nbuf; myphase = nphase; opened = true; }
            if (++nbuf == NUM_ACC) { nbuf = 0; nphase ^= 1; }
          }
          if (ps.b_open) {
            if (which == 1) { mybuf = nbuf; myphase = nphase; opened = true; }
            if (++nbuf == NUM_ACC) { nbuf = 0; nphase ^= 1; }
          }
#ifdef DLUX_DEBUG_TIMING
          const long long t0_ = clock64();
          mbar_wait_cluster(fullG_bar(sg), pg);
          const long long t1_ = clock64();
          mbar_wait_cluster(fullA_bar(sa), pa);
          const long long t2_ = clock64();
          if (opened) mbar_wait_cluster(tempty_bar(mybuf), myphase ^ 1);
          const long long t3_ = clock64();
          dbg_g += t1_ - t0_; dbg_a += t2_ - t1_; dbg_t += t3_ - t2_;
#else
          // (acquire at cluster scope: the peer CTA's warps arrive on these too)
          mbar_wait_cluster(fullG_bar(sg), pg);
          mbar_wait_cluster(fullA_bar(sa), pa);
          if (opened) mbar_wait_cluster(tempty_bar(mybuf), myphase ^ 1);  // the drain warpgroup(s) released the buffer
#endif
          tc_fence_after();
          if (elect_one()) {
            issue_tile_chunk(tmem_base + mybuf * ACC_COLS, tmem_base + (uint32_t)(G_BASE_COL + sg * G_COLS),
                             smem_base + sa * A_BYTES + which * TILE_BYTES, opened);
            // when these MMAs retire: smem slot released in both CTAs, phasor stage released
            // (each barrier also counts the other issuer warp's commit), partial handed over
            constexpr uint16_t BOTH = (1u << CLUSTER) - 1;
            umma_commit_mc2(emptyA_bar(sa), BOTH);
            umma_commit_mc2(emptyG_bar(sg), BOTH);
            if (which ? ps.b_close : ps.a_close) umma_commit_mc2(tfull_bar(mybuf), BOTH);
          }
          __syncwarp();
          if (++sa == A_STAGES) { sa = 0; pa ^= 1; }
          if (++sg == G_STAGES) { sg = 0; pg ^= 1; }
        }
      }
#ifdef DLUX_DEBUG_TIMING
      if (blockIdx.x == 0 && lane == 0)
        printf("MMA%d K=%d rows=%d: total %lld  wait fullG %lld  fullA %lld  tempty %lld\n", which, p.K, p.rows,
               clock64() - dbg_start, dbg_g, dbg_a, dbg_t);
#endif
    }
  } else if (warp < NUM_EPI_WARPS) {
    // ===================== drain + epilogue (WG0: tile a, WG1: tile b) =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_EPI));
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int which = warp >> 2;     // 0: tile a, 1: tile b
    uint32_t nbuf = 0, nphase = 0, mybuf = 0, myphase = 0;  // mirrors the MMA issuers' acquisition sequence
#ifdef DLUX_DEBUG_TIMING
    long long dbg_w = 0, dbg_e = 0, dbg_n = 0;
    const long long dbg_start = clock64();
#endif
    float* stg = reinterpret_cast<float*>(smem_gen + RING_BYTES + BAR_BYTES + warp * STG_WARP_BYTES);
    const uint32_t stg_addr = smem_base + RING_BYTES + BAR_BYTES + warp * STG_WARP_BYTES;
    for (int unit = cl_id; unit < n_units; unit += n_cl) {
      int ustage = 0, item, t;
      if (FUSED) decode_fused(ftp, unit, ustage, item, t);
      else decode_unit<SPARSE>(tp, units_per_item, unit, item, t);
      const TcParams& tp = ftp.s[FUSED ? ustage : 0];
      const GemmParams& p = tp.g;
      const int m0 = (t % tp.tiles_mp) * (2 * BM) + which * BM;
      const int nq0 = ((t / tp.tiles_mp) * CLUSTER + (int)crank) * NB + q * 16;  // first output row of this warp's lane quarter
      float tot[BM];
#pragma unroll
      for (int j = 0; j < BM; ++j) tot[j] = 0.0f;
      const ChunkWalk<SPARSE> cw(tp, t % tp.tiles_mp);     // (tp = this unit's stage)
      int prev = -1, kc = (SPARSE && cw.n > 0) ? cw.at(0) : -1;
      for (int ci = 0; ci < cw.n; ++ci) {
        PartialSchedule ps;
        if (SPARSE) {
          const int next = cw.next_of(ci);
          ps = partial_schedule(prev, kc, next, tp.k_chunks);
          prev = kc;
          kc = next;
        } else {
          ps = partial_schedule_dense(ci, tp.k_chunks);
        }
        if (ps.a_open) {
          if (which == 0) { mybuf = nbuf; myphase = nphase; }
          if (++nbuf == NUM_ACC) { nbuf = 0; nphase ^= 1; }
        }
        if (ps.b_open) {
          if (which == 1) { mybuf = nbuf; myphase = nphase; }
          if (++nbuf == NUM_ACC) { nbuf = 0; nphase ^= 1; }
        }
        if (!(which ? ps.b_close : ps.a_close)) continue;
        const uint32_t buf = mybuf;
#ifdef DLUX_DEBUG_TIMING
        const long long tw0_ = clock64();
        mbar_wait(tfull_bar(buf), myphase);
        dbg_w += clock64() - tw0_;
#else
        mbar_wait(tfull_bar(buf), myphase);
#endif
        tc_fence_after();
        const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + buf * ACC_COLS;
#pragma unroll
        for (int c = 0; c < BM; c += 16) {
          uint32_t v0[16];
          tmem_ld16(t0 + c, v0);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) tot[c + j] += __uint_as_float(v0[j]);
        }
        tc_fence_before();
        mbar_arrive_cluster(tempty_bar(buf) + lead_delta);
      }
#ifdef DLUX_DEBUG_TIMING
      const long long te0_ = clock64();
#endif
      if (FUSED && ustage == 0 && item >= ftp.ring) {   // the ring slot's previous tenant has been consumed
        if (lane == 0) wait_counter(ftp.consumed + (item - ftp.ring), ftp.consumed_target);
        __syncwarp();
      }
      if (m0 < p.rows) {  // warp-uniform condition
        if (p.mode == EPI_PLANES)
          tile_epilogue_tma(p, item, FUSED ? item % ftp.ring : item, nq0, m0, lane, tot, stg_addr, &omap0, &omap1,
                            FUSED);
        else if (p.mode == EPI_PSF)
          tile_epilogue_psf(p, item, nq0, m0, lane, tot, stg_addr, &omapc);
        else if (tp.c64_tma)
          tile_epilogue_c64_tma(p, item, nq0, m0, lane, tot, stg_addr, &omapc);
        else
          tile_epilogue_c64_lsu(p, item, nq0, m0, lane, tot, stg);
      }
      if (FUSED && lane == 0) {
        if (ustage == 0) {      // my share of the intermediate is in memory: one of the ready[item] arrivals
          bulk_wait_all();
          asm volatile("fence.proxy.async;" ::: "memory");
          publish_counter(ftp.ready + item);
        } else {                // this unit no longer reads its ring slot
          publish_counter(ftp.consumed + item);
        }
      }
#ifdef DLUX_DEBUG_TIMING
      dbg_e += clock64() - te0_; ++dbg_n;
#endif
    }
    if (lane == 0) bulk_wait_all();  // my bulk stores have completed before the CTA retires
#ifdef DLUX_DEBUG_TIMING
    if (blockIdx.x == 0 && lane == 0 && (warp & 3) == 0)
      printf("DRAIN%d: total %lld  wait tfull %lld  epilogue %lld  units %lld\n", which, clock64() - dbg_start, dbg_w,
             dbg_e, dbg_n);
#endif
  } else if (warp >= FIRST_CONV_WARP) {
    // ===================== operand converters (WG5) =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_CONV));
    const int cw = warp - FIRST_CONV_WARP;          // tile = cw / 2, plane = cw % 2
    const uint32_t off_f = (uint32_t)((cw >> 1) * TILE_BYTES + (cw & 1) * PLANE_BYTES);
    const uint32_t off_b = (uint32_t)((cw >> 1) * TILE_BYTES + BPL_BASE + 2 * (cw & 1) * BPLANE_BYTES);
    int cstage = 0;
    uint32_t cphase = 0;
#ifdef DLUX_DEBUG_TIMING
    long long dbg_cw = 0, dbg_cc = 0;
    const long long dbg_cstart = clock64();
#endif
    for (int unit = cl_id; unit < n_units; unit += n_cl) {
      int n_walk = tp.k_chunks;
      if (FUSED) {
        int ustage, item_, t_;
        decode_fused(ftp, unit, ustage, item_, t_);
        n_walk = ftp.s[ustage].k_chunks;
      }
      if (SPARSE) {
        int item_, t_;
        decode_unit<SPARSE>(tp, units_per_item, unit, item_, t_);
        n_walk = ChunkWalk<SPARSE>(tp, t_ % tp.tiles_mp).n;
      }
      for (int ci = 0; ci < n_walk; ++ci) {
#ifdef DLUX_DEBUG_TIMING
        const long long tc0_ = clock64();
#endif
        mbar_wait(rawfull_bar(cstage), cphase);
#ifdef DLUX_DEBUG_TIMING
        const long long tc1_ = clock64();
#endif
        uint8_t* sb = smem_gen + cstage * A_BYTES;
        convert_plane(sb + off_f, sb + off_b, lane);
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(fullA_bar(cstage) + lead_delta);
#ifdef DLUX_DEBUG_TIMING
        dbg_cw += tc1_ - tc0_; dbg_cc += clock64() - tc1_;
#endif
        if (++cstage == A_STAGES) { cstage = 0; cphase ^= 1; }
      }
    }
#ifdef DLUX_DEBUG_TIMING
    if (blockIdx.x == 0 && lane == 0 && cw == 0)
      printf("CONV: total %lld  wait rawfull %lld  convert+arrive %lld\n", clock64() - dbg_cstart, dbg_cw, dbg_cc);
#endif
  } else {
    // ===================== phasor generators =====================
    // Two warps per TMEM lane quarter: WG3 produces k-step 0 (k 0..7) of every chunk, WG4
    // k-step 1.  Lane 2j (Re row of phasor column j) needs (G1, G2) = (cos, -sin), lane 2j+1
    // (Im row) needs (sin, cos) = (cos, -sin) of the angle minus a quarter turn (an exact
    // quadrant shift inside the sincos).
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_GEN));
    const int q = warp & 3;                          // TMEM lane quarter
    const int ks = (warp - FIRST_GEN_WARP) >> 2;     // which k-step of the chunk
    const float dft_inv = DFT ? 1.0f / p.dft_period : 0.0f;
    const int jcol = (q * 32 + lane) >> 1;           // phasor column within the tile
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    int stage = 0;
    uint32_t phase = 0;
#ifdef DLUX_DEBUG_TIMING
    long long dbg_gw = 0;
    const long long dbg_gstart = clock64();
#endif
    for (int unit = cl_id; unit < n_units; unit += n_cl) {
      int ustage = 0, item, t;
      if (FUSED) decode_fused(ftp, unit, ustage, item, t);
      else decode_unit<SPARSE>(tp, units_per_item, unit, item, t);
      const TcParams& tp = ftp.s[FUSED ? ustage : 0];
      const GemmParams& p = tp.g;
      const ChunkWalk<SPARSE> cw(tp, t % tp.tiles_mp);
      const int n = ((t / tp.tiles_mp) * CLUSTER + (int)crank) * NB + jcol;
      const float* kv = p.kvec + (size_t)item * p.kvec_stride;
      const float u = (n < p.n_out) ? __ldg(p.nvec + (size_t)item * p.nvec_stride + n) : 0.0f;
      // The two lanes of a pair (2j, 2j+1) need the same 8 phasors in different roles: each
      // evaluates 4 of them (even lane k 0..3 of the k-step, odd lane k 4..7) and passes the
      // other lane what it needs with one shuffle per value.
      const int half = lane & 1;
      // four consecutive k per lane: one 16-byte load when the coordinate rows allow it (K and the row pitch
      // multiples of 4, checked on the host), scalar loads otherwise
      const bool kv4 = tp.kvec_vec4 != 0;
      const int kofs = ks * UMMA_K + half * 4;
      float xk[4];  // my four k coordinates of this chunk, prefetched one chunk ahead
      const int kc0 = (SPARSE && cw.n > 0) ? cw.at(0) : 0;
      load_k4(kv, kc0 * BK + kofs, p.K, kv4, xk);
      for (int ci = 0; ci < cw.n; ++ci) {
        // Evaluate into registers first, THEN wait for the TMEM stage: with only two phasor
        // stages the evaluation must overlap the MMAs that still read the stage.
        float g1[8], g2[8];  // (G1, G2) of my row for the 8 k of the k-step
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float sn, cs;
          fast_sincos_turn(DFT ? dft_arg(p.sign2pi, xk[j], u, p.dft_period, dft_inv) : phase_arg(p.sign2pi, xk[j], u),
                           &sn, &cs);
          // even lane (Re row): (G1, G2) = (cos, -sin), k 0..3 mine, k 4..7 the partner's;
          // odd lane (Im row):  (G1, G2) = (sin,  cos), k 4..7 mine, k 0..3 the partner's
          const float ps = __shfl_xor_sync(0xFFFFFFFFu, sn, 1);
          const float pc = __shfl_xor_sync(0xFFFFFFFFu, cs, 1);
          g1[j] = half ? ps : cs;
          g1[4 + j] = half ? sn : pc;
          g2[j] = half ? pc : -sn;
          g2[4 + j] = half ? cs : -ps;
        }
        float g1h[8], g2h[8];
        uint32_t pk[4][4];  // packed bf16: G1_hi, G1_lo, G2_hi, G2_lo, 8 k -> 4 columns each
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          float g1l2[2], g2l2[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int j = 2 * jj + e;
            g1h[j] = tf32_hi(g1[j]);
            g1l2[e] = g1[j] - g1h[j];
            g2h[j] = tf32_hi(g2[j]);
            g2l2[e] = g2[j] - g2h[j];
          }
          pk[0][jj] = pack_bf16(g1h[2 * jj], g1h[2 * jj + 1]);
          pk[1][jj] = pack_bf16(g1l2[0], g1l2[1]);
          pk[2][jj] = pack_bf16(g2h[2 * jj], g2h[2 * jj + 1]);
          pk[3][jj] = pack_bf16(g2l2[0], g2l2[1]);
        }
        int knext0 = (ci + 1) * BK;     // dense: the next chunk follows
        if (SPARSE) {
          const int kc_next = cw.next_of(ci);
          knext0 = kc_next < 0 ? p.K : kc_next * BK;
        }
        load_k4(kv, knext0 + kofs, p.K, kv4, xk);   // next chunk's coordinates (latency hidden behind the wait)
#ifdef DLUX_DEBUG_TIMING
        const long long tg1_ = clock64();
        mbar_wait(emptyG_bar(stage), phase ^ 1);
        dbg_gw += clock64() - tg1_;
#else
        mbar_wait(emptyG_bar(stage), phase ^ 1);
#endif
        tc_fence_after();
        const uint32_t g0 = tmem_base + lane_addr + (uint32_t)(G_BASE_COL + stage * G_COLS);
        // (a warp-uniform branch: every column offset of the six stores is then a compile-time constant)
        if (ks == 0) store_phasors<0>(g0, g1h, g2h, pk);
        else store_phasors<1>(g0, g1h, g2h, pk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (elect_one()) mbar_arrive_cluster(fullG_bar(stage) + lead_delta);
        if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
      }
    }
#ifdef DLUX_DEBUG_TIMING
    if (blockIdx.x == 0 && lane == 0 && (warp == FIRST_GEN_WARP || warp == FIRST_GEN_WARP + 5))
      printf("GEN%d: total %lld  wait emptyG %lld\n", warp, clock64() - dbg_gstart, dbg_gw);
#endif
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA exits while its peer may still multicast into it
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}


// ------------------------------------------------------------------ tensor-pipe peak probe
// MMA-only microbenchmark of the instruction shapes gemm_tc_kernel issues: CTA pairs
// (cta_group::2), M256 x N128, A operand resident in tensor memory, B operand resident in
// shared memory -- no TMA, no generators, no epilogue.  One elected thread of every leader CTA
// issues batches of 32 MMAs back to back and keeps two batches in flight.  bench.py times it
// with CUDA events to MEASURE the dense tf32 / bf16 tcgen05 rate the roofline is quoted
// against (profiles/tf32_peak.json).  kind: 0 = kind::tf32 (K = 8), 1 = kind::f16 bf16
// (K = 16), 2 = the kernel's own mix (4 tf32 + 4 bf16 per 16-k chunk and tile).
constexpr int PROBE_THREADS = 128;
constexpr int PROBE_BATCH = 32;
constexpr int PROBE_SMEM = 2 * PLANE_BYTES + 4 * BPLANE_BYTES + 1024 + 64;

__global__ void __launch_bounds__(PROBE_THREADS, 1)
tc_peak_probe_kernel(int kind, int n_batches, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_base = base + 2 * PLANE_BYTES + 4 * BPLANE_BYTES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + 2 * PLANE_BYTES + 4 * BPLANE_BYTES + 32);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // operand bits: pseudo-random finite values (idle-zero operands would understate the power draw)
  uint32_t h = (uint32_t)(blockIdx.x * PROBE_THREADS + threadIdx.x) * 2654435761u + 12345u;
  for (int i = threadIdx.x; i < (2 * PLANE_BYTES) / 4; i += PROBE_THREADS) {
    h = h * 1664525u + 1013904223u;
    reinterpret_cast<float*>(gen)[i] = tf32_hi(((float)(h >> 8) * (1.0f / 8388608.0f)) - 1.0f);
  }
  for (int i = threadIdx.x; i < (4 * BPLANE_BYTES) / 4; i += PROBE_THREADS) {
    h = h * 1664525u + 1013904223u;
    const float a = ((float)(h >> 8) * (1.0f / 8388608.0f)) - 1.0f;
    reinterpret_cast<uint32_t*>(gen + 2 * PLANE_BYTES)[i] = pack_bf16(a, -a * 0.5f);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    mbar_init(bar_base, 1);
    mbar_init(bar_base + 8, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // phasor-like A operand: every warp fills its lane quarter of the two phasor stages
  {
    const uint32_t la = tmem_base + ((uint32_t)(warp * 32) << 16) + G_BASE_COL;
    for (int c = 0; c < G_STAGES * G_COLS; c += 8) {
      float v[8];
      const bool packed = (c % G_COLS) >= GB_BASE;   // the bf16 planes of a stage hold packed pairs
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        h = h * 1664525u + 1013904223u;
        const float a = ((float)(h >> 8) * (1.0f / 8388608.0f)) - 1.0f;
        v[j] = packed ? __uint_as_float(pack_bf16(a, 0.25f - a)) : tf32_hi(a);
      }
      tmem_st8(la + c, v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  if (cluster_ctarank() == 0 && warp == 1) {
    const uint64_t d_rh = make_desc_sw64(base), d_ih = make_desc_sw64(base + PLANE_BYTES);
    const uint64_t b0 = make_desc_sw32(base + 2 * PLANE_BYTES), b1 = make_desc_sw32(base + 2 * PLANE_BYTES + BPLANE_BYTES);
    const uint64_t b2 = make_desc_sw32(base + 2 * PLANE_BYTES + 2 * BPLANE_BYTES);
    const uint64_t b3 = make_desc_sw32(base + 2 * PLANE_BYTES + 3 * BPLANE_BYTES);
    constexpr uint64_t KS = (UMMA_K * 4) >> 4;
    uint32_t ph[2] = {0, 0};
    for (int b = 0; b < n_batches; ++b) {
      const int slot = b & 1;
      if (b >= 2) { mbar_wait(bar_base + 8 * slot, ph[slot]); ph[slot] ^= 1; }
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int m = 0; m < PROBE_BATCH; m += 8) {
          const uint32_t d = tmem_base + (uint32_t)(((b * (PROBE_BATCH / 8) + m / 8) % NUM_ACC) * ACC_COLS);
          const uint32_t g0 = tmem_base + (uint32_t)(G_BASE_COL + ((m / 8) & 1) * G_COLS);
          const uint32_t gb = g0 + GB_BASE;
          if (kind == 0) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                umma_tf32_ts2(d, g0 + ks * UMMA_K, d_rh + ks * KS, IDESC, 1u);
                umma_tf32_ts2(d, g0 + BK + ks * UMMA_K, d_ih + ks * KS, IDESC, 1u);
              }
          } else if (kind == 1) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              umma_bf16_ts2(d, gb + 1 * GB_COLS, b0, IDESC_BF16, 1u);
              umma_bf16_ts2(d, gb + 0 * GB_COLS, b1, IDESC_BF16, 1u);
              umma_bf16_ts2(d, gb + 3 * GB_COLS, b2, IDESC_BF16, 1u);
              umma_bf16_ts2(d, gb + 2 * GB_COLS, b3, IDESC_BF16, 1u);
            }
          } else {
            issue_tile_chunk(d, g0, base, false);
          }
        }
        umma_commit_mc2(bar_base + 8 * slot, (uint16_t)1);   // arrives on the leader's barrier
      }
      __syncwarp();
    }
    for (int b = n_batches > 2 ? n_batches - 2 : 0; b < n_batches; ++b) {   // drain the last two batches
      const int slot = b & 1;
      mbar_wait(bar_base + 8 * slot, ph[slot]);
      ph[slot] ^= 1;
    }
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  if (sink && blockIdx.x == 0 && warp == 0) {   // read one accumulator word so the work is observable
    uint32_t v[16];
    tmem_ld16(tmem_base, v);
    tmem_ld_wait();
    if (lane == 0) sink[0] = __uint_as_float(v[0]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
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
  bool ready = false;
};

// Kernel attributes (the dynamic shared-memory opt-in) and the SM count are PER DEVICE: one
// state per device ordinal, initialised on the first launch on that device.
TcState& tc_state() {
  constexpr int MAX_DEV = 64;
  static TcState states[MAX_DEV];
  static std::mutex mu;
  static TcState bad;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) {
    bad.rc = DLUX_ERR_CUDA;
    return bad;
  }
  std::lock_guard<std::mutex> lk(mu);
  TcState& st = states[dev];
  if (st.ready) return st;
  st.ready = true;
  cudaDeviceGetAttribute(&st.num_sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&st.cc_major, cudaDevAttrComputeCapabilityMajor, dev);
  if (st.cc_major != 10) { st.rc = DLUX_ERR_UNSUPPORTED; return st; }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || !fn) {
    st.rc = DLUX_ERR_CUDA;
    return st;
  }
  st.encode = (EncodeTiledFn)fn;
  if (cudaFuncSetAttribute(gemm_tc_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           SMEM_BYTES) != cudaSuccess ||
      cudaFuncSetAttribute(gemm_tc_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           SMEM_BYTES) != cudaSuccess ||
      cudaFuncSetAttribute(gemm_tc_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           SMEM_BYTES) != cudaSuccess ||
      cudaFuncSetAttribute(gemm_tc_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           SMEM_BYTES) != cudaSuccess) {
    st.rc = DLUX_ERR_CUDA;
    return st;
  }
  return st;
}

}  // namespace

size_t gemm_tc_workspace_bytes() { return 0; }

namespace {
// data operand planes [n][rows][K] as TMA-load tensors (box 16 k x 64 rows, SWIZZLE_64B)
int encode_in_maps(TcState& s, float* const hi[2], int K, int rows, cuuint64_t n, CUtensorMap maps[2]) {
  for (int i = 0; i < 2; ++i) {
    const cuuint64_t pitch = pitch4(K);  // row strides are multiples of 16 bytes
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, n};
    cuuint64_t strides[2] = {pitch * 4, pitch * 4 * (cuuint64_t)rows};
    cuuint32_t box[3] = {BK, ROWS_CTA, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = s.encode(&maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)hi[i], dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      fprintf(stderr, "[dlux_b200] cuTensorMapEncodeTiled failed: %d\n", (int)r);
      return DLUX_ERR_CUDA;
    }
  }
  return DLUX_OK;
}
// output planes of EPI_PLANES, written by TMA stores: [n][n_out][rows], box 16 x 16
int encode_plane_out_maps(TcState& s, float* const hi[2], int rows, int n_out, cuuint64_t n, CUtensorMap omaps[2]) {
  for (int i = 0; i < 2; ++i) {
    const cuuint64_t pitch = pitch4(rows);
    cuuint64_t dims[3] = {(cuuint64_t)rows, (cuuint64_t)n_out, n};
    cuuint64_t strides[2] = {pitch * 4, pitch * 4 * (cuuint64_t)n_out};
    cuuint32_t box[3] = {16, 16, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = s.encode(&omaps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)hi[i], dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      fprintf(stderr, "[dlux_b200] cuTensorMapEncodeTiled (output plane %d) failed: %d\n", i, (int)r);
      return DLUX_ERR_CUDA;
    }
  }
  return DLUX_OK;
}
// the final result of a stage: EPI_C64 as a float32 tensor [n_items][n_out][2 * rows] (box 32 x 16), or EPI_PSF as
// the 2-d image [n_out][rows] that receives reduce-adds
int encode_result_map(TcState& s, const GemmParams& p, CUtensorMap* omapc, int* c64_tma) {
  *c64_tma = (p.mode == EPI_C64 && (p.rows % 2) == 0 && ((uintptr_t)p.out_c64 & 15) == 0) ? 1 : 0;
  if (*c64_tma) {
    cuuint64_t dims[3] = {(cuuint64_t)2 * p.rows, (cuuint64_t)p.n_out, (cuuint64_t)p.n_items};
    cuuint64_t strides[2] = {(cuuint64_t)p.rows * 8, (cuuint64_t)p.rows * 8 * (cuuint64_t)p.n_out};
    cuuint32_t box[3] = {32, 16, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = s.encode(omapc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)p.out_c64, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      fprintf(stderr, "[dlux_b200] cuTensorMapEncodeTiled (complex64 output) failed: %d\n", (int)r);
      return DLUX_ERR_CUDA;
    }
  }
  if (p.mode == EPI_PSF) {
    if (!p.out_psf || !p.item_w || (p.rows % 4) != 0 || ((uintptr_t)p.out_psf & 15)) return DLUX_ERR_ARG;
    cuuint64_t dims[2] = {(cuuint64_t)p.rows, (cuuint64_t)p.n_out};
    cuuint64_t strides[1] = {(cuuint64_t)p.rows * 4};
    cuuint32_t box[2] = {16, 16};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = s.encode(omapc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)p.out_psf, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      fprintf(stderr, "[dlux_b200] cuTensorMapEncodeTiled (image) failed: %d\n", (int)r);
      return DLUX_ERR_CUDA;
    }
  }
  return DLUX_OK;
}
int fill_tc_params(const GemmParams& p, int c64_tma, TcParams* tp) {
  tp->g = p;
  tp->c64_tma = c64_tma;
  tp->tiles_mp = (p.rows + 2 * BM - 1) / (2 * BM);
  tp->tiles_np = (p.n_out + CLUSTER * NB - 1) / (CLUSTER * NB);
  tp->units_per_item = tp->tiles_mp * tp->tiles_np;   // (sparse: the kernel replaces it by the device-side list length)
  const long long total = (long long)tp->units_per_item * p.n_items;
  if (total > 1073741823LL) return DLUX_ERR_SHAPE;
  tp->n_units = (int)total;
  tp->k_chunks = (p.K + BK - 1) / BK;
  tp->kvec_vec4 = (p.K % 4 == 0 && p.kvec_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(p.kvec) & 15) == 0) ? 1 : 0;
  return DLUX_OK;
}
template <class Kernel>
int launch_tc(TcState& s, Kernel kernel, int n_units, cudaStream_t st, const CUtensorMap (&m)[7], const FusedTc& f) {
  const int max_clusters = s.num_sms / CLUSTER;
  const int n_clusters = n_units < max_clusters ? n_units : max_clusters;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(n_clusters * CLUSTER);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, m[0], m[1], m[2], m[3], m[4], m[5], m[6], f);
  if (e != cudaSuccess) {
    fprintf(stderr, "[dlux_b200] cudaLaunchKernelEx(gemm_tc): %s\n", cudaGetErrorString(e));
    note_cuda_error((int)e);
    return DLUX_ERR_CUDA;
  }
  note_launch();
  return check_launch("gemm_tc");
}
}  // namespace

int launch_gemm_tc(const GemmParams& p, cudaStream_t st) {
  if (p.n_items <= 0) return DLUX_OK;
  TcState& s = tc_state();
  if (s.rc != DLUX_OK) return s.rc;
  CUtensorMap m[7];
  int rc = encode_in_maps(s, p.a.hi, p.K, p.rows, (cuuint64_t)(p.n_data > 0 ? p.n_data : p.n_items), &m[0]);
  if (rc) return rc;
  m[2] = m[0]; m[3] = m[1];     // unused by the C64 / PSF epilogues
  if (p.mode == EPI_PLANES && (rc = encode_plane_out_maps(s, p.out.hi, p.rows, p.n_out, (cuuint64_t)p.n_items, &m[2])))
    return rc;
  m[4] = m[0];
  int c64_tma = 0;
  if ((rc = encode_result_map(s, p, &m[4], &c64_tma))) return rc;
  m[5] = m[0]; m[6] = m[1];     // stage-2 operand of the fused launch: unused
  FusedTc f{};
  if ((rc = fill_tc_params(p, c64_tma, &f.s[0]))) return rc;
  f.s[1] = f.s[0];
  const bool sparse = p.chunk_cnt != nullptr || p.unit_list != nullptr;
  const bool dft = p.dft_period > 0.0f;
  if (dft && sparse) return DLUX_ERR_ARG;
  const int n_units = f.s[0].n_units;
  if (dft) return launch_tc(s, gemm_tc_kernel<false, true, false>, n_units, st, m, f);
  if (sparse) return launch_tc(s, gemm_tc_kernel<true, false, false>, n_units, st, m, f);
  return launch_tc(s, gemm_tc_kernel<false, false, false>, n_units, st, m, f);
}

int gemm_tc_fused_ring(const GemmParams& g1, const GemmParams& g2, int max_ring, int* lag_out) {
  TcState& s = tc_state();
  if (s.rc != DLUX_OK) return 0;
  TcParams a, b;
  if (fill_tc_params(g1, 0, &a) || fill_tc_params(g2, 0, &b)) return 0;
  const int n_cl = s.num_sms / CLUSTER;
  // stage 2 of an item starts >= two "waves" of units after its stage 1, so that its operand is complete
  // by the time the clusters get there (waits are then the exception, not the rule)
  static const double waves = [] {
    const char* e = getenv("DLUX_B200_FUSE_WAVES");
    const double v = e ? atof(e) : 0.0;
    return v > 0.0 ? v : 2.0;
  }();
  const int per_item = a.units_per_item + b.units_per_item;
  int lag = (int)((waves * n_cl + per_item - 1) / per_item) + 1;
  if (lag > g1.n_items) lag = g1.n_items;
  int ring = lag + 2;
  if (ring > g1.n_items) ring = g1.n_items;
  if (ring > max_ring) return 0;          // (lag < ring must hold unless the ring covers every item)
  if (lag_out) *lag_out = lag;
  return ring;
}

// Stage 1 (EPI_PLANES into a ring of `ring` item slots at g1.out) and stage 2 (whose data operand g2.a is that ring)
// of the same items in one persistent launch.  sync_ws: 2 * n_items ints (zeroed here).
int launch_gemm_tc_fused(const GemmParams& g1, const GemmParams& g2, int ring, int lag, int* sync_ws, cudaStream_t st) {
  if (g1.n_items <= 0) return DLUX_OK;
  TcState& s = tc_state();
  if (s.rc != DLUX_OK) return s.rc;
  if (g1.mode != EPI_PLANES || g2.mode == EPI_PLANES || g1.n_items != g2.n_items || g2.K != g1.rows ||
      g2.rows != g1.n_out || g1.sign2pi != g2.sign2pi || g1.dft_period > 0.0f || g1.chunk_cnt || g1.unit_list ||
      g2.chunk_cnt || g2.unit_list || ring < 1 || lag < 1 || (lag >= ring && ring < g1.n_items))
    return DLUX_ERR_ARG;
  CUtensorMap m[7];
  int rc = encode_in_maps(s, g1.a.hi, g1.K, g1.rows, (cuuint64_t)(g1.n_data > 0 ? g1.n_data : g1.n_items), &m[0]);
  if (rc) return rc;
  if ((rc = encode_plane_out_maps(s, g1.out.hi, g1.rows, g1.n_out, (cuuint64_t)ring, &m[2]))) return rc;
  m[4] = m[0];
  int c64_tma = 0;
  if ((rc = encode_result_map(s, g2, &m[4], &c64_tma))) return rc;
  if ((rc = encode_in_maps(s, g1.out.hi, g2.K, g2.rows, (cuuint64_t)ring, &m[5]))) return rc;
  FusedTc f{};
  if ((rc = fill_tc_params(g1, 0, &f.s[0])) || (rc = fill_tc_params(g2, c64_tma, &f.s[1]))) return rc;
  f.fused = 1;
  f.n_items = g1.n_items;
  f.lag = lag;
  f.ring = ring;
  const long long total = (long long)f.s[0].n_units + f.s[1].n_units;
  if (total > 2147483647LL) return DLUX_ERR_SHAPE;
  f.n_units_total = (int)total;
  f.ready = sync_ws;
  f.consumed = sync_ws + g1.n_items;
  f.ready_target = f.s[0].units_per_item * CLUSTER * NUM_EPI_WARPS;
  f.consumed_target = f.s[1].units_per_item * CLUSTER * NUM_EPI_WARPS;
  if ((rc = launch_zero(reinterpret_cast<float*>(sync_ws), 2 * (size_t)g1.n_items, st))) return rc;
  return launch_tc(s, gemm_tc_kernel<false, false, true>, f.n_units_total, st, m, f);
}

// Launches the MMA-only probe; returns the real FLOPs it executes (0 on error) through *flops.
int launch_tc_peak_probe(int kind, int n_batches, float* sink, double* flops, cudaStream_t st) {
  TcState& s = tc_state();
  if (s.rc != DLUX_OK) return s.rc;
  if (kind < 0 || kind > 2 || n_batches < 1) return DLUX_ERR_ARG;
  static std::once_flag once[64];
  int dev = 0;
  cudaGetDevice(&dev);
  cudaError_t attr_rc = cudaSuccess;
  std::call_once(once[dev & 63], [&] {
    attr_rc = cudaFuncSetAttribute(tc_peak_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PROBE_SMEM);
  });
  if (attr_rc != cudaSuccess) return DLUX_ERR_CUDA;
  const int n_clusters = s.num_sms / CLUSTER;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(n_clusters * CLUSTER);
  cfg.blockDim = dim3(PROBE_THREADS);
  cfg.dynamicSmemBytes = PROBE_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, tc_peak_probe_kernel, kind, n_batches, sink);
  if (e != cudaSuccess) {
    note_cuda_error((int)e);
    return DLUX_ERR_CUDA;
  }
  // per MMA: 2 * M(256) * N(128) * K real FLOPs; K = 8 (tf32) or 16 (bf16); the mix is half and half
  const double per_mma = 2.0 * UMMA_M * BM * (kind == 0 ? 8.0 : kind == 1 ? 16.0 : 12.0);
  if (flops) *flops = per_mma * PROBE_BATCH * (double)n_batches * n_clusters;
  note_launch();
  return check_launch("tc_peak_probe");
}

}  // namespace dlux
