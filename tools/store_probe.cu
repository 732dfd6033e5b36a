// Burst-store probe: how fast can the SMs push an epilogue-sized burst (256 KiB per CTA) out
// through TMA tensor stores, as a function of (a) how many SMs store at the same time and (b) the
// contiguous row length of the boxes?  gemm_tc_kernel's EPI_PLANES epilogue writes 16 x 16 boxes
// (64-byte rows of the fp32 planes, 32-byte rows of the bf16 planes) and all 148 CTAs reach their
// epilogue at the same moment; the measured ~12 k cycles per unit = 21 B/clk/SM.  This tool tells
// whether that is a per-SM limit or a shared (L2 fabric) limit that staggering the CTAs relieves.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/store_probe tools/store_probe.cu -lcuda
//   ./gpurun_out/store_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256, 1)
store_probe(const __grid_constant__ CUtensorMap map, int boxes_per_warp, int box_cols, int active_mod,
            int col_tiles, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (blockIdx.x % active_mod != 0) return;
  const int cta = blockIdx.x / active_mod;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* stg = reinterpret_cast<float*>(smem + warp * 8192);
  for (int i = lane; i < 2048; i += 32) stg[i] = (float)(i + warp);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const long long t0 = clock64();
  if (lane == 0) {
    for (int b = 0; b < boxes_per_warp; ++b) {
      const int gbox = (cta * 8 + warp) * boxes_per_warp + b;   // global box index
      const int c0 = (gbox % col_tiles) * box_cols;
      const int r0 = (gbox / col_tiles) * 16;
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                   ::"l"(&map), "r"(smem_u32(stg) + (b & 1) * 4096), "r"(c0), "r"(r0) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  int dev = 0, sms = 0, khz = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn encode = (EncodeTiledFn)fn;
  const size_t per_cta = 256 << 10;
  const int cols = 1024;                                        // fp32 row pitch 4 KiB, as the C3 intermediate
  const size_t rows = (size_t)sms * per_cta / (cols * 4);
  float* dst;
  CK(cudaMalloc(&dst, rows * cols * 4));
  long long* cyc;
  CK(cudaMalloc(&cyc, sms * sizeof(long long)));
  CK(cudaFuncSetAttribute(store_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  printf("SMs %d, nominal clock %d MHz, burst %zu KiB per CTA, rows of %d floats\n", sms, khz / 1000, per_cta >> 10, cols);
  printf("%10s %10s %12s %12s %14s %12s\n", "box_cols", "row_bytes", "active_SMs", "time_us", "agg_GB/s", "B/clk/SM");
  for (int box_cols : {8, 16, 32, 64, 128}) {
    const int box_bytes = box_cols * 4 * 16;
    if (box_bytes > 4096 * 2) continue;
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, 16};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dst, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    const int boxes_per_warp = (int)(per_cta / 8 / box_bytes);
    for (int active_mod : {1, 2, 4, 8}) {
      float best = 1e30f;
      long long best_cyc = 0;
      for (int rep = 0; rep < 6; ++rep) {
        CK(cudaMemsetAsync(cyc, 0, sms * sizeof(long long)));
        CK(cudaEventRecord(e0));
        store_probe<<<sms, 256, 200 << 10>>>(map, boxes_per_warp, box_cols, active_mod, cols / box_cols, cyc);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        std::vector<long long> h(sms);
        CK(cudaMemcpy(h.data(), cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (auto v : h) mx = v > mx ? v : mx;
        if (rep > 0 && ms < best) { best = ms; best_cyc = mx; }
      }
      const int active = (sms + active_mod - 1) / active_mod;
      const double bytes = (double)active * per_cta;
      printf("%10d %10d %12d %12.1f %14.1f %12.1f   (max CTA cycles %lld)\n", box_cols, box_cols * 4, active,
             best * 1e3, bytes / (best * 1e-3) / 1e9, (double)per_cta / (double)best_cyc, best_cyc);
    }
  }
  return 0;
}
