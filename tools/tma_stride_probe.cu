// Does a TMA tiled load with elementStrides = {2, 1} gather every second 16-bit element, starting at an
// ODD coordinate, into a dense SWIZZLE_32B box?  (Would let the bf16 "hi" operand plane be a strided view
// of the float32 plane: the upper halves of the words, at zero instruction cost.)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap map, int c0, int r0, uint16_t* out, int nbytes) {
  __shared__ __align__(1024) uint8_t buf[4096];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) buf[i] = 0xEE;
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(nbytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(buf)), "l"(&map), "r"(smem_u32(&bar)), "r"(c0), "r"(r0) : "memory");
  }
  uint32_t done = 0;
  long long t0 = clock64();
  while (!done && clock64() - t0 < 20000000LL) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(buf)[i];
  if (threadIdx.x == 0) out[2048] = (uint16_t)done;
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn encode = (EncodeTiledFn)fn;
  const int rows = 64, K = 64;                    // fp32 plane [rows][K]; word (r, k) = (hi16 = 0x8000 | (r << 8) | k, lo16 = k)
  std::vector<uint32_t> h(rows * K);
  for (int r = 0; r < rows; ++r) for (int k = 0; k < K; ++k) h[r * K + k] = ((uint32_t)(0x8000 | (r << 8) | k) << 16) | (uint32_t)k;
  uint32_t* d; CK(cudaMalloc(&d, h.size() * 4)); CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  uint16_t* out; CK(cudaMalloc(&out, 2049 * 2));
  for (int variant = 0; variant < 4; ++variant) {
    // bf16 view: dim0 = 2K halves per row
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)2 * K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 4};
    cuuint32_t box[2] = {32, 64};                 // traverse 32 halves with stride 2 -> 16 elements; 64 rows
    cuuint32_t estr[2] = {2, 1};
    CUtensorMapSwizzle sw = (variant & 1) ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    int c0 = (variant & 2) ? 2 * 16 + 1 : 1;      // k0 = 16 or 0, odd start = the upper halves
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d (swizzle %s, c0 %d): encode rc %d\n", variant, (variant & 1) ? "32B" : "none", c0, (int)r);
    if (r != CUDA_SUCCESS) continue;
    probe<<<1, 128>>>(map, c0, 0, out, 16 * 64 * 2);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("  kernel error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<uint16_t> o(2049);
    CK(cudaMemcpy(o.data(), out, 2049 * 2, cudaMemcpyDeviceToHost));
    printf("  completed=%d; row0: ", o[2048]);
    for (int i = 0; i < 20; ++i) printf("%04x ", o[i]);
    printf("\n  row1: ");
    for (int i = 16; i < 36; ++i) printf("%04x ", o[i]);
    printf("\n  row5: ");
    for (int i = 80; i < 96; ++i) printf("%04x ", o[i]);
    printf("\n");
  }
  return 0;
}
