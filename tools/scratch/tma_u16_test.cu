// standalone bisect: 2-D TMA box load of a uint16 tensor (32 x 32 box, no swizzle) into shared memory
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap map, int c0, int c1, uint16_t* out) {
  __shared__ __align__(128) uint16_t tile[32 * 32];
  __shared__ __align__(8) uint64_t bar;
  uint32_t sb = (uint32_t)__cvta_generic_to_shared(&bar), st = (uint32_t)__cvta_generic_to_shared(tile);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sb));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb), "r"(2048) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(st), "l"(&map),
                 "r"(sb), "r"(c0), "r"(c1)
                 : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}" ::"r"(sb) : "memory");
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = tile[i];
}
int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int rows = 100, cols = 2500, pitch = (variant == 1 ? 2560 : 2504);
  std::vector<uint16_t> h((size_t)rows * pitch);
  for (int r = 0; r < rows; ++r) for (int c = 0; c < pitch; ++c) h[(size_t)r * pitch + c] = (uint16_t)(r * 100 + c % 100);
  uint16_t *d, *o;
  cudaMalloc(&d, h.size() * 2); cudaMalloc(&o, 2048);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn fn = (EncodeTiledFn)p;
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)pitch * 2};
  cuuint32_t box[2] = {32, 32}, es[2] = {1, 1};
  CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  (variant == 2 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %d\n", (int)r);
  k<<<1, 128>>>(map, variant == 4 ? 32 : 37, variant == 4 ? 8 : 11, o);
  printf("variant %d\n", variant);
  cudaError_t e = cudaDeviceSynchronize();
  printf("run: %s\n", cudaGetErrorString(e));
  std::vector<uint16_t> res(1024);
  cudaMemcpy(res.data(), o, 2048, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int i = 0; i < 32; ++i) for (int j = 0; j < 32; ++j) bad += res[i * 32 + j] != h[(size_t)(11 + i) * pitch + 37 + j];
  printf("mismatches %d (first %d expect %d)\n", bad, res[0], h[(size_t)11 * pitch + 37]);
  return 0;
}
