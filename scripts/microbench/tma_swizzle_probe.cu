// tma_swizzle_probe.cu -- where does a SWIZZLE_128B TMA box land when its shared-memory destination is 128-byte aligned but
// NOT aligned to the 1024-byte swizzle atom?  (The halo-tile convolution wants boxes of {64 ch, 10 px, R rows} at row pitch
// 1280 B, i.e. destinations at multiples of 1280 B.)  Expectation if TMA, like tcgen05.mma, derives the XOR phase from the
// absolute address: chunk j of the 128-byte row at address A sits at A + ((j ^ ((A >> 7) & 7)) << 4).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_swizzle_probe tma_swizzle_probe.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tm, int dst_off, int rows_px, uint16_t* out /*[rows_px*64]*/, int* err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  uint8_t* gen = smem + (base - smem_u32(smem));
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(gen)[i] = 0xFFFFFFFFu;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t bytes = rows_px * 128;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(base + dst_off), "l"(&tm), "r"(smem_u32(&bar)), "r"(0), "r"(0), "r"(0) : "memory");
  }
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    if (!done && ++spins > (1u << 22)) { if (threadIdx.x == 0) *err = 1; return; }
  }
  // un-swizzle with the ABSOLUTE-address rule and write the logical [pixel][channel] array
  for (int e = threadIdx.x; e < rows_px * 8; e += blockDim.x) {
    const int p = e >> 3, j = e & 7;
    const uint32_t rowaddr = base + dst_off + p * 128;
    const uint8_t* src = gen + dst_off + p * 128 + ((j ^ ((rowaddr >> 7) & 7)) << 4);
    for (int k = 0; k < 8; ++k) out[p * 64 + j * 8 + k] = reinterpret_cast<const uint16_t*>(src)[k];
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int C = 64, W = 10, H = 18;                          // global tensor [H][W][C] fp16, value = index
  std::vector<uint16_t> h(C * W * H);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint16_t)i;
  uint16_t *d_in, *d_out; int* d_err;
  cudaMalloc(&d_in, h.size() * 2); cudaMalloc(&d_out, h.size() * 2); cudaMalloc(&d_err, 4);
  cudaMemcpy(d_in, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  PFN_encodeTiled enc = (PFN_encodeTiled)fn;
  const int smem = 48 * 1024 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int offs[] = {0, 128, 1280, 1280 * 3, 1280 * 15, 1024 * 5 + 384};
  const int rowsets[] = {18, 15, 3};
  for (int rows : rowsets)
    for (int off : offs) {
      CUtensorMap tm;
      cuuint64_t gd[3] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H};
      cuuint64_t gs[2] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2};
      cuuint32_t bx[3] = {64, 10, (cuuint32_t)rows}, es[3] = {1, 1, 1};
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, d_in, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
      cudaMemset(d_out, 0xEE, h.size() * 2); cudaMemset(d_err, 0, 4);
      probe<<<1, 128, smem>>>(tm, off, rows * W, d_out, d_err);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("rows %2d dst +%5d: %s\n", rows, off, cudaGetErrorString(e)); return 1; }
      std::vector<uint16_t> o(rows * W * C); int herr = 0;
      cudaMemcpy(o.data(), d_out, o.size() * 2, cudaMemcpyDeviceToHost); cudaMemcpy(&herr, d_err, 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (size_t i = 0; i < o.size(); ++i) bad += o[i] != (uint16_t)i;
      printf("rows %2d dst +%5d (mod 1024 = %4d): %s (%d of %zu elements differ)%s\n", rows, off, off % 1024,
             bad ? "MISMATCH" : "absolute-address swizzle OK", bad, o.size(), herr ? " [timeout]" : "");
    }
  return 0;
}
