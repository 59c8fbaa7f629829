// umma_swizzle_probe.cu -- what does tcgen05.mma read for a K-major SWIZZLE_128B operand whose start address is NOT
// aligned to the 1024-byte swizzle atom, and whose stride between 8-row groups (SBO) is not a multiple of 1024 bytes?
//
// Why: the fused conv1a -> conv1b kernel wants ONE copy of the conv1a halo tile in shared memory (18 rows x 10 pixels x
// 64 channels, pixel pitch 128 B, row pitch 1280 B) and nine tap views of it: tap (ky,kx) = the same bytes read from
// start + (ky*10 + kx)*128 with SBO = 1280.  Whether that works depends on how the hardware forms the XOR phase:
//   "absolute": phase = (byte address >> 7) & 7             -> data written with the absolute-address swizzle, base_offset 0
//   "relative": phase = (row within 8-row group + base_offset) & 7 -> data written with phase = pixel column & 7, base_offset kx
// The probe runs one 128x64x64 MMA per (variant, tap) and compares with the exact integer result.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_swizzle_probe umma_swizzle_probe.cu
// run:   for v in 0 1 2 3 4 5; do ./umma_swizzle_probe $v; done      (one process per variant: a trap must not hide the rest)
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int ROWS = 18, COLS = 10, NPIX = ROWS * COLS;   // halo tile of a 16 x 8 output tile
constexpr int N = 64, K = 64;

__host__ __device__ inline int a_val(int p, int k) { return ((p * 5 + k * 3) % 11) - 5; }
__host__ __device__ inline int b_val(int n, int k) { return ((n * 7 + k) % 9) - 4; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) probe_kernel(int variant, int ky, int kx, float* out /*[128][64]*/) {
  extern __shared__ uint8_t raw[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  uint8_t* A = gen;                      // NPIX * 128 B = 23040 B
  uint8_t* B = gen + 24 * 1024;          // 64 rows * 128 B, 1024-aligned
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool layout_abs = (variant == 0 || variant == 1 || variant == 5);
  for (int e = tid; e < NPIX * 8; e += 128) {
    const int p = e >> 3, j = e & 7, col = p % COLS;
    const uint32_t addr = base + p * 128;
    const int phase = layout_abs ? ((addr >> 7) & 7) : (col & 7);
    __half h[8];
    for (int i = 0; i < 8; ++i) h[i] = __float2half((float)a_val(p, j * 8 + i));
    *reinterpret_cast<uint4*>(A + p * 128 + ((j ^ phase) * 16)) = *reinterpret_cast<uint4*>(h);
  }
  for (int e = tid; e < N * 8; e += 128) {
    const int n = e >> 3, j = e & 7;
    __half h[8];
    for (int i = 0; i < 8; ++i) h[i] = __float2half((float)b_val(n, j * 8 + i));
    *reinterpret_cast<uint4*>(B + n * 128 + ((j ^ (n & 7)) * 16)) = *reinterpret_cast<uint4*>(h);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> visible to the MMA's async proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t start = base + (uint32_t)((ky * COLS + kx) * 128);
    int bo = 0;
    if (variant == 1) bo = (start >> 7) & 7;
    if (variant == 2) bo = kx;
    if (variant == 3) bo = (8 - kx) & 7;
    if (variant == 5) bo = (8 - ((start >> 7) & 7)) & 7;
    const uint64_t sbo = (uint64_t)((COLS * 128) >> 4);            // 1280 B between 8-pixel groups (tile rows)
    const uint64_t adesc = (uint64_t)((start >> 4) & 0x3FFF) | (1ull << 16) | (sbo << 32) | (1ull << 46) |
                           ((uint64_t)bo << 49) | (2ull << 61);
    const uint32_t bstart = base + 24 * 1024;
    const uint64_t bdesc = (uint64_t)((bstart >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int k = 0; k < K / 16; ++k) {
      const uint64_t adv = (uint64_t)((k * 32) >> 4);
      const uint32_t acc = k ? 1u : 0u;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(adesc + adv), "l"(bdesc + adv), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {
    uint32_t done = 0, spins = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
      if (!done && ++spins > (1u << 22)) __trap();
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int n0 = 0; n0 < N; n0 += 16) {
    uint32_t v[16];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + n0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) out[tid * N + n0 + i] = __uint_as_float(v[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const char* names[6] = {"abs-layout bo=0", "abs-layout bo=(start>>7)&7", "col-layout bo=kx", "col-layout bo=(8-kx)&7",
                          "col-layout bo=0", "abs-layout bo=-(start>>7)&7"};
  float* d_out;
  cudaMalloc(&d_out, 128 * N * sizeof(float));
  const int smem = 24 * 1024 + 8 * 1024 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  static float h[128 * N];
  printf("variant %d (%s):", variant, names[variant]);
  int total_bad = 0;
  for (int ky = 0; ky < 3; ++ky)
    for (int kx = 0; kx < 3; ++kx) {
      cudaMemset(d_out, 0xff, sizeof(h));
      probe_kernel<<<1, 128, smem>>>(variant, ky, kx, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf(" [tap %d,%d: %s]\n", ky, kx, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int m = 0; m < 128; ++m) {
        const int r = m >> 3, c = m & 7, p = (r + ky) * COLS + c + kx;
        for (int n = 0; n < N; ++n) {
          int ref = 0;
          for (int k = 0; k < K; ++k) ref += a_val(p, k) * b_val(n, k);
          if (h[m * N + n] != (float)ref) ++bad;
        }
      }
      printf(" (%d,%d)=%s", ky, kx, bad ? "BAD" : "ok");
      if (bad) printf("[%d]", bad);
      total_bad += bad;
    }
  printf("  => %s\n", total_bad ? "MISMATCH" : "ALL TAPS EXACT");
  return 0;
}
