// umma_rate_probe.cu -- cycles per tcgen05.mma (kind::f16, cta_group::1, M = 128, K = 16, both operands from shared memory)
// as a function of N, of how many independent accumulators the issue stream rotates over, of the operand layout (1024-byte
// atoms vs the 1280-byte halo pitch of the fused convolution) and of operand re-use between consecutive instructions.
//
// Why: the 64-channel convolution layers issue, per K step, one N = 128 MMA and one N = 64 MMA into ONE accumulator tile and
// measure ~115 cycles per K step where the tensor-pipe arithmetic needs 96.  Is that a dependent-accumulate latency (then
// two tiles interleaved would hide it), shared-memory operand bandwidth (then only fewer bytes per MAC help), or the issue
// rate of the one issuing thread?
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_rate_probe umma_rate_probe.cu
// run:   ./umma_rate_probe            (prints one line per configuration; one CTA, the numbers are SM cycles)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Op { uint64_t a, b; uint32_t d, idesc; };
constexpr int MAXOPS = 8;
struct Cfg { Op op[MAXOPS]; int nops; int iters; };

__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int NOPS>
__global__ void __launch_bounds__(128, 1) rate_kernel(Cfg c, long long* out) {
  extern __shared__ uint8_t raw[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (uint32_t e = tid * 16; e < 160 * 1024; e += 128 * 16) *reinterpret_cast<uint4*>(raw + (base - smem_u32(raw)) + e) = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    Op op[NOPS];
#pragma unroll
    for (int i = 0; i < NOPS; ++i) { op[i] = c.op[i]; op[i].a += (uint64_t)(base >> 4); op[i].b += (uint64_t)(base >> 4); op[i].d += tmem; }
    const long long t0 = clock64();
    for (int it = 0; it < c.iters; ++it) {
#pragma unroll
      for (int i = 0; i < NOPS; ++i) mma(op[i].d, op[i].a, op[i].b, op[i].idesc, 1u);
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    const long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// descriptor WITHOUT the base address (added on the device): K-major SWIZZLE_128B, byte offset `off`, SBO bytes
static uint64_t desc(uint32_t off, uint32_t sbo) {
  return (uint64_t)((off >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
static uint32_t idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

static long long* d_out;
template <int NOPS>
static void run(const char* name, Cfg c) {
  c.nops = NOPS; c.iters = 2048 / NOPS * 2;
  const int smem = 161 * 1024 + 1024;
  cudaFuncSetAttribute(rate_kernel<NOPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long h[2] = {0, 0};
  for (int rep = 0; rep < 2; ++rep) {
    rate_kernel<NOPS><<<1, 128, smem>>>(c, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-72s %s\n", name, cudaGetErrorString(e)); exit(1); }
  }
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  const double n = (double)c.iters * NOPS;
  printf("%-72s issue %7.1f  complete %7.1f cycles per MMA   (%.1f per group of %d)\n", name, h[0] / n, h[1] / n, h[1] / n * NOPS, NOPS);
}

int main() {
  cudaMalloc(&d_out, 16);
  // shared-memory map (offsets from the 1024-aligned base): A region 0 .. 64 KB, B region 64 KB .. 160 KB
  const uint32_t A0 = 0, B0 = 64 * 1024;
  auto A = [&](int tile, int kslice) { return desc(A0 + tile * 16384 + kslice * 32, 1024); };           // 128 rows x 128 B tiles
  auto Ah = [&](int tap, int kslice) { return desc(A0 + ((tap / 3) * 10 + tap % 3) * 128 + kslice * 32, 1280); };   // halo views
  auto B = [&](int tile, int kslice) { return desc(B0 + tile * 32768 + kslice * 32, 1024); };           // up to 256 rows x 128 B
  Cfg c;
  // --- one MMA shape, one accumulator, same operands every time
  for (int n : {64, 128, 256}) {
    char nm[128]; snprintf(nm, sizeof nm, "N=%d, 1 accumulator, same A/B K-slice", n);
    for (int i = 0; i < 4; ++i) c.op[i] = {A(0, 0), B(0, 0), 0u, idesc(n)};
    run<4>(nm, c);
  }
  // --- K-slices advancing as in a real K loop
  for (int n : {64, 128, 256}) {
    char nm[128]; snprintf(nm, sizeof nm, "N=%d, 1 accumulator, K-slices 0..3", n);
    for (int i = 0; i < 4; ++i) c.op[i] = {A(0, i), B(0, i), 0u, idesc(n)};
    run<4>(nm, c);
  }
  // --- rotating over independent accumulators
  for (int n : {64, 128}) {
    for (int nacc : {2, 4}) {
      char nm[128]; snprintf(nm, sizeof nm, "N=%d, %d accumulators round-robin, K-slices 0..3", n, nacc);
      for (int i = 0; i < 8; ++i) c.op[i] = {A(i % nacc, (i / nacc) % 4), B(0, (i / nacc) % 4), (uint32_t)((i % nacc) * 128), idesc(n)};
      run<8>(nm, c);
    }
  }
  // --- the convolution's pattern: wide N=128 (A_hi x [W_hi|W_lo]) then N=64 (A_lo x W_hi) into the cross half
  for (int i = 0; i < 4; ++i) { c.op[2 * i] = {A(0, i), B(0, i), 0u, idesc(128)}; c.op[2 * i + 1] = {A(1, i), B(0, i), 64u, idesc(64)}; }
  run<8>("conv pattern: N=128 (A_hi) + N=64 (A_lo -> cross half), one tile", c);
  for (int i = 0; i < 4; ++i) { c.op[2 * i] = {Ah(4, i), B(0, i), 0u, idesc(128)}; c.op[2 * i + 1] = {Ah(4, i) + (uint64_t)(32768 >> 4), B(0, i), 64u, idesc(64)}; }
  run<8>("conv pattern, A = halo view (pitch 1280 B, unaligned start)", c);
  // two tiles interleaved: tile 0 -> columns 0..127, tile 1 -> columns 128..255
  for (int i = 0; i < 2; ++i) {
    c.op[4 * i + 0] = {A(0, i), B(0, i), 0u, idesc(128)};   c.op[4 * i + 1] = {A(2, i), B(0, i), 128u, idesc(128)};
    c.op[4 * i + 2] = {A(1, i), B(0, i), 64u, idesc(64)};   c.op[4 * i + 3] = {A(3, i), B(0, i), 192u, idesc(64)};
  }
  run<8>("conv pattern, two tiles interleaved (wide0 wide1 lo0 lo1)", c);
  for (int i = 0; i < 2; ++i) {
    c.op[4 * i + 0] = {A(0, i), B(0, i), 0u, idesc(128)};   c.op[4 * i + 1] = {A(1, i), B(0, i), 64u, idesc(64)};
    c.op[4 * i + 2] = {A(2, i), B(0, i), 128u, idesc(128)}; c.op[4 * i + 3] = {A(3, i), B(0, i), 192u, idesc(64)};
  }
  run<8>("conv pattern, two tiles interleaved (wide0 lo0 wide1 lo1)", c);
  // lo product into its own columns (no overlap with the wide accumulator)
  for (int i = 0; i < 4; ++i) { c.op[2 * i] = {A(0, i), B(0, i), 0u, idesc(128)}; c.op[2 * i + 1] = {A(1, i), B(0, i), 128u, idesc(64)}; }
  run<8>("conv pattern, lo product into separate columns", c);
  // one N=192 MMA instead (A_hi x [W_hi|W_lo|...]) -- what a single wider instruction costs
  for (int i = 0; i < 4; ++i) c.op[i] = {A(0, i), B(0, i), 0u, idesc(192)};
  run<4>("N=192, 1 accumulator, K-slices 0..3", c);
  // different B per instruction (weights of different taps)
  for (int i = 0; i < 8; ++i) c.op[i] = {A(0, i % 4), B(i % 3, i % 4), 0u, idesc(128)};
  run<8>("N=128, 1 accumulator, B from 3 different tiles", c);
  return 0;
}
