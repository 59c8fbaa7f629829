// fp64_rate.cu -- measures the per-SM issue rate of DFMA vs FFMA (dependent-free chains, 8 warps per SM like the
// pose-graph solver's CTAs).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rate fp64_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <typename T>
__global__ void fma_chain(T* out, int iters, long long* cycles) {
  T a[8];
  for (int i = 0; i < 8; ++i) a[i] = (T)(threadIdx.x + i) * (T)1e-3;
  const T b = (T)1.000001, c = (T)1e-7;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = a[i] * b + c;      // 8 independent chains per thread
  }
  __syncthreads();
  const long long t1 = clock64();
  T s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <typename T>
static void run(const char* name, int threads) {
  T* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(T)); cudaMalloc(&cyc, 148 * sizeof(long long));
  const int iters = 4096;
  fma_chain<T><<<148, threads>>>(out, iters, cyc);
  fma_chain<T><<<148, threads>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  const double warp_instr = (double)iters * 8 * (threads / 32);
  printf("%s threads/CTA=%4d: %.0f cycles for %.0f warp-FMAs per SM -> %.2f cycles per warp-instruction per SM (%.1f lanes/clk/SM)\n",
         name, threads, avg, warp_instr, avg / warp_instr, 32.0 * warp_instr / avg);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int th : {32, 128, 256, 1024}) { run<double>("DFMA", th); run<float>("FFMA", th); }
  return 0;
}
