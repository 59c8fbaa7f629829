// conv_umma.cu -- SuperPoint convolutions on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a only.
//
// Layers: swarm_loop/superpoint.ipynb:143-158 of the reference (3x3 pad 1 and 1x1 convolutions, NHWC here).
//
// Implicit GEMM, one CTA tile = 8 x 16 output pixels (M = 128) x all output channels (N = Cout, 64..256):
//   * A operand: for every filter tap (ky,kx) and every 64-channel slab, ONE TMA box {64 ch, 16 x, 8 y, 1 image}
//     fetched at the tap-shifted coordinate; out-of-image elements are zero-filled by the TMA unit, which is the
//     convolution's zero padding.  The box lands in shared memory as 128 rows x 128 B with the 128-byte swizzle,
//     i.e. exactly the canonical K-major SWIZZLE_128B UMMA layout (im2col staging is done by the copy engine).
//   * B operand: the weight slab [N][64] of the same tap, K-major SWIZZLE_128B, by TMA.
//   * D: fp32 accumulators in TMEM (128 lanes x N columns), double buffered so the epilogue of tile t overlaps
//     the MMAs of tile t+1.
// Precision: parity with the fp32 oracle needs ~1e-6 relative error, which fp16/bf16 operands cannot give.  Every
// fp32 operand x is carried as TWO fp16 planes  hi = fp16(s*x), lo = fp16(s*x - hi)  (s a power of two, exact), and
// each K step computes the three products  hi*hi + lo*hi + hi*lo  (the dropped lo*lo term is 2^-22 relative): hi*hi goes
// to a MAIN accumulator, the two cross products to a CROSS accumulator (see the truncation note at the MMA issuer), and
// for 2N <= 256 one MMA of width 2N covers hi*hi and hi*lo at once.  Both planes together are 4 bytes per element -- the
// same HBM/L2 footprint as fp32 activations -- and the fp16 MACs cost 1.5x one TF32 MMA.  Measured against the oracle:
// see tests/test_gpu_superpoint.py.
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane each), warp 2 = TMEM
// allocator, warps 4-11 = epilogue (tcgen05.ld -> bias/ReLU -> re-split -> NHWC stores; two warps per TMEM lane quarter).
// Persistent over tiles.
#include <cuda.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "kernels.cuh"
#include "conv_umma.cuh"
#include "umma_ptx.cuh"

namespace osb {

constexpr int UM_TH = 8, UM_TW = 16;            // output tile (pixels)
constexpr int UM_KC = 64;                       // fp16 channels per K slab (= 128 bytes = one swizzle row)

// Shared-memory plan.  For a 3x3 layer ONE A box per (kx, 64-channel slab) carries 10 rows (tile + vertical halo):
// the three vertical taps ky = 0,1,2 read it at row offsets ky*16 rows = ky*2048 bytes -- a multiple of the 1024-byte
// swizzle atom, so the same SWIZZLE_128B descriptor applies with only the start address moved.  That cuts the
// activation traffic from 9 to 3.75 tile-loads per tile (L2 -> shared memory is what bounds this kernel).
// The weights of each tap stream through their own ring.
constexpr int UM_A_SLOT = (UM_TH + 2) * UM_TW * 128;   // 20 KB per plane: 10 rows x 16 px x 128 B
constexpr int UM_A_SLOTS = 2;
constexpr int UM_THREADS = 384;                 // 12 warps: TMA, MMA, TMEM allocator, (idle), 8 x epilogue
// RES = true (64 -> 64 channel 3x3 layers: conv1b, conv2a, conv2b = 65 % of the network's FLOPs): the 9 taps' weight
// planes (144 KB) stay resident in shared memory for the CTA's whole life instead of streaming through a ring, which
// removes more than half of the remaining L2 -> shared-memory traffic.
template <int N, bool RES>
struct UmmaCfg {
  static constexpr int B_BYTES = N * 128;                        // one weight plane of one tap / slab
  static constexpr int B_SLOT = 2 * B_BYTES;                     // hi + lo
  static constexpr int B_SLOTS = RES ? 9 : (N <= 64) ? 6 : (N <= 80) ? 5 : (N <= 128) ? 4 : 2;
  static constexpr int A_RING = UM_A_SLOTS * 2 * UM_A_SLOT;      // 80 KB
  // two fp32 accumulators per tile (see the precision note in the kernel): 2N TMEM columns per buffer
  static constexpr int NBUF = (4 * N <= 512) ? 2 : 1;
  static constexpr int TMEM_COLS = (2 * N * NBUF <= 128) ? 128 : (2 * N * NBUF <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = A_RING + B_SLOTS * B_SLOT + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

struct UmmaArgs {
  const float* bias;       // [N]
  __half* out_hi;          // NHWC planes of the next layer (or null)
  __half* out_lo;
  float* out_f32;          // fp32 output [pixels][out_cstride] (or null)
  int H, W, B;
  int ks;                  // 1 or 3
  int cin_slabs;           // Cin / 64
  int out_c;               // channels stored per pixel (<= N)
  int out_cstride;         // channel stride of the destination
  float inv_scale;         // 1 / (act_scale * w_scale)
  float out_scale;         // scale of the stored fp16 planes
  int relu;                // 0 none, 1 ReLU, 2 ReLU6
  int n_off;               // first output channel of this launch (Cout > 256 runs as several N <= 256 passes)
  int pool;                // 1: fused 2x2 max-pool, the planes written are [B][H/2][W/2][C]
  int n_split;             // a layer wider than the kernel's N runs as n_split work items per tile (N channels each)
  int epi;                 // 0: bias/act/pool + store; 1: detector head -- softmax over 65 logits, drop the dustbin,
                           //    8x8 pixel shuffle straight into the heat map `out_f32` ([B][8H][8W])
};

template <int N, bool RES, bool SPLIT>
__global__ void __launch_bounds__(UM_THREADS, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                 const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo, UmmaArgs P) {
  using Cfg = UmmaCfg<N, RES>;
  constexpr int AS = UM_A_SLOTS, BS = Cfg::B_SLOTS, NBUF = Cfg::NBUF;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_base = smem_base + Cfg::A_RING;
  const uint32_t bar_base = b_base + BS * Cfg::B_SLOT;                   // 8-byte barriers
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (AS + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * AS + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * AS + BS + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * AS + 2 * BS + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * AS + 2 * BS + 2 + a); };
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_x = (P.W + UM_TW - 1) / UM_TW, tiles_y = (P.H + UM_TH - 1) / UM_TH;
  const int n_tiles = P.B * tiles_x * tiles_y;
  // work item = (tile, channel block): items of one tile are adjacent, so concurrent CTAs share its activations in L2
  const int n_split = SPLIT ? P.n_split : 1;          // compile-time 1 for ordinary layers: no div / mod per item
  const int n_items = n_tiles * n_split;
  const int halo = P.ks / 2;
  const uint32_t a_box_bytes = (uint32_t)(UM_TH + 2 * halo) * UM_TW * 128;   // bytes of one A plane box

  if (threadIdx.x == 0) {
    for (int s = 0; s < AS; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < BS; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8); }
    // (with NBUF == 1 only index 0 is used; with RES the b_full barriers are filled once and b_empty stays unused)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  // programmatic dependent launch: the next layer's CTAs may be scheduled as ours retire (they still wait for this whole
  // grid in their own griddepcontrol.wait before touching activations)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0 && elect_one()) {
    // ===================== TMA producer =====================
    int as = 0; uint32_t aph = 0;
    int bs = 0; uint32_t bph = 0;
    if (RES) {
      // all 9 taps (one 64-channel slab) once: slot t holds tap t.  Issued BEFORE the dependency wait: weights are
      // constants, so under programmatic dependent launch they stream in while the previous layer is still draining.
      for (int t = 0; t < 9; ++t) {
        const uint32_t sb = b_base + t * Cfg::B_SLOT;
        mbar_expect_tx(b_full(t), Cfg::B_SLOT);
        tma_load_3d(sb, &tm_w_hi, b_full(t), 0, P.n_off, t);
        tma_load_3d(sb + Cfg::B_BYTES, &tm_w_lo, b_full(t), 0, P.n_off, t);
      }
    }
    // the activations are the previous kernel's output: wait for the whole grid we depend on (no-op without PDL)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int tile = SPLIT ? item / n_split : item, n_off = SPLIT ? P.n_off + (item % n_split) * N : P.n_off;
      const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
      const int x0 = tx * UM_TW, y0 = ty * UM_TH;
      for (int kx = 0; kx < P.ks; ++kx) {
        for (int cs = 0; cs < P.cin_slabs; ++cs) {
          // activation box: tile rows + vertical halo at the kx-shifted column, shared by the ks vertical taps
          mbar_wait(a_empty(as), aph ^ 1);
          const uint32_t sa = smem_base + as * (2 * UM_A_SLOT);
          mbar_expect_tx(a_full(as), 2 * a_box_bytes);
          tma_load_4d(sa, &tm_a_hi, a_full(as), cs * UM_KC, x0 + kx - halo, y0 - halo, b);
          tma_load_4d(sa + UM_A_SLOT, &tm_a_lo, a_full(as), cs * UM_KC, x0 + kx - halo, y0 - halo, b);
          if (++as == AS) { as = 0; aph ^= 1; }
          for (int ky = 0; ky < P.ks && !RES; ++ky) {
            mbar_wait(b_empty(bs), bph ^ 1);
            const uint32_t sb = b_base + bs * Cfg::B_SLOT;
            mbar_expect_tx(b_full(bs), Cfg::B_SLOT);
            tma_load_3d(sb, &tm_w_hi, b_full(bs), cs * UM_KC, n_off, ky * P.ks + kx);
            tma_load_3d(sb + Cfg::B_BYTES, &tm_w_lo, b_full(bs), cs * UM_KC, n_off, ky * P.ks + kx);
            if (++bs == BS) { bs = 0; bph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 && elect_one()) {
    // ===================== MMA issuer =====================
    // instruction descriptor (cute::UMMA::InstrDescriptor): D = f32 (1 << 4), A = B = f16 (0), K-major both,
    // N >> 3 at bit 17, M >> 4 at bit 24
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr bool WIDE = (2 * N <= 256);
    constexpr uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * N) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    // Precision: the tensor core TRUNCATES its fp32 accumulator after every MMA (measured: ~1e-5 relative per layer when
    // all three products of the split share one accumulator).  The exact-in-fp32 hi*hi products therefore go to a MAIN
    // accumulator (one truncation per K=16 step) and the two small cross products lo*hi, hi*lo to a second, CROSS
    // accumulator whose magnitude -- and truncation error -- is 2^-11 of the main one; the epilogue adds them in fp32.
    int as = 0; uint32_t aph = 0;
    int bs = 0; uint32_t bph = 0;
    int acc = 0; uint32_t acc_phase = 0;
    if (RES)
      for (int t = 0; t < 9; ++t) mbar_wait(b_full(t), 0);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      mbar_wait(tempty_bar(acc), acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_main = tmem_base + (uint32_t)(acc * 2 * N);
      const uint32_t d_cross = d_main + (uint32_t)N;
      uint32_t first = 1;
      for (int kx = 0; kx < P.ks; ++kx) {
        for (int cs = 0; cs < P.cin_slabs; ++cs) {
          mbar_wait(a_full(as), aph);
          tc_fence_after();
          const uint32_t sa = smem_base + as * (2 * UM_A_SLOT);
          for (int ky = 0; ky < P.ks; ++ky) {
            uint32_t sb;
            if (RES) {
              sb = b_base + (ky * 3 + kx) * Cfg::B_SLOT;
            } else {
              mbar_wait(b_full(bs), bph);
              tc_fence_after();
              sb = b_base + bs * Cfg::B_SLOT;
            }
            // vertical tap ky reads the box from tile row ky on: + ky * 16 px * 128 B = ky * 2048 B (2 swizzle atoms)
            const uint64_t a_hi = umma_desc_sw128(sa + ky * (UM_TW * 128));
            const uint64_t a_lo = umma_desc_sw128(sa + UM_A_SLOT + ky * (UM_TW * 128));
            const uint64_t b_hi = umma_desc_sw128(sb), b_lo = umma_desc_sw128(sb + Cfg::B_BYTES);
#pragma unroll
            for (int k = 0; k < UM_KC / 16; ++k) {
              const uint64_t adv = (uint64_t)(k * 32 >> 4);     // advance 16 fp16 = 32 bytes inside the swizzle row
              const uint32_t accum = (first && k == 0) ? 0u : 1u;
              if (WIDE) {
                // the weight slot is [W_hi (N rows) | W_lo (N rows)] and the accumulators are [main (N cols) | cross (N cols)]:
                // ONE MMA of width 2N computes hi*hi -> main and hi*lo -> cross, fetching the activation operand once
                // (at N = 64 the instruction is bound by shared-memory operand bandwidth, not by the tensor pipe)
                umma_f16(d_main, a_hi + adv, b_hi + adv, idesc2, accum);
                umma_f16(d_cross, a_lo + adv, b_hi + adv, idesc, 1u);
              } else {
                umma_f16(d_main, a_hi + adv, b_hi + adv, idesc, accum);
                umma_f16(d_cross, a_lo + adv, b_hi + adv, idesc, accum);
                umma_f16(d_cross, a_hi + adv, b_lo + adv, idesc, 1u);
              }
            }
            first = 0;
            if (!RES) {
              umma_commit(b_empty(bs));                          // weight slot free when these MMAs retire
              if (++bs == BS) { bs = 0; bph ^= 1; }
            }
          }
          umma_commit(a_empty(as));                              // activation box free after its last vertical tap
          if (++as == AS) { as = 0; aph ^= 1; }
        }
      }
      umma_commit(tfull_bar(acc));                               // accumulators complete -> epilogue
      if (++acc == NBUF) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    // eight warps: two per TMEM lane quarter, which take the 16-column chunks alternately -- one warp per quarter leaves
    // the epilogue, not the MMAs, as the pace of the 64-channel layers (a chunk is a long dependent chain: TMEM load,
    // arithmetic, conversions, stores)
    const int q = warp & 3;                                    // TMEM lane quarter this warp may access
    const int eset = (warp - 4) >> 2;                          // 0 / 1: which chunks of the quarter
    const int m = q * 32 + lane;                               // output pixel within the tile
    const int r = m / UM_TW, c = m % UM_TW;
    int acc = 0; uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int tile = SPLIT ? item / n_split : item, n_off = SPLIT ? P.n_off + (item % n_split) * N : P.n_off;
      const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
      const int y = ty * UM_TH + r, x = tx * UM_TW + c;
      const bool inside = (y < P.H) && (x < P.W);
      const size_t pix = ((size_t)b * P.H + y) * P.W + x;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 * N);
      // fused 2x2 max-pool: the warp owns tile rows 2q, 2q+1 (lane = (row & 1) * 16 + col); the pooled pixel of
      // (even row, even col) is the max over lanes l, l+1, l+16, l+17 -> two shuffle steps, writer lanes l < 16, l even
      const int py = ty * (UM_TH / 2) + q, px = tx * (UM_TW / 2) + (lane >> 1);
      const bool pool_writer = P.pool && lane < 16 && !(lane & 1) && py < (P.H >> 1) && px < (P.W >> 1);
      const size_t ppix = ((size_t)b * (P.H >> 1) + py) * (P.W >> 1) + px;
      if (P.epi == 1) {
        if (eset == 0) {
        // fused detector head (superpoint.ipynb:190-198): the thread holds one cell's 65 logits in TMEM.  Three passes over
        // the columns (max, sum in channel order, normalise + pixel shuffle) -- the arithmetic of sp_softmax_shuffle_kernel,
        // so the heat map is bit-identical to the two-kernel path.
        auto logit = [&](uint32_t a, uint32_t c, int ch) {
          return fmaf(__uint_as_float(a) + __uint_as_float(c), P.inv_scale, __ldg(P.bias + ch));
        };
        float mx = -INFINITY;
#pragma unroll 1
        for (int n0 = 0; n0 < 80; n0 += 16) {
          uint32_t v[16], vc[16];
          tmem_ld16(t_row + n0, v);
          tmem_ld16(t_row + N + n0, vc);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (n0 + i < 65) mx = fmaxf(mx, logit(v[i], vc[i], n0 + i));
        }
        float sum = 0.f;
#pragma unroll 1
        for (int n0 = 0; n0 < 80; n0 += 16) {
          uint32_t v[16], vc[16];
          tmem_ld16(t_row + n0, v);
          tmem_ld16(t_row + N + n0, vc);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (n0 + i < 65) sum += expf(logit(v[i], vc[i], n0 + i) - mx);
        }
        const int W8 = P.W * 8;
        float* out = P.out_f32 + ((size_t)b * P.H * 8 + (size_t)y * 8) * W8 + (size_t)x * 8;
#pragma unroll 1
        for (int n0 = 0; n0 < 64; n0 += 16) {
          uint32_t v[16], vc[16];
          tmem_ld16(t_row + n0, v);
          tmem_ld16(t_row + N + n0, vc);
          if (!inside) continue;
          float e[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) e[i] = expf(logit(v[i], vc[i], n0 + i) - mx) / sum;
#pragma unroll
          for (int r2 = 0; r2 < 2; ++r2) {
            float4* dst = reinterpret_cast<float4*>(out + (size_t)(n0 / 8 + r2) * W8);
            dst[0] = make_float4(e[8 * r2], e[8 * r2 + 1], e[8 * r2 + 2], e[8 * r2 + 3]);
            dst[1] = make_float4(e[8 * r2 + 4], e[8 * r2 + 5], e[8 * r2 + 6], e[8 * r2 + 7]);
          }
        }
        }      // (the second warp of the quarter has nothing to do for this head: it only releases the accumulator)
      } else
#pragma unroll 1
      for (int n0 = eset * 16; n0 < N; n0 += 32) {
        uint32_t v[16], vc[16];
        tmem_ld16(t_row + n0, v);
        tmem_ld16(t_row + N + n0, vc);
        if (n_off - P.n_off + n0 >= P.out_c) continue;              // warp-uniform
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float a = fmaf(__uint_as_float(v[i]) + __uint_as_float(vc[i]), P.inv_scale, __ldg(P.bias + n_off + n0 + i));
          if (P.relu) a = fmaxf(a, 0.f);
          if (P.relu == 2) a = fminf(a, 6.f);
          f[i] = a;
        }
        bool store = inside;
        size_t opix = pix;
        if (P.pool) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            f[i] = fmaxf(f[i], __shfl_down_sync(0xffffffffu, f[i], 1));
            f[i] = fmaxf(f[i], __shfl_down_sync(0xffffffffu, f[i], 16));
          }
          store = pool_writer; opix = ppix;
        }
        if (!store) continue;
        if (P.out_f32) {
          float* dst = P.out_f32 + opix * P.out_cstride + n_off + n0;
#pragma unroll
          for (int i = 0; i < 2; ++i)
            st_global_256(dst + 8 * i, __float_as_uint(f[8 * i]), __float_as_uint(f[8 * i + 1]), __float_as_uint(f[8 * i + 2]),
                          __float_as_uint(f[8 * i + 3]), __float_as_uint(f[8 * i + 4]), __float_as_uint(f[8 * i + 5]),
                          __float_as_uint(f[8 * i + 6]), __float_as_uint(f[8 * i + 7]));
        } else {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float s0 = f[2 * i] * P.out_scale, s1 = f[2 * i + 1] * P.out_scale;
            // packed conversions (cvt.rn.f16x2.f32: the roundings of two scalar conversions, off the slow F2F pipe)
            const __half2 hp = __floats2half2_rn(s0, s1);
            const float2 hf = __half22float2(hp);
            const __half2 lp = __floats2half2_rn(s0 - hf.x, s1 - hf.y);
            hi[i] = *reinterpret_cast<const uint32_t*>(&hp);
            lo[i] = *reinterpret_cast<const uint32_t*>(&lp);
          }
          st_global_256(P.out_hi + opix * P.out_cstride + n_off + n0, hi[0], hi[1], hi[2], hi[3], hi[4], hi[5], hi[6], hi[7]);
          st_global_256(P.out_lo + opix * P.out_cstride + n_off + n0, lo[0], lo[1], lo[2], lo[3], lo[4], lo[5], lo[6], lo[7]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));              // 8 epilogue warps -> count 8
      if (++acc == NBUF) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
  }
}

// --------------------------------------------------------------------------------------------------------------
// first layer (Cin = 1) and 2x2 max-pool on split planes, re-split after the fp32 op
// --------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_store8(__half* hi, __half* lo, const float* f, float scale) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float s0 = f[2 * i] * scale, s1 = f[2 * i + 1] * scale;
    const __half h0 = __float2half_rn(s0), h1 = __float2half_rn(s1);
    const __half l0 = __float2half_rn(s0 - __half2float(h0)), l1 = __float2half_rn(s1 - __half2float(h1));
    h[i] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
    l[i] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
  }
  *reinterpret_cast<uint4*>(hi) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo) = make_uint4(l[0], l[1], l[2], l[3]);
}

// conv1a (Cin = 1) + bias + ReLU -> split fp16 planes.  Thread = (8 output channels, one pixel column of a 32 x 8 tile):
// the 72 weights of its channels live in registers for the whole tile, the inputs (after the u8 -> f32 LUT) are staged
// once per CTA in shared memory and slide down the column through registers (3 broadcast LDS per pixel instead of one
// LDS per FMA pair), and the 8 threads of a pixel write its 128-byte channel vector as eight adjacent 16-byte chunks --
// a warp stores 4 pixels x 128 B contiguously per plane, no staging of the output.  The plane scale (a power of two) is
// folded into weights and bias, and the 9 taps (ky-major) accumulate onto the bias.
constexpr int CF_TW = 32, CF_TH = 8;
__global__ void __launch_bounds__(256)
conv_first_split_kernel(const float* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ lut,
                        const uint8_t* __restrict__ img, __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                        int H, int W, float out_scale) {
  __shared__ float sin_[CF_TH + 2][CF_TW + 2];
  const int tid = threadIdx.x, b = blockIdx.z, x0 = blockIdx.x * CF_TW, y0 = blockIdx.y * CF_TH;
  const uint8_t* ib = img + (size_t)b * H * W;
  for (int e = tid; e < (CF_TH + 2) * (CF_TW + 2); e += 256) {
    const int r = e / (CF_TW + 2), c = e % (CF_TW + 2);
    const int gy = y0 + r - 1, gx = x0 + c - 1;
    sin_[r][c] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(lut + ib[(size_t)gy * W + gx]) : 0.f;
  }
  const int cg = tid & 7, px = tid >> 3;
  float wr[9][8], br[8];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + t * 64 + cg * 8));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + t * 64 + cg * 8 + 4));
    wr[t][0] = w0.x; wr[t][1] = w0.y; wr[t][2] = w0.z; wr[t][3] = w0.w;
    wr[t][4] = w1.x; wr[t][5] = w1.y; wr[t][6] = w1.z; wr[t][7] = w1.w;
#pragma unroll
    for (int j = 0; j < 8; ++j) wr[t][j] *= out_scale;          // power of two: exact, commutes with every rounding below
  }
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + cg * 8));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + cg * 8 + 4));
    br[0] = b0.x; br[1] = b0.y; br[2] = b0.z; br[3] = b0.w; br[4] = b1.x; br[5] = b1.y; br[6] = b1.z; br[7] = b1.w;
#pragma unroll
    for (int j = 0; j < 8; ++j) br[j] *= out_scale;
  }
  __syncthreads();
  const int x = x0 + px;
  float in[3][3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { in[0][k] = sin_[0][px + k]; in[1][k] = sin_[1][px + k]; }
#pragma unroll
  for (int r = 0; r < CF_TH; ++r) {
#pragma unroll
    for (int k = 0; k < 3; ++k) in[2][k] = sin_[r + 2][px + k];
    const int y = y0 + r;
    if (y < H && x < W) {
      uint32_t h[4], l[4];
#pragma unroll
      for (int j2 = 0; j2 < 4; ++j2) {
        float a0 = br[2 * j2], a1 = br[2 * j2 + 1];                // bias first: 9 taps accumulate onto it
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          a0 = fmaf(in[t / 3][t % 3], wr[t][2 * j2], a0);
          a1 = fmaf(in[t / 3][t % 3], wr[t][2 * j2 + 1], a1);
        }
        const float s0 = fmaxf(a0, 0.f), s1 = fmaxf(a1, 0.f);
        const __half h0 = __float2half_rn(s0), h1 = __float2half_rn(s1);
        const __half l0 = __float2half_rn(s0 - __half2float(h0)), l1 = __float2half_rn(s1 - __half2float(h1));
        h[j2] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        l[j2] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
      }
      const size_t chunk = (((size_t)b * H + y) * W + x) * 8 + cg;          // 16-byte chunk index inside the plane
      reinterpret_cast<uint4*>(out_hi)[chunk] = make_uint4(h[0], h[1], h[2], h[3]);
      reinterpret_cast<uint4*>(out_lo)[chunk] = make_uint4(l[0], l[1], l[2], l[3]);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { in[0][k] = in[1][k]; in[1][k] = in[2][k]; }
  }
}

// --------------------------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

osb_status umma_make_tmap(CUtensorMap* tm, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                          const uint32_t* box) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("conv_umma", "cuTensorMapEncodeTiled entry point not available"); return OSB_ERR_CUDA; }
  cuuint64_t gd[5]; cuuint64_t gs[4]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[128];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    set_error("conv_umma", buf);
    return OSB_ERR_CUDA;
  }
  return OSB_OK;
}

osb_status umma_layer_upload(UmmaLayer* L, const float* w_oihw, const float* bias, int cin, int cout, int ks,
                             float w_scale) {
  L->cin = cin; L->cout = cout; L->ks = ks; L->taps = ks * ks; L->w_scale = w_scale;
  L->n_pad = (cout <= 64) ? 64 : (cout <= 80) ? 80 : (cout <= 128) ? 128 : (cout <= 256) ? 256 : 512;
  OSB_REQUIRE(cin % UM_KC == 0 && cout <= 512, "tcgen05 conv: Cin must be a multiple of 64 and Cout <= 512");
  const size_t n = (size_t)L->taps * L->n_pad * cin;
  std::vector<__half> hi(n, __float2half(0.f)), lo(n, __float2half(0.f));
  std::vector<float> bp(L->n_pad, 0.f);
  for (int o = 0; o < cout; ++o) {
    bp[o] = bias[o];
    for (int c = 0; c < cin; ++c)
      for (int t = 0; t < L->taps; ++t) {
        const float s = w_oihw[((size_t)o * cin + c) * L->taps + t] * w_scale;
        const __half h = __float2half_rn(s);
        const size_t idx = ((size_t)t * L->n_pad + o) * cin + c;
        hi[idx] = h;
        lo[idx] = __float2half_rn(s - __half2float(h));
      }
  }
  OSB_CUDA(cudaMalloc(&L->w_hi, n * sizeof(__half)));
  OSB_CUDA(cudaMalloc(&L->w_lo, n * sizeof(__half)));
  OSB_CUDA(cudaMalloc(&L->bias, L->n_pad * sizeof(float)));
  OSB_CUDA(cudaMemcpy(L->w_hi, hi.data(), n * sizeof(__half), cudaMemcpyHostToDevice));
  OSB_CUDA(cudaMemcpy(L->w_lo, lo.data(), n * sizeof(__half), cudaMemcpyHostToDevice));
  OSB_CUDA(cudaMemcpy(L->bias, bp.data(), L->n_pad * sizeof(float), cudaMemcpyHostToDevice));
  const uint64_t dims[3] = {(uint64_t)cin, (uint64_t)L->n_pad, (uint64_t)L->taps};
  const uint64_t strides[2] = {(uint64_t)cin * 2, (uint64_t)cin * L->n_pad * 2};
  const uint32_t box[3] = {UM_KC, (uint32_t)std::min(L->n_pad, 256), 1};
  osb_status s;
  if ((s = umma_make_tmap(&L->tm_hi, L->w_hi, 3, dims, strides, box)) != OSB_OK) return s;
  if ((s = umma_make_tmap(&L->tm_lo, L->w_lo, 3, dims, strides, box)) != OSB_OK) return s;
  if (L->n_pad == 64) {
    const uint32_t box32[3] = {UM_KC, 32, 1};
    if ((s = umma_make_tmap(&L->tm_hi32, L->w_hi, 3, dims, strides, box32)) != OSB_OK) return s;
  }
  if (L->n_pad >= 256) {                       // 128-row boxes: the layer as n_pad / 128 work items per tile
    const uint32_t box128[3] = {UM_KC, 128, 1};
    if ((s = umma_make_tmap(&L->tm_hi128, L->w_hi, 3, dims, strides, box128)) != OSB_OK) return s;
    if ((s = umma_make_tmap(&L->tm_lo128, L->w_lo, 3, dims, strides, box128)) != OSB_OK) return s;
  }
  return OSB_OK;
}

void umma_layer_free(UmmaLayer* L) {
  cudaFree(L->w_hi); cudaFree(L->w_lo); cudaFree(L->bias);
  L->w_hi = L->w_lo = nullptr; L->bias = nullptr;
}

osb_status umma_act_maps(CUtensorMap* hi, CUtensorMap* lo, __half* p_hi, __half* p_lo, int B, int H, int W, int C,
                         int ks) {
  const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  const uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
  // the box carries the vertical halo of the layer that READS these planes (ks x ks filter)
  const uint32_t box[4] = {UM_KC, UM_TW, (uint32_t)(UM_TH + 2 * (ks / 2)), 1};
  osb_status s;
  if ((s = umma_make_tmap(hi, p_hi, 4, dims, strides, box)) != OSB_OK) return s;
  return umma_make_tmap(lo, p_lo, 4, dims, strides, box);
}

// OSB_CONV_PDL=1 launches the convolutions with programmatic dependent launch.  Off by default: measured (r01f) it
// costs throughput here -- the dependent layer's CTAs take the SMs the concurrent NetVLAD stream was filling
// (2.10 ms per keyframe with it, 1.93 ms without).
static const bool g_conv_pdl = [] { const char* e = getenv("OSB_CONV_PDL"); return e && atoi(e) != 0; }();

// OSB_CONV_NSPLIT=0: layers of 256 / 512 output channels run through the N = 256 kernel (A/B switch)
static const bool g_conv_nsplit = [] { const char* e = getenv("OSB_CONV_NSPLIT"); return !(e && atoi(e) == 0); }();

template <int N, bool RES, bool SPLIT = false>
static osb_status launch_umma(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const UmmaLayer& L, const UmmaArgs& P,
                              cudaStream_t st, int max_ctas, bool box128 = false) {
  using Cfg = UmmaCfg<N, RES>;
  OSB_SMEM_OPT_IN((conv_umma_kernel<N, RES, SPLIT>), Cfg::SMEM_BYTES);
  const int tiles = P.B * cdiv(P.W, UM_TW) * cdiv(P.H, UM_TH) * P.n_split;
  // persistent CTAs, one per SM; `max_ctas` leaves SMs free for a kernel running beside this one on another stream
  const int grid = std::min(tiles, persistent_ctas(max_ctas));
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(UM_THREADS); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES; cfg.stream = st;
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_conv_pdl ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  OSB_CUDA(cudaLaunchKernelEx(&cfg, conv_umma_kernel<N, RES, SPLIT>, a_hi, a_lo, box128 ? L.tm_hi128 : L.tm_hi,
                              box128 ? L.tm_lo128 : L.tm_lo, P));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return OSB_OK;
}

osb_status umma_conv_forward(const UmmaLayer& L, const CUtensorMap& a_hi, const CUtensorMap& a_lo, int B, int H, int W,
                             float act_scale, __half* out_hi, __half* out_lo, float* out_f32, int out_c, int out_cstride,
                             float out_scale, int relu, int pool, cudaStream_t st, int max_ctas) {
  UmmaArgs P;
  P.pool = pool;
  OSB_REQUIRE(!pool || (H % 2 == 0 && W % 2 == 0), "fused max-pool needs even H and W");
  P.bias = L.bias; P.out_hi = out_hi; P.out_lo = out_lo; P.out_f32 = out_f32;
  P.H = H; P.W = W; P.B = B; P.ks = L.ks; P.cin_slabs = L.cin / UM_KC;
  P.out_c = out_c; P.out_cstride = out_cstride;
  P.inv_scale = 1.0f / (act_scale * L.w_scale); P.out_scale = out_scale; P.relu = relu;
  OSB_REQUIRE(out_c % 16 == 0 && out_c <= L.n_pad && out_cstride % 8 == 0, "tcgen05 conv: bad output channel layout");
  P.n_off = 0; P.n_split = 1; P.epi = 0;
  switch (L.n_pad) {
    case 64:
      if (L.ks == 3 && L.cin == UM_KC) return launch_umma<64, true>(a_hi, a_lo, L, P, st, max_ctas);   // weights resident
      return launch_umma<64, false>(a_hi, a_lo, L, P, st, max_ctas);
    case 80: return launch_umma<80, false>(a_hi, a_lo, L, P, st, max_ctas);
    case 128: return launch_umma<128, false>(a_hi, a_lo, L, P, st, max_ctas);
    case 256:
      if (g_conv_nsplit) {                        // 2 items of 128 channels per tile: TMEM double-buffered (the N = 256
        P.n_split = 2;                            // kernel is single-buffered), finer work items for the 320-tile layers
        return launch_umma<128, false, true>(a_hi, a_lo, L, P, st, max_ctas, true);
      }
      return launch_umma<256, false>(a_hi, a_lo, L, P, st, max_ctas);
    case 512: {
      if (g_conv_nsplit) {
        P.n_split = 4;
        return launch_umma<128, false, true>(a_hi, a_lo, L, P, st, max_ctas, true);
      }
      // two N = 256 passes over the same activations
      P.out_c = 256;
      osb_status s = launch_umma<256, false>(a_hi, a_lo, L, P, st, max_ctas);
      if (s != OSB_OK) return s;
      P.n_off = 256;
      return launch_umma<256, false>(a_hi, a_lo, L, P, st, max_ctas);
    }
  }
  set_error("umma_conv_forward", "unsupported N");
  return OSB_ERR_INVALID;
}

// detector head: convPb (256 -> 65, 1x1) with the softmax + 8x8 pixel shuffle fused into the epilogue; `semi` is the heat
// map [B][8H][8W]
osb_status umma_conv_softmax_forward(const UmmaLayer& L, const CUtensorMap& a_hi, const CUtensorMap& a_lo, int B, int H, int W,
                                     float act_scale, float* semi, cudaStream_t st, int max_ctas) {
  OSB_REQUIRE(L.n_pad == 80 && L.cout == 65 && L.ks == 1, "fused detector head expects the 65-logit 1x1 layer");
  UmmaArgs P;
  P.pool = 0; P.bias = L.bias; P.out_hi = nullptr; P.out_lo = nullptr; P.out_f32 = semi;
  P.H = H; P.W = W; P.B = B; P.ks = L.ks; P.cin_slabs = L.cin / UM_KC;
  P.out_c = 80; P.out_cstride = 80;
  P.inv_scale = 1.0f / (act_scale * L.w_scale); P.out_scale = 1.f; P.relu = 0;
  P.n_off = 0; P.n_split = 1; P.epi = 1;
  return launch_umma<80, false>(a_hi, a_lo, L, P, st, max_ctas);
}

// depthwise 3x3 (pad 1, stride s) + bias + ReLU6 on fp32 NHWC input, output as split fp16 planes for the pointwise
// tcgen05 conv that follows; one thread per (output pixel, 8 channels)
__global__ void dwconv3x3_split_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                       const float* __restrict__ x, __half* __restrict__ out_hi,
                                       __half* __restrict__ out_lo, int H, int W, int Ho, int Wo, int C, int stride,
                                       float out_scale, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int C8 = C >> 3;
  const int c = (int)(i % C8) * 8;
  int64_t p = i / C8;
  const int ox = (int)(p % Wo); p /= Wo;
  const int oy = (int)(p % Ho);
  const int b = (int)(p / Ho);
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int gy = oy * stride + ky - 1, gx = ox * stride + kx - 1;
      if (gy < 0 || gy >= H || gx < 0 || gx >= W) continue;
      const float4* xp = reinterpret_cast<const float4*>(x + (((size_t)b * H + gy) * W + gx) * C + c);
      const float4* wp = reinterpret_cast<const float4*>(w + (size_t)(ky * 3 + kx) * C + c);
      const float4 v0 = xp[0], v1 = xp[1], w0 = wp[0], w1 = wp[1];
      a[0] = fmaf(v0.x, w0.x, a[0]); a[1] = fmaf(v0.y, w0.y, a[1]); a[2] = fmaf(v0.z, w0.z, a[2]); a[3] = fmaf(v0.w, w0.w, a[3]);
      a[4] = fmaf(v1.x, w1.x, a[4]); a[5] = fmaf(v1.y, w1.y, a[5]); a[6] = fmaf(v1.z, w1.z, a[6]); a[7] = fmaf(v1.w, w1.w, a[7]);
    }
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = fminf(fmaxf(a[j] + bias[c + j], 0.f), 6.f);
  split_store8(out_hi + (size_t)i * 8, out_lo + (size_t)i * 8, a, out_scale);
}

// stride-1 variant: one thread per (4 consecutive output pixels of a row, 8 channels).  The 3 x 6 input window is read
// once (36 float4 instead of 72 for four single-pixel threads) and the 9 x 8 weights once per thread; same tap order per
// output as the kernel above, so the planes are bit-identical.
__global__ void dwconv3x3_split_s1x4_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                            const float* __restrict__ x, __half* __restrict__ out_hi,
                                            __half* __restrict__ out_lo, int H, int W, int C, float out_scale,
                                            int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int C8 = C >> 3, Wg = (W + 3) >> 2;
  const int c = (int)(i % C8) * 8;
  int64_t p = i / C8;
  const int ox0 = (int)(p % Wg) * 4; p /= Wg;
  const int oy = (int)(p % H);
  const int b = (int)(p / H);
  float4 wv[9][2];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float4* wp = reinterpret_cast<const float4*>(w + (size_t)t * C + c);
    wv[t][0] = __ldg(wp); wv[t][1] = __ldg(wp + 1);
  }
  float a[4][8];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int j = 0; j < 8; ++j) a[q][j] = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int gy = oy + ky - 1;
    if (gy < 0 || gy >= H) continue;
    float4 v[6][2];
#pragma unroll
    for (int cx = 0; cx < 6; ++cx) {
      const int gx = ox0 + cx - 1;
      if (gx >= 0 && gx < W) {
        const float4* xp = reinterpret_cast<const float4*>(x + (((size_t)b * H + gy) * W + gx) * C + c);
        v[cx][0] = xp[0]; v[cx][1] = xp[1];
      } else {
        v[cx][0] = v[cx][1] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int gx = ox0 + q + kx - 1;
        if (gx < 0 || gx >= W) continue;              // (skipped like the single-pixel kernel: no +0 added)
        const float4 v0 = v[q + kx][0], v1 = v[q + kx][1], w0 = wv[ky * 3 + kx][0], w1 = wv[ky * 3 + kx][1];
        a[q][0] = fmaf(v0.x, w0.x, a[q][0]); a[q][1] = fmaf(v0.y, w0.y, a[q][1]);
        a[q][2] = fmaf(v0.z, w0.z, a[q][2]); a[q][3] = fmaf(v0.w, w0.w, a[q][3]);
        a[q][4] = fmaf(v1.x, w1.x, a[q][4]); a[q][5] = fmaf(v1.y, w1.y, a[q][5]);
        a[q][6] = fmaf(v1.z, w1.z, a[q][6]); a[q][7] = fmaf(v1.w, w1.w, a[q][7]);
      }
  }
  float bb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bb[j] = __ldg(bias + c + j);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int ox = ox0 + q;
    if (ox >= W) break;
#pragma unroll
    for (int j = 0; j < 8; ++j) a[q][j] = fminf(fmaxf(a[q][j] + bb[j], 0.f), 6.f);
    const size_t o = ((((size_t)b * H + oy) * W + ox) * C8 + (c >> 3)) * 8;
    split_store8(out_hi + o, out_lo + o, a[q], out_scale);
  }
}

osb_status umma_dwconv_forward(const float* w_tap_c, const float* bias, const float* x, __half* out_hi, __half* out_lo,
                               int B, int H, int W, int C, int stride, float out_scale, cudaStream_t st) {
  const int Ho = H / stride, Wo = W / stride;
  if (stride == 1) {
    const int64_t total4 = (int64_t)B * H * ((W + 3) / 4) * (C / 8);
    OSB_LAUNCH(dwconv3x3_split_s1x4_kernel, (unsigned)cdiv64(total4, 128), 128, 0, st, w_tap_c, bias, x, out_hi, out_lo,
               H, W, C, out_scale, total4);
    OSB_CHECK_LAUNCH();
    return OSB_OK;
  }
  const int64_t total = (int64_t)B * Ho * Wo * (C / 8);
  OSB_LAUNCH(dwconv3x3_split_kernel, (unsigned)cdiv64(total, 256), 256, 0, st, w_tap_c, bias, x, out_hi, out_lo, H, W,
             Ho, Wo, C, stride, out_scale, total);
  OSB_CHECK_LAUNCH();
  return OSB_OK;
}

osb_status umma_first_forward(const float* w_tap_cout, const float* bias, const float* lut, const uint8_t* img,
                              __half* out_hi, __half* out_lo, int B, int H, int W, float out_scale, cudaStream_t st) {
  dim3 grid(cdiv(W, CF_TW), cdiv(H, CF_TH), B);
  OSB_LAUNCH(conv_first_split_kernel, grid, 256, 0, st, w_tap_cout, bias, lut, img, out_hi, out_lo, H, W, out_scale);
  OSB_CHECK_LAUNCH();
  return OSB_OK;
}

}  // namespace osb
