// conv1_fused.cu -- SuperPoint's first two layers as ONE kernel: conv1a (1 -> 64, 3x3, ReLU) is computed inside the SM and
// feeds conv1b (64 -> 64, 3x3, ReLU, fused 2x2 max-pool) on the tensor cores without ever leaving shared memory.
// (swarm_loop/superpoint.ipynb:143-146,165-167 of the reference)
//
// Unfused, conv1a wrote its output as split-fp16 planes (629 MB per 8-image keyframe) that conv1b immediately re-read:
// a 0.17 ms kernel plus 4.8x avoidable HBM traffic in the dominant layer.  Here the only HBM traffic of the pair is the u8
// image in and the pooled conv1b planes out.
//
// Tile = 16 rows x 8 pixels of conv1b output (M = 128, GEMM row m = r*8 + c).  Its conv1a halo tile is 18 rows x 10 pixels
// x 64 channels, kept ONCE in shared memory as two fp16 planes (hi, lo), pixel pitch 128 B, row pitch 1280 B, each pixel's
// eight 16-byte channel chunks XOR-swizzled with bits [7,10) of their absolute shared-memory address -- the K-major
// SWIZZLE_128B pattern.  All nine filter taps are views of that one copy: tap (ky,kx) is the UMMA descriptor whose start
// address is moved by (ky*10 + kx)*128 bytes, with SBO = 1280 B between the 8-pixel row groups.  That the tensor core
// derives its XOR phase from the absolute address (so that a start address inside a 1024-byte swizzle atom and an SBO that
// is not a multiple of 1024 read the right bytes) was measured on the B200 with scripts/microbench/umma_swizzle_probe.cu
// (profiles/r02_umma_swizzle_probe.txt: variant 0 exact for all nine taps).
//
// Shared memory (227 KB): conv1b weights resident, 9 taps x [W_hi | W_lo] = 144 KB; halo rows: 33 rows x 1280 B x 2 planes
// = 82.5 KB -- not the 36 rows two independent tiles would need.  Tiles therefore alternate between window 0 = rows
// [0,18) and window 1 = rows [15,33): 15 of a tile's 18 halo rows can be produced while the previous tile's MMAs are
// still reading theirs, the 3 overlapping rows are computed early, held in registers and stored as soon as those MMAs
// have retired.
//
// Warp roles (480 threads = 15 warps, 128 registers): warp 14 = TMEM allocator, weight TMA and MMA issuer (one lane chosen
// with elect.sync: guarded by `lane == 0` the compiler wraps every tcgen05.mma in an election loop, 9 dependent instructions
// per MMA, and the issuing thread -- not the tensor pipe -- sets the pace); warps 6-13 epilogue, two per TMEM lane quarter
// taking the 16-column chunks alternately (TMEM -> bias/ReLU -> 2x2 max-pool by shuffles -> re-split -> NHWC planes of
// conv2a's input); warps 0-5 conv1a producers with lane = halo pixel: warps 0-4 the 150 pixels of the 15 private rows,
// warp 5 the 30 pixels of the shared rows (all 64 channels computed into registers BEFORE it waits for the previous tile's
// MMAs).  With lane = pixel every lane of a warp uses the same weight: the 576 weights + 64 biases travel as a kernel
// parameter and reach the FFMA2s (packed fp32 pairs) through the constant bank and uniform registers; a thread converts its
// pixel's nine inputs once (u8 patch of the tile, 20 x 12 bytes, staged in shared memory one tile ahead).  The fp32 FMAs run
// in the order of conv_first_split_kernel, so the fused and unfused paths are bit-identical.  DESIGN.md section 3 lists what
// bound the kernel at each stage of its life and the measurements behind each change.
#include "conv_umma.cuh"
#include <type_traits>
#include "umma_ptx.cuh"

namespace osb {

constexpr int F1_TH = 16, F1_TW = 8;                    // conv1b output tile
constexpr int F1_HR = F1_TH + 2, F1_HC = F1_TW + 2;     // conv1a halo tile: 18 x 10 pixels
constexpr int F1_PITCH = F1_HC * 128;                   // bytes per halo row per plane
constexpr int F1_ROWS = 33;                             // window 0 = rows [0,18), window 1 = rows [15,33)
constexpr int F1_WIN1 = F1_ROWS - F1_HR;                // 15: first row of window 1
constexpr int F1_PLANE = F1_ROWS * F1_PITCH;            // 42 240 B
constexpr int F1_W_SLOT = 2 * 64 * 128;                 // one tap: [W_hi (64 rows) | W_lo (64 rows)] x 128 B
constexpr int F1_W_BYTES = 9 * F1_W_SLOT;               // 147 456 B
constexpr int F1_NPROD = 6;                             // producer warps: lane = halo pixel, warp = block of 32 pixels
constexpr int F1_EPI0 = 6;                              // warps 6..13: epilogue, two per TMEM lane quarter (= warp & 3)
constexpr int F1_MMAW = 14;                             // warp 14: TMEM allocation, weight TMA, MMA issue
constexpr int F1_THREADS = 15 * 32;                     // 480 (16-warp allocation: 128 registers per thread)
constexpr int F1_BAR_OFF = F1_W_BYTES + 2 * F1_PLANE;
constexpr int F1_PR = F1_HR + 2, F1_PC = F1_HC + 2;     // u8 input patch of a tile: 20 rows x 12 columns
constexpr int F1_PATCH_OFF = F1_BAR_OFF + 80;           // 9 mbarriers + the TMEM base slot, then the patch
constexpr int F1_SMEM = F1_PATCH_OFF + F1_PR * F1_PC;
static_assert(F1_SMEM <= 227 * 1024, "shared memory plan exceeds 227 KB");
static_assert(F1_PR * F1_PC <= 2 * F1_NPROD * 32, "two patch bytes per producer thread");

// measurement only (make EXTRA=-DF1_ABLATE=<bits> after touching this file, results are WRONG when set): 1 producers do not store, 2 no patch loads,
// 4 epilogue only drains TMEM, 8 no lo*hi MMA, 16 producers do not compute.  profiles/r02_f1_ablate.txt
#ifndef F1_ABLATE
#define F1_ABLATE 0
#endif
constexpr int kAblate = F1_ABLATE;

struct Fused1Args {
  const uint8_t* img;      // [B][H][W]
  const float* w1a;        // conv1a weights [tap][64]
  const float* b1a;        // conv1a bias [64]
  const float* bias;       // conv1b bias [64]
  __half* out_hi;          // pooled conv1b output planes [B][H/2][W/2][64]
  __half* out_lo;
  int H, W, B;
  float alpha;             // (float)(1/255): u8 -> f32 as cv::Mat::convertTo does (superpoint_tensorrt.cpp:127)
  float act_scale;         // scale of the conv1a planes (power of two)
  float inv_scale;         // 1 / (act_scale * w_scale)
  float out_scale;         // scale of the stored planes
  unsigned long long* dbg; // optional [16] cycle counters of CTA 0 (null = off): see umma_conv1_fused_forward
  int pool;                // 1: fused 2x2 max-pool (output [B][H/2][W/2][64]); 0: output [B][H][W][64]   (FIRST = false only)
};

__device__ __forceinline__ uint64_t umma_desc_sw128_sbo(uint32_t saddr, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void st_shared_128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// FIRST = true : the halo tile is COMPUTED (conv1a from the u8 image, 6 producer warps) -- SuperPoint's first two layers.
// FIRST = false: the halo tile is LOADED -- the same kernel as a 64 -> 64 3x3 layer on split-fp16 planes (conv2a, conv2b):
//                one TMA box per window part instead of the three kx-shifted boxes of conv_umma_kernel (46 KB of
//                activations per tile instead of 120 KB through L2 and shared memory).  A window is two boxes per plane:
//                {64 ch, 10 px, 15 rows} for the rows private to it and {64 ch, 10 px, 3 rows} for the rows it shares with
//                the other window, the latter issued only after the previous tile's MMAs have retired.  TMA writes
//                SWIZZLE_128B by absolute shared-memory address too, 128-byte aligned destinations suffice
//                (scripts/microbench/tma_swizzle_probe.cu, profiles/r02_tma_swizzle_probe.txt).
template <bool FIRST>
__global__ void __launch_bounds__(F1_THREADS, 1)
conv64_halo_kernel(const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                   const __grid_constant__ CUtensorMap tm_a15_hi, const __grid_constant__ CUtensorMap tm_a15_lo,
                   const __grid_constant__ CUtensorMap tm_a3_hi, const __grid_constant__ CUtensorMap tm_a3_lo,
                   const __grid_constant__ Conv1aW W1, Fused1Args P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) __trap();                              // the plan has no slack for re-aligning
  const uint32_t a_hi_base = base + F1_W_BYTES, a_lo_base = a_hi_base + F1_PLANE;
  const uint32_t bar_base = base + F1_BAR_OFF;
  const uint32_t b_full = bar_base;
  auto a_full = [&](int w) { return bar_base + 8u * (1 + w); };
  auto mma_done = [&](int w) { return bar_base + 8u * (3 + w); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (5 + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (7 + a); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_raw + F1_BAR_OFF + 72);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_x = (P.W + F1_TW - 1) / F1_TW, tiles_y = (P.H + F1_TH - 1) / F1_TH;
  const int n_tiles = P.B * tiles_x * tiles_y;

  if (threadIdx.x == 0) {
    mbar_init(b_full, 1);
    for (int w = 0; w < 2; ++w) { mbar_init(a_full(w), FIRST ? F1_NPROD : 1); mbar_init(mma_done(w), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == F1_MMAW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(bar_base + 72u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == F1_MMAW && elect_one()) {
    // ===================== conv1b weights: all 9 taps once, resident for the CTA's life =====================
    mbar_expect_tx(b_full, F1_W_BYTES);
    for (int t = 0; t < 9; ++t) {
      const uint32_t sb = base + t * F1_W_SLOT;
      tma_load_3d(sb, &tm_w_hi, b_full, 0, 0, t);
      tma_load_3d(sb + F1_W_SLOT / 2, &tm_w_lo, b_full, 0, 0, t);
    }
    // ===================== MMA issuer (same thread) =====================
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);    // N = 64
    constexpr uint32_t idesc2 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // N = 128: [main | cross]
    int acc = 0; uint32_t acc_phase = 0;
    uint32_t i = 0;
    mbar_wait(b_full, 0);
    const bool prof = P.dbg != nullptr && blockIdx.x == 0;
    long long c_te = 0, c_af = 0, c_is = 0, t0 = 0, t1 = 0, t2 = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++i) {
      const int w = i & 1;
      if (prof) t0 = clock64();
      mbar_wait(tempty_bar(acc), acc_phase ^ 1);
      if (prof) t1 = clock64();
      mbar_wait(a_full(w), (i >> 1) & 1);
      if (prof) t2 = clock64();
      tc_fence_after();
      const uint32_t d_main = tmem_base + (uint32_t)(acc * 128), d_cross = d_main + 64;
      const uint32_t row0 = w ? F1_WIN1 : 0;
      uint32_t first = 1;
      // tap order (kx outer, ky inner, K ascending) = the accumulation order of conv_umma_kernel: bit-identical sums
#pragma unroll 1
      for (int kx = 0; kx < 3; ++kx) {                       // (not unrolled: 27 hoisted descriptors would spill)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const uint32_t off = ((row0 + ky) * F1_HC + kx) * 128;
          const uint64_t a_hi = umma_desc_sw128_sbo(a_hi_base + off, F1_PITCH);
          const uint64_t a_lo = umma_desc_sw128_sbo(a_lo_base + off, F1_PITCH);
          const uint32_t sb = base + (ky * 3 + kx) * F1_W_SLOT;
          const uint64_t b_hi = umma_desc_sw128(sb);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);
            // one MMA of width 128 over [W_hi | W_lo]: hi*hi -> main, hi*lo -> cross; then lo*hi -> cross
            umma_f16(d_main, a_hi + adv, b_hi + adv, idesc2, (first && k == 0) ? 0u : 1u);
            if (!(kAblate & 8)) umma_f16(d_cross, a_lo + adv, b_hi + adv, idesc, 1u);
          }
          first = 0;
        }
      }
      umma_commit(mma_done(w));                    // the halo window may be overwritten
      umma_commit(tfull_bar(acc));                 // accumulators complete -> epilogue
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if (prof) { c_te += t1 - t0; c_af += t2 - t1; c_is += clock64() - t2; }
    }
    if (prof) { P.dbg[4] = c_te; P.dbg[5] = c_af; P.dbg[6] = c_is; P.dbg[9] = i; }
  } else if (warp >= F1_EPI0 && warp < F1_EPI0 + 8) {
    // ===================== epilogue: eight warps, two per TMEM lane quarter, taking the 16-column chunks alternately (one
    //                       warp per quarter is slower than the MMAs: a chunk is one long dependent chain) =====================
    const int q = warp & 3;                        // TMEM lane quarter = tile rows 4q .. 4q+3 (lane = (row & 3) * 8 + col)
    int acc = 0; uint32_t acc_phase = 0;
    const int Hp = P.H >> 1, Wp = P.W >> 1;
    const bool prof = P.dbg != nullptr && blockIdx.x == 0 && warp == F1_EPI0 && lane == 0;
    const int eset = (warp - F1_EPI0) >> 2;
    constexpr int NSET = 2;
    long long c_wait = 0, c_work = 0, t0 = 0, t1 = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      if (prof) t0 = clock64();
      const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
      // pooled pixel of (even row, even col): max over lanes l, l+1, l+8, l+9; writer lanes have row and col even
      const bool pool = FIRST || P.pool;
      const int py = ty * (F1_TH / 2) + 2 * q + (lane >> 4), px = tx * (F1_TW / 2) + ((lane & 7) >> 1);
      const int y = ty * F1_TH + 4 * q + (lane >> 3), x = tx * F1_TW + (lane & 7);
      const bool writer = pool ? (!(lane & 8) && !(lane & 1) && py < Hp && px < Wp) : (y < P.H && x < P.W);
      const size_t ppix = pool ? ((size_t)b * Hp + py) * Wp + px : ((size_t)b * P.H + y) * P.W + x;
      mbar_wait(tfull_bar(acc), acc_phase);
      if (prof) t1 = clock64();
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 128);
#pragma unroll 1
      for (int n0 = eset * 16; n0 < 64; n0 += 16 * NSET) {
        uint32_t v[16], vc[16];
        tmem_ld16(t_row + n0, v);
        tmem_ld16(t_row + 64 + n0, vc);
        if (kAblate & 4) continue;
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float a = fmaf(__uint_as_float(v[i]) + __uint_as_float(vc[i]), P.inv_scale, __ldg(P.bias + n0 + i));
          f[i] = fmaxf(a, 0.f);
        }
        if (pool) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            f[i] = fmaxf(f[i], __shfl_down_sync(0xffffffffu, f[i], 1));
            f[i] = fmaxf(f[i], __shfl_down_sync(0xffffffffu, f[i], 8));
          }
        }
        if (!writer) continue;
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float s0 = f[2 * i] * P.out_scale, s1 = f[2 * i + 1] * P.out_scale;
          const __half2 hp = __floats2half2_rn(s0, s1);
          const float2 hf = __half22float2(hp);
          const __half2 lp = __floats2half2_rn(s0 - hf.x, s1 - hf.y);
          hi[i] = *reinterpret_cast<const uint32_t*>(&hp);
          lo[i] = *reinterpret_cast<const uint32_t*>(&lp);
        }
        st_global_256(P.out_hi + ppix * 64 + n0, hi[0], hi[1], hi[2], hi[3], hi[4], hi[5], hi[6], hi[7]);
        st_global_256(P.out_lo + ppix * 64 + n0, lo[0], lo[1], lo[2], lo[3], lo[4], lo[5], lo[6], lo[7]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if (prof) { c_wait += t1 - t0; c_work += clock64() - t1; }
    }
    if (prof) { P.dbg[7] = c_wait; P.dbg[8] = c_work; }
  } else if (!FIRST && warp == 0 && elect_one()) {
    // ===================== halo windows by TMA (64 -> 64 layers on split planes) =====================
    asm volatile("griddepcontrol.wait;" ::: "memory");       // the planes are the previous kernel's output (no-op without PDL)
    uint32_t i = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++i) {
      const int w = i & 1;
      const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
      const int x0 = tx * F1_TW - 1, y0 = ty * F1_TH - 1;     // image coordinates of halo pixel (0, 0)
      // halo rows private to the window: 0..14 of window 0 (smem rows 0..14), 3..17 of window 1 (smem rows 18..32);
      // shared: 15..17 of window 0 = 0..2 of window 1 (smem rows 15..17)
      const int prow = w ? 3 : 0, srow = w ? 0 : 15;
      const uint32_t pdst = (uint32_t)((w ? F1_WIN1 + 3 : 0) * F1_PITCH), sdst = (uint32_t)(F1_WIN1 * F1_PITCH);
      if (i >= 2) mbar_wait(mma_done(w), ((i >> 1) - 1) & 1);
      mbar_expect_tx(a_full(w), 2u * F1_HR * F1_PITCH);
      tma_load_4d(a_hi_base + pdst, &tm_a15_hi, a_full(w), 0, x0, y0 + prow, b);
      tma_load_4d(a_lo_base + pdst, &tm_a15_lo, a_full(w), 0, x0, y0 + prow, b);
      if (i >= 1) mbar_wait(mma_done(w ^ 1), ((i - 1) >> 1) & 1);
      tma_load_4d(a_hi_base + sdst, &tm_a3_hi, a_full(w), 0, x0, y0 + srow, b);
      tma_load_4d(a_lo_base + sdst, &tm_a3_lo, a_full(w), 0, x0, y0 + srow, b);
    }
  } else if (FIRST && warp < F1_NPROD) {
    // ===================== conv1a producers (6 warps) =====================
    // lane = halo pixel.  Warps 0..4 take the 150 pixels of the rows private to the window, warp 5 the 30 pixels of the
    // three rows shared with the other window, which may only be overwritten after the previous tile's MMAs have retired.
    // A thread converts its pixel's nine inputs once and runs the 64 channels in eight groups of eight: 36 FFMA2 per group
    // with the weights from the constant bank, then ReLU, the fp16 split and one 16-byte store per plane (the eight pixels
    // of a quarter-warp sit in eight consecutive 128-byte rows, whose swizzle phases differ: conflict-free).
    const int pw = warp;
    const bool sh = pw == 5;
    const int q = sh ? lane : pw * 32 + lane;                 // pixel index inside its part
    const bool active = sh ? lane < 30 : q < 150;
    const int qc = sh ? min(q, 29) : min(q, 149);
    const int rr = qc / F1_HC, cc = qc - rr * F1_HC;
    const int ptid = warp * 32 + lane;                        // 0 .. 191: stages patch bytes ptid and ptid + 192
    uint8_t* patch = smem_raw + F1_PATCH_OFF;
    asm volatile("griddepcontrol.wait;" ::: "memory");       // the image is the previous kernel's output (no-op without PDL)
    // input patch of a tile: rows y0-2 .. y0+17, columns x0-2 .. x0+9 of the u8 image, zero outside (conv1a's padding)
    auto patch_byte = [&](int tile, int idx) -> uint32_t {
      const int pr = idx / F1_PC, pc = idx - pr * F1_PC;
      const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
      const int gy = ty * F1_TH - 2 + pr, gx = tx * F1_TW - 2 + pc;
      const bool in = idx < F1_PR * F1_PC && gy >= 0 && gy < P.H && gx >= 0 && gx < P.W;
      return in ? (uint32_t)__ldg(P.img + ((size_t)b * P.H + gy) * P.W + gx) : 0u;
    };
    constexpr int NPT = F1_NPROD * 32;
    int tile = blockIdx.x;
    patch[ptid] = (uint8_t)(tile < n_tiles ? patch_byte(tile, ptid) : 0u);
    if (ptid + NPT < F1_PR * F1_PC) patch[ptid + NPT] = (uint8_t)(tile < n_tiles ? patch_byte(tile, ptid + NPT) : 0u);
    asm volatile("bar.sync 1, 192;" ::: "memory");
    uint32_t i = 0;
    const bool prof = P.dbg != nullptr && blockIdx.x == 0 && lane == 0;        // every producer warp: dbg[10 + warp] = its busy cycles
    long long c_w1 = 0, c_cmp = 0, t0 = 0, t1 = 0, t2 = 0;
    const long long t_begin = clock64();
    for (; tile < n_tiles; tile += gridDim.x, ++i) {
      const int w = i & 1;
      const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y;
      const int next = tile + gridDim.x;
      // next tile's patch bytes: in flight for the whole tile
      const uint32_t nb0 = next < n_tiles ? patch_byte(next, ptid) : 0u;
      const uint32_t nb1 = next < n_tiles ? patch_byte(next, ptid + NPT) : 0u;
      // halo row of the pixel and its byte offset in a plane: private rows 0..14 of window 0 sit in buffer rows 0..14, rows
      // 3..17 of window 1 in buffer rows 18..32; the shared rows (15..17 of window 0 = 0..2 of window 1) in rows 15..17
      const int r = sh ? rr + (w ? 0 : 15) : rr + (w ? 3 : 0);
      const uint32_t off = (uint32_t)(((sh ? F1_WIN1 : (w ? F1_WIN1 + 3 : 0)) * F1_HC + qc) * 128);
      const uint32_t ah = a_hi_base + off, al = a_lo_base + off;
      const uint32_t ph = (ah >> 7) & 7u, pl = (al >> 7) & 7u;          // swizzle phases of the pixel's two rows
      // outside the image: conv1b's zero padding (only tiles on the image border have such pixels)
      const int iy = ty * F1_TH - 1 + r, ix = tx * F1_TW - 1 + cc;
      const bool valid = iy >= 0 && iy < P.H && ix >= 0 && ix < P.W;
      float in[9];
      {
        const uint8_t* pp = patch + r * F1_PC + cc;
#pragma unroll
        for (int t = 0; t < 9; ++t)
          in[t] = (kAblate & 2) ? (float)(t + r) : __fmul_rn(__uint2float_rn((uint32_t)pp[(t / 3) * F1_PC + t % 3]), P.alpha);
      }
      if (prof) t0 = clock64();
      // the rows may be overwritten once the MMAs that read them have retired: the tile two back (same window) for the
      // private rows -- long gone -- and the PREVIOUS tile for the shared ones, which therefore sit on the critical path
      // between two tiles' MMAs: their warp computes all 64 channels into registers first and only then waits, so that
      // nothing but the stores is left to do when the rows are released
      if (!sh && i >= 2) mbar_wait(mma_done(w), ((i >> 1) - 1) & 1);
      if (prof) t1 = clock64();
      // one group of eight channels -> packed split fp16 (hi, lo); taps ascending for every channel: the fma chain of
      // conv_first_split_kernel, bit for bit
      auto group = [&](auto G, uint32_t (&h)[4], uint32_t (&l)[4]) {
        constexpr int C0 = decltype(G)::value * 8;
        uint64_t acc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = pk2(W1.b[C0 + 2 * j], W1.b[C0 + 2 * j + 1]);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const uint64_t vv = pk2(in[t], in[t]);
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[j] = fma2(vv, pk2(W1.w[t][C0 + 2 * j], W1.w[t][C0 + 2 * j + 1]), acc[j]);
        }
#pragma unroll
        for (int j2 = 0; j2 < 4; ++j2) {
          float a0, a1;
          upk2(acc[j2], a0, a1);
          const float s0 = fmaxf(a0, 0.f), s1 = fmaxf(a1, 0.f);
          // one packed conversion per pair (cvt.rn.f16x2.f32: same roundings as two scalar conversions)
          const __half2 hp = __floats2half2_rn(s0, s1);
          const float2 hf = __half22float2(hp);
          float d0, d1;
          upk2(sub2(pk2(s0, s1), pk2(hf.x, hf.y)), d0, d1);
          const __half2 lp = __floats2half2_rn(d0, d1);
          h[j2] = valid ? *reinterpret_cast<const uint32_t*>(&hp) : 0u;
          l[j2] = valid ? *reinterpret_cast<const uint32_t*>(&lp) : 0u;
        }
      };
      auto put = [&](int g, const uint32_t (&h)[4], const uint32_t (&l)[4]) {
        if (!active || (kAblate & 1)) return;
        st_shared_128(ah + (((uint32_t)g ^ ph) << 4), h[0], h[1], h[2], h[3]);
        st_shared_128(al + (((uint32_t)g ^ pl) << 4), l[0], l[1], l[2], l[3]);
      };
      if (!(kAblate & 16)) {
        if (!sh) {
          auto run = [&](auto G) { uint32_t h[4], l[4]; group(G, h, l); put(decltype(G)::value, h, l); };
          run(std::integral_constant<int, 0>{}); run(std::integral_constant<int, 1>{});
          run(std::integral_constant<int, 2>{}); run(std::integral_constant<int, 3>{});
          run(std::integral_constant<int, 4>{}); run(std::integral_constant<int, 5>{});
          run(std::integral_constant<int, 6>{}); run(std::integral_constant<int, 7>{});
        } else {
          uint32_t hq[8][4], lq[8][4];
          group(std::integral_constant<int, 0>{}, hq[0], lq[0]); group(std::integral_constant<int, 1>{}, hq[1], lq[1]);
          group(std::integral_constant<int, 2>{}, hq[2], lq[2]); group(std::integral_constant<int, 3>{}, hq[3], lq[3]);
          group(std::integral_constant<int, 4>{}, hq[4], lq[4]); group(std::integral_constant<int, 5>{}, hq[5], lq[5]);
          group(std::integral_constant<int, 6>{}, hq[6], lq[6]); group(std::integral_constant<int, 7>{}, hq[7], lq[7]);
          if (i >= 1) mbar_wait(mma_done(w ^ 1), ((i - 1) >> 1) & 1);
#pragma unroll
          for (int g = 0; g < 8; ++g) put(g, hq[g], lq[g]);
        }
      } else if (sh && i >= 1) {
        mbar_wait(mma_done(w ^ 1), ((i - 1) >> 1) & 1);
      }
      if (prof) t2 = clock64();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full(w));
      // hand the patch over to the next tile: everyone has read this one, then everyone sees the next
      asm volatile("bar.sync 1, 192;" ::: "memory");
      patch[ptid] = (uint8_t)nb0;
      if (ptid + NPT < F1_PR * F1_PC) patch[ptid + NPT] = (uint8_t)nb1;
      asm volatile("bar.sync 1, 192;" ::: "memory");
      if (prof) { c_w1 += t1 - t0; c_cmp += t2 - t1; }
    }
    if (prof && warp == 0) { P.dbg[0] = c_w1; P.dbg[1] = c_cmp; P.dbg[2] = 0; P.dbg[3] = clock64() - t_begin; }
    if (prof) P.dbg[10 + warp] = c_cmp + c_w1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == F1_MMAW) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
  }
}

// conv1a + ReLU + conv1b + ReLU + 2x2 max-pool: u8 images -> the split planes conv2a reads.  L1b = conv1b's weights
// (n_pad 64, Cin 64, 3x3) as uploaded by umma_layer_upload; w1a [tap][64], b1a [64] fp32.
osb_status umma_conv1_fused_forward(const UmmaLayer& L1b, const float* w1a_host, const float* b1a_host, const uint8_t* img, int B,
                                    int H, int W, float act_scale, __half* out_hi, __half* out_lo, float out_scale, cudaStream_t st,
                                    int max_ctas, unsigned long long* dbg) {
  OSB_REQUIRE(L1b.n_pad == 64 && L1b.cin == 64 && L1b.ks == 3, "fused first layers expect the 64 -> 64 3x3 layer");
  OSB_REQUIRE(H % 2 == 0 && W % 2 == 0, "fused max-pool needs even H and W");
  Fused1Args P;
  P.img = img; P.w1a = nullptr; P.b1a = nullptr; P.bias = L1b.bias; P.out_hi = out_hi; P.out_lo = out_lo;
  P.H = H; P.W = W; P.B = B;
  P.alpha = (float)(1.0 / 255.0);
  P.dbg = dbg; P.pool = 1;
  Conv1aW W1;                                             // the plane scale is a power of two: the products are exact
  for (int t = 0; t < 9; ++t)
    for (int c = 0; c < 64; ++c) W1.w[t][c] = w1a_host[t * 64 + c] * act_scale;
  for (int c = 0; c < 64; ++c) W1.b[c] = b1a_host[c] * act_scale;
  P.act_scale = act_scale; P.inv_scale = 1.0f / (act_scale * L1b.w_scale); P.out_scale = out_scale;
  OSB_SMEM_OPT_IN(conv64_halo_kernel<true>, F1_SMEM);
  const int tiles = B * cdiv(W, F1_TW) * cdiv(H, F1_TH);
  const int grid = std::min(tiles, persistent_ctas(max_ctas));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(F1_THREADS); cfg.dynamicSmemBytes = F1_SMEM; cfg.stream = st;
  // (the activation maps are unused by the FIRST instantiation: the weight maps stand in)
  OSB_CUDA(cudaLaunchKernelEx(&cfg, conv64_halo_kernel<true>, L1b.tm_hi, L1b.tm_lo, L1b.tm_hi, L1b.tm_lo, L1b.tm_hi, L1b.tm_lo, W1, P));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return OSB_OK;
}

// TMA descriptors of a 64-channel activation tensor for the halo-window kernel: boxes {64 ch, 10 px, 15 rows} and
// {64 ch, 10 px, 3 rows} on each plane ([B][H][W][64] fp16)
osb_status umma_halo_maps(HaloMaps* M, __half* p_hi, __half* p_lo, int B, int H, int W) {
  const uint64_t dims[4] = {64, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  const uint64_t strides[3] = {64 * 2, (uint64_t)W * 64 * 2, (uint64_t)H * W * 64 * 2};
  const uint32_t box15[4] = {64, F1_HC, F1_WIN1, 1}, box3[4] = {64, F1_HC, F1_HR - F1_WIN1, 1};
  osb_status s;
  if ((s = umma_make_tmap(&M->a15_hi, p_hi, 4, dims, strides, box15)) != OSB_OK) return s;
  if ((s = umma_make_tmap(&M->a15_lo, p_lo, 4, dims, strides, box15)) != OSB_OK) return s;
  if ((s = umma_make_tmap(&M->a3_hi, p_hi, 4, dims, strides, box3)) != OSB_OK) return s;
  return umma_make_tmap(&M->a3_lo, p_lo, 4, dims, strides, box3);
}

// 64 -> 64 3x3 convolution + bias + ReLU (+ optional 2x2 max-pool) on split planes with the halo-window kernel
osb_status umma_conv64_halo_forward(const UmmaLayer& L, const HaloMaps& M, int B, int H, int W, float act_scale, __half* out_hi,
                                    __half* out_lo, float out_scale, int pool, cudaStream_t st, int max_ctas,
                                    unsigned long long* dbg) {
  OSB_REQUIRE(L.n_pad == 64 && L.cin == 64 && L.ks == 3, "halo-window kernel expects a 64 -> 64 3x3 layer");
  OSB_REQUIRE(!pool || (H % 2 == 0 && W % 2 == 0), "fused max-pool needs even H and W");
  Fused1Args P;
  P.img = nullptr; P.w1a = nullptr; P.b1a = nullptr; P.bias = L.bias; P.out_hi = out_hi; P.out_lo = out_lo;
  P.H = H; P.W = W; P.B = B; P.alpha = 0.f; P.dbg = dbg; P.pool = pool;
  P.act_scale = act_scale; P.inv_scale = 1.0f / (act_scale * L.w_scale); P.out_scale = out_scale;
  OSB_SMEM_OPT_IN(conv64_halo_kernel<false>, F1_SMEM);
  const int tiles = B * cdiv(W, F1_TW) * cdiv(H, F1_TH);
  const int grid = std::min(tiles, persistent_ctas(max_ctas));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(F1_THREADS); cfg.dynamicSmemBytes = F1_SMEM; cfg.stream = st;
  static const Conv1aW no_w1 = {};
  OSB_CUDA(cudaLaunchKernelEx(&cfg, conv64_halo_kernel<false>, L.tm_hi, L.tm_lo, M.a15_hi, M.a15_lo, M.a3_hi, M.a3_lo, no_w1, P));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return OSB_OK;
}

}  // namespace osb
