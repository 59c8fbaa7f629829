// conv64_pair.cu -- the 64 -> 64 3x3 layers of SuperPoint (conv1a+conv1b fused, conv2a, conv2b: 66 % of the network's
// FLOPs) on CTA PAIRS: tcgen05.mma.cta_group::2, M = 256.
//
// Same algorithm as conv1_fused.cu (one shared-memory halo tile per output tile, nine descriptor views, split-fp16 with a
// main and a cross accumulator), but two CTAs of a cluster -- two SMs of one TPC -- issue their MMAs together: the
// instruction covers the leader's 16 x 8 tile (rows 0..127 of D, in the leader's TMEM) and the peer's tile (rows 128..255,
// in the peer's TMEM), reads each CTA's halo tile from that CTA's shared memory, and reads the B operand HALF from each:
//   MMA 1 (N = 128, [W_hi | W_lo] -> [main | cross]): rows 0..63 = W_hi from the leader, rows 64..127 = W_lo from the peer;
//   MMA 2 (N = 64,  W_hi -> cross with A_lo):         rows 0..31 from the leader, rows 32..63 from the peer.
// So a CTA keeps 12 KB of weights per tap instead of 16 KB (108 KB instead of 144 KB), which is what makes room for TWO full
// halo windows (2 x 18 rows x 2 planes = 90 KB): a tile's halo can be produced / loaded entirely while the previous tile's
// MMAs run -- the single-CTA kernel had to share three rows between its windows and paid a bubble per tile for them -- and a
// CTA fetches 11 KB of operands per K step instead of 14 KB, which moves the N = 64 MMA from shared-memory-bandwidth bound
// to tensor-pipe bound.
// Barriers: producers and epilogue warps of BOTH CTAs arrive on the leader's mbarriers (mapa + mbarrier.arrive
// .shared::cluster); the leader's tcgen05.commit multicasts "accumulators full" and "halo window free" to both CTAs.
// FIRST = true: halo tile computed from the u8 image (conv1a, 6 producer warps per CTA); false: loaded by TMA (conv2a/2b).
// Bit-identical to conv_umma_kernel<64,RES> / conv1_fused.cu: same operand values, same accumulation order per output row.
#include "conv_umma.cuh"
#include "umma_ptx.cuh"
#include <type_traits>

namespace osb {

constexpr int P2_TH = 16, P2_TW = 8;
constexpr int P2_HR = P2_TH + 2, P2_HC = P2_TW + 2;
constexpr int P2_PITCH = P2_HC * 128;                   // 1280 B per halo row per plane
constexpr int P2_WIN = P2_HR * P2_PITCH;                // 23 040 B: one window of one plane
constexpr int P2_PLANE = 2 * P2_WIN;                    // two windows
constexpr int P2_WX = 64 * 128, P2_WY = 32 * 128;       // weight regions of a tap: X = 64 rows, Y = 32 rows
constexpr int P2_W_SLOT = P2_WX + P2_WY;                // 12 288 B
constexpr int P2_W_BYTES = 9 * P2_W_SLOT;               // 110 592 B
constexpr int P2_NPROD = 6;                             // producer warps: lane = halo pixel (180 of 192 lanes)
constexpr int P2_EPI0 = 6;                              // warps 6..13: epilogue, two per TMEM lane quarter
constexpr int P2_MMAW = 14;                             // warp 14: TMEM allocation, weight TMA, MMA issue (leader)
constexpr int P2_THREADS = 15 * 32;
constexpr int P2_BAR_OFF = P2_W_BYTES + 2 * P2_PLANE;   // 202 752
constexpr int P2_PR = P2_HR + 2, P2_PC = P2_HC + 2;     // u8 patch 20 x 12
constexpr int P2_PATCH_OFF = P2_BAR_OFF + 128;          // 12 mbarriers + TMEM slot
constexpr int P2_SMEM = P2_PATCH_OFF + 2 * 256;         // two patches (one tile ahead)

struct PairArgs {
  const uint8_t* img; const float* w1a; const float* b1a; const float* bias;
  __half* out_hi; __half* out_lo;
  int H, W, B;
  float alpha, act_scale, inv_scale, out_scale;
  int pool;
  unsigned long long* dbg;
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a barrier that threads of the PEER CTA arrive on: acquire at cluster scope
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done, spins = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done && ++spins > (1u << 24)) __trap();
  } while (!done);
}
__device__ __forceinline__ uint32_t ld_shared_cluster_u32(uint32_t cluster_addr) {
  uint32_t v; asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(cluster_addr) : "memory"); return v;
}
__device__ __forceinline__ void umma_f16_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {        // arrives on `bar` in BOTH CTAs of the pair
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ uint64_t p2_desc_sbo(uint32_t saddr, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void p2_st_shared_128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <bool FIRST>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P2_THREADS, 1)
conv64_pair_kernel(const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                   const __grid_constant__ CUtensorMap tm_w_hi32, const __grid_constant__ CUtensorMap tm_a_hi,
                   const __grid_constant__ CUtensorMap tm_a_lo, const __grid_constant__ Conv1aW W1, PairArgs P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) __trap();
  const uint32_t rank = cluster_ctarank();                   // 0 = leader (issues the MMAs)
  const uint32_t a_hi_base = base + P2_W_BYTES, a_lo_base = a_hi_base + P2_PLANE;
  const uint32_t bar_base = base + P2_BAR_OFF;
  const uint32_t b_full = bar_base;                          // local: own weights landed
  const uint32_t w_ready = bar_base + 8;                     // leader's: both CTAs' weights landed (count 2)
  auto a_full = [&](int w) { return bar_base + 8u * (2 + w); };     // leader's: both halo tiles of the pair are complete
  auto a_local = [&](int w) { return bar_base + 8u * (4 + w); };    // local TMA completion (FIRST = false)
  auto mma_done = [&](int w) { return bar_base + 8u * (6 + w); };   // both: window w may be overwritten
  auto tfull_bar = [&](int a) { return bar_base + 8u * (8 + a); };  // both: accumulators complete
  auto tempty_bar = [&](int a) { return bar_base + 8u * (10 + a); };// leader's: both epilogues have drained buffer a (count 8)
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_raw + P2_BAR_OFF + 96);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_x = (P.W + P2_TW - 1) / P2_TW, tiles_y = (P.H + P2_TH - 1) / P2_TH;
  const int n_tiles = P.B * tiles_x * tiles_y;
  const int n_pairs = (n_tiles + 1) >> 1;
  const int n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;

  if (threadIdx.x == 0) {
    mbar_init(b_full, 1);
    mbar_init(w_ready, 2);
    for (int w = 0; w < 2; ++w) {
      mbar_init(a_full(w), FIRST ? 2 * P2_NPROD : 2);
      mbar_init(a_local(w), 1);
      mbar_init(mma_done(w), 1);
    }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == P2_MMAW) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(bar_base + 96u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();                                        // barriers of both CTAs initialised, TMEM allocated in both
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // the pair's accumulators must sit at the same TMEM columns in both CTAs (one joint allocation): anything else is a bug
  if (threadIdx.x == 0 && ld_shared_cluster_u32(mapa_u32(bar_base + 96u, rank ^ 1u)) != tmem_base) __trap();
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // tile of this CTA in pair-iteration i; the second tile of an odd last pair repeats the last tile and stores nothing
  auto tile_of = [&](int pair) { return 2 * pair + (int)rank; };

  if (warp == P2_MMAW && elect_one()) {
    // ===================== weights: this CTA's share of every tap, resident for the kernel's life =====================
    // leader: X = W_hi rows 0..63, Y = W_hi rows 0..31;  peer: X = W_lo rows 0..63, Y = W_hi rows 32..63
    mbar_expect_tx(b_full, P2_W_BYTES);
    for (int t = 0; t < 9; ++t) {
      const uint32_t sb = base + t * P2_W_SLOT;
      tma_load_3d(sb, rank == 0 ? &tm_w_hi : &tm_w_lo, b_full, 0, 0, t);
      tma_load_3d(sb + P2_WX, &tm_w_hi32, b_full, 0, rank == 0 ? 0 : 32, t);
    }
    mbar_wait(b_full, 0);
    mbar_arrive_cluster(mapa_u32(w_ready, 0));
    if (rank == 0) {
      // ===================== MMA issuer (leader only) =====================
      constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);    // M 256, N 64
      constexpr uint32_t idesc2 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);  // M 256, N 128
      int acc = 0; uint32_t acc_phase = 0;
      uint32_t i = 0;
      mbar_wait_cluster(w_ready, 0);
      const bool prof = P.dbg != nullptr && cluster_id == 0;
      long long c_te = 0, c_af = 0, c_is = 0, t0 = 0, t1 = 0, t2 = 0;
      for (int pair = cluster_id; pair < n_pairs; pair += n_clusters, ++i) {
        const int w = i & 1;
        if (prof) t0 = clock64();
        mbar_wait_cluster(tempty_bar(acc), acc_phase ^ 1);
        if (prof) t1 = clock64();
        mbar_wait_cluster(a_full(w), (i >> 1) & 1);
        if (prof) t2 = clock64();
        tc_fence_after();
        const uint32_t d_main = tmem_base + (uint32_t)(acc * 128), d_cross = d_main + 64;
        uint32_t first = 1;
#pragma unroll 1
        for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const uint32_t off = (uint32_t)w * P2_WIN + (uint32_t)((ky * P2_HC + kx) * 128);
            const uint64_t a_hi = p2_desc_sbo(a_hi_base + off, P2_PITCH);
            const uint64_t a_lo = p2_desc_sbo(a_lo_base + off, P2_PITCH);
            const uint32_t sb = base + (ky * 3 + kx) * P2_W_SLOT;
            const uint64_t b_x = umma_desc_sw128(sb), b_y = umma_desc_sw128(sb + P2_WX);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adv = (uint64_t)(k * 32 >> 4);
              umma_f16_2cta(d_main, a_hi + adv, b_x + adv, idesc2, (first && k == 0) ? 0u : 1u);   // hi*[W_hi | W_lo]
              umma_f16_2cta(d_cross, a_lo + adv, b_y + adv, idesc, 1u);                            // lo*W_hi -> cross
            }
            first = 0;
          }
        }
        umma_commit_2cta(mma_done(w));
        umma_commit_2cta(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        if (prof) { c_te += t1 - t0; c_af += t2 - t1; c_is += clock64() - t2; }
      }
      if (prof) { P.dbg[4] = c_te; P.dbg[5] = c_af; P.dbg[6] = c_is; P.dbg[9] = i; }
    }
  } else if (warp >= P2_EPI0 && warp < P2_EPI0 + 8) {
    // ===================== epilogue (each CTA drains its own TMEM; eight warps, two per lane quarter, take the
    //                       16-column chunks alternately) =====================
    const int q = warp & 3;
    const int eset = (warp - P2_EPI0) >> 2;
    constexpr int NSET = 2;
    int acc = 0; uint32_t acc_phase = 0;
    const int Hp = P.H >> 1, Wp = P.W >> 1;
    const uint32_t tempty0 = mapa_u32(tempty_bar(0), 0), tempty1 = mapa_u32(tempty_bar(1), 0);
    const bool prof = P.dbg != nullptr && blockIdx.x == 0 && warp == P2_EPI0 && lane == 0;
    long long c_wait = 0, c_work = 0, t0 = 0, t1 = 0;
    for (int pair = cluster_id; pair < n_pairs; pair += n_clusters) {
      const int tile_raw = tile_of(pair);
      const bool active = tile_raw < n_tiles;
      const int tile = active ? tile_raw : n_tiles - 1;
      const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
      const bool pool = FIRST || P.pool;
      const int py = ty * (P2_TH / 2) + 2 * q + (lane >> 4), px = tx * (P2_TW / 2) + ((lane & 7) >> 1);
      const int y = ty * P2_TH + 4 * q + (lane >> 3), x = tx * P2_TW + (lane & 7);
      const bool writer = active && (pool ? (!(lane & 8) && !(lane & 1) && py < Hp && px < Wp) : (y < P.H && x < P.W));
      const size_t ppix = pool ? ((size_t)b * Hp + py) * Wp + px : ((size_t)b * P.H + y) * P.W + x;
      if (prof) t0 = clock64();
      mbar_wait(tfull_bar(acc), acc_phase);
      if (prof) t1 = clock64();
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 128);
#pragma unroll 1
      for (int n0 = eset * 16; n0 < 64; n0 += 16 * NSET) {
        uint32_t v[16], vc[16];
        tmem_ld16(t_row + n0, v);
        tmem_ld16(t_row + 64 + n0, vc);
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
      if (lane == 0) mbar_arrive_cluster(acc ? tempty1 : tempty0);      // the leader's barrier: 4 warps x 2 CTAs
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if (prof) { c_wait += t1 - t0; c_work += clock64() - t1; }
    }
    if (prof) { P.dbg[7] = c_wait; P.dbg[8] = c_work; }
  } else if (!FIRST && warp == 0 && elect_one()) {
    // ===================== halo tiles by TMA: one box {64 ch, 10 px, 18 rows} per plane =====================
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t af0 = mapa_u32(a_full(0), 0), af1 = mapa_u32(a_full(1), 0);
    uint32_t i = 0;
    for (int pair = cluster_id; pair < n_pairs; pair += n_clusters, ++i) {
      const int w = i & 1;
      const int tile = min(tile_of(pair), n_tiles - 1);
      const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
      if (i >= 2) mbar_wait(mma_done(w), ((i >> 1) - 1) & 1);
      mbar_expect_tx(a_local(w), 2u * P2_WIN);
      tma_load_4d(a_hi_base + w * P2_WIN, &tm_a_hi, a_local(w), 0, tx * P2_TW - 1, ty * P2_TH - 1, b);
      tma_load_4d(a_lo_base + w * P2_WIN, &tm_a_lo, a_local(w), 0, tx * P2_TW - 1, ty * P2_TH - 1, b);
      mbar_wait(a_local(w), (i >> 1) & 1);                    // landed here -> tell the leader
      mbar_arrive_cluster(w ? af1 : af0);
    }
  } else if (FIRST && warp < P2_NPROD) {
    // ===================== conv1a producers (6 warps per CTA): lane = halo pixel, all 64 channels in eight groups of
    // eight, weights from the constant bank (conv1_fused.cu has the same producers; here every window is a full one) =====
    const int q = warp * 32 + lane;
    const bool active = q < P2_HR * P2_HC;
    const int qc = min(q, P2_HR * P2_HC - 1);
    const int r = qc / P2_HC, c = qc - r * P2_HC;
    const int ptid = warp * 32 + lane;                       // stages patch bytes ptid and ptid + 192
    constexpr int NPT = P2_NPROD * 32;
    uint8_t* patch0 = smem_raw + P2_PATCH_OFF;
    const uint32_t af0 = mapa_u32(a_full(0), 0), af1 = mapa_u32(a_full(1), 0);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    auto patch_byte = [&](int tile, int idx) -> uint32_t {
      const int pr = idx / P2_PC, pc = idx - pr * P2_PC;
      const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
      const int gy = ty * P2_TH - 2 + pr, gx = tx * P2_TW - 2 + pc;
      const bool in = idx < P2_PR * P2_PC && gy >= 0 && gy < P.H && gx >= 0 && gx < P.W;
      return in ? (uint32_t)__ldg(P.img + ((size_t)b * P.H + gy) * P.W + gx) : 0u;
    };
    int pair = cluster_id;
    if (pair < n_pairs) {
      const int t0 = min(tile_of(pair), n_tiles - 1);
      patch0[ptid] = (uint8_t)patch_byte(t0, ptid);
      if (ptid + NPT < P2_PR * P2_PC) patch0[ptid + NPT] = (uint8_t)patch_byte(t0, ptid + NPT);
    }
    asm volatile("bar.sync 1, 192;" ::: "memory");
    uint32_t i = 0;
    const bool prof = P.dbg != nullptr && blockIdx.x == 0 && warp == 0 && lane == 0;
    long long c_w1 = 0, c_cmp = 0, t0 = 0, t1 = 0, t2 = 0;
    const long long t_begin = clock64();
    for (; pair < n_pairs; pair += n_clusters, ++i) {
      const int w = i & 1;
      const int tile = min(tile_of(pair), n_tiles - 1);
      const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y;
      const int next = pair + n_clusters;
      const int ntile = min(tile_of(next), n_tiles - 1);
      const uint32_t nb0 = next < n_pairs ? patch_byte(ntile, ptid) : 0u;
      const uint32_t nb1 = next < n_pairs ? patch_byte(ntile, ptid + NPT) : 0u;
      const uint8_t* patch = patch0 + (i & 1) * 256;
      uint8_t* patch_next = patch0 + ((i + 1) & 1) * 256;            // last read in iteration i-1, before its closing barrier
      const uint32_t off = (uint32_t)w * P2_WIN + (uint32_t)(qc * 128);
      const uint32_t ah = a_hi_base + off, al = a_lo_base + off;
      const uint32_t ph = (ah >> 7) & 7u, pl = (al >> 7) & 7u;
      const int iy = ty * P2_TH - 1 + r, ix = tx * P2_TW - 1 + c;
      const bool valid = iy >= 0 && iy < P.H && ix >= 0 && ix < P.W;
      float in[9];
      {
        const uint8_t* pp = patch + r * P2_PC + c;
#pragma unroll
        for (int t = 0; t < 9; ++t) in[t] = __fmul_rn(__uint2float_rn((uint32_t)pp[(t / 3) * P2_PC + t % 3]), P.alpha);
      }
      if (prof) t0 = clock64();
      if (i >= 2) mbar_wait(mma_done(w), ((i >> 1) - 1) & 1);
      if (prof) t1 = clock64();
      auto run = [&](auto G) {
        constexpr int C0 = decltype(G)::value * 8;
        uint64_t acc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = pk2(W1.b[C0 + 2 * j], W1.b[C0 + 2 * j + 1]);
#pragma unroll
        for (int t = 0; t < 9; ++t) {                         // taps ascending per channel: conv_first_split_kernel's fma chain
          const uint64_t vv = pk2(in[t], in[t]);
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[j] = fma2(vv, pk2(W1.w[t][C0 + 2 * j], W1.w[t][C0 + 2 * j + 1]), acc[j]);
        }
        uint32_t h[4], l[4];
#pragma unroll
        for (int j2 = 0; j2 < 4; ++j2) {
          float a0, a1;
          upk2(acc[j2], a0, a1);
          const float s0 = fmaxf(a0, 0.f), s1 = fmaxf(a1, 0.f);
          const __half2 hp = __floats2half2_rn(s0, s1);
          const float2 hf = __half22float2(hp);
          float d0, d1;
          upk2(sub2(pk2(s0, s1), pk2(hf.x, hf.y)), d0, d1);
          const __half2 lp = __floats2half2_rn(d0, d1);
          h[j2] = valid ? *reinterpret_cast<const uint32_t*>(&hp) : 0u;
          l[j2] = valid ? *reinterpret_cast<const uint32_t*>(&lp) : 0u;
        }
        if (active) {
          const uint32_t g = (uint32_t)decltype(G)::value;
          p2_st_shared_128(ah + ((g ^ ph) << 4), h[0], h[1], h[2], h[3]);
          p2_st_shared_128(al + ((g ^ pl) << 4), l[0], l[1], l[2], l[3]);
        }
      };
      run(std::integral_constant<int, 0>{}); run(std::integral_constant<int, 1>{});
      run(std::integral_constant<int, 2>{}); run(std::integral_constant<int, 3>{});
      run(std::integral_constant<int, 4>{}); run(std::integral_constant<int, 5>{});
      run(std::integral_constant<int, 6>{}); run(std::integral_constant<int, 7>{});
      if (prof) t2 = clock64();
      asm volatile("fence.proxy.async;" ::: "memory");         // generic-proxy stores -> visible to the pair's MMA (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(w ? af1 : af0);
      patch_next[ptid] = (uint8_t)nb0;                         // the next tile's patch bytes (their loads had the whole tile to land)
      if (ptid + NPT < P2_PR * P2_PC) patch_next[ptid + NPT] = (uint8_t)nb1;
      asm volatile("bar.sync 1, 192;" ::: "memory");           // patch_next written by all, this patch read by all
      if (prof) { c_w1 += t1 - t0; c_cmp += t2 - t1; }
    }
    if (prof) { P.dbg[0] = c_w1; P.dbg[1] = c_cmp; P.dbg[2] = 0; P.dbg[3] = clock64() - t_begin; }
  }
  tc_fence_before();
  cluster_sync_all();                                          // nobody leaves while the peer may still signal or read
  if (warp == P2_MMAW) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
  }
}

static osb_status launch_pair(bool first, const UmmaLayer& L, const CUtensorMap& a_hi, const CUtensorMap& a_lo, const Conv1aW& W1,
                              const PairArgs& P, cudaStream_t st, int max_ctas) {
  const int tiles = P.B * cdiv(P.W, P2_TW) * cdiv(P.H, P2_TH);
  const int pairs = (tiles + 1) / 2;
  int ctas = persistent_ctas(max_ctas);
  ctas = std::max(2, std::min(ctas & ~1, 2 * pairs));          // whole clusters
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(P2_THREADS); cfg.dynamicSmemBytes = P2_SMEM; cfg.stream = st;
  if (first) {
    OSB_SMEM_OPT_IN(conv64_pair_kernel<true>, P2_SMEM);
    OSB_CUDA(cudaLaunchKernelEx(&cfg, conv64_pair_kernel<true>, L.tm_hi, L.tm_lo, L.tm_hi32, a_hi, a_lo, W1, P));
  } else {
    OSB_SMEM_OPT_IN(conv64_pair_kernel<false>, P2_SMEM);
    OSB_CUDA(cudaLaunchKernelEx(&cfg, conv64_pair_kernel<false>, L.tm_hi, L.tm_lo, L.tm_hi32, a_hi, a_lo, W1, P));
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return OSB_OK;
}

osb_status umma_pair_first_forward(const UmmaLayer& L1b, const float* w1a_host, const float* b1a_host, const uint8_t* img, int B,
                                   int H, int W, float act_scale, __half* out_hi, __half* out_lo, float out_scale, cudaStream_t st,
                                   int max_ctas, unsigned long long* dbg) {
  OSB_REQUIRE(L1b.n_pad == 64 && L1b.cin == 64 && L1b.ks == 3, "pair kernel expects the 64 -> 64 3x3 layer");
  OSB_REQUIRE(H % 2 == 0 && W % 2 == 0, "fused max-pool needs even H and W");
  PairArgs P;
  P.img = img; P.w1a = nullptr; P.b1a = nullptr; P.bias = L1b.bias; P.out_hi = out_hi; P.out_lo = out_lo;
  P.H = H; P.W = W; P.B = B; P.alpha = (float)(1.0 / 255.0); P.pool = 1; P.dbg = dbg;
  P.act_scale = act_scale; P.inv_scale = 1.0f / (act_scale * L1b.w_scale); P.out_scale = out_scale;
  Conv1aW W1;                                             // the plane scale is a power of two: the products are exact
  for (int t = 0; t < 9; ++t)
    for (int c = 0; c < 64; ++c) W1.w[t][c] = w1a_host[t * 64 + c] * act_scale;
  for (int c = 0; c < 64; ++c) W1.b[c] = b1a_host[c] * act_scale;
  return launch_pair(true, L1b, L1b.tm_hi, L1b.tm_lo, W1, P, st, max_ctas);     // (activation maps unused)
}

// descriptors of a 64-channel activation tensor for the pair kernel: one box {64 ch, 10 px, 18 rows} per plane
osb_status umma_pair_maps(CUtensorMap* hi, CUtensorMap* lo, __half* p_hi, __half* p_lo, int B, int H, int W) {
  const uint64_t dims[4] = {64, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  const uint64_t strides[3] = {64 * 2, (uint64_t)W * 64 * 2, (uint64_t)H * W * 64 * 2};
  const uint32_t box[4] = {64, P2_HC, P2_HR, 1};
  osb_status s = umma_make_tmap(hi, p_hi, 4, dims, strides, box);
  if (s != OSB_OK) return s;
  return umma_make_tmap(lo, p_lo, 4, dims, strides, box);
}

osb_status umma_pair_conv64_forward(const UmmaLayer& L, const CUtensorMap& a_hi, const CUtensorMap& a_lo, int B, int H, int W,
                                    float act_scale, __half* out_hi, __half* out_lo, float out_scale, int pool, cudaStream_t st,
                                    int max_ctas, unsigned long long* dbg) {
  OSB_REQUIRE(L.n_pad == 64 && L.cin == 64 && L.ks == 3, "pair kernel expects a 64 -> 64 3x3 layer");
  OSB_REQUIRE(!pool || (H % 2 == 0 && W % 2 == 0), "fused max-pool needs even H and W");
  PairArgs P;
  P.img = nullptr; P.w1a = nullptr; P.b1a = nullptr; P.bias = L.bias; P.out_hi = out_hi; P.out_lo = out_lo;
  P.H = H; P.W = W; P.B = B; P.alpha = 0.f; P.pool = pool; P.dbg = dbg;
  P.act_scale = act_scale; P.inv_scale = 1.0f / (act_scale * L.w_scale); P.out_scale = out_scale;
  static const Conv1aW no_w1 = {};
  return launch_pair(false, L, a_hi, a_lo, no_w1, P, st, max_ctas);
}

}  // namespace osb
