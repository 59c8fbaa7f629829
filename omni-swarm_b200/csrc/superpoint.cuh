// superpoint.cuh -- internal SuperPoint / NetVLAD objects (shared by the C-ABI wrappers and the keyframe front-end)
#pragma once
#include "common.cuh"
#include "kernels.cuh"
#include "conv_umma.cuh"

namespace osb {

size_t sp_expected_weights();
size_t nv_expected_weights();

struct SuperPoint {
  int W = 0, H = 0, Wc = 0, Hc = 0, max_num = 0, max_batch = 0, last_batch = 0;
  float thres = 0.f;
  cudaStream_t stream = nullptr;
  // weights
  float *w1a = nullptr, *b1a = nullptr, *lut = nullptr, *pca_compT = nullptr, *pca_mean_d = nullptr;
  std::vector<float> w1a_host, b1a_host;          // conv1a [tap][64] and bias: the fused first-layers kernel takes them as a kernel parameter
  ConvLayer L[12];
  // activations / outputs (device)
  uint8_t* d_img = nullptr;
  float *actA = nullptr, *actB = nullptr, *d_logits = nullptr, *d_semi = nullptr, *d_desc = nullptr;
  KeypointScratch ks;
  int32_t* d_nk = nullptr;
  float *d_kpts = nullptr, *d_conf = nullptr, *d_out = nullptr;

  // tensor-core path (conv_umma.cu): weights as split fp16 planes, one pair of TMA descriptors per conv input
  bool use_umma = true;
  UmmaLayer UL[12];
  CUtensorMap tmA[12], tmB[12];     // [layer] -> (hi, lo) descriptors of that layer's INPUT planes
  __half *in_hi[12] = {}, *in_lo[12] = {};
  // per-layer timing (debug / bench): ev[i] is recorded after launch i of the network when `layer_prof` is set
  bool layer_prof = false;
  cudaEvent_t lev[20] = {};
  int n_lev = 0;
  void mark(cudaStream_t st) { if (layer_prof && n_lev < 20) { if (!lev[n_lev]) cudaEventCreate(&lev[n_lev]); cudaEventRecord(lev[n_lev++], st); } }

  osb_status init(const float* weights, size_t n_weights, int width, int height, float thres, int max_num,
                  const float* pca_comp, const float* pca_mean, int max_batch);
  void release();
  // keypoint extraction needs only the detector head: when a KpJob is passed, the network launches it on `kp_stream`
  // as soon as the heat map exists and runs the descriptor head beside it (on B fewer SMs); `st` re-joins before return
  struct KpJob { int32_t* nk; float* kpts; float* conf; };
  cudaStream_t kp_stream = nullptr;
  cudaEvent_t ev_semi = nullptr, ev_kp = nullptr;
  bool overlap_kp = true;
  unsigned long long* d_f1dbg = nullptr;   // cycle counters of the fused first-layers kernel (filled while layer_prof is on)
  CUtensorMap pairA[4], pairB[4]; // [2], [3]: conv2a / conv2b inputs for the CTA-pair kernel (one 18-row box per plane)
  bool pair_first = false;        // OSB_SP_PAIR=1: the first layers on CTA pairs too; =2: only conv2a / conv2b; =3: only conv2b
  bool pair_2a = true;
  bool pair64 = false;            // OSB_SP_PAIR=1: conv1a+1b, conv2a, conv2b on CTA pairs (conv64_pair.cu, tcgen05.mma.cta_group::2).
                                  // Bit-identical, measured SLOWER than the single-CTA kernels (r02: conv1 0.57 vs 0.47 ms, conv2a
                                  // 0.120 vs 0.116 ms; the pair's MMA stream ran at 134 cycles per K step against 114) -- kept as a switch
  HaloMaps halo[4];               // [2], [3]: conv2a / conv2b inputs for the halo-window kernel
  bool halo64 = false;            // OSB_SP_HALO64=1: conv2a / conv2b through conv64_halo_kernel<false> (one TMA halo window per tile:
                                  // 2.6x less activation traffic, but the 3 rows the two windows share are loaded after the previous
                                  // tile's MMAs and cost a ~1200-cycle bubble per tile: 0.118 vs 0.116 ms, r02) -- kept as a switch
  bool fuse_first = true;          // conv1a computed inside conv1b's kernel (conv1_fused.cu; OSB_SP_FUSE1=0: two kernels)
  bool fused_softmax = true;       // detector-head softmax + pixel shuffle in convPb's epilogue (OSB_SP_FUSED_SOFTMAX=0: two kernels)
  osb_status network(const uint8_t* img_dev, int B, cudaStream_t st, const KpJob* kp = nullptr);
  osb_status network_umma(const uint8_t* img_dev, int B, cudaStream_t st, const KpJob* kp);
  osb_status keypoints(int B, const KpJob& kp, cudaStream_t st);
  osb_status descriptors(int B, const KpJob& kp, float* out, cudaStream_t st);
  // network + keypoints + descriptors (what inference() is)
  osb_status forward(const uint8_t* img_dev, int B, int32_t* nk, float* kpts, float* conf, float* out, cudaStream_t st);
  osb_status postprocess(int B, int32_t* nk, float* kpts, float* conf, float* out, cudaStream_t st);
  osb_status infer_dev(const uint8_t* img_dev, int B, int32_t* nk, float* kpts, float* out, cudaStream_t st);
};

struct NetVLAD {
  int W = 0, H = 0, max_batch = 0;
  cudaStream_t stream = nullptr;
  float *w0 = nullptr, *b0 = nullptr, *lut = nullptr, *pw0_kc = nullptr, *pw0_b = nullptr;
  struct Block { float *dw = nullptr, *dwb = nullptr; ConvLayer pw; int cin = 0, cout = 0, stride = 1; } blk[7];
  ConvLayer proj, assign;
  float* centroids = nullptr;
  uint8_t* d_img = nullptr;
  float *actA = nullptr, *actB = nullptr, *d_assign = nullptr, *d_out = nullptr;
  float *d_mu = nullptr, *d_part = nullptr, *d_psum = nullptr;
  // tensor-core pointwise path: blocks 1..6 and the projection read split fp16 planes written by the depthwise kernel
  bool use_umma = true;
  UmmaLayer upw[7], uproj;
  CUtensorMap tmA[8], tmB[8];           // [block] (7 = projection) descriptors of the pointwise conv's input planes
  __half *pl_hi[8] = {}, *pl_lo[8] = {};
  __half* planes = nullptr;

  osb_status init(const float* weights, size_t n_weights, int width, int height, int max_batch);
  void release();
  osb_status infer_dev(const uint8_t* img_dev, int B, float* out_dev, cudaStream_t st);
};

}  // namespace osb
