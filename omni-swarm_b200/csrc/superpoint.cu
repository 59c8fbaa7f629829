// superpoint.cu -- osb_superpoint: the device-resident replacement of class SuperPointTensorRT
// (swarm_loop/include/swarm_loop/superpoint_tensorrt.h:20-28, swarm_loop/src/superpoint_tensorrt.cpp:91-230).
// Network: swarm_loop/superpoint.ipynb:135-205.  Everything from the u8 image to {keypoints, 64-d descriptors}
// stays in HBM; the reference copied 1.2 MB + 4.9 MB per image back to the host and post-processed on the CPU
// (swarm_loop/src/tensorrt_generic.cpp:58-75).
#include "superpoint.cuh"
#include <stdlib.h>

namespace osb {

static const int SP_CIN[12] = {1, 64, 64, 64, 64, 128, 128, 128, 128, 256, 128, 256};
static const int SP_COUT[12] = {64, 64, 64, 64, 128, 128, 128, 128, 256, 65, 256, 256};
static const int SP_KS[12] = {3, 3, 3, 3, 3, 3, 3, 3, 3, 1, 3, 1};
// power-of-two scales of the split-fp16 planes (exact): activations (post-ReLU, O(1)) x 16, weights (O(0.05)) x 1024
constexpr float SP_ACT_SCALE = 16.f;
constexpr float SP_W_SCALE = 1024.f;

size_t sp_expected_weights() {
  size_t n = 0;
  for (int i = 0; i < 12; ++i) n += (size_t)SP_COUT[i] * SP_CIN[i] * SP_KS[i] * SP_KS[i] + SP_COUT[i];
  return n;
}

osb_status SuperPoint::init(const float* weights, size_t n_weights, int width, int height, float thres_, int max_num_,
                            const float* pca_comp, const float* pca_mean, int max_batch_) {
  OSB_REQUIRE(weights && pca_comp && pca_mean, "null weights / pca");
  OSB_REQUIRE(n_weights == sp_expected_weights(), "weight blob has the wrong length (expected 1300865 floats)");
  OSB_REQUIRE(width > 0 && height > 0 && width % 8 == 0 && height % 8 == 0, "width/height must be multiples of 8");
  OSB_REQUIRE(max_num_ > 0 && max_num_ <= 8192 && max_batch_ > 0, "bad max_num / max_batch");
  W = width; H = height; thres = thres_; max_num = max_num_; max_batch = max_batch_;
  Hc = H / 8; Wc = W / 8;
  OSB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  {
    int least = 0, greatest = 0;
    OSB_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    OSB_CUDA(cudaStreamCreateWithPriority(&kp_stream, cudaStreamNonBlocking, greatest));
  }
  OSB_CUDA(cudaEventCreateWithFlags(&ev_semi, cudaEventDisableTiming));
  OSB_CUDA(cudaEventCreateWithFlags(&ev_kp, cudaEventDisableTiming));
  if (const char* e = getenv("OSB_SP_OVERLAP")) overlap_kp = atoi(e) != 0;
  if (const char* e = getenv("OSB_SP_FUSED_SOFTMAX")) fused_softmax = atoi(e) != 0;
  if (const char* e = getenv("OSB_SP_FUSE1")) fuse_first = atoi(e) != 0;
  if (const char* e = getenv("OSB_SP_HALO64")) halo64 = atoi(e) != 0;
  if (const char* e = getenv("OSB_SP_PAIR")) { pair64 = atoi(e) != 0; pair_first = atoi(e) == 1; pair_2a = atoi(e) != 3; }
  // ---- weights ----
  const float* p = weights;
  {
    // conv1a: [64][1][3][3] -> [tap][64]
    std::vector<float> w9(9 * 64);
    for (int o = 0; o < 64; ++o)
      for (int t = 0; t < 9; ++t) w9[t * 64 + o] = p[o * 9 + t];
    OSB_CUDA(cudaMalloc(&w1a, 9 * 64 * sizeof(float)));
    OSB_CUDA(cudaMalloc(&b1a, 64 * sizeof(float)));
    OSB_CUDA(cudaMemcpy(w1a, w9.data(), 9 * 64 * sizeof(float), cudaMemcpyHostToDevice));
    OSB_CUDA(cudaMemcpy(b1a, p + 64 * 9, 64 * sizeof(float), cudaMemcpyHostToDevice));
    w1a_host = w9;
    b1a_host.assign(p + 64 * 9, p + 64 * 9 + 64);
    p += 64 * 9 + 64;
  }
  {
    // OSB_SP_CONV=ffma selects the fp32 CUDA-core convolutions (debug / A-B parity); default = tcgen05 path
    const char* e = getenv("OSB_SP_CONV");
    use_umma = !(e && strcmp(e, "ffma") == 0);
  }
  for (int i = 1; i < 12; ++i) {
    const size_t nw = (size_t)SP_COUT[i] * SP_CIN[i] * SP_KS[i] * SP_KS[i];
    osb_status s = conv_layer_upload(&L[i], p, p + nw, SP_CIN[i], SP_COUT[i], SP_KS[i]);
    if (s != OSB_OK) return s;
    if (use_umma) {
      s = umma_layer_upload(&UL[i], p, p + nw, SP_CIN[i], SP_COUT[i], SP_KS[i], SP_W_SCALE);
      if (s != OSB_OK) return s;
    }
    p += nw + SP_COUT[i];
  }
  {
    // u8 -> f32 * (1/255): cv::Mat::convertTo(CV_32F, 1/255.0) computes (float)v * (float)alpha
    // (superpoint_tensorrt.cpp:127)
    std::vector<float> l(256);
    const float alpha = (float)(1.0 / 255.0);
    for (int v = 0; v < 256; ++v) l[v] = (float)v * alpha;
    OSB_CUDA(cudaMalloc(&lut, 256 * sizeof(float)));
    OSB_CUDA(cudaMemcpy(lut, l.data(), 256 * sizeof(float), cudaMemcpyHostToDevice));
  }
  {
    std::vector<float> ct(256 * 64);
    for (int o = 0; o < 64; ++o)
      for (int c = 0; c < 256; ++c) ct[c * 64 + o] = pca_comp[o * 256 + c];
    OSB_CUDA(cudaMalloc(&pca_compT, 256 * 64 * sizeof(float)));
    OSB_CUDA(cudaMalloc(&pca_mean_d, 256 * sizeof(float)));
    OSB_CUDA(cudaMemcpy(pca_compT, ct.data(), 256 * 64 * sizeof(float), cudaMemcpyHostToDevice));
    OSB_CUDA(cudaMemcpy(pca_mean_d, pca_mean, 256 * sizeof(float), cudaMemcpyHostToDevice));
  }
  // ---- activations ----
  const size_t B = max_batch, HW = (size_t)H * W;
  OSB_CUDA(cudaMalloc(&d_img, B * HW));
  OSB_CUDA(cudaMalloc(&actA, B * HW * 64 * sizeof(float)));
  OSB_CUDA(cudaMalloc(&actB, B * HW * 64 * sizeof(float)));
  OSB_CUDA(cudaMalloc(&d_logits, B * Hc * Wc * 80 * sizeof(float)));
  if (use_umma) {
    // input geometry of every conv layer: which ping-pong buffer it reads and its [H][W][C]
    // (layer order: 1 conv1b 2 conv2a 3 conv2b 4 conv3a 5 conv3b 6 conv4a 7 conv4b 8 convPa 9 convPb 10 convDa 11 convDb)
    // with the pools fused into the conv epilogues the layers simply alternate between the two buffers;
    // conv4b's output x (layer 8 input) stays in B while convPa writes A, so convDa (10) reads B, convDb (11) reads A
    const int in_buf[12] = {-1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0};      // 0 = actA, 1 = actB
    const int in_div[12] = {0, 1, 2, 2, 4, 4, 8, 8, 8, 8, 8, 8};
    for (int i = 1; i < 12; ++i) {
      const int h = H / in_div[i], w = W / in_div[i], c = SP_CIN[i];
      __half* base = reinterpret_cast<__half*>(in_buf[i] == 0 ? actA : actB);
      in_hi[i] = base;
      in_lo[i] = base + (size_t)max_batch * h * w * c;
      osb_status s = umma_act_maps(&tmA[i], &tmB[i], in_hi[i], in_lo[i], max_batch, h, w, c, SP_KS[i]);
      if (s != OSB_OK) return s;
      if (i == 2 || i == 3) {
        s = umma_halo_maps(&halo[i], in_hi[i], in_lo[i], max_batch, h, w);
        if (s != OSB_OK) return s;
        s = umma_pair_maps(&pairA[i], &pairB[i], in_hi[i], in_lo[i], max_batch, h, w);
        if (s != OSB_OK) return s;
      }
    }
  }
  OSB_CUDA(cudaMalloc(&d_f1dbg, 16 * sizeof(unsigned long long)));
  OSB_CUDA(cudaMemset(d_f1dbg, 0, 16 * sizeof(unsigned long long)));
  OSB_CUDA(cudaMalloc(&d_semi, B * HW * sizeof(float)));
  OSB_CUDA(cudaMalloc(&d_desc, B * Hc * Wc * 256 * sizeof(float)));
  OSB_CUDA(cudaMalloc(&ks.state, B * HW));
  OSB_CUDA(cudaMalloc(&ks.surv, B * HW));
  OSB_CUDA(cudaMalloc(&ks.cand, B * HW * sizeof(int32_t)));
  OSB_CUDA(cudaMalloc(&ks.skey, B * HW * sizeof(unsigned long long)));
  OSB_CUDA(cudaMalloc(&ks.cmask, B * 2 * HW * sizeof(unsigned long long)));
  OSB_CUDA(cudaMalloc(&ks.counts, B * 8 * sizeof(int32_t)));
  OSB_CUDA(cudaMalloc(&ks.cnorm, B * 256 * sizeof(float)));
  OSB_CUDA(cudaMalloc(&d_nk, B * sizeof(int32_t)));
  OSB_CUDA(cudaMalloc(&d_kpts, B * max_num * 2 * sizeof(float)));
  OSB_CUDA(cudaMalloc(&d_conf, B * max_num * sizeof(float)));
  OSB_CUDA(cudaMalloc(&d_out, B * max_num * 64 * sizeof(float)));
  OSB_CUDA(cudaMemset(d_nk, 0, B * sizeof(int32_t)));
  OSB_CUDA(cudaMemset(ks.surv, 0, B * HW));
  return OSB_OK;
}

void SuperPoint::release() {
  cudaFree(w1a); cudaFree(b1a); cudaFree(lut); cudaFree(pca_compT); cudaFree(pca_mean_d);
  for (int i = 1; i < 12; ++i) { conv_layer_free(&L[i]); umma_layer_free(&UL[i]); }
  for (int i = 0; i < 20; ++i) if (lev[i]) cudaEventDestroy(lev[i]);
  cudaFree(d_img); cudaFree(actA); cudaFree(actB); cudaFree(d_logits); cudaFree(d_semi); cudaFree(d_desc);
  cudaFree(ks.state); cudaFree(ks.surv); cudaFree(ks.cand); cudaFree(ks.skey); cudaFree(ks.cmask); cudaFree(ks.counts); cudaFree(ks.cnorm);
  cudaFree(d_nk); cudaFree(d_kpts); cudaFree(d_conf); cudaFree(d_out); cudaFree(d_f1dbg);
  if (stream) cudaStreamDestroy(stream);
  if (kp_stream) cudaStreamDestroy(kp_stream);
  if (ev_semi) cudaEventDestroy(ev_semi);
  if (ev_kp) cudaEventDestroy(ev_kp);
}

// tensor-core network: every activation is a pair of fp16 planes (hi, lo) scaled by SP_ACT_SCALE; the planes of a
// layer's output live in the ping-pong buffer the next layer's TMA descriptors point at.
osb_status SuperPoint::network_umma(const uint8_t* img_dev, int B, cudaStream_t st, const KpJob* kp) {
  osb_status s;
#define RUN(x) do { s = (x); if (s != OSB_OK) return s; } while (0)
  const float SA = SP_ACT_SCALE;
  n_lev = 0;
  mark(st);
  // conv layer i at resolution h x w; its output (optionally 2x2 max-pooled in the epilogue) becomes the input planes
  // of layer `out_layer`
  auto conv = [&](int i, int h, int w, int out_layer, int pool) {
    return umma_conv_forward(UL[i], tmA[i], tmB[i], B, h, w, SA, in_hi[out_layer], in_lo[out_layer], nullptr,
                             SP_COUT[i], SP_COUT[i], SA, 1, pool, st);
  };
  // cycle counters of the 64 -> 64 kernels (osb_superpoint_read what = 5): OSB_F1_DEBUG=<1|2|3> selects conv1 / conv2a / conv2b;
  // off by default because the clock reads cost a few percent of the kernel
  static const int dbg_layer = [] { const char* e = getenv("OSB_F1_DEBUG"); return e ? atoi(e) : 0; }();
  if (fuse_first && pair64 && pair_first) {
    mark(st);
    RUN(umma_pair_first_forward(UL[1], w1a_host.data(), b1a_host.data(), img_dev, B, H, W, SA, in_hi[2], in_lo[2], SA, st, 0,
                                (layer_prof && dbg_layer == 1) ? d_f1dbg : nullptr));     // conv1a+conv1b+pool -> B
    mark(st);
  } else if (fuse_first) {
    mark(st);                                                                             // (conv1a has no launch of its own)
    RUN(umma_conv1_fused_forward(UL[1], w1a_host.data(), b1a_host.data(), img_dev, B, H, W, SA, in_hi[2], in_lo[2], SA, st, 0,
                                 (layer_prof && dbg_layer == 1) ? d_f1dbg : nullptr));   // conv1a+conv1b+pool -> B
    mark(st);
  } else {
    RUN(umma_first_forward(w1a, b1a, lut, img_dev, in_hi[1], in_lo[1], B, H, W, SA, st)); // conv1a            -> A
    mark(st);
    RUN(conv(1, H, W, 2, 1));                                                             // conv1b + pool     -> B
    mark(st);
  }
  if (pair64) {
    if (pair_2a)
      RUN(umma_pair_conv64_forward(UL[2], pairA[2], pairB[2], B, H / 2, W / 2, SA, in_hi[3], in_lo[3], SA, 0, st, 0,
                                   (layer_prof && dbg_layer == 2) ? d_f1dbg : nullptr));                // conv2a   -> A
    else
      RUN(conv(2, H / 2, W / 2, 3, 0));
    mark(st);
    RUN(umma_pair_conv64_forward(UL[3], pairA[3], pairB[3], B, H / 2, W / 2, SA, in_hi[4], in_lo[4], SA, 1, st, 0,
                                 (layer_prof && dbg_layer == 3) ? d_f1dbg : nullptr));                  // conv2b + pool -> B
    mark(st);
  } else if (halo64) {
    RUN(umma_conv64_halo_forward(UL[2], halo[2], B, H / 2, W / 2, SA, in_hi[3], in_lo[3], SA, 0, st, 0,
                                 (layer_prof && dbg_layer == 2) ? d_f1dbg : nullptr));                  // conv2a   -> A
    mark(st);
    RUN(umma_conv64_halo_forward(UL[3], halo[3], B, H / 2, W / 2, SA, in_hi[4], in_lo[4], SA, 1, st, 0,
                                 (layer_prof && dbg_layer == 3) ? d_f1dbg : nullptr));                  // conv2b + pool -> B
    mark(st);
  } else {
    RUN(conv(2, H / 2, W / 2, 3, 0));                                                     // conv2a            -> A
    mark(st);
    RUN(conv(3, H / 2, W / 2, 4, 1));                                                     // conv2b + pool     -> B
    mark(st);
  }
  RUN(conv(4, H / 4, W / 4, 5, 0));                                                       // conv3a            -> A
  mark(st);
  RUN(conv(5, H / 4, W / 4, 6, 1));                                                       // conv3b + pool     -> B
  mark(st);
  RUN(conv(6, Hc, Wc, 7, 0));                                                             // conv4a            -> A
  mark(st);
  RUN(conv(7, Hc, Wc, 8, 0));                                                             // conv4b            -> B (x)
  mark(st);
  RUN(conv(8, Hc, Wc, 9, 0));                                                             // convPa            -> A
  mark(st);
  if (fused_softmax) {
    RUN(umma_conv_softmax_forward(UL[9], tmA[9], tmB[9], B, Hc, Wc, SA, d_semi, st));      // convPb + softmax + pixel shuffle
    mark(st);
  } else {
    RUN(umma_conv_forward(UL[9], tmA[9], tmB[9], B, Hc, Wc, SA, nullptr, nullptr, d_logits, 80, 80, 1.f, 0, 0, st));   // convPb
    mark(st);
    RUN(sp_softmax_shuffle(d_logits, 80, d_semi, B, Hc, Wc, st));
  }
  // the keypoint kernel (one CTA per image, latency-bound) runs beside the descriptor head, which leaves it B SMs
  const bool fork = kp && overlap_kp && !layer_prof && kp_stream;
  int head_ctas = 0;
  if (fork) {
    OSB_CUDA(cudaEventRecord(ev_semi, st));
    OSB_CUDA(cudaStreamWaitEvent(kp_stream, ev_semi, 0));
    RUN(keypoints(B, *kp, kp_stream));
    OSB_CUDA(cudaEventRecord(ev_kp, kp_stream));
    head_ctas = std::max(1, persistent_ctas() - B);
  }
  RUN(umma_conv_forward(UL[10], tmA[10], tmB[10], B, Hc, Wc, SA, in_hi[11], in_lo[11], nullptr, SP_COUT[10], SP_COUT[10],
                        SA, 1, 0, st, head_ctas));                                        // convDa (reads B)  -> A
  mark(st);
  RUN(umma_conv_forward(UL[11], tmA[11], tmB[11], B, Hc, Wc, SA, nullptr, nullptr, d_desc, 256, 256, 1.f, 0, 0, st,
                        head_ctas));                                                      // convDb
  mark(st);
  RUN(l2norm_cells(d_desc, (int64_t)B * Hc * Wc, 256, st));
  if (fork) OSB_CUDA(cudaStreamWaitEvent(st, ev_kp, 0));
  else if (kp) RUN(keypoints(B, *kp, st));
#undef RUN
  return OSB_OK;
}

// the network: u8 images (device) -> d_semi, d_desc
osb_status SuperPoint::network(const uint8_t* img_dev, int B, cudaStream_t st, const KpJob* kp) {
  if (use_umma) return network_umma(img_dev, B, st, kp);
  osb_status s;
#define RUN(x) do { s = (x); if (s != OSB_OK) return s; } while (0)
  RUN(conv_first_forward(w1a, b1a, lut, img_dev, actA, B, H, W, 64, 1, ACT_RELU, st));      // conv1a
  RUN(conv_forward(L[1], actA, actB, B, H, W, 64, ACT_RELU, st));                            // conv1b
  RUN(maxpool2x2_forward(actB, actA, B, H, W, 64, st));
  RUN(conv_forward(L[2], actA, actB, B, H / 2, W / 2, 64, ACT_RELU, st));                    // conv2a
  RUN(conv_forward(L[3], actB, actA, B, H / 2, W / 2, 64, ACT_RELU, st));                    // conv2b
  RUN(maxpool2x2_forward(actA, actB, B, H / 2, W / 2, 64, st));
  RUN(conv_forward(L[4], actB, actA, B, H / 4, W / 4, 128, ACT_RELU, st));                   // conv3a
  RUN(conv_forward(L[5], actA, actB, B, H / 4, W / 4, 128, ACT_RELU, st));                   // conv3b
  RUN(maxpool2x2_forward(actB, actA, B, H / 4, W / 4, 128, st));
  RUN(conv_forward(L[6], actA, actB, B, Hc, Wc, 128, ACT_RELU, st));                         // conv4a
  RUN(conv_forward(L[7], actB, actA, B, Hc, Wc, 128, ACT_RELU, st));                         // conv4b
  RUN(conv_forward(L[8], actA, actB, B, Hc, Wc, 256, ACT_RELU, st));                         // convPa
  RUN(conv_forward(L[9], actB, d_logits, B, Hc, Wc, 72, ACT_NONE, st));                      // convPb (65 -> stride 72)
  RUN(conv_forward(L[10], actA, actB, B, Hc, Wc, 256, ACT_RELU, st));                        // convDa
  RUN(conv_forward(L[11], actB, d_desc, B, Hc, Wc, 256, ACT_NONE, st));                      // convDb
  RUN(l2norm_cells(d_desc, (int64_t)B * Hc * Wc, 256, st));
  RUN(sp_softmax_shuffle(d_logits, 72, d_semi, B, Hc, Wc, st));
  if (kp) RUN(keypoints(B, *kp, st));
#undef RUN
  return OSB_OK;
}

osb_status SuperPoint::keypoints(int B, const KpJob& kp, cudaStream_t st) {
  return sp_keypoints(d_semi, B, H, W, thres, max_num, ks, kp.nk, kp.kpts, kp.conf, st);
}

osb_status SuperPoint::descriptors(int B, const KpJob& kp, float* out, cudaStream_t st) {
  return sp_descriptors(d_desc, B, H, W, kp.nk, kp.kpts, max_num, pca_compT, pca_mean_d, ks.cnorm, out, st);
}

osb_status SuperPoint::postprocess(int B, int32_t* nk, float* kpts, float* conf, float* out, cudaStream_t st) {
  const KpJob kp{nk, kpts, conf};
  osb_status s = keypoints(B, kp, st);
  if (s != OSB_OK) return s;
  return descriptors(B, kp, out, st);
}

osb_status SuperPoint::forward(const uint8_t* img_dev, int B, int32_t* nk, float* kpts, float* conf, float* out,
                               cudaStream_t st) {
  OSB_REQUIRE(B > 0 && B <= max_batch, "batch out of range");
  const KpJob kp{nk, kpts, conf};
  osb_status s = network(img_dev, B, st, &kp);
  if (s != OSB_OK) return s;
  last_batch = B;
  return descriptors(B, kp, out, st);
}

osb_status SuperPoint::infer_dev(const uint8_t* img_dev, int B, int32_t* nk, float* kpts, float* out, cudaStream_t st) {
  return forward(img_dev, B, nk, kpts, d_conf, out, st);
}

}  // namespace osb

using namespace osb;

struct osb_superpoint {
  int device = 0;
  SuperPoint sp;
  std::mutex mu;
};

extern "C" osb_status osb_superpoint_create(osb_superpoint** out, const float* weights, size_t n_weights, int width,
                                            int height, float thres, int max_num, const float* pca_comp,
                                            const float* pca_mean, int max_batch) {
  OSB_REQUIRE(out != nullptr, "null out");
  osb_status s = require_device();
  if (s != OSB_OK) return s;
  osb_superpoint* h = new osb_superpoint();
  h->device = current_device();
  s = h->sp.init(weights, n_weights, width, height, thres, max_num, pca_comp, pca_mean, max_batch);
  if (s != OSB_OK) { h->sp.release(); delete h; return s; }
  *out = h;
  return OSB_OK;
}

extern "C" osb_status osb_superpoint_destroy(osb_superpoint* h) {
  if (!h) return OSB_OK;
  h->sp.release();
  delete h;
  return OSB_OK;
}

static osb_status sp_copy_out(SuperPoint& sp, int B, int32_t* n_kpts, float* kpts, float* desc, cudaStream_t st) {
  OSB_CUDA(cudaMemcpyAsync(n_kpts, sp.d_nk, B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  OSB_CUDA(cudaMemcpyAsync(kpts, sp.d_kpts, (size_t)B * sp.max_num * 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
  OSB_CUDA(cudaMemcpyAsync(desc, sp.d_out, (size_t)B * sp.max_num * 64 * sizeof(float), cudaMemcpyDeviceToHost, st));
  OSB_CUDA(cudaStreamSynchronize(st));
  return OSB_OK;
}

extern "C" osb_status osb_superpoint_infer(osb_superpoint* h, const uint8_t* images, int batch, int32_t* n_kpts,
                                           float* kpts, float* desc) {
  OSB_REQUIRE(h && images && n_kpts && kpts && desc, "null argument");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  SuperPoint& sp = h->sp;
  OSB_REQUIRE(batch > 0 && batch <= sp.max_batch, "batch out of range");
  cudaStream_t st = sp.stream;
  OSB_CUDA(cudaMemcpyAsync(sp.d_img, images, (size_t)batch * sp.H * sp.W, cudaMemcpyHostToDevice, st));
  osb_status s = sp.infer_dev(sp.d_img, batch, sp.d_nk, sp.d_kpts, sp.d_out, st);
  if (s != OSB_OK) return s;
  return sp_copy_out(sp, batch, n_kpts, kpts, desc, st);
}

extern "C" osb_status osb_superpoint_infer_dev(osb_superpoint* h, const uint8_t* images_dev, int batch,
                                               int32_t* n_kpts_dev, float* kpts_dev, float* desc_dev, void* stream) {
  OSB_REQUIRE(h && images_dev && n_kpts_dev && kpts_dev && desc_dev, "null argument");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  return h->sp.infer_dev(images_dev, batch, n_kpts_dev, kpts_dev, desc_dev, (cudaStream_t)stream);
}

extern "C" osb_status osb_superpoint_postprocess(osb_superpoint* h, const float* semi, const float* desc_nchw,
                                                 int batch, int32_t* n_kpts, float* kpts, float* desc) {
  OSB_REQUIRE(h && semi && desc_nchw && n_kpts && kpts && desc, "null argument");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  SuperPoint& sp = h->sp;
  OSB_REQUIRE(batch > 0 && batch <= sp.max_batch, "batch out of range");
  cudaStream_t st = sp.stream;
  const size_t HW = (size_t)sp.H * sp.W, dn = (size_t)256 * sp.Hc * sp.Wc;
  OSB_CUDA(cudaMemcpyAsync(sp.d_semi, semi, batch * HW * sizeof(float), cudaMemcpyHostToDevice, st));
  OSB_CUDA(cudaMemcpyAsync(sp.actA, desc_nchw, batch * dn * sizeof(float), cudaMemcpyHostToDevice, st));
  osb_status s = nchw_to_nhwc(sp.actA, sp.d_desc, batch, 256, sp.Hc, sp.Wc, st);
  if (s != OSB_OK) return s;
  sp.last_batch = batch;
  s = sp.postprocess(batch, sp.d_nk, sp.d_kpts, sp.d_conf, sp.d_out, st);
  if (s != OSB_OK) return s;
  return sp_copy_out(sp, batch, n_kpts, kpts, desc, st);
}

extern "C" osb_status osb_superpoint_set_profiling(osb_superpoint* h, int enable) {
  OSB_REQUIRE(h != nullptr, "null handle");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  h->sp.layer_prof = enable != 0;
  h->sp.n_lev = 0;
  return OSB_OK;
}

extern "C" osb_status osb_superpoint_layer_ms(osb_superpoint* h, float* ms, int n) {
  OSB_REQUIRE(h != nullptr && ms != nullptr && n >= 12, "need room for 12 layer times");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  SuperPoint& sp = h->sp;
  OSB_CUDA(cudaStreamSynchronize(sp.stream));
  for (int i = 0; i < n; ++i) ms[i] = 0.f;
  for (int i = 0; i + 1 < sp.n_lev && i < n; ++i) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, sp.lev[i], sp.lev[i + 1]) == cudaSuccess) ms[i] = t; else cudaGetLastError();
  }
  return OSB_OK;
}

extern "C" osb_status osb_superpoint_read(osb_superpoint* h, int what, int image, float* out, size_t n_floats) {
  OSB_REQUIRE(h && out, "null argument");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  SuperPoint& sp = h->sp;
  OSB_REQUIRE(image >= 0 && image < sp.max_batch, "image index out of range");
  cudaStream_t st = sp.stream;
  const size_t HW = (size_t)sp.H * sp.W, dn = (size_t)256 * sp.Hc * sp.Wc;
  if (what == 0) {
    OSB_REQUIRE(n_floats == HW, "semi needs H*W floats");
    OSB_CUDA(cudaMemcpyAsync(out, sp.d_semi + image * HW, HW * sizeof(float), cudaMemcpyDeviceToHost, st));
  } else if (what == 1) {
    OSB_REQUIRE(n_floats == dn, "desc needs 256*H/8*W/8 floats");
    osb_status s = nhwc_to_nchw(sp.d_desc + image * dn, sp.actB, 1, 256, sp.Hc, sp.Wc, st);
    if (s != OSB_OK) return s;
    OSB_CUDA(cudaMemcpyAsync(out, sp.actB, dn * sizeof(float), cudaMemcpyDeviceToHost, st));
  } else if (what == 2) {
    OSB_REQUIRE(n_floats == (size_t)sp.max_num, "conf needs max_num floats");
    OSB_CUDA(cudaMemcpyAsync(out, sp.d_conf + (size_t)image * sp.max_num, sp.max_num * sizeof(float),
                             cudaMemcpyDeviceToHost, st));
  } else if (what == 3) {
    OSB_REQUIRE(n_floats == HW, "survivor plane needs H*W floats");
    std::vector<uint8_t> tmp(HW);
    OSB_CUDA(cudaMemcpyAsync(tmp.data(), sp.ks.surv + image * HW, HW, cudaMemcpyDeviceToHost, st));
    OSB_CUDA(cudaStreamSynchronize(st));
    for (size_t i = 0; i < HW; ++i) out[i] = (float)tmp[i];
    return OSB_OK;
  } else if (what == 4) {
    OSB_REQUIRE(n_floats == 8, "counts needs 8 floats");
    int32_t c[8];
    OSB_CUDA(cudaMemcpyAsync(c, sp.ks.counts + image * 8, sizeof(c), cudaMemcpyDeviceToHost, st));
    OSB_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < 8; ++i) out[i] = (float)c[i];
    return OSB_OK;
  } else if (what == 5) {
    OSB_REQUIRE(n_floats == 16, "fused-kernel counters need 16 floats");
    unsigned long long c[16];
    OSB_CUDA(cudaMemcpyAsync(c, sp.d_f1dbg, sizeof(c), cudaMemcpyDeviceToHost, st));
    OSB_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < 16; ++i) out[i] = (float)c[i];
    return OSB_OK;
  } else {
    set_error("osb_superpoint_read", "unknown `what`");
    return OSB_ERR_INVALID;
  }
  OSB_CUDA(cudaStreamSynchronize(st));
  return OSB_OK;
}
