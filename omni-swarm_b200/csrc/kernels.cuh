// kernels.cuh -- internal (non-ABI) launchers shared between the translation units of libomniswarm_b200.
#pragma once
#include "common.cuh"

namespace osb {

// ---- match.cu ------------------------------------------------------------------------------------------------
int db_scan_grid(int64_t n, int64_t* chunk_out);
// n_dev (optional): true row count in device memory, n is then an upper bound used to size the grid
osb_status db_search_device(const float* rows, int64_t n, const int64_t* n_dev, int dim, const float* q_dev, int nq,
                            int k, float* part_scores, int64_t* part_ids, unsigned int* done, float* scores_dev, int64_t* ids_dev,
                            cudaStream_t st);
// q/t: device tables of n_pairs pointers to [<=max_n][64] descriptor blocks
// outputs (qi/ti/dout/map_out) are [n_pairs][out_stride]
osb_status bf_match_device(int n_pairs, int max_n, int out_stride, const float* const* q, const int32_t* nq,
                           const float* const* t,
                           const int32_t* nt, float* dist_scratch, int32_t* qi, int32_t* ti, float* dout,
                           int32_t* n_out, int32_t* map_out, cudaStream_t st);

// ---- conv_ffma.cu --------------------------------------------------------------------------------------------
enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_RELU6 = 2 };

// Weights of one dense conv layer repacked for the FFMA kernel: [tap][Cin][Cout_pad] + bias[Cout_pad].
struct ConvLayer {
  int cin = 0, cout = 0, cout_pad = 0, ks = 1;
  float* w = nullptr;     // device
  float* b = nullptr;     // device
};
osb_status conv_layer_upload(ConvLayer* L, const float* w_oihw, const float* bias, int cin, int cout, int ks);
void conv_layer_free(ConvLayer* L);

// y[B][H][W][out_cstride] = act(conv_ks(x[B][H][W][Cin]) + b)   (NHWC, stride 1, "same" padding)
osb_status conv_forward(const ConvLayer& L, const float* x, float* y, int B, int H, int W, int out_cstride,
                        int act, cudaStream_t st);
// first layer: u8 image [B][H][W] -> [B][H/stride][W/stride][COUT], 3x3 pad 1, input value LUT (256 floats, device)
osb_status conv_first_forward(const float* w_tap_cout /*[9][cout]*/, const float* bias, const float* lut,
                              const uint8_t* img, float* y, int B, int H, int W, int cout, int stride, int act,
                              cudaStream_t st);
osb_status maxpool2x2_forward(const float* x, float* y, int B, int H, int W, int C, cudaStream_t st);
// depthwise 3x3 pad 1 stride s: w [9][C], NHWC
osb_status dwconv3x3_forward(const float* w_tap_c, const float* bias, const float* x, float* y, int B, int H, int W,
                             int C, int stride, int act, cudaStream_t st);

// ---- postproc.cu ---------------------------------------------------------------------------------------------
// detector head: logits [B][Hc][Wc][cstride>=65] -> semi [B][Hc*8][Wc*8]  (softmax over 65, drop dustbin, 8x8 shuffle)
osb_status sp_softmax_shuffle(const float* logits, int cstride, float* semi, int B, int Hc, int Wc, cudaStream_t st);
// descriptor head: in-place L2 normalisation over `C` channels of every cell of [cells][C]
osb_status l2norm_cells(float* x, int64_t cells, int C, cudaStream_t st);

struct KeypointScratch {   // per handle, sized for max_batch images
  uint8_t* state = nullptr;     // [B][H*W]
  uint8_t* surv = nullptr;      // [B][H*W]
  int32_t* cand = nullptr;      // [B][H*W]
  unsigned long long* skey = nullptr;  // [B][H*W]
  unsigned long long* cmask = nullptr; // [B][2][H*W] earlier / later stronger-neighbour bit masks per candidate
  bool write_surv = true;              // maintain the survivor plane (parity hook `read(3)`); off in the front-end
  int32_t* counts = nullptr;    // [B][8]: M candidates, S survivors, rounds, reserved, SM cycles of phases 1..4
  float* cnorm = nullptr;       // [B][256]
};
osb_status sp_keypoints(const float* semi, int B, int H, int W, float thres, int max_num, KeypointScratch& ks,
                        int32_t* n_kpts, float* kpts, float* conf, cudaStream_t st);
// desc_nhwc [B][Hc][Wc][256]; kpts [B][max_num][2]; out [B][max_num][64]
osb_status sp_descriptors(const float* desc_nhwc, int B, int H, int W, const int32_t* n_kpts, const float* kpts,
                          int max_num, const float* pca_compT /*[256][64]*/, const float* pca_mean, float* cnorm, float* out,
                          cudaStream_t st);
// layout helper: [B][C][h][w] -> [B][h][w][C]
osb_status nchw_to_nhwc(const float* in, float* out, int B, int C, int h, int w, cudaStream_t st);
osb_status nhwc_to_nchw(const float* in, float* out, int B, int C, int h, int w, cudaStream_t st);

// ---- geom.cu ---------------------------------------------------------------------------------------------------
// homography-RANSAC inlier masks of n_pairs correspondence sets ([n_pairs][max_n] float2 old / new points)
// stereo triangulation of the matched up/down keypoints of n_dirs directions (lift.cu; loop_cam.cpp:393-432)
osb_status stereo_lift_device(const float* kp_up, const float* kp_down, const int32_t* match, const int32_t* n_up,
                              const int32_t* n_down, int n_dirs, int max_n, const double* K, const double* pose_up,
                              const double* pose_down, double triangle_thres, int min_pts, float* pts3d, uint8_t* flag_up,
                              uint8_t* flag_down, cudaStream_t st);
osb_status homography_ransac_device(const float* src_dev, const float* dst_dev, const int32_t* n_dev, int n_pairs, int max_n,
                                    float thresh, uint32_t seed, uint8_t* mask_dev, int32_t* n_inl_dev, int32_t* winner_dev,
                                    cudaStream_t st, unsigned int* scratch /* 2 * n_pairs words, zero between launches */);

}  // namespace osb
