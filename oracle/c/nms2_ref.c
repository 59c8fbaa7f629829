/* ORACLE (test infrastructure, NOT product code) -- plain-C restatement of SuperPointTensorRT::getKeyPoints + NMS2
 * (/root/reference/swarm_loop/src/superpoint_tensorrt.cpp:164-189, 237-310).
 *
 * Second, independent statement of the same loops as oracle/frontend_ref.py::nms2 (numpy); tests/test_oracle_pins.py
 * checks the two against each other and against tests/golden/postproc.npz.  Also the CPU post-processing of bench.py's
 * TRT-like baseline leg, so that leg pays C speed for this stage like the reference does.  Only tests/, smoke() and
 * bench.py's baseline legs may load it (oracle/__init__.py); the product library never links it.
 *
 * Defined where the reference is undefined: a neighbour whose flat address falls outside the H*W buffer is skipped
 * (cv::Mat::at is unchecked in release builds); std::sort ties keep raster order (stable insertion into the top list).
 * Literal otherwise: strict '>' threshold, row-major candidate order (cv::findNonZero), CV_16UC1 index plane that wraps
 * above 65535 candidates (:246,260), strict '<' suppression that may overwrite an earlier survivor (:273-276), border = 0.
 *
 * build: gcc -O2 -shared -fPIC -o libnms2_ref.so nms2_ref.c      (oracle/c/Makefile; __graft_entry__.build() runs it)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* prob [H][W] f32 -> up to max_num keypoints (x, y) as floats, ordered by descending confidence; conf_out [max_num];
 * returns the number of keypoints, or -1 when out of memory.  n_cand_out (may be NULL) receives the candidate count. */
int osb_ref_get_keypoints(const float* prob, int H, int W, float thres, int max_num, int dist_thresh, float* kpts_out,
                          float* conf_out, int* n_cand_out) {
  const long HW = (long)H * W;
  uint8_t* grid = (uint8_t*)calloc((size_t)HW, 1);
  uint16_t* inds = (uint16_t*)calloc((size_t)HW, sizeof(uint16_t));
  float* confp = (float*)calloc((size_t)HW, sizeof(float));
  int32_t* cand = (int32_t*)malloc((size_t)HW * sizeof(int32_t));
  if (!grid || !inds || !confp || !cand) { free(grid); free(inds); free(confp); free(cand); return -1; }
  /* getKeyPoints (:164-189): mask = prob > thres, findNonZero = raster order */
  long M = 0;
  for (long L = 0; L < HW; ++L)
    if (prob[L] > thres) cand[M++] = (int32_t)L;
  if (n_cand_out) *n_cand_out = (int)M;
  /* NMS2 scatter (:245-263) */
  for (long i = 0; i < M; ++i) {
    const long L = cand[i];
    grid[L] = 1;
    inds[L] = (uint16_t)(i & 0xFFFF);
    confp[L] = prob[L];
  }
  /* greedy raster-order suppression (:265-283) */
  for (long i = 0; i < M; ++i) {
    const long L = cand[i];
    if (grid[L] != 1) continue;
    const float c = confp[L];
    for (int k = -dist_thresh; k <= dist_thresh; ++k)
      for (int j = -dist_thresh; j <= dist_thresh; ++j) {
        if (j == 0 && k == 0) continue;
        const long N = L + (long)k * W + j;          /* the flat address Mat::at computes: columns wrap into the next row */
        if (N < 0 || N >= HW) continue;
        if (confp[N] < c) grid[N] = 0;
      }
    grid[L] = 2;
  }
  /* survivors in raster order (:287-302), sorted by confidence descending (:304), first max_num (:305-308).
   * Insertion into a bounded list keeps raster order among equal confidences (= a stable sort). */
  int n = 0;
  for (long L = 0; L < HW; ++L) {
    if (grid[L] != 2) continue;
    const float c = confp[L];
    if (n == max_num && !(c > conf_out[n - 1])) continue;
    int pos = n < max_num ? n : max_num - 1;
    while (pos > 0 && conf_out[pos - 1] < c) {
      conf_out[pos] = conf_out[pos - 1];
      kpts_out[2 * pos] = kpts_out[2 * (pos - 1)];
      kpts_out[2 * pos + 1] = kpts_out[2 * (pos - 1) + 1];
      --pos;
    }
    const long src = cand[inds[L]];                  /* the u16 index plane: wraps above 65535 candidates */
    conf_out[pos] = c;
    kpts_out[2 * pos] = (float)(src % W);
    kpts_out[2 * pos + 1] = (float)(src / W);
    if (n < max_num) ++n;
  }
  free(grid); free(inds); free(confp); free(cand);
  return n;
}
