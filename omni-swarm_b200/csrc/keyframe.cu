// keyframe.cu -- osb_frontend: the per-keyframe pipeline kept resident on the GPU.
//
//   extract  = LoopCam::on_flattened_images -> generate_stereo_image_descriptor for every direction
//              (swarm_loop/src/loop_cam.cpp:178-229, 341-523): SuperPoint on the up and down image, NetVLAD on the
//              up image (extractor_img_desc_deepnet :524-585, incl. the STEREO_FISHEYE bottom-quarter blanking
//              :535-538), stereo cross-check match up<->down (match_HFNet_local_features :141-174).
//   ingest   = LoopDetector::add_to_database (swarm_loop/src/loop_detector.cpp:150-173).
//   query    = query_fisheyeframe_from_database + query_from_database (:176-287) and the
//              per-direction matcher of compute_correspond_features (:431-470, :539-567).
//
// Row counters and the image-id -> (frame, direction) maps live in device memory, so a keyframe is processed with
// a single host synchronisation at the very end (the reference synchronised after every engine call,
// swarm_loop/src/tensorrt_generic.cpp:73).
#include "superpoint.cuh"
#include <atomic>

namespace osb {

struct DbDev {                 // device-resident database (one for own keyframes, one for remote ones)
  int64_t ntotal;              // rows (faiss ntotal)
  int nframes;
  int overflow;                // set when a row could not be added
};

// Row-count feedback of the ingest kernel, written into mapped pinned host memory (no copy, no synchronisation): the host's
// bounds are conservative (every ingested record is charged to BOTH stores because its drone_id is only known on the
// device); the last ingest that has actually run reports the true counts together with how much had been charged when it
// was enqueued, so bound = count + what was charged since.  seq_begin / seq_end make a torn read detectable.
struct FeFeedback {
  volatile long long seq_begin;
  volatile long long n_local, n_remote, charged;
  volatile long long seq_end;
};

struct DbStore {
  int64_t cap = 0;
  int64_t upper = 0;           // host-side upper bound of ntotal (exact after every synchronising call)
  DbDev* dev = nullptr;
  float* rows = nullptr;       // [cap][4096]
  float* ldesc = nullptr;      // [cap][max_num][64]
  float* kpts = nullptr;       // [cap][max_num][2]   landmarks_2d of the row (geometric filter)
  int32_t* smatch = nullptr;   // [cap][max_num]      landmarks_flag of the row (non-zero = the landmark has a 3-D point)
  int32_t* nk = nullptr;       // [cap]
  int32_t* row_frame = nullptr;// [cap]
  int32_t* row_dir = nullptr;  // [cap]
  int32_t* frame_rows = nullptr;  // [cap][4]
  int32_t* frame_msg = nullptr;   // [cap]   msg_id of the keyframe (imgid2fisheye -> fisheyeframe_database key)
  int32_t* frame_drone = nullptr; // [cap]   drone_id of the keyframe
  float* part_scores = nullptr;
  int64_t* part_ids = nullptr;
  unsigned int* done = nullptr;
  float* top_scores = nullptr;    // [KMAX]
  int64_t* top_ids = nullptr;     // [KMAX]
};

constexpr int FE_KMAX = 32;

// blank the bottom quarter of every image (loop_cam.cpp:535-538)
__global__ void fe_blank_kernel(uint8_t* __restrict__ img, int H, int W, int n_img) {
  const int rows0 = H * 3 / 4, nrow = H / 4;
  const size_t per = (size_t)nrow * W;
  const size_t total = per * n_img;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t im = i / per, off = i % per;
    img[im * (size_t)H * W + (size_t)rows0 * W + off] = 0;
  }
}

// assemble the keyframe record from the SuperPoint outputs (batch order: up[0..n_dirs), down[0..n_dirs))
__global__ void fe_pack_kernel(osb_keyframe_record* __restrict__ rec, int drone_id, int msg_id, int n_dirs, int max_num,
                               const int32_t* __restrict__ nk, const float* __restrict__ kpts,
                               const float* __restrict__ desc, const int32_t* __restrict__ stereo_map,
                               int accept_min_3d_pts, const float* __restrict__ l3d /*[n_dirs][max_num][3] or null*/,
                               const uint8_t* __restrict__ lflag /*[n_dirs][max_num] or null*/) {
  const int d = blockIdx.x;
  const int tid = threadIdx.x;
  if (d >= n_dirs) {          // unused directions: zero counts
    if (tid == 0 && d < OSB_MAX_DIRS) { rec->n_kpts[d] = 0; rec->n_kpts_down[d] = 0; }
    return;
  }
  const int n = nk[d];
  if (d == 0 && tid == 0) { rec->drone_id = drone_id; rec->msg_id = msg_id; rec->n_dirs = n_dirs; rec->reserved = 0; }
  if (tid == 0) { rec->n_kpts[d] = n; rec->n_kpts_down[d] = nk[n_dirs + d]; }
  for (int i = tid; i < OSB_MAX_KPTS * OSB_FEATURE_DESC_SIZE; i += blockDim.x) {
    const int r = i / OSB_FEATURE_DESC_SIZE;
    rec->local_desc[d][r][i % OSB_FEATURE_DESC_SIZE] =
        (r < n) ? desc[((size_t)d * max_num + r) * OSB_FEATURE_DESC_SIZE + (i % OSB_FEATURE_DESC_SIZE)] : 0.f;
  }
  for (int i = tid; i < OSB_MAX_KPTS; i += blockDim.x) {
    const bool ok = i < n;
    rec->kpts[d][i][0] = ok ? kpts[((size_t)d * max_num + i) * 2] : 0.f;
    rec->kpts[d][i][1] = ok ? kpts[((size_t)d * max_num + i) * 2 + 1] : 0.f;
    // the stereo match is skipped when landmarks_2d.size() <= ACCEPT_MIN_3D_PTS (loop_cam.cpp:385-391)
    const int sm = (ok && n > accept_min_3d_pts) ? stereo_map[(size_t)d * max_num + i] : -1;
    rec->stereo_match[d][i] = sm;
    // landmarks_flag / landmarks_3d (loop_cam.cpp:405-432): triangulated on the device when the cameras are known
    const bool fl = ok && (lflag ? lflag[(size_t)d * max_num + i] != 0 : sm >= 0);
    rec->landmarks_flag[d][i] = fl ? 1 : 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) rec->landmarks_3d[d][i][k] = (fl && l3d) ? l3d[((size_t)d * max_num + i) * 3 + k] : 0.f;
  }
}

// add_to_database for a batch of records: phase 1 (one thread) assigns rows in record/direction order
__global__ void fe_assign_kernel(const osb_keyframe_record* __restrict__ recs, int n_records, int skip, int self_id,
                                 DbDev* __restrict__ local, DbDev* __restrict__ remote, long long cap,
                                 int32_t* __restrict__ l_row_frame, int32_t* __restrict__ l_row_dir,
                                 int32_t* __restrict__ l_frame_rows, int32_t* __restrict__ l_frame_msg,
                                 int32_t* __restrict__ r_row_frame, int32_t* __restrict__ r_row_dir,
                                 int32_t* __restrict__ r_frame_rows, int32_t* __restrict__ r_frame_msg,
                                 int32_t* __restrict__ l_frame_drone, int32_t* __restrict__ r_frame_drone,
                                 int32_t* __restrict__ assign /*[n_records][4]: row | (remote<<30), or -1*/,
                                 FeFeedback* __restrict__ fb, long long seq, long long charged) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  for (int r = 0; r < n_records; ++r) {
    for (int d = 0; d < OSB_MAX_DIRS; ++d) assign[r * OSB_MAX_DIRS + d] = -1;
    if (r == skip) continue;
    const osb_keyframe_record* rec = recs + r;
    const bool is_remote = rec->drone_id != self_id;
    DbDev* db = is_remote ? remote : local;
    int32_t* row_frame = is_remote ? r_row_frame : l_row_frame;
    int32_t* row_dir = is_remote ? r_row_dir : l_row_dir;
    int32_t* frame_rows = is_remote ? r_frame_rows : l_frame_rows;
    int32_t* frame_msg = is_remote ? r_frame_msg : l_frame_msg;
    int32_t* frame_drone = is_remote ? r_frame_drone : l_frame_drone;
    if (db->nframes >= cap) { db->overflow = 1; continue; }
    const int fs = db->nframes++;
    frame_msg[fs] = rec->msg_id;
    frame_drone[fs] = rec->drone_id;
    for (int d = 0; d < OSB_MAX_DIRS; ++d) {
      frame_rows[fs * OSB_MAX_DIRS + d] = -1;
      if (d >= rec->n_dirs || rec->n_kpts[d] <= 0) continue;         // landmark_num > 0 (loop_detector.cpp:153)
      if (db->ntotal >= cap) { db->overflow = 1; continue; }
      const int row = (int)db->ntotal++;
      row_frame[row] = fs; row_dir[row] = d;
      frame_rows[fs * OSB_MAX_DIRS + d] = row;
      assign[r * OSB_MAX_DIRS + d] = row | (is_remote ? (1 << 30) : 0);
    }
  }
  fb->seq_begin = seq;
  __threadfence_system();
  fb->n_local = local->ntotal; fb->n_remote = remote->ntotal; fb->charged = charged;
  __threadfence_system();
  fb->seq_end = seq;
}

// phase 2: copy global + local descriptors of every assigned (record, direction) into its row
__global__ void fe_copy_rows_kernel(const osb_keyframe_record* __restrict__ recs, const int32_t* __restrict__ assign,
                                    int max_num, float* __restrict__ l_rows, float* __restrict__ l_ldesc,
                                    int32_t* __restrict__ l_nk, float* __restrict__ r_rows, float* __restrict__ r_ldesc,
                                    int32_t* __restrict__ r_nk, float* __restrict__ l_kpts, int32_t* __restrict__ l_sm,
                                    float* __restrict__ r_kpts, int32_t* __restrict__ r_sm) {
  const int r = blockIdx.x / OSB_MAX_DIRS, d = blockIdx.x % OSB_MAX_DIRS;
  const int a = assign[blockIdx.x];
  if (a < 0) return;
  const bool is_remote = (a >> 30) & 1;
  const int row = a & ((1 << 30) - 1);
  const osb_keyframe_record* rec = recs + r;
  float* rows = is_remote ? r_rows : l_rows;
  float* ldesc = is_remote ? r_ldesc : l_ldesc;
  int32_t* nk = is_remote ? r_nk : l_nk;
  const float4* g = reinterpret_cast<const float4*>(&rec->global_desc[d][0]);
  float4* gd = reinterpret_cast<float4*>(rows + (size_t)row * OSB_DEEP_DESC_SIZE);
  for (int i = threadIdx.x; i < OSB_DEEP_DESC_SIZE / 4; i += blockDim.x) gd[i] = g[i];
  const int n = min(rec->n_kpts[d], max_num);
  const float4* l = reinterpret_cast<const float4*>(&rec->local_desc[d][0][0]);
  float4* ld = reinterpret_cast<float4*>(ldesc + (size_t)row * max_num * OSB_FEATURE_DESC_SIZE);
  for (int i = threadIdx.x; i < n * OSB_FEATURE_DESC_SIZE / 4; i += blockDim.x) ld[i] = l[i];
  float* kp = (is_remote ? r_kpts : l_kpts) + (size_t)row * max_num * 2;
  int32_t* sm = (is_remote ? r_sm : l_sm) + (size_t)row * max_num;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    kp[2 * i] = rec->kpts[d][i][0]; kp[2 * i + 1] = rec->kpts[d][i][1];
    sm[i] = rec->landmarks_flag[d][i];
  }
  if (threadIdx.x == 0) nk[row] = n;
}

struct QueryParams {
  int self_id, n_dirs, query_dir, match_index_dist, init_mode, nonkeyframe, k_remote, k_local, max_num;
  double inner_product_thres, init_mode_product_thres;
};

// literal query_from_database (loop_detector.cpp:199-242) on a top-k result held in device memory
__device__ int fe_search_rule(const float* scores, const int64_t* ids, int k, int64_t ntotal, int max_index,
                              double thres, int index_offset, double* distance) {
  int ret = -1;
  for (int i = 0; i < k; ++i) {
    const int64_t lab = ids[i];
    if (lab < 0) continue;
    // imgid2fisheye holds every row ever added (loop_detector.cpp:155), so the membership test always passes
    ret = (int)lab + index_offset;
    if (lab <= ntotal - max_index && (double)scores[i] > thres) {
      *distance = (double)scores[i];
      return ret;
    }
  }
  return ret;
}

// one thread: the acceptance rule of query_from_database / query_fisheyeframe_from_database and the direction
// pairing of compute_correspond_features (loop_detector.cpp:455-465); fills the matcher's pointer tables.
__global__ void fe_query_rule_kernel(QueryParams qp, const osb_keyframe_record* __restrict__ rec,
                                     const DbDev* __restrict__ local, const DbDev* __restrict__ remote,
                                     const float* __restrict__ l_scores, const int64_t* __restrict__ l_ids,
                                     const float* __restrict__ r_scores, const int64_t* __restrict__ r_ids,
                                     const int32_t* __restrict__ l_row_frame, const int32_t* __restrict__ l_row_dir,
                                     const int32_t* __restrict__ l_frame_rows, const int32_t* __restrict__ l_nk,
                                     const float* __restrict__ l_ldesc,
                                     const int32_t* __restrict__ r_row_frame, const int32_t* __restrict__ r_row_dir,
                                     const int32_t* __restrict__ r_frame_rows, const int32_t* __restrict__ r_nk,
                                     const float* __restrict__ r_ldesc,
                                     osb_loop_result* __restrict__ res, const float** __restrict__ qptr,
                                     const float** __restrict__ tptr, int32_t* __restrict__ nq, int32_t* __restrict__ nt,
                                     const float* __restrict__ l_kpts, const int32_t* __restrict__ l_sm,
                                     const float* __restrict__ r_kpts, const int32_t* __restrict__ r_sm,
                                     const float** __restrict__ g_qk, const float** __restrict__ g_tk,
                                     const int32_t** __restrict__ g_qflag,
                                     const int32_t* __restrict__ l_frame_msg, const int32_t* __restrict__ l_frame_drone,
                                     const int32_t* __restrict__ r_frame_msg, const int32_t* __restrict__ r_frame_drone) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const bool own = rec->drone_id == qp.self_id;
  const double thres = qp.init_mode ? qp.init_mode_product_thres : qp.inner_product_thres;
  double distance = -1.0;                                   // loop_detector.cpp:263
  int id = -1;
  const int64_t db_size = local->ntotal + remote->ntotal;
  // on_image_recv gate (loop_detector.cpp:93) and landmark_num > 0 of the queried direction (:262)
  const bool gate = (db_size > qp.match_index_dist || qp.init_mode || !own) && rec->n_kpts[qp.query_dir] > 0;
  if (gate) {
    if (own) {                                              // :182-190
      const int r = fe_search_rule(r_scores, r_ids, qp.k_remote, remote->ntotal, 1, thres, OSB_REMOTE_MAGIN_NUMBER, &distance);
      if (!qp.nonkeyframe)
        id = fe_search_rule(l_scores, l_ids, 5 + qp.match_index_dist, local->ntotal, qp.match_index_dist, thres, 0, &distance);
      else if (r != -1) id = r;
    } else {                                                // :191-195
      id = fe_search_rule(l_scores, l_ids, 5 + 1, local->ntotal, 1, thres, 0, &distance);
    }
  }
  const bool accepted = (id != -1) && (distance > -1.0);    // :265 (best_distance = -1)
  res->hit_id = id;
  res->hit_score = (float)distance;
  res->accepted = accepted ? 1 : 0;
  res->hit_dir = -1;
  res->swapped = 0;
  res->hit_msg_id = -1;
  res->hit_drone_id = -1;
  for (int j = 0; j < OSB_MAX_DIRS; ++j) {
    nq[j] = 0; nt[j] = 0; qptr[j] = nullptr; tptr[j] = nullptr;
    g_qk[j] = nullptr; g_tk[j] = nullptr; g_qflag[j] = nullptr;
    res->dir_new[j] = -1; res->dir_old[j] = -1;
  }
  if (!accepted) return;
  const bool hit_remote = id >= OSB_REMOTE_MAGIN_NUMBER;
  const int row = hit_remote ? id - OSB_REMOTE_MAGIN_NUMBER : id;
  const int32_t* row_frame = hit_remote ? r_row_frame : l_row_frame;
  const int32_t* row_dir = hit_remote ? r_row_dir : l_row_dir;
  const int32_t* frame_rows = hit_remote ? r_frame_rows : l_frame_rows;
  const int32_t* nk = hit_remote ? r_nk : l_nk;
  const float* ldesc = hit_remote ? r_ldesc : l_ldesc;
  const int fs = row_frame[row];
  const int direction_old = row_dir[row];                   // imgid2dir (:275)
  res->hit_dir = direction_old;
  // imgid2fisheye[best_image_id] -> the keyframe's msg_id, fisheyeframe_database[msg_id].drone_id (loop_detector.cpp:272-275):
  // what the host adapter needs to find the old FisheyeFrameDescriptor_t for compute_loop
  res->hit_msg_id = (hit_remote ? r_frame_msg : l_frame_msg)[fs];
  res->hit_drone_id = (hit_remote ? r_frame_drone : l_frame_drone)[fs];
  // compute_loop(new, old) -- or (old, new) when the hit comes from the remote database and the keyframe is ours
  // (loop_detector.cpp:113-118): the first argument plays "new_frame_desc" (the matcher's query side).
  const bool swapped = hit_remote && own;
  res->swapped = swapped ? 1 : 0;
  const int main_new = swapped ? direction_old : qp.query_dir;
  const int main_old = swapped ? qp.query_dir : direction_old;
  int slot = 0;
  for (int _dn = main_new; _dn < main_new + OSB_MAX_DIRS; ++_dn) {     // :455-465
    const int dir_new = _dn % OSB_MAX_DIRS;
    const int dir_old = ((main_old - main_new + OSB_MAX_DIRS) % OSB_MAX_DIRS + _dn) % OSB_MAX_DIRS;
    // "new" side / "old" side descriptor blocks
    const int dir_rec = swapped ? dir_old : dir_new;          // direction taken from the current keyframe record
    const int dir_db = swapped ? dir_new : dir_old;           // direction taken from the database frame
    if (dir_rec >= rec->n_dirs) continue;
    const int row_db = frame_rows[fs * OSB_MAX_DIRS + dir_db];
    const int n_rec = rec->n_kpts[dir_rec];
    const int n_db = row_db >= 0 ? nk[row_db] : 0;
    if (n_rec <= 0 || n_db <= 0) continue;                    // both landmark_num > 0 (:461)
    const float* p_rec = &rec->local_desc[dir_rec][0][0];
    const float* p_db = ldesc + (size_t)row_db * qp.max_num * OSB_FEATURE_DESC_SIZE;
    res->dir_new[slot] = dir_new; res->dir_old[slot] = dir_old;
    if (swapped) { qptr[slot] = p_db; nq[slot] = n_db; tptr[slot] = p_rec; nt[slot] = n_rec; }
    else { qptr[slot] = p_rec; nq[slot] = n_rec; tptr[slot] = p_db; nt[slot] = n_db; }
    {   // 2-D landmarks and 3-D flags of the two sides, for the geometric filter (loop_detector.cpp:569-598)
      const float* k_rec = &rec->kpts[dir_rec][0][0];
      const int32_t* f_rec = &rec->landmarks_flag[dir_rec][0];
      const float* k_db = (hit_remote ? r_kpts : l_kpts) + (size_t)row_db * qp.max_num * 2;
      const int32_t* f_db = (hit_remote ? r_sm : l_sm) + (size_t)row_db * qp.max_num;
      g_qk[slot] = swapped ? k_db : k_rec; g_tk[slot] = swapped ? k_rec : k_db; g_qflag[slot] = swapped ? f_db : f_rec;
    }
    ++slot;
  }
}

// geometric filter, step 1 (loop_detector.cpp:569-586): keep, in match order, the matches whose NEW (query-side) landmark
// has a 3-D flag; gather old_2d / new_2d of the kept matches.  One CTA per direction-pair slot; ordered compaction by
// ballot prefix.
__global__ void __launch_bounds__(256)
fe_geo_gather_kernel(const osb_loop_result* __restrict__ res, const float* const* __restrict__ g_qk,
                     const float* const* __restrict__ g_tk, const int32_t* const* __restrict__ g_qflag,
                     float2* __restrict__ src, float2* __restrict__ dst, int32_t* __restrict__ kept,
                     int32_t* __restrict__ n_kept) {
  __shared__ int warp_cnt[8];
  const int slot = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = (res->dir_new[slot] >= 0) ? res->n_matches[slot] : 0;
  bool keep = false;
  int qi = 0, ti = 0;
  if (tid < n) {
    qi = res->match_new[slot][tid]; ti = res->match_old[slot][tid];
    keep = g_qflag[slot][qi] != 0;
  }
  const unsigned bal = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) warp_cnt[warp] = __popc(bal);
  __syncthreads();
  int base = 0, total = 0;
  for (int w = 0; w < 8; ++w) { if (w < warp) base += warp_cnt[w]; total += warp_cnt[w]; }
  if (keep) {
    const int pos = base + __popc(bal & ((1u << lane) - 1u));
    const float* tk = g_tk[slot];
    const float* qk = g_qk[slot];
    src[slot * OSB_MAX_KPTS + pos] = make_float2(tk[2 * ti], tk[2 * ti + 1]);      // old_2d
    dst[slot * OSB_MAX_KPTS + pos] = make_float2(qk[2 * qi], qk[2 * qi + 1]);      // new_2d
    kept[slot * OSB_MAX_KPTS + pos] = tid;
  }
  if (tid == 0) n_kept[slot] = total;
}

// step 3 (:590-597): reduceVector by the RANSAC mask
__global__ void __launch_bounds__(256)
fe_geo_apply_kernel(osb_loop_result* __restrict__ res, const int32_t* __restrict__ kept, const int32_t* __restrict__ n_kept,
                    const uint8_t* __restrict__ mask) {
  __shared__ int warp_cnt[8];
  const int slot = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = n_kept[slot];
  const bool valid = n >= 4;                                  // else the reference returns false (:598-600)
  const bool keep = valid && tid < n && mask[slot * OSB_MAX_KPTS + tid] != 0;
  const unsigned bal = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) warp_cnt[warp] = __popc(bal);
  __syncthreads();
  int base = 0, total = 0;
  for (int w = 0; w < 8; ++w) { if (w < warp) base += warp_cnt[w]; total += warp_cnt[w]; }
  if (keep) {
    const int pos = base + __popc(bal & ((1u << lane) - 1u));
    const int m = kept[slot * OSB_MAX_KPTS + tid];
    res->geo_new[slot][pos] = res->match_new[slot][m];
    res->geo_old[slot][pos] = res->match_old[slot][m];
  }
  if (tid == 0) { res->n_geo[slot] = total; res->geo_valid[slot] = (res->dir_new[slot] >= 0 && valid) ? 1 : 0; }
}

}  // namespace osb

using namespace osb;

struct osb_frontend {
  osb_frontend_config cfg;
  int device = 0;
  std::mutex mu;
  cudaStream_t stream = nullptr;
  SuperPoint sp;
  NetVLAD nv;
  DbStore db[2];                 // 0 local, 1 remote
  uint8_t* d_img = nullptr;      // [2*n_dirs][H][W]
  // stereo matcher state
  const float** d_st_q = nullptr; const float** d_st_t = nullptr;   // static pointer tables into sp.d_out
  int32_t *d_st_qi = nullptr, *d_st_ti = nullptr, *d_st_n = nullptr, *d_st_map = nullptr;
  float *d_st_dist = nullptr, *d_dist_scratch = nullptr;
  // query state
  const float** d_q_q = nullptr; const float** d_q_t = nullptr;
  int32_t *d_q_nq = nullptr, *d_q_nt = nullptr;
  float* d_q_dist = nullptr;
  // geometric filter scratch (cfg.geometric_filter)
  const float** d_g_qk = nullptr; const float** d_g_tk = nullptr; const int32_t** d_g_qflag = nullptr;
  float *d_g_src = nullptr, *d_g_dst = nullptr;
  int32_t *d_g_kept = nullptr, *d_g_nkept = nullptr, *d_g_ninl = nullptr, *d_g_win = nullptr;
  uint8_t* d_g_mask = nullptr;
  unsigned int* d_g_scratch = nullptr;
  // stereo triangulation (osb_frontend_set_cameras)
  bool have_cameras = false;
  double K[4] = {0, 0, 0, 0}, triangle_thres = 0.006;
  double left_ext[OSB_MAX_DIRS][7], right_ext[OSB_MAX_DIRS][7], pose_drone[7] = {0, 0, 0, 1, 0, 0, 0};
  double* d_cam_pose = nullptr;  // [2][n_dirs][7]: pose_drone * left / right extrinsics of the current keyframe
  float* d_l3d = nullptr;        // [n_dirs][max_num][3]
  uint8_t *d_lflag_up = nullptr, *d_lflag_down = nullptr;
  int32_t* d_assign = nullptr;   // [max_records][4]
  int max_records = 64;
  osb_keyframe_record* d_record = nullptr;   // used by process()
  osb_loop_result* d_result = nullptr;
  // stage profiling: ev[i] marks the START of stage i, ev[8] the end of the last one
  cudaStream_t stream2 = nullptr;                 // NetVLAD runs here, overlapped with the keypoint kernels
  cudaStream_t stream_sp = nullptr;               // SuperPoint runs here at the highest stream priority (null: caller's stream)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_sp = nullptr;
  DbDev* h_cnt = nullptr;                         // pinned [2]: row counters read back without blocking the driver (a copy
                                                  // to pageable memory would hold other host threads' launches until it ran)
  FeFeedback* fb_host = nullptr;                  // mapped pinned memory + its device alias
  FeFeedback* fb_dev = nullptr;
  long long ingest_seq = 0, fb_min_seq = 1;       // feedback older than fb_min_seq predates a load / reset and is ignored
  long long charged = 0;                          // rows charged to each store by ingests so far
  cudaEvent_t ev_ingest = nullptr;                // recorded after the last ingest on ITS stream
  bool ingest_pending = false;
  bool profiling = false;
  cudaEvent_t ev[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool ev_valid[9] = {false, false, false, false, false, false, false, false, false};
};

static inline void fe_mark(osb_frontend* h, int i, cudaStream_t st) {
  if (!h->profiling) return;
  if (!h->ev[i]) cudaEventCreate(&h->ev[i]);
  cudaEventRecord(h->ev[i], st);
  h->ev_valid[i] = true;
}

static osb_status dbstore_alloc(DbStore& s, int64_t cap, int max_num) {
  s.cap = cap; s.upper = 0;
  OSB_CUDA(cudaMalloc(&s.dev, sizeof(DbDev)));
  OSB_CUDA(cudaMemset(s.dev, 0, sizeof(DbDev)));
  OSB_CUDA(cudaMalloc(&s.rows, (size_t)cap * OSB_DEEP_DESC_SIZE * sizeof(float)));
  OSB_CUDA(cudaMalloc(&s.ldesc, (size_t)cap * max_num * OSB_FEATURE_DESC_SIZE * sizeof(float)));
  OSB_CUDA(cudaMalloc(&s.kpts, (size_t)cap * max_num * 2 * sizeof(float)));
  OSB_CUDA(cudaMalloc(&s.smatch, (size_t)cap * max_num * sizeof(int32_t)));
  OSB_CUDA(cudaMemset(s.kpts, 0, (size_t)cap * max_num * 2 * sizeof(float)));
  OSB_CUDA(cudaMemset(s.smatch, 1, (size_t)cap * max_num * sizeof(int32_t)));   // rows loaded without geometry: every landmark flagged (non-zero)
  OSB_CUDA(cudaMalloc(&s.nk, cap * sizeof(int32_t)));
  OSB_CUDA(cudaMalloc(&s.row_frame, cap * sizeof(int32_t)));
  OSB_CUDA(cudaMalloc(&s.row_dir, cap * sizeof(int32_t)));
  OSB_CUDA(cudaMalloc(&s.frame_rows, cap * OSB_MAX_DIRS * sizeof(int32_t)));
  OSB_CUDA(cudaMalloc(&s.frame_msg, cap * sizeof(int32_t)));
  OSB_CUDA(cudaMalloc(&s.frame_drone, cap * sizeof(int32_t)));
  int64_t chunk;
  const int gmax = db_scan_grid(cap, &chunk);
  OSB_CUDA(cudaMalloc(&s.part_scores, (size_t)8 * gmax * FE_KMAX * sizeof(float)));
  OSB_CUDA(cudaMalloc(&s.part_ids, (size_t)8 * gmax * FE_KMAX * sizeof(int64_t)));
  OSB_CUDA(cudaMalloc(&s.done, sizeof(unsigned int)));
  OSB_CUDA(cudaMemset(s.done, 0, sizeof(unsigned int)));
  OSB_CUDA(cudaMalloc(&s.top_scores, FE_KMAX * sizeof(float)));
  OSB_CUDA(cudaMalloc(&s.top_ids, FE_KMAX * sizeof(int64_t)));
  OSB_CUDA(cudaMemset(s.top_ids, 0xFF, FE_KMAX * sizeof(int64_t)));     // "no result" (-1) until the first scan of this store
  return OSB_OK;
}

static void dbstore_free(DbStore& s) {
  cudaFree(s.dev); cudaFree(s.rows); cudaFree(s.ldesc); cudaFree(s.kpts); cudaFree(s.smatch); cudaFree(s.nk); cudaFree(s.row_frame); cudaFree(s.row_dir);
  cudaFree(s.frame_rows); cudaFree(s.frame_msg); cudaFree(s.frame_drone); cudaFree(s.part_scores); cudaFree(s.part_ids); cudaFree(s.done);
  cudaFree(s.top_scores); cudaFree(s.top_ids);
}

extern "C" osb_status osb_frontend_create(osb_frontend** out, const osb_frontend_config* cfg, const float* sp_weights,
                                          size_t n_sp_weights, const float* pca_comp, const float* pca_mean,
                                          const float* nv_weights, size_t n_nv_weights) {
  OSB_REQUIRE(out && cfg && sp_weights && pca_comp && pca_mean && nv_weights, "null argument");
  OSB_REQUIRE(cfg->n_dirs >= 1 && cfg->n_dirs <= OSB_MAX_DIRS, "n_dirs must be 1..4");
  OSB_REQUIRE(cfg->max_num >= 1 && cfg->max_num <= OSB_MAX_KPTS, "max_num must be 1..200");
  OSB_REQUIRE(cfg->query_dir >= 0 && cfg->query_dir < cfg->n_dirs, "query_dir out of range");
  OSB_REQUIRE(cfg->db_capacity > 0 && cfg->db_capacity < (1 << 30), "bad db_capacity");
  OSB_REQUIRE(cfg->match_index_dist >= 1 && 5 + cfg->match_index_dist <= FE_KMAX, "bad match_index_dist");
  osb_status s = require_device();
  if (s != OSB_OK) return s;
  osb_frontend* h = new osb_frontend();
  h->cfg = *cfg;
  h->device = current_device();
  const int nd = cfg->n_dirs, mn = cfg->max_num;
#define FE_TRY(x) do { s = (x); if (s != OSB_OK) { osb_frontend_destroy(h); return s; } } while (0)
#define FE_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { set_error("osb_frontend_create", cudaGetErrorString(e_)); osb_frontend_destroy(h); return OSB_ERR_CUDA; } } while (0)
  FE_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  FE_CUDA(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
  {
    // SuperPoint is the critical path, NetVLAD (stream2, default = lowest priority) only has to finish before the pack:
    // the convolution CTAs of SuperPoint are dispatched first whenever both streams have work.  Opt-in (OSB_FE_PRIO=1):
    // measured (r01g) it only moves NetVLAD's ~0.11 ms from inside the SuperPoint phase to after it (1.98 vs 1.95 ms).
    const char* e = getenv("OSB_FE_PRIO");
    if (e && atoi(e) != 0) {
      int least = 0, greatest = 0;
      FE_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
      FE_CUDA(cudaStreamCreateWithPriority(&h->stream_sp, cudaStreamNonBlocking, greatest));
    }
  }
  FE_CUDA(cudaEventCreateWithFlags(&h->ev_sp, cudaEventDisableTiming));
  FE_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  FE_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  FE_CUDA(cudaEventCreateWithFlags(&h->ev_ingest, cudaEventDisableTiming));
  FE_CUDA(cudaHostAlloc((void**)&h->h_cnt, 2 * sizeof(DbDev), cudaHostAllocDefault));
  FE_CUDA(cudaHostAlloc((void**)&h->fb_host, sizeof(FeFeedback), cudaHostAllocMapped));
  memset((void*)h->fb_host, 0, sizeof(FeFeedback));
  FE_CUDA(cudaHostGetDevicePointer((void**)&h->fb_dev, (void*)h->fb_host, 0));
  FE_TRY(h->sp.init(sp_weights, n_sp_weights, cfg->width, cfg->height, cfg->sp_thres, mn, pca_comp, pca_mean, 2 * nd));
  h->sp.ks.write_surv = false;       // the survivor plane is only a parity hook of the standalone SuperPoint handle
  FE_TRY(h->nv.init(nv_weights, n_nv_weights, cfg->width, cfg->height, nd));
  FE_TRY(dbstore_alloc(h->db[0], cfg->db_capacity, mn));
  FE_TRY(dbstore_alloc(h->db[1], cfg->db_capacity, mn));
  const size_t HW = (size_t)cfg->width * cfg->height;
  FE_CUDA(cudaMalloc(&h->d_img, 2 * nd * HW));
  FE_CUDA(cudaMalloc(&h->d_st_q, OSB_MAX_DIRS * sizeof(float*)));
  FE_CUDA(cudaMalloc(&h->d_st_t, OSB_MAX_DIRS * sizeof(float*)));
  FE_CUDA(cudaMalloc(&h->d_q_q, OSB_MAX_DIRS * sizeof(float*)));
  FE_CUDA(cudaMalloc(&h->d_q_t, OSB_MAX_DIRS * sizeof(float*)));
  FE_CUDA(cudaMalloc(&h->d_st_qi, OSB_MAX_DIRS * mn * sizeof(int32_t)));
  FE_CUDA(cudaMalloc(&h->d_st_ti, OSB_MAX_DIRS * mn * sizeof(int32_t)));
  FE_CUDA(cudaMalloc(&h->d_st_map, OSB_MAX_DIRS * mn * sizeof(int32_t)));
  FE_CUDA(cudaMalloc(&h->d_st_n, OSB_MAX_DIRS * sizeof(int32_t)));
  FE_CUDA(cudaMalloc(&h->d_st_dist, OSB_MAX_DIRS * mn * sizeof(float)));
  FE_CUDA(cudaMalloc(&h->d_q_dist, OSB_MAX_DIRS * OSB_MAX_KPTS * sizeof(float)));
  FE_CUDA(cudaMalloc(&h->d_g_qk, OSB_MAX_DIRS * sizeof(float*)));
  FE_CUDA(cudaMalloc(&h->d_g_tk, OSB_MAX_DIRS * sizeof(float*)));
  FE_CUDA(cudaMalloc(&h->d_g_qflag, OSB_MAX_DIRS * sizeof(int32_t*)));
  FE_CUDA(cudaMalloc(&h->d_g_src, OSB_MAX_DIRS * OSB_MAX_KPTS * 2 * sizeof(float)));
  FE_CUDA(cudaMalloc(&h->d_g_dst, OSB_MAX_DIRS * OSB_MAX_KPTS * 2 * sizeof(float)));
  FE_CUDA(cudaMalloc(&h->d_g_kept, OSB_MAX_DIRS * OSB_MAX_KPTS * sizeof(int32_t)));
  FE_CUDA(cudaMalloc(&h->d_g_nkept, OSB_MAX_DIRS * sizeof(int32_t)));
  FE_CUDA(cudaMalloc(&h->d_g_ninl, OSB_MAX_DIRS * sizeof(int32_t)));
  FE_CUDA(cudaMalloc(&h->d_g_win, OSB_MAX_DIRS * sizeof(int32_t)));
  FE_CUDA(cudaMalloc(&h->d_g_mask, OSB_MAX_DIRS * OSB_MAX_KPTS));
  FE_CUDA(cudaMalloc(&h->d_g_scratch, 2 * OSB_MAX_DIRS * sizeof(unsigned int)));
  FE_CUDA(cudaMemset(h->d_g_scratch, 0, 2 * OSB_MAX_DIRS * sizeof(unsigned int)));
  FE_CUDA(cudaMalloc(&h->d_dist_scratch, (size_t)OSB_MAX_DIRS * mn * mn * sizeof(float)));
  FE_CUDA(cudaMalloc(&h->d_q_nq, OSB_MAX_DIRS * sizeof(int32_t)));
  FE_CUDA(cudaMalloc(&h->d_q_nt, OSB_MAX_DIRS * sizeof(int32_t)));
  FE_CUDA(cudaMalloc(&h->d_assign, h->max_records * OSB_MAX_DIRS * sizeof(int32_t)));
  FE_CUDA(cudaMalloc(&h->d_cam_pose, 2 * OSB_MAX_DIRS * 7 * sizeof(double)));
  FE_CUDA(cudaMalloc(&h->d_l3d, (size_t)OSB_MAX_DIRS * mn * 3 * sizeof(float)));
  FE_CUDA(cudaMalloc(&h->d_lflag_up, (size_t)OSB_MAX_DIRS * mn));
  FE_CUDA(cudaMalloc(&h->d_lflag_down, (size_t)OSB_MAX_DIRS * mn));
  FE_CUDA(cudaMalloc(&h->d_record, sizeof(osb_keyframe_record)));
  FE_CUDA(cudaMalloc(&h->d_result, sizeof(osb_loop_result)));
  {
    const float* q[OSB_MAX_DIRS] = {nullptr, nullptr, nullptr, nullptr};
    const float* t[OSB_MAX_DIRS] = {nullptr, nullptr, nullptr, nullptr};
    for (int d = 0; d < nd; ++d) {
      q[d] = h->sp.d_out + (size_t)d * mn * OSB_FEATURE_DESC_SIZE;
      t[d] = h->sp.d_out + (size_t)(nd + d) * mn * OSB_FEATURE_DESC_SIZE;
    }
    FE_CUDA(cudaMemcpy(h->d_st_q, q, sizeof(q), cudaMemcpyHostToDevice));
    FE_CUDA(cudaMemcpy(h->d_st_t, t, sizeof(t), cudaMemcpyHostToDevice));
  }
#undef FE_TRY
#undef FE_CUDA
  *out = h;
  return OSB_OK;
}

extern "C" osb_status osb_frontend_destroy(osb_frontend* h) {
  if (!h) return OSB_OK;
  h->sp.release(); h->nv.release();
  dbstore_free(h->db[0]); dbstore_free(h->db[1]);
  cudaFree(h->d_img); cudaFree(h->d_st_q); cudaFree(h->d_st_t); cudaFree(h->d_q_q); cudaFree(h->d_q_t);
  cudaFree(h->d_st_qi); cudaFree(h->d_st_ti); cudaFree(h->d_st_map); cudaFree(h->d_st_n); cudaFree(h->d_st_dist);
  cudaFree(h->d_g_qk); cudaFree(h->d_g_tk); cudaFree(h->d_g_qflag); cudaFree(h->d_g_src); cudaFree(h->d_g_dst);
  cudaFree(h->d_g_kept); cudaFree(h->d_g_nkept); cudaFree(h->d_g_ninl); cudaFree(h->d_g_win); cudaFree(h->d_g_mask); cudaFree(h->d_g_scratch);
  cudaFree(h->d_q_dist); cudaFree(h->d_dist_scratch); cudaFree(h->d_q_nq); cudaFree(h->d_q_nt); cudaFree(h->d_assign);
  cudaFree(h->d_record); cudaFree(h->d_result);
  cudaFree(h->d_cam_pose); cudaFree(h->d_l3d); cudaFree(h->d_lflag_up); cudaFree(h->d_lflag_down);
  for (int i = 0; i < 9; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->ev_ingest) cudaEventDestroy(h->ev_ingest);
  if (h->h_cnt) cudaFreeHost(h->h_cnt);
  if (h->fb_host) cudaFreeHost((void*)h->fb_host);
  if (h->stream2) cudaStreamDestroy(h->stream2);
  if (h->stream_sp) cudaStreamDestroy(h->stream_sp);
  if (h->ev_sp) cudaEventDestroy(h->ev_sp);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return OSB_OK;
}

// a * b for poses (x y z, qw qx qy qz), host side
static void pose_compose_host(const double* a, const double* b, double* o) {
  const double w = a[3], x = a[4], y = a[5], z = a[6];
  const double cx = y * b[2] - z * b[1], cy = z * b[0] - x * b[2], cz = x * b[1] - y * b[0];
  const double dx = y * cz - z * cy, dy = z * cx - x * cz, dz = x * cy - y * cx;
  o[0] = a[0] + b[0] + 2.0 * (w * cx + dx);
  o[1] = a[1] + b[1] + 2.0 * (w * cy + dy);
  o[2] = a[2] + b[2] + 2.0 * (w * cz + dz);
  o[3] = w * b[3] - x * b[4] - y * b[5] - z * b[6];
  o[4] = w * b[4] + x * b[3] + y * b[6] - z * b[5];
  o[5] = w * b[5] - x * b[6] + y * b[3] + z * b[4];
  o[6] = w * b[6] + x * b[5] - y * b[4] + z * b[3];
}

static osb_status fe_extract_dev(osb_frontend* h, const uint8_t* img_dev /*[2*nd][H][W] up then down*/, int32_t msg_id,
                                 osb_keyframe_record* record_dev, cudaStream_t st) {
  const osb_frontend_config& c = h->cfg;
  const int nd = c.n_dirs, mn = c.max_num;
  osb_status s;
  if (c.zero_bottom_quarter) {
    OSB_LAUNCH(fe_blank_kernel, 256, 256, 0, st, const_cast<uint8_t*>(img_dev), c.height, c.width, 2 * nd);
    OSB_CHECK_LAUNCH();
  }
  // fork: NetVLAD is independent of SuperPoint (it only reads the images).  It runs on a second stream for the whole
  // SuperPoint phase: its element-wise / depthwise kernels co-reside with the persistent convolution CTAs, its
  // tensor-core launches fill the wave tails, and while the keypoint kernels run (one CTA per image) it owns the
  // other 140 SMs.  It joins before the record is packed.
  OSB_CUDA(cudaEventRecord(h->ev_fork, st));
  OSB_CUDA(cudaStreamWaitEvent(h->stream2, h->ev_fork, 0));
  // NetVLAD on the up images writes straight into the record (image_desc, loop_cam.cpp:553-556)
  // (OSB_FE_SKIP_NV=1 is a measurement switch only -- it leaves the global descriptors stale -- used to attribute the
  //  cost of sharing the GPU with NetVLAD)
  static const bool skip_nv = [] { const char* e = getenv("OSB_FE_SKIP_NV"); return e && atoi(e) != 0; }();
  if (!skip_nv && (s = h->nv.infer_dev(img_dev, nd, &record_dev->global_desc[0][0], h->stream2)) != OSB_OK) return s;
  OSB_CUDA(cudaEventRecord(h->ev_join, h->stream2));
  cudaStream_t sps = h->stream_sp ? h->stream_sp : st;
  if (sps != st) OSB_CUDA(cudaStreamWaitEvent(sps, h->ev_fork, 0));
  fe_mark(h, 0, sps);
  // network; the keypoint kernel is forked beside the descriptor head inside (superpoint.cu) and joined before return
  const SuperPoint::KpJob kp{h->sp.d_nk, h->sp.d_kpts, h->sp.d_conf};
  if ((s = h->sp.network(img_dev, 2 * nd, sps, &kp)) != OSB_OK) return s;
  h->sp.last_batch = 2 * nd;
  fe_mark(h, 1, sps);
  if ((s = h->sp.descriptors(2 * nd, kp, h->sp.d_out, sps)) != OSB_OK) return s;
  fe_mark(h, 2, sps);
  if (sps != st) {
    OSB_CUDA(cudaEventRecord(h->ev_sp, sps));
    OSB_CUDA(cudaStreamWaitEvent(st, h->ev_sp, 0));
  }
  OSB_CUDA(cudaStreamWaitEvent(st, h->ev_join, 0));      // join (stage 2 = the part of NetVLAD that was not hidden)
  fe_mark(h, 3, st);
  // stereo match up[d] <-> down[d] (loop_cam.cpp:388)
  if ((s = bf_match_device(nd, mn, mn, h->d_st_q, h->sp.d_nk, h->d_st_t, h->sp.d_nk + nd, h->d_dist_scratch,
                           h->d_st_qi, h->d_st_ti, h->d_st_dist, h->d_st_n, h->d_st_map, st)) != OSB_OK) return s;
  if (h->have_cameras) {
    // pose_up / pose_down = pose_drone * extrinsics (loop_cam.cpp:394-396), then the per-keypoint triangulation (:398-432)
    double cam[2][OSB_MAX_DIRS][7];
    for (int d = 0; d < nd; ++d) {
      pose_compose_host(h->pose_drone, h->left_ext[d], cam[0][d]);
      pose_compose_host(h->pose_drone, h->right_ext[d], cam[1][d]);
    }
    OSB_CUDA(cudaMemcpyAsync(h->d_cam_pose, cam[0], (size_t)nd * 7 * sizeof(double), cudaMemcpyHostToDevice, st));
    OSB_CUDA(cudaMemcpyAsync(h->d_cam_pose + OSB_MAX_DIRS * 7, cam[1], (size_t)nd * 7 * sizeof(double), cudaMemcpyHostToDevice, st));
    if ((s = stereo_lift_device(h->sp.d_kpts, h->sp.d_kpts + (size_t)nd * mn * 2, h->d_st_map, h->sp.d_nk, h->sp.d_nk + nd, nd, mn,
                                h->K, h->d_cam_pose, h->d_cam_pose + OSB_MAX_DIRS * 7, h->triangle_thres, c.accept_min_3d_pts,
                                h->d_l3d, h->d_lflag_up, h->d_lflag_down, st)) != OSB_OK) return s;
  }
  OSB_LAUNCH(fe_pack_kernel, OSB_MAX_DIRS, 256, 0, st, record_dev, c.self_id, msg_id, nd, mn, h->sp.d_nk, h->sp.d_kpts,
             h->sp.d_out, h->d_st_map, c.accept_min_3d_pts, h->have_cameras ? h->d_l3d : nullptr,
             h->have_cameras ? h->d_lflag_up : nullptr);
  OSB_CHECK_LAUNCH();
  fe_mark(h, 4, st);
  return OSB_OK;
}

extern "C" osb_status osb_frontend_extract_dev(osb_frontend* h, const uint8_t* images_up_dev,
                                               const uint8_t* images_down_dev, int32_t msg_id,
                                               osb_keyframe_record* record_dev, void* stream) {
  OSB_REQUIRE(h && images_up_dev && images_down_dev && record_dev, "null argument");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t half = (size_t)h->cfg.n_dirs * h->cfg.width * h->cfg.height;
  OSB_CUDA(cudaMemcpyAsync(h->d_img, images_up_dev, half, cudaMemcpyDeviceToDevice, st));
  OSB_CUDA(cudaMemcpyAsync(h->d_img + half, images_down_dev, half, cudaMemcpyDeviceToDevice, st));
  return fe_extract_dev(h, h->d_img, msg_id, record_dev, st);
}

extern "C" osb_status osb_frontend_extract(osb_frontend* h, const uint8_t* images_up, const uint8_t* images_down,
                                           int32_t msg_id, osb_keyframe_record* record_dev, void* stream) {
  OSB_REQUIRE(h && images_up && images_down && record_dev, "null argument");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t half = (size_t)h->cfg.n_dirs * h->cfg.width * h->cfg.height;
  OSB_CUDA(cudaMemcpyAsync(h->d_img, images_up, half, cudaMemcpyHostToDevice, st));
  OSB_CUDA(cudaMemcpyAsync(h->d_img + half, images_down, half, cudaMemcpyHostToDevice, st));
  return fe_extract_dev(h, h->d_img, msg_id, record_dev, st);
}

static osb_status fe_refresh_counts(osb_frontend* h, cudaStream_t st);

// lower the host bounds with what the most recent EXECUTED ingest reported (no synchronisation; see FeFeedback)
static void fe_tighten_bounds(osb_frontend* h) {
  const FeFeedback* fb = h->fb_host;
  const long long e = fb->seq_end;
  std::atomic_thread_fence(std::memory_order_acquire);
  const long long nl = fb->n_local, nr = fb->n_remote, ch = fb->charged;
  std::atomic_thread_fence(std::memory_order_acquire);
  const long long b = fb->seq_begin;
  if (b != e || e < h->fb_min_seq) return;
  const long long since = h->charged - ch;
  h->db[0].upper = std::min<int64_t>(h->db[0].upper, nl + since);
  h->db[1].upper = std::min<int64_t>(h->db[1].upper, nr + since);
}

// side: which store the host CHARGES for the batch (its bounds): 0 = unknown (both), 1 = all records are this drone's own,
// 2 = all records are foreign.  The kernel always routes by drone_id.
static osb_status fe_ingest(osb_frontend* h, const osb_keyframe_record* recs, int n_records, int skip, cudaStream_t st,
                            int side = 0) {
  OSB_REQUIRE(n_records >= 0 && n_records <= h->max_records, "too many records in one ingest (max 64)");
  if (n_records == 0) return OSB_OK;
  DbStore &L = h->db[0], &R = h->db[1];
  fe_tighten_bounds(h);
  const int64_t chg_l = side == 2 ? 0 : (int64_t)n_records * OSB_MAX_DIRS, chg_r = side == 1 ? 0 : (int64_t)n_records * OSB_MAX_DIRS;
  if (L.upper + chg_l > L.cap || R.upper + chg_r > R.cap) {
    // the upper bounds are conservative (every record charged to both databases): make them exact, and count which
    // database each record of this batch really goes to, before giving up
    osb_status rs = fe_refresh_counts(h, st);
    if (rs != OSB_OK) return rs;
    std::vector<int32_t> hdr((size_t)n_records * 4);
    OSB_CUDA(cudaMemcpy2DAsync(hdr.data(), 16, recs, sizeof(osb_keyframe_record), 16, n_records, cudaMemcpyDeviceToHost, st));
    OSB_CUDA(cudaStreamSynchronize(st));
    int64_t n_loc = 0, n_rem = 0;
    for (int r = 0; r < n_records; ++r) {
      if (r == skip) continue;
      (hdr[(size_t)r * 4] == h->cfg.self_id ? n_loc : n_rem) += OSB_MAX_DIRS;
    }
    if (L.upper + n_loc > L.cap || R.upper + n_rem > R.cap) {
      set_error("osb_frontend_ingest", "database capacity exceeded");
      return OSB_ERR_CAPACITY;
    }
  }
  fe_mark(h, 4, st);
  OSB_LAUNCH(fe_assign_kernel, 1, 32, 0, st, recs, n_records, skip, h->cfg.self_id, L.dev, R.dev, (long long)L.cap,
             L.row_frame, L.row_dir, L.frame_rows, L.frame_msg, R.row_frame, R.row_dir, R.frame_rows, R.frame_msg,
             L.frame_drone, R.frame_drone, h->d_assign, h->fb_dev, h->ingest_seq + 1,
             h->charged + (long long)n_records * OSB_MAX_DIRS);
  OSB_CHECK_LAUNCH();
  ++h->ingest_seq;
  h->charged += (long long)n_records * OSB_MAX_DIRS;
  OSB_LAUNCH(fe_copy_rows_kernel, n_records * OSB_MAX_DIRS, 256, 0, st, recs, h->d_assign, h->cfg.max_num, L.rows,
             L.ldesc, L.nk, R.rows, R.ldesc, R.nk, L.kpts, L.smatch, R.kpts, R.smatch);
  OSB_CHECK_LAUNCH();
  L.upper += chg_l;
  R.upper += chg_r;
  OSB_CUDA(cudaEventRecord(h->ev_ingest, st));
  h->ingest_pending = true;
  fe_mark(h, 5, st);
  return OSB_OK;
}

extern "C" osb_status osb_frontend_ingest_own(osb_frontend* h, const osb_keyframe_record* record_dev, void* stream) {
  OSB_REQUIRE(h && record_dev, "null argument");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  return fe_ingest(h, record_dev, 1, -1, (cudaStream_t)stream, 1);
}

extern "C" osb_status osb_frontend_ingest(osb_frontend* h, const osb_keyframe_record* records_dev, int n_records,
                                          int skip, void* stream) {
  OSB_REQUIRE(h && records_dev, "null argument");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  return fe_ingest(h, records_dev, n_records, skip, (cudaStream_t)stream);
}

static osb_status fe_query(osb_frontend* h, const osb_keyframe_record* rec, int init_mode, int nonkeyframe,
                           osb_loop_result* res, cudaStream_t st) {
  const osb_frontend_config& c = h->cfg;
  DbStore &L = h->db[0], &R = h->db[1];
  const float* q = &rec->global_desc[c.query_dir][0];
  const int k_local = 5 + c.match_index_dist, k_remote = 5 + 1;    // SEARCH_NEAREST_NUM + max_index
  osb_status s;
  fe_mark(h, 5, st);
  fe_tighten_bounds(h);
  // a store that has never received a row needs no scan: its top-k list still holds the -1 labels it was created with
  // (upper is an upper bound of ntotal, so 0 is exact)
  if (R.upper > 0 &&
      (s = db_search_device(R.rows, std::min(R.upper, R.cap), &R.dev->ntotal, OSB_DEEP_DESC_SIZE, q, 1, k_remote,
                            R.part_scores, R.part_ids, R.done, R.top_scores, R.top_ids, st)) != OSB_OK) return s;
  if ((s = db_search_device(L.rows, std::min(L.upper, L.cap), &L.dev->ntotal, OSB_DEEP_DESC_SIZE, q, 1, k_local,
                            L.part_scores, L.part_ids, L.done, L.top_scores, L.top_ids, st)) != OSB_OK) return s;
  fe_mark(h, 6, st);
  QueryParams qp;
  qp.self_id = c.self_id; qp.n_dirs = c.n_dirs; qp.query_dir = c.query_dir; qp.match_index_dist = c.match_index_dist;
  qp.init_mode = init_mode; qp.nonkeyframe = nonkeyframe; qp.k_remote = k_remote; qp.k_local = k_local;
  qp.max_num = c.max_num;
  qp.inner_product_thres = c.inner_product_thres; qp.init_mode_product_thres = c.init_mode_product_thres;
  OSB_LAUNCH(fe_query_rule_kernel, 1, 32, 0, st, qp, rec, L.dev, R.dev, L.top_scores, L.top_ids, R.top_scores, R.top_ids,
             L.row_frame, L.row_dir, L.frame_rows, L.nk, L.ldesc, R.row_frame, R.row_dir, R.frame_rows, R.nk, R.ldesc,
             res, h->d_q_q, h->d_q_t, h->d_q_nq, h->d_q_nt, L.kpts, L.smatch, R.kpts, R.smatch, h->d_g_qk, h->d_g_tk,
             h->d_g_qflag, L.frame_msg, L.frame_drone, R.frame_msg, R.frame_drone);
  OSB_CHECK_LAUNCH();
  // per-direction cross-check match new vs old (loop_detector.cpp:564-567); empty pairs produce n = 0
  s = bf_match_device(OSB_MAX_DIRS, c.max_num, OSB_MAX_KPTS, h->d_q_q, h->d_q_nq, h->d_q_t, h->d_q_nt,
                      h->d_dist_scratch, &res->match_new[0][0], &res->match_old[0][0], h->d_q_dist,
                      &res->n_matches[0], nullptr, st);
  if (s != OSB_OK) return s;
  if (c.geometric_filter) {
    // loop_detector.cpp:569-598: 3-D-flag filter, homography RANSAC mask, reduceVector -- all direction pairs at once
    OSB_LAUNCH(fe_geo_gather_kernel, OSB_MAX_DIRS, 256, 0, st, res, h->d_g_qk, h->d_g_tk, h->d_g_qflag,
               reinterpret_cast<float2*>(h->d_g_src), reinterpret_cast<float2*>(h->d_g_dst), h->d_g_kept, h->d_g_nkept);
    OSB_CHECK_LAUNCH();
    if ((s = homography_ransac_device(h->d_g_src, h->d_g_dst, h->d_g_nkept, OSB_MAX_DIRS, OSB_MAX_KPTS, 3.0f,
                                      (uint32_t)c.ransac_seed, h->d_g_mask, h->d_g_ninl, h->d_g_win, st, h->d_g_scratch)) != OSB_OK) return s;
    OSB_LAUNCH(fe_geo_apply_kernel, OSB_MAX_DIRS, 256, 0, st, res, h->d_g_kept, h->d_g_nkept, h->d_g_mask);
    OSB_CHECK_LAUNCH();
  }
  fe_mark(h, 7, st);
  return OSB_OK;
}

extern "C" osb_status osb_frontend_query(osb_frontend* h, const osb_keyframe_record* record_dev, int init_mode,
                                         int nonkeyframe, osb_loop_result* result_dev, void* stream) {
  OSB_REQUIRE(h && record_dev && result_dev, "null argument");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  return fe_query(h, record_dev, init_mode, nonkeyframe, result_dev, (cudaStream_t)stream);
}

// exact host-side row counts.  The counts are read on `st`, which need not be the stream that carried the last ingest
// (streams are non-blocking): wait for that ingest first, or the bound could be lowered below the true count while the
// rows are still being appended -- the scan grid is sized from it.
static osb_status fe_refresh_counts(osb_frontend* h, cudaStream_t st) {
  if (h->ingest_pending) OSB_CUDA(cudaStreamWaitEvent(st, h->ev_ingest, 0));
  OSB_CUDA(cudaMemcpyAsync(&h->h_cnt[0], h->db[0].dev, sizeof(DbDev), cudaMemcpyDeviceToHost, st));
  OSB_CUDA(cudaMemcpyAsync(&h->h_cnt[1], h->db[1].dev, sizeof(DbDev), cudaMemcpyDeviceToHost, st));
  OSB_CUDA(cudaStreamSynchronize(st));
  h->db[0].upper = h->h_cnt[0].ntotal; h->db[1].upper = h->h_cnt[1].ntotal;
  h->ingest_pending = false;
  h->fb_min_seq = h->ingest_seq + 1;            // exact now: older feedback carries nothing new
  return OSB_OK;
}

extern "C" osb_status osb_frontend_process(osb_frontend* h, const uint8_t* images_up, const uint8_t* images_down,
                                           int32_t msg_id, osb_keyframe_record* record_host,
                                           osb_loop_result* result_host) {
  OSB_REQUIRE(h && images_up && images_down, "null argument");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  cudaStream_t st = h->stream;
  const size_t half = (size_t)h->cfg.n_dirs * h->cfg.width * h->cfg.height;
  OSB_CUDA(cudaMemcpyAsync(h->d_img, images_up, half, cudaMemcpyHostToDevice, st));
  OSB_CUDA(cudaMemcpyAsync(h->d_img + half, images_down, half, cudaMemcpyHostToDevice, st));
  osb_status s;
  if ((s = fe_extract_dev(h, h->d_img, msg_id, h->d_record, st)) != OSB_OK) return s;
  if ((s = fe_ingest(h, h->d_record, 1, -1, st, 1)) != OSB_OK) return s;   // add_to_database (loop_detector.cpp:89): own record
  if ((s = fe_query(h, h->d_record, 0, 0, h->d_result, st)) != OSB_OK) return s;
  if (record_host) OSB_CUDA(cudaMemcpyAsync(record_host, h->d_record, sizeof(osb_keyframe_record), cudaMemcpyDeviceToHost, st));
  if (result_host) OSB_CUDA(cudaMemcpyAsync(result_host, h->d_result, sizeof(osb_loop_result), cudaMemcpyDeviceToHost, st));
  return fe_refresh_counts(h, st);     // the one synchronisation of the keyframe
}

extern "C" osb_status osb_frontend_set_cameras(osb_frontend* h, const double* intrinsics, const double* left_extrinsics,
                                               const double* right_extrinsics, double triangle_thres) {
  OSB_REQUIRE(h && intrinsics && left_extrinsics && right_extrinsics, "null argument");
  OSB_REQUIRE(intrinsics[0] > 0 && intrinsics[1] > 0 && triangle_thres > 0, "bad intrinsics / threshold");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  for (int i = 0; i < 4; ++i) h->K[i] = intrinsics[i];
  for (int d = 0; d < h->cfg.n_dirs; ++d)
    for (int i = 0; i < 7; ++i) { h->left_ext[d][i] = left_extrinsics[d * 7 + i]; h->right_ext[d][i] = right_extrinsics[d * 7 + i]; }
  h->triangle_thres = triangle_thres;
  h->have_cameras = true;
  return OSB_OK;
}

extern "C" osb_status osb_frontend_set_drone_pose(osb_frontend* h, const double* pose_drone) {
  OSB_REQUIRE(h && pose_drone, "null argument");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  for (int i = 0; i < 7; ++i) h->pose_drone[i] = pose_drone[i];
  return OSB_OK;
}

extern "C" osb_status osb_frontend_set_profiling(osb_frontend* h, int enable) {
  OSB_REQUIRE(h != nullptr, "null handle");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  h->profiling = enable != 0;
  for (int i = 0; i < 9; ++i) h->ev_valid[i] = false;
  return OSB_OK;
}

extern "C" osb_status osb_frontend_stage_ms(osb_frontend* h, float* ms8) {
  OSB_REQUIRE(h != nullptr && ms8 != nullptr, "null argument");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  for (int i = 0; i < 8; ++i) {
    ms8[i] = 0.f;
    if (i < 7 && h->ev_valid[i] && h->ev_valid[i + 1]) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, h->ev[i], h->ev[i + 1]) == cudaSuccess) ms8[i] = t; else cudaGetLastError();
    }
  }
  return OSB_OK;
}

extern "C" osb_status osb_frontend_finish(osb_frontend* h, void* stream) {
  OSB_REQUIRE(h != nullptr, "null handle");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  return fe_refresh_counts(h, (cudaStream_t)stream);
}

extern "C" int64_t osb_frontend_db_size(osb_frontend* h, int remote) {
  if (!h) return -1;
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  if (fe_refresh_counts(h, h->stream) != OSB_OK) return -1;
  return h->db[remote ? 1 : 0].upper;
}

extern "C" osb_status osb_frontend_db_reset(osb_frontend* h) {
  OSB_REQUIRE(h != nullptr, "null handle");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  for (int i = 0; i < 2; ++i) {
    OSB_CUDA(cudaMemsetAsync(h->db[i].dev, 0, sizeof(DbDev), h->stream));
    OSB_CUDA(cudaMemsetAsync(h->db[i].top_ids, 0xFF, FE_KMAX * sizeof(int64_t), h->stream));
    h->db[i].upper = 0;
  }
  h->fb_min_seq = h->ingest_seq + 1;
  OSB_CUDA(cudaStreamSynchronize(h->stream));
  return OSB_OK;
}

// landmarks_2d and stereo_match (>= 0 <=> landmarks_flag) of rows [first_row, first_row + n) that were put in with
// osb_frontend_db_load: the inputs of the geometric filter when such a row is the loop hit
extern "C" osb_status osb_frontend_db_set_geometry(osb_frontend* h, int remote, int64_t first_row, int64_t n,
                                                   const float* kpts, const int32_t* stereo_match) {
  OSB_REQUIRE(h && kpts && stereo_match && n >= 0 && first_row >= 0, "bad arguments");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  cudaStream_t st = h->stream;
  if (fe_refresh_counts(h, st) != OSB_OK) return OSB_ERR_CUDA;
  DbStore& S = h->db[remote ? 1 : 0];
  OSB_REQUIRE(first_row + n <= S.upper, "rows out of range");
  const int mn = h->cfg.max_num;
  OSB_CUDA(cudaMemcpyAsync(S.kpts + (size_t)first_row * mn * 2, kpts, (size_t)n * mn * 2 * sizeof(float),
                           cudaMemcpyHostToDevice, st));
  std::vector<int32_t> flags((size_t)n * mn);
  for (size_t i = 0; i < flags.size(); ++i) flags[i] = stereo_match[i] >= 0 ? 1 : 0;       // landmarks_flag of the rows
  OSB_CUDA(cudaMemcpyAsync(S.smatch + (size_t)first_row * mn, flags.data(), flags.size() * sizeof(int32_t),
                           cudaMemcpyHostToDevice, st));
  OSB_CUDA(cudaStreamSynchronize(st));
  return OSB_OK;
}

extern "C" osb_status osb_frontend_db_load(osb_frontend* h, int remote, int64_t n, const float* global_desc,
                                           const float* local_desc, const int32_t* n_kpts) {
  OSB_REQUIRE(h && global_desc && n >= 0, "bad arguments");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard dg(h->device);
  cudaStream_t st = h->stream;
  if (fe_refresh_counts(h, st) != OSB_OK) return OSB_ERR_CUDA;
  DbStore& S = h->db[remote ? 1 : 0];
  const int64_t base = S.upper;
  if (base + n > S.cap) { set_error("osb_frontend_db_load", "database capacity exceeded"); return OSB_ERR_CAPACITY; }
  const int mn = h->cfg.max_num, qd = h->cfg.query_dir;
  DbDev hd;
  OSB_CUDA(cudaMemcpy(&hd, S.dev, sizeof(DbDev), cudaMemcpyDeviceToHost));
  // the frame tables are [cap] too, and a keyframe without keypoints adds a frame but no row (nframes may exceed ntotal)
  if ((int64_t)hd.nframes + n > S.cap) { set_error("osb_frontend_db_load", "frame table capacity exceeded"); return OSB_ERR_CAPACITY; }
  OSB_CUDA(cudaMemcpyAsync(S.rows + (size_t)base * OSB_DEEP_DESC_SIZE, global_desc,
                           (size_t)n * OSB_DEEP_DESC_SIZE * sizeof(float), cudaMemcpyHostToDevice, st));
  std::vector<int32_t> nk(n), rf(n), rd(n), fr((size_t)n * OSB_MAX_DIRS, -1), fm(n, -1);
  for (int64_t i = 0; i < n; ++i) {
    nk[i] = (local_desc && n_kpts) ? std::min(n_kpts[i], mn) : 0;
    rf[i] = hd.nframes + (int)i; rd[i] = qd;
    fr[(size_t)i * OSB_MAX_DIRS + qd] = (int32_t)(base + i);
  }
  if (local_desc)
    OSB_CUDA(cudaMemcpyAsync(S.ldesc + (size_t)base * mn * OSB_FEATURE_DESC_SIZE, local_desc,
                             (size_t)n * mn * OSB_FEATURE_DESC_SIZE * sizeof(float), cudaMemcpyHostToDevice, st));
  // rows loaded without geometry: landmarks at the origin, every landmark flagged (osb_frontend_db_set_geometry fills them)
  OSB_CUDA(cudaMemsetAsync(S.kpts + (size_t)base * mn * 2, 0, (size_t)n * mn * 2 * sizeof(float), st));
  OSB_CUDA(cudaMemsetAsync(S.smatch + (size_t)base * mn, 1, (size_t)n * mn * sizeof(int32_t), st));
  OSB_CUDA(cudaMemcpyAsync(S.nk + base, nk.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  OSB_CUDA(cudaMemcpyAsync(S.row_frame + base, rf.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  OSB_CUDA(cudaMemcpyAsync(S.row_dir + base, rd.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  OSB_CUDA(cudaMemcpyAsync(S.frame_rows + (size_t)hd.nframes * OSB_MAX_DIRS, fr.data(), fr.size() * sizeof(int32_t),
                           cudaMemcpyHostToDevice, st));
  OSB_CUDA(cudaMemcpyAsync(S.frame_msg + hd.nframes, fm.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  std::vector<int32_t> fd(n, remote ? -1 : h->cfg.self_id);
  OSB_CUDA(cudaMemcpyAsync(S.frame_drone + hd.nframes, fd.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  hd.ntotal += n; hd.nframes += (int)n;
  OSB_CUDA(cudaMemcpyAsync(S.dev, &hd, sizeof(DbDev), cudaMemcpyHostToDevice, st));
  OSB_CUDA(cudaStreamSynchronize(st));
  S.upper = hd.ntotal;
  return OSB_OK;
}
