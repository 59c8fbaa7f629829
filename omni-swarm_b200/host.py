"""Host-side mirror of the reference's C++ call sites, one class per replaced object.

Names, argument meaning and error behaviour follow the reference (paths relative to /root/reference):
  SuperPoint        <- SuperPointTensorRT            swarm_loop/include/swarm_loop/superpoint_tensorrt.h:20-28
  NetVLAD           <- MobileNetVLADTensorRT         swarm_loop/include/swarm_loop/mobilenetvlad_tensorrt.h:10-21
  IndexFlatIP       <- faiss::IndexFlatIP            swarm_loop/include/swarm_loop/loop_detector.h:27-29
  BFMatcher         <- cv::BFMatcher(NORM_L2, true)  swarm_loop/src/loop_cam.cpp:147-150
  PoseGraphSolver   <- SwarmLocalizationSolver::solve_once  swarm_localization/src/swarm_localization_solver.cpp:1668
  KeyframeFrontend  <- LoopCam::on_flattened_images + LoopDetector::on_image_recv database work
All compute happens in libomniswarm_b200.so; these classes only marshal numpy arrays.
"""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import lib as _l


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class SuperPoint:
    """`inference(image) -> (keypoints [N,2] f32 (x,y) by descending confidence, descriptors [N,64])`."""

    def __init__(self, weights: np.ndarray, pca_comp: np.ndarray, pca_mean: np.ndarray, width: int, height: int,
                 thres: float = 0.015, max_num: int = 200, max_batch: int = 8):
        self._lib = _l.load()
        self.width, self.height, self.thres, self.max_num, self.max_batch = width, height, thres, max_num, max_batch
        w, pc, pm = _f32(weights).reshape(-1), _f32(pca_comp), _f32(pca_mean)
        assert pc.shape == (64, 256) and pm.shape == (256,)
        self._h = C.c_void_p()
        _l.check(self._lib.osb_superpoint_create(C.byref(self._h), _l.ptr(w), w.size, width, height, thres, max_num,
                                                 _l.ptr(pc), _l.ptr(pm), max_batch))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.osb_superpoint_destroy(self._h)
            self._h = None

    __del__ = close

    def _outputs(self, B):
        return (np.zeros(B, np.int32), np.zeros((B, self.max_num, 2), np.float32),
                np.zeros((B, self.max_num, 64), np.float32))

    def inference_batch(self, images: np.ndarray):
        """images [B,H,W] uint8 -> list of (kpts, desc)."""
        images = np.ascontiguousarray(images, dtype=np.uint8)
        # the reference asserts the image size (superpoint_tensorrt.cpp:122)
        assert images.ndim == 3 and images.shape[1:] == (self.height, self.width), \
            "Input image must have same size with network"
        B = images.shape[0]
        n, k, d = self._outputs(B)
        _l.check(self._lib.osb_superpoint_infer(self._h, _l.ptr(images), B, _l.ptr(n), _l.ptr(k), _l.ptr(d)))
        return [(k[b, :n[b]].copy(), d[b, :n[b]].copy()) for b in range(B)]

    def inference(self, image: np.ndarray):
        return self.inference_batch(image[None])[0]

    def postprocess(self, semi: np.ndarray, desc_nchw: np.ndarray):
        """parity hook: getKeyPoints + NMS2 + computeDescriptors on caller-supplied engine outputs."""
        semi, desc_nchw = _f32(semi), _f32(desc_nchw)
        if semi.ndim == 2:
            semi, desc_nchw = semi[None], desc_nchw[None]
        B = semi.shape[0]
        n, k, d = self._outputs(B)
        _l.check(self._lib.osb_superpoint_postprocess(self._h, _l.ptr(semi), _l.ptr(desc_nchw), B, _l.ptr(n),
                                                      _l.ptr(k), _l.ptr(d)))
        return [(k[b, :n[b]].copy(), d[b, :n[b]].copy()) for b in range(B)]

    LAYERS = ["conv1a", "conv1b+pool", "conv2a", "conv2b+pool", "conv3a", "conv3b+pool", "conv4a", "conv4b", "convPa",
              "convPb", "convDa", "convDb"]

    def layer_ms(self, images: np.ndarray) -> dict:
        """device time of every network layer for one batch (tensor-core path)."""
        _l.check(self._lib.osb_superpoint_set_profiling(self._h, 1))
        self.inference_batch(images)
        ms = np.zeros(12, np.float32)
        _l.check(self._lib.osb_superpoint_layer_ms(self._h, _l.ptr(ms), 12))
        _l.check(self._lib.osb_superpoint_set_profiling(self._h, 0))
        return {k: float(v) for k, v in zip(self.LAYERS, ms)}

    def read(self, what: str, image: int = 0) -> np.ndarray:
        H, W = self.height, self.width
        shapes = {"semi": (0, (H, W)), "desc": (1, (256, H // 8, W // 8)), "conf": (2, (self.max_num,)),
                  "survivors": (3, (H, W)), "counts": (4, (8,)), "fused1_cycles": (5, (16,))}
        code, shape = shapes[what]
        out = np.zeros(shape, np.float32)
        _l.check(self._lib.osb_superpoint_read(self._h, code, image, _l.ptr(out), out.size))
        return out


class NetVLAD:
    """`inference(image) -> [4096] f32` (mobilenetvlad_tensorrt.cpp:4-15)."""

    def __init__(self, weights: np.ndarray, width: int, height: int, max_batch: int = 4):
        self._lib = _l.load()
        self.width, self.height = width, height
        w = _f32(weights).reshape(-1)
        self._h = C.c_void_p()
        _l.check(self._lib.osb_netvlad_create(C.byref(self._h), _l.ptr(w), w.size, width, height, max_batch))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.osb_netvlad_destroy(self._h)
            self._h = None

    __del__ = close

    def inference_batch(self, images: np.ndarray) -> np.ndarray:
        images = np.ascontiguousarray(images, dtype=np.uint8)
        assert images.ndim == 3 and images.shape[1:] == (self.height, self.width)
        out = np.zeros((images.shape[0], _l.DEEP_DESC_SIZE), np.float32)
        _l.check(self._lib.osb_netvlad_infer(self._h, _l.ptr(images), images.shape[0], _l.ptr(out)))
        return out

    def inference(self, image: np.ndarray) -> np.ndarray:
        return self.inference_batch(image[None])[0]


class IndexFlatIP:
    """faiss::IndexFlatIP look-alike: `add(x)`, `search(q, k) -> (D, I)`, `ntotal`."""

    def __init__(self, d: int, capacity: int = 16384):
        self._lib = _l.load()
        self.d = d
        self._h = C.c_void_p()
        _l.check(self._lib.osb_db_create(C.byref(self._h), d, capacity))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.osb_db_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def ntotal(self) -> int:
        return int(self._lib.osb_db_size(self._h))

    def add(self, x: np.ndarray) -> int:
        x = _f32(x).reshape(-1, self.d)
        first = C.c_int64(-1)
        _l.check(self._lib.osb_db_add(self._h, x.shape[0], _l.ptr(x), C.byref(first)))
        return int(first.value)

    def search(self, q: np.ndarray, k: int):
        q = _f32(q).reshape(-1, self.d)
        D = np.zeros((q.shape[0], k), np.float32)
        I = np.zeros((q.shape[0], k), np.int64)
        _l.check(self._lib.osb_db_search(self._h, q.shape[0], _l.ptr(q), k, _l.ptr(D), _l.ptr(I)))
        return D, I

    def search_dev(self, q_ptr: int, nq: int, k: int, scores_ptr: int, ids_ptr: int, stream: int):
        """device pointers in / out, no synchronisation (osb_db_search_dev)."""
        _l.check(self._lib.osb_db_search_dev(self._h, nq, C.c_void_p(q_ptr), k, C.c_void_p(scores_ptr),
                                             C.c_void_p(ids_ptr), C.c_void_p(stream)))

    def add_dev(self, x_ptr: int, n: int, stream: int) -> int:
        first = C.c_int64(-1)
        _l.check(self._lib.osb_db_add_dev(self._h, n, C.c_void_p(x_ptr), C.byref(first), C.c_void_p(stream)))
        return int(first.value)

    def reset(self):
        _l.check(self._lib.osb_db_reset(self._h))


class BFMatcher:
    """cv::BFMatcher(cv::NORM_L2, crossCheck=True): `match(query, train) -> (queryIdx, trainIdx, distance)`."""

    def __init__(self, max_pairs: int = 8, max_n: int = 200, dim: int = 64):
        self._lib = _l.load()
        self.max_pairs, self.max_n, self.dim = max_pairs, max_n, dim
        self._h = C.c_void_p()
        _l.check(self._lib.osb_matcher_create(C.byref(self._h), max_pairs, max_n, dim))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.osb_matcher_destroy(self._h)
            self._h = None

    __del__ = close

    def match_batch(self, queries, trains):
        P = len(queries)
        q = np.zeros((P, self.max_n, self.dim), np.float32)
        t = np.zeros((P, self.max_n, self.dim), np.float32)
        nq = np.array([len(x) for x in queries], np.int32)
        nt = np.array([len(x) for x in trains], np.int32)
        for p in range(P):
            q[p, :nq[p]] = queries[p]
            t[p, :nt[p]] = trains[p]
        qi = np.zeros((P, self.max_n), np.int32); ti = np.zeros((P, self.max_n), np.int32)
        dist = np.zeros((P, self.max_n), np.float32); n = np.zeros(P, np.int32)
        _l.check(self._lib.osb_matcher_match(self._h, P, _l.ptr(q), _l.ptr(nq), _l.ptr(t), _l.ptr(nt), _l.ptr(qi),
                                             _l.ptr(ti), _l.ptr(dist), _l.ptr(n)))
        return [(qi[p, :n[p]].copy(), ti[p, :n[p]].copy(), dist[p, :n[p]].copy()) for p in range(P)]

    def match(self, query: np.ndarray, train: np.ndarray):
        return self.match_batch([query], [train])[0]


def homography_ransac(src_list, dst_list, thresh: float = 3.0, seed: int = 0):
    """cv::findHomography(old_2d, new_2d, CV_RANSAC, 3, mask) for several correspondence sets at once
    (loop_detector.cpp:589-598): src_list[i] = old_2d, dst_list[i] = new_2d, [n_i, 2] float32 ->
    list of (mask uint8 [n_i], n_inliers, winning hypothesis)."""
    lib = _l.load()
    n_pairs = len(src_list)
    max_n = max(1, max(len(a) for a in src_list))
    src = np.zeros((n_pairs, max_n, 2), np.float32); dst = np.zeros((n_pairs, max_n, 2), np.float32)
    n = np.zeros(n_pairs, np.int32)
    for i, (a, b) in enumerate(zip(src_list, dst_list)):
        n[i] = len(a)
        if len(a):
            src[i, :len(a)] = a; dst[i, :len(a)] = b
    mask = np.zeros((n_pairs, max_n), np.uint8); ninl = np.zeros(n_pairs, np.int32); win = np.zeros(n_pairs, np.int32)
    _l.check(lib.osb_homography_ransac(_l.ptr(src), _l.ptr(dst), _l.ptr(n), n_pairs, max_n, float(thresh), int(seed),
                                       _l.ptr(mask), _l.ptr(ninl), _l.ptr(win)))
    return [(mask[i, :n[i]].copy(), int(ninl[i]), int(win[i])) for i in range(n_pairs)]


class PoseGraphSolver:
    """Flat-array form of SwarmLocalizationSolver::solve_once: `solve(graph) -> (poses, summary)`."""

    def __init__(self, max_nodes: int = 4096, max_factors: int = 32768):
        self._lib = _l.load()
        self._h = C.c_void_p()
        _l.check(self._lib.osb_solver_create(C.byref(self._h), max_nodes, max_factors))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.osb_solver_destroy(self._h)
            self._h = None

    __del__ = close

    def default_options(self) -> _l.SolveOptions:
        o = _l.SolveOptions()
        self._lib.osb_solve_default_options(C.byref(o))
        return o

    @staticmethod
    def _arrays(g):
        return (np.ascontiguousarray(g["fixed"], np.uint8), np.ascontiguousarray(g["ftype"], np.int32),
                np.ascontiguousarray(g["ia"], np.int32), np.ascontiguousarray(g["ib"], np.int32),
                np.ascontiguousarray(g["payload"], np.float64), np.ascontiguousarray(g["huber"], np.uint8))

    def solve(self, g: dict, options: _l.SolveOptions | None = None, init: np.ndarray | None = None):
        fixed, ftype, ia, ib, payload, huber = self._arrays(g)
        poses = np.ascontiguousarray(g["init"] if init is None else init, np.float64).copy()
        opt = options if options is not None else self.default_options()
        summ = _l.SolveSummary()
        _l.check(self._lib.osb_solver_solve(self._h, poses.shape[0], _l.ptr(poses), _l.ptr(fixed), len(ftype),
                                            _l.ptr(ftype), _l.ptr(ia), _l.ptr(ib), _l.ptr(payload), _l.ptr(huber),
                                            C.byref(opt), C.byref(summ)))
        return poses, summ

    # ---- resident graph: the factor list stays on the device between solves ----
    def graph_clear(self):
        _l.check(self._lib.osb_solver_graph_clear(self._h))

    def graph_add_nodes(self, poses: np.ndarray, fixed: np.ndarray | None = None) -> int:
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 4)
        fx = None if fixed is None else np.ascontiguousarray(fixed, np.uint8)
        first = C.c_int32(-1)
        _l.check(self._lib.osb_solver_graph_add_nodes(self._h, poses.shape[0], _l.ptr(poses),
                                                      None if fx is None else _l.ptr(fx), C.byref(first)))
        return int(first.value)

    def graph_add_factors(self, ftype, ia, ib, payload, huber):
        ftype = np.ascontiguousarray(ftype, np.int32); ia = np.ascontiguousarray(ia, np.int32)
        ib = np.ascontiguousarray(ib, np.int32); huber = np.ascontiguousarray(huber, np.uint8)
        payload = np.ascontiguousarray(payload, np.float64)
        _l.check(self._lib.osb_solver_graph_add_factors(self._h, len(ftype), _l.ptr(ftype), _l.ptr(ia), _l.ptr(ib),
                                                        _l.ptr(payload), _l.ptr(huber)))

    def graph_set_fixed(self, node: int, fixed: bool = True):
        _l.check(self._lib.osb_solver_graph_set_fixed(self._h, node, int(fixed)))

    def graph_set_poses(self, first: int, poses: np.ndarray):
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 4)
        _l.check(self._lib.osb_solver_graph_set_poses(self._h, first, poses.shape[0], _l.ptr(poses)))

    def graph_get_poses(self, first: int = 0, n: int | None = None) -> np.ndarray:
        if n is None:
            n = self.graph_size()[0] - first
        out = np.zeros((n, 4), np.float64)
        _l.check(self._lib.osb_solver_graph_get_poses(self._h, first, n, _l.ptr(out)))
        return out

    def graph_size(self) -> tuple[int, int]:
        a, b = C.c_int32(0), C.c_int32(0)
        _l.check(self._lib.osb_solver_graph_size(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def graph_drop_oldest(self, n_nodes: int):
        _l.check(self._lib.osb_solver_graph_drop_oldest(self._h, n_nodes))

    def solve_resident(self, options: _l.SolveOptions | None = None) -> _l.SolveSummary:
        opt = options if options is not None else self.default_options()
        summ = _l.SolveSummary()
        _l.check(self._lib.osb_solver_solve_resident(self._h, C.byref(opt), C.byref(summ)))
        return summ

    def phase_cycles(self) -> dict:
        c = np.zeros(12, np.float64)
        _l.check(self._lib.osb_solver_phase_cycles(self._h, _l.ptr(c)))
        names = ["factor", "barrier", "node1", "reduce1", "node2", "reduce2", "cg_iterations", "kernel", "ctas",
                 "cluster", "j_in_smem", "threads"]
        d = dict(zip(names, c.tolist()))
        d["chain_preconditioner"] = float((int(d["j_in_smem"]) >> 1) & 1)
        d["inner_fp32"] = float((int(d["j_in_smem"]) >> 2) & 1)
        d["j_in_smem"] = float(int(d["j_in_smem"]) & 1)
        return d

    def chain_cycles(self) -> np.ndarray:
        c = np.zeros(128, np.float64)
        _l.check(self._lib.osb_solver_chain_cycles(self._h, _l.ptr(c)))
        return c.reshape(16, 8)

    @staticmethod
    def chain_plan(g: dict):
        """Host-only: the solver's internal node numbering (greedy maximum-weight path cover) ->
        (order [n] internal -> caller's node id, link [n] uint8)."""
        lib = _l.load()
        fixed = np.ascontiguousarray(g["fixed"], np.uint8)
        ftype = np.ascontiguousarray(g["ftype"], np.int32)
        ia = np.ascontiguousarray(g["ia"], np.int32); ib = np.ascontiguousarray(g["ib"], np.int32)
        payload = np.ascontiguousarray(g["payload"], np.float64)
        n = fixed.shape[0]
        order = np.zeros(n, np.int32); link = np.zeros(n, np.uint8)
        _l.check(lib.osb_solver_chain_plan(n, _l.ptr(fixed), len(ftype), _l.ptr(ftype), _l.ptr(ia), _l.ptr(ib),
                                           _l.ptr(payload), _l.ptr(order), _l.ptr(link)))
        return order, link

    def linearize(self, g: dict, poses: np.ndarray):
        _, ftype, ia, ib, payload, _ = self._arrays(g)
        poses = np.ascontiguousarray(poses, np.float64)
        m = len(ftype)
        r = np.zeros((m, 4)); Ja = np.zeros((m, 4, 4)); Jb = np.zeros((m, 4, 4))
        _l.check(self._lib.osb_solver_linearize(self._h, poses.shape[0], _l.ptr(poses), m, _l.ptr(ftype), _l.ptr(ia),
                                                _l.ptr(ib), _l.ptr(payload), _l.ptr(r), _l.ptr(Ja), _l.ptr(Jb)))
        return r, Ja, Jb


class KeyframeFrontend:
    """The per-keyframe pipeline (extract -> ingest -> query) on one GPU."""

    def __init__(self, sp_weights, pca_comp, pca_mean, nv_weights, width=640, height=480, n_dirs=4, max_num=200,
                 sp_thres=0.015, self_id=0, db_capacity=16384, inner_product_thres=0.3, init_mode_product_thres=0.2,
                 match_index_dist=5, query_dir=None, zero_bottom_quarter=True, accept_min_3d_pts=10,
                 geometric_filter=False, ransac_seed=0):
        self._lib = _l.load()
        cfg = _l.FrontendConfig()
        cfg.width, cfg.height, cfg.n_dirs, cfg.max_num = width, height, n_dirs, max_num
        cfg.sp_thres, cfg.self_id, cfg.db_capacity = sp_thres, self_id, db_capacity
        cfg.inner_product_thres, cfg.init_mode_product_thres = inner_product_thres, init_mode_product_thres
        cfg.match_index_dist = match_index_dist
        cfg.query_dir = (1 if n_dirs > 1 else 0) if query_dir is None else query_dir
        cfg.zero_bottom_quarter = int(zero_bottom_quarter)
        cfg.accept_min_3d_pts = accept_min_3d_pts
        cfg.geometric_filter, cfg.ransac_seed = int(geometric_filter), int(ransac_seed)
        self.cfg = cfg
        spw, nvw = _f32(sp_weights).reshape(-1), _f32(nv_weights).reshape(-1)
        pc, pm = _f32(pca_comp), _f32(pca_mean)
        self._h = C.c_void_p()
        _l.check(self._lib.osb_frontend_create(C.byref(self._h), C.byref(cfg), _l.ptr(spw), spw.size, _l.ptr(pc),
                                               _l.ptr(pm), _l.ptr(nvw), nvw.size))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.osb_frontend_destroy(self._h)
            self._h = None

    __del__ = close

    def process(self, images_up: np.ndarray, images_down: np.ndarray, msg_id: int):
        """HOST images [n_dirs,H,W] u8 -> (KeyframeRecord, LoopResult); one synchronisation."""
        up = np.ascontiguousarray(images_up, np.uint8); down = np.ascontiguousarray(images_down, np.uint8)
        rec, res = _l.KeyframeRecord(), _l.LoopResult()
        _l.check(self._lib.osb_frontend_process(self._h, _l.ptr(up), _l.ptr(down), msg_id, C.byref(rec), C.byref(res)))
        return rec, res

    def process_raw(self, up_ptr: int, down_ptr: int, msg_id: int, rec_ptr: int, res_ptr: int):
        """same, with raw HOST pointers (pinned buffers owned by the caller)."""
        _l.check(self._lib.osb_frontend_process(self._h, C.c_void_p(up_ptr), C.c_void_p(down_ptr), msg_id,
                                                C.c_void_p(rec_ptr), C.c_void_p(res_ptr)))

    def extract(self, up_ptr: int, down_ptr: int, msg_id: int, record_dev: int, stream: int, device_images=False):
        fn = self._lib.osb_frontend_extract_dev if device_images else self._lib.osb_frontend_extract
        _l.check(fn(self._h, C.c_void_p(up_ptr), C.c_void_p(down_ptr), msg_id, C.c_void_p(record_dev), C.c_void_p(stream)))

    def ingest(self, records_dev: int, n_records: int, skip: int, stream: int):
        _l.check(self._lib.osb_frontend_ingest(self._h, C.c_void_p(records_dev), n_records, skip, C.c_void_p(stream)))

    def ingest_own(self, record_dev: int, stream: int):
        """add_to_database of the record `extract` has just written (this drone's own)"""
        _l.check(self._lib.osb_frontend_ingest_own(self._h, C.c_void_p(record_dev), C.c_void_p(stream)))

    def query(self, record_dev: int, result_dev: int, stream: int, init_mode=False, nonkeyframe=False):
        _l.check(self._lib.osb_frontend_query(self._h, C.c_void_p(record_dev), int(init_mode), int(nonkeyframe),
                                              C.c_void_p(result_dev), C.c_void_p(stream)))

    def finish(self, stream: int):
        _l.check(self._lib.osb_frontend_finish(self._h, C.c_void_p(stream)))

    STAGES = ["superpoint_net(+keypoints beside descriptor head)", "descriptors", "netvlad_unhidden", "stereo_pack", "add_to_database", "db_scan",
              "rule_local_match"]

    def set_profiling(self, enable: bool):
        _l.check(self._lib.osb_frontend_set_profiling(self._h, int(enable)))

    def stage_ms(self) -> dict:
        ms = np.zeros(8, np.float32)
        _l.check(self._lib.osb_frontend_stage_ms(self._h, _l.ptr(ms)))
        return {k: float(v) for k, v in zip(self.STAGES, ms)}

    def set_cameras(self, intrinsics, left_extrinsics, right_extrinsics, triangle_thres=0.006):
        """stereo triangulation inside extract: the record then carries landmarks_3d / landmarks_flag (loop_cam.cpp:393-432)"""
        K = np.ascontiguousarray(intrinsics, np.float64)
        le, re = np.ascontiguousarray(left_extrinsics, np.float64), np.ascontiguousarray(right_extrinsics, np.float64)
        _l.check(self._lib.osb_frontend_set_cameras(self._h, _l.ptr(K), _l.ptr(le), _l.ptr(re), float(triangle_thres)))

    def set_drone_pose(self, pose_drone):
        p = np.ascontiguousarray(pose_drone, np.float64)
        _l.check(self._lib.osb_frontend_set_drone_pose(self._h, _l.ptr(p)))

    def db_size(self, remote=False) -> int:
        return int(self._lib.osb_frontend_db_size(self._h, int(remote)))

    def db_reset(self):
        _l.check(self._lib.osb_frontend_db_reset(self._h))

    def db_load(self, global_desc: np.ndarray, local_desc=None, n_kpts=None, remote=False):
        g = _f32(global_desc)
        ld = None if local_desc is None else _f32(local_desc)
        nk = None if n_kpts is None else np.ascontiguousarray(n_kpts, np.int32)
        _l.check(self._lib.osb_frontend_db_load(self._h, int(remote), g.shape[0], _l.ptr(g), _l.ptr(ld), _l.ptr(nk)))

    def db_set_geometry(self, first_row: int, kpts: np.ndarray, stereo_match: np.ndarray, remote=False):
        """landmarks_2d [n][max_num][2] and stereo_match [n][max_num] of rows put in with db_load"""
        k = _f32(kpts); sm = np.ascontiguousarray(stereo_match, np.int32)
        _l.check(self._lib.osb_frontend_db_set_geometry(self._h, int(remote), first_row, k.shape[0], _l.ptr(k), _l.ptr(sm)))


def stereo_lift(kp_up, kp_down, stereo_match, n_up, n_down, intrinsics, pose_up, pose_down, triangle_thres=0.006,
                accept_min_3d_pts=0):
    """loop_cam.cpp:393-432 for n_dirs directions: kp_* [n_dirs,max_n,2] f32, stereo_match [n_dirs,max_n] i32,
    poses [n_dirs,7] -> (pts3d [n_dirs,max_n,3] f32, flag_up, flag_down [n_dirs,max_n] u8)"""
    lib = _l.load()
    ku, kd = _f32(kp_up), _f32(kp_down)
    nd, mn = ku.shape[0], ku.shape[1]
    sm = np.ascontiguousarray(stereo_match, np.int32)
    nu, ndn = np.ascontiguousarray(n_up, np.int32), np.ascontiguousarray(n_down, np.int32)
    K = np.ascontiguousarray(intrinsics, np.float64)
    pu, pd = np.ascontiguousarray(pose_up, np.float64), np.ascontiguousarray(pose_down, np.float64)
    pts = np.zeros((nd, mn, 3), np.float32); fu = np.zeros((nd, mn), np.uint8); fd = np.zeros((nd, mn), np.uint8)
    _l.check(lib.osb_stereo_lift(_l.ptr(ku), _l.ptr(kd), _l.ptr(sm), _l.ptr(nu), _l.ptr(ndn), nd, mn, _l.ptr(K), _l.ptr(pu),
                                 _l.ptr(pd), float(triangle_thres), int(accept_min_3d_pts), _l.ptr(pts), _l.ptr(fu), _l.ptr(fd)))
    return pts, fu, fd


def depth_lift(kp, n, depth_mm, intrinsics, pose_cam, near=0.3, far=10.0, accept_min_3d_pts=0):
    """loop_cam.cpp:276-302: kp [n_dirs,max_n,2], depth_mm [n_dirs,H,W] u16, pose_cam [n_dirs,7] -> (pts3d, flag)"""
    lib = _l.load()
    k = _f32(kp)
    nd, mn = k.shape[0], k.shape[1]
    nn = np.ascontiguousarray(n, np.int32)
    dep = np.ascontiguousarray(depth_mm, np.uint16)
    K = np.ascontiguousarray(intrinsics, np.float64)
    pc = np.ascontiguousarray(pose_cam, np.float64)
    pts = np.zeros((nd, mn, 3), np.float32); fl = np.zeros((nd, mn), np.uint8)
    _l.check(lib.osb_depth_lift(_l.ptr(k), _l.ptr(nn), nd, mn, _l.ptr(dep), dep.shape[1], dep.shape[2], _l.ptr(K), _l.ptr(pc),
                                float(near), float(far), int(accept_min_3d_pts), _l.ptr(pts), _l.ptr(fl)))
    return pts, fl


def pnp_ransac(cases, max_n: int | None = None):
    """LoopDetector::compute_relative_pose + check_loop_odometry_consistency (loop_detector.cpp:294-413) for a batch of loop
    candidates.  cases: list of dicts with X [n,3], uv [n,2] and the osb_pnp_params fields (prior, extrinsic, drone_pose_now,
    drone_pose_old [7]; iterations, thresh, seed, is_4dof, min_loop_num, rperr_thres, accept_loop_yaw_rad, max_loop_dis;
    optional same_drone, odom_rel [7], cov [6,6], odometry_consistency_threshold) -> list of (mask uint8 [n], PnpResult)."""
    lib = _l.load()
    nc = len(cases)
    if max_n is None:
        max_n = max(1, max(len(c["X"]) for c in cases))
    p3 = np.zeros((nc, max_n, 3), np.float32); p2 = np.zeros((nc, max_n, 2), np.float32)
    n = np.zeros(nc, np.int32)
    prm = (_l.PnpParams * nc)()
    for i, c in enumerate(cases):
        k = len(c["X"]); n[i] = k
        if k:
            p3[i, :k] = c["X"]; p2[i, :k] = c["uv"]
        p = prm[i]
        p.iterations, p.reproj_thresh, p.seed = int(c.get("iterations", 100)), float(c.get("thresh", 3.0)), int(c.get("seed", 0))
        p.is_4dof, p.min_loop_num, p.same_drone = int(c.get("is_4dof", 1)), int(c.get("min_loop_num", 15)), int(c.get("same_drone", 0))
        p.rperr_thres, p.accept_loop_yaw_rad = float(c.get("rperr_thres", 0.1)), float(c.get("accept_loop_yaw_rad", 0.8))
        p.max_loop_dis = float(c.get("max_loop_dis", 5.0))
        p.odometry_consistency_threshold = float(c.get("odometry_consistency_threshold", 10.0))
        for name in ("prior", "extrinsic", "drone_pose_now", "drone_pose_old"):
            getattr(p, name)[:] = [float(x) for x in c[name]]
        p.odom_rel[:] = [float(x) for x in c.get("odom_rel", [0, 0, 0, 1, 0, 0, 0])]
        p.odom_edge_cov[:] = [float(x) for x in np.asarray(c.get("cov", np.eye(6)), np.float64).reshape(-1)]
    mask = np.zeros((nc, max_n), np.uint8)
    res = (_l.PnpResult * nc)()
    _l.check(lib.osb_pnp_ransac(_l.ptr(p3), _l.ptr(p2), _l.ptr(n), nc, max_n, prm, _l.ptr(mask), res))
    return [(mask[i, :n[i]].copy(), res[i]) for i in range(nc)]


def pcm_outlier_rejection(edges, pcm_thres: float, odom_pos_cov_per_m: float, odom_ang_cov_per_m: float,
                          want_matrices: bool = False):
    """SwarmLocalOutlierRejection::OutlierRejectionLoopEdgesPCM (swarm_outlier_rejection.cpp:173-297) for the loop edges of
    one drone pair: `edges` = list of dicts (id_a, id_b, rel [7], cov [6,6], odom_a [7], odom_b [7], len_a, len_b) in
    insertion order -> indices of the kept loops in maxCliqueHeu's order (+ adjacency and smd matrices on request)."""
    lib = _l.load()
    n = len(edges)
    arr = (_l.LoopEdge * n)()
    for i, e in enumerate(edges):
        a = arr[i]
        a.id_a, a.id_b, a.len_a, a.len_b = int(e["id_a"]), int(e["id_b"]), float(e["len_a"]), float(e["len_b"])
        a.rel_pose[:] = [float(x) for x in e["rel"]]
        a.cov[:] = [float(x) for x in np.asarray(e["cov"], np.float64).reshape(-1)]
        a.odom_a[:] = [float(x) for x in e["odom_a"]]
        a.odom_b[:] = [float(x) for x in e["odom_b"]]
    clique = np.zeros(n, np.int32)
    size = C.c_int32(0)
    adj = np.zeros((n, n), np.uint8) if want_matrices else None
    smd = np.zeros((n, n), np.float64) if want_matrices else None
    _l.check(lib.osb_pcm(arr, n, float(pcm_thres), float(odom_pos_cov_per_m), float(odom_ang_cov_per_m), _l.ptr(clique),
                         C.byref(size), _l.ptr(adj), _l.ptr(smd)))
    out = clique[:size.value].copy()
    return (out, adj, smd) if want_matrices else out


class Swarm:
    """The swarm-wide keyframe exchange behind the C ABI (osb_swarm_*): one ncclAllGather of the fixed-size keyframe
    record per round; replaces LoopNet::broadcast_fisheye_desc / image_desc_callback (loop_net.cpp:20-120,142-172).
    `unique_id()` on rank 0, hand the 128 bytes to the other ranks, then `Swarm(id, rank, world)` everywhere."""

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * _l.SWARM_ID_BYTES)()
        _l.check(_l.load().osb_swarm_unique_id(buf))
        return bytes(buf)

    def __init__(self, uid: bytes | None, rank: int, world: int):
        self._lib = _l.load()
        self._h = C.c_void_p()
        idbuf = None
        if uid is not None:
            assert len(uid) == _l.SWARM_ID_BYTES
            idbuf = (C.c_uint8 * _l.SWARM_ID_BYTES).from_buffer_copy(uid)
        _l.check(self._lib.osb_swarm_init(C.byref(self._h), idbuf, rank, world))
        self.rank, self.world = rank, world
        self.transport = "p2p copy engines" if self._lib.osb_swarm_transport(self._h) == 1 else "nccl"

    def close(self):
        if getattr(self, "_h", None):
            self._lib.osb_swarm_destroy(self._h)
            self._h = None

    __del__ = close

    def exchange(self, record_dev: int, gathered_dev: int, stream: int):
        """this rank's record -> gathered[world] in rank order, enqueued on `stream` (osb_swarm_exchange)"""
        _l.check(self._lib.osb_swarm_exchange(self._h, C.c_void_p(record_dev), C.c_void_p(gathered_dev), C.c_void_p(stream)))

    def exchange_async(self, record_dev: int, gathered_dev: int, stream: int):
        """the same on the handle's own stream, behind an event of `stream` (osb_swarm_exchange_async)"""
        _l.check(self._lib.osb_swarm_exchange_async(self._h, C.c_void_p(record_dev), C.c_void_p(gathered_dev), C.c_void_p(stream)))

    def wait(self, stream: int):
        """`stream` waits for the last exchange_async (osb_swarm_wait)"""
        _l.check(self._lib.osb_swarm_wait(self._h, C.c_void_p(stream)))


def launch_count() -> int:
    return int(_l.load().osb_launch_count())
