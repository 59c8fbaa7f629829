"""ctypes binding of csrc/libomniswarm_b200.so (the C ABI declared in include/omniswarm_b200.h).

There is no CPU fallback: `load()` raises if the library has not been built, and every `create`
returns OSB_ERR_NO_DEVICE without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libomniswarm_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_NO_DEVICE = 0, 1, 2, 3, 4
MAX_DIRS, MAX_KPTS, FEATURE_DESC_SIZE, DEEP_DESC_SIZE = 4, 200, 64, 4096
REMOTE_MAGIN_NUMBER = 1000000
SWARM_ID_BYTES = 128
PAYLOAD_LEN = 24


class OsbError(RuntimeError):
    def __init__(self, status, text):
        super().__init__(f"libomniswarm_b200 status {status}: {text}")
        self.status = status


class SolveOptions(C.Structure):
    _fields_ = [("max_iterations", C.c_int32), ("max_pcg_iterations", C.c_int32), ("max_time_s", C.c_double),
                ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double),
                ("parameter_tolerance", C.c_double), ("pcg_tolerance", C.c_double),
                ("initial_trust_radius", C.c_double), ("preconditioner", C.c_int32), ("inner_precision", C.c_int32)]


class SolveSummary(C.Structure):
    _fields_ = [("initial_cost", C.c_double), ("final_cost", C.c_double), ("solve_ms", C.c_double),
                ("iterations", C.c_int32), ("pcg_iterations", C.c_int32), ("n_residuals", C.c_int32),
                ("termination", C.c_int32)]


class KeyframeRecord(C.Structure):
    _fields_ = [("drone_id", C.c_int32), ("msg_id", C.c_int32), ("n_dirs", C.c_int32), ("reserved", C.c_int32),
                ("n_kpts", C.c_int32 * MAX_DIRS), ("n_kpts_down", C.c_int32 * MAX_DIRS),
                ("global_desc", (C.c_float * DEEP_DESC_SIZE) * MAX_DIRS),
                ("local_desc", ((C.c_float * FEATURE_DESC_SIZE) * MAX_KPTS) * MAX_DIRS),
                ("kpts", ((C.c_float * 2) * MAX_KPTS) * MAX_DIRS),
                ("stereo_match", (C.c_int32 * MAX_KPTS) * MAX_DIRS),
                ("landmarks_3d", ((C.c_float * 3) * MAX_KPTS) * MAX_DIRS),
                ("landmarks_flag", (C.c_int32 * MAX_KPTS) * MAX_DIRS)]


class LoopResult(C.Structure):
    _fields_ = [("hit_id", C.c_int32), ("hit_dir", C.c_int32), ("hit_score", C.c_float), ("accepted", C.c_int32),
                ("swapped", C.c_int32), ("hit_msg_id", C.c_int32), ("hit_drone_id", C.c_int32), ("dir_new", C.c_int32 * MAX_DIRS), ("dir_old", C.c_int32 * MAX_DIRS),
                ("n_matches", C.c_int32 * MAX_DIRS), ("match_new", (C.c_int32 * MAX_KPTS) * MAX_DIRS),
                ("match_old", (C.c_int32 * MAX_KPTS) * MAX_DIRS),
                ("geo_valid", C.c_int32 * MAX_DIRS), ("n_geo", C.c_int32 * MAX_DIRS),
                ("geo_new", (C.c_int32 * MAX_KPTS) * MAX_DIRS), ("geo_old", (C.c_int32 * MAX_KPTS) * MAX_DIRS)]


class LoopEdge(C.Structure):
    _fields_ = [("id_a", C.c_int32), ("id_b", C.c_int32), ("rel_pose", C.c_double * 7), ("cov", C.c_double * 36),
                ("odom_a", C.c_double * 7), ("odom_b", C.c_double * 7), ("len_a", C.c_double), ("len_b", C.c_double)]


class PnpParams(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("reproj_thresh", C.c_float), ("seed", C.c_uint32), ("is_4dof", C.c_int32),
                ("min_loop_num", C.c_int32), ("same_drone", C.c_int32), ("rperr_thres", C.c_double),
                ("accept_loop_yaw_rad", C.c_double), ("max_loop_dis", C.c_double),
                ("odometry_consistency_threshold", C.c_double), ("prior", C.c_double * 7), ("extrinsic", C.c_double * 7),
                ("drone_pose_now", C.c_double * 7), ("drone_pose_old", C.c_double * 7), ("odom_rel", C.c_double * 7),
                ("odom_edge_cov", C.c_double * 36)]


class PnpResult(C.Structure):
    _fields_ = [("pnp_success", C.c_int32), ("n_inliers", C.c_int32), ("winner", C.c_int32), ("verified", C.c_int32),
                ("odometry_consistent", C.c_int32), ("reserved", C.c_int32), ("rperr", C.c_double), ("md", C.c_double),
                ("pose_cam", C.c_double * 7), ("dp_old_to_new", C.c_double * 4)]


class FrontendConfig(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("n_dirs", C.c_int32), ("max_num", C.c_int32),
                ("sp_thres", C.c_float), ("self_id", C.c_int32), ("db_capacity", C.c_int32),
                ("inner_product_thres", C.c_double), ("init_mode_product_thres", C.c_double),
                ("match_index_dist", C.c_int32), ("query_dir", C.c_int32), ("zero_bottom_quarter", C.c_int32),
                ("accept_min_3d_pts", C.c_int32), ("geometric_filter", C.c_int32), ("ransac_seed", C.c_int32)]


RECORD_BYTES = C.sizeof(KeyframeRecord)
RESULT_BYTES = C.sizeof(LoopResult)

_P = C.c_void_p
_SIG = {
    "osb_last_error": (C.c_char_p, []),
    "osb_version": (C.c_char_p, []),
    "osb_device_count": (C.c_int, []),
    "osb_launch_count": (C.c_int64, []),
    "osb_set_sm_budget": (None, [C.c_int]),
    "osb_superpoint_create": (C.c_int, [C.POINTER(_P), _P, C.c_size_t, C.c_int, C.c_int, C.c_float, C.c_int, _P, _P, C.c_int]),
    "osb_superpoint_destroy": (C.c_int, [_P]),
    "osb_superpoint_infer": (C.c_int, [_P, _P, C.c_int, _P, _P, _P]),
    "osb_superpoint_infer_dev": (C.c_int, [_P, _P, C.c_int, _P, _P, _P, _P]),
    "osb_superpoint_postprocess": (C.c_int, [_P, _P, _P, C.c_int, _P, _P, _P]),
    "osb_superpoint_read": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_size_t]),
    "osb_superpoint_set_profiling": (C.c_int, [_P, C.c_int]),
    "osb_superpoint_layer_ms": (C.c_int, [_P, _P, C.c_int]),
    "osb_netvlad_create": (C.c_int, [C.POINTER(_P), _P, C.c_size_t, C.c_int, C.c_int, C.c_int]),
    "osb_netvlad_destroy": (C.c_int, [_P]),
    "osb_netvlad_infer": (C.c_int, [_P, _P, C.c_int, _P]),
    "osb_netvlad_infer_dev": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "osb_db_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int64]),
    "osb_db_destroy": (C.c_int, [_P]),
    "osb_db_add": (C.c_int, [_P, C.c_int64, _P, C.POINTER(C.c_int64)]),
    "osb_db_add_dev": (C.c_int, [_P, C.c_int64, _P, C.POINTER(C.c_int64), _P]),
    "osb_db_search": (C.c_int, [_P, C.c_int64, _P, C.c_int, _P, _P]),
    "osb_db_search_dev": (C.c_int, [_P, C.c_int64, _P, C.c_int, _P, _P, _P]),
    "osb_topk_merge_dev": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P]),
    "osb_db_size": (C.c_int64, [_P]),
    "osb_db_reset": (C.c_int, [_P]),
    "osb_homography_ransac": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_float, C.c_uint32, _P, _P, _P]),
    "osb_homography_ransac_dev": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_float, C.c_uint32, _P, _P, _P, _P]),
    "osb_matcher_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int]),
    "osb_matcher_destroy": (C.c_int, [_P]),
    "osb_matcher_match": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P]),
    "osb_matcher_match_dev": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "osb_solve_default_options": (None, [C.POINTER(SolveOptions)]),
    "osb_solver_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int]),
    "osb_solver_destroy": (C.c_int, [_P]),
    "osb_solver_solve": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, _P, _P, _P, _P, _P, C.POINTER(SolveOptions), C.POINTER(SolveSummary)]),
    "osb_solver_graph_clear": (C.c_int, [_P]),
    "osb_solver_graph_add_nodes": (C.c_int, [_P, C.c_int, _P, _P, C.POINTER(C.c_int32)]),
    "osb_solver_graph_add_factors": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P]),
    "osb_solver_graph_set_fixed": (C.c_int, [_P, C.c_int, C.c_int]),
    "osb_solver_graph_set_poses": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "osb_solver_graph_get_poses": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "osb_solver_graph_size": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "osb_solver_graph_drop_oldest": (C.c_int, [_P, C.c_int]),
    "osb_solver_solve_resident": (C.c_int, [_P, _P, _P]),
    "osb_solver_phase_cycles": (C.c_int, [_P, _P]),
    "osb_solver_chain_cycles": (C.c_int, [_P, _P]),
    "osb_solver_chain_plan": (C.c_int, [C.c_int, _P, C.c_int, _P, _P, _P, _P, _P, _P]),
    "osb_solver_linearize": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, _P, _P, _P, _P, _P, _P]),
    "osb_frontend_create": (C.c_int, [C.POINTER(_P), C.POINTER(FrontendConfig), _P, C.c_size_t, _P, _P, _P, C.c_size_t]),
    "osb_frontend_destroy": (C.c_int, [_P]),
    "osb_frontend_extract": (C.c_int, [_P, _P, _P, C.c_int32, _P, _P]),
    "osb_frontend_extract_dev": (C.c_int, [_P, _P, _P, C.c_int32, _P, _P]),
    "osb_frontend_ingest": (C.c_int, [_P, _P, C.c_int, C.c_int, _P]),
    "osb_frontend_ingest_own": (C.c_int, [_P, _P, _P]),
    "osb_frontend_query": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P]),
    "osb_frontend_process": (C.c_int, [_P, _P, _P, C.c_int32, _P, _P]),
    "osb_frontend_finish": (C.c_int, [_P, _P]),
    "osb_frontend_set_profiling": (C.c_int, [_P, C.c_int]),
    "osb_frontend_stage_ms": (C.c_int, [_P, _P]),
    "osb_frontend_db_size": (C.c_int64, [_P, C.c_int]),
    "osb_frontend_db_reset": (C.c_int, [_P]),
    "osb_frontend_db_load": (C.c_int, [_P, C.c_int, C.c_int64, _P, _P, _P]),
    "osb_frontend_db_set_geometry": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int64, _P, _P]),
    "osb_stereo_lift": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, _P, _P, _P, C.c_double, C.c_int, _P, _P, _P]),
    "osb_stereo_lift_dev": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, _P, _P, _P, C.c_double, C.c_int, _P, _P, _P, _P]),
    "osb_depth_lift": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, C.c_int, C.c_int, _P, _P, C.c_double, C.c_double, C.c_int, _P, _P]),
    "osb_depth_lift_dev": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, C.c_int, C.c_int, _P, _P, C.c_double, C.c_double, C.c_int, _P, _P, _P]),
    "osb_frontend_set_cameras": (C.c_int, [_P, _P, _P, _P, C.c_double]),
    "osb_frontend_set_drone_pose": (C.c_int, [_P, _P]),
    "osb_pnp_ransac": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P, _P, _P]),
    "osb_pnp_ransac_dev": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P, _P, _P, _P]),
    "osb_pcm": (C.c_int, [_P, C.c_int, C.c_double, C.c_double, C.c_double, _P, _P, _P, _P]),
    "osb_pcm_dev": (C.c_int, [_P, C.c_int, C.c_double, C.c_double, C.c_double, _P, _P, _P, _P, _P]),
    "osb_swarm_unique_id": (C.c_int, [_P]),
    "osb_swarm_init": (C.c_int, [C.POINTER(_P), _P, C.c_int, C.c_int]),
    "osb_swarm_destroy": (C.c_int, [_P]),
    "osb_swarm_exchange": (C.c_int, [_P, _P, _P, _P]),
    "osb_swarm_exchange_async": (C.c_int, [_P, _P, _P, _P]),
    "osb_swarm_wait": (C.c_int, [_P, _P]),
    "osb_swarm_rank": (C.c_int, [_P]),
    "osb_swarm_transport": (C.c_int, [_P]),
    "osb_swarm_world": (C.c_int, [_P]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into csrc/libomniswarm_b200.so (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libomniswarm_b200.so failed")
    return LIB_PATH


def exported_symbols():
    return sorted(_SIG)


def load():
    """Load the shared library and attach the signatures.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIG.items():
        fn = getattr(lib, name)          # AttributeError here = symbol missing from the build
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int):
    if status != OK:
        raise OsbError(status, load().osb_last_error().decode(errors="replace"))


def ptr(a):
    """numpy array / ctypes object / int (device pointer) -> c_void_p"""
    import numpy as np
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
        return a.ctypes.data_as(C.c_void_p)
    return C.cast(C.byref(a), C.c_void_p)
