"""ORACLE (test infrastructure): ctypes binding of oracle/c/nms2_ref.c, the plain-C restatement of getKeyPoints + NMS2
(superpoint_tensorrt.cpp:164-189, 237-310).  `get_keypoints` has the signature of frontend_ref.get_keypoints."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
_PATH = os.path.join(_DIR, "libnms2_ref.so")
_lib = None


def build() -> str:
    r = subprocess.run(["make", "-C", _DIR], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building oracle/c failed:\n" + r.stdout + r.stderr)
    return _PATH


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            build()
        _lib = C.CDLL(_PATH)
        _lib.osb_ref_get_keypoints.restype = C.c_int
        _lib.osb_ref_get_keypoints.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p,
                                               C.c_void_p, C.POINTER(C.c_int)]
    return _lib


def get_keypoints(prob: np.ndarray, thres: float, max_num: int, dist_thresh: int = 4):
    """prob [H,W] f32 -> (kpts [n,2] f32 (x,y) by descending confidence, conf [n])."""
    lib = load()
    prob = np.ascontiguousarray(prob, np.float32)
    H, W = prob.shape
    k = np.zeros((max_num, 2), np.float32)
    c = np.zeros(max_num, np.float32)
    ncand = C.c_int(0)
    n = lib.osb_ref_get_keypoints(prob.ctypes.data, H, W, C.c_float(thres), max_num, dist_thresh, k.ctypes.data,
                                  c.ctypes.data, C.byref(ncand))
    if n < 0:
        raise MemoryError("osb_ref_get_keypoints")
    return k[:n].copy(), c[:n].copy()
