"""ORACLE (test infrastructure, NOT product code) -- the REFERENCE's own `FMC::maxCliqueHeu`, compiled from its sources in
place (oracle/ref_build/Makefile -> oracle/_ref/libfmc_ref.so) and called through ctypes.  Used to pin the restatement in
oracle/pcm_ref.py (`max_clique_heu`), which is what the CUDA path is compared with."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libfmc_ref.so")
REFERENCE = "/root/reference"


def build() -> bool:
    """compile oracle/_ref/libfmc_ref.so when the reference tree is present; -> whether the library exists afterwards"""
    if os.path.isdir(REFERENCE):
        subprocess.run(["make", "-C", os.path.join(_HERE, "ref_build")], check=True, capture_output=True)
    return os.path.exists(_SO)


def available() -> bool:
    return os.path.exists(_SO)


def max_clique_heu(adj: np.ndarray):
    """adjacency matrix (symmetric 0/1, zero diagonal) -> (clique vertex list in the reference's order, its size); neighbour lists
    ascending, as OutlierRejectionLoopEdgesPCM fills pcm_graph"""
    lib = C.CDLL(_SO)
    lib.fmc_max_clique_heu.restype = C.c_int
    n = adj.shape[0]
    nbr = [np.nonzero(adj[v])[0].astype(np.int32) for v in range(n)]
    offs = np.zeros(n + 1, np.int32)
    offs[1:] = np.cumsum([len(x) for x in nbr])
    edges = np.concatenate(nbr).astype(np.int32) if offs[-1] else np.zeros(1, np.int32)
    out = np.zeros(max(n, 1), np.int32)
    size = lib.fmc_max_clique_heu(n, offs.ctypes.data_as(C.c_void_p), edges.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert size >= -1, "maxCliqueHeu returned a size that differs from its clique list"
    return out[:max(size, 0)].tolist(), size
