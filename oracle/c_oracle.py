"""TEST INFRASTRUCTURE — ctypes loader for the C restatement (oracle/d3f_oracle.c)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Iterable, Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, '_build', 'libd3f_oracle.so')
_lib = None


def build() -> str:
    src = os.path.join(HERE, 'd3f_oracle.c')
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(['make', '-C', HERE, '-s'], check=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        try:
            build()
        except Exception:
            if not os.path.exists(LIB):
                raise
        _lib = C.CDLL(LIB)
        _lib.d3f_oracle_eval.restype = C.c_int
        _lib.d3f_oracle_threads.restype = C.c_int
    return _lib


def threads() -> int:
    return int(load().d3f_oracle_threads())


def field_eval(pts, pose, K, depth, H, W, maps: Optional[Dict[str, np.ndarray]] = None,
               return_names: Iterable[str] = (), mu: float = 0.02, eval_dist: bool = False,
               return_inter: bool = False) -> Dict[str, np.ndarray]:
    """Same signature and results layout as oracle.field_oracle.field_eval."""
    lib = load()
    maps = maps or {}
    names = [] if eval_dist else list(return_names)
    pts = np.ascontiguousarray(pts, np.float32)
    pose = np.ascontiguousarray(pose, np.float32)
    K = np.ascontiguousarray(K, np.float32)
    depth = np.ascontiguousarray(depth, np.float32)
    n, V = pts.shape[0], pose.shape[0]
    vols = [np.ascontiguousarray(maps[k]) for k in names]
    for v in vols:
        assert v.dtype in (np.float32, np.uint8) and v.ndim == 4 and v.shape[0] == V
    dist = np.empty(n, np.float32)
    valid = np.empty(n, np.uint8)
    outs = [np.empty((n, v.shape[3]), np.float32) for v in vols]
    inters = [np.empty((V, n, v.shape[3]), np.float32) for v in vols] if return_inter else []
    nk = len(vols)
    vp = C.c_void_p

    def parr(arrs):
        a = (vp * max(len(arrs), 1))()
        for i, x in enumerate(arrs):
            a[i] = x.ctypes.data
        return a

    def iarr(vals):
        return (C.c_int * max(len(vals), 1))(*vals)

    rc = lib.d3f_oracle_eval(C.c_int(V), C.c_int(H), C.c_int(W), vp(pose.ctypes.data), vp(K.ctypes.data),
                             vp(depth.ctypes.data), vp(pts.ctypes.data), C.c_int64(n), C.c_int(nk), parr(vols),
                             iarr([0 if v.dtype == np.float32 else 1 for v in vols]),
                             iarr([v.shape[1] for v in vols]), iarr([v.shape[2] for v in vols]),
                             iarr([v.shape[3] for v in vols]), C.c_float(mu), C.c_int(1 if eval_dist else 0),
                             vp(dist.ctypes.data), vp(valid.ctypes.data), parr(outs),
                             parr(inters) if return_inter else None)
    assert rc == 0, rc
    res = {'dist': dist, 'valid_mask': valid.astype(bool)}
    for i, k in enumerate(names):
        res[k] = outs[i]
        if return_inter:
            res[k + '_inter'] = inters[i]
    return res
