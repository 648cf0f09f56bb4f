"""ctypes wrapper of oracle/ba_ref_c.c (single-threaded C restatement of the localBA solve path).
TEST / BASELINE INFRASTRUCTURE ONLY - see oracle/ba_ref.py for the reference file:line map."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libov2oracle_ba.so"
_lib = None


class _Pb(C.Structure):
    _fields_ = [("ncam", C.c_int), ("npts", C.c_int), ("nobs", C.c_int), ("K", C.c_void_p), ("pose", C.c_void_p),
                ("pose_const", C.c_void_p), ("lm_anchor_cam", C.c_void_p), ("lm_anchor_px", C.c_void_p),
                ("lm_invdepth", C.c_void_p), ("obs_cam", C.c_void_p), ("obs_lm", C.c_void_p), ("obs_px", C.c_void_p)]


def build(force: bool = False) -> Path:
    src = _HERE / "ba_ref_c.c"
    _SO.parent.mkdir(exist_ok=True)
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["gcc", "-O2", "-march=native", "-shared", "-fPIC", "-o", str(_SO), str(src), "-lm"])
    return _SO


def local_ba(pb: dict, max_iters_robust=5, max_iters_refine=10, huber_th=5.9915, function_tolerance=1e-3,
             use_robust=True, apply_l2_after_robust=True):
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        _lib.ov2_oracle_local_ba.argtypes = [C.POINTER(_Pb), C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                                             C.c_void_p, C.c_void_p]
    keep = {k: np.ascontiguousarray(pb[k]) for k in ("K", "pose", "pose_const", "lm_anchor_cam", "lm_anchor_px",
                                                     "lm_invdepth", "obs_cam", "obs_lm", "obs_px")}
    p = _Pb(len(keep["pose"]), len(keep["lm_invdepth"]), len(keep["obs_cam"]),
            *[keep[k].ctypes.data for k in ("K", "pose", "pose_const", "lm_anchor_cam", "lm_anchor_px", "lm_invdepth",
                                            "obs_cam", "obs_lm", "obs_px")])
    res = np.zeros(8)
    flags = np.zeros(len(keep["obs_cam"]), np.uint8)
    _lib.ov2_oracle_local_ba(C.byref(p), int(max_iters_robust), int(max_iters_refine), float(huber_th),
                             float(function_tolerance), int(bool(use_robust)), int(bool(apply_l2_after_robust)),
                             res.ctypes.data, flags.ctypes.data)
    pb["pose"][...] = keep["pose"]
    pb["lm_invdepth"][...] = keep["lm_invdepth"]
    term = {0: "CONVERGENCE", 1: "NO_CONVERGENCE", 2: "FAILURE"}[int(res[6])]
    return dict(iters_robust=int(res[0]), iters_refine=int(res[1]), initial_cost=res[2], final_cost=res[3],
                n_outliers_first=int(res[4]), n_outliers_second=int(res[5]), termination=term, flags=flags)
