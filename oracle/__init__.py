"""CPU oracle for the Gaussian-splatting rasterizer hot path -- TEST INFRASTRUCTURE ONLY.

PARITY PARTLY PINNED: the reference's CUDA rasterizer (`submodules/diff-gaussian-rasterization`, MonoGS
"-w-pose" fork) and `simple-knn` are absent from /root/reference (`.MISSING_LARGE_BLOBS:1`), and the
reference holds no tests or golden vectors, so the rasterizer's internal arithmetic
is UNPINNED: the C files in this directory restate the published algorithm;
`tests/test_oracle_autograd.py` checks the analytic backward against float64 autograd of the forward.
Pinned by RUNNING reference-held Python (utils/pose_utils.py, camera_utils.py, slam_utils.py ->
tests/golden/reference_pin.npz): the pose-gradient conventions and the tracking chain around the rasterizer
(tests/test_reference_pin.py).  `oracle/build_ref.py` builds the reference's own CUDA into oracle/_ref/ if its
sources ever appear under /root/reference (today: absent).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py` (cpu_baseline / `--impl reference`) may import
this package.  The product (`lvd_gs-slam_b200/`) never does and has no CPU fallback.

numpy in, numpy out; thin ctypes binding over `oracle/_build/liboracle.so` (built by `oracle/Makefile`).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

FLAG_EXACT_PP = 1
FLAG_OPACITY_GRAD = 2
TILE = 16


def build(force: bool = False) -> str:
    """Compile the C oracle with gcc (see oracle/Makefile)."""
    srcs = [os.path.join(_HERE, f) for f in ("raster_oracle.c", "knn_oracle.c")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oracle_count_instances.restype = C.c_int64
    return _lib


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_num_threads(n: int) -> int:
    """OpenMP threads the oracle uses from now on (torchrun sets OMP_NUM_THREADS=1 in every rank's environment)."""
    lib().oracle_set_num_threads(C.c_int(int(n)))
    return num_threads()


def _f(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def tile_grid(W, H):
    return (W + TILE - 1) // TILE, (H + TILE - 1) // TILE


def rasterize_forward(means3D, opacities, scales=None, rotations=None, shs=None, colors_precomp=None,
                      cov3D_precomp=None, *, viewmatrix, projmatrix, campos, bg, W, H, tanfovx, tanfovy,
                      sh_degree=0, scale_modifier=1.0, want_margin=True):
    """Forward pass; returns every intermediate the parity tests compare (dict of numpy arrays)."""
    L = lib()
    means3D = _f(means3D); P = means3D.shape[0]
    opacities = _f(opacities).reshape(-1)
    scales = _f(scales); rotations = _f(rotations); shs = _f(shs)
    colors_precomp = _f(colors_precomp); cov3D_precomp = _f(cov3D_precomp)
    view = _f(viewmatrix).reshape(-1); proj = _f(projmatrix).reshape(-1)
    campos = _f(campos).reshape(-1); bg = _f(bg).reshape(-1)
    M = 0 if shs is None else shs.shape[1]
    o = dict(P=P, W=W, H=H, M=M, D=sh_degree)
    o["radii"] = np.zeros(P, np.int32); o["means2D"] = np.zeros((P, 2), np.float32)
    o["depths"] = np.zeros(P, np.float32); o["cov3D"] = np.zeros((P, 6), np.float32)
    o["conic_opacity"] = np.zeros((P, 4), np.float32); o["rgb"] = np.zeros((P, 3), np.float32)
    o["clamped"] = np.zeros(P, np.uint8); o["rect"] = np.zeros((P, 4), np.int32)
    o["tiles_touched"] = np.zeros(P, np.uint32)
    L.oracle_preprocess(C.c_int(P), C.c_int(sh_degree), C.c_int(M), _p(means3D), _p(scales), _p(rotations),
                        _p(opacities), _p(shs), _p(colors_precomp), _p(cov3D_precomp), C.c_float(scale_modifier),
                        _p(view), _p(proj), _p(campos), C.c_int(W), C.c_int(H), C.c_float(tanfovx),
                        C.c_float(tanfovy), _p(o["radii"]), _p(o["means2D"]), _p(o["depths"]), _p(o["cov3D"]),
                        _p(o["conic_opacity"]), _p(o["rgb"]), _p(o["clamped"]), _p(o["rect"]),
                        _p(o["tiles_touched"]))
    R = int(L.oracle_count_instances(C.c_int(P), _p(o["tiles_touched"])))
    gx, gy = tile_grid(W, H)
    o["R"] = R; o["tile_bits"] = int(L.oracle_tile_bits(C.c_int(W), C.c_int(H)))
    n = max(R, 1)
    o["keys_unsorted"] = np.zeros(n, np.uint64); o["vals_unsorted"] = np.zeros(n, np.uint32)
    o["keys_sorted"] = np.zeros(n, np.uint64); o["point_list"] = np.zeros(n, np.uint32)
    o["ranges"] = np.zeros((gx * gy, 2), np.uint32)
    L.oracle_bin(C.c_int(P), C.c_int(W), C.c_int(H), _p(o["radii"]), _p(o["depths"]), _p(o["rect"]),
                 _p(o["tiles_touched"]), C.c_int64(R), _p(o["keys_unsorted"]), _p(o["vals_unsorted"]),
                 _p(o["keys_sorted"]), _p(o["point_list"]), _p(o["ranges"]))
    for k in ("keys_unsorted", "vals_unsorted", "keys_sorted", "point_list"):
        o[k] = o[k][:R]
    o["color"] = np.zeros((3, H, W), np.float32); o["depth"] = np.zeros((1, H, W), np.float32)
    o["opacity"] = np.zeros((1, H, W), np.float32); o["final_T"] = np.zeros((H, W), np.float32)
    o["n_contrib"] = np.zeros((H, W), np.uint32); o["n_touched"] = np.zeros(P, np.int32)
    o["margin"] = np.zeros((H, W), np.float32) if want_margin else None
    pl = o["point_list"] if R else np.zeros(1, np.uint32)
    L.oracle_blend_forward(C.c_int(P), C.c_int(W), C.c_int(H), _p(o["ranges"]), _p(pl), _p(o["means2D"]),
                           _p(o["conic_opacity"]), _p(o["rgb"]), _p(o["depths"]), _p(bg), _p(o["color"]),
                           _p(o["depth"]), _p(o["opacity"]), _p(o["final_T"]), _p(o["n_contrib"]),
                           _p(o["n_touched"]), _p(o["margin"]))
    o["_in"] = dict(means3D=means3D, opacities=opacities, scales=scales, rotations=rotations, shs=shs,
                    colors_precomp=colors_precomp, cov3D_precomp=cov3D_precomp, view=view, proj=proj,
                    campos=campos, bg=bg, tanfovx=tanfovx, tanfovy=tanfovy, scale_modifier=scale_modifier)
    return o


def rasterize_backward(fwd, dL_dcolor, dL_ddepth=None, dL_dopacity_img=None, *, projmatrix_raw, flags=0):
    """Backward pass from a `rasterize_forward` result; returns the nine upstream gradient tensors + internals."""
    L = lib()
    P, W, H, M, D, R = fwd["P"], fwd["W"], fwd["H"], fwd["M"], fwd["D"], fwd["R"]
    i = fwd["_in"]
    dL_dcolor = _f(dL_dcolor).reshape(3, H, W)
    dL_ddepth = None if dL_ddepth is None else _f(dL_ddepth).reshape(H, W)
    dL_dopacity_img = None if dL_dopacity_img is None else _f(dL_dopacity_img).reshape(H, W)
    praw = _f(projmatrix_raw).reshape(-1)
    g = {}
    g["dL_dmean2D"] = np.zeros((P, 2), np.float32); g["dL_dconic"] = np.zeros((P, 3), np.float32)
    g["dL_dopacity"] = np.zeros(P, np.float32); g["dL_dcolor"] = np.zeros((P, 3), np.float32)
    g["dL_ddepth"] = np.zeros(P, np.float32)
    pl = fwd["point_list"] if R else np.zeros(1, np.uint32)
    L.oracle_blend_backward(C.c_int(P), C.c_int(W), C.c_int(H), C.c_int64(R), _p(fwd["ranges"]), _p(pl),
                            _p(fwd["means2D"]), _p(fwd["conic_opacity"]), _p(fwd["rgb"]), _p(fwd["depths"]),
                            _p(i["bg"]), _p(fwd["final_T"]), _p(fwd["n_contrib"]), _p(dL_dcolor), _p(dL_ddepth),
                            _p(dL_dopacity_img), C.c_int(flags), _p(g["dL_dmean2D"]), _p(g["dL_dconic"]),
                            _p(g["dL_dopacity"]), _p(g["dL_dcolor"]), _p(g["dL_ddepth"]))
    g["dL_dmeans3D"] = np.zeros((P, 3), np.float32); g["dL_dcov3D"] = np.zeros((P, 6), np.float32)
    g["dL_dsh"] = np.zeros((P, M, 3), np.float32) if i["shs"] is not None else None
    has_sr = i["cov3D_precomp"] is None
    g["dL_dscales"] = np.zeros((P, 3), np.float32) if has_sr else None
    g["dL_drots"] = np.zeros((P, 4), np.float32) if has_sr else None
    g["dL_dtau"] = np.zeros((P, 6), np.float32)
    L.oracle_preprocess_backward(C.c_int(P), C.c_int(D), C.c_int(M), _p(i["means3D"]), _p(fwd["radii"]),
                                 _p(i["shs"]), _p(fwd["clamped"]), _p(i["scales"]), _p(i["rotations"]),
                                 _p(i["cov3D_precomp"]), C.c_float(i["scale_modifier"]), _p(fwd["cov3D"]),
                                 _p(i["view"]), _p(i["proj"]), _p(praw), _p(i["campos"]), C.c_int(W), C.c_int(H),
                                 C.c_float(i["tanfovx"]), C.c_float(i["tanfovy"]), C.c_int(flags),
                                 _p(g["dL_dmean2D"]), _p(g["dL_dconic"]), _p(g["dL_dcolor"]), _p(g["dL_ddepth"]),
                                 _p(g["dL_dmeans3D"]), _p(g["dL_dcov3D"]), _p(g["dL_dsh"]), _p(g["dL_dscales"]),
                                 _p(g["dL_drots"]), _p(g["dL_dtau"]))
    g["dL_dcolors_precomp"] = g["dL_dcolor"] if i["colors_precomp"] is not None else None
    g["grad_rho"] = g["dL_dtau"][:, :3].astype(np.float64).sum(0).astype(np.float32)
    g["grad_theta"] = g["dL_dtau"][:, 3:].astype(np.float64).sum(0).astype(np.float32)
    return g


def mark_visible(means3D, viewmatrix):
    means3D = _f(means3D); P = means3D.shape[0]
    out = np.zeros(P, np.uint8)
    lib().oracle_mark_visible(C.c_int(P), _p(means3D), _p(_f(viewmatrix).reshape(-1)), _p(out))
    return out.astype(bool)


def dist2(points):
    """simple_knn.distCUDA2 restatement: mean squared distance to the 3 nearest neighbours (brute force)."""
    pts = _f(points); P = pts.shape[0]
    out = np.zeros(P, np.float32)
    lib().oracle_dist2(C.c_int(P), _p(pts), _p(out))
    return out
