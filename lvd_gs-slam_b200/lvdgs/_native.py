"""Builds (nvcc, sm_100a) and loads `liblvdgs.so`, the C-ABI library declared in include/lvdgs.h.

There is no CPU or eager-PyTorch fallback: if the library cannot be built or loaded, every op raises.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC_DIR = os.path.join(PKG_DIR, "csrc")
BUILD_DIR = os.path.join(PKG_DIR, "build")
SO_PATH = os.path.join(PKG_DIR, "liblvdgs.so")
SOURCES = ["api.cu", "preprocess.cu", "radix_sort.cu", "tile_sort.cu", "slam_ops.cu", "blend_forward.cu", "blend_backward.cu",
           "preprocess_backward.cu", "knn.cu", "adam.cu", "exchange.cu", "cub_compare.cu", "peak.cu", "ssim_loss.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"] + os.environ.get("LVDGS_NVCC_DEFS", "").split()

_lib = None


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def _stale():
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    deps = [os.path.join(SRC_DIR, f) for f in os.listdir(SRC_DIR)] + \
           [os.path.join(os.path.dirname(PKG_DIR), "include", "lvdgs.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a and link liblvdgs.so in-tree (cross-compiles without a GPU)."""
    if not force and not _stale():
        return SO_PATH
    nvcc = _nvcc()
    if nvcc is None:
        raise RuntimeError("liblvdgs.so is missing/stale and nvcc was not found; there is no CPU fallback")
    os.makedirs(BUILD_DIR, exist_ok=True)
    # one builder at a time (torchrun starts one process per GPU against the same tree); the others wait, then find
    # the library fresh
    import fcntl
    lock = open(os.path.join(BUILD_DIR, ".lock"), "w")
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        if not force and not _stale():
            return SO_PATH
        return _build_locked(nvcc, force, verbose)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _build_locked(nvcc, force, verbose):
    hdrs = [os.path.join(SRC_DIR, f) for f in os.listdir(SRC_DIR) if f.endswith((".cuh", ".h"))] + \
           [os.path.join(os.path.dirname(PKG_DIR), "include", "lvdgs.h")]
    hdr_t = max(os.path.getmtime(h) for h in hdrs)

    def compile_one(src):
        s = os.path.join(SRC_DIR, src)
        o = os.path.join(BUILD_DIR, src.replace(".cu", ".o"))
        if not force and os.path.exists(o) and os.path.getmtime(o) > max(os.path.getmtime(s), hdr_t):
            return o
        cmd = [nvcc, "-c", *NVCC_FLAGS, "-o", o, s]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(" ".join(cmd))
        return o

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO_PATH, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return SO_PATH


class RasterParams(C.Structure):
    _fields_ = [("P", C.c_int32), ("sh_degree", C.c_int32), ("sh_coeffs", C.c_int32), ("width", C.c_int32),
                ("height", C.c_int32), ("tan_fovx", C.c_float), ("tan_fovy", C.c_float),
                ("scale_modifier", C.c_float), ("prefiltered", C.c_int32), ("debug", C.c_int32),
                ("flags", C.c_int32)]


class GeomLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in ("depths", "means2D", "conic_opacity", "rgbd", "rect", "tiles_touched",
                                          "point_offsets", "clamped", "scan_state", "visible_list", "total")]


class BinningLayout(C.Structure):
    _fields_ = [("keys", C.c_size_t * 2), ("vals", C.c_size_t * 2), ("sort_ws", C.c_size_t),
                ("sorted_sel", C.c_size_t), ("total", C.c_size_t)]


class ImgLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in ("final_T", "n_contrib", "ranges", "tile_order", "tile_grid", "sort_hist", "tile_cursor", "total")]


RESIZE_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_int32, C.c_size_t)


class StaticBuffers(C.Structure):        # lvdgs_static_buffers (include/lvdgs.h)
    _fields_ = [("base", C.c_void_p * 3), ("capacity", C.c_size_t * 3), ("fallback", RESIZE_FN), ("fallback_user", C.c_void_p)]

EXPORTS = ["lvdgs_version", "lvdgs_last_error", "lvdgs_set_device", "lvdgs_launch_count", "lvdgs_reset_launch_count", "lvdgs_tail_rerun_count",
           "lvdgs_profile_begin", "lvdgs_profile_end",
           "lvdgs_get_geom_layout", "lvdgs_get_binning_layout", "lvdgs_get_img_layout", "lvdgs_rasterize_forward",
           "lvdgs_backward_scratch_bytes", "lvdgs_rasterize_backward", "lvdgs_mark_visible",
           "lvdgs_dist2_workspace_bytes", "lvdgs_dist2", "lvdgs_adam_step", "lvdgs_sort_workspace_bytes", "lvdgs_sort_pairs",
           "lvdgs_cub_sort_workspace_bytes", "lvdgs_cub_sort_pairs",
           "lvdgs_fused_loss_workspace_bytes", "lvdgs_fused_loss", "lvdgs_covis_counts", "lvdgs_n_obs",
           "lvdgs_compact_workspace_bytes", "lvdgs_compact_count", "lvdgs_compact_move", "lvdgs_pose_step", "lvdgs_gather_rows",
           "lvdgs_fp32_peak", "lvdgs_gaussian_activate", "lvdgs_gaussian_activation_backward",
           "lvdgs_masked_ssim_loss_workspace_bytes", "lvdgs_masked_ssim_loss", "lvdgs_exchange_adam", "lvdgs_zero_async", "lvdgs_static_resize"]


def lib():
    """The loaded library (built on demand when nvcc is present).  Raises if unavailable -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    # LVDGS_SO=<path>: load a variant built by scripts/build_variant.py (kernel-tuning experiments), no staleness check
    so = os.environ.get("LVDGS_SO") or SO_PATH
    if so == SO_PATH and _stale() and _nvcc() is not None:
        build()
    if not os.path.exists(so):
        raise RuntimeError(f"{so} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(so)
    vp, i32, i64, sz, f = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t, C.c_float
    L.lvdgs_version.restype = C.c_int
    L.lvdgs_last_error.restype = C.c_char_p
    L.lvdgs_set_device.argtypes = [C.c_int]
    L.lvdgs_launch_count.restype = i64
    L.lvdgs_reset_launch_count.restype = None
    L.lvdgs_tail_rerun_count.restype = i64
    L.lvdgs_profile_begin.argtypes = [vp]
    L.lvdgs_profile_end.argtypes = [vp, C.c_char_p, sz, C.POINTER(f), i32]
    L.lvdgs_get_geom_layout.argtypes = [i32, C.POINTER(GeomLayout)]
    L.lvdgs_get_binning_layout.argtypes = [i64, C.POINTER(BinningLayout)]
    L.lvdgs_get_img_layout.argtypes = [i32, i32, C.POINTER(ImgLayout)]
    L.lvdgs_rasterize_forward.argtypes = [C.POINTER(RasterParams)] + [vp] * 12 + [RESIZE_FN, vp, i64] + [vp] * 5 + \
                                         [C.POINTER(i64), C.POINTER(i64), vp]
    L.lvdgs_backward_scratch_bytes.argtypes = [i32, i64]
    L.lvdgs_backward_scratch_bytes.restype = sz
    L.lvdgs_rasterize_backward.argtypes = [C.POINTER(RasterParams)] + [vp] * 17 + [i64, i64, vp, vp, vp, sz] + [vp] * 11
    L.lvdgs_mark_visible.argtypes = [i32, vp, vp, vp, vp, vp]
    L.lvdgs_dist2_workspace_bytes.argtypes = [i32]
    L.lvdgs_dist2_workspace_bytes.restype = sz
    L.lvdgs_dist2.argtypes = [i32, vp, vp, vp, sz, vp]
    L.lvdgs_adam_step.argtypes = [i64, vp, vp, vp, vp, i32, C.POINTER(i64), C.POINTER(f), C.c_double, C.c_double, C.c_double, i32, vp]
    L.lvdgs_exchange_adam.argtypes = [i32, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), i64, i64, vp, vp, i32, C.POINTER(i64),
                                      C.POINTER(f), C.POINTER(i64), i64, C.c_double, C.c_double, C.c_double, i32, vp, vp, vp, i32, vp]
    L.lvdgs_zero_async.argtypes = [vp, sz, vp]
    L.lvdgs_static_resize.argtypes = [vp, i32, sz]
    L.lvdgs_static_resize.restype = vp
    L.lvdgs_sort_workspace_bytes.argtypes = [i64]
    L.lvdgs_sort_workspace_bytes.restype = sz
    L.lvdgs_sort_pairs.argtypes = [i64, vp, vp, vp, vp, i32, vp, sz, C.POINTER(i32), vp]
    L.lvdgs_cub_sort_workspace_bytes.argtypes = [i64, i32]
    L.lvdgs_cub_sort_workspace_bytes.restype = sz
    L.lvdgs_cub_sort_pairs.argtypes = [i64, vp, vp, vp, vp, i32, vp, sz, vp]
    L.lvdgs_fused_loss_workspace_bytes.restype = sz
    L.lvdgs_fused_loss.argtypes = [i32, i32] + [vp] * 7 + [f, f, f, i32] + [vp] * 5 + [sz, vp]
    L.lvdgs_covis_counts.argtypes = [i64, vp, vp, i32, vp, vp]
    L.lvdgs_n_obs.argtypes = [i64, i32, vp, i32, vp, vp]
    L.lvdgs_compact_workspace_bytes.argtypes = [i64]
    L.lvdgs_compact_workspace_bytes.restype = sz
    L.lvdgs_compact_count.argtypes = [i64, vp, vp, sz, C.POINTER(vp), vp]
    L.lvdgs_compact_move.argtypes = [i64, vp, vp, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(i32), vp]
    L.lvdgs_gather_rows.argtypes = [i64, vp, i64, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(i32), vp]
    L.lvdgs_gaussian_activate.argtypes = [i64] + [vp] * 7
    L.lvdgs_gaussian_activation_backward.argtypes = [i64] + [vp] * 8
    L.lvdgs_masked_ssim_loss_workspace_bytes.argtypes = [i32, i32]
    L.lvdgs_masked_ssim_loss_workspace_bytes.restype = sz
    L.lvdgs_masked_ssim_loss.argtypes = [i32, i32] + [vp] * 6 + [f, f] + [vp] * 4 + [sz, vp]
    L.lvdgs_fp32_peak.argtypes = [i32, i32, i32, vp, C.POINTER(C.c_double), vp]
    L.lvdgs_pose_step.argtypes = [vp, vp, vp, f, f, f, C.c_double, C.c_double, C.c_double, i32, f, vp]
    _lib = L
    return L


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed: {lib().lvdgs_last_error().decode()}")


def ptr(t):
    """Device pointer of a torch tensor (or None) as c_void_p."""
    return None if t is None else C.c_void_p(t.data_ptr())


def debug_views(bufs, P: int, R: int, W: int, H: int, capacity=None):
    """numpy copies of the arrays inside the three opaque buffers (parity tests / debugging only).

    `bufs` maps LVDGS_BUF_{GEOM,BINNING,IMG} -> torch uint8 tensor as handed out by the resize callback."""
    import numpy as np
    L = lib()
    out = {}

    def grab(buf, off, nbytes, dtype, shape):
        raw = buf[off:off + nbytes].cpu().numpy()
        return raw.view(dtype).reshape(shape).copy()

    gl, bl, il = GeomLayout(), BinningLayout(), ImgLayout()
    L.lvdgs_get_geom_layout(P, C.byref(gl))
    L.lvdgs_get_binning_layout(capacity if capacity is not None else R, C.byref(bl))
    L.lvdgs_get_img_layout(W, H, C.byref(il))
    g = bufs.get(0)
    if g is not None and P > 0:
        out["depths"] = grab(g, gl.depths, 4 * P, np.float32, (P,))
        m4 = grab(g, gl.means2D, 16 * P, np.float32, (P, 4))
        out["means2D"] = np.ascontiguousarray(m4[:, :2])
        out["extent"] = np.ascontiguousarray(m4[:, 2:])
        out["conic_opacity"] = grab(g, gl.conic_opacity, 16 * P, np.float32, (P, 4))
        out["rgbd"] = grab(g, gl.rgbd, 16 * P, np.float32, (P, 4))
        out["rect"] = grab(g, gl.rect, 8 * P, np.int16, (P, 4)).astype(np.int32)
        out["tiles_touched"] = grab(g, gl.tiles_touched, 4 * P, np.uint32, (P,))
        out["point_offsets"] = grab(g, gl.point_offsets, 4 * P, np.uint32, (P,))
        out["clamped"] = grab(g, gl.clamped, P, np.uint8, (P,))
    b = bufs.get(1)
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    if b is not None and R > 0:
        sel = int(grab(b, bl.sorted_sel, 4, np.int32, (1,))[0])
        out["sorted_sel"] = sel
        out["keys_unsorted"] = grab(b, bl.keys[0], 8 * R, np.uint64, (R,)) if sel == 1 else None
        out["keys_sorted"] = grab(b, bl.keys[sel], 8 * R, np.uint64, (R,))
        out["point_list"] = grab(b, bl.vals[sel], 4 * R, np.uint32, (R,))
    i = bufs.get(2)
    if i is not None:
        out["final_T"] = grab(i, il.final_T, 4 * W * H, np.float32, (H, W))
        out["n_contrib"] = grab(i, il.n_contrib, 4 * W * H, np.uint32, (H, W))
        out["ranges"] = grab(i, il.ranges, 8 * tiles, np.uint32, (tiles, 2))
    return out


def profile_begin(stream_ptr):
    check(lib().lvdgs_profile_begin(C.c_void_p(stream_ptr)), "lvdgs_profile_begin")


def profile_end(stream_ptr, max_entries=4096):
    """-> list of (kernel name, milliseconds) for every launch since profile_begin."""
    names = C.create_string_buffer(64 * max_entries)
    ms = (C.c_float * max_entries)()
    n = lib().lvdgs_profile_end(C.c_void_p(stream_ptr), names, len(names), ms, max_entries)
    if n < 0:
        raise RuntimeError("lvdgs_profile_end failed")
    nm = names.value.decode().split("\n")[:n]
    return [(nm[i], float(ms[i])) for i in range(n)]
