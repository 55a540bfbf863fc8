"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports every symbol
include/lvdgs.h declares; the Python plugin surface has the reference's names and error behaviour; buffer layouts are
sane.  No compute calls (there is no GPU here)."""
import ctypes as C
import os
import re

import pytest
import torch

from lvdgs import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol():
    so = _native.build()
    assert os.path.exists(so)
    L = _native.lib()
    hdr = open(os.path.join(ROOT, "include", "lvdgs.h")).read()
    declared = set(re.findall(r"\b(lvdgs_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/lvdgs.h but not exported"
    assert declared == set(_native.EXPORTS)
    assert L.lvdgs_version() >= 100


def test_sass_is_sm100a_only():
    out = os.popen(f"cuobjdump -lelf {_native.SO_PATH} 2>/dev/null").read()
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_layouts_are_aligned_and_monotone():
    L = _native.lib()
    gl, bl, il = _native.GeomLayout(), _native.BinningLayout(), _native.ImgLayout()
    assert L.lvdgs_get_geom_layout(1000, C.byref(gl)) == 0
    offs = [gl.depths, gl.means2D, gl.conic_opacity, gl.rgbd, gl.rect, gl.tiles_touched, gl.point_offsets, gl.clamped, gl.total]
    assert offs == sorted(offs) and all(o % 256 == 0 for o in offs)
    assert gl.conic_opacity - gl.means2D >= 16 * 1000
    assert L.lvdgs_get_binning_layout(12345, C.byref(bl)) == 0
    assert bl.keys[1] - bl.keys[0] >= 8 * 12345 and bl.vals[1] - bl.vals[0] >= 4 * 12345 and bl.total > bl.sort_ws
    assert L.lvdgs_get_img_layout(1241, 376, C.byref(il)) == 0
    assert il.n_contrib - il.final_T >= 4 * 1241 * 376 and il.total - il.ranges >= 8 * 78 * 24
    assert L.lvdgs_backward_scratch_bytes(1000, 0) >= 1000 * 48


def test_plugin_surface_names_and_errors():
    import diff_gaussian_rasterization as dgr
    from simple_knn._C import distCUDA2
    from gaussian_splatting.gaussian_renderer import render, render_with_custom_resolution  # noqa: F401
    fields = dgr.GaussianRasterizationSettings._fields
    assert fields == ("image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix",
                      "projmatrix", "projmatrix_raw", "sh_degree", "campos", "prefiltered", "debug")
    rs = dgr.GaussianRasterizationSettings(8, 8, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), torch.eye(4), 0,
                                           torch.zeros(3), False, False)
    rast = dgr.GaussianRasterizer(rs)
    z = torch.zeros(4, 3)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        rast(means3D=z, means2D=z, opacities=z[:, :1], scales=z, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        rast(means3D=z, means2D=z, opacities=z[:, :1], shs=torch.zeros(4, 1, 3), colors_precomp=z, scales=z, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        rast(means3D=z, means2D=z, opacities=z[:, :1], shs=torch.zeros(4, 1, 3), scales=z, rotations=torch.zeros(4, 4),
             cov3D_precomp=torch.zeros(4, 6))
    # no CPU fallback anywhere on the product path
    with pytest.raises(RuntimeError, match="no CPU path"):
        rast(means3D=z, means2D=z, opacities=z[:, :1], shs=torch.zeros(4, 1, 3), scales=z, rotations=torch.zeros(4, 4))
    with pytest.raises(RuntimeError, match="no CPU path"):
        distCUDA2(torch.zeros(5, 3))


def test_render_shim_returns_none_for_empty_map():
    from gaussian_splatting.gaussian_renderer import render

    class PC:
        get_xyz = torch.zeros(0, 3)

    assert render(object(), PC(), None, torch.zeros(3)) is None


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "lvd_gs-slam_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(import oracle|from oracle)", src, re.M), os.path.join(d, f)


def test_slam_ops_reject_cpu_tensors():
    """The ops either side of the rasterizer have no CPU path either."""
    import torch
    from lvdgs import slam_ops
    from lvdgs.mapping import ShardedMapper
    with pytest.raises(RuntimeError, match="no CPU path"):
        slam_ops.fused_loss(torch.zeros(3, 4, 4), gt_image=torch.zeros(3, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU path"):
        slam_ops.covisibility(torch.zeros(8, dtype=torch.bool), torch.zeros(8, dtype=torch.bool))
    with pytest.raises(RuntimeError, match="no CPU path"):
        slam_ops.compact_rows(torch.ones(8, dtype=torch.bool), [torch.zeros(8, 3)])
    with pytest.raises(RuntimeError, match="no CPU path"):
        ShardedMapper(8, device="cpu").prune(torch.ones(8, dtype=torch.bool))


def _header():
    return open(os.path.join(ROOT, "include", "lvdgs.h")).read()


def test_ctypes_prototypes_have_the_headers_arity():
    """Every bound function takes as many arguments as include/lvdgs.h declares (a silent mismatch would corrupt the
    call frame instead of raising)."""
    L = _native.lib()
    text = re.sub(r"/\*.*?\*/", "", _header(), flags=re.S)
    protos = dict(re.findall(r"\b(lvdgs_\w+)\s*\(([^;{]*?)\)\s*;", text))
    assert set(_native.EXPORTS) <= set(protos)
    for name in _native.EXPORTS:
        params = protos[name].strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        fn = getattr(L, name)
        if fn.argtypes is not None:
            assert len(fn.argtypes) == n, (name, len(fn.argtypes), n)


def test_pose_state_offsets_match_the_header():
    """lvdgs.tracking addresses fields of lvdgs_pose_state by float offset."""
    from lvdgs import tracking as trk
    text = re.sub(r"/\*.*?\*/", "", _header(), flags=re.S)
    body = re.search(r"typedef struct lvdgs_pose_state \{(.*?)\} lvdgs_pose_state;", text, re.S).group(1)
    off, offsets = 0, {}
    for typ, name, dim in re.findall(r"(float|int32_t)\s+(\w+)(?:\[(\d+)\])?\s*;", body):
        offsets[name] = off
        off += int(dim) if dim else 1
    want = dict(view=trk._VIEW, proj=trk._PROJ, proj_raw=trk._PRAW, campos=trk._CAMPOS, R=trk._R, T=trk._T,
                exposure=trk._EXPO, adam_m=trk._M, adam_v=trk._V, step=trk._STEP, converged=trk._CONV, tau_norm=trk._TAUN)
    for k, v in want.items():
        assert offsets[k] == v, (k, offsets[k], v)
    assert off == trk._SIZE
    with pytest.raises(RuntimeError, match="no CPU path"):
        trk.PoseTracker(8, 32, 32, 1.0, 1.0, device="cpu")
