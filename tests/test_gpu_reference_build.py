"""The sm_100a kernels against the reference's OWN CUDA rasterizer -- runs only where oracle/build_ref.py found the
reference's submodule sources and built oracle/_ref (today they are absent from /root/reference, so this test SKIPS and
parity stays pinned through the reference's Python files only).  Bars are north_star's: radii and
n_touched bit-exact, images within 1e-5, parameter / pose gradients within 1e-3 per element."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import build_ref  # noqa: E402

from lvdgs import synth  # noqa: E402
from gpu_harness import run_cuda, grad_mismatch  # noqa: E402

pytestmark = pytest.mark.gpu
REF_SO = build_ref.available()["diff_gaussian_rasterization_ref"]


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.skipif(REF_SO is None, reason="oracle/_ref not built: the reference's rasterizer sources are absent (.MISSING_LARGE_BLOBS)")
def test_cuda_path_matches_the_reference_build():
    ref = _load(REF_SO, "diff_gaussian_rasterization_ref")
    dev = "cuda"
    cam = synth.make_camera("kitti")
    sc = synth.make_scene(100_000, cam, seed=11)
    bg = np.zeros(3, np.float32)
    gc, gd = synth.make_upstream_grads(cam)
    out, _, g = run_cuda(sc, cam, bg, grads=(gc, gd, None), debug=False)
    t = lambda a: torch.tensor(np.ascontiguousarray(a), device=dev)
    empty = torch.empty(0, device=dev)
    args = (t(bg), t(sc["means3D"]), empty, t(sc["opacities"]), t(sc["scales"]), t(sc["rotations"]), 1.0, empty,
            t(cam.world_view_transform), t(cam.full_proj_transform), t(cam.projection_matrix), cam.tanfovx, cam.tanfovy,
            cam.image_height, cam.image_width, t(sc["shs"]), 0, t(cam.camera_center), False, False)
    R, color, radii, geom, binning, img, depth, opacity, n_touched = ref.rasterize_gaussians(*args)
    assert np.array_equal(radii.cpu().numpy(), out["radii"])
    assert np.array_equal(n_touched.cpu().numpy(), out["n_touched"])
    for a, b in ((color, out["color"]), (depth, out["depth"]), (opacity, out["opacity"])):
        assert np.abs(a.cpu().numpy() - b).max() <= 1e-5
    grads = ref.rasterize_gaussians_backward(t(bg), t(sc["means3D"]), radii, empty, t(sc["scales"]), t(sc["rotations"]), 1.0, empty,
                                             t(cam.world_view_transform), t(cam.full_proj_transform), t(cam.projection_matrix),
                                             cam.tanfovx, cam.tanfovy, t(gc), t(gd), t(sc["shs"]), 0, t(cam.camera_center), geom, R,
                                             binning, img, False)
    g_means2D, g_colors, g_opac, g_means3D, g_cov, g_sh, g_scales, g_rots, g_tau = grads
    for name, ref_g in (("means3D", g_means3D), ("opacities", g_opac), ("scales", g_scales), ("rotations", g_rots), ("shs", g_sh)):
        assert grad_mismatch(g[name], ref_g.cpu().numpy().reshape(g[name].shape)) == 0, name
    tau = g_tau.view(-1, 6).sum(0).cpu().numpy()
    np.testing.assert_allclose(np.concatenate([g["rho"].ravel(), g["theta"].ravel()]), tau, rtol=1e-3, atol=1e-3 * np.abs(tau).max())
