"""Golden fixtures (tests/golden/raster_small.npz, produced by tests/golden/make_golden.py from the oracle):
on CPU they pin the oracle against regressions; on the GPU they anchor the CUDA path to a committed file."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden import build_case, run_oracle  # noqa: E402

G = np.load(os.path.join(HERE, "golden", "raster_small.npz"))


def test_generator_inputs_are_reproducible():
    cam, sc, bg, gc, gd = build_case()
    for k in ("means3D", "opacities", "scales", "rotations", "shs"):
        np.testing.assert_array_equal(sc[k], G[k])
    np.testing.assert_array_equal(gc, G["grad_color"])
    np.testing.assert_array_equal(cam.full_proj_transform, G["projmatrix"])


def test_oracle_reproduces_golden():
    cam, sc, bg, gc, gd = build_case()
    fwd, g = run_oracle(cam, sc, bg, gc, gd)
    for k in ("radii", "n_touched", "n_contrib", "keys_sorted", "point_list", "ranges"):
        np.testing.assert_array_equal(fwd[k], G[k])
    for k in ("color", "depth", "opacity"):
        np.testing.assert_allclose(fwd[k], G[k], rtol=0, atol=1e-6)     # expf may differ in the last ulp across libm builds
    for k in ("dL_dmeans3D", "dL_dopacity", "dL_dscales", "dL_drots", "dL_dsh", "grad_rho", "grad_theta"):
        np.testing.assert_allclose(g[k], G[k], rtol=1e-4, atol=1e-7)
    assert G["keys_sorted"].size > 1000 and (G["radii"] > 0).sum() > 100


@pytest.mark.gpu
def test_cuda_matches_golden():
    import torch
    import diff_gaussian_rasterization as dgr
    dev = "cuda"
    t = lambda a, rg=False: torch.tensor(np.asarray(a), dtype=torch.float32, device=dev, requires_grad=rg)
    rs = dgr.GaussianRasterizationSettings(
        image_height=48, image_width=80, tanfovx=float(G["tanfov"][0]), tanfovy=float(G["tanfov"][1]), bg=t(G["bg"]),
        scale_modifier=1.0, viewmatrix=t(G["viewmatrix"]), projmatrix=t(G["projmatrix"]), projmatrix_raw=t(G["projmatrix_raw"]),
        sh_degree=1, campos=t(G["campos"]), prefiltered=False, debug=False)
    means, opac, scales, rots, shs = (t(G[k], True) for k in ("means3D", "opacities", "scales", "rotations", "shs"))
    m2d = torch.zeros_like(means, requires_grad=True)
    theta = torch.zeros(3, device=dev, requires_grad=True); rho = torch.zeros(3, device=dev, requires_grad=True)
    color, radii, depth, opacity, n_touched = dgr.GaussianRasterizer(rs)(
        means3D=means, means2D=m2d, opacities=opac, shs=shs, scales=scales, rotations=rots, theta=theta, rho=rho)
    torch.autograd.backward([color, depth], [t(G["grad_color"]), t(G["grad_depth"])])
    ok = G["margin"] > 1e-5
    np.testing.assert_array_equal(radii.cpu().numpy(), G["radii"])
    assert np.abs(color.detach().cpu().numpy()[:, ok] - G["color"][:, ok]).max() < 1e-5
    assert np.abs(opacity.detach().cpu().numpy()[0][ok] - G["opacity"][0][ok]).max() < 1e-5
    d = depth.detach().cpu().numpy()[0]
    assert (np.abs(d - G["depth"][0]) / np.maximum(1, np.abs(G["depth"][0])))[ok].max() < 1e-5
    if ok.all():
        np.testing.assert_array_equal(n_touched.cpu().numpy(), G["n_touched"])
    rel = lambda a, b: np.abs(np.asarray(a, np.float64) - b).max() / (np.abs(b).max() + 1e-30)
    assert rel(means.grad.cpu().numpy(), G["dL_dmeans3D"]) < 1e-3
    assert rel(opac.grad.cpu().numpy().reshape(-1), G["dL_dopacity"]) < 1e-3
    assert rel(scales.grad.cpu().numpy(), G["dL_dscales"]) < 1e-3
    assert rel(rots.grad.cpu().numpy(), G["dL_drots"]) < 1e-3
    assert rel(shs.grad.cpu().numpy(), G["dL_dsh"]) < 1e-3
    assert rel(m2d.grad.cpu().numpy()[:, :2], G["dL_dmean2D"]) < 1e-3
    assert rel(rho.grad.cpu().numpy(), G["grad_rho"]) < 1e-3 and rel(theta.grad.cpu().numpy(), G["grad_theta"]) < 1e-3
