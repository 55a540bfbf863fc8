"""Pins the oracle's analytic backward (oracle/raster_oracle.c) against float64 autograd of the forward.

With flags = EXACT_PP | OPACITY_GRAD the analytic backward must be the true derivative (SURVEY.md A.6 items 1 and 3
switched off); with flags = 0 (upstream behaviour) the only terms allowed to move are the pose gradient and
the dropped opacity-image gradient, which is asserted separately.
"""
import numpy as np
import pytest
import torch

import oracle
from lvdgs import synth
from f64_reference import forward as f64_forward


def _scene(N, sh_degree, seed, centered=False, name="tiny"):
    W, H = 72, 56
    cam = synth.Cam(W, H, 60.0, 58.0, W / 2.0 if centered else W / 2.0 - 3.3, H / 2.0 if centered else H / 2.0 + 2.1,
                    np.eye(3), np.zeros(3))
    a = np.radians(7.0)
    cam.R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]) @ \
        np.array([[1, 0, 0], [0, np.cos(0.05), -np.sin(0.05)], [0, np.sin(0.05), np.cos(0.05)]])
    cam.T = np.array([0.1, -0.05, 0.3])
    sc = synth.make_scene(N, cam, seed=seed, sh_degree=sh_degree, behind_frac=0.05)
    # move the cloud into this camera's frame so that it fills the frustum, keep everything inside 1.3x FoV
    W2C = synth.getWorld2View2(cam.R, cam.T)
    pc = sc["means3D"].astype(np.float64)
    pw = (np.linalg.inv(W2C) @ np.concatenate([pc, np.ones((N, 1))], 1).T).T[:, :3]
    sc["means3D"] = pw.astype(np.float32)
    sc["scales"] *= 2.0
    sc["opacities"] = np.clip(sc["opacities"], 0.05, 0.9)
    return cam, sc


@pytest.mark.parametrize("sh_degree", [0, 3])
def test_backward_matches_f64_autograd(sh_degree):
    N = 160
    cam, sc = _scene(N, sh_degree, seed=3)
    bg = np.array([0.2, 0.5, 0.1], np.float32)
    kw = dict(viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, campos=cam.camera_center,
              bg=bg, W=cam.image_width, H=cam.image_height, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
              sh_degree=sh_degree)
    fwd = oracle.rasterize_forward(sc["means3D"], sc["opacities"], sc["scales"], sc["rotations"], sc["shs"], **kw)
    assert fwd["R"] > 500 and (fwd["radii"] > 0).sum() > 50
    rng = np.random.default_rng(7)
    H, W = cam.image_height, cam.image_width
    gc = rng.normal(0, 1, (3, H, W)).astype(np.float32)
    gd = rng.normal(0, 1, (1, H, W)).astype(np.float32) * 0.3
    go = rng.normal(0, 1, (1, H, W)).astype(np.float32)
    flags = oracle.FLAG_EXACT_PP | oracle.FLAG_OPACITY_GRAD
    g = oracle.rasterize_backward(fwd, gc, gd, go, projmatrix_raw=cam.projection_matrix, flags=flags)

    dt = torch.float64
    t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), dtype=dt)
    means = t(sc["means3D"]).requires_grad_()
    scales = t(sc["scales"]).requires_grad_()
    rots = t(sc["rotations"]).requires_grad_()
    opac = t(sc["opacities"]).requires_grad_()
    shs = t(sc["shs"]).requires_grad_()
    tau = torch.zeros(6, dtype=dt, requires_grad=True)
    W2C = t(synth.getWorld2View2(cam.R, cam.T))
    Pr = t(synth.getProjectionMatrix2(0.01, 100.0, cam.cx, cam.cy, cam.fx, cam.fy, W, H))
    color, dimg, oimg = f64_forward(means, scales, rots, opac, shs, tau, W2C=W2C, Pr=Pr, campos=t(cam.camera_center),
                                    bg=bg, W=W, H=H, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, sh_degree=sh_degree,
                                    fwd=fwd)
    # forward agreement (float32 oracle vs float64) away from knife-edge pixels
    ok = fwd["margin"] > 1e-4
    assert ok.mean() > 0.98
    np.testing.assert_allclose(fwd["color"][:, ok], color.detach().numpy()[:, ok], atol=2e-5)
    np.testing.assert_allclose(fwd["depth"][0][ok], dimg.detach().numpy()[ok], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(fwd["opacity"][0][ok], oimg.detach().numpy()[ok], atol=2e-5)
    loss = (color * t(gc)).sum() + (dimg * t(gd[0])).sum() + (oimg * t(go[0])).sum()
    loss.backward()

    def close(name, a, b, rtol=2e-3):
        a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
        scale = np.abs(b).max() + 1e-30
        err = np.abs(a - b).max() / scale
        print(f"{name}: relerr {err:.2e} scale {scale:.3e}")
        assert err < rtol, f"{name}: max err / max |ref| = {err:.3e}"

    close("means3D", g["dL_dmeans3D"], means.grad.numpy())
    close("scales", g["dL_dscales"], scales.grad.numpy())
    close("rots", g["dL_drots"], rots.grad.numpy())
    close("opacity", g["dL_dopacity"], opac.grad.numpy().reshape(-1))
    close("sh", g["dL_dsh"], shs.grad.numpy())
    if sh_degree == 0:   # A.6 item 7: upstream's campos pose term for degree>0 is not the true derivative
        close("rho", g["grad_rho"], tau.grad.numpy()[:3])
        close("theta", g["grad_theta"], tau.grad.numpy()[3:])

    # upstream behaviour (flags=0): only pose (principal point) and the dropped opacity-image gradient may differ
    g0 = oracle.rasterize_backward(fwd, gc, gd, None, projmatrix_raw=cam.projection_matrix, flags=0)
    g1 = oracle.rasterize_backward(fwd, gc, gd, None, projmatrix_raw=cam.projection_matrix, flags=oracle.FLAG_EXACT_PP)
    for k in ("dL_dmeans3D", "dL_dscales", "dL_drots", "dL_dopacity", "dL_dsh", "dL_dcov3D"):
        np.testing.assert_array_equal(g0[k], g1[k])
    assert np.abs(g0["grad_theta"] - g1["grad_theta"]).max() > 0   # cx != W/2 here


def test_pose_flag_is_noop_for_centred_principal_point():
    cam, sc = _scene(120, 0, seed=5, centered=True)
    kw = dict(viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, campos=cam.camera_center,
              bg=np.zeros(3), W=cam.image_width, H=cam.image_height, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy)
    fwd = oracle.rasterize_forward(sc["means3D"], sc["opacities"], sc["scales"], sc["rotations"], sc["shs"], **kw)
    gc, gd = synth.make_upstream_grads(cam)
    g0 = oracle.rasterize_backward(fwd, gc, gd, projmatrix_raw=cam.projection_matrix, flags=0)
    g1 = oracle.rasterize_backward(fwd, gc, gd, projmatrix_raw=cam.projection_matrix, flags=oracle.FLAG_EXACT_PP)
    np.testing.assert_allclose(g0["dL_dtau"], g1["dL_dtau"], rtol=1e-5, atol=1e-12)
