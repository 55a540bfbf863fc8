"""The sm_100a path against numbers produced by RUNNING reference-held code (tests/golden/reference_pin.npz, made by
tests/golden/make_reference_golden.py from /root/reference/utils/{pose_utils,camera_utils,slam_utils}.py):

  * the tracking loop of utils/slam_frontend.py:1468-1521 -- Camera -> render -> get_loss_tracking -> backward -> Adam ->
    update_pose -- replayed on the CUDA rasterizer must walk the trajectory the reference code walked with the oracle as
    its rasterizer (first iterations to gradient tolerance, the rest to a few learning rates);
  * the CUDA pose gradient with LVDGS_FLAGS=3 (true derivative) must equal finite differences taken through the
    reference's own SE3_exp (utils/pose_utils.py:56-68);
  * the fused loss kernel must reproduce the reference losses and their autograd gradients on the recorded inputs;
  * one mapping iteration through the reference's get_loss_mapping: CUDA parameter gradients == oracle-rasterizer
    gradients, per element.
Camera / losses / update_pose come from tests/ref_conventions.load(): the reference's own modules where /root/reference
is mounted, the restatements pinned against them (tests/test_reference_pin.py) on the GPU box.
"""
import os

import numpy as np
import pytest
import torch

import ref_conventions as rc
from golden.make_reference_golden import TRACK_CFG, fd_scene, loss_case, tracking_scene
from gpu_harness import run_cuda
from test_reference_pin import GOLD, run_tracking_chain

pytestmark = pytest.mark.gpu


def elementwise_close(a, b, rtol=1e-3, what=""):
    """north_star's gradient bar, per element: |a - b| <= rtol |b| + rtol * median|b| (the median term is the absolute
    floor for elements that are themselves rounding noise)."""
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    nz = np.abs(b[b != 0])
    atol = rtol * (np.median(nz) if nz.size else 0.0)
    bad = np.abs(a - b) > rtol * np.abs(b) + atol
    assert not bad.any(), f"{what}: {bad.sum()} of {b.size} elements off, worst {np.abs(a - b)[bad].max():.3e} vs |b| {np.abs(b)[bad][np.argmax(np.abs(a - b)[bad])]:.3e}"


def test_cuda_tracking_loop_walks_the_reference_trajectory():
    from gaussian_splatting.gaussian_renderer import render
    pu, su, cu, _src = rc.load()
    iters = GOLD["trk_loss"].shape[0]
    rec = run_tracking_chain(render, pu, su, cu, "cuda", iters=iters)
    assert abs(rec["target_sum"] - float(GOLD["trk_target_sum"])) <= 1e-5 * abs(float(GOLD["trk_target_sum"]))
    # first iteration: identical inputs -> loss to 1e-5, pose gradient to the gradient bar
    assert abs(rec["loss"][0] - GOLD["trk_loss"][0]) <= 1e-5 * GOLD["trk_loss"][0]
    g0 = np.concatenate([rec["g_trans"][0], rec["g_rot"][0]]); w0 = np.concatenate([GOLD["trk_g_trans"][0], GOLD["trk_g_rot"][0]])
    elementwise_close(g0, w0, 1e-3, "pose gradient, iteration 0")
    # the whole walk: Adam's steps are lr * g / |g|-like, so float noise moves a step by a fraction of a learning rate
    lr_t = TRACK_CFG["Training"]["lr"]["cam_trans_delta"]
    for it in range(iters):
        assert abs(rec["loss"][it] - GOLD["trk_loss"][it]) <= 2e-3 * GOLD["trk_loss"][0]
        assert np.abs(rec["T"][it] - GOLD["trk_T"][it]).max() <= 0.5 * lr_t, it
        assert np.abs(rec["R"][it] - GOLD["trk_R"][it]).max() <= 2e-4, it
    e0 = np.linalg.norm(rec["T"][0] - GOLD["trk_true_T"]); e1 = np.linalg.norm(rec["T"][-1] - GOLD["trk_true_T"])
    assert e1 < 0.25 * e0


def test_cuda_pose_gradient_equals_finite_differences_through_the_reference_se3_exp(monkeypatch):
    import diff_gaussian_rasterization as dgr
    monkeypatch.setattr(dgr, "FLAGS", 3)            # LVDGS_FLAG_EXACT_PP | LVDGS_FLAG_OPACITY_GRAD: the true derivative
    cam, sc, gc, gd, go = fd_scene()
    bg = np.array([0.2, 0.5, 0.1], np.float32)
    out, internals, g = run_cuda(sc, cam, bg, grads=(gc, gd, go))
    assert int(internals["n_contrib"].sum()) == int(GOLD["fd_n_contrib_sum"])
    got = np.concatenate([g["rho"].ravel(), g["theta"].ravel()])
    want = GOLD["fd_dL_dtau"]
    assert np.all(np.abs(got - want) <= 1e-3 * np.abs(want) + 1e-3 * np.median(np.abs(want))), (got, want)
    np.testing.assert_allclose(GOLD["fd_dL_dtau_autograd"], want, rtol=1e-5)


@pytest.mark.parametrize("name", ["trk_rgb", "trk_rgbd", "map_rgb", "map_rgbd"])
def test_fused_loss_kernel_reproduces_the_reference_loss_vectors(name):
    from lvdgs import slam_ops
    kind, mode = {"trk_rgb": ("tracking", True), "trk_rgbd": ("tracking", False), "map_rgb": ("mapping", "rgb"),
                  "map_rgbd": ("mapping", "rgbd")}[name]
    _, _, cu, _ = rc.load()
    for seed in (1, 2):
        d = loss_case(100 + seed)
        t = lambda k: torch.tensor(d[k], device="cuda", requires_grad=True)
        image, depth, opacity = t("image"), t("depth"), t("opacity")
        H, W = d["gt"].shape[1:]
        cam = cu.Camera(0, torch.tensor(d["gt"], device="cuda"), None, d["mono"], torch.eye(4), torch.eye(4), 1., 1., 0., 0., 1., 1.,
                        H, W, device="cuda")
        cam.grad_mask = torch.tensor(d["grad_mask"], device="cuda")
        cam.exposure_a.data.fill_(float(d["a"])); cam.exposure_b.data.fill_(float(d["b"]))
        cfg = {"Training": {"monocular": bool(mode is True or mode == "rgb"), "rgb_boundary_threshold": 0.01, "alpha": 0.9},
               "Dataset": {"depth_loss": False}}
        if kind == "tracking":
            loss = slam_ops.get_loss_tracking(cfg, image, depth, opacity, cam)
        else:
            loss = slam_ops.get_loss_mapping(cfg, image, cam, depth=depth, monodepth=(mode == "rgbd"))
        loss.backward()
        k = f"loss_{name}_{seed}"
        assert abs(float(loss) - float(GOLD[k + "_value"])) <= 2e-6 * abs(float(GOLD[k + "_value"]))
        z = lambda x: np.zeros(x.shape, np.float32) if x.grad is None else x.grad.cpu().numpy()
        for got, key in ((z(image), "_gimage"), (z(depth), "_gdepth"), (z(opacity), "_gopacity")):
            np.testing.assert_allclose(got, GOLD[k + key], rtol=2e-5, atol=1e-9)
        np.testing.assert_allclose(cam.exposure_a.grad.cpu().numpy(), GOLD[k + "_ga"], rtol=2e-5, atol=1e-8)
        np.testing.assert_allclose(cam.exposure_b.grad.cpu().numpy(), GOLD[k + "_gb"], rtol=2e-5, atol=1e-8)


def test_one_mapping_iteration_through_the_reference_loss_matches_the_oracle_rasterizer():
    """utils/slam_backend.py:184-306 in miniature: render -> get_loss_mapping (rgb + mono depth) -> backward, once with the
    CUDA rasterizer and once with the oracle standing in for it; every parameter gradient per element."""
    import importlib
    import sys
    import oracle
    from lvdgs import synth
    pu, su, cu, _src = rc.load()
    c, sc, _tau0, _gm = tracking_scene()
    rng = np.random.default_rng(3)
    H, W = c.image_height, c.image_width
    gt = torch.tensor(rng.uniform(0, 1, (3, H, W)).astype(np.float32))
    mono = rng.uniform(1, 40, (H, W)).astype(np.float32)
    cfg = {"Training": {"monocular": True, "rgb_boundary_threshold": 0.01, "alpha": 0.95}, "Dataset": {"depth_loss": True}}

    def one(render, dev):
        cam = rc.make_camera(cu, c, dev, image=gt.to(dev), mono_depth=mono)
        pc = rc.Gaussians(sc, dev)
        for name in ("get_xyz", "get_opacity", "get_scaling", "get_rotation", "get_features"):
            getattr(pc, name).requires_grad_()
        pkg = render(cam, pc, rc.Pipe(), torch.zeros(3, device=dev))
        loss = su.get_loss_mapping(cfg, pkg["render"], cam, depth=pkg["depth"], monodepth=True)
        loss.backward()
        g = {n: getattr(pc, n).grad.cpu().numpy() for n in ("get_xyz", "get_opacity", "get_scaling", "get_rotation", "get_features")}
        g["viewspace"] = pkg["viewspace_points"].grad.cpu().numpy()
        return float(loss), g, pkg["n_touched"].cpu().numpy(), pkg["radii"].cpu().numpy()

    from gaussian_splatting.gaussian_renderer import render as cuda_render
    loss_c, g_c, nt_c, radii_c = one(cuda_render, "cuda")
    # the oracle-backed stand-in, in a private copy of the shim module so that the product modules stay as they are
    import oracle_rasterizer
    saved = {k: sys.modules.get(k) for k in ("diff_gaussian_rasterization", "gaussian_splatting.gaussian_renderer")}
    try:
        oracle_rasterizer.install()
        cpu_render = importlib.import_module("gaussian_splatting.gaussian_renderer").render
        loss_o, g_o, nt_o, radii_o = one(cpu_render, "cpu")
    finally:
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
    assert abs(loss_c - loss_o) <= 1e-5 * abs(loss_o)
    np.testing.assert_array_equal(radii_c, radii_o)
    assert (nt_c != nt_o).mean() < 2e-3            # knife-edge pixels may move single counts
    for name in g_o:
        # Gaussians with a knife-edge pixel in reach are compared like the rest here: the loss gradient is smooth in them
        a, b = g_c[name], g_o[name]
        scale = np.median(np.abs(b[b != 0])) if (b != 0).any() else 0.0
        bad = np.abs(a - b) > 2e-3 * np.abs(b) + 2e-3 * scale
        assert bad.mean() < 2e-3, (name, bad.mean())
