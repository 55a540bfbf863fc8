"""Generates tests/golden/reference_pin.npz by RUNNING the reference's own Python from /root/reference (read-only):
utils/pose_utils.py (SO3_exp, V, SE3_exp, update_pose), utils/camera_utils.py (Camera and its derived matrices) and
utils/slam_utils.py (get_loss_tracking*, get_loss_mapping*).  These three files are the only reference-held code on
or next to the rasterizer hot path (the rasterizer itself is absent, /root/reference/.MISSING_LARGE_BLOBS:1); they
define the conventions the pose gradient has to obey.

    python tests/golden/make_reference_golden.py          (in the build container; needs /root/reference, no GPU)

Sections of the fixture:
  se3_* / up_* / cam_*   known-answer vectors of SE3_exp, update_pose and the Camera matrices;
  loss_*                 the four loss variants with autograd gradients on small random images;
  trk_*                  a 30-iteration tracking loop exactly as utils/slam_frontend.py:1468-1521 runs it -- REAL Camera,
                         the shim's render(), REAL get_loss_tracking, loss.backward(), torch Adam, REAL update_pose -- with
                         the CPU oracle standing in for the CUDA rasterizer (tests/oracle_rasterizer.py); per-iteration
                         loss, pose gradients and poses.  The GPU test replays it on the sm_100a kernels;
  fd_*                   dL/dtau in float64 THROUGH THE REFERENCE'S SE3_exp (tau -> SE3_exp(tau) T_w2c -> float64 forward), by
                         central differences (blend decisions frozen) and by autograd -- the number the CUDA pose gradient
                         (LVDGS_FLAGS=3) and the oracle's analytic one must reproduce.
The reference hard-codes `.cuda()`; on this GPU-less box `Tensor.cuda` is made the identity (tests/ref_conventions.py).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
ROOT = os.path.dirname(TESTS)
for p in (ROOT, os.path.join(ROOT, "lvd_gs-slam_b200"), TESTS):
    if p not in sys.path:
        sys.path.insert(0, p)

import ref_conventions as rc          # noqa: E402
from lvdgs import synth               # noqa: E402

OUT = os.path.join(HERE, "reference_pin.npz")
TRACK_CFG = {"Training": {"monocular": True, "rgb_boundary_threshold": 0.01, "alpha": 0.98,
                          "lr": {"cam_rot_delta": 0.003, "cam_trans_delta": 0.001}},
             "Dataset": {"depth_loss": True}}          # configs/mono/KITTI/base_config.yaml:12-15,20-56


def tracking_scene():
    """The small scene of the trk_* section (also rebuilt by the tests from the same seeds)."""
    c = synth.make_camera("mast3r_kitti")
    sc = synth.make_scene(6000, c, seed=5)
    sc["opacities"] = np.clip(sc["opacities"] * 1.5, 0.3, 0.99).astype(np.float32)
    tau0 = np.array([0.03, -0.02, 0.04, np.radians(0.4), np.radians(-0.3), np.radians(0.2)], np.float32)
    rng = np.random.default_rng(17)
    grad_mask = rng.uniform(0, 1, (1, c.image_height, c.image_width)) > 0.35
    return c, sc, tau0, grad_mask


def fd_scene():
    """The tiny scene of the fd_* section (tests/test_oracle_autograd.py uses the same construction)."""
    W, H = 72, 56
    cam = synth.Cam(W, H, 60.0, 58.0, W / 2.0 - 3.3, H / 2.0 + 2.1, np.eye(3), np.zeros(3))
    a = np.radians(7.0)
    cam.R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]) @ \
        np.array([[1, 0, 0], [0, np.cos(0.05), -np.sin(0.05)], [0, np.sin(0.05), np.cos(0.05)]])
    cam.T = np.array([0.1, -0.05, 0.3])
    N = 160
    sc = synth.make_scene(N, cam, seed=3, sh_degree=0, behind_frac=0.05)
    W2C = synth.getWorld2View2(cam.R, cam.T)
    pw = (np.linalg.inv(W2C) @ np.concatenate([sc["means3D"].astype(np.float64), np.ones((N, 1))], 1).T).T[:, :3]
    sc["means3D"] = pw.astype(np.float32)
    sc["scales"] *= 2.0
    sc["opacities"] = np.clip(sc["opacities"], 0.05, 0.9)
    rng = np.random.default_rng(7)
    gc = rng.normal(0, 1, (3, H, W)).astype(np.float32)
    gd = (rng.normal(0, 1, (1, H, W)) * 0.3).astype(np.float32)
    go = rng.normal(0, 1, (1, H, W)).astype(np.float32)
    return cam, sc, gc, gd, go


def loss_case(seed, H=20, W=28):
    rng = np.random.default_rng(seed)
    f = lambda *s: rng.uniform(0, 1, s).astype(np.float32)
    gt = f(3, H, W)
    gt[:, :3, :5] = 0.0                                   # pixels under the rgb boundary threshold
    mono = (f(H, W) * 20).astype(np.float32)
    mono[5:8, 2:9] = 0.0                                  # invalid depth
    return dict(image=f(3, H, W), depth=(f(1, H, W) * 20).astype(np.float32), opacity=f(1, H, W), gt=gt, mono=mono,
                grad_mask=f(1, H, W) > 0.4, a=np.float32(0.07), b=np.float32(-0.02))


def run_losses(su, cu, out):
    variants = {"trk_rgb": ("tracking", True), "trk_rgbd": ("tracking", False), "map_rgb": ("mapping", "rgb"),
                "map_rgbd": ("mapping", "rgbd")}
    for name, (kind, mode) in variants.items():
        for seed in (1, 2):
            d = loss_case(100 + seed)
            t = lambda k: torch.tensor(d[k], requires_grad=True)
            image, depth, opacity = t("image"), t("depth"), t("opacity")
            H, W = d["gt"].shape[1:]
            cam = cu.Camera(0, torch.tensor(d["gt"]), None, d["mono"], torch.eye(4), torch.eye(4), 1., 1., 0., 0., 1., 1., H, W, device="cpu")
            cam.grad_mask = torch.tensor(d["grad_mask"])
            cam.exposure_a.data.fill_(float(d["a"])); cam.exposure_b.data.fill_(float(d["b"]))
            cfg = {"Training": {"monocular": bool(mode is True or mode == "rgb"), "rgb_boundary_threshold": 0.01, "alpha": 0.9},
                   "Dataset": {"depth_loss": False}}
            if kind == "tracking":
                loss = su.get_loss_tracking(cfg, image, depth, opacity, cam)
            else:
                loss = su.get_loss_mapping(cfg, image, cam, depth=depth, monodepth=(mode == "rgbd"))
            loss.backward()
            z = lambda x: np.zeros_like(x.detach().numpy()) if x.grad is None else x.grad.numpy()
            k = f"loss_{name}_{seed}"
            out[k + "_value"] = np.float32(loss.item())
            out[k + "_gimage"], out[k + "_gdepth"], out[k + "_gopacity"] = z(image), z(depth), z(opacity)
            out[k + "_ga"], out[k + "_gb"] = cam.exposure_a.grad.numpy().copy(), cam.exposure_b.grad.numpy().copy()


def run_tracking(pu, su, cu, out, iters=30):
    import oracle_rasterizer
    oracle_rasterizer.install()
    from gaussian_splatting.gaussian_renderer import render
    c, sc, tau0, grad_mask = tracking_scene()
    pc = rc.Gaussians(sc, "cpu")
    bg = torch.zeros(3)
    true_cam = rc.make_camera(cu, c, "cpu")
    with torch.no_grad():
        target = render(true_cam, pc, rc.Pipe(), bg)["render"].clone()
    cam = rc.make_camera(cu, c, "cpu", image=target)
    cam.grad_mask = torch.tensor(grad_mask)
    T0 = pu.SE3_exp(torch.tensor(tau0)) @ torch.eye(4)
    cam.update_RT(T0[:3, :3].contiguous(), T0[:3, 3].contiguous())
    lr = TRACK_CFG["Training"]["lr"]
    opt = torch.optim.Adam([{"params": [cam.cam_rot_delta], "lr": lr["cam_rot_delta"]},
                            {"params": [cam.cam_trans_delta], "lr": lr["cam_trans_delta"]},
                            {"params": [cam.exposure_a], "lr": 0.01}, {"params": [cam.exposure_b], "lr": 0.01}])
    rec = dict(loss=[], g_rot=[], g_trans=[], g_a=[], g_b=[], R=[], T=[], a=[], b=[])
    for it in range(iters):                               # utils/slam_frontend.py:1492-1521
        pkg = render(cam, pc, rc.Pipe(), bg)
        opt.zero_grad()
        loss = su.get_loss_tracking(TRACK_CFG, pkg["render"], pkg["depth"], pkg["opacity"], cam)
        loss.backward()
        rec["loss"].append(loss.item())
        rec["g_rot"].append(cam.cam_rot_delta.grad.numpy().copy()); rec["g_trans"].append(cam.cam_trans_delta.grad.numpy().copy())
        rec["g_a"].append(cam.exposure_a.grad.numpy().copy()); rec["g_b"].append(cam.exposure_b.grad.numpy().copy())
        with torch.no_grad():
            opt.step()
            pu.update_pose(cam)
        rec["R"].append(cam.R.detach().numpy().copy()); rec["T"].append(cam.T.detach().numpy().copy())
        rec["a"].append(cam.exposure_a.detach().numpy().copy()); rec["b"].append(cam.exposure_b.detach().numpy().copy())
    for k, v in rec.items():
        out["trk_" + k] = np.asarray(v, np.float32)
    out["trk_target_sum"] = np.float64(target.double().sum().item())   # the target itself is re-rendered by the tests
    out["trk_true_R"], out["trk_true_T"] = true_cam.R.numpy().copy(), true_cam.T.numpy().copy()
    print("tracking: loss %.5f -> %.5f, |T - T_true| %.4f -> %.4f" % (
        rec["loss"][0], rec["loss"][-1], np.linalg.norm(T0[:3, 3].numpy() - out["trk_true_T"]),
        np.linalg.norm(rec["T"][-1] - out["trk_true_T"])))


def run_fd(pu, out):
    import oracle
    from f64_reference import forward as f64_forward
    cam, sc, gc, gd, go = fd_scene()
    bg = np.array([0.2, 0.5, 0.1], np.float32)
    H, W = cam.image_height, cam.image_width
    fwd = oracle.rasterize_forward(sc["means3D"], sc["opacities"], sc["scales"], sc["rotations"], sc["shs"],
                                   viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform,
                                   campos=cam.camera_center, bg=bg, W=W, H=H, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy)
    dt = torch.float64
    t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), dtype=dt)
    W2C = t(synth.getWorld2View2(cam.R, cam.T))
    Pr = t(synth.getProjectionMatrix2(0.01, 100.0, cam.cx, cam.cy, cam.fx, cam.fy, W, H))
    args = [t(sc[k]) for k in ("means3D", "scales", "rotations", "opacities", "shs")]

    def loss_of(tau, cache=None):
        V = pu.SE3_exp(tau) @ W2C                                   # the REFERENCE's exponential map, float64
        color, dimg, oimg = f64_forward(*args, torch.zeros(6, dtype=dt), W2C=V, Pr=Pr, campos=t(cam.camera_center), bg=bg,
                                        W=W, H=H, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, sh_degree=0, fwd=fwd,
                                        valid_cache=cache)
        return (color * t(gc)).sum() + (dimg * t(gd[0])).sum() + (oimg * t(go[0])).sum()

    # (i) autograd THROUGH the reference's SE3_exp at tau = 0
    tau = torch.zeros(6, dtype=dt, requires_grad=True)
    loss_of(tau).backward()
    g_auto = tau.grad.numpy().copy()
    # (ii) central differences through the same function, blend decisions frozen at their tau = 0 values (the analytic
    # backward, like autograd, does not differentiate the 1/255 and power <= 0 cut-offs)
    cache = {}
    with torch.no_grad():
        loss_of(torch.zeros(6, dtype=dt), cache)
        h, g_fd = 1e-6, np.zeros(6)
        for k in range(6):
            e = torch.zeros(6, dtype=dt); e[k] = h
            g_fd[k] = float(loss_of(e, cache) - loss_of(-e, cache)) / (2 * h)
    assert np.abs(g_fd - g_auto).max() <= 1e-5 * np.abs(g_auto).max(), (g_fd, g_auto)
    out["fd_dL_dtau"] = g_fd                                        # [rho(3); theta(3)], utils/pose_utils.py:59-60
    out["fd_dL_dtau_autograd"] = g_auto
    out["fd_n_contrib_sum"] = np.int64(fwd["n_contrib"].sum())      # guards against a silently different scene
    print("fd dL/dtau:", g_fd, "autograd:", g_auto)


def main():
    pu, su, cu = rc.import_reference(cpu=True)
    out = {}
    rng = np.random.default_rng(0)
    # ---- SE3_exp / SO3_exp / V ----
    taus = rng.normal(0, 0.3, (10, 6)).astype(np.float32)
    taus[0] = 0.0
    taus[1, 3:] = [3e-6, -2e-6, 1e-6]                     # small-angle branch (angle < 1e-5)
    taus[2, 3:] = [0.0, 0.0, 2.5]                         # large rotation
    out["se3_tau"] = taus
    out["se3_T"] = np.stack([pu.SE3_exp(torch.tensor(x)).numpy() for x in taus])
    out["so3_R"] = np.stack([pu.SO3_exp(torch.tensor(x[3:])).numpy() for x in taus])
    out["so3_V"] = np.stack([pu.V(torch.tensor(x[3:])).numpy() for x in taus])
    # ---- update_pose + Camera matrices ----
    cams = []
    for k in range(4):
        c = synth.make_camera("kitti", k=k + 1)
        cam = rc.make_camera(cu, c, "cpu")
        d_rot = (rng.normal(0, 0.02, 3)).astype(np.float32) if k else np.array([1e-6, 2e-6, -1e-6], np.float32)
        d_tr = (rng.normal(0, 0.05, 3)).astype(np.float32) if k else np.array([2e-5, 1e-5, 3e-5], np.float32)
        rec = dict(R0=cam.R.numpy().copy(), T0=cam.T.numpy().copy(), d_rot=d_rot, d_tr=d_tr,
                   wvt0=cam.world_view_transform.numpy().copy(), fpt0=cam.full_proj_transform.numpy().copy(),
                   cc0=cam.camera_center.numpy().copy(), proj=cam.projection_matrix.numpy().copy())
        cam.cam_rot_delta.data[:] = torch.tensor(d_rot); cam.cam_trans_delta.data[:] = torch.tensor(d_tr)
        with torch.no_grad():
            conv = pu.update_pose(cam)
        rec.update(R1=cam.R.detach().numpy().copy(), T1=cam.T.detach().numpy().copy(), conv=np.bool_(bool(conv)),
                   wvt1=cam.world_view_transform.detach().numpy().copy(), fpt1=cam.full_proj_transform.detach().numpy().copy(),
                   cc1=cam.camera_center.detach().numpy().copy(),
                   deltas_after=np.concatenate([cam.cam_rot_delta.detach().numpy(), cam.cam_trans_delta.detach().numpy()]))
        cams.append(rec)
    for k in cams[0]:
        out["up_" + k] = np.stack([c[k] for c in cams])
    run_losses(su, cu, out)
    run_fd(pu, out)
    run_tracking(pu, su, cu, out)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
