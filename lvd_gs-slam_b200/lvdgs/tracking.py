"""PoseTracker -- the frontend's per-frame tracking loop (utils/slam_frontend.py:1466-1533) as a device-resident loop over
the C ABI.

The reference runs, per iteration: render() -> get_loss_tracking -> loss.backward() -> torch.optim.Adam.step on
(cam_rot_delta, cam_trans_delta, exposure_a, exposure_b) -> update_pose (utils/pose_utils.py:70-87), i.e. the rasterizer
plus ~100 small torch kernels, a fresh [N,3] zeros tensor, autograd bookkeeping and two extra host synchronisations.
Here one iteration is 6 + 1 + 3 + 1 launches with no allocation:
    lvdgs_rasterize_forward  -> lvdgs_fused_loss (tracking rgb loss, exposure, grad_mask; dL/dcolor and dL/dexposure)
    -> lvdgs_rasterize_backward (LVDGS_FLAG_POSE_ONLY: dL/dtau only) -> lvdgs_pose_step (Adam + SE3 update + camera matrices)
on a camera block that lives on the device (lvdgs_pose_state); the only host wait is the rasterizer's read-back of the
instance count, and the convergence flag of the previous step rides along with it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native
from .engine import RasterEngine

# offsets (in floats) inside lvdgs_pose_state (include/lvdgs.h)
_VIEW, _PROJ, _PRAW, _CAMPOS, _R, _T, _EXPO, _M, _V, _STEP, _CONV, _TAUN, _SIZE = 0, 16, 32, 48, 52, 61, 64, 68, 76, 84, 85, 86, 88
LOSS_OPACITY_WEIGHT = 1
LOSS_DEPTH_NEEDS_OPAQUE = 2


class _StateCamera:
    """The camera tensors the engine reads, as views into the device-resident pose state."""

    def __init__(self, state, W, H, tanfovx, tanfovy, bg):
        self.W, self.H, self.tanfovx, self.tanfovy = int(W), int(H), float(tanfovx), float(tanfovy)
        self.view, self.proj, self.proj_raw = state[_VIEW:_VIEW + 16], state[_PROJ:_PROJ + 16], state[_PRAW:_PRAW + 16]
        self.campos, self.bg = state[_CAMPOS:_CAMPOS + 3], bg


class PoseTracker:
    def __init__(self, P, W, H, tanfovx, tanfovy, sh_coeffs=1, sh_degree=0, device="cuda", lr_rot=0.003, lr_trans=0.001,
                 lr_exposure=0.01, betas=(0.9, 0.999), eps=1e-8, converged_threshold=1e-4, optimise_exposure=True,
                 rgb_boundary_threshold=0.01, bg=(0.0, 0.0, 0.0)):
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("PoseTracker: CUDA only (there is no CPU path)")
        self.L = _native.lib()
        self.eng = RasterEngine(P, W, H, sh_coeffs=sh_coeffs, sh_degree=sh_degree, device=self.dev, slots=1)
        self.W, self.H = W, H
        self.lr = (float(lr_rot), float(lr_trans), float(lr_exposure))
        self.betas, self.eps, self.threshold = betas, float(eps), float(converged_threshold)
        self.optimise_exposure = optimise_exposure
        self.rgb_thr = float(rgb_boundary_threshold)
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.state = torch.zeros(_SIZE, **f32)
        self.bg = torch.tensor(bg, **f32)
        self.cam = _StateCamera(self.state, W, H, tanfovx, tanfovy, self.bg)
        self.g_color = torch.empty(3, H, W, **f32)
        self.g_depth = None                      # allocated by the first RGB-D frame
        self.loss_out = torch.zeros(4, **f32)
        self.loss_ws = torch.zeros(self.L.lvdgs_fused_loss_workspace_bytes(), dtype=torch.uint8, device=self.dev)
        self._flag_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._state_i32 = self.state.view(torch.int32)

    # ---- camera ----
    def set_camera(self, R, T, projection_matrix, exposure_a=0.0, exposure_b=0.0):
        """R, T: world -> camera (utils/camera_utils.py: Camera.R / Camera.T); projection_matrix: the camera's
        `projection_matrix` (P^T, utils/slam_frontend.py:1743-1749).  Resets the optimiser state, like the reference's
        fresh torch.optim.Adam per tracked frame."""
        R = np.asarray(R.detach().cpu() if torch.is_tensor(R) else R, dtype=np.float64).reshape(3, 3)
        T = np.asarray(T.detach().cpu() if torch.is_tensor(T) else T, dtype=np.float64).reshape(3)
        praw = np.asarray(projection_matrix.detach().cpu() if torch.is_tensor(projection_matrix) else projection_matrix,
                          dtype=np.float64).reshape(4, 4)
        w2c = np.eye(4); w2c[:3, :3] = R; w2c[:3, 3] = T
        view = w2c.T
        st = np.zeros(_SIZE, np.float32)
        st[_VIEW:_VIEW + 16] = view.reshape(-1)
        st[_PROJ:_PROJ + 16] = (view @ praw).reshape(-1)
        st[_PRAW:_PRAW + 16] = praw.reshape(-1)
        st[_CAMPOS:_CAMPOS + 3] = -R.T @ T
        st[_R:_R + 9] = R.reshape(-1)
        st[_T:_T + 3] = T
        st[_EXPO], st[_EXPO + 1] = float(exposure_a), float(exposure_b)      # python floats or 1-element tensors (Camera.exposure_a/b)
        self.state.copy_(torch.from_numpy(st))

    @property
    def R(self):
        return self.state[_R:_R + 9].view(3, 3)

    @property
    def T(self):
        return self.state[_T:_T + 3]

    @property
    def exposure(self):
        return self.state[_EXPO:_EXPO + 2]

    # ---- one frame ----
    def track(self, means3D, opacities, scales, rotations, shs, gt_image, grad_mask=None, iters=100,
              stop_when_converged=True, gt_depth=None, alpha=0.95):
        """Runs up to `iters` tracking iterations against `gt_image` [3,H,W] and returns a dict with the number of pose
        steps taken, the last loss (device scalar), and the engine's render / depth / opacity of the last forward.
        gt_depth None (the monocular configs LVD-GS ships: get_loss_tracking always returns get_loss_tracking_rgb for
        them, utils/slam_utils.py:45-49): the rgb loss of utils/slam_utils.py:53-62 with the exposure model of :43.
        gt_depth [H,W] or [1,H,W] (RGB-D configs): get_loss_tracking_rgbd (:65-83), alpha * rgb + (1 - alpha) * depth term
        over the pixels with gt_depth > 0.01 and rendered opacity > 0.95; the depth gradient then enters the pose-only
        backward (ten sums per (warp, Gaussian) instead of the six moments of the rgb-only loop).
        Difference from the reference loop: when the step converges, the returned render is one forward at the UPDATED
        pose (the reference returns the render_pkg from before the last pose update, utils/slam_frontend.py:1523-1533)."""
        L, eng, cam, p = self.L, self.eng, self.cam, _native.ptr
        stream = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        gt = gt_image if (gt_image.dtype is torch.float32 and gt_image.is_contiguous()) else gt_image.contiguous().float()
        gm = None if grad_mask is None else grad_mask.contiguous().float()
        gtd = None
        w_rgb, w_depth, flags = 1.0, 0.0, LOSS_OPACITY_WEIGHT
        if gt_depth is not None:
            gtd = gt_depth.reshape(self.H, self.W).contiguous().float()
            w_rgb, w_depth, flags = float(alpha), 1.0 - float(alpha), LOSS_OPACITY_WEIGHT | LOSS_DEPTH_NEEDS_OPAQUE
            if self.g_depth is None:
                self.g_depth = torch.empty(self.H, self.W, dtype=torch.float32, device=self.dev)
        g_depth = self.g_depth if gtd is not None else None
        expo = self.state[_EXPO:_EXPO + 2]
        g_expo = self.loss_out[1:3] if self.optimise_exposure else None
        sl = eng.slots[0]
        steps = 0
        self._flag_host.zero_()
        for it in range(1, iters + 1):
            eng.forward(cam, means3D, opacities, scales, rotations, shs)        # host waits for R here ...
            if stop_when_converged and steps > 0 and int(self._flag_host[0]) != 0:
                break                                                           # ... so the previous step's flag has landed
            _native.check(L.lvdgs_fused_loss(self.W, self.H, p(sl.color), p(sl.depth) if gtd is not None else None, p(sl.opacity),
                                             p(gt), p(gtd), p(gm), p(expo), self.rgb_thr, w_rgb, w_depth, flags, p(self.g_color),
                                             p(g_depth), None, p(self.loss_out), p(self.loss_ws), self.loss_ws.numel(), stream),
                          "lvdgs_fused_loss")
            eng.backward(cam, means3D, opacities, scales, rotations, shs, self.g_color, g_depth, None, pose_only=True)
            _native.check(L.lvdgs_pose_step(p(self.state), p(sl.g_tau), p(g_expo), self.lr[0], self.lr[1], self.lr[2],
                                            self.betas[0], self.betas[1], self.eps, it, self.threshold, stream),
                          "lvdgs_pose_step")
            steps = it
            self._flag_host.copy_(self._state_i32[_CONV:_CONV + 1], non_blocking=True)
        return {"steps": steps, "loss": self.loss_out[0], "render": sl.color, "depth": sl.depth, "opacity": sl.opacity,
                "radii": sl.radii, "n_touched": sl.n_touched}
