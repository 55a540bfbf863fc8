"""BASELINE configs[1]: KITTI-shaped tracking -- 1241x376, 300k Gaussians, the per-frame pose-gradient render loop
(100 iterations of render -> tracking loss -> backward -> Adam on (cam_rot_delta, cam_trans_delta) -> update_pose),
through the reference-facing shim gaussian_splatting.gaussian_renderer.render, as utils/slam_frontend.py:1468-1533 does.
Early exit disabled for timing.  Prints one JSON line."""
import json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lvd_gs-slam_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from lvdgs import synth
from gaussian_splatting.gaussian_renderer import render
from test_gpu_shim_tracking import Cam, Gaussians, Pipe, update_pose, SE3_exp

dev = "cuda"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
c = synth.make_camera("kitti")
sc = synth.make_scene(N, c, seed=0)
pc = Gaussians(sc, dev)
bg = torch.zeros(3, device=dev)
with torch.no_grad():
    target = render(Cam(c, dev), pc, Pipe(), bg)["render"].clone()
H, W = c.image_height, c.image_width

def track_frame():
    cam = Cam(c, dev)
    tau0 = torch.tensor([0.02, -0.01, 0.03, math.radians(0.3), math.radians(-0.2), math.radians(0.1)], device=dev)
    T0 = SE3_exp(tau0)
    cam.R, cam.T = T0[:3, :3].contiguous(), T0[:3, 3].contiguous()
    opt = torch.optim.Adam([{"params": [cam.cam_rot_delta], "lr": 0.003}, {"params": [cam.cam_trans_delta], "lr": 0.001}])
    for it in range(iters):
        pkg = render(cam, pc, Pipe(), bg)
        loss = (pkg["opacity"] * (pkg["render"] - target).abs()).mean()
        opt.zero_grad()
        loss.backward()
        with torch.no_grad():
            opt.step()
            update_pose(cam)
    return float(loss)

track_frame()                                   # warm-up frame
torch.cuda.synchronize()
ts = []
for _ in range(3):
    t0 = time.perf_counter(); l = track_frame(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
t = float(np.median(ts))
print(json.dumps({"workload": "kitti_tracking", "gaussians": N, "image": [W, H], "iters_per_frame": iters,
                  "ms_per_iter": 1e3 * t / iters, "iters_per_s": iters / t, "mpix_per_s": iters * H * W / t / 1e6,
                  "frame_ms": 1e3 * t, "final_loss": l}))
