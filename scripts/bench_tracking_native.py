"""BASELINE configs[1] through the device-resident loop (lvdgs.tracking.PoseTracker); compare scripts/bench_tracking.py
(the same loop through the plugin surface + torch).  Prints one JSON line."""
import json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lvd_gs-slam_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from lvdgs import synth, tracking as trk
from gaussian_splatting.gaussian_renderer import render
from test_gpu_shim_tracking import Cam, Gaussians, Pipe, SE3_exp

dev = "cuda"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
c = synth.make_camera("kitti")
pc = Gaussians(synth.make_scene(N, c, seed=0), dev)
true_cam = Cam(c, dev)
with torch.no_grad():
    target = render(true_cam, pc, Pipe(), torch.zeros(3, device=dev))["render"].clone()
H, W = c.image_height, c.image_width
tr = trk.PoseTracker(N, W, H, c.tanfovx, c.tanfovy, device=dev, lr_rot=0.003, lr_trans=0.001, rgb_boundary_threshold=-1.0)
T0 = SE3_exp(torch.tensor([0.02, -0.01, 0.03, math.radians(0.3), math.radians(-0.2), math.radians(0.1)], device=dev))

def track_frame():
    tr.set_camera(T0[:3, :3], T0[:3, 3], true_cam.projection_matrix)
    return tr.track(pc.get_xyz, pc.get_opacity, pc.get_scaling, pc.get_rotation, pc.get_features, target, iters=iters,
                    stop_when_converged=False)

track_frame(); torch.cuda.synchronize()
ts = []
for _ in range(3):
    t0 = time.perf_counter(); out = track_frame(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
t = float(np.median(ts))
print(json.dumps({"workload": "kitti_tracking_native", "gaussians": N, "image": [W, H], "iters_per_frame": iters,
                  "ms_per_iter": 1e3 * t / iters, "iters_per_s": iters / t, "mpix_per_s": iters * H * W / t / 1e6,
                  "frame_ms": 1e3 * t, "final_loss": float(out["loss"])}))
