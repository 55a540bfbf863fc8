"""Per-kernel device times of the device-resident tracking loop (BASELINE configs[1]: 300k Gaussians, 1241x376)."""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lvd_gs-slam_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from lvdgs import synth, tracking as trk, _native
from gaussian_splatting.gaussian_renderer import render
from test_gpu_shim_tracking import Cam, Gaussians, Pipe, SE3_exp
dev = "cuda"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
c = synth.make_camera("kitti")
pc = Gaussians(synth.make_scene(N, c, seed=0), dev)
true_cam = Cam(c, dev)
with torch.no_grad():
    target = render(true_cam, pc, Pipe(), torch.zeros(3, device=dev))["render"].clone()
tr = trk.PoseTracker(N, c.image_width, c.image_height, c.tanfovx, c.tanfovy, device=dev, rgb_boundary_threshold=-1.0)
T0 = SE3_exp(torch.tensor([0.02, -0.01, 0.03, math.radians(0.3), math.radians(-0.2), math.radians(0.1)], device=dev))
args = (pc.get_xyz, pc.get_opacity, pc.get_scaling, pc.get_rotation, pc.get_features, target)
tr.set_camera(T0[:3, :3], T0[:3, 3], true_cam.projection_matrix)
tr.track(*args, iters=20, stop_when_converged=False)
torch.cuda.synchronize()
t0 = time.perf_counter(); tr.track(*args, iters=100, stop_when_converged=False); torch.cuda.synchronize()
print("ms/iter", (time.perf_counter() - t0) * 10)
stream = torch.cuda.current_stream().cuda_stream
_native.profile_begin(stream)
tr.track(*args, iters=10, stop_when_converged=False)
agg = {}
for k, v in _native.profile_end(stream):
    agg[k] = agg.get(k, 0.0) + v / 10
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print(f"{k:28s} {v*1e3:8.1f} us/iter")
print("sum", sum(agg.values()) * 1e3)
