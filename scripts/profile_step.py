"""Per-kernel device times of one fwd+bwd (CUDA events after every launch, lvdgs_profile_*).  Usage:
python scripts/profile_step.py [N] [camera] [iters]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lvd_gs-slam_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from lvdgs import synth, _native
from gpu_harness import settings_for
import diff_gaussian_rasterization as dgr

N = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
camname = sys.argv[2] if len(sys.argv) > 2 else "kitti"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
cam = synth.make_camera(camname)
sc = synth.make_scene(N, cam, seed=0)
dev = "cuda"
t = lambda a: torch.tensor(a, device=dev, requires_grad=True)
means3D, opac, scales, rots, shs = t(sc["means3D"]), t(sc["opacities"]), t(sc["scales"]), t(sc["rotations"]), t(sc["shs"])
theta = torch.zeros(3, device=dev, requires_grad=True); rho = torch.zeros(3, device=dev, requires_grad=True)
gc, gd = synth.make_upstream_grads(cam)
gc = torch.tensor(gc, device=dev); gd = torch.tensor(gd, device=dev)
rast = dgr.GaussianRasterizer(settings_for(cam, (0, 0, 0), 0))
stream = torch.cuda.current_stream().cuda_stream

def step():
    m2d = torch.zeros_like(means3D, requires_grad=True)
    color, radii, depth, opacity, n_touched = rast(means3D=means3D, means2D=m2d, opacities=opac, shs=shs, scales=scales,
                                                   rotations=rots, theta=theta, rho=rho)
    torch.autograd.backward([color, depth], [gc, gd])
    return color, radii, n_touched

for _ in range(3):
    color, radii, n_touched = step()
torch.cuda.synchronize()
R = color.grad_fn.num_rendered if color.grad_fn is not None else -1
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
H, W = cam.image_height, cam.image_width
agg = {}
for _ in range(iters):
    _native.profile_begin(stream)
    step()
    for name, t_ms in _native.profile_end(stream):
        agg.setdefault(name, []).append(t_ms)
print(json.dumps(dict(N=N, cam=camname, R=int(R), visible=int((radii > 0).sum()), ms_fwd_bwd=ms, mpix_s=H * W / ms / 1e3)))
tot = 0
for name, v in agg.items():
    per_step = sum(v) / iters
    tot += per_step
    print(f"{name:28s} launches/step {len(v)/iters:5.1f}  ms/step {per_step:8.4f}")
print(f"{'sum':28s} {tot:8.4f} ms")
