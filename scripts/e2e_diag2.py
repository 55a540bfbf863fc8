"""Where does host time go in the plugin path?  Times the C calls vs the Python around them."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lvd_gs-slam_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from lvdgs import synth, _native
import diff_gaussian_rasterization as dgr
from gpu_harness import settings_for
dev = "cuda"
cams = [synth.make_camera("kitti", k) for k in range(8)]
sc = synth.make_scene(500_000, cams[0], seed=0)
t = lambda a: torch.tensor(a, device=dev, requires_grad=True)
params = [t(sc[k]) for k in ("means3D", "opacities", "scales", "rotations", "shs")]
H, W = cams[0].image_height, cams[0].image_width
img = torch.rand(3, H, W, device=dev); dep = torch.rand(1, H, W, device=dev) * 50
rss = [settings_for(c, (0, 0, 0), 0) for c in cams]
L = _native.lib()
acc = {"cf": 0.0, "cb": 0.0}
of, ob = L.lvdgs_rasterize_forward, L.lvdgs_rasterize_backward
class Wrap:
    def __init__(self, f, key): self.f, self.key = f, key
    def __call__(self, *a):
        t0 = time.perf_counter(); r = self.f(*a); acc[self.key] += time.perf_counter() - t0; return r
L.lvdgs_rasterize_forward = Wrap(of, "cf"); L.lvdgs_rasterize_backward = Wrap(ob, "cb")
def step():
    tt = dict(fwd=0.0, loss=0.0, bwd=0.0)
    for p in params: p.grad = None
    for k in range(8):
        t0 = time.perf_counter()
        theta = torch.zeros(3, device=dev, requires_grad=True); rho = torch.zeros(3, device=dev, requires_grad=True)
        m2d = torch.zeros_like(params[0], requires_grad=True)
        color, radii, depth, opacity, n_touched = dgr.GaussianRasterizer(rss[k])(means3D=params[0], means2D=m2d, opacities=params[1], shs=params[4], scales=params[2], rotations=params[3], theta=theta, rho=rho)
        t1 = time.perf_counter()
        loss = 0.9 * (color - img).abs().mean() + 0.1 * (depth - dep).abs().mean()
        t2 = time.perf_counter()
        loss.backward()
        t3 = time.perf_counter()
        tt["fwd"] += t1 - t0; tt["loss"] += t2 - t1; tt["bwd"] += t3 - t2
    torch.cuda.synchronize()
    return tt
for i in range(8):
    acc["cf"] = acc["cb"] = 0.0
    t0 = time.perf_counter(); tt = step(); dt = time.perf_counter() - t0
    if i >= 4:
        print(f"step {dt*1e3:6.2f} ms | per view: fwd_py {tt['fwd']/8*1e3:.3f} (C call {acc['cf']/8*1e3:.3f}) loss_py {tt['loss']/8*1e3:.3f} bwd_py {tt['bwd']/8*1e3:.3f} (C call {acc['cb']/8*1e3:.3f})")
