"""cProfile of the plugin path (render + torch loss + backward) to see where the HOST time of the e2e step goes."""
import cProfile, pstats, os, sys, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lvd_gs-slam_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from lvdgs import synth
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
def step():
    for p in params: p.grad = None
    for k in range(8):
        theta = torch.zeros(3, device=dev, requires_grad=True); rho = torch.zeros(3, device=dev, requires_grad=True)
        m2d = torch.zeros_like(params[0], requires_grad=True)
        color, radii, depth, opacity, n_touched = dgr.GaussianRasterizer(rss[k])(means3D=params[0], means2D=m2d, opacities=params[1], shs=params[4], scales=params[2], rotations=params[3], theta=theta, rho=rho)
        loss = 0.9 * (color - img).abs().mean() + 0.1 * (depth - dep).abs().mean()
        loss.backward()
    torch.cuda.synchronize()
for _ in range(5): step()
pr = cProfile.Profile(); pr.enable()
for _ in range(20): step()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(35); print(s.getvalue()[:6000])
