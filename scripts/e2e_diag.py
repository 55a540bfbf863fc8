"""Diagnose host-side cost of the plugin path: per-step wall times and allocator activity."""
import os, sys, time
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
def step(detail=False):
    for p in params: p.grad = None
    ts = []
    for k in range(8):
        t0 = time.perf_counter()
        theta = torch.zeros(3, device=dev, requires_grad=True); rho = torch.zeros(3, device=dev, requires_grad=True)
        m2d = torch.zeros_like(params[0], requires_grad=True)
        color, radii, depth, opacity, n_touched = dgr.GaussianRasterizer(rss[k])(means3D=params[0], means2D=m2d, opacities=params[1], shs=params[4], scales=params[2], rotations=params[3], theta=theta, rho=rho)
        t1 = time.perf_counter()
        loss = 0.9 * (color - img).abs().mean() + 0.1 * (depth - dep).abs().mean()
        loss.backward()
        t2 = time.perf_counter()
        ts.append((t1 - t0, t2 - t1))
    torch.cuda.synchronize()
    return ts
for i in range(12):
    st = torch.cuda.memory_stats()
    a0 = st.get("num_device_alloc", 0); r0 = st.get("num_alloc_retries", 0)
    t0 = time.perf_counter(); ts = step(); dt = time.perf_counter() - t0
    st = torch.cuda.memory_stats()
    print(f"step {i}: {dt*1e3:7.2f} ms  cudaMallocs {st.get('num_device_alloc',0)-a0} retries {st.get('num_alloc_retries',0)-r0} reserved {torch.cuda.memory_reserved()/2**20:.0f} MiB  fwd_host {sum(a for a,_ in ts)*1e3:.2f} bwd_host {sum(b for _,b in ts)*1e3:.2f}")
