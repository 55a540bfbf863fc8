"""Where the plugin-path mapping step (bench.py step_e2e) spends its time: device time (CUDA events) and host wall time
of the three phases -- 8 x (render + loss), one backward through all views, optimiser step.
python scripts/e2e_breakdown.py [steps] [N] [camera]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lvd_gs-slam_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from lvdgs import synth, slam_ops
import diff_gaussian_rasterization as dgr
from gpu_harness import settings_for
dev = "cuda"
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
N = int(sys.argv[2]) if len(sys.argv) > 2 else 500_000      # a tiny N (2000) leaves only the host overhead
camname = sys.argv[3] if len(sys.argv) > 3 else "kitti"
cams = [synth.make_camera(camname, k) for k in range(8)]
sc = synth.make_scene(N, cams[0], seed=0)
t = lambda a: torch.tensor(a, device=dev, requires_grad=True)
params = [t(sc[k]) for k in ("means3D", "opacities", "scales", "rotations", "shs")]
opt = torch.optim.Adam(params, lr=1e-8)
H, W = cams[0].image_height, cams[0].image_width
img = torch.rand(3, H, W, device=dev); dep = torch.rand(1, H, W, device=dev) * 50
rss = [settings_for(c, (0, 0, 0), 0) for c in cams]
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
def step(rec=None):
    opt.zero_grad(set_to_none=True)
    w0 = time.perf_counter(); ev[0].record()
    total = None
    for k in range(8):
        theta = torch.zeros(3, device=dev, requires_grad=True); rho = torch.zeros(3, device=dev, requires_grad=True)
        m2d = torch.zeros_like(params[0], requires_grad=True)
        color, radii, depth, opacity, n_touched = dgr.GaussianRasterizer(rss[k])(means3D=params[0], means2D=m2d, opacities=params[1], shs=params[4], scales=params[2], rotations=params[3], theta=theta, rho=rho)
        loss = slam_ops.fused_loss(color, depth, gt_image=img, gt_depth=dep, rgb_boundary_threshold=-1.0, w_rgb=0.9, w_depth=0.1)
        total = loss if total is None else total + loss
    w1 = time.perf_counter(); ev[1].record()
    total.backward()
    w2 = time.perf_counter(); ev[2].record()
    opt.step()
    w3 = time.perf_counter(); ev[3].record()
    torch.cuda.synchronize()
    w4 = time.perf_counter()
    if rec is not None:
        rec.append(dict(dev_fwd=ev[0].elapsed_time(ev[1]), dev_bwd=ev[1].elapsed_time(ev[2]), dev_opt=ev[2].elapsed_time(ev[3]),
                        host_fwd=(w1 - w0) * 1e3, host_bwd=(w2 - w1) * 1e3, host_opt=(w3 - w2) * 1e3, wall=(w4 - w0) * 1e3))
for _ in range(5): step()
rec = []
for _ in range(steps): step(rec)
print(json.dumps({'N': N, 'cam': camname, **{k: float(np.median([r[k] for r in rec])) for k in rec[0]}}))
