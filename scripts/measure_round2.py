"""Round-2 measurements that BASELINE.json's configs[3] / configs[4] and BASELINE.md section 3 ask for, written as JSON
(gpurun_out/r02_measurements.json -> committed as profiles/r02_measurements.json).  One B200, CUDA events, median of the
timed repetitions after warm-up; nvidia-smi clocks / throttle reasons sampled before and after.

  sort        K4 comparator: the hand-written onesweep (lvdgs_sort_pairs) vs cub::DeviceRadixSort::SortPairs (the library
              call the reference makes) on (tile | depth) keys, n = 1.34 M ... 60 M; plus, inside `sweep`, emission + sort
              of the default tile-segment path against the onesweep path on real frames
  nuscenes    configs[3]: 1600x900, 2 M Gaussians, forward + backward per kernel
  dist2       configs[3] map initialisation: simple_knn.distCUDA2 at 7.3 k / 14.6 k (the reference's per-keyframe sizes,
              configs/mono/KITTI/base_config.yaml:16-17) and 2 M points
  sweep       configs[4]: 1920x1080, N = 100 k ... 8 M Gaussians, mapping iterations (fwd + bwd + exchange_and_update) WITH
              densify / prune churn between iterations: every iteration clones 1 %, splits 1 % and prunes 2 % of the map
              (seeded), so every per-Gaussian array changes size

Usage (GPU box): python scripts/measure_round2.py [sort] [nuscenes] [dist2] [sweep]
"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lvd_gs-slam_b200"), os.path.join(ROOT, "scripts")]
import numpy as np
import torch

from lvdgs import _native, synth
from lvdgs.engine import RasterEngine, ViewCamera
from lvdgs.mapping import ShardedMapper

L = _native.lib()
dev = torch.device("cuda")
what = set(sys.argv[1:]) or {"sort", "nuscenes", "dist2", "sweep"}
out = {}


def clocks():
    q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    r = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
    return r


def timed(fn, warm=3, reps=7):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


out["clocks_before"] = clocks()

if "sort" in what:
    import importlib.util
    rows = []
    p = _native.ptr
    for n in (1_340_000, 5_000_000, 20_000_000, 60_000_000):
        rng = np.random.default_rng(0)
        tiles = rng.integers(0, 1872, n, dtype=np.uint64)
        depth = np.exp(rng.uniform(np.log(0.2), np.log(100.0), n)).astype(np.float32).view(np.uint32).astype(np.uint64)
        keys = torch.from_numpy(((tiles << np.uint64(32)) | depth).view(np.int64)).to(dev)
        vals = torch.arange(n, dtype=torch.int32, device=dev)
        k0, k1, v0, v1 = keys.clone(), torch.empty_like(keys), vals.clone(), torch.empty_like(vals)
        ws = torch.empty(L.lvdgs_sort_workspace_bytes(n), dtype=torch.uint8, device=dev)
        wc = torch.empty(max(1, L.lvdgs_cub_sort_workspace_bytes(n, 43)), dtype=torch.uint8, device=dev)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        sel = C.c_int32(0)
        res = {"n": n}
        for name in ("onesweep_ours", "cub_device_radix_sort"):
            ts = []
            for it in range(10):
                k0.copy_(keys); v0.copy_(vals)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if name == "onesweep_ours":
                    L.lvdgs_sort_pairs(n, p(k0), p(k1), p(v0), p(v1), 43, p(ws), ws.numel(), C.byref(sel), st)
                else:
                    L.lvdgs_cub_sort_pairs(n, p(k0), p(k1), p(v0), p(v1), 43, p(wc), wc.numel(), st)
                e1.record(); e1.synchronize()
                if it >= 3:
                    ts.append(e0.elapsed_time(e1))
            res[name + "_ms"] = float(np.median(ts))
        res["onesweep_GBs_algorithmic"] = (8 + 6 * 24) * n / (res["onesweep_ours_ms"] * 1e-3) / 1e9
        res["speedup_vs_cub"] = res["cub_device_radix_sort_ms"] / res["onesweep_ours_ms"]
        rows.append(res)
        print("sort", json.dumps(res), flush=True)
        del keys, vals, k0, k1, v0, v1, ws, wc
    out["sort"] = rows

if "nuscenes" in what:
    cam = synth.make_camera("nuscenes")
    N = 2_000_000
    sc = synth.make_scene(N, cam, seed=0)
    t = lambda a: torch.tensor(a, device=dev)
    m, o, s, r, sh = t(sc["means3D"]), t(sc["opacities"]), t(sc["scales"]), t(sc["rotations"]), t(sc["shs"])
    H, W = cam.image_height, cam.image_width
    gc, gd = synth.make_upstream_grads(cam)
    gc, gd = t(gc), t(gd)
    vc = ViewCamera(cam, dev)
    eng = RasterEngine(N, W, H, device=dev, slots=1)

    def step():
        eng.forward(vc, m, o, s, r, sh); eng.backward(vc, m, o, s, r, sh, gc, gd)
    ms = timed(step)
    stream = torch.cuda.current_stream().cuda_stream
    prof = {}
    for _ in range(3):
        _native.profile_begin(stream)
        step()
        for k, v in _native.profile_end(stream):
            prof[k] = prof.get(k, 0.0) + v / 3
    out["nuscenes_2m"] = dict(workload="BASELINE configs[3]: 1600x900, 2 M Gaussians, depth + opacity outputs, fwd+bwd", ms_fwd_bwd=ms,
                              mpix_per_s=H * W / ms / 1e3, R=int(eng.R), visible=int((eng.radii > 0).sum()), pairs=eng.pair_count(),
                              kernels_ms={k: round(v, 4) for k, v in prof.items()})
    print("nuscenes", json.dumps(out["nuscenes_2m"]), flush=True)
    del eng, m, o, s, r, sh

if "dist2" in what:
    from simple_knn._C import distCUDA2
    rows = []
    for n, how in ((7_291, "one KITTI keyframe, pcd_downsample 64"), (14_582, "initialisation, pcd_downsample_init 32"),
                   (2_000_000, "BASELINE configs[3]: the whole 2 M map")):
        cam = synth.make_camera("kitti")
        pts = torch.tensor(synth.make_scene(n, cam, seed=1)["means3D"], device=dev)
        ms = timed(lambda: distCUDA2(pts), warm=2, reps=5)
        rows.append(dict(points=n, what=how, ms=ms, mpoints_per_s=n / ms / 1e3, GBs_algorithmic=16.0 * n / (ms * 1e-3) / 1e9))
        print("dist2", json.dumps(rows[-1]), flush=True)
    out["dist2"] = rows

if "sweep" in what:
    cam = synth.make_camera("hd")
    H, W = cam.image_height, cam.image_width
    gc, gd = synth.make_upstream_grads(cam)
    gc, gd = torch.tensor(gc, device=dev), torch.tensor(gd, device=dev)
    vc = ViewCamera(cam, dev)
    rows = []
    for N in (100_000, 500_000, 2_000_000, 8_000_000):
        sc = synth.make_scene(N, cam, seed=0)
        mapper = ShardedMapper(N, sh_coeffs=1, device=dev, lrs={k: v * 1e-4 for k, v in
                                                               dict(means3D=1.6e-4, shs=2.5e-3, opacity=5e-2, scales=1e-3, rotations=1e-3).items()})
        mapper.load(means3D=sc["means3D"], shs=sc["shs"], opacity=sc["opacities"], scales=sc["scales"], rotations=sc["rotations"])
        eng = RasterEngine(N, W, H, device=dev, slots=1, grad_flat=mapper.new_grad_block())
        gdev = torch.Generator(device=dev).manual_seed(N)
        t_render, t_churn, sizes, Rs = [], [], [], []
        for it in range(8):
            args = [mapper.view(k) for k in ("means3D", "opacity", "scales", "rotations", "shs")]
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            eng.forward(vc, *args); eng.backward(vc, *args, gc, gd, accumulate=True)
            mapper.exchange_and_update(eng.grad_flat)
            e1.record()
            torch.cuda.synchronize()
            t_host = time.perf_counter()
            # churn: clone 1 %, split 1 % (two children at perturbed positions, scales / 1.6, parent removed), prune 2 %
            P = mapper.P
            perm = torch.randperm(P, generator=gdev, device=dev)
            n1 = max(1, P // 100)
            n_prune = 2 * n1 if it % 2 == 0 else n1            # the map breathes: -0 % / +1 % alternately
            clone_idx, split_idx, prune_idx = perm[:n1], perm[n1:2 * n1], perm[2 * n1:2 * n1 + n_prune]
            mapper.densify_clone(clone_idx)
            rep = split_idx.repeat(2)
            new_xyz = mapper.view("means3D")[rep] + 0.01 * torch.randn(rep.numel(), 3, generator=gdev, device=dev)
            mapper.densify_clone(rep, overrides={"means3D": new_xyz, "scales": torch.log(mapper.view("scales")[rep] / 1.6)})
            keep = torch.ones(mapper.P, dtype=torch.bool, device=dev)
            keep[split_idx] = False; keep[prune_idx] = False
            mapper.prune(keep)
            eng.set_num_gaussians(mapper.P, mapper.new_grad_block())
            e2.record(); e2.synchronize()
            if it >= 2:
                t_render.append(e0.elapsed_time(e1)); t_churn.append(1e3 * (time.perf_counter() - t_host))
            sizes.append(mapper.P); Rs.append(int(eng.R))
        row = dict(N_start=N, image=[W, H], N_after_each_iteration=sizes, R=Rs[-1], ms_iteration_fwd_bwd_update=float(np.median(t_render)),
                   ms_churn_clone_split_prune_host_wall=float(np.median(t_churn)), mpix_per_s=H * W / float(np.median(t_render)) / 1e3,
                   churn="per iteration: clone 1 %, split 1 % into 2 (parents removed), prune 2 %, seeded; every array changes size")
        rows.append(row)
        print("sweep", json.dumps(row), flush=True)
        del eng, mapper
        torch.cuda.empty_cache()
    out["sweep_1080p_with_churn"] = rows

out["clocks_after"] = clocks()
out["clocks_columns"] = "sm MHz, max sm MHz, power W, hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap"
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
path = os.path.join(ROOT, "gpurun_out", "r02_measurements.json")
prev = {}
if os.path.exists(path):
    try:
        prev = json.load(open(path))
    except Exception:
        prev = {}
prev.update(out)
json.dump(prev, open(path, "w"), indent=1)
print("wrote", path)
