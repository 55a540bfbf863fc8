"""BASELINE configs[4]: scale sweep at 1920x1080, N = 100k .. 8M Gaussians, forward + backward through the engine, for the
default tile-segment sort and for LVDGS_FLAG_GLOBAL_SORT (onesweep).  Checks that both give the same point list and
prints per-N: instances, longest tile list, ms fwd+bwd, Mpix/s.  Usage: python scripts/scale_sweep.py [N ...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lvd_gs-slam_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from lvdgs import synth, _native
from lvdgs.engine import RasterEngine, ViewCamera

Ns = [int(a) for a in sys.argv[1:]] or [100_000, 500_000, 2_000_000, 8_000_000]
dev = torch.device("cuda")
cam = synth.make_camera("hd")
H, W = cam.image_height, cam.image_width
gc, gd = synth.make_upstream_grads(cam)
gc, gd = torch.tensor(gc, device=dev), torch.tensor(gd, device=dev)
vc = ViewCamera(cam, dev)
for N in Ns:
    sc = synth.make_scene(N, cam, seed=0)
    t = lambda a: torch.tensor(a, device=dev)
    m, o, s, r, sh = t(sc["means3D"]), t(sc["opacities"]), t(sc["scales"]), t(sc["rotations"]), t(sc["shs"])
    res = {}
    for flags, name in ((0, "tile_sort"), (16, "global_sort")):
        eng = RasterEngine(N, W, H, device=dev, flags=flags, slots=1)
        for _ in range(3):
            eng.zero_grads(); eng.forward(vc, m, o, s, r, sh); eng.backward(vc, m, o, s, r, sh, gc, gd)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 5
        e0.record()
        for _ in range(iters):
            eng.zero_grads(); eng.forward(vc, m, o, s, r, sh); eng.backward(vc, m, o, s, r, sh, gc, gd)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        stream = torch.cuda.current_stream().cuda_stream
        _native.profile_begin(stream)
        eng.forward(vc, m, o, s, r, sh); eng.backward(vc, m, o, s, r, sh, gc, gd)
        prof = {}
        for k, v in _native.profile_end(stream):
            prof[k] = prof.get(k, 0.0) + v
        views = _native.debug_views(eng.slots[0].arena, N, eng.R, W, H, capacity=eng.slots[0].capacity)
        rg = views["ranges"].astype(np.int64)
        res[name] = dict(ms=ms, point_list=views["point_list"], prof=prof, longest=int((rg[:, 1] - rg[:, 0]).max()))
        del eng
    same = bool(np.array_equal(res["tile_sort"]["point_list"], res["global_sort"]["point_list"]))
    srt = lambda p: sum(v for k, v in p.items() if "sort" in k or "emit" in k)
    print(json.dumps(dict(N=N, image=[W, H], R=int(res["tile_sort"]["point_list"].size), longest_list=res["tile_sort"]["longest"],
                          point_lists_identical=same,
                          tile_sort=dict(ms_fwd_bwd=round(res["tile_sort"]["ms"], 3), mpix_s=round(H * W / res["tile_sort"]["ms"] / 1e3, 1),
                                         emit_plus_sort_ms=round(srt(res["tile_sort"]["prof"]), 3)),
                          global_sort=dict(ms_fwd_bwd=round(res["global_sort"]["ms"], 3), mpix_s=round(H * W / res["global_sort"]["ms"] / 1e3, 1),
                                           emit_plus_sort_ms=round(srt(res["global_sort"]["prof"]), 3)))), flush=True)
