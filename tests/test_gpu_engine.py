"""RasterEngine (allocation-free C-ABI driver): the two-stream, two-slot pipelined run over several views must produce
the same accumulated gradient block as the plugin surface rendering the views one after the other."""
import numpy as np
import pytest
import torch

from lvdgs import synth, _native
from lvdgs.engine import RasterEngine, ViewCamera
from gpu_harness import run_cuda, rel_err

pytestmark = pytest.mark.gpu


def test_pipelined_views_match_sequential_plugin():
    dev = torch.device("cuda")
    cams = [synth.make_camera("mast3r_kitti", k) for k in range(5)]
    sc = synth.make_scene(30_000, cams[0], seed=8)
    H, W, P = cams[0].image_height, cams[0].image_width, 30_000
    rng = np.random.default_rng(1)
    gcs = [rng.normal(0, 1, (3, H, W)).astype(np.float32) for _ in cams]
    gds = [rng.normal(0, 1, (1, H, W)).astype(np.float32) for _ in cams]
    # reference: plugin surface, one view at a time, gradients summed in float64
    ref = {k: 0.0 for k in ("means3D", "opacities", "scales", "rotations", "shs")}
    outs, per_view = [], []
    for cam, gc, gd in zip(cams, gcs, gds):
        out, _, g = run_cuda(sc, cam, np.zeros(3, np.float32), grads=(gc, gd, None), debug=False)
        outs.append(out)
        per_view.append((g["means2D"], np.concatenate([g["rho"].reshape(-1), g["theta"].reshape(-1)])))
        for k in ref:
            ref[k] = ref[k] + g[k].astype(np.float64)
    t = lambda a: torch.tensor(a, device=dev)
    means3D, opac, scales, rots, shs = t(sc["means3D"]), t(sc["opacities"]), t(sc["scales"]), t(sc["rotations"]), t(sc["shs"])
    eng = RasterEngine(P, W, H, device=dev)
    vcs = [ViewCamera(c, dev) for c in cams]
    tg = [(t(gc), t(gd)) for gc, gd in zip(gcs, gds)]
    seen = {}

    def upstream(k, slot):
        seen[k] = (slot.color.clone(), slot.radii.clone(), slot.n_touched.clone())
        return tg[k][0], tg[k][1], None

    for rep in range(3):        # run 0 sizes the arenas (exact mode), later runs use the speculative launch
        eng.zero_grads()
        eng.run_views(vcs, means3D, opac, scales, rots, shs, upstream)
        torch.cuda.synchronize()
        got = {k: v.detach().cpu().numpy() for k, v in eng.grads.items()}
        assert rel_err(got["means3D"].reshape(P, 3), ref["means3D"]) < 1e-4
        assert rel_err(got["opacity"], ref["opacities"].reshape(-1)) < 1e-4
        assert rel_err(got["scales"].reshape(P, 3), ref["scales"]) < 1e-4
        assert rel_err(got["rotations"].reshape(P, 4), ref["rotations"]) < 1e-4
        assert rel_err(got["shs"].reshape(P, 1, 3), ref["shs"]) < 1e-4
        # per-view outputs of the last view of slot 0 (view 4): the backward walks the visible list only, the rows of
        # culled Gaussians must still read as zero
        assert rel_err(eng.slots[0].g_means2D.cpu().numpy(), per_view[4][0]) < 1e-4
        assert rel_err(eng.slots[0].g_tau.cpu().numpy(), per_view[4][1]) < 1e-4
        culled = outs[4]["radii"] == 0
        assert culled.any() and float(eng.slots[0].g_means2D[torch.tensor(culled, device=dev)].abs().max()) == 0.0
        for k, out in enumerate(outs):
            np.testing.assert_array_equal(seen[k][0].cpu().numpy(), out["color"])
            np.testing.assert_array_equal(seen[k][1].cpu().numpy(), out["radii"])
            np.testing.assert_array_equal(seen[k][2].cpu().numpy(), out["n_touched"])


def test_single_view_window_runs_on_the_callers_stream_and_matches_the_plugin():
    """A window of ONE view (a rank of an 8-GPU mapping step) takes run_views' single-stream path: same gradient block,
    same per-view outputs as the plugin surface."""
    dev = torch.device("cuda")
    cam = synth.make_camera("mast3r_kitti", 2)
    P = 20_000
    sc = synth.make_scene(P, cam, seed=12)
    H, W = cam.image_height, cam.image_width
    rng = np.random.default_rng(4)
    gc, gd = rng.normal(0, 1, (3, H, W)).astype(np.float32), rng.normal(0, 1, (1, H, W)).astype(np.float32)
    out, _, g = run_cuda(sc, cam, np.zeros(3, np.float32), grads=(gc, gd, None), debug=False)
    t = lambda a: torch.tensor(a, device=dev)
    args = [t(sc[k]) for k in ("means3D", "opacities", "scales", "rotations", "shs")]
    eng = RasterEngine(P, W, H, device=dev)
    assert eng.view_streams(1)[0] == torch.cuda.current_stream(dev) and eng.view_streams(2)[0] == eng.s_fwd
    tg = (t(gc), t(gd))
    for rep in range(3):
        eng.zero_grads()
        eng.run_views([ViewCamera(cam, dev)], *args, lambda k, slot: (tg[0], tg[1], None))
        torch.cuda.synchronize()
        np.testing.assert_array_equal(eng.color.cpu().numpy(), out["color"])
        assert rel_err(eng.grads["means3D"].cpu().numpy().reshape(P, 3), g["means3D"]) < 1e-5
        assert rel_err(eng.grads["rotations"].cpu().numpy().reshape(P, 4), g["rotations"]) < 1e-5
        assert rel_err(eng.g_tau.cpu().numpy(), np.concatenate([g["rho"].reshape(-1), g["theta"].reshape(-1)])) < 1e-5


def test_alternating_image_sizes_never_rerun_the_speculative_tail():
    """tracking renders 1241x376 and, per frame, one 512x144 depth image (utils/init_pose.py:145) on the
    same thread.  The long-list history is kept per (device, image size), so after warm-up neither shape disturbs the
    other's speculation: zero repeated tails, and a second engine of another size on the same thread changes nothing."""
    dev = torch.device("cuda")
    L = _native.lib()
    big, small = synth.make_camera("kitti"), synth.make_camera("mast3r_kitti")
    N = 120_000
    sc = synth.make_scene(N, big, seed=4)
    sc["scales"] *= 2.5                                        # long tile lists at full resolution, short ones at 512x144
    t = lambda a: torch.tensor(a, device=dev)
    args = [t(sc[k]) for k in ("means3D", "opacities", "scales", "rotations", "shs")]
    e_big = RasterEngine(N, big.image_width, big.image_height, device=dev, slots=1)
    e_small = RasterEngine(N, small.image_width, small.image_height, device=dev, slots=1)
    v_big, v_small = ViewCamera(big, dev), ViewCamera(small, dev)
    for _ in range(3):                                         # warm-up: capacity hints and list-length histories settle
        e_big.forward(v_big, *args); e_small.forward(v_small, *args)
    ref_big, ref_small = e_big.color.clone(), e_small.depth.clone()
    before = L.lvdgs_tail_rerun_count()
    for _ in range(6):
        e_big.forward(v_big, *args); e_small.forward(v_small, *args)
    torch.cuda.synchronize()
    assert L.lvdgs_tail_rerun_count() == before
    assert torch.equal(e_big.color, ref_big) and torch.equal(e_small.depth, ref_small)
