"""GPU parity of the SURVEY 8f "next" rows built so far (N3 fused losses, N4 covisibility counts, N1 prune compaction)
against the reference code they replace (cited per test; tests/ref_conventions.py loads or restates it).  Losses: float32 sums in a different
order -> 1e-5 relative on the loss, 1e-6 absolute on the (O(1/HW)) gradients; integer / row-move results exact."""
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

H, W = 37, 53


def _viewpoint(rng, dev):
    vp = types.SimpleNamespace()
    gt = rng.uniform(0, 1, (3, H, W)).astype(np.float32)
    gt[:, :5, :7] = 0.001                                       # below rgb_boundary_threshold
    vp.original_image = torch.tensor(gt, device=dev)
    md = rng.uniform(0.5, 40, (H, W)).astype(np.float32)
    md[10:14, 20:30] = 0.0                                      # invalid mono depth
    vp.mono_depth = md
    vp.grad_mask = torch.tensor((rng.uniform(0, 1, (1, H, W)) > 0.3), device=dev)
    vp.exposure_a = torch.tensor([0.07], device=dev, requires_grad=True)
    vp.exposure_b = torch.tensor([-0.02], device=dev, requires_grad=True)
    return vp


# ---- the reference's losses (utils/slam_utils.py:42-121): its own functions when /root/reference is mounted, else the
# restatements pinned against them by tests/test_reference_pin.py ----
import ref_conventions as rc

_SU = rc.load()[1]
ref_tracking, ref_mapping = _SU.get_loss_tracking, _SU.get_loss_mapping


def _inputs(rng, dev):
    mk = lambda a: torch.tensor(a.astype(np.float32), device=dev, requires_grad=True)
    image = mk(rng.uniform(0, 1, (3, H, W)))
    depth = mk(rng.uniform(0.5, 40, (1, H, W)))
    opacity = mk(rng.uniform(0.5, 1.0, (1, H, W)))
    return image, depth, opacity


def _grads(loss, leaves):
    gs = torch.autograd.grad(loss * 1.7, leaves, allow_unused=True)     # a non-trivial grad_output
    return [None if g is None else g.detach().cpu().numpy() for g in gs]


@pytest.mark.parametrize("mode", ["track_mono", "track_rgbd", "map_rgbd", "map_rgb", "map_init"])
def test_fused_losses_match_the_reference_expressions(mode):
    from lvdgs import slam_ops
    dev = torch.device("cuda")
    rng = np.random.default_rng(7)
    vp = _viewpoint(rng, dev)
    image, depth, opacity = _inputs(rng, dev)
    config = {"Training": {"monocular": mode != "track_rgbd", "rgb_boundary_threshold": 0.01, "alpha": 0.9}, "Dataset": {"depth_loss": False}}
    leaves = [image, depth, opacity, vp.exposure_a, vp.exposure_b]
    if mode.startswith("track"):
        ours = slam_ops.get_loss_tracking(config, image, depth, opacity, vp)
        ref = ref_tracking(config, image, depth, opacity, vp)
    else:
        kw = dict(depth=depth, initialization=mode == "map_init", monodepth=mode != "map_rgb")
        ours = slam_ops.get_loss_mapping(config, image, vp, **kw)
        ref = ref_mapping(config, image, vp, **kw)
    assert abs(float(ours) - float(ref)) <= 1e-5 * abs(float(ref))
    for go, gr in zip(_grads(ours, leaves), _grads(ref, leaves)):
        if gr is None:
            assert go is None or np.abs(go).max() == 0
        else:
            assert go is not None
            np.testing.assert_allclose(go, gr, rtol=2e-5, atol=1e-7 if gr.size > 1 else 1e-6)


def test_fused_loss_is_deterministic_and_feeds_the_rasterizer_layout():
    from lvdgs import slam_ops
    dev = torch.device("cuda")
    rng = np.random.default_rng(9)
    vp = _viewpoint(rng, dev)
    image, depth, opacity = _inputs(rng, dev)
    config = {"Training": {"monocular": True, "rgb_boundary_threshold": 0.01}}
    vals = [float(slam_ops.get_loss_mapping(config, image, vp, depth=depth)) for _ in range(5)]
    assert len(set(vals)) == 1
    with pytest.raises(RuntimeError):
        slam_ops.get_loss_mapping(config, image.cpu(), vp, depth=depth.cpu())


def test_covisibility_and_n_obs():
    """utils/slam_frontend.py:1598-1603,1631-1639 and utils/slam_backend.py:322-325."""
    from lvdgs import slam_ops
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(0)
    for n in (0, 1, 31, 1000, 300_001):
        a = (torch.rand(n, generator=g) > 0.4).to(dev)
        b = (torch.rand(n, generator=g) > 0.7).to(dev)
        want = [int(a.count_nonzero()), int(b.count_nonzero()), int((a & b).count_nonzero()), int((a | b).count_nonzero())]
        for conv in (lambda t: t, lambda t: t.to(torch.int32) * 5, lambda t: t.long()):      # bool, n_touched-like int32, .long()
            assert slam_ops.covisibility(conv(a), conv(b)).tolist() == want
    masks = [(torch.rand(5000, generator=g) > 0.5).long().to(dev) for _ in range(8)]
    n_obs = slam_ops.accumulate_n_obs(masks)
    assert torch.equal(n_obs.long(), torch.stack(masks).sum(0))


@pytest.mark.parametrize("n", [0, 1, 1023, 1024, 1025, 200_000])
def test_compact_rows_equals_boolean_indexing(n):
    """GaussianModel.prune_points: every tensor indexed by the same mask (utils/slam_backend.py:128-145)."""
    from lvdgs import slam_ops
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(n)
    tensors = [torch.randn(n, w, generator=g).to(dev) for w in (3, 3, 1, 3, 4, 48)] + [torch.randn(n, 1, 3, generator=g).to(dev)]
    for frac in (0.0, 0.3, 1.0):
        keep = (torch.rand(n, generator=g) < frac).to(dev)
        out = slam_ops.compact_rows(keep, tensors)
        for t, o in zip(tensors, out):
            assert torch.equal(o, t[keep])


def test_mapper_prune_keeps_replica_state_consistent():
    from lvdgs.mapping import ShardedMapper, GROUPS
    dev = torch.device("cuda")
    P = 5000
    m = ShardedMapper(P, sh_coeffs=1, device=dev)
    g = torch.Generator(device="cpu").manual_seed(1)
    m.param_flat.copy_(torch.randn(m.param_flat.numel(), generator=g))
    m.exp_avg.copy_(torch.randn(m.param_flat.numel(), generator=g))
    m.exp_avg_sq.copy_(torch.rand(m.param_flat.numel(), generator=g))
    m.denom.copy_(torch.rand(P, generator=g))
    before = {n: m.raw_view(n).clone() for n in GROUPS}
    avg_before = m.exp_avg[m.slices["rotations"]].view(P, 4).clone()
    denom_before = m.denom.clone()
    keep = (torch.rand(P, generator=g) > 0.25).to(dev)
    P2 = m.prune(keep)
    assert P2 == int(keep.sum()) and m.P == P2 and 14 * P2 <= m.param_flat.numel() < 14 * P2 + 32
    for n in GROUPS:
        assert torch.equal(m.raw_view(n), before[n][keep])
    # the activated copies the rasterizer reads follow the pruned raw block
    assert torch.allclose(m.view("opacity"), torch.sigmoid(m.raw_view("opacity")), atol=1e-6)
    assert torch.allclose(m.view("rotations"), torch.nn.functional.normalize(m.raw_view("rotations")), atol=1e-6)
    assert torch.equal(m.exp_avg[m.slices["rotations"]].view(P2, 4), avg_before[keep])
    assert torch.equal(m.denom, denom_before[keep])
    m.adam_step(torch.ones_like(m.param_flat))          # the fused optimiser runs on the pruned block


def test_gather_rows_and_mapper_densify():
    """densify_and_clone / densify_and_split's row copies (utils/slam_backend.py:359-376): exact against t[index]."""
    from lvdgs import slam_ops
    from lvdgs.mapping import ShardedMapper, GROUPS
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(3)
    n = 10_000
    tensors = [torch.randn(n, w, generator=g).to(dev) for w in (3, 3, 1, 3, 4)] + [torch.randn(n, 16, 3, generator=g).to(dev)]
    for m in (0, 1, 777, 25_000):
        idx = torch.randint(0, n, (m,), generator=g).to(dev)
        for t, o in zip(tensors, slam_ops.gather_rows(idx, tensors)):
            assert torch.equal(o, t[idx])
    P = 4000
    mp = ShardedMapper(P, sh_coeffs=1, device=dev)
    mp.param_flat.copy_(torch.randn(mp.param_flat.numel(), generator=g))
    mp.exp_avg.copy_(torch.randn(mp.param_flat.numel(), generator=g))
    before = {k: mp.raw_view(k).clone() for k in GROUPS}
    avg_before = mp.exp_avg[mp.slices["scales"]].view(P, 3).clone()
    idx = torch.randint(0, P, (900,), generator=g).to(dev)
    new_means = torch.randn(900, 3, generator=g).to(dev)
    P2 = mp.densify_clone(idx, overrides={"means3D": new_means})
    assert P2 == P + 900 and 14 * P2 <= mp.param_flat.numel() < 14 * P2 + 32 and mp.denom.numel() == P2
    for k in GROUPS:
        v = mp.raw_view(k)
        assert torch.equal(v[:P], before[k])
        assert torch.equal(v[P:], new_means if k == "means3D" else before[k][idx])
    assert torch.equal(mp.exp_avg[mp.slices["scales"]].view(P2, 3)[:P], avg_before)
    assert float(mp.exp_avg[mp.slices["scales"]].view(P2, 3)[P:].abs().max()) == 0.0
    mp.adam_step(torch.ones_like(mp.param_flat))


def test_mapper_optimises_raw_parameters_like_gaussian_model():
    """the reference's GaussianModel runs Adam on logit-opacity / log-scale / un-normalised quaternions and
    renders their activations.  60 steps at the REAL learning rates (configs/mono/KITTI/base_config.yaml:59-66) with a
    gradient that depends on the activated values: the mapper (lvdgs_gaussian_activate -> gradient with respect to the
    activations -> lvdgs_gaussian_activation_backward -> lvdgs_adam_step) against torch autograd + torch.optim.Adam on
    the same raw leaves."""
    from lvdgs.mapping import ShardedMapper, GROUPS
    dev = torch.device("cuda")
    P = 3000
    g = torch.Generator(device="cpu").manual_seed(11)
    init = {"means3D": torch.randn(P, 3, generator=g), "shs": torch.randn(P, 1, 3, generator=g),
            "opacity": torch.rand(P, 1, generator=g) * 0.9 + 0.05, "scales": torch.rand(P, 3, generator=g) * 0.5 + 0.01,
            "rotations": torch.randn(P, 4, generator=g)}
    target = {k: (torch.rand_like(v) * 0.8 + 0.1).to(dev) for k, v in init.items()}
    lrs = {"means3D": 1.6e-4, "shs": 2.5e-3, "opacity": 5e-2, "scales": 1e-3, "rotations": 1e-3}
    m = ShardedMapper(P, sh_coeffs=1, device=dev, lrs=lrs, eps=1e-15)
    m.load(**{k: v.numpy() for k, v in init.items()})
    # torch side: raw leaves, activations as GaussianModel's get_* properties
    raw = {"means3D": init["means3D"].clone(), "shs": init["shs"].clone(), "opacity": torch.log(init["opacity"] / (1 - init["opacity"])),
           "scales": torch.log(init["scales"]), "rotations": init["rotations"].clone()}
    raw = {k: v.to(dev).requires_grad_() for k, v in raw.items()}
    opt = torch.optim.Adam([{"params": [raw[k]], "lr": lrs[k]} for k in GROUPS], eps=1e-15)
    act = lambda r: {"means3D": r["means3D"], "shs": r["shs"], "opacity": torch.sigmoid(r["opacity"]), "scales": torch.exp(r["scales"]),
                     "rotations": torch.nn.functional.normalize(r["rotations"])}
    grad = m.new_grad_block()
    for it in range(60):
        # dL/d(activated) for L = sum over groups of 0.5 |a - target|^2 * 100 / P: what a rasterizer backward would hand over
        for k in GROUPS:
            grad[m.slices[k]] = ((m.view(k).reshape(target[k].shape) - target[k]) * (100.0 / P)).reshape(-1)
        m.exchange_and_update(grad)
        assert float(grad.abs().max()) == 0.0                        # left zeroed for the next iteration
        opt.zero_grad()
        a = act(raw)
        loss = sum(0.5 * ((a[k] - target[k]) ** 2).sum() for k in GROUPS) * (100.0 / P)
        loss.backward()
        opt.step()
    for k in GROUPS:
        np.testing.assert_allclose(m.raw_view(k).cpu().numpy(), raw[k].detach().reshape(P, -1).cpu().numpy(), rtol=2e-4, atol=2e-5)
    a = act(raw)
    assert float(m.view("opacity").min()) > 0 and float(m.view("opacity").max()) < 1
    np.testing.assert_allclose(m.view("rotations").norm(dim=1).cpu().numpy(), 1.0, atol=1e-5)
    np.testing.assert_allclose(m.view("scales").cpu().numpy(), a["scales"].detach().cpu().numpy(), rtol=2e-4)


def test_mapper_gaussian_model_maintenance():
    """SURVEY 8f N1 remainder on the raw-parameter block: extend_from_pcd_seq's insertion with distCUDA2 scale initialisation
    (caller utils/slam_backend.py:75-78), densify_and_prune's clone / split / prune decisions (:359-370) and
    reset_opacity_nonvisible (:372-376) -- against the same rules written with torch indexing."""
    from lvdgs.mapping import ShardedMapper
    from simple_knn._C import distCUDA2
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(5)
    m = ShardedMapper(0, sh_coeffs=1, device=dev)
    pts = (torch.rand(3000, 3, generator=g) * 4).to(dev)
    cols = torch.rand(3000, 3, generator=g).to(dev)
    assert m.extend_from_points(pts, cols, point_size=1.0, kf_id=7) == 3000
    d2 = torch.clamp_min(distCUDA2(pts), 1e-7)
    assert torch.allclose(m.view("scales"), torch.sqrt(d2)[:, None].repeat(1, 3), rtol=1e-5)
    assert torch.allclose(m.view("opacity"), torch.full((3000, 1), 0.5, device=dev), atol=1e-6)
    assert torch.equal(m.view("rotations"), torch.tensor([1.0, 0, 0, 0], device=dev).repeat(3000, 1))
    assert torch.allclose(m.view("shs")[:, 0] * 0.28209479177387814 + 0.5, cols, atol=1e-6)
    assert int((m.unique_kfIDs == 7).sum()) == 3000 and float(m.exp_avg.abs().max()) == 0.0
    # a second keyframe keeps the first one's rows and Adam moments in place
    m.exp_avg.fill_(0.25)
    assert m.extend_from_points(pts[:500] + 10.0, cols[:500], kf_id=9) == 3500
    assert torch.equal(m.view("means3D")[:3000], pts) and float(m.exp_avg[m.slices["scales"]].view(3500, 3)[:3000].min()) == 0.25
    assert float(m.exp_avg[m.slices["scales"]].view(3500, 3)[3000:].abs().max()) == 0.0

    # ---- densify_and_prune ----
    P = m.P
    m.grad_norm_accum.copy_(torch.rand(P, generator=g).to(dev) * 4e-4); m.denom.fill_(1.0)
    m.denom[:10] = 0.0                                                   # never-seen Gaussians: NaN gradient -> 0
    m.max_radii2D.copy_(torch.rand(P, generator=g).to(dev) * 30)
    m.params["opacity"].copy_(torch.logit(torch.rand(P, generator=g).clamp(0.01, 0.99)).to(dev)); m.activate()
    extent, thr, min_op, max_screen = 20.0, 2e-4, 0.3, 25.0
    scales0, opac0, means0 = m.view("scales").clone(), m.view("opacity").clone().reshape(-1), m.view("means3D").clone()
    grads = (m.grad_norm_accum / m.denom).nan_to_num(0.0)
    hot, small = grads >= thr, scales0.max(1).values <= 0.01 * extent
    n_clone, n_split = int((hot & small).sum()), int((hot & ~small).sum())
    assert n_clone > 0 and n_split > 0
    radii0 = m.max_radii2D.clone()
    P2 = m.densify_and_prune(thr, min_op, extent, max_screen, generator=torch.Generator(device="cpu").manual_seed(1))
    # expected survivors: originals that were not split and pass the prune rules, clones and split children likewise
    keep_orig = ~(hot & ~small) & ~((opac0 < min_op) | (radii0 > max_screen) | (scales0.max(1).values > 0.1 * extent))
    clone_keep = ~((opac0 < min_op) | (scales0.max(1).values > 0.1 * extent))[hot & small]           # clones start with max_radii2D = 0
    child_scales = (scales0 / 1.6)[hot & ~small].repeat(2, 1)
    child_keep = ~((opac0[hot & ~small].repeat(2) < min_op) | (child_scales.max(1).values > 0.1 * extent))
    assert P2 == int(keep_orig.sum()) + int(clone_keep.sum()) + int(child_keep.sum())
    n0 = int(keep_orig.sum())
    assert torch.equal(m.view("means3D")[:n0], means0[keep_orig])
    assert torch.allclose(m.view("scales")[n0 + int(clone_keep.sum()):], child_scales[child_keep], rtol=1e-5)
    assert float(m.grad_norm_accum.abs().max()) == 0.0 and m.unique_kfIDs.numel() == P2

    # ---- reset_opacity_nonvisible ----
    vis = [torch.rand(P2, generator=g).to(dev) > 0.7 for _ in range(3)]
    before = m.view("opacity").clone().reshape(-1)
    m.exp_avg[m.slices["opacity"]].fill_(1.0)
    m.reset_opacity_nonvisible(vis)
    seen = vis[0] | vis[1] | vis[2]
    after = m.view("opacity").reshape(-1)
    assert torch.equal(after[seen], before[seen]) and torch.allclose(after[~seen], torch.full_like(after[~seen], 0.4), atol=1e-6)
    assert float(m.exp_avg[m.slices["opacity"]].abs().max()) == 0.0


def _reference_masked_loss(image, gt_image, static_mask, background, depth, mono_depth, lambda_dssim, depth_lambda):
    """utils/slam_backend.py:199-261, statement by statement (the shape-normalising branches reduce to these lines for
    [1,H,W] depth / [H,W] mono depth / [H,W] mask), with gaussian_splatting.utils.loss_utils' l1_loss / ssim."""
    from gaussian_splatting.utils.loss_utils import l1_loss, ssim
    masked_image = image.clone()
    masked_gt = gt_image.clone()
    for c in range(3):
        masked_image[c][~static_mask] = background[c]
        masked_gt[c][~static_mask] = background[c]
    Ll1 = l1_loss(masked_image, masked_gt)
    ssim_loss = 1.0 - ssim(masked_image, masked_gt)
    loss = (1.0 - lambda_dssim) * Ll1 + lambda_dssim * ssim_loss
    if depth is not None and mono_depth is not None:
        d = depth.squeeze(0)
        depth_mask = static_mask & (mono_depth > 0) & (d > 0)
        if depth_mask.any():
            loss = loss + depth_lambda * torch.abs(d[depth_mask] - mono_depth[depth_mask]).mean()
    return loss


@pytest.mark.parametrize("shape,with_depth,bg", [((37, 53), True, (0.0, 0.0, 0.0)), ((144, 512), True, (0.2, 0.5, 0.1)),
                                                 ((64, 48), False, (0.0, 0.0, 0.0)), ((376, 1241), True, (0.0, 0.0, 0.0))])
def test_masked_ssim_mapping_loss_matches_the_reference_expression(shape, with_depth, bg):
    """SURVEY 8f N3, the branch LVD-GS's dynamic-object masking takes (utils/slam_backend.py:199-261): value to 1e-5
    relative and every gradient element to 1e-3 of the tensor's scale against torch autograd of the reference expression
    evaluated in float64 on the CPU (cudnn's TF32 convolutions would be the less exact side)."""
    from lvdgs import slam_ops
    Hh, Ww = shape
    rng = np.random.default_rng(Hh * 7 + Ww)
    dev = torch.device("cuda")
    base = rng.uniform(0, 1, (3, Hh, Ww))
    image_np = np.clip(base + rng.normal(0, 0.08, base.shape), 0, 1).astype(np.float32)      # a render that resembles its target
    gt_np = base.astype(np.float32)
    mask_np = rng.uniform(0, 1, (Hh, Ww)) > 0.2
    mask_np[Hh // 3: Hh // 2, Ww // 4: Ww // 2] = False                                      # a solid dynamic object
    depth_np = rng.uniform(-0.5, 30, (1, Hh, Ww)).astype(np.float32)
    mono_np = rng.uniform(-1.0, 30, (Hh, Ww)).astype(np.float32)
    image = torch.tensor(image_np, device=dev, requires_grad=True)
    depth = torch.tensor(depth_np, device=dev, requires_grad=True) if with_depth else None
    loss, terms = slam_ops.masked_mapping_loss(image, torch.tensor(gt_np, device=dev), torch.tensor(mask_np, device=dev),
                                               torch.tensor(bg, device=dev), depth=depth,
                                               mono_depth=torch.tensor(mono_np, device=dev) if with_depth else None,
                                               lambda_dssim=0.2, depth_lambda=0.1, return_terms=True)
    (loss * 1.3).backward()
    img64 = torch.tensor(image_np, dtype=torch.float64, requires_grad=True)
    dep64 = torch.tensor(depth_np, dtype=torch.float64, requires_grad=True) if with_depth else None
    ref = _reference_masked_loss(img64, torch.tensor(gt_np, dtype=torch.float64), torch.tensor(mask_np),
                                 torch.tensor(bg, dtype=torch.float64), dep64,
                                 torch.tensor(mono_np, dtype=torch.float64) if with_depth else None, 0.2, 0.1)
    (ref * 1.3).backward()
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref)), (float(loss), float(ref))
    gi, gr = image.grad.cpu().numpy().astype(np.float64), img64.grad.numpy()
    assert np.abs(gi - gr).max() <= 1e-3 * np.abs(gr).max()
    assert np.all(gi[:, ~mask_np] == 0)                                                      # painted pixels carry no gradient
    if with_depth:
        np.testing.assert_allclose(depth.grad.cpu().numpy(), dep64.grad.numpy(), rtol=1e-5, atol=1e-12)
        assert float(terms[4]) == float((mask_np & (mono_np > 0) & (depth_np[0] > 0)).sum())
    # deterministic
    again = slam_ops.masked_mapping_loss(image.detach(), torch.tensor(gt_np, device=dev), torch.tensor(mask_np, device=dev),
                                         torch.tensor(bg, device=dev), depth=None if depth is None else depth.detach(),
                                         mono_depth=torch.tensor(mono_np, device=dev) if with_depth else None)
    assert float(again) == float(loss)


def test_single_gpu_fused_update_equals_the_three_kernel_sequence():
    """One GPU: ShardedMapper.exchange_and_update runs lvdgs_exchange_adam with a world of one (chain rule + Adam +
    activations in one launch); it must reproduce lvdgs_gaussian_activation_backward -> lvdgs_adam_step ->
    lvdgs_gaussian_activate (the expressions are the same; only fused-multiply-add contraction may differ)."""
    from lvdgs.mapping import ShardedMapper, GROUPS
    dev = torch.device("cuda")
    P = 5003                                        # not a multiple of 4: exercises the padded group boundaries
    g = torch.Generator(device="cpu").manual_seed(5)
    init = {"means3D": torch.randn(P, 3, generator=g), "shs": torch.randn(P, 1, 3, generator=g),
            "opacity": torch.rand(P, 1, generator=g) * 0.9 + 0.05, "scales": torch.rand(P, 3, generator=g) * 0.5 + 0.01,
            "rotations": torch.randn(P, 4, generator=g)}
    a, b = (ShardedMapper(P, sh_coeffs=1, device=dev) for _ in range(2))
    for m in (a, b):
        m.load(**{k: v.numpy() for k, v in init.items()})
    ga, gb = a.new_grad_block(), b.new_grad_block()
    for it in range(5):
        grad = torch.randn(a.total, generator=g).to(dev) * 1e-3
        for name, (off, ln) in a.layout.items():    # the padding between the groups carries no gradient
            end = min((o for o, _ in a.layout.values() if o > off), default=a.total)
            grad[off + ln:end] = 0
        ga.copy_(grad); gb.copy_(grad)
        a.exchange_and_update(ga)                   # fused
        b.activation_backward(gb); b.adam_step(gb); gb.zero_()      # the sequence (adam_step activates)
        assert float(ga.abs().max()) == 0.0
    assert a.t == b.t == 5
    for x, y, what in ((a.param_flat, b.param_flat, "params"), (a.exp_avg, b.exp_avg, "exp_avg"), (a.exp_avg_sq, b.exp_avg_sq, "exp_avg_sq")):
        np.testing.assert_allclose(x.cpu().numpy(), y.cpu().numpy(), rtol=2e-6, atol=1e-9, err_msg=what)
    for k in GROUPS:       # what the rasterizer reads (the float4 padding behind a group is never read by anyone)
        np.testing.assert_allclose(a.view(k).cpu().numpy(), b.view(k).cpu().numpy(), rtol=2e-6, atol=1e-9, err_msg=k)
