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
