"""Host logic of the keyframe-sharded mapping iteration (lvdgs.mapping) on CPU: sharding, the SUM all-reduce of the
contiguous gradient block over gloo at world_size 2, and bit-identical replicas after the Adam step."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lvdgs.mapping import ShardedMapper, shard_keyframes, group_widths


def test_shard_keyframes_partitions_the_window():
    for n in (1, 7, 8, 10):
        for world in (1, 2, 3, 4, 8):
            shards = [shard_keyframes(n, world, r) for r in range(world)]
            flat = sorted(k for s in shards for k in s)
            assert flat == list(range(n))
            assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1


def torch_adam(m, grad):
    """Reference Adam in plain torch (float32), the stand-in optimiser for the CPU host-logic tests and the checker of
    the fused CUDA kernel in test_gpu_adam_matches_torch."""
    b1, b2 = m.betas
    m.exp_avg.mul_(b1).add_(grad, alpha=1 - b1)
    m.exp_avg_sq.mul_(b2).addcmul_(grad, grad, value=1 - b2)
    bc1, bc2 = 1 - b1 ** m.t, 1 - b2 ** m.t
    denom = (m.exp_avg_sq / bc2).sqrt_().add_(m.eps)
    m.param_flat.addcdiv_(m.exp_avg * m.lr_flat / bc1, denom, value=-1.0)


def _fake_view_grad(mapper, k, P):
    """Deterministic stand-in for 'render view k and back-propagate': depends on the view and on the parameters."""
    g = torch.Generator().manual_seed(100 + k)
    base = torch.randn(14 * P + 64, generator=g)[:mapper.param_flat.numel()]      # independent of the tail padding
    return base * 1e-2 + 0.1 * torch.sin(mapper.param_flat * (k + 1))


def _init_params(mapper, P):
    g = torch.Generator().manual_seed(7)
    mapper.param_flat.copy_(torch.randn(14 * P + 64, generator=g)[:mapper.param_flat.numel()])


def _run_rank(rank, world, port, P, n_views, iters, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mapper = ShardedMapper(P, sh_coeffs=1, device="cpu", optimizer_fn=torch_adam)
        _init_params(mapper, P)
        grad = torch.zeros_like(mapper.param_flat)
        owned = []
        for it in range(iters):
            extra = [n_views + int(x) for x in torch.randperm(4, generator=torch.Generator().manual_seed(it))[:2]]
            mine = mapper.step(n_views, lambda k: grad.add_(_fake_view_grad(mapper, k, P)), grad, extra_views=extra)
            owned.append(mine)
            vis = torch.zeros(P, dtype=torch.bool); vis[rank::world] = True
            mapper.add_densification_stats(torch.full((P, 2), float(rank + 1)), vis, torch.full((P,), 3 + rank))
        mapper.reduce_stats()
        torch.save(dict(params=mapper.param_flat.clone(), owned=owned, accum=mapper.grad_norm_accum.clone(),
                        denom=mapper.denom.clone(), radii=mapper.max_radii2D.clone()), os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.timeout(180)
def test_two_ranks_match_single_process(tmp_path):
    P, n_views, iters, world = 257, 8, 3, 2
    mp.spawn(_run_rank, args=(world, _free_port(), P, n_views, iters, str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(tmp_path / "r0.pt")
    r1 = torch.load(tmp_path / "r1.pt")
    # replicas are bit-identical: same all-reduced gradient, same Adam arithmetic
    assert torch.equal(r0["params"], r1["params"])
    # every view of every iteration was rendered by exactly one rank
    for it in range(iters):
        assert sorted(r0["owned"][it] + r1["owned"][it]) == sorted(set(r0["owned"][it] + r1["owned"][it]))
        assert len(r0["owned"][it]) + len(r1["owned"][it]) == n_views + 2
    # single-process reference: all views on one rank, no collective
    ref = ShardedMapper(P, sh_coeffs=1, device="cpu", optimizer_fn=torch_adam)
    _init_params(ref, P)
    grad = torch.zeros_like(ref.param_flat)
    for it in range(iters):
        extra = [n_views + int(x) for x in torch.randperm(4, generator=torch.Generator().manual_seed(it))[:2]]
        ref.step(n_views, lambda k: grad.add_(_fake_view_grad(ref, k, P)), grad, extra_views=extra)
    # float32 sums in a different association (per-rank partial sums, then all-reduce): equal to rounding
    n = ref.param_flat.numel()                  # the 2-rank block carries up to 4 more floats of tail padding (reduce-scatterable)
    assert torch.allclose(r0["params"][:n], ref.param_flat, rtol=1e-5, atol=1e-6)
    # densification side-band: SUM of norms / counts, MAX of radii
    assert torch.equal(r0["accum"], r1["accum"]) and torch.equal(r0["radii"], r1["radii"])
    expect_denom = torch.full((P,), float(iters))
    assert torch.equal(r0["denom"], expect_denom)
    assert float(r0["radii"].max()) == 4.0


def test_block_layout_matches_engine_gradient_block():
    """The mapper's parameter block and RasterEngine.grad_flat use the same layout (lvdgs.engine.block_layout: group order,
    widths, 16-byte aligned group starts), so the reduced gradient block is consumed by adam_step without repacking."""
    from lvdgs.engine import block_layout
    P, M = 10, 4
    m = ShardedMapper(P, sh_coeffs=M, device="cpu")
    layout, total = block_layout(P, M)
    w = group_widths(M)
    off = 0
    for name in ("means3D", "shs", "opacity", "scales", "rotations"):
        assert m.slices[name] == slice(off, off + w[name] * P) == slice(layout[name][0], layout[name][0] + layout[name][1])
        off += (w[name] * P + 3) // 4 * 4
    assert m.param_flat.numel() == total == off
    assert m.view("shs").shape == (P, M, 3) and m.view("rotations").shape == (P, 4)
    assert m.view("means3D").data_ptr() == m.param_flat.data_ptr()
    # a reduce-scatterable block: the padded total is a multiple of 4 x world
    for world in (2, 3, 8):
        assert block_layout(257, 1, multiple=4 * world)[1] % (4 * world) == 0


def test_cpu_parameters_without_stand_in_optimizer_raise():
    m = ShardedMapper(4, device="cpu")
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.adam_step(torch.zeros_like(m.param_flat))


@pytest.mark.gpu
def test_gpu_adam_matches_torch():
    P = 10_007
    g = torch.Generator().manual_seed(1)
    a = ShardedMapper(P, sh_coeffs=1, device="cuda", eps=1e-8)
    b = ShardedMapper(P, sh_coeffs=1, device="cpu", eps=1e-8, optimizer_fn=torch_adam)
    init = torch.randn(a.param_flat.numel(), generator=g)
    a.param_flat.copy_(init); b.param_flat.copy_(init)
    for it in range(4):
        grad = torch.randn(init.numel(), generator=g) * (10.0 ** (it - 2))
        a.adam_step(grad.cuda()); b.adam_step(grad.clone())
    torch.cuda.synchronize()
    for sl in a.slices.values():              # group by group: the alignment padding between groups is nobody's parameter
        assert torch.allclose(a.param_flat[sl].cpu(), b.param_flat[sl], rtol=1e-5, atol=1e-7)
        assert torch.allclose(a.exp_avg_sq[sl].cpu(), b.exp_avg_sq[sl], rtol=1e-5, atol=0)
