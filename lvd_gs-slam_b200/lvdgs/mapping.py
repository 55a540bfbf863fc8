"""Keyframe-sharded mapping iteration over several GPUs -- NEW functionality: the reference maps
on one GPU (README.md:93-104, CUDA_VISIBLE_DEVICES=0).

One mapping iteration of the reference (utils/slam_backend.py:153-389) renders every keyframe of the window
(<= 8, plus 2 random older keyframes) against the same Gaussian map, SUMS the per-view losses
(`loss_mapping +=`, :266/:300), back-propagates once and takes one Adam step.  The views are independent given the
map, so here: one process per GPU, the map replicated, keyframe k owned by rank k mod world; every rank accumulates
the parameter gradients of its views into one contiguous float32 block (lvdgs.engine.RasterEngine.grad_flat: the
backward kernel adds into it, there is no pack kernel) and the only data-path collective is ONE SUM all-reduce of
that block per iteration (56 B per Gaussian at SH degree 0), followed by the identical Adam update on every rank --
replicas stay bit-identical without a broadcast.  Densification statistics ride along: the accumulated 2-D
gradient norms and their denominators are SUM-reduced, `max_radii2D` is MAX-reduced (utils/slam_backend.py:350-357).
Per-keyframe visibility masks (`n_touched > 0`, :311-315) stay on the owning rank.

The collective / sharding logic is device-agnostic on purpose: the GPU path hands it RasterEngine's gradient block
over NCCL, the CPU tests (gloo, world_size 2) hand it a stand-in gradient function and a stand-in optimiser.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

# parameter groups in block order, with their widths per Gaussian (M = SH coefficients per Gaussian)
GROUPS = ("means3D", "shs", "opacity", "scales", "rotations")


def group_widths(sh_coeffs: int = 1) -> Dict[str, int]:
    return {"means3D": 3, "shs": 3 * sh_coeffs, "opacity": 1, "scales": 3, "rotations": 4}


def shard_keyframes(n_views: int, world: int, rank: int) -> List[int]:
    """Keyframe k is rendered by rank k mod world (round-robin keeps ranks balanced when n_views % world != 0)."""
    return list(range(rank, n_views, world))


class ShardedMapper:
    """Replicated flat parameter block + Adam state; gradient exchange over torch.distributed."""

    def __init__(self, P: int, sh_coeffs: int = 1, device="cpu", lrs: Optional[Dict[str, float]] = None,
                 betas=(0.9, 0.999), eps: float = 1e-15, group=None, optimizer_fn=None):
        """optimizer_fn(mapper, grad_flat): stand-in optimiser for host-logic tests on CPU tensors.  The product path
        (CUDA tensors) always runs the fused lvdgs_adam_step kernel; there is no CPU implementation in this package."""
        self.P, self.M = P, sh_coeffs
        self.optimizer_fn = optimizer_fn
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        widths = group_widths(sh_coeffs)
        total = sum(widths.values()) * P
        self.param_flat = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.exp_avg = torch.zeros_like(self.param_flat)
        self.exp_avg_sq = torch.zeros_like(self.param_flat)
        self.lr_flat = torch.empty_like(self.param_flat)
        # 3DGS / MonoGS default learning rates (the reference reads them from configs/mono/*/base_config.yaml opt_params)
        lrs = lrs or {"means3D": 1.6e-4, "shs": 2.5e-3, "opacity": 5e-2, "scales": 1e-3, "rotations": 1e-3}
        self.params: Dict[str, torch.Tensor] = {}
        self.slices: Dict[str, slice] = {}
        off = 0
        for name in GROUPS:
            n = widths[name] * P
            self.slices[name] = slice(off, off + n)
            self.params[name] = self.param_flat[off:off + n]
            self.lr_flat[off:off + n] = lrs[name]
            off += n
        self.betas, self.eps, self.t = betas, eps, 0
        self.grad_norm_accum = torch.zeros(P, dtype=torch.float32, device=self.device)
        self.denom = torch.zeros(P, dtype=torch.float32, device=self.device)
        self.max_radii2D = torch.zeros(P, dtype=torch.float32, device=self.device)

    # ---- views of the parameter block in the rasterizer's input layout ----
    def view(self, name: str) -> torch.Tensor:
        w = group_widths(self.M)[name]
        t = self.params[name]
        if name == "shs":
            return t.view(self.P, self.M, 3)
        return t.view(self.P, w)

    def load(self, **arrays):
        for name, a in arrays.items():
            self.params[name].copy_(torch.as_tensor(a, dtype=torch.float32).reshape(-1))

    # ---- collectives ----
    def reduce_gradients(self, grad_flat: torch.Tensor) -> torch.Tensor:
        """SUM over ranks of the contiguous gradient block (the mapping loss is a sum over views)."""
        if self.world > 1:
            dist.all_reduce(grad_flat, op=dist.ReduceOp.SUM, group=self.group)
        return grad_flat

    def reduce_stats(self):
        """Densification statistics: SUM the per-view gradient-norm accumulators and counters, MAX the radii."""
        if self.world > 1:
            dist.all_reduce(self.grad_norm_accum, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(self.denom, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(self.max_radii2D, op=dist.ReduceOp.MAX, group=self.group)

    def add_densification_stats(self, viewspace_grad_xy: torch.Tensor, visible: torch.Tensor, radii: torch.Tensor):
        """Local accumulation for one of this rank's views (GaussianModel.add_densification_stats, called once per
        view at utils/slam_backend.py:350-357: the norm is taken per view, before any summation)."""
        self.grad_norm_accum += torch.where(visible, viewspace_grad_xy.norm(dim=-1), torch.zeros_like(self.denom))
        self.denom += visible.to(torch.float32)
        self.max_radii2D = torch.where(visible, torch.maximum(self.max_radii2D, radii.to(torch.float32)), self.max_radii2D)

    # ---- optimiser ----
    def adam_step(self, grad_flat: torch.Tensor):
        """Adam on the whole block with per-group learning rates: one fused kernel of the C ABI (lvdgs_adam_step).
        Identical inputs on every rank -> identical result, so the replicas never need a broadcast."""
        self.t += 1
        if self.optimizer_fn is not None:
            return self.optimizer_fn(self, grad_flat)
        if not self.param_flat.is_cuda:
            raise RuntimeError("ShardedMapper.adam_step: parameters must live on a CUDA device (no CPU path)")
        import ctypes as C
        from . import _native
        L = _native.lib()
        ends = (C.c_int64 * len(GROUPS))(*[self.slices[n].stop for n in GROUPS])
        if not hasattr(self, "_lrs_c"):
            self._lrs_c = (C.c_float * len(GROUPS))(*[float(self.lr_flat[self.slices[n].start]) for n in GROUPS])
        rc = L.lvdgs_adam_step(self.param_flat.numel(), _native.ptr(self.param_flat), _native.ptr(grad_flat),
                               _native.ptr(self.exp_avg), _native.ptr(self.exp_avg_sq), len(GROUPS), ends, self._lrs_c,
                               self.betas[0], self.betas[1], self.eps, self.t,
                               C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        _native.check(rc, "lvdgs_adam_step")

    # ---- map maintenance ----
    def prune(self, keep: torch.Tensor) -> int:
        """GaussianModel.prune_points on the replicated block (callers utils/slam_backend.py:128-145,322-339): drops the
        rows where `keep` is zero from every parameter group, both Adam moments and the densification statistics -- one
        mask scan + two launches (lvdgs_compact_*) instead of one torch index kernel per tensor.  Every rank must call it
        with the identical mask (it is derived from all-reduced statistics), so the replicas stay identical.
        Returns the new number of Gaussians."""
        if not self.param_flat.is_cuda:
            raise RuntimeError("ShardedMapper.prune: parameters must live on a CUDA device (no CPU path)")
        from .slam_ops import compact_rows
        widths = group_widths(self.M)
        srcs = []
        for flat in (self.param_flat, self.exp_avg, self.exp_avg_sq):
            srcs += [flat[self.slices[n]].view(self.P, widths[n]) for n in GROUPS]
        srcs += [self.grad_norm_accum.view(self.P, 1), self.denom.view(self.P, 1), self.max_radii2D.view(self.P, 1)]
        new = compact_rows(keep, srcs)
        P2 = new[0].shape[0]
        ng = len(GROUPS)
        lr_of = {n: float(self.lr_flat[self.slices[n].start]) if self.P else 0.0 for n in GROUPS}
        flats = [torch.cat([t.reshape(-1) for t in new[i * ng:(i + 1) * ng]]) if P2 else
                 torch.zeros(0, dtype=torch.float32, device=self.device) for i in range(3)]
        self.param_flat, self.exp_avg, self.exp_avg_sq = flats
        self.grad_norm_accum, self.denom, self.max_radii2D = (t.reshape(-1) for t in new[3 * ng:])
        self.lr_flat = torch.empty_like(self.param_flat)
        self.P, off = P2, 0
        for name in GROUPS:
            n = widths[name] * P2
            self.slices[name] = slice(off, off + n)
            self.params[name] = self.param_flat[off:off + n]
            self.lr_flat[off:off + n] = lr_of[name]
            off += n
        return P2

    def densify_clone(self, index: torch.Tensor, overrides: Optional[Dict[str, torch.Tensor]] = None) -> int:
        """Appends copies of the Gaussians `index` (int64) to the replicated block -- GaussianModel.densify_and_clone, and
        with `overrides` = {"means3D": new positions, "scales": new scales} the append half of densify_and_split (callers
        utils/slam_backend.py:359-376).  New rows get zero Adam moments and zero densification statistics, like
        densification_postfix.  The rows of all groups are gathered in one launch (lvdgs_gather_rows) straight into the new
        block.  Every rank must call it with identical arguments.  Returns the new number of Gaussians."""
        if not self.param_flat.is_cuda:
            raise RuntimeError("ShardedMapper.densify_clone: parameters must live on a CUDA device (no CPU path)")
        from .slam_ops import gather_rows
        widths = group_widths(self.M)
        k = int(index.numel())
        P2 = self.P + k
        lr_of = {n: float(self.lr_flat[self.slices[n].start]) if self.P else 0.0 for n in GROUPS}
        new_flat = torch.empty(sum(widths.values()) * P2, dtype=torch.float32, device=self.device)
        new_m, new_v = torch.zeros_like(new_flat), torch.zeros_like(new_flat)
        new_slices, off, tails, srcs = {}, 0, [], []
        for name in GROUPS:
            w = widths[name]
            new_slices[name] = slice(off, off + w * P2)
            old = slice(self.slices[name].start, self.slices[name].stop)
            new_flat[off:off + w * self.P] = self.param_flat[old]
            new_m[off:off + w * self.P] = self.exp_avg[old]
            new_v[off:off + w * self.P] = self.exp_avg_sq[old]
            tails.append(new_flat[off + w * self.P:off + w * P2].view(k, w))
            srcs.append(self.param_flat[old].view(self.P, w))
            off += w * P2
        gather_rows(index, srcs, out=tails)
        for name, t in (overrides or {}).items():
            tails[GROUPS.index(name)].copy_(t.reshape(k, widths[name]))
        pad = torch.zeros(k, dtype=torch.float32, device=self.device)
        self.grad_norm_accum = torch.cat([self.grad_norm_accum, pad])
        self.denom = torch.cat([self.denom, pad])
        self.max_radii2D = torch.cat([self.max_radii2D, pad])
        self.param_flat, self.exp_avg, self.exp_avg_sq = new_flat, new_m, new_v
        self.lr_flat = torch.empty_like(new_flat)
        self.P, self.slices = P2, new_slices
        for name in GROUPS:
            self.params[name] = self.param_flat[self.slices[name]]
            self.lr_flat[self.slices[name]] = lr_of[name]
        return P2

    # ---- one mapping iteration ----
    def step(self, n_views: int, render_and_grad: Callable[[int], None], grad_flat: torch.Tensor,
             zero: Optional[Callable[[], None]] = None, extra_views: Sequence[int] = ()):
        """render_and_grad(k) must ADD view k's parameter gradients into `grad_flat`.
        `extra_views`: the reference's 2 random older keyframes (utils/slam_backend.py:275); the caller draws them
        with a generator seeded identically on every rank, they are sharded like the window."""
        if zero is not None:
            zero()
        else:
            grad_flat.zero_()
        views = list(range(n_views)) + list(extra_views)
        mine = [views[i] for i in shard_keyframes(len(views), self.world, self.rank)]
        for k in mine:
            render_and_grad(k)
        self.reduce_gradients(grad_flat)
        self.adam_step(grad_flat)
        return mine
