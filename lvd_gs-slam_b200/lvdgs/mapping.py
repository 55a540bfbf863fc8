"""Keyframe-sharded mapping iteration over several GPUs -- NEW functionality: the reference maps
on one GPU (README.md:93-104, CUDA_VISIBLE_DEVICES=0).

One mapping iteration of the reference (utils/slam_backend.py:153-389) renders every keyframe of the window
(<= 8, plus 2 random older keyframes) against the same Gaussian map, SUMS the per-view losses
(`loss_mapping +=`, :266/:300), back-propagates once and takes one Adam step.  The views are independent given the
map, so here: one process per GPU, the map replicated, keyframe k owned by rank k mod world; every rank accumulates
the parameter gradients of its views into one contiguous float32 block (lvdgs.engine.RasterEngine.grad_flat: the
backward kernel adds into it, there is no pack kernel).  The exchange step per iteration (NCCL):

    reduce-scatter (SUM) of the gradient block  ->  fused Adam on this rank's 1/world slice  ->  all-gather of the
    updated parameters,

i.e. the optimiser state is sharded (ZeRO-1): the same bytes cross NVLink as in an all-reduce, but the Adam kernel
touches 1/world of the block on every rank and the gradient block is re-zeroed off the critical path.
On NVLink-connected GPUs of one node the three steps (plus the activation chain rule before) are ONE kernel
(lvdgs_exchange_adam, csrc/exchange.cu): the parameter, activation and gradient blocks live in symmetric memory
(torch.distributed._symmetric_memory); every rank sums its slice of all gradient blocks inside the NVSwitch
(multimem.ld_reduce on the multicast mapping; peer loads without multicast support or at two ranks), updates it and
writes the new raw parameters once, replicated by the switch to every rank (multimem.st; peer stores otherwise); two
device-side barriers bracket the launch, the activations are recomputed locally behind the second one
(LVDGS_EXCHANGE_LOCAL_ACT=0: stored by the kernel instead), and the gradient block can be cleared on a side stream
(`defer_zero`, `grad_ready`).  On ONE GPU the same kernel with a world of one replaces the three launches (chain rule,
Adam, activations).  LVDGS_P2P_EXCHANGE=0, or a failed rendezvous, selects the NCCL sequence above; LVDGS_MULTICAST=0/1
forces peer / multicast access.  With a backend
that has no reduce-scatter (gloo, the CPU tests) the block is all-reduced and every rank runs the identical full update.
Either way the replicas stay bit-identical without a broadcast.

Parametrisation: like the reference's GaussianModel the block holds RAW parameters -- logit opacity, log
scale, un-normalised quaternion -- and the rasterizer is fed their activations (lvdgs_gaussian_activate); the gradients
it returns are taken back through the activations (lvdgs_gaussian_activation_backward) before the exchange, so Adam
works on the same variables as `GaussianModel.optimizer` (opacity can never leave (0, 1), rotations stay unit).

Densification statistics ride along: the accumulated 2-D gradient norms and their denominators are SUM-reduced,
`max_radii2D` is MAX-reduced (utils/slam_backend.py:350-357).  Per-keyframe visibility masks (`n_touched > 0`, :311-315)
stay on the owning rank.

The collective / sharding logic is device-agnostic on purpose: the GPU path hands it RasterEngine's gradient block
over NCCL, the CPU tests (gloo, world_size 2) hand it a stand-in gradient function and a stand-in optimiser.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from .engine import GROUPS, block_layout

ACTIVATED = ("opacity", "scales", "rotations")     # groups whose raw value differs from what the rasterizer consumes


def group_widths(sh_coeffs: int = 1) -> Dict[str, int]:
    return {"means3D": 3, "shs": 3 * sh_coeffs, "opacity": 1, "scales": 3, "rotations": 4}


def shard_keyframes(n_views: int, world: int, rank: int) -> List[int]:
    """Keyframe k is rendered by rank k mod world (round-robin keeps ranks balanced when n_views % world != 0)."""
    return list(range(rank, n_views, world))


class ShardedMapper:
    """Replicated flat parameter block + Adam state; gradient exchange over torch.distributed."""

    def __init__(self, P: int, sh_coeffs: int = 1, device="cpu", lrs: Optional[Dict[str, float]] = None,
                 betas=(0.9, 0.999), eps: float = 1e-15, group=None, optimizer_fn=None, raw: bool = True):
        """raw: the block holds GaussianModel's raw parameters and `view()` returns their activations (the default).
        raw=False keeps the round-1 behaviour (the block holds what the rasterizer consumes) for host-logic tests.
        optimizer_fn(mapper, grad_flat): stand-in optimiser for host-logic tests on CPU tensors.  The product path
        (CUDA tensors) always runs the fused lvdgs_adam_step kernel; there is no CPU implementation in this package."""
        self.M = sh_coeffs
        self.optimizer_fn = optimizer_fn
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.raw = bool(raw) and self.device.type == "cuda"
        # 3DGS / MonoGS default learning rates (the reference reads them from configs/mono/*/base_config.yaml opt_params)
        self.lrs = dict(lrs or {"means3D": 1.6e-4, "shs": 2.5e-3, "opacity": 5e-2, "scales": 1e-3, "rotations": 1e-3})
        self.betas, self.eps, self.t = betas, eps, 0
        self._p2p = None           # symmetric-memory handles of the peer-memory exchange (None: NCCL sequence)
        self.exchange_mode = "nccl"     # "multicast" / "peer" once the peer-memory exchange has run
        import os
        # activations of the updated parameters: recomputed locally after the exchange (default) or stored by the exchange
        # kernel into every rank's activated block (LVDGS_EXCHANGE_LOCAL_ACT=0)
        self.local_activation = os.environ.get("LVDGS_EXCHANGE_LOCAL_ACT", "1") != "0"
        self.grad_ready = None          # with defer_zero: the event after which the gradient block is zero again
        self.time_exchange, self._phase_log = False, []
        self._grad_block = None
        self._alloc(P)
        self.grad_norm_accum = torch.zeros(P, dtype=torch.float32, device=self.device)
        self.denom = torch.zeros(P, dtype=torch.float32, device=self.device)
        self.max_radii2D = torch.zeros(P, dtype=torch.float32, device=self.device)
        self._zero_stream = None
        self._zero_done = None

    # ---- storage ----
    def _want_p2p(self) -> bool:
        import os
        return (self.world > 1 and self.device.type == "cuda" and self.raw and self.optimizer_fn is None
                and os.environ.get("LVDGS_P2P_EXCHANGE", "1") != "0" and dist.is_initialized()
                and dist.get_backend(self.group) == "nccl")

    def _alloc(self, P: int):
        self.P = P
        self.layout, self.total = block_layout(P, self.M, multiple=4 * self.world)
        z = lambda n: torch.zeros(n, dtype=torch.float32, device=self.device)
        self.exp_avg, self.exp_avg_sq = z(self.total), z(self.total)
        self._p2p, self._grad_block = None, None
        self._symm = {}
        if self._want_p2p():
            # parameter / activation / gradient blocks in symmetric memory: every rank can address every rank's copy
            try:
                import torch.distributed._symmetric_memory as symm_mem
                act_len = sum((self.layout[n][1] + 3) & ~3 for n in ACTIVATED) + 4 * self.world
                for name, n in (("param", self.total), ("act", max(act_len, 4)), ("grad", self.total)):
                    t = symm_mem.empty(n, dtype=torch.float32, device=self.device)
                    t.zero_()
                    self._symm[name] = (t, symm_mem.rendezvous(t, self.group if self.group is not None else dist.group.WORLD))
                self._p2p = {k: h for k, (t, h) in self._symm.items()}
            except Exception as e:       # no peer access / no symmetric-memory support on this box: NCCL sequence
                import warnings
                warnings.warn(f"lvdgs: peer-memory exchange unavailable ({type(e).__name__}: {e}); using NCCL reduce-scatter / all-gather")
                self._symm, self._p2p = {}, None
        self.param_flat = self._symm["param"][0] if self._p2p else z(self.total)
        self._bind()
        self.moments_sharded = False

    def _bind(self):
        self.slices = {n: slice(off, off + ln) for n, (off, ln) in self.layout.items()}
        self.params = {n: self.param_flat[sl] for n, sl in self.slices.items()}
        if self.raw:      # activated copies of the three groups the rasterizer does not take raw
            self.act_layout, act_total, off = {}, 0, 0
            for n in ACTIVATED:
                ln = self.layout[n][1]
                self.act_layout[n] = (off, ln)
                off += (ln + 3) & ~3
            self.act_flat = self._symm["act"][0] if self._p2p else torch.zeros(max(off, 4), dtype=torch.float32, device=self.device)
            self.act = {n: self.act_flat[o:o + ln] for n, (o, ln) in self.act_layout.items()}
        if self.optimizer_fn is not None:      # per-element learning rate, only for the stand-in optimisers of the CPU tests
            self.lr_flat = torch.zeros_like(self.param_flat)
            for n, sl in self.slices.items():
                self.lr_flat[sl] = float(self.lrs[n])
        # reduce-scatter slice of this rank
        self.shard_len = self.total // self.world
        self.shard = slice(self.rank * self.shard_len, (self.rank + 1) * self.shard_len)

    def new_grad_block(self) -> torch.Tensor:
        """A zeroed gradient block with this mapper's layout and padding (hand it to RasterEngine(grad_flat=...))."""
        if self._p2p:              # the symmetric gradient block (one per mapper): the peer-memory exchange reads it remotely
            self._grad_block = self._symm["grad"][0]
            self._grad_block.zero_()
            return self._grad_block
        return torch.zeros(self.total, dtype=torch.float32, device=self.device)

    # ---- views of the parameter block in the rasterizer's input layout ----
    def view(self, name: str) -> torch.Tensor:
        w = group_widths(self.M)[name]
        t = self.act[name] if (self.raw and name in ACTIVATED) else self.params[name]
        if name == "shs":
            return t.view(self.P, self.M, 3)
        return t.view(self.P, w)

    def raw_view(self, name: str) -> torch.Tensor:
        return self.params[name].view(self.P, -1)

    def load(self, **arrays):
        """Loads ACTIVATED values (opacity in (0,1), positive scales, rotations as given), as a scene provides them; with
        raw storage they are converted once: logit, log, identity (GaussianModel.create_pcd does the same)."""
        for name, a in arrays.items():
            v = torch.as_tensor(a, dtype=torch.float32).reshape(-1).to(self.device)
            if self.raw and name == "opacity":
                v = torch.log(v / (1.0 - v))
            elif self.raw and name == "scales":
                v = torch.log(v)
            self.params[name].copy_(v)
        self.activate()

    def _lib(self):
        import ctypes as C
        from . import _native
        return C, _native, _native.lib()

    def _stream(self):
        import ctypes as C
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def activate(self):
        """raw -> what the rasterizer consumes (sigmoid / exp / normalize), one fused kernel."""
        if not self.raw or self.P == 0:
            return
        C, _native, L = self._lib()
        rc = L.lvdgs_gaussian_activate(self.P, _native.ptr(self.params["opacity"]), _native.ptr(self.params["scales"]),
                                       _native.ptr(self.params["rotations"]), _native.ptr(self.act["opacity"]),
                                       _native.ptr(self.act["scales"]), _native.ptr(self.act["rotations"]), self._stream())
        _native.check(rc, "lvdgs_gaussian_activate")

    def activation_backward(self, grad_flat: torch.Tensor):
        """gradients w.r.t. the activated values (what the backward kernel accumulated) -> w.r.t. the raw parameters."""
        if not self.raw or self.P == 0:
            return
        C, _native, L = self._lib()
        g = lambda n: _native.ptr(grad_flat[self.slices[n]])
        rc = L.lvdgs_gaussian_activation_backward(self.P, _native.ptr(self.act["opacity"]), _native.ptr(self.act["scales"]),
                                                  _native.ptr(self.act["rotations"]), _native.ptr(self.params["rotations"]),
                                                  g("opacity"), g("scales"), g("rotations"), self._stream())
        _native.check(rc, "lvdgs_gaussian_activation_backward")

    # ---- collectives ----
    def _can_scatter(self) -> bool:
        return self.world > 1 and self.param_flat.is_cuda and dist.get_backend(self.group) == "nccl"

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
    def _adam_range(self, grads: torch.Tensor, lo: int, hi: int):
        """Fused Adam (lvdgs_adam_step) on block elements [lo, hi) with per-group learning rates; `grads` holds the
        gradients of exactly that range."""
        C, _native, L = self._lib()
        ends, lr = [], []
        for n in GROUPS:      # group ends clipped to the range, relative to its start (padding inherits the group before it)
            off, ln = self.layout[n]
            nxt = min((o for o, _ in self.layout.values() if o > off), default=self.total)
            ends.append(min(max(nxt, lo), hi) - lo)
            lr.append(float(self.lrs[n]))
        ends[-1] = hi - lo
        rc = L.lvdgs_adam_step(hi - lo, _native.ptr(self.param_flat[lo:hi]), _native.ptr(grads), _native.ptr(self.exp_avg[lo:hi]),
                               _native.ptr(self.exp_avg_sq[lo:hi]), len(GROUPS), (C.c_int64 * len(GROUPS))(*ends),
                               (C.c_float * len(GROUPS))(*lr), self.betas[0], self.betas[1], self.eps, self.t, self._stream())
        _native.check(rc, "lvdgs_adam_step")

    def adam_step(self, grad_flat: torch.Tensor):
        """Adam on the whole block with an already reduced gradient: one fused kernel of the C ABI (lvdgs_adam_step).
        Identical inputs on every rank -> identical result, so the replicas never need a broadcast."""
        self.t += 1
        if self.optimizer_fn is not None:
            return self.optimizer_fn(self, grad_flat)
        if not self.param_flat.is_cuda:
            raise RuntimeError("ShardedMapper.adam_step: parameters must live on a CUDA device (no CPU path)")
        self._adam_range(grad_flat, 0, self.total)
        self.activate()

    def exchange_and_update(self, grad_flat: torch.Tensor, defer_zero: bool = False):
        """The exchange step of one iteration: raw-parameter chain rule, gradient SUM over the ranks, Adam, activations.
        NCCL: reduce-scatter -> Adam on this rank's slice -> all-gather of the parameters; the gradient block is zeroed
        for the next iteration on a side stream while the all-gather runs."""
        if self._p2p and grad_flat is self._grad_block:
            return self._exchange_p2p(grad_flat, defer_zero)
        if (self.world == 1 and self.raw and self.optimizer_fn is None and self.param_flat.is_cuda and grad_flat.is_cuda
                and grad_flat.numel() >= self.total and not (grad_flat.data_ptr() & 15)):
            return self._exchange_single(grad_flat)
        self.activation_backward(grad_flat)
        if not self._can_scatter():
            self.reduce_gradients(grad_flat)
            self.adam_step(grad_flat)
            grad_flat.zero_()
            return
        self.t += 1
        if not hasattr(self, "_g_shard") or self._g_shard.numel() != self.shard_len:
            self._g_shard = torch.empty(self.shard_len, dtype=torch.float32, device=self.device)
        dist.reduce_scatter_tensor(self._g_shard, grad_flat, op=dist.ReduceOp.SUM, group=self.group)
        cur = torch.cuda.current_stream(self.device)
        if self._zero_stream is None:
            self._zero_stream = torch.cuda.Stream(self.device)
            self._zero_done = torch.cuda.Event()
        self._zero_stream.wait_stream(cur)                       # the reduce-scatter has consumed the block
        with torch.cuda.stream(self._zero_stream):
            grad_flat.zero_()
            self._zero_done.record(self._zero_stream)
        self._adam_range(self._g_shard, self.shard.start, self.shard.stop)
        self.moments_sharded = True                              # only this rank's slice of exp_avg / exp_avg_sq is current
        dist.all_gather_into_tensor(self.param_flat, self.param_flat[self.shard], group=self.group)   # in place
        self.activate()
        cur.wait_event(self._zero_done)

    def _group_ends(self):
        ends = []
        for n in GROUPS:
            off, _ = self.layout[n]
            ends.append(min((o for o, _ in self.layout.values() if o > off), default=self.total))
        ends[-1] = self.total
        return ends

    def _exchange_single(self, grad_flat: torch.Tensor):
        """One GPU: the same kernel with a world of one -- chain rule, Adam and activations in ONE launch instead of three
        (lvdgs_gaussian_activation_backward, lvdgs_adam_step, lvdgs_gaussian_activate)."""
        C, _native, L = self._lib()
        self.t += 1
        one = lambda t: (C.c_void_p * 1)(t.data_ptr())
        ends = (C.c_int64 * len(GROUPS))(*self._group_ends())
        lr = (C.c_float * len(GROUPS))(*[float(self.lrs[n]) for n in GROUPS])
        act_off = (C.c_int64 * 3)(*[self.act_layout[n][0] for n in ACTIVATED])
        rc = L.lvdgs_exchange_adam(1, 0, one(grad_flat), one(self.param_flat), one(self.act_flat), 0, self.total,
                                   _native.ptr(self.exp_avg), _native.ptr(self.exp_avg_sq), len(GROUPS), ends, lr, act_off,
                                   self.act_flat.numel(), self.betas[0], self.betas[1], self.eps, self.t, None, None, None, 1,
                                   self._stream())
        _native.check(rc, "lvdgs_exchange_adam")
        grad_flat.zero_()

    def _exchange_p2p(self, grad_flat: torch.Tensor, defer_zero: bool = False):
        """barrier | ONE kernel: peer reads of the gradient slice, chain rule, Adam, peer stores of parameters + activations | barrier."""
        C, _native, L = self._lib()
        self.t += 1
        hp, ha, hg = self._p2p["param"], self._p2p["act"], self._p2p["grad"]
        if not hasattr(self, "_p2p_args") or self._p2p_args[0] is not hp:
            import os
            W = self.world
            base = lambda h: [int(x) + int(getattr(h, "offset", 0)) for x in h.buffer_ptrs]
            arr = lambda h: (C.c_void_p * W)(*base(h))
            ends = self._group_ends()
            act_off = (C.c_int64 * 3)(*[self.act_layout[n][0] for n in ACTIVATED])
            # NVSwitch multicast mappings of the three blocks (0 when the box has no multicast support): the sum over the
            # ranks and the replication of the results then happen inside the switch (multimem.ld_reduce / multimem.st)
            mc = [int(getattr(h, "multicast_ptr", 0) or 0) for h in (hg, hp, ha)]
            # (two ranks: every byte crosses the switch once either way, and peer loads / stores were measured faster)
            if os.environ.get("LVDGS_MULTICAST", "1" if W > 2 else "0") == "0" or not all(mc):
                mc = [0, 0, 0]
            else:
                mc = [m + int(getattr(h, "offset", 0)) for m, h in zip(mc, (hg, hp, ha))]
            self.exchange_mode = "multicast" if mc[0] else "peer"
            self._p2p_args = (hp, arr(hg), arr(hp), arr(ha), (C.c_int64 * len(GROUPS))(*ends), act_off,
                              [C.c_void_p(m) if m else None for m in mc])
        _, g_arr, p_arr, a_arr, ends, act_off, mc = self._p2p_args
        lr = (C.c_float * len(GROUPS))(*[float(self.lrs[n]) for n in GROUPS])
        ev = self._phase_events(5) if self.time_exchange else None
        if ev: ev[0].record()
        hg.barrier(channel=0)                                  # every rank's views have accumulated into its gradient block
        if ev: ev[1].record()
        rc = L.lvdgs_exchange_adam(self.world, self.rank, g_arr, p_arr, a_arr, self.shard.start, self.shard.stop,
                                   _native.ptr(self.exp_avg), _native.ptr(self.exp_avg_sq), len(GROUPS), ends, lr, act_off,
                                   self.act_flat.numel(), self.betas[0], self.betas[1], self.eps, self.t, mc[0], mc[1], mc[2],
                                   2 if self.local_activation else 1, self._stream())
        _native.check(rc, "lvdgs_exchange_adam")
        self.moments_sharded = True
        if ev: ev[2].record()
        hg.barrier(channel=1)                                  # all stores have landed, all gradient slices have been read
        if ev: ev[3].record()
        cur = torch.cuda.current_stream(self.device)
        if defer_zero:
            # the gradient block is cleared on a side stream; the caller makes its first backward wait for `grad_ready`
            # (RasterEngine.run_views(..., bwd_wait=mapper.grad_ready)) instead of this stream
            if self._zero_stream is None:
                self._zero_stream = torch.cuda.Stream(self.device)
            after = torch.cuda.Event(); after.record(cur)
            self._zero_stream.wait_event(after)
            with torch.cuda.stream(self._zero_stream):
                grad_flat.zero_()
                self.grad_ready = torch.cuda.Event(); self.grad_ready.record(self._zero_stream)
        else:
            grad_flat.zero_()
            self.grad_ready = None
        if self.local_activation:
            self.activate()                                    # one local pass instead of 8 of 14 floats per Gaussian over NVLink
        if ev: ev[4].record()

    def _phase_events(self, n):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        self._phase_log.append(evs)
        return evs

    def exchange_phase_ms(self):
        """Median device time of the phases of the peer-memory exchange recorded since `time_exchange` was switched on:
        barrier | exchange kernel | barrier | gradient-block memset."""
        import statistics
        torch.cuda.synchronize(self.device)
        rows = [[a.elapsed_time(b) for a, b in zip(e[:-1], e[1:])] for e in self._phase_log]
        self._phase_log = []
        if not rows:
            return None
        names = ("barrier_before", "exchange_kernel", "barrier_after", "zero_gradients_and_local_activation")
        return {n: statistics.median(r[k] for r in rows) for k, n in enumerate(names)}

    def gather_moments(self):
        """Makes exp_avg / exp_avg_sq complete on every rank again (before prune / densify move rows around)."""
        if self.moments_sharded and self.world > 1:
            for m in (self.exp_avg, self.exp_avg_sq):
                dist.all_gather_into_tensor(m, m[self.shard].clone(), group=self.group)
        self.moments_sharded = False

    # ---- map maintenance ----
    def _rows(self, flat):
        w = group_widths(self.M)
        return [flat[self.slices[n]].view(self.P, w[n]) for n in GROUPS]

    def _rebuild(self, P2: int, new_rows, stats):
        """Installs new per-group row tensors (params, exp_avg, exp_avg_sq: 3 x 5 tensors of P2 rows) as the block."""
        ng = len(GROUPS)
        self._alloc(P2)
        for i, flat in enumerate((self.param_flat, self.exp_avg, self.exp_avg_sq)):
            for j, n in enumerate(GROUPS):
                if P2:
                    flat[self.slices[n]].copy_(new_rows[i * ng + j].reshape(-1))
        self.grad_norm_accum, self.denom, self.max_radii2D = stats
        self.activate()

    def prune(self, keep: torch.Tensor) -> int:
        """GaussianModel.prune_points on the replicated block (callers utils/slam_backend.py:128-145,322-339): drops the
        rows where `keep` is zero from every parameter group, both Adam moments and the densification statistics -- one
        mask scan + two launches (lvdgs_compact_*) instead of one torch index kernel per tensor.  Every rank must call it
        with the identical mask (it is derived from all-reduced statistics), so the replicas stay identical.
        Returns the new number of Gaussians."""
        if not self.param_flat.is_cuda:
            raise RuntimeError("ShardedMapper.prune: parameters must live on a CUDA device (no CPU path)")
        from .slam_ops import compact_rows
        self.gather_moments()
        srcs = self._rows(self.param_flat) + self._rows(self.exp_avg) + self._rows(self.exp_avg_sq)
        srcs += [self.grad_norm_accum.view(self.P, 1), self.denom.view(self.P, 1), self.max_radii2D.view(self.P, 1)]
        new = compact_rows(keep, srcs)
        P2 = new[0].shape[0]
        ng = len(GROUPS)
        self._rebuild(P2, new[:3 * ng], tuple(t.reshape(-1) for t in new[3 * ng:]))
        return P2

    def densify_clone(self, index: torch.Tensor, overrides: Optional[Dict[str, torch.Tensor]] = None) -> int:
        """Appends copies of the Gaussians `index` (int64) to the replicated block -- GaussianModel.densify_and_clone, and
        with `overrides` = {"means3D": new positions, "scales": new scales} the append half of densify_and_split (callers
        utils/slam_backend.py:359-376).  Overrides are in the block's own parametrisation (raw log-scales with raw=True).
        New rows get zero Adam moments and zero densification statistics, like densification_postfix.  The rows of all
        groups are gathered in one launch (lvdgs_gather_rows).  Every rank must call it with identical arguments.
        Returns the new number of Gaussians."""
        if not self.param_flat.is_cuda:
            raise RuntimeError("ShardedMapper.densify_clone: parameters must live on a CUDA device (no CPU path)")
        from .slam_ops import gather_rows
        self.gather_moments()
        widths = group_widths(self.M)
        k = int(index.numel())
        old_P = self.P
        old = (self._rows(self.param_flat), self._rows(self.exp_avg), self._rows(self.exp_avg_sq))
        tails = [torch.empty(k, widths[n], dtype=torch.float32, device=self.device) for n in GROUPS]
        gather_rows(index, old[0], out=tails)
        for name, t in (overrides or {}).items():
            tails[GROUPS.index(name)].copy_(t.reshape(k, widths[name]))
        zero = lambda n: torch.zeros(k, widths[n], dtype=torch.float32, device=self.device)
        rows = [torch.cat([o, t]) for o, t in zip(old[0], tails)]
        rows += [torch.cat([o, zero(n)]) for o, n in zip(old[1], GROUPS)]
        rows += [torch.cat([o, zero(n)]) for o, n in zip(old[2], GROUPS)]
        pad = torch.zeros(k, dtype=torch.float32, device=self.device)
        stats = (torch.cat([self.grad_norm_accum, pad]), torch.cat([self.denom, pad]), torch.cat([self.max_radii2D, pad]))
        self._rebuild(old_P + k, rows, stats)
        return self.P

    # ---- GaussianModel maintenance on the block (GaussianModel itself is not in /root/reference, the
    # semantics are those of the 3DGS / MonoGS model its callers assume) ----
    @staticmethod
    def _inverse_sigmoid(x):
        return torch.log(x / (1.0 - x))

    def extend_from_points(self, points: torch.Tensor, colors: torch.Tensor, point_size: float = 1.0,
                           init_opacity: float = 0.5, kf_id: int = -1) -> int:
        """GaussianModel.extend_from_pcd_seq (caller utils/slam_backend.py:75-78), after the keyframe has been back-projected:
        appends one Gaussian per point -- isotropic scale from the mean squared distance to the 3 nearest neighbours
        (simple_knn.distCUDA2 -> lvdgs_dist2; log(sqrt(clamp_min(d2, 1e-7) * point_size))), identity rotation, opacity
        inverse_sigmoid(0.5), SH-0 colour (rgb - 0.5) / C0 -- with zero Adam moments.  points [n,3], colors [n,3] in [0,1].
        Every rank must call it with identical arguments.  Returns the new number of Gaussians."""
        if not self.param_flat.is_cuda or not self.raw:
            raise RuntimeError("ShardedMapper.extend_from_points: needs the raw-parameter block on a CUDA device (no CPU path)")
        from simple_knn._C import distCUDA2
        pts = points.to(self.device, torch.float32).contiguous()
        n = pts.shape[0]
        d2 = torch.clamp_min(distCUDA2(pts), 1e-7) * point_size
        new = {"means3D": pts, "scales": torch.log(torch.sqrt(d2))[:, None].repeat(1, 3),
               "rotations": torch.tensor([1.0, 0, 0, 0], device=self.device).repeat(n, 1),
               "opacity": self._inverse_sigmoid(torch.full((n, 1), float(init_opacity), device=self.device))}
        shs = torch.zeros(n, self.M, 3, device=self.device)
        shs[:, 0] = (colors.to(self.device, torch.float32) - 0.5) / 0.28209479177387814
        new["shs"] = shs.reshape(n, -1)
        self.gather_moments()
        old = (self._rows(self.param_flat), self._rows(self.exp_avg), self._rows(self.exp_avg_sq))
        widths = group_widths(self.M)
        zero = lambda name: torch.zeros(n, widths[name], dtype=torch.float32, device=self.device)
        rows = [torch.cat([o, new[name].reshape(n, widths[name])]) for o, name in zip(old[0], GROUPS)]
        rows += [torch.cat([o, zero(name)]) for o, name in zip(old[1], GROUPS)]
        rows += [torch.cat([o, zero(name)]) for o, name in zip(old[2], GROUPS)]
        pad = torch.zeros(n, dtype=torch.float32, device=self.device)
        stats = (torch.cat([self.grad_norm_accum, pad]), torch.cat([self.denom, pad]), torch.cat([self.max_radii2D, pad]))
        kf = getattr(self, "unique_kfIDs", torch.full((self.P,), -1, dtype=torch.int32, device=self.device))
        self._rebuild(self.P + n, rows, stats)
        self.unique_kfIDs = torch.cat([kf, torch.full((n,), int(kf_id), dtype=torch.int32, device=self.device)])
        return self.P

    def reset_opacity_nonvisible(self, visibility_filters: Sequence[torch.Tensor], value: float = 0.4):
        """GaussianModel.reset_opacity_nonvisible (caller utils/slam_backend.py:372-376): every Gaussian that none of the
        given views saw gets opacity `value` (raw: inverse_sigmoid(0.4)); the opacity group's Adam moments restart from zero
        (replace_tensor_to_optimizer).  Visibility comes from lvdgs_n_obs over the masks (one launch)."""
        if not self.param_flat.is_cuda or not self.raw:
            raise RuntimeError("ShardedMapper.reset_opacity_nonvisible: needs the raw-parameter block on a CUDA device")
        from .slam_ops import accumulate_n_obs as n_obs
        seen = n_obs(list(visibility_filters)) > 0 if len(visibility_filters) else torch.zeros(self.P, dtype=torch.bool, device=self.device)
        raw = self.params["opacity"]
        raw.copy_(torch.where(seen, raw, self._inverse_sigmoid(torch.full_like(raw, float(value)))))
        self.gather_moments()
        self.exp_avg[self.slices["opacity"]].zero_()
        self.exp_avg_sq[self.slices["opacity"]].zero_()
        self.activate()

    def densify_and_prune(self, max_grad: float, min_opacity: float, extent: float, max_screen_size: Optional[float],
                          percent_dense: float = 0.01, n_split: int = 2, generator: Optional[torch.Generator] = None) -> int:
        """GaussianModel.densify_and_prune (caller utils/slam_backend.py:359-370) on the replicated block, from the all-reduced
        statistics (call reduce_stats() first when world > 1):
          grads = grad_norm_accum / denom;  clone where grads >= max_grad and max(scale) <= percent_dense * extent;
          split (n_split samples from N(0, scale) in the Gaussian's frame, scales / (0.8 n_split)) where grads >= max_grad and
          max(scale) > percent_dense * extent, the split originals removed;  prune opacity < min_opacity, and -- when
          max_screen_size is given -- max_radii2D > max_screen_size or max(scale) > 0.1 * extent.
        `generator`: seeded identically on every rank, so every replica draws the same split samples.  The statistics are
        reset afterwards.  Returns the new number of Gaussians."""
        if not self.param_flat.is_cuda or not self.raw:
            raise RuntimeError("ShardedMapper.densify_and_prune: needs the raw-parameter block on a CUDA device")
        grads = self.grad_norm_accum / self.denom
        grads[grads.isnan()] = 0.0
        max_scale = self.view("scales").max(dim=1).values
        hot = grads >= max_grad
        small = max_scale <= percent_dense * extent
        clone_idx = torch.nonzero(hot & small).reshape(-1)
        split_idx = torch.nonzero(hot & ~small).reshape(-1)
        P0 = self.P
        kf = getattr(self, "unique_kfIDs", torch.full((self.P,), -1, dtype=torch.int32, device=self.device))
        if clone_idx.numel():
            self.densify_clone(clone_idx)
            kf = torch.cat([kf, kf[clone_idx]])
        if split_idx.numel():
            rep = split_idx.repeat(n_split)
            scales = self.view("scales")[rep]
            samples = torch.randn(scales.shape, generator=generator, device=self.device if generator is None or generator.device.type == "cuda" else "cpu").to(self.device) * scales
            q = self.view("rotations")[rep]
            r, x, y, z = q.unbind(1)
            R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                             2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                             2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1).reshape(-1, 3, 3)
            new_xyz = torch.bmm(R, samples[..., None]).squeeze(-1) + self.view("means3D")[rep]
            new_scales = torch.log(scales / (0.8 * n_split))
            self.densify_clone(rep, overrides={"means3D": new_xyz, "scales": new_scales})
            kf = torch.cat([kf, kf[rep]])
        keep = torch.ones(self.P, dtype=torch.bool, device=self.device)
        keep[split_idx] = False                                          # the originals of the split Gaussians go
        prune = self.view("opacity").reshape(-1) < min_opacity
        if max_screen_size:
            prune |= (self.max_radii2D > max_screen_size) | (self.view("scales").max(dim=1).values > 0.1 * extent)
        keep &= ~prune
        self.prune(keep)
        self.unique_kfIDs = kf[keep]
        self.grad_norm_accum.zero_(); self.denom.zero_(); self.max_radii2D.zero_()
        return self.P

    # ---- one mapping iteration ----
    def step(self, n_views: int, render_and_grad: Callable[[int], None], grad_flat: torch.Tensor,
             zero: Optional[Callable[[], None]] = None, extra_views: Sequence[int] = ()):
        """render_and_grad(k) must ADD view k's parameter gradients (with respect to the values `view()` returns) into
        `grad_flat`, which must be zero on entry the first time; the exchange step leaves it zeroed for the next call.
        `extra_views`: the reference's 2 random older keyframes (utils/slam_backend.py:275); the caller draws them
        with a generator seeded identically on every rank, they are sharded like the window."""
        if zero is not None:
            zero()
        views = list(range(n_views)) + list(extra_views)
        mine = [views[i] for i in shard_keyframes(len(views), self.world, self.rank)]
        for k in mine:
            render_and_grad(k)
        self.exchange_and_update(grad_flat)
        return mine
