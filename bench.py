#!/usr/bin/env python
"""bench.py -- fwd+bwd render throughput of the B200 rasterizer on BASELINE.json's headline configuration.

Workload (`config.workload` = "kitti_window8_500k"): one STEP = forward + backward of the 8 keyframe views of a
mapping window (1241x376 each, SURVEY.md section 8d cameras k=0..7) against a replicated map of 500 000 synthetic
Gaussians, parameter gradients summed over the views (the mapping loss is a sum, utils/slam_backend.py:266,300),
followed by the Adam update of the map (utils/slam_backend.py:378-380).
metric = fwd+bwd render Mpix/s = 8*H*W / step time.  At N GPUs the 8 views are sharded over the ranks
(8/N each) and the [P,14] parameter-gradient block is SUM-allreduced over NCCL at the end of the step
(strong scaling: the window is fixed).

  value  : device-resident inputs, lean C-ABI engine (lvdgs.engine.RasterEngine), CUDA-event timed
  e2e    : the reference-facing plugin surface (diff_gaussian_rasterization.GaussianRasterizer + torch loss),
           per view H2D of the target image/depth and camera from pinned host memory, D2H of loss + pose gradient
  roofline / cpu_baseline : see DESIGN.md section 7
  --impl reference : the oracle (CPU restatement of the reference's algorithm; the reference's CUDA sources are
           absent) timed on the host cores for the same metric, one view per step
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "lvd_gs-slam_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np

# keep stdout to the single JSON line the driver parses: NCCL prints its version banner there at VERSION level
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")       # whatever NCCL logs (its banner included) goes to stderr

METRIC = "fwd+bwd render Mpix/s at 1241x376, 500k Gaussians"
CPU_BASELINE_DETAIL = ("C/OpenMP restatement of the reference's algorithm (oracle/raster_oracle.c), not the pure-torch "
                       "transcription BASELINE.md mentions -- a stronger baseline; the reference's own CUDA sources are absent")
WORKLOADS = {"kitti_window8_500k": dict(N=500_000, cam="kitti", views=8),
             "kitti_window8_1m": dict(N=1_000_000, cam="kitti", views=8)}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="kitti_window8_500k", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_view(sc, cam, gc, gd):
    """One view of the workload through the CPU oracle (forward + backward).  Returns seconds."""
    import oracle
    t0 = time.perf_counter()
    fwd = oracle.rasterize_forward(sc["means3D"], sc["opacities"], sc["scales"], sc["rotations"], sc["shs"],
                                   viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform,
                                   campos=cam.camera_center, bg=np.zeros(3, np.float32), W=cam.image_width,
                                   H=cam.image_height, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, want_margin=False)
    oracle.rasterize_backward(fwd, gc, gd, projmatrix_raw=cam.projection_matrix)
    return time.perf_counter() - t0


def bench_config(args, wl, W, H, world, views_per_rank):
    """`config` of the JSON line -- the SAME keys and values in both arms (the driver compares them)."""
    return {"workload": args.workload, "gaussians": wl["N"], "image": [W, H], "views_per_step": wl["views"], "sh_degree": 0,
            "l2": "inputs larger than L2: each view streams its own sorted instance list, and the 8 views of a step "
                  "rewrite >300 MB of binning state between revisits"}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, wl):
    """Reference arm: the reference's algorithm on the host cores (oracle port; the reference's own CUDA build is
    impossible here -- its sources are absent), same metric/config; each step is a bounded sample of the workload: one
    view of the 8-view window (fwd+bwd, full resolution), Mpix/s normalises."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    from lvdgs import synth
    oracle.build()
    cores = oracle.set_num_threads(host_cores())            # torchrun exports OMP_NUM_THREADS=1: ask for the cores explicitly
    cams = [synth.make_camera(wl["cam"], k) for k in range(wl["views"])]
    sc = synth.make_scene(wl["N"], cams[0], seed=0)
    gc, gd = synth.make_upstream_grads(cams[0])
    H, W = cams[0].image_height, cams[0].image_width
    for i in range(args.warmup):
        cpu_oracle_view(sc, cams[i % len(cams)], gc, gd)
    t = 0.0
    for i in range(args.steps):
        t += cpu_oracle_view(sc, cams[i % len(cams)], gc, gd)
    ms = 1e3 * t / max(args.steps, 1)
    val = H * W / (ms * 1e-3) / 1e6
    sample = "one keyframe view (fwd+bwd, full 1241x376) of the 8-view window per step, OpenMP over tiles / Gaussians"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Mpix/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args, wl, W, H, 1, 1),
            "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": cores, "kind": "port", "sample": sample,
                             "detail": CPU_BASELINE_DETAIL},
            "e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def measure_fp32_peak(L, dev, n_sm):
    """Measured FP32 FMA-pipe peak: one launch of n_sm x 8 blocks x 256 threads x 4096 x 16 independent FMAs, scalar and
    packed, best of 5, CUDA events on the launching stream."""
    import ctypes as C
    import torch
    out = torch.zeros(4, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    res = {}
    for packed, key in ((0, "ffma_tflops"), (1, "ffma2_tflops")):
        best = 0.0
        for _ in range(6):
            fmas = C.c_double(0.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.lvdgs_fp32_peak(n_sm * 8, 4096, packed, C.c_void_p(out.data_ptr()), C.byref(fmas), stream)
            e1.record(); e1.synchronize()
            best = max(best, 2.0 * fmas.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        res[key] = best
    return res


def main():
    args = parse()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl)

    import torch
    import torch.distributed as dist
    from lvdgs import synth, _native
    from lvdgs.engine import RasterEngine, ViewCamera
    from lvdgs.mapping import ShardedMapper

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert wl["views"] % world == 0, "the 8-view window must divide over the ranks"
    L = _native.lib()

    cams = [synth.make_camera(wl["cam"], k) for k in range(wl["views"])]
    my_views = list(range(rank, wl["views"], world))          # keyframe k -> rank k mod world
    sc = synth.make_scene(wl["N"], cams[0], seed=0)            # identical replicated map on every rank
    H, W = cams[0].image_height, cams[0].image_width
    t = lambda a: torch.tensor(a, dtype=torch.float32, device=dev).contiguous()
    # the replicated map lives in ONE flat block of RAW parameters (lvdgs.mapping.ShardedMapper: logit opacity, log scale,
    # un-normalised rotations, like the reference's GaussianModel) updated by the fused Adam kernel.  The learning rates
    # are scaled down 1e-4 only because the upstream gradients of this workload are fixed random images: at the real rates
    # 25 steps of Adam along them would random-walk the map and the timed steps would not see the same workload
    # (tests/test_gpu_slam_ops.py::test_mapper_optimises_raw_parameters_like_gaussian_model runs the real rates)
    base_lr = {"means3D": 1.6e-4, "shs": 2.5e-3, "opacity": 5e-2, "scales": 1e-3, "rotations": 1e-3}
    gc_np, gd_np = synth.make_upstream_grads(cams[0])
    gc, gd = t(gc_np), t(gd_np)
    vcs = {k: ViewCamera(cams[k], dev) for k in my_views}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def build_job(scene, n_gauss):
        """Mapper (replicated raw parameter block + sharded Adam) and engine (persistent arenas, two-stream view pipeline)."""
        mp_ = ShardedMapper(n_gauss, sh_coeffs=1, device=dev, lrs={k: v * 1e-4 for k, v in base_lr.items()})
        mp_.load(means3D=scene["means3D"], shs=scene["shs"], opacity=scene["opacities"], scales=scene["scales"], rotations=scene["rotations"])
        eng_ = RasterEngine(n_gauss, W, H, sh_coeffs=1, sh_degree=0, device=dev, slots=int(os.environ.get("LVDGS_SLOTS", "2")),
                            grad_flat=mp_.new_grad_block())
        return mp_, eng_

    fixed_upstream = lambda k, slot: (gc, gd, None)           # fixed synthetic dL/dcolor, dL/ddepth
    my_vcs = [vcs[k] for k in my_views]

    def make_step(mp_, eng_):
        def step():
            args_ = [mp_.view(k) for k in ("means3D", "opacity", "scales", "rotations", "shs")]
            # forward of view k+1 overlaps the backward of view k (two streams, two buffer slots); gradients accumulate
            eng_.run_views(my_vcs, *args_, fixed_upstream, bwd_wait=mp_.grad_ready)
            # SUM over the keyframe shards, chain rule to the raw parameters, Adam on this rank's slice, new parameters to
            # every rank, activations (one kernel over NVSwitch multicast / peer memory, or the NCCL sequence); the
            # gradient block is cleared on a side stream, the next step's first backward waits for that (grad_ready)
            mp_.exchange_and_update(eng_.grad_flat, defer_zero=True)
        return step

    def time_steps(step, steps, warmup):
        for _ in range(warmup):
            step()
        barrier()
        e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0_.record()
        for _ in range(steps):
            step()
        e1_.record()
        barrier()
        mine = e0_.elapsed_time(e1_) / steps
        tms_ = torch.tensor([mine], device=dev)
        per_rank = [mine]
        if world > 1:
            allr = [torch.zeros(1, device=dev) for _ in range(world)]
            dist.all_gather(allr, tms_)
            per_rank = [float(x.item()) for x in allr]
            dist.all_reduce(tms_, op=dist.ReduceOp.MAX)
        return float(tms_.item()), per_rank

    # ---------------- value: device-resident, C-ABI engine ----------------
    mapper, eng = build_job(sc, P := wl["N"])
    means3D, opac, scales, rots, shs = (mapper.view(k) for k in ("means3D", "opacity", "scales", "rotations", "shs"))
    step_resident = make_step(mapper, eng)
    for _ in range(args.warmup + 3):      # W warm-up steps + 3 more: arena growth, capacity hints and the first NCCL / symmetric-memory
        step_resident()                   # calls of the process settle before the timed region (all untimed)
    barrier()
    if world > 1:                         # the collective used by the timing barrier itself is warm too
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    L.lvdgs_reset_launch_count()
    ms, ms_per_rank = time_steps(step_resident, args.steps, 0)
    launches = int(L.lvdgs_launch_count())
    value = wl["views"] * H * W / (ms * 1e-3) / 1e6
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    # SURVEY 8(d)'s per-view figure beside the pipelined window: H*W / (t_fwd + t_bwd) of ONE view at a time on one stream
    per_view = None
    if rank == 0:
        ts_ = []
        for k in my_views[:4]:
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                eng.forward(vcs[k], means3D, opac, scales, rots, shs)
                eng.backward(vcs[k], means3D, opac, scales, rots, shs, gc, gd, accumulate=True)
            e1.record(); e1.synchronize()
            ts_.append(e0.elapsed_time(e1) / 5)
        eng.zero_grads()
        per_view = {"ms_fwd_bwd": float(np.median(ts_)), "mpix_per_s": H * W / (float(np.median(ts_)) * 1e-3) / 1e6,
                    "what": "one view at a time on one stream (no overlap between views), median over this rank's first views"}

    # where a step's time goes on THIS rank count: the rendering of the rank's views alone and the exchange step alone
    # (raw-parameter chain rule, NCCL reduce-scatter, Adam on the slice, all-gather, activations), each as the max over ranks
    def views_only():
        eng.run_views(my_vcs, *[mapper.view(k) for k in ("means3D", "opacity", "scales", "rotations", "shs")], fixed_upstream)
    ms_views, ms_views_ranks = time_steps(views_only, max(5, args.steps // 2), 2)
    ms_exch, _ = time_steps(lambda: mapper.exchange_and_update(eng.grad_flat), max(5, args.steps // 2), 2)
    exch_phases = None
    if mapper._p2p:       # device time of the phases of the peer-memory exchange on this rank (CUDA events around each)
        mapper.time_exchange = True
        time_steps(lambda: mapper.exchange_and_update(eng.grad_flat), max(5, args.steps // 2), 0)
        mapper.time_exchange = False
        exch_phases = mapper.exchange_phase_ms()
    step_split = {"views_ms": ms_views, "views_ms_by_rank": ms_views_ranks, "exchange_ms": ms_exch, "exchange_phases_ms_rank0": exch_phases,
                  "what": "timed apart, max over ranks: this rank's views through RasterEngine.run_views; "
                          "ShardedMapper.exchange_and_update (gradient sum over the ranks, activation chain rule, Adam on the shard, "
                          "parameters to every rank, activations; exchange_phases_ms_rank0 = its phases on rank 0 from CUDA events)"}

    # BASELINE configs[2] names 1 M Gaussians for the mapping window: the same step on that map (extra key; the headline
    # metric is quoted at 500 k)
    mapping_1m = None
    if args.workload == "kitti_window8_500k" and os.environ.get("LVDGS_BENCH_1M", "1") != "0":
        sc1 = synth.make_scene(1_000_000, cams[0], seed=0)
        m1, e1m = build_job(sc1, 1_000_000)
        ms1, ms1_ranks = time_steps(make_step(m1, e1m), max(5, args.steps // 2), 3)
        mapping_1m = {"workload": "kitti_window8_1m (BASELINE configs[2])", "ms_per_step": ms1, "iters_per_s": 1e3 / ms1,
                      "mpix_per_s": wl["views"] * H * W / (ms1 * 1e-3) / 1e6, "ms_per_step_by_rank": ms1_ranks}
        del m1, e1m, sc1
        torch.cuda.empty_cache()

    # ---------------- e2e: plugin surface, host buffers ----------------
    import diff_gaussian_rasterization as dgr
    from lvdgs import slam_ops
    params = [x.detach().clone().requires_grad_() for x in (means3D, opac, scales, rots, shs)]
    e2e_opt = torch.optim.Adam(params, lr=1e-8)
    rng = np.random.default_rng(2)
    host = {}
    for k in my_views:
        host[k] = dict(img=torch.from_numpy(rng.uniform(0, 1, (3, H, W)).astype(np.float32)).pin_memory(),
                       dep=torch.from_numpy(rng.uniform(1, 50, (1, H, W)).astype(np.float32)).pin_memory(),
                       cam=torch.from_numpy(np.concatenate([cams[k].world_view_transform.ravel(), cams[k].full_proj_transform.ravel(),
                                                            cams[k].projection_matrix.ravel(), cams[k].camera_center.ravel(),
                                                            np.zeros(1, np.float32)]).astype(np.float32)).pin_memory())
    res_host = torch.zeros(len(my_views), 8).pin_memory()
    bgt = torch.zeros(3, device=dev)
    h2d = sum(v["img"].numel() + v["dep"].numel() + v["cam"].numel() for v in host.values()) * 4
    d2h = res_host.numel() * 4

    copy_stream = torch.cuda.Stream(dev)
    # one staging slot per view of the rank (7.5 MB each): a step queues every upload on the copy stream up front -- the
    # cameras first (a render needs only its camera), then the target image / depth of each view (needed by its loss)
    stage = [dict(img=torch.empty(3, H, W, device=dev), dep=torch.empty(1, H, W, device=dev), cam=torch.empty(52, device=dev),
                  cam_ready=torch.cuda.Event(), ready=torch.cuda.Event(), free=torch.cuda.Event()) for _ in my_views]

    def prefetch_all():
        """H2D of every view's camera, then target image / depth, on the copy stream (slot j = view j of this rank)."""
        with torch.cuda.stream(copy_stream):
            for j, k in enumerate(my_views):
                st = stage[j]
                copy_stream.wait_event(st["free"])              # the previous step has consumed the slot
                st["cam"].copy_(host[k]["cam"], non_blocking=True)
                st["cam_ready"].record(copy_stream)
            for j, k in enumerate(my_views):
                st = stage[j]
                st["img"].copy_(host[k]["img"], non_blocking=True); st["dep"].copy_(host[k]["dep"], non_blocking=True)
                st["ready"].record(copy_stream)

    torch_loss = os.environ.get("LVDGS_E2E_TORCH_LOSS", "0") == "1"

    def step_e2e():
        """One mapping iteration the way utils/slam_backend.py:167-306 runs it: render every view of the window, sum the
        per-view losses (`loss_mapping +=`), ONE backward through all of them, one optimiser step."""
        cur = torch.cuda.current_stream()
        e2e_opt.zero_grad(set_to_none=True)
        prefetch_all()
        total, poses = None, []
        for j, k in enumerate(my_views):
            st = stage[j]
            cur.wait_event(st["cam_ready"])
            img, dep, cmv = st["img"], st["dep"], st["cam"]     # the slot is not rewritten before the step's backward has run
            rs = dgr.GaussianRasterizationSettings(
                image_height=H, image_width=W, tanfovx=cams[k].tanfovx, tanfovy=cams[k].tanfovy, bg=bgt, scale_modifier=1.0,
                viewmatrix=cmv[0:16].view(4, 4), projmatrix=cmv[16:32].view(4, 4), projmatrix_raw=cmv[32:48].view(4, 4),
                sh_degree=0, campos=cmv[48:51], prefiltered=False, debug=False)
            theta = torch.zeros(3, device=dev, requires_grad=True); rho = torch.zeros(3, device=dev, requires_grad=True)
            m2d = torch.zeros_like(params[0], requires_grad=True)
            color, radii, depth, opacity, n_touched = dgr.GaussianRasterizer(rs)(
                means3D=params[0], means2D=m2d, opacities=params[1], shs=params[4], scales=params[2], rotations=params[3],
                theta=theta, rho=rho)
            # mapping loss of utils/slam_utils.py:107-121 (0.9 L1 rgb + 0.1 L1 depth on the valid pixels) through the
            # package's fused loss op (lvdgs.slam_ops); LVDGS_E2E_TORCH_LOSS=1 uses the torch expression
            cur.wait_event(st["ready"])                         # the loss is the first consumer of the targets
            if torch_loss:
                loss = 0.9 * (color - img).abs().mean() + 0.1 * (depth - dep).abs().mean()
            else:
                loss = slam_ops.fused_loss(color, depth, gt_image=img, gt_depth=dep, rgb_boundary_threshold=-1.0,
                                           w_rgb=0.9, w_depth=0.1)
            total = loss if total is None else total + loss
            poses.append((loss, rho, theta, st))
        total.backward()
        for j, (loss, rho, theta, st) in enumerate(poses):
            res_host[j, 0:1].copy_(loss.detach().reshape(1), non_blocking=True)
            res_host[j, 1:4].copy_(rho.grad, non_blocking=True)
            res_host[j, 4:7].copy_(theta.grad, non_blocking=True)
        if world > 1:
            for p_ in params:
                dist.all_reduce(p_.grad)
        e2e_opt.step()
        for st in stage:
            st["free"].record(cur)                              # cameras (read by the backward) and targets consumed
        torch.cuda.current_stream().synchronize()               # the step's result is on the host

    for _ in range(max(3, args.warmup)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1) / args.steps
    tms = torch.tensor([e2e_ms], device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    e2e_ms = float(tms.item())
    e2e_val = wl["views"] * H * W / (e2e_ms * 1e-3) / 1e6
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- e2e over the C ABI engine: same host buffers, same copies, views pipelined ----------------
    # The plugin surface above has to follow the reference's single-stream program order (all renders, one backward).  A
    # caller written against the C ABI (lvdgs.engine.RasterEngine) can overlap the views; this is that path with the SAME
    # per-step host traffic: H2D of every view's target image / depth / camera from pinned memory, the fused mapping loss
    # (lvdgs_fused_loss) between forward and backward, D2H of every view's loss and pose gradient.
    from lvdgs import slam_ops as _so
    # one staging slot per view of the rank (7.5 MB each): every upload of a step is queued on the copy stream up front and
    # only waits for the previous step's use of its own slot
    NS = len(my_views)
    st_c = [dict(img=torch.empty(3, H, W, device=dev), dep=torch.empty(1, H, W, device=dev), cam=torch.empty(52, device=dev),
                 ready=torch.cuda.Event(), free=torch.cuda.Event(), g_img=torch.empty(3, H, W, device=dev),
                 g_dep=torch.empty(1, H, W, device=dev), out=torch.empty(4, device=dev)) for _ in range(NS)]
    res_c = torch.zeros(len(my_views), 8).pin_memory()

    class _StagedCamera:
        """ViewCamera over a staging slot: the five device arrays are views into the 52 floats copied from the host."""
        def __init__(self, cam, st):
            self.W, self.H, self.tanfovx, self.tanfovy = cam.image_width, cam.image_height, cam.tanfovx, cam.tanfovy
            c_ = st["cam"]
            self.view, self.proj, self.proj_raw, self.campos, self.bg = c_[0:16], c_[16:32], c_[32:48], c_[48:51], bgt

    staged = [[_StagedCamera(cams[k], st_c[j]) for j, k in enumerate(my_views)]]

    def prefetch_c(j):
        st, hb = st_c[j], host[my_views[j]]
        copy_stream.wait_event(st["free"])
        with torch.cuda.stream(copy_stream):
            st["img"].copy_(hb["img"], non_blocking=True); st["dep"].copy_(hb["dep"], non_blocking=True)
            st["cam"].copy_(hb["cam"], non_blocking=True)
            st["ready"].record(copy_stream)

    def upstream_c(j, slot):
        """Runs on the backward stream after view j's forward: loss + upstream gradients from the staged targets."""
        st = st_c[j]
        _so.fused_loss_into(slot.color, slot.depth, st["img"], st["dep"], st["g_img"], st["g_dep"], st["out"],
                            rgb_boundary_threshold=-1.0, w_rgb=0.9, w_depth=0.1)
        res_c[j, 0:1].copy_(st["out"][0:1], non_blocking=True)
        return st["g_img"], st["g_dep"], None

    def after_view_c(j, slot):
        res_c[j, 1:7].copy_(slot.g_tau, non_blocking=True)
        st_c[j]["free"].record(torch.cuda.current_stream())           # targets + camera of this slot consumed (fwd and bwd)

    def step_c_abi():
        args_ = [mapper.view(k) for k in ("means3D", "opacity", "scales", "rotations", "shs")]
        for j in range(NS):
            prefetch_c(j)
        # the forward stream waits for view j's camera right before the view is launched; the backward stream (loss) follows it
        fwd_stream = eng.view_streams(NS)[0]
        eng.run_views(staged[0], *args_, upstream_c, on_view=after_view_c,
                      before_view=lambda j: fwd_stream.wait_event(st_c[j]["ready"]))
        mapper.exchange_and_update(eng.grad_flat)
        torch.cuda.current_stream().synchronize()                      # the step's results are on the host

    for _ in range(max(3, args.warmup)):
        step_c_abi()
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_c_abi()
    e1.record()
    barrier()
    c_ms = e0.elapsed_time(e1) / args.steps
    tms = torch.tensor([c_ms], device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    c_ms = float(tms.item())
    e2e_c_abi = {"value": wl["views"] * H * W / (c_ms * 1e-3) / 1e6, "unit": "Mpix/s", "ms_per_step": c_ms,
                 "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": res_c.numel() * 4,
                 "path": "lvdgs.engine.RasterEngine.run_views over the C ABI (forward of view k+1 overlaps backward of view k) + "
                         "lvdgs_fused_loss between them + ShardedMapper.exchange_and_update; per view H2D of target image / depth / "
                         "camera from pinned memory on a copy stream, D2H of loss and pose gradient"}

    # ---------------- per-kernel device times + roofline (rank 0, outside the timed regions) ----------------
    roof, kernels = None, []
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        sm_max = peaks.get("sm_max_mhz", 1965.0)
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        fp32_computed = n_sm * 128 * 2 * sm_max * 1e6 / 1e12    # TFLOP/s at max clock
        fp32_meas = measure_fp32_peak(L, dev, n_sm)             # FFMA / FFMA2 micro-kernel, CUDA events, best of 5
        fp32_peak = max(fp32_meas["ffma_tflops"], fp32_meas["ffma2_tflops"])
        stream = torch.cuda.current_stream().cuda_stream
        agg, pairs, Rs, vis = {}, 0, 0, 0
        reps = 3
        for _ in range(reps):
            for k in my_views:
                _native.profile_begin(stream)
                eng.forward(vcs[k], means3D, opac, scales, rots, shs)
                eng.backward(vcs[k], means3D, opac, scales, rots, shs, gc, gd, accumulate=True)
                for name, t_ms in _native.profile_end(stream):
                    agg.setdefault(name, []).append(t_ms)
                pairs += eng.pair_count(); Rs += eng.R; vis += int((eng.radii > 0).sum().item())
        nview = reps * len(my_views)
        pairs /= nview; Rs /= nview; vis /= nview
        passes = 6
        alg = {  # algorithmic work per view (SURVEY.md section 8d / DESIGN.md section 6)
            "preprocess_forward": ("hbm", 20.0 * (P - vis) + 131.0 * vis),
            "binning_count": ("hbm", 4.0 * P + 12.0 * vis),
            "emit_keys": ("hbm", 8.0 * Rs + 20.0 * P),            # one 8-byte (depth | Gaussian) word per instance
            "tile_sort": ("hbm", 20.0 * Rs),                      # 8 B read + 12 B (key, value) written per instance;
                                                                  # the network itself runs in shared memory (ALU-bound)
            "sort_onesweep": ("hbm", passes * 24.0 * Rs),         # only with LVDGS_FLAG_GLOBAL_SORT
            "blend_forward": ("fp32", 28.0 * pairs),
            "blend_backward": ("fp32", 70.0 * pairs),
            "preprocess_backward": ("hbm", 190.0 * vis + 52.0 * (P - vis)),
            "adam_step": ("hbm", 28.0 * 14 * P / max(len(my_views), 1)),
        }
        for name, v in agg.items():
            per_view_ms = sum(v) / nview
            ent = {"kernel": name, "launches_per_view": len(v) / nview, "ms_per_view": per_view_ms}
            if name in alg and per_view_ms > 0:
                bound, work = alg[name]
                if bound == "hbm":
                    ach = work / (per_view_ms * 1e-3) / 1e9
                    ent.update(bound="hbm", achieved=ach, peak=hbm_peak, unit="GB/s", frac=ach / hbm_peak)
                else:
                    ach = work / (per_view_ms * 1e-3) / 1e12
                    ent.update(bound="fp32", achieved=ach, peak=fp32_peak, unit="TFLOP/s", frac=ach / fp32_peak)
            kernels.append(ent)
        kernels.sort(key=lambda e: -e["ms_per_view"])
        dom = next((e for e in kernels if "bound" in e), None)
        traffic, ncu = None, None
        try:   # one `ncu --set full` capture per kernel, committed under profiles/ (per launch, like `achieved`)
            ncu_all = json.load(open(os.path.join(ROOT, "profiles", "ncu_kernel_metrics.json")))
            for e in kernels:      # DRAM bytes of the captured launch (view 0) beside the algorithmic bytes of the mean view
                if e["kernel"] in ncu_all and e.get("bound") == "hbm":
                    e["ncu_dram_bytes_view0"] = ncu_all[e["kernel"]]["dram_bytes"]
                    e["algorithmic_bytes_per_view"] = alg[e["kernel"]][1]
            ncu = ncu_all.get(dom["kernel"])
            traffic = ncu["dram_bytes"]
        except Exception:
            pass
        if dom:
            roof = {"kernel": dom["kernel"], "bound": dom["bound"], "achieved": dom["achieved"], "peak": dom["peak"],
                    "unit": dom["unit"], "frac": dom["frac"], "traffic": traffic,
                    "ms_per_launch": dom["ms_per_view"] / dom["launches_per_view"],
                    "peak_source": hbm_src if dom["bound"] == "hbm" else
                    f"measured in this run: lvdgs_fp32_peak micro-kernel, FFMA {fp32_meas['ffma_tflops']:.1f} / FFMA2 {fp32_meas['ffma2_tflops']:.1f} "
                    f"TFLOP/s (computed {n_sm} SMs x 128 lanes x 2 x {sm_max:.0f} MHz = {fp32_computed:.1f}; MEASURED_PEAKS.json has no FP32 "
                    "figure; tensor cores unused: the blend is FP32 FMA/MUFU issue-bound, not a dense contraction)",
                    "fp32_peak_measured": fp32_meas,
                    "algorithmic": {"pairs_per_view": pairs, "instances_per_view": Rs, "visible_per_view": vis},
                    "ncu": ncu,
                    "note": "frac counts only the FLOPs of the A.3/A.4 blend math; the kernel's issue slots (ncu "
                            "smsp__issue_active, in `ncu`) are the pipe-utilisation figure north_star asks for"}

    # ---------------- cpu baseline (rank 0, N=1): oracle port on the host cores, one view ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle
        oracle.build()
        cpu_oracle_view(sc, cams[0], gc_np, gd_np)               # warm (page-in, OpenMP pool)
        ts = [cpu_oracle_view(sc, cams[k], gc_np, gd_np) for k in (0, 3, 6)]
        tv = float(np.median(ts))
        cpu = {"value": H * W / tv / 1e6, "unit": "Mpix/s", "cores": oracle.set_num_threads(host_cores()), "kind": "port", "detail": CPU_BASELINE_DETAIL,
               "sample": "3 of the 8 views (k=0,3,6), full 1241x376 fwd+bwd each, median; C oracle with OpenMP",
               "seconds_per_view": tv}

    # ---------------- BASELINE configs[1]: the tracking loop, device-resident (rank 0, N=1; extra key, not the metric) ----------------
    tracking = None
    if rank == 0 and world == 1:
        try:
            from lvdgs.tracking import PoseTracker
            NT, iters_t = 300_000, 100
            sct = synth.make_scene(NT, cams[0], seed=0)
            tt = lambda a: torch.tensor(a, dtype=torch.float32, device=dev).contiguous()
            gm, go, gs, gr, gsh = tt(sct["means3D"]), tt(sct["opacities"]), tt(sct["scales"]), tt(sct["rotations"]), tt(sct["shs"])
            eng_t = RasterEngine(NT, W, H, sh_coeffs=1, sh_degree=0, device=dev, slots=1)
            eng_t.forward(ViewCamera(cams[0], dev), gm, go, gs, gr, gsh)
            target = eng_t.color.clone()
            del eng_t
            trk_ = PoseTracker(NT, W, H, cams[0].tanfovx, cams[0].tanfovy, device=dev, lr_rot=0.003, lr_trans=0.001,
                               rgb_boundary_threshold=-1.0)
            R1, T1 = cams[1].R, cams[1].T                       # start from the neighbouring keyframe's pose
            ts_ = []
            for rep in range(3):
                trk_.set_camera(R1, T1, cams[0].projection_matrix)
                torch.cuda.synchronize(); t0_ = time.perf_counter()
                out_t = trk_.track(gm, go, gs, gr, gsh, target, iters=iters_t, stop_when_converged=False)
                torch.cuda.synchronize(); ts_.append(time.perf_counter() - t0_)
            tsec = float(np.median(ts_[1:]))
            tracking = {"workload": "kitti_tracking_300k (BASELINE configs[1])", "iters_per_frame": iters_t,
                        "ms_per_iter": 1e3 * tsec / iters_t, "iters_per_s": iters_t / tsec,
                        "mpix_per_s": iters_t * H * W / tsec / 1e6,
                        "path": "lvdgs.tracking.PoseTracker: rasterizer fwd + fused tracking loss + pose-only bwd + lvdgs_pose_step, host wall clock"}
        except Exception as e:                                   # never let the extra measurement break the metric line
            tracking = {"error": repr(e)[:200]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": bench_config(args, wl, W, H, world, len(my_views)),
                "layout": {"views_per_rank": len(my_views),
                           "parallelism": (f"keyframes sharded over {world} rank(s); exchange = " +
                                           (("ONE kernel over NVSwitch multicast memory (lvdgs_exchange_adam: in-switch multimem.ld_reduce of the gradient slice, chain rule, Adam on the shard, multimem.st of parameters + activations to every rank)"
                                             if mapper.exchange_mode == "multicast" else
                                             "ONE kernel over NVLink peer memory (lvdgs_exchange_adam: peer loads of the gradient slice, chain rule, Adam on the shard, peer stores of parameters + activations)")
                                            if mapper._p2p else "NCCL reduce-scatter of the [P,14] gradients + Adam on the shard + all-gather")) if world > 1 else "1 GPU"},
                "e2e": {"value": e2e_val, "unit": "Mpix/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "path": "diff_gaussian_rasterization.GaussianRasterizer (autograd) + "
                                + ("torch L1 loss" if torch_loss else "lvdgs.slam_ops.fused_loss (mapping rgbd loss)") + ", losses summed over the window and one backward (as utils/slam_backend.py:167-306) + torch Adam; "
                                "every view's H2D (camera, then target image / depth) queued on a copy stream at the start of the step"},
                "e2e_c_abi": e2e_c_abi,
                "gpu_launches": launches, "mapping_iters_per_s": 1e3 / ms, "ms_per_step_by_rank": ms_per_rank,
                "per_view_unpipelined": per_view, "step_split": step_split, "mapping_1m": mapping_1m, "clocks": clocks, "roofline": roof,
                "kernels": kernels, "cpu_baseline": cpu, "tracking": tracking}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
