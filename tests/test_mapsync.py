"""Versioned shared map snapshot (lvdgs.mapsync) replacing clone_obj + mp.Queue pickling of the whole map
(utils/slam_backend.py:470-480, utils/multiprocessing_utils.py:21-31, utils/slam_frontend.py:1690-1697): slot protocol and
cross-process visibility on host shared memory (CPU suite) and on one CUDA allocation opened over CUDA IPC (GPU)."""
import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from lvdgs.mapsync import MapPublisher, MapSubscriber, N_SLOTS


def _arrays(P, M, seed, device):
    g = torch.Generator().manual_seed(seed)
    mk = lambda *s: torch.randn(*s, generator=g).to(device)
    return {"means3D": mk(P, 3), "shs": mk(P, M, 3), "opacity": torch.rand(P, 1, generator=g).to(device),
            "scales": torch.rand(P, 3, generator=g).to(device), "rotations": mk(P, 4)}


def _check(view, arrays):
    for attr, name in (("get_xyz", "means3D"), ("get_features", "shs"), ("get_opacity", "opacity"), ("get_scaling", "scales"),
                       ("get_rotation", "rotations")):
        assert torch.equal(getattr(view, attr).cpu(), arrays[name].cpu()), name


def test_slot_protocol_never_overwrites_what_a_reader_holds():
    pub = MapPublisher(capacity=1000, sh_coeffs=1, device="cpu")
    sub = MapSubscriber(pub.handle())
    assert sub.acquire() is None                                      # nothing published yet
    a1 = _arrays(700, 1, 1, "cpu")
    assert pub.publish(a1) == 1
    v1 = sub.acquire()
    assert v1.version == 1 and v1.P == 700
    _check(v1, a1)
    seen = {v1.slot}
    for k in range(2, 9):                                             # the reader keeps holding v1's slot all along
        ak = _arrays(300 + 50 * k, 1, k, "cpu")
        assert pub.publish(ak) == k
        _check(v1, a1)                                                # untouched, zero copies were made for it
        seen.add(int(pub.header[1]))
    assert seen == set(range(N_SLOTS))                                # the other two slots alternate
    sub.release()
    v = sub.acquire()
    assert v.version == 8 and v.P == 700
    _check(v, ak)
    with pytest.raises(ValueError):
        pub.publish(_arrays(1001, 1, 0, "cpu"))
    pub.close()


def _reader(handle, q_in, q_out):
    sub = MapSubscriber(handle)
    while True:
        cmd = q_in.get()
        if cmd == "stop":
            break
        view = sub.acquire()
        out = dict(version=view.version, P=view.P, xyz_sum=float(view.get_xyz.double().sum()),
                   rot_last=view.get_rotation[-1].cpu().numpy().tolist())
        sub.release()
        q_out.put(out)


def _cross_process(device):
    ctx = mp.get_context("spawn")
    pub = MapPublisher(capacity=5000, sh_coeffs=1, device=device)
    q_in, q_out = ctx.Queue(), ctx.Queue()
    p = ctx.Process(target=_reader, args=(pub.handle(), q_in, q_out))     # the handle crosses ONCE, like a queue message
    p.start()
    try:
        for k in range(1, 5):
            a = _arrays(1000 * k, 1, 10 + k, device)
            pub.publish(a)
            if device != "cpu":
                torch.cuda.synchronize()
            q_in.put("read")
            got = q_out.get(timeout=120)
            assert got["version"] == k and got["P"] == 1000 * k
            assert abs(got["xyz_sum"] - float(a["means3D"].double().sum())) < 1e-6 * (1 + abs(got["xyz_sum"]))
            np.testing.assert_array_equal(np.float32(got["rot_last"]), a["rotations"][-1].cpu().numpy())
    finally:
        q_in.put("stop")
        p.join(timeout=60)
        pub.close()
    assert p.exitcode == 0


@pytest.mark.timeout(300)
def test_reader_process_sees_published_versions_host_memory():
    _cross_process("cpu")


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_reader_process_sees_published_versions_cuda_ipc():
    _cross_process("cuda")


@pytest.mark.gpu
def test_published_view_renders_like_the_source_map():
    """The subscriber's views go straight into gaussian_renderer.render (the frontend's tracking render)."""
    import ref_conventions as rc
    from lvdgs import synth
    from gaussian_splatting.gaussian_renderer import render
    c = synth.make_camera("mast3r_kitti")
    sc = synth.make_scene(8000, c, seed=2)
    pc = rc.Gaussians(sc, "cuda")
    pub = MapPublisher(capacity=10_000, device="cuda")
    sub = MapSubscriber(pub.handle())
    pub.publish({"means3D": pc.get_xyz, "shs": pc.get_features, "opacity": pc.get_opacity, "scales": pc.get_scaling,
                 "rotations": pc.get_rotation})
    view = sub.acquire()
    _, _, cu, _ = rc.load()
    cam = rc.make_camera(cu, c, "cuda")
    bg = torch.zeros(3, device="cuda")
    with torch.no_grad():
        a, b = render(cam, pc, rc.Pipe(), bg), render(cam, view, rc.Pipe(), bg)
    assert torch.equal(a["render"], b["render"]) and torch.equal(a["n_touched"], b["n_touched"])
    sub.release()
    pub.close()
