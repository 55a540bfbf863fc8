"""GPU parity for the two stand-alone pieces of the C ABI: the onesweep radix sort (vs numpy stable sort and vs
cub::DeviceRadixSort, the library call the reference makes) and distCUDA2 (vs the brute-force oracle / cKDTree)."""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle
from lvdgs import _native

pytestmark = pytest.mark.gpu


def _sort(keys_np, vals_np, end_bit, which="ours"):
    L = _native.lib()
    n = len(keys_np)
    dev = "cuda"
    k0 = torch.from_numpy(keys_np.view(np.int64)).to(dev)
    v0 = torch.from_numpy(vals_np.view(np.int32)).to(dev)
    k1 = torch.empty_like(k0); v1 = torch.empty_like(v0)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = _native.ptr
    if which == "ours":
        ws = torch.empty(L.lvdgs_sort_workspace_bytes(n), dtype=torch.uint8, device=dev)
        sel = C.c_int32(0)
        rc = L.lvdgs_sort_pairs(n, p(k0), p(k1), p(v0), p(v1), end_bit, p(ws), ws.numel(), C.byref(sel), stream)
        _native.check(rc, "sort")
        torch.cuda.synchronize()
        ks, vs = (k0, v0) if sel.value == 0 else (k1, v1)
    else:
        ws = torch.empty(max(1, L.lvdgs_cub_sort_workspace_bytes(n, end_bit)), dtype=torch.uint8, device=dev)
        rc = L.lvdgs_cub_sort_pairs(n, p(k0), p(k1), p(v0), p(v1), end_bit, p(ws), ws.numel(), stream)
        _native.check(rc, "cub sort")
        torch.cuda.synchronize()
        ks, vs = k1, v1
    return ks.cpu().numpy().view(np.uint64), vs.cpu().numpy().view(np.uint32)


@pytest.mark.parametrize("n,end_bit,dup", [(1, 43, False), (31, 43, True), (4095, 43, False), (4096, 45, True),
                                           (4097, 40, True), (100_003, 43, True), (1_500_000, 43, False),
                                           (300_000, 64, False), (70_000, 9, True), (5_000_000, 45, True)])
def test_sort_matches_stable_reference(n, end_bit, dup):
    rng = np.random.default_rng(n + end_bit)
    tile_bits = max(end_bit - 32, 0)
    tiles = rng.integers(0, 1 << tile_bits, n, dtype=np.uint64) if tile_bits else np.zeros(n, np.uint64)
    if dup:   # many equal keys: stability is observable through the values
        depth = rng.integers(0, 50, n).astype(np.float32) * 0.37 + 0.5
    else:
        depth = np.exp(rng.uniform(np.log(0.2), np.log(100.0), n)).astype(np.float32)
    dbits = depth.view(np.uint32).astype(np.uint64)
    if end_bit < 32:
        dbits &= np.uint64((1 << end_bit) - 1)
    keys = (tiles << np.uint64(32)) | dbits
    if end_bit == 64:
        keys = rng.integers(0, 2 ** 63, n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, n, dtype=np.uint64)
    vals = np.arange(n, dtype=np.uint32)
    order = np.argsort(keys, kind="stable")
    ks, vs = _sort(keys, vals, end_bit, "ours")
    np.testing.assert_array_equal(ks, keys[order])
    np.testing.assert_array_equal(vs, vals[order])
    kc, vc = _sort(keys, vals, end_bit, "cub")
    np.testing.assert_array_equal(kc, ks)
    np.testing.assert_array_equal(vc, vs)


@pytest.mark.parametrize("P", [4, 1000, 2049, 14_600, 60_000])
def test_dist2_matches_oracle(P):
    from simple_knn._C import distCUDA2
    rng = np.random.default_rng(P)
    pts = (rng.normal(0, 1, (P, 3)) * np.array([5.0, 1.0, 20.0])).astype(np.float32)
    if P >= 1000:
        pts[:50] = pts[50:100]          # exact duplicates -> zero distances
    got = distCUDA2(torch.tensor(pts, device="cuda")).cpu().numpy()
    ref = oracle.dist2(pts)
    np.testing.assert_allclose(got, ref, rtol=2e-6, atol=1e-12)
    from scipy.spatial import cKDTree
    d, _ = cKDTree(pts.astype(np.float64)).query(pts.astype(np.float64), k=4)
    np.testing.assert_allclose(got, (d[:, 1:] ** 2).mean(1), rtol=1e-4, atol=1e-9)


def test_dist2_large_surface_like_cloud_matches_kdtree():
    """2 M points on a 2.5-D surface (what back-projected keyframe depth maps look like; BASELINE configs[3] size):
    exactness against cKDTree, the all-pairs oracle being out of reach at this size."""
    from simple_knn._C import distCUDA2
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(0)
    P = 2_000_000
    u, v = rng.uniform(-40, 40, P), rng.uniform(-10, 10, P)
    z = 20 + 5 * np.sin(0.2 * u) + 0.05 * rng.normal(0, 1, P) + np.where(rng.uniform(0, 1, P) < 0.1, 30.0, 0.0)
    pts = np.stack([u, v, z], 1).astype(np.float32)
    got = distCUDA2(torch.tensor(pts, device="cuda")).cpu().numpy()
    sub = rng.choice(P, 50_000, replace=False)
    d, _ = cKDTree(pts.astype(np.float64)).query(pts[sub].astype(np.float64), k=4)
    np.testing.assert_allclose(got[sub], (d[:, 1:] ** 2).mean(1), rtol=2e-4, atol=1e-9)
