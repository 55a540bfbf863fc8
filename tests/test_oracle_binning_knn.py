"""Independent checks of the oracle's integer stages (it is the checker of every bit-exact GPU test, so its own binning
and kNN are cross-checked here against brute-force numpy / scipy restatements of SURVEY.md App. A.2 and App. B):
tile rects from (mean2D, radius), the per-tile lists as a stable sort of (tile | depth bits) keys in emission order,
the tile ranges, and distCUDA2 against a KD-tree."""
import numpy as np
import pytest

import oracle
from lvdgs import synth


def _fwd(name, N, seed, k=None):
    cam = synth.make_camera(name, k)
    sc = synth.make_scene(N, cam, seed=seed)
    if k is not None:
        sc["means3D"][:, 2] += 2.0
    fwd = oracle.rasterize_forward(sc["means3D"], sc["opacities"], sc["scales"], sc["rotations"], sc["shs"],
                                   viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform,
                                   campos=cam.camera_center, bg=np.zeros(3, np.float32), W=cam.image_width,
                                   H=cam.image_height, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy)
    return cam, fwd


@pytest.mark.parametrize("name,N,seed,k", [("mast3r_kitti", 6000, 1, None), ("vga", 9000, 2, 3)])
def test_oracle_binning_is_the_stable_sort_of_the_emitted_keys(name, N, seed, k):
    cam, f = _fwd(name, N, seed, k)
    W, H = cam.image_width, cam.image_height
    gx, gy = (W + 15) // 16, (H + 15) // 16
    radii, m2d = f["radii"].astype(np.int64), f["means2D"].astype(np.float32)
    # getRect (App. A.1): tile rect of the square of half-width `radius` around the pixel centre, clamped to the grid;
    # the int casts truncate toward zero like the C code
    r32 = radii.astype(np.float32)
    x0 = np.clip(((m2d[:, 0] - r32) / np.float32(16)).astype(np.int64), 0, gx)
    y0 = np.clip(((m2d[:, 1] - r32) / np.float32(16)).astype(np.int64), 0, gy)
    x1 = np.clip(((m2d[:, 0] + r32 + np.float32(15)) / np.float32(16)).astype(np.int64), 0, gx)
    y1 = np.clip(((m2d[:, 1] + r32 + np.float32(15)) / np.float32(16)).astype(np.int64), 0, gy)
    touched = np.where(radii > 0, (x1 - x0) * (y1 - y0), 0)
    np.testing.assert_array_equal(touched, f["tiles_touched"])
    vis = np.nonzero(touched > 0)[0]
    np.testing.assert_array_equal(np.stack([x0, y0, x1, y1], 1)[vis], f["rect"][vis])
    # duplicateWithKeys (A.2): Gaussian-major, then y, then x; stable sort by key
    keys, vals = [], []
    dbits = f["depths"].view(np.uint32).astype(np.uint64)
    for i in vis:
        for y in range(y0[i], y1[i]):
            for x in range(x0[i], x1[i]):
                keys.append((np.uint64(y * gx + x) << np.uint64(32)) | dbits[i]); vals.append(i)
    keys, vals = np.array(keys, np.uint64), np.array(vals, np.uint32)
    order = np.argsort(keys, kind="stable")
    assert f["R"] == keys.size
    np.testing.assert_array_equal(keys[order], f["keys_sorted"])
    np.testing.assert_array_equal(vals[order], f["point_list"])
    # identifyTileRanges (A.2)
    tiles = (keys[order] >> np.uint64(32)).astype(np.int64)
    ranges = np.zeros((gx * gy, 2), np.int64)
    for t in np.unique(tiles):
        idx = np.nonzero(tiles == t)[0]
        ranges[t] = (idx[0], idx[-1] + 1)
    np.testing.assert_array_equal(ranges, f["ranges"].astype(np.int64))


def test_oracle_dist2_matches_a_kdtree():
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(0)
    pts = np.concatenate([rng.normal(0, 1, (3000, 3)), rng.normal(5, 0.01, (200, 3)), np.zeros((3, 3))]).astype(np.float32)
    d, _ = cKDTree(pts.astype(np.float64)).query(pts.astype(np.float64), k=4)          # self + 3 neighbours
    want = (d[:, 1:] ** 2).mean(1)
    got = oracle.dist2(pts)
    np.testing.assert_allclose(got, want, rtol=2e-4, atol=1e-10)
