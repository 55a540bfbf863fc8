"""Generates tests/golden/raster_small.npz: seeded inputs + the ORACLE's outputs for one small scene.

There are no reference-owned golden vectors for this path (the reference has no tests and its rasterizer sources are
absent), and the reference cannot be imported or compiled here, so these fixtures are produced by
oracle/raster_oracle.c -- whose backward is pinned against float64 autograd in tests/test_oracle_autograd.py.  They
freeze the oracle's behaviour (a regression pin, checked on CPU) and give the GPU tests a second, file-based anchor.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lvd_gs-slam_b200")]
import oracle  # noqa: E402
from lvdgs import synth  # noqa: E402


def build_case():
    cam = synth.Cam(80, 48, 70.0, 66.0, 37.3, 26.1, np.eye(3), np.zeros(3))
    a = np.radians(4.0)
    cam.R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    cam.T = np.array([0.05, -0.02, 0.4])
    sc = synth.make_scene(400, cam, seed=42, sh_degree=1)
    sc["scales"] *= 1.5
    bg = np.array([0.1, 0.3, 0.2], np.float32)
    rng = np.random.default_rng(43)
    gc = rng.normal(0, 1, (3, 48, 80)).astype(np.float32)
    gd = (rng.normal(0, 1, (1, 48, 80)) * 0.2).astype(np.float32)
    return cam, sc, bg, gc, gd


def run_oracle(cam, sc, bg, gc, gd):
    fwd = oracle.rasterize_forward(sc["means3D"], sc["opacities"], sc["scales"], sc["rotations"], sc["shs"],
                                   viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform,
                                   campos=cam.camera_center, bg=bg, W=cam.image_width, H=cam.image_height,
                                   tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, sh_degree=1)
    g = oracle.rasterize_backward(fwd, gc, gd, projmatrix_raw=cam.projection_matrix, flags=0)
    return fwd, g


if __name__ == "__main__":
    cam, sc, bg, gc, gd = build_case()
    fwd, g = run_oracle(cam, sc, bg, gc, gd)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "raster_small.npz")
    np.savez_compressed(
        out, means3D=sc["means3D"], opacities=sc["opacities"], scales=sc["scales"], rotations=sc["rotations"], shs=sc["shs"],
        bg=bg, grad_color=gc, grad_depth=gd, viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform,
        projmatrix_raw=cam.projection_matrix, campos=cam.camera_center, tanfov=np.array([cam.tanfovx, cam.tanfovy]),
        color=fwd["color"], depth=fwd["depth"], opacity=fwd["opacity"], radii=fwd["radii"], n_touched=fwd["n_touched"],
        n_contrib=fwd["n_contrib"], keys_sorted=fwd["keys_sorted"], point_list=fwd["point_list"], ranges=fwd["ranges"],
        margin=fwd["margin"], dL_dmeans3D=g["dL_dmeans3D"], dL_dmean2D=g["dL_dmean2D"], dL_dopacity=g["dL_dopacity"],
        dL_dscales=g["dL_dscales"], dL_drots=g["dL_drots"], dL_dsh=g["dL_dsh"], grad_rho=g["grad_rho"],
        grad_theta=g["grad_theta"])
    print("wrote", out, os.path.getsize(out), "bytes; R =", fwd["R"], "visible =", int((fwd["radii"] > 0).sum()))
