"""The algebra behind blend_backward.cu / preprocess_backward.cu, checked on the CPU against the oracle's blend backward
(which follows SURVEY.md App. A.4 term by term):

 * the colour / depth blended BEHIND a Gaussian enters dL/dalpha only through S = <B, dL/dpixel>, and S obeys
   S <- alpha <c, dL/dpixel> + (1 - alpha) S  (one scalar recurrence instead of four);
 * a pixel that does not blend a Gaussian may run the same arithmetic with alpha = G = 0;
 * the ten per-Gaussian sums of A.4 are linear in the moments S_0 = sum m, S_x = sum m dx, S_y, S_xx, S_xy, S_yy of
   m = G dL/dalpha (d = mean2D - pixel) plus the depth and rgb sums:
   dL_dmean2D = -o (A S_x + B S_y, B S_x + C S_y) (W/2, H/2),  dL_dconic = -o/2 (S_xx, S_xy, S_yy),  dL_dopacity = S_0.
The model below is a float64 numpy transcription of the kernel's per-pixel loop."""
import numpy as np

import oracle
from lvdgs import synth


def test_moment_formulation_equals_the_reference_sums():
    W, H = 64, 48
    cam = synth.Cam(W, H, 55.0, 52.0, W / 2.0 - 2.2, H / 2.0 + 1.7, np.eye(3), np.zeros(3))
    sc = synth.make_scene(260, cam, seed=9)
    sc["scales"] *= 2.5
    bg = np.array([0.3, 0.1, 0.6], np.float32)
    f = oracle.rasterize_forward(sc["means3D"], sc["opacities"], sc["scales"], sc["rotations"], sc["shs"],
                                 viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform,
                                 campos=cam.camera_center, bg=bg, W=W, H=H, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy)
    rng = np.random.default_rng(2)
    gc = rng.normal(0, 1, (3, H, W)).astype(np.float32)
    gd = rng.normal(0, 1, (1, H, W)).astype(np.float32)
    go = rng.normal(0, 1, (1, H, W)).astype(np.float32)
    ref = oracle.rasterize_backward(f, gc, gd, go, projmatrix_raw=cam.projection_matrix, flags=oracle.FLAG_OPACITY_GRAD)

    P = f["P"]
    mom = np.zeros((P, 6)); dcol = np.zeros((P, 3)); ddep = np.zeros(P)
    co = f["conic_opacity"].astype(np.float64); m2d = f["means2D"].astype(np.float64)
    rgb = f["rgb"].astype(np.float64); dep = f["depths"].astype(np.float64)
    gx = (W + 15) // 16
    for py in range(H):
        for px in range(W):
            r0, r1 = f["ranges"][(py // 16) * gx + px // 16]
            last = int(f["n_contrib"][py, px])
            Tf = float(f["final_T"][py, px])
            dp = np.array([gc[0, py, px], gc[1, py, px], gc[2, py, px], gd[0, py, px]], np.float64)
            bgd = float(bg.astype(np.float64) @ dp[:3]) - float(go[0, py, px])     # d(1 - T_final)/dalpha = +T_final/(1-alpha)
            T, S = Tf, 0.0
            for k in range(int(r1 - r0) - 1, -1, -1):          # every entry of the tile's list, rearmost first
                g = int(f["point_list"][r0 + k])
                dx, dy = m2d[g, 0] - px, m2d[g, 1] - py
                A, B, C, o = co[g]
                power = -0.5 * (A * dx * dx + C * dy * dy) - B * dx * dy
                G = np.exp(power)
                alpha = min(0.99, o * G)
                ok = k < last and power <= 0.0 and alpha >= 1.0 / 255.0
                if not ok:
                    G, alpha = 0.0, 0.0                        # branch-free: same arithmetic, nothing moves
                inv = 1.0 / (1.0 - alpha)
                T = T * inv
                c = np.array([rgb[g, 0], rgb[g, 1], rgb[g, 2], dep[g]])
                cdp = float(c @ dp)
                dL_dalpha = (cdp - S) * T - Tf * inv * bgd
                S = alpha * cdp + (1.0 - alpha) * S
                m = G * dL_dalpha
                mom[g] += (m * dx, m * dy, m * dx * dx, m * dx * dy, m * dy * dy, m)
                w = alpha * T
                dcol[g] += w * dp[:3]; ddep[g] += w * dp[3]
    A, B, C, o = co[:, 0], co[:, 1], co[:, 2], co[:, 3]
    dmean = np.stack([-0.5 * W * o * (A * mom[:, 0] + B * mom[:, 1]), -0.5 * H * o * (B * mom[:, 0] + C * mom[:, 1])], 1)
    dconic = np.stack([-0.5 * o * mom[:, 2], -0.5 * o * mom[:, 3], -0.5 * o * mom[:, 4]], 1)

    def close(a, b, name):
        err = np.abs(a - b).max() / (np.abs(b).max() + 1e-30)
        assert err < 2e-4, (name, err)

    assert (np.abs(ref["dL_dopacity"]) > 0).sum() > 50
    close(dmean, ref["dL_dmean2D"], "dL_dmean2D")
    close(dconic, ref["dL_dconic"], "dL_dconic")
    close(mom[:, 5], ref["dL_dopacity"], "dL_dopacity")
    close(dcol, ref["dL_dcolor"], "dL_dcolor")
    close(ddep, ref["dL_ddepth"], "dL_ddepth")
