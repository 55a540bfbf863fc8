"""float64 torch restatement of the rasterizer FORWARD, differentiable by autograd.

Used only to check the C oracle's hand-written analytic backward (SURVEY.md section 8c self-check (i)/(ii)).
Discrete decisions (tile lists, depth order, per-pixel contributor count) are taken from the oracle's forward
and held constant, which is what the analytic backward assumes (A.6 item 4).
The camera is perturbed on the left, T = Exp(tau) T_w2c with tau = [rho; theta], following
/root/reference/utils/pose_utils.py:22-87 (restated here in float64).
"""
import math

import torch

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435]


def skew(x):
    z = torch.zeros((), dtype=x.dtype)
    return torch.stack([torch.stack([z, -x[2], x[1]]), torch.stack([x[2], z, -x[0]]), torch.stack([-x[1], x[0], z])])


def SE3_exp(tau):
    rho, theta = tau[:3], tau[3:]
    Wm = skew(theta)
    W2 = Wm @ Wm
    I = torch.eye(3, dtype=tau.dtype)
    # tau is evaluated at 0: small-angle branch of pose_utils.SO3_exp / V
    R = I + Wm + 0.5 * W2
    V = I + 0.5 * Wm + (1.0 / 6.0) * W2
    T = torch.eye(4, dtype=tau.dtype)
    T = T.clone()
    T[:3, :3] = R
    T[:3, 3] = V @ rho
    return T


def eval_sh(deg, sh, dirs):
    r = SH_C0 * sh[:, 0]
    if deg > 0:
        x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
        r = r - SH_C1 * y * sh[:, 1] + SH_C1 * z * sh[:, 2] - SH_C1 * x * sh[:, 3]
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            r = (r + SH_C2[0] * xy * sh[:, 4] + SH_C2[1] * yz * sh[:, 5] + SH_C2[2] * (2 * zz - xx - yy) * sh[:, 6]
                 + SH_C2[3] * xz * sh[:, 7] + SH_C2[4] * (xx - yy) * sh[:, 8])
            if deg > 2:
                r = (r + SH_C3[0] * y * (3 * xx - yy) * sh[:, 9] + SH_C3[1] * xy * z * sh[:, 10]
                     + SH_C3[2] * y * (4 * zz - xx - yy) * sh[:, 11] + SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12]
                     + SH_C3[4] * x * (4 * zz - xx - yy) * sh[:, 13] + SH_C3[5] * z * (xx - yy) * sh[:, 14]
                     + SH_C3[6] * x * (xx - 3 * yy) * sh[:, 15])
    return torch.clamp_min(r + 0.5, 0.0)


def forward(means3D, scales, rotations, opacities, shs, tau, *, W2C, Pr, campos, bg, W, H, tanfovx, tanfovy,
            sh_degree, fwd, colors_precomp=None, valid_cache=None):
    """W2C, Pr: math-convention 4x4 float64 (NOT transposed). fwd: oracle forward dict (constants).
    valid_cache: optional dict; the per-tile blend decisions (power <= 0, alpha >= 1/255) are stored in it on the first
    call and reused afterwards, so that finite differences do not step across those discontinuities."""
    dt = torch.float64
    V = SE3_exp(tau) @ W2C
    Pj = Pr @ V
    N = means3D.shape[0]
    ph = torch.cat([means3D, torch.ones(N, 1, dtype=dt)], 1)
    pv = ph @ V.T
    hom = ph @ Pj.T
    pw = 1.0 / (hom[:, 3] + 1e-7)
    ndc = hom[:, :2] * pw[:, None]
    pix = torch.stack([((ndc[:, 0] + 1) * W - 1) * 0.5, ((ndc[:, 1] + 1) * H - 1) * 0.5], 1)
    r, x, y, z = rotations.unbind(1)
    Rq = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                      2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                      2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1).reshape(N, 3, 3)
    A = Rq * scales[:, None, :]
    Sigma = A @ A.transpose(1, 2)
    fx, fy = W / (2 * tanfovx), H / (2 * tanfovy)
    tz = pv[:, 2]
    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    txc = torch.clamp(pv[:, 0] / tz, -limx, limx) * tz
    tyc = torch.clamp(pv[:, 1] / tz, -limy, limy) * tz
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz, zero, -fx * txc / (tz * tz), zero, fy / tz, -fy * tyc / (tz * tz)], 1).reshape(N, 2, 3)
    Mm = J @ V[:3, :3]
    cov = Mm @ Sigma @ Mm.transpose(1, 2)
    a = cov[:, 0, 0] + 0.3
    b = cov[:, 0, 1]
    c = cov[:, 1, 1] + 0.3
    det = a * c - b * b
    conic = torch.stack([c / det, -b / det, a / det], 1)
    if colors_precomp is not None:
        rgb = colors_precomp
    else:
        d = means3D - campos[None]
        d = d / d.norm(dim=1, keepdim=True)
        rgb = eval_sh(sh_degree, shs, d)
    depth = tz
    gx = (W + 15) // 16
    gy = (H + 15) // 16
    ranges = fwd["ranges"]
    plist = torch.from_numpy(fwd["point_list"].astype("int64"))
    ncontrib = torch.from_numpy(fwd["n_contrib"].astype("int64"))
    color = torch.zeros(3, H, W, dtype=dt)
    dimg = torch.zeros(H, W, dtype=dt)
    oimg = torch.zeros(H, W, dtype=dt)
    bgt = torch.as_tensor(bg, dtype=dt)
    for ty in range(gy):
        for tx in range(gx):
            r0, r1 = int(ranges[ty * gx + tx, 0]), int(ranges[ty * gx + tx, 1])
            y0, y1 = ty * 16, min(ty * 16 + 16, H)
            x0, x1 = tx * 16, min(tx * 16 + 16, W)
            ys, xs = torch.meshgrid(torch.arange(y0, y1), torch.arange(x0, x1), indexing="ij")
            pf = torch.stack([xs.reshape(-1), ys.reshape(-1)], 1).to(dt)
            npx = pf.shape[0]
            if r1 > r0:
                ids = plist[r0:r1]
                dxy = pix[ids][None, :, :] - pf[:, None, :]
                con = conic[ids][None]
                power = -0.5 * (con[..., 0] * dxy[..., 0] ** 2 + con[..., 2] * dxy[..., 1] ** 2) - con[..., 1] * dxy[..., 0] * dxy[..., 1]
                alpha = torch.clamp_max(opacities[ids].reshape(1, -1) * torch.exp(power), 0.99)
                k = torch.arange(r1 - r0)[None, :]
                valid = (power <= 0) & (alpha >= 1.0 / 255.0) & (k < ncontrib[y0:y1, x0:x1].reshape(-1, 1))
                if valid_cache is not None:
                    valid = valid_cache.setdefault((ty, tx), valid.detach())
                aeff = torch.where(valid, alpha, torch.zeros_like(alpha))
                Tincl = torch.cumprod(1 - aeff, 1)
                Texcl = torch.cat([torch.ones(npx, 1, dtype=dt), Tincl[:, :-1]], 1)
                w = aeff * Texcl
                Tfin = Tincl[:, -1]
                col = w @ rgb[ids]
                dep = w @ depth[ids]
            else:
                Tfin = torch.ones(npx, dtype=dt)
                col = torch.zeros(npx, 3, dtype=dt)
                dep = torch.zeros(npx, dtype=dt)
            col = col + Tfin[:, None] * bgt[None]
            color[:, y0:y1, x0:x1] = col.T.reshape(3, y1 - y0, x1 - x0)
            dimg[y0:y1, x0:x1] = dep.reshape(y1 - y0, x1 - x0)
            oimg[y0:y1, x0:x1] = (1 - Tfin).reshape(y1 - y0, x1 - x0)
    return color, dimg, oimg
