"""Shared helpers for the GPU parity tests: run the CUDA path through the public plugin surface
(diff_gaussian_rasterization -> ctypes -> liblvdgs.so C ABI) and the oracle on the same inputs."""
import numpy as np
import torch

import oracle
from lvdgs import _native, synth


def settings_for(cam, bg, sh_degree, device="cuda", debug=False):
    from diff_gaussian_rasterization import GaussianRasterizationSettings
    t = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32, device=device)
    return GaussianRasterizationSettings(
        image_height=cam.image_height, image_width=cam.image_width, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
        bg=t(bg), scale_modifier=1.0, viewmatrix=t(cam.world_view_transform), projmatrix=t(cam.full_proj_transform),
        projmatrix_raw=t(cam.projection_matrix), sh_degree=sh_degree, campos=t(cam.camera_center), prefiltered=False,
        debug=debug)


def run_cuda(sc, cam, bg, grads=None, device="cuda", use_precomp_color=False, use_precomp_cov=None, debug=True):
    """Returns (outputs dict of numpy, internals dict of numpy, grads dict of numpy or None)."""
    import diff_gaussian_rasterization as dgr
    t = lambda a, rg=True: torch.tensor(np.asarray(a), dtype=torch.float32, device=device, requires_grad=rg)
    means3D = t(sc["means3D"]); opac = t(sc["opacities"])
    scales = rots = cov = shs = cp = None
    if use_precomp_cov is not None:
        cov = t(use_precomp_cov)
    else:
        scales = t(sc["scales"]); rots = t(sc["rotations"])
    if use_precomp_color:
        cp = t(sc["colors_precomp"])
    else:
        shs = t(sc["shs"])
    means2D = torch.zeros_like(means3D, requires_grad=True)
    theta = torch.zeros(3, device=device, requires_grad=True)
    rho = torch.zeros(3, device=device, requires_grad=True)
    rs = settings_for(cam, bg, sc.get("sh_degree", 0), device, debug=debug)
    rast = dgr.GaussianRasterizer(rs)
    # keep a handle on the opaque buffers: patch through a subclass of the autograd ctx is awkward, so re-run the
    # forward by hand for the internals
    color, radii, depth, opacity, n_touched = rast(means3D=means3D, means2D=means2D, opacities=opac, shs=shs,
                                                   colors_precomp=cp, scales=scales, rotations=rots, cov3D_precomp=cov,
                                                   theta=theta, rho=rho)
    torch.cuda.synchronize()
    out = dict(color=color.detach().cpu().numpy(), radii=radii.cpu().numpy(), depth=depth.detach().cpu().numpy(),
               opacity=opacity.detach().cpu().numpy(), n_touched=n_touched.cpu().numpy())
    ctx_bufs = dgr.debug_buffers(color.grad_fn) if hasattr(color.grad_fn, "arena") else None
    R = color.grad_fn.num_rendered if hasattr(color.grad_fn, "num_rendered") else None
    internals = None
    if ctx_bufs is not None:
        internals = _native.debug_views(ctx_bufs, means3D.shape[0], R, cam.image_width, cam.image_height,
                                        capacity=color.grad_fn.capacity)
        internals["R"] = R
    g = None
    if grads is not None:
        gc, gd, go = grads
        loss = (color * torch.tensor(gc, device=device)).sum()
        if gd is not None:
            loss = loss + (depth * torch.tensor(gd, device=device)).sum()
        if go is not None:
            loss = loss + (opacity * torch.tensor(go, device=device)).sum()
        loss.backward()
        torch.cuda.synchronize()
        n = lambda x: None if x is None or x.grad is None else x.grad.detach().cpu().numpy()
        g = dict(means3D=n(means3D), means2D=n(means2D), opacities=n(opac), scales=n(scales), rotations=n(rots),
                 cov3D=n(cov), shs=n(shs), colors_precomp=n(cp), theta=n(theta), rho=n(rho))
    return out, internals, g


def run_oracle(sc, cam, bg, grads=None, flags=0, use_precomp_color=False, use_precomp_cov=None):
    kw = dict(viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, campos=cam.camera_center,
              bg=bg, W=cam.image_width, H=cam.image_height, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
              sh_degree=sc.get("sh_degree", 0))
    fwd = oracle.rasterize_forward(sc["means3D"], sc["opacities"],
                                   None if use_precomp_cov is not None else sc["scales"],
                                   None if use_precomp_cov is not None else sc["rotations"],
                                   None if use_precomp_color else sc["shs"],
                                   sc["colors_precomp"] if use_precomp_color else None,
                                   use_precomp_cov, **kw)
    g = None
    if grads is not None:
        gc, gd, go = grads
        g = oracle.rasterize_backward(fwd, gc, gd, go, projmatrix_raw=cam.projection_matrix, flags=flags)
    return fwd, g


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-30)


def tainted_gaussians(fwd, thr=1e-5):
    """Gaussians that can reach a knife-edge pixel (oracle margin <= thr): in that pixel's tile list and with the pixel
    inside their 3-sigma square.  Their blend decisions there legitimately depend on the last ulp of exp(), so gradient
    comparisons set them aside (as the forward comparison sets the pixels aside)."""
    W, H = fwd["W"], fwd["H"]
    gx = (W + 15) // 16
    tainted = np.zeros(fwd["P"], bool)
    ys, xs = np.nonzero(~(fwd["margin"] > thr))
    for y, x in zip(ys, xs):
        r0, r1 = fwd["ranges"][(y // 16) * gx + x // 16]
        ids = fwd["point_list"][r0:r1]
        m, r = fwd["means2D"][ids], fwd["radii"][ids].astype(np.float32) + 1.0
        tainted[ids[(np.abs(m[:, 0] - x) <= r) & (np.abs(m[:, 1] - y) <= r)]] = True
    return tainted


def grad_mismatch(a, b, rows=None, rtol=1e-3):
    """Fraction of elements violating north_star's gradient bar PER ELEMENT: |a - b| <= rtol |b| + rtol * median|b|
    (median over the non-zero reference elements: the absolute floor for elements that are themselves rounding noise).
    rows: optional boolean mask of the rows (Gaussians) to compare."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    a = a.reshape(b.shape)
    if rows is not None:
        a, b = a[rows], b[rows]
    nz = np.abs(b[b != 0])
    atol = rtol * (np.median(nz) if nz.size else 0.0)
    bad = np.abs(a - b) > rtol * np.abs(b) + atol
    return float(bad.mean()) if bad.size else 0.0
