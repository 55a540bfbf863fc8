"""`simple_knn._C.distCUDA2` over liblvdgs.so (include/lvdgs.h: lvdgs_dist2).  CUDA only; no CPU fallback."""
import ctypes as C

import torch

from lvdgs import _native


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    """points [P,3] float32 CUDA -> [P] float32: mean squared distance to the 3 nearest neighbours."""
    if not points.is_cuda:
        raise RuntimeError("distCUDA2 (B200): points must be a CUDA tensor; there is no CPU path")
    L = _native.lib()
    pts = points.detach()
    if pts.dtype != torch.float32 or not pts.is_contiguous():
        pts = pts.contiguous().float()
    P = pts.shape[0]
    out = torch.empty(P, dtype=torch.float32, device=pts.device)
    ws_bytes = L.lvdgs_dist2_workspace_bytes(P)
    ws = torch.empty(max(int(ws_bytes), 1), dtype=torch.uint8, device=pts.device)
    if pts.device.index is not None:
        L.lvdgs_set_device(pts.device.index)
    stream = C.c_void_p(torch.cuda.current_stream(pts.device).cuda_stream)
    rc = L.lvdgs_dist2(P, _native.ptr(pts), _native.ptr(out), _native.ptr(ws), C.c_size_t(ws.numel()), stream)
    _native.check(rc, "lvdgs_dist2")
    return out
