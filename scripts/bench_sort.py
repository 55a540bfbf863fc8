"""Onesweep radix sort (ours) vs cub::DeviceRadixSort::SortPairs (the reference's K4 library call) on (tile|depth) keys."""
import ctypes as C, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lvd_gs-slam_b200")]
import numpy as np, torch
from lvdgs import _native
L = _native.lib(); p = _native.ptr
dev = "cuda"
def bench(n, end_bit=43, iters=20):
    rng = np.random.default_rng(0)
    tiles = rng.integers(0, 1872, n, dtype=np.uint64)
    depth = np.exp(rng.uniform(np.log(0.2), np.log(100.0), n)).astype(np.float32).view(np.uint32).astype(np.uint64)
    keys = torch.from_numpy(((tiles << np.uint64(32)) | depth).view(np.int64)).to(dev)
    vals = torch.arange(n, dtype=torch.int32, device=dev)
    k0, k1, v0, v1 = keys.clone(), torch.empty_like(keys), vals.clone(), torch.empty_like(vals)
    ws = torch.empty(L.lvdgs_sort_workspace_bytes(n), dtype=torch.uint8, device=dev)
    wc = torch.empty(max(1, L.lvdgs_cub_sort_workspace_bytes(n, end_bit)), dtype=torch.uint8, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    sel = C.c_int32(0)
    res = {}
    for name in ("ours", "cub"):
        ts = []
        for it in range(iters + 3):
            k0.copy_(keys); v0.copy_(vals)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if name == "ours":
                L.lvdgs_sort_pairs(n, p(k0), p(k1), p(v0), p(v1), end_bit, p(ws), ws.numel(), C.byref(sel), st)
            else:
                L.lvdgs_cub_sort_pairs(n, p(k0), p(k1), p(v0), p(v1), end_bit, p(wc), wc.numel(), st)
            e1.record(); torch.cuda.synchronize()
            if it >= 3: ts.append(e0.elapsed_time(e1))
        res[name] = float(np.median(ts))
    res["n"] = n
    res["ours_GBs_algorithmic"] = (8 + 6 * 24) * n / (res["ours"] * 1e-3) / 1e9
    return res
for n in [int(x) for x in (sys.argv[1:] or ["1340000", "5000000", "20000000", "60000000"])]:
    print(json.dumps(bench(n)))
