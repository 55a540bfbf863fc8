import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lvd_gs-slam_b200")]
import numpy as np, torch
from lvdgs import _native
L = _native.lib(); p = _native.ptr
n = 1340000
rng = np.random.default_rng(0)
tiles = rng.integers(0, 1872, n, dtype=np.uint64)
depth = np.exp(rng.uniform(np.log(0.2), np.log(100.0), n)).astype(np.float32).view(np.uint32).astype(np.uint64)
keys = torch.from_numpy(((tiles << np.uint64(32)) | depth).view(np.int64)).cuda()
vals = torch.arange(n, dtype=torch.int32, device="cuda")
k1, v1 = torch.empty_like(keys), torch.empty_like(vals)
ws = torch.empty(L.lvdgs_sort_workspace_bytes(n), dtype=torch.uint8, device="cuda")
sel = C.c_int32(0)
for _ in range(3):
    k0, v0 = keys.clone(), vals.clone()
    L.lvdgs_sort_pairs(n, p(k0), p(k1), p(v0), p(v1), 43, p(ws), ws.numel(), C.byref(sel), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
nb = (n + 4095) // 4096
buf = (C.c_longlong * (nb * 8))()
L.lvdgs_debug_sort_timing.argtypes = [C.c_void_p, C.c_int]
print("rc", L.lvdgs_debug_sort_timing(buf, nb * 8))
t = np.array(buf, dtype=np.int64).reshape(nb, 8)
d = np.diff(t, axis=1)
names = ["load+n", "rank", "digit-scan+lookback", "smem scatter keys", "write keys", "scatter vals", "write vals"]
print("blocks", nb, "(last pass of the sort); median cycles per phase [p10, p50, p90]:")
for i, nm in enumerate(names):
    print(f"  {nm:24s} {np.percentile(d[:, i], 10):9.0f} {np.percentile(d[:, i], 50):9.0f} {np.percentile(d[:, i], 90):9.0f}")
tot = t[:, 7] - t[:, 0]
print("  total after ticket       ", np.percentile(tot, 10), np.percentile(tot, 50), np.percentile(tot, 90))
print("  kernel span (max end - min start) cycles:", t[:, 7].max() - t[:, 0].min())
