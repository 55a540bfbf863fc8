"""Build a tuning variant of the library next to the product one: python scripts/build_variant.py NAME "-DFOO=1 -DBAR=2"
-> lvd_gs-slam_b200/variants/liblvdgs_NAME.so; run anything with LVDGS_SO=<that path> to use it."""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lvd_gs-slam_b200")]
from lvdgs import _native as N
name, defs = sys.argv[1], (sys.argv[2].split() if len(sys.argv) > 2 else [])
out_dir = os.path.join(N.PKG_DIR, "variants"); bdir = os.path.join(N.BUILD_DIR, "variant_" + name)
os.makedirs(out_dir, exist_ok=True); os.makedirs(bdir, exist_ok=True)
def cc(src):
    o = os.path.join(bdir, src.replace(".cu", ".o"))
    subprocess.run([N._nvcc(), "-c", *N.NVCC_FLAGS, *defs, "-o", o, os.path.join(N.SRC_DIR, src)], check=True)
    return o
with ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(cc, N.SOURCES))
so = os.path.join(out_dir, f"liblvdgs_{name}.so")
subprocess.run([N._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", so, *objs], check=True)
print(so)
