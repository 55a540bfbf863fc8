"""Rebuild liblvdgs.so with each set of -D knobs given on the command line and print the per-kernel profile of one
500k-Gaussian KITTI view for it.  Usage (on a GPU box): python scripts/try_defs.py "" "-DLVDGS_CURSOR_STRIDE=32" ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = os.environ.get("TRY_ARGS", "500000 kitti 10").split()
for defs in sys.argv[1:]:
    env = dict(os.environ, LVDGS_NVCC_DEFS=defs)
    subprocess.run([sys.executable, "-c", "import sys; sys.path[:0]=[%r, %r]; from lvdgs import _native; _native.build(force=True)"
                    % (ROOT, os.path.join(ROOT, "lvd_gs-slam_b200"))], check=True, env=env)
    print("=== defs: %r" % defs, flush=True)
    subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "profile_step.py"), *args], env=env)
