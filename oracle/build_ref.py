"""oracle/_ref: the reference's OWN rasterizer, compiled from its sources where they lie -- if they are ever there.

Test infrastructure (like everything under oracle/).  The CUDA sources of `diff-gaussian-rasterization` and `simple-knn`
live in `submodules.zip`, a blob that is absent from /root/reference (`.MISSING_LARGE_BLOBS:1`), so today this script
reports "absent" and exits 0; parity stays PINNED ONLY THROUGH THE PYTHON FILES either side of the rasterizer
(DESIGN.md section 2).  Should the submodules appear (unzipped under /root/reference/submodules/), this builds them with
torch.utils.cpp_extension straight from that read-only tree into oracle/_ref/ (git-ignored, travels to the GPU box), and
tests/test_gpu_reference_build.py compares the sm_100a kernels with them on identical inputs at north_star's
tolerances.  No reference source is copied into the repository.

    python oracle/build_ref.py            # -> oracle/_ref/{diff_gaussian_rasterization_ref,simple_knn_ref}*.so or "absent"
"""
import glob
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("LVDGS_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
CANDIDATES = {
    "diff_gaussian_rasterization_ref": ("submodules/diff-gaussian-rasterization", ["rasterize_points.cu", "ext.cpp", "cuda_rasterizer/*.cu"],
                                        ["third_party/glm"]),
    "simple_knn_ref": ("submodules/simple-knn", ["simple_knn.cu", "spatial.cu", "ext.cpp"], []),
}


def find(sub):
    for base in (REF_ROOT, os.path.join(REF_ROOT, "LVD_GS-SLAM")):
        d = os.path.join(base, sub)
        if os.path.isdir(d):
            return d
    return None


def build() -> dict:
    """Returns {name: path of the built module or None}."""
    built = {}
    for name, (sub, patterns, incs) in CANDIDATES.items():
        src_dir = find(sub)
        if src_dir is None:
            built[name] = None
            continue
        sources = sorted(sum((glob.glob(os.path.join(src_dir, p)) for p in patterns), []))
        if not sources:
            built[name] = None
            continue
        from torch.utils.cpp_extension import load
        os.makedirs(OUT, exist_ok=True)
        # the extension's own module name is baked into its PYBIND11_MODULE(TORCH_EXTENSION_NAME, ...): load() defines it
        load(name=name, sources=sources, extra_include_paths=[os.path.join(src_dir, i) for i in incs],
             extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"], build_directory=OUT,
             is_python_module=False, verbose=False)
        hits = glob.glob(os.path.join(OUT, name + "*.so"))
        built[name] = hits[0] if hits else None
    return built


def available() -> dict:
    return {name: (glob.glob(os.path.join(OUT, name + "*.so")) or [None])[0] for name in CANDIDATES}


if __name__ == "__main__":
    res = build()
    for k, v in res.items():
        print(f"{k}: {v if v else 'absent (reference sources not under ' + REF_ROOT + ')'}")
    sys.exit(0)
