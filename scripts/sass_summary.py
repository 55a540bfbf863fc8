"""Static evidence for profiles/: per kernel, ptxas resource usage (registers, shared memory, spills) and the SASS
instruction mix (cuobjdump) -- e.g. the packed-FP32 FFMA2/FMUL2/FADD2 of the blend kernels and the VIMNMX comparators of the
tile sort.  No GPU needed.  Usage: python scripts/sass_summary.py > profiles/r01_sass_summary.txt"""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "lvd_gs-slam_b200", "csrc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--expt-relaxed-constexpr", "-Xptxas", "-v"]
KEEP = ("FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "MUFU", "VIMNMX", "FMNMX", "ISETP", "FSETP", "SEL", "FSEL", "LDS", "STS",
        "LDG", "STG", "REDG", "ATOMG", "SHFL", "VOTE", "BAR", "IMAD", "LOP3")
for f in sorted(os.listdir(SRC)):
    if not f.endswith(".cu") or f == "cub_compare.cu":
        continue
    with tempfile.TemporaryDirectory() as td:
        o = os.path.join(td, "k.o")
        r = subprocess.run(["nvcc", "-c", *FLAGS, os.path.join(SRC, f), "-o", o], capture_output=True, text=True)
        info = r.stderr
        res = {}
        for m in re.finditer(r"Compiling entry function '(\S+)'.*?\n.*?\n.*?(\d+) bytes spill stores.*?\n.*?Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?", info):
            res[m.group(1)] = (int(m.group(3)), int(m.group(4) or 0), int(m.group(2)))
        sass = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
    print(f"## {f}")
    cur, mix = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1); mix[cur] = collections.Counter(); continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            mix[cur][m.group(1)] += 1
    for fn, c in mix.items():
        dem = subprocess.run(["cu++filt", fn], capture_output=True, text=True).stdout.strip() or fn
        name = re.sub(r"\(.*", "", dem.replace("(bool)", "").replace("(int)", "")).replace("void lvdgs::", "").replace("lvdgs::", "")
        regs, smem, spill = res.get(fn, (None, None, None))
        total = sum(c.values())
        shown = ", ".join(f"{k} {c[k]}" for k in KEEP if c.get(k))
        print(f"- {name}: {regs} regs, {smem} B static smem, {spill} B spills, {total} SASS instructions: {shown}")
    print()
