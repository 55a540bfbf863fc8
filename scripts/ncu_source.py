"""Per-source-line executed warp instructions from an ncu report with --import-source: python scripts/ncu_source.py rep [top]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# find header
hi = next(i for i, r in enumerate(rows) if "Source" in r and any("Instructions Executed" in c for c in r))
hdr = rows[hi]
ci = hdr.index("Source"); ie = next(i for i, c in enumerate(hdr) if c == "Instructions Executed")
ws = next((i for i, c in enumerate(hdr) if c.startswith("Warp Stall Sampling (All")), None)
tot = 0; lines = []
for r in rows[hi + 1:]:
    if len(r) <= ie: continue
    try: n = int(r[ie])
    except ValueError: continue
    tot += n
    s = int(r[ws]) if ws is not None and r[ws].isdigit() else 0
    lines.append((n, s, r[ci].strip()[:110]))
print("total instructions", tot)
for n, s, src in sorted(lines, reverse=True)[:top]:
    print(f"{n:12d} {100*n/tot:5.1f}%  stall_samples {s:6d}  {src}")
