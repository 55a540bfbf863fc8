"""Per-source-line totals (stall samples, executed warp instructions) from an ncu report captured with --import-source:
python scripts/ncu_source.py report.ncu-rep [top] [by=stall|inst]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25; by = sys.argv[3] if len(sys.argv) > 3 else "stall"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
agg = collections.OrderedDict()
hdr = None
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        iS = hdr.index("Warp Stall Sampling (All Samples)"); iI = hdr.index("Instructions Executed"); iT = hdr.index("Thread Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr) - 5: continue
    try:
        key = (r[0], r[1].strip()[:105]); st = int(r[iS] or 0); ins = int(r[iI] or 0); th = int(r[iT] or 0)
    except ValueError:
        continue
    a = agg.setdefault(key, [0, 0, 0]); a[0] += st; a[1] += ins; a[2] += th
ts = sum(a[0] for a in agg.values()) or 1; ti = sum(a[1] for a in agg.values()) or 1
print(f"total stall samples {ts}, warp instructions {ti}")
idx = 0 if by == "stall" else 1
for (ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][idx])[:top]:
    print(f"{100*a[0]/ts:5.1f}% stall {100*a[1]/ti:5.1f}% inst  lanes {a[2]/max(a[1],1):4.1f}  L{ln:>4} {src}")
