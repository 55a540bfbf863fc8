"""profiles/ncu_kernel_metrics.json (read by bench.py's roofline entry) from the condensed ncu summaries:
python scripts/ncu_metrics_json.py kernel=profiles/r01_ncu_<...>.txt ..."""
import json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_path = os.path.join(ROOT, "profiles", "ncu_kernel_metrics.json")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
try:
    out = json.load(open(out_path))
except Exception:
    out = {}
for arg in sys.argv[1:]:
    name, path = arg.split("=")
    vals = {}
    for line in open(os.path.join(ROOT, path)):
        m = re.match(r"\s+(\S+) \[(.*?)\] = (\S+)", line)
        if m:
            vals[m.group(1)] = (m.group(2), float(m.group(3).replace(",", "")))
    g = lambda k: vals.get(k, ("", 0.0))[1]
    dram = sum(vals[k][1] * UNIT.get(vals[k][0], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum") if k in vals)
    out[name] = {
        "dram_bytes": dram, "duration_us_cold": g("gpu__time_duration.sum"),
        "smsp_issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "pipe_fma_pct": g("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "pipe_alu_pct": g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "pipe_xu_pct": g("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        "pipe_lsu_pct": g("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "inst_executed": g("smsp__inst_executed.sum"),
        "registers": g("launch__registers_per_thread"),
        "source": f"{path} (ncu --set full --clock-control none, view 0 of the 500k KITTI workload)"}
json.dump(out, open(out_path, "w"), indent=1)
print(json.dumps(out, indent=1))
