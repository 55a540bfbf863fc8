"""Instruction-mix probes of the FP32 pipes (csrc/peak.cu): warp instructions per clock per SM sub-partition for scalar
FFMA, packed FFMA2 / FMUL2 / FADD2, an ALU-pipe select, and their mixes.  Usage (GPU box): python scripts/pipe_probe.py"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lvd_gs-slam_b200")]
import torch
from lvdgs import _native
L = _native.lib()
dev = torch.device("cuda")
props = torch.cuda.get_device_properties(dev)
n_sm = props.multi_processor_count
out = torch.zeros(4, device=dev)
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
names = {0: "16 FFMA", 1: "16 FFMA2", 2: "16 FMUL2", 3: "16 FADD2", 4: "8 FFMA2 + 8 FFMA", 5: "8 FFMA2 + 8 (SETP+SELP)",
         6: "16 (SETP+SELP)", 7: "8 FFMA + 8 (SETP+SELP)"}
instr_per_round = {0: 16, 1: 16, 2: 16, 3: 16, 4: 16, 5: 24, 6: 32, 7: 24}
import subprocess
clk = float(subprocess.run(["nvidia-smi", "--query-gpu=clocks.max.sm", "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.split()[0])
res = {}
for blocks_per_sm in (8, 4):
    for mode in range(8):
        best = 1e9
        for _ in range(6):
            work = C.c_double(0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.lvdgs_fp32_peak(n_sm * blocks_per_sm, 4096, mode, C.c_void_p(out.data_ptr()), C.byref(work), stream)
            e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        warp_instr = n_sm * blocks_per_sm * 8 * 4096 * instr_per_round[mode]
        ipc = warp_instr / (best * 1e-3 * clk * 1e6) / (n_sm * 4)
        res[f"{names[mode]} @ {blocks_per_sm * 8} warps/SM"] = {"ms": best, "warp_instr_per_clk_per_smsp": round(ipc, 3)}
        print(f"{names[mode]:28s} {blocks_per_sm * 8:3d} warps/SM  {best:8.3f} ms   {ipc:.3f} warp-instr/clk/SMSP (at {clk:.0f} MHz)", flush=True)
json.dump({"sm_mhz_max": clk, "n_sm": n_sm, "probes": res}, open(os.path.join(ROOT, "gpurun_out", "pipe_probe.json"), "w"), indent=1)
