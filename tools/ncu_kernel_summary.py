#!/usr/bin/env python
"""Key figures of one ncu --set full capture (first kernel in the report): time, registers, occupancy, pipe utilisation, stall mix,
dram bytes, and thread-level FP64 flops from the source page (DFMA x 2 + DMUL + DADD + DMMA.8x8x4 x 512 per warp instruction).
usage: ncu_kernel_summary.py <report.ncu-rep> [pulse_steps]  -> JSON on stdout"""
import csv, io, json, subprocess, sys
rep = sys.argv[1]
pulse_steps = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h, v = r[0], r[2]
m = dict(zip(h, v))
def g(k):
    try:
        return float(m[k].replace(",", ""))
    except Exception:
        return None
out = {"kernel": m.get("Kernel Name"), "gpu_time_us": g("gpu__time_duration.sum"), "registers": g("launch__registers_per_thread"),
       "ctas_per_sm_by_registers": g("launch__occupancy_limit_registers"),
       "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
       "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
       "fp64_pipe_active_pct": g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
       "dmma_pipe_active_pct": g("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"),
       "warp_instructions": g("smsp__inst_executed.sum"),
       "dram_bytes_read": g("dram__bytes_read.sum"), "dram_bytes_write": g("dram__bytes_write.sum"),
       "dram_units": (r[1][h.index("dram__bytes_read.sum")], r[1][h.index("dram__bytes_write.sum")]),
       "l2_hit_rate_pct": g("lts__t_sector_hit_rate.pct"),
       "stalls_per_issue": {k.split("stalled_")[1].split("_per_issue")[0]: g(k) for k in h
                            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and (g(k) or 0) > 0.2}}
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
it, ie = hdr.index("Thread Instructions Executed"), hdr.index("Instructions Executed")
fl = 0.0
cnt = {}
for row in rows[2:]:
    if not row or not row[0].startswith("0x"):
        continue
    toks = row[1].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    base = op.split(".")[0]
    if base in ("DFMA", "DMUL", "DADD", "DMMA"):
        cnt[base] = cnt.get(base, 0) + (int(row[ie] or 0) if base == "DMMA" else int(row[it] or 0))
fl = 2.0 * cnt.get("DFMA", 0) + cnt.get("DMUL", 0) + cnt.get("DADD", 0) + 512.0 * cnt.get("DMMA", 0)
out["fp64_counts"] = cnt
out["fp64_flops"] = fl
if pulse_steps:
    out["fp64_flops_per_pulse_step"] = fl / pulse_steps
if out["gpu_time_us"]:
    out["fp64_tflops_under_ncu"] = fl / (out["gpu_time_us"] * 1e-6) / 1e12
print(json.dumps(out, indent=1))
