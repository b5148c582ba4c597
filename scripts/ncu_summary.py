"""Summarise an .ncu-rep (first kernel): duration, DRAM bytes, pipes, stall reasons, top stalled source lines."""
import csv, subprocess, sys, io

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
units = dict(zip(hdr, rows[1]))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0      # n-th profiled kernel of the report
vals = rows[2 + which]
m = dict(zip(hdr, vals))
def g(k):
    try: return float(m[k].replace(",", ""))
    except Exception: return float("nan")
print("kernel:", m.get("Kernel Name"), "grid", m.get("Grid Size"), "block", m.get("Block Size"))
print(f"duration_us={g('gpu__time_duration.sum')/1e3 if g('gpu__time_duration.sum')>1e4 else g('gpu__time_duration.sum'):.2f} ({m.get('gpu__time_duration.sum')})")
for k in ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
          "lts__throughput.avg.pct_of_peak_sustained_elapsed",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "smsp__inst_executed.sum",
          "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
          "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]:
    if k in m: print(f"  {k} = {m[k]} {units.get(k, '')}")
st = []
for h, v in m.items():
    if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
        try: st.append((float(v.replace(",", "")), h.split("stalled_")[1].split("_per_issue")[0]))
        except Exception: pass
print("stalls (warps per issue):", ", ".join(f"{n}={f:.2f}" for f, n in sorted(st, reverse=True)[:8]))
