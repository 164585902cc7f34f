"""Summarise an ncu report: headline metrics + the top stall locations (needs -lineinfo builds).
usage: python tools_dev/ncu_top.py report.ncu-rep [n_top]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, v = rows[0], rows[-1]
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "lts__t_bytes.sum", "lts__t_bytes.sum.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "smsp__cycles_elapsed.avg.per_second"]
units = rows[1] if len(rows) > 2 else [""] * len(h)
for k, u, x in zip(h, units, v):
    if k in keys:
        print(f"{k:90s} {x} {u}")
for k, x in zip(h, v):
    if k.startswith("smsp__pcsamp_warps_issue_stalled") and not k.endswith("not_issued") and float(x or 0) > 0:
        print(f"{k:90s} {x}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; data = rows[2:]
ia, isrc, isamp, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r[isamp] or 0) for r in data)
print("total samples", tot)
for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:ntop]:
    st = sorted(((h[i], int(r[i] or 0)) for i in stall_cols if int(r[i] or 0) > 0), key=lambda x: -x[1])[:3]
    print(r[ia][-5:], f"{int(r[isamp]) / tot * 100:5.1f}%", r[iex], r[isrc][:70], st)
