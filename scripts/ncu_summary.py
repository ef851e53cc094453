"""Summarise an .ncu-rep: key raw metrics + hottest SASS lines by stall samples. usage: ncu_summary.py file.ncu-rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 18
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct"]
for vals in rows[2:]:
    for h, u, v in zip(hdr, units, vals):
        if h in want:
            print(f"  {h} [{u}] = {v}")
    print("  stalls:", ", ".join(f"{h.split('issue_stalled_')[1].split('_per')[0]}={float(v):.1f}" for h, v in zip(hdr, vals)
                                   if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(v or 0) > 0.5))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[hi]
i_src, i_s, i_ex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = [(int(r[i_s] or 0), r[i_src].strip(), int(r[i_ex] or 0)) for r in rows[hi + 1:] if len(r) > i_ex and r[i_s].isdigit()]
tot = sum(d[0] for d in data) or 1
print(f"  total stall samples {tot}")
order = sorted(range(len(data)), key=lambda k: -data[k][0])[:top]
for k in sorted(order):
    s, sc, ex = data[k]
    print(f"   {k:5d} {100 * s / tot:5.1f}% ex={ex:8d}  {sc}")
