"""ncu report(s) -> markdown table of the metrics the profiles/ summaries quote.  usage: summarize_ncu.py out.md title rep [rep ...]"""
import csv, subprocess, sys, os
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum", "launch__grid_size",
           "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


out_md, title, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
lines = [f"# {title}\n", "`ncu --set full --clock-control none`; one column per captured launch; GB/s = (dram read + write) / duration.\n"]
for rep in reps:
    hdr, units, data = raw(rep)
    kn = hdr.index("Kernel Name")
    lines += [f"## {os.path.basename(rep)}\n", "| metric | unit | " + " | ".join(str(i) for i in range(len(data))) + " |", "|---|---|" + "---|" * len(data),
              "| kernel |  | " + " | ".join(r[kn].split("(")[0][-44:].replace("|", "/") for r in data) + " |"]
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            lines.append(f"| {m} | {units[i]} | " + " | ".join(("%.6g" % num(r[i])) if num(r[i]) is not None else r[i] for r in data) + " |")
    it, ir_, iw = hdr.index("gpu__time_duration.sum"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    gbs = []
    for r in data:
        t_us = num(r[it]) * TIME.get(units[it], 1.0)
        b = num(r[ir_]) * UNIT.get(units[ir_], 1) + num(r[iw]) * UNIT.get(units[iw], 1)
        gbs.append("%.0f" % (b / t_us / 1e3))
    lines.append("| achieved DRAM traffic rate | GB/s | " + " | ".join(gbs) + " |\n")
open(out_md, "w").write("\n".join(lines) + "\n")
print(out_md, "written")
