"""Turn gpurun_out/<tag>_*.ncu-rep and <tag>_launches.csv into the committed summaries under profiles/."""
import csv, subprocess, sys, json, collections, os
tag = sys.argv[1] if len(sys.argv) > 1 else "r1d"
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size",
           "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active"]
def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]
def table(rep, title, note):
    hdr, units, data = raw(rep)
    kn = hdr.index("Kernel Name")
    lines = [f"## {os.path.basename(rep)} — {title}", "", note, "", "| metric | unit | " + " | ".join(str(i) for i in range(len(data))) + " |",
             "|---|---|" + "---|" * len(data), "| Kernel Name |  | " + " | ".join(r[kn][:48].replace("|", "/") for r in data) + " |"]
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            vals = []
            for r in data:
                try: vals.append(f"{float(r[i].replace(',', '')):.6g}")
                except ValueError: vals.append(r[i])
            lines.append(f"| {m} | {units[i]} | " + " | ".join(vals) + " |")
    return "\n".join(lines) + "\n", (hdr, units, data)
md = [f"# ncu --set full captures ({tag}, `--clock-control none --import-source on`), key metrics per launch\n"]
t, (hdr, units, data) = table(f"gpurun_out/{tag}_qgemm.ncu-rep", "the four GEMMs of decoder block 0", "QKV: QUANT, o_proj: RESID, w1||w3: ACTMUL, w2: RESID/CTA-pair (batch 8 x seq 1024, TinyLlama shapes).")
md.append(t)
ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
def tobytes(v, u):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
tr = [tobytes(r[ir], units[ir]) + tobytes(r[iw], units[iw]) for r in data]
json.dump({"kernel": "qgemm_kernel", "source": f"profiles/{tag}_ncu_summary.md ({tag}_qgemm.ncu-rep, ncu --set full, block 0)",
           "bytes_per_launch": tr, "avg_bytes_per_launch": sum(tr) / len(tr)}, open(f"profiles/{tag}_qgemm_traffic.json", "w"), indent=1)
md.append(table(f"gpurun_out/{tag}_attn.ncu-rep", "qnorm, qrope, qattn (single-QK-pass kernel) of block 0", "")[0])
md.append(table(f"gpurun_out/{tag}_decode.ncu-rep", "decode step, first 12 launches of layer 0/1",
                "qnorm_row, qgemv (QKV) + epilogue, qattn_decode (cluster of 8), qgemv (o) + epilogue, qnorm_row, qgemv (w1||w3) + epilogue, qgemv (w2) + epilogue, qnorm_row (TinyLlama shapes, batch 8, position 1016).")[0])
md.append(table(f"gpurun_out/{tag}_fgemv.ncu-rep", "fp32 lm_head GEMV of the decode step", "")[0])
open(f"profiles/{tag}_ncu_summary.md", "w").write("\n".join(md))
# launch list
rows = list(csv.reader(open(f"gpurun_out/{tag}_launches.csv")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]; kn, mv, mn = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
agg = collections.OrderedDict(); n = 0
for r in rows[hi + 1:]:
    if len(r) > mv and r[mn] == "gpu__time_duration.sum":
        a = agg.setdefault(r[kn][:70], [0, 0.0]); a[0] += 1; a[1] += float(r[mv].replace(",", "")) / 1e6; n += 1
tot = sum(a[1] for a in agg.values())
out = [f"# ncu launch list of one bench step ({tag}: `ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off python bench.py --profile-step`)\n",
       f"total {tot:.3f} ms over {n} launches (cold-cache, serialised: compare shares, not absolutes; raw rows in {tag}_launches.csv)\n",
       "| kernel | launches | ms | share |", "|---|---|---|---|"]
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {c} | {t:.3f} | {100 * t / tot:.1f}% |")
open(f"profiles/{tag}_launches.md", "w").write("\n".join(out) + "\n")
print("\n".join(out[:14]))
