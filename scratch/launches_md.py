"""ncu `--metrics gpu__time_duration.sum --csv` log -> markdown table of kernel, launches, total ms, share.  usage: launches_md.py in.csv out.md "title" """
import csv, sys, collections
src, dst, title = sys.argv[1:4]
rows = []
with open(src, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rd:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", ""))
    v = v / 1e6 if r[iu] in ("ns", "nsecond") else (v / 1e3 if r[iu] in ("us", "usecond") else v)
    a = agg.setdefault(r[ik], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values()); n = sum(a[0] for a in agg.values())
with open(dst, "w") as f:
    f.write(f"# {title}\n\ntotal {tot:.3f} ms over {n} launches (ncu per-launch times are cold-cache and serialised: compare shares, not absolutes)\n\n")
    f.write("| kernel | launches | ms | share |\n|---|---|---|---|\n")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k[:80]}` | {c} | {t:.3f} | {100 * t / tot:.1f}% |\n")
print(dst, "written")
