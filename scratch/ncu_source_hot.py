"""ncu report -> where the warp-stall samples of a kernel sit in its SASS stream.  usage: ncu_source_hot.py report.ncu-rep [min_pct]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
h = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rd = csv.reader(io.StringIO("\n".join(lines[h:])))
hdr = next(rd)
isrc, isamp, iinst = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
rows = []
for r in rd:
    if len(r) <= isamp:
        continue
    try:
        rows.append((r[isrc].strip(), float(r[isamp] or 0), float(r[iinst] or 0)))
    except ValueError:
        pass
tot = sum(s for _, s, _ in rows) or 1
print(f"{len(rows)} SASS instructions, {tot:.0f} samples")
cum = 0.0
marks = ("BAR", "UCGABAR", "LDG", "LDS", "STG", "STS", "MUFU", "SHFL", "EXIT", "BRA", "ATOM", "RED", "LD.E", "LDC")
for i, (src, s, n) in enumerate(rows):
    cum += s
    if 100 * s / tot >= min_pct:
        print(f"#{i:5d}  {100 * s / tot:5.1f}%  cum {100 * cum / tot:5.1f}%  exec {n:8.0f}  {src[:110]}")
