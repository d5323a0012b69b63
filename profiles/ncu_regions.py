"""Stall samples by code region (classified by execution count relative to the innermost loop) and top stall sites.
    python profiles/ncu_regions.py x.ncu-rep [ntop]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
f = lambda r, k: float(r[ix[k]] or 0)
mx = max(f(r, "Instructions Executed") for r in data)
tot = sum(f(r, "# Samples") for r in data); toti = sum(f(r, "Instructions Executed") for r in data)
cls = [("x>=0.9", 0.9, 2), ("0.25-0.9", 0.25, 0.9), ("0.05-0.25", 0.05, 0.25), ("0.002-0.05", 0.002, 0.05), ("rest", -1, 0.002)]
for name, lo, hi in cls:
    sel = [r for r in data if lo <= f(r, "Instructions Executed") / mx < hi]
    sh = lambda k: 100 * sum(f(r, k) for r in sel) / tot
    print(f"{name:12s} static={len(sel):6d} samples={sh('# Samples'):5.1f}% dyn={100 * sum(f(r, 'Instructions Executed') for r in sel) / toti:5.1f}% "
          f"short_sb={sh('stall_short_sb'):5.1f}% wait={sh('stall_wait'):5.1f}% long_sb={sh('stall_long_sb'):4.1f}% math={sh('stall_math'):4.1f}% sel={sh('stall_selected'):5.1f}% notsel={sh('stall_not_selected'):4.1f}% noinst={sh('stall_no_inst'):4.1f}% br={sh('stall_branch_resolving'):4.1f}%")
top = sorted(enumerate(data), key=lambda t: -f(t[1], "# Samples"))[:ntop]
for i, r in sorted(top):
    print(f"{i:6d} x={f(r, 'Instructions Executed') / mx:6.3f} s={100 * f(r, '# Samples') / tot:5.2f}% ssb={int(f(r, 'stall_short_sb')):6d} wait={int(f(r, 'stall_wait')):6d} lsb={int(f(r, 'stall_long_sb')):5d} {r[ix['Source']][:64]:64s} <- {data[i - 1][ix['Source']][:44]}")
