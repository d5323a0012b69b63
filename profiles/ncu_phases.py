"""Stall samples ALONG the hot loop: the SASS of the kernel is cut into windows of `win` instructions (only instructions executed at
least `xmin` times the innermost count), and per window the share of all samples, the issue/stall split and marker opcodes are printed.
    python profiles/ncu_phases.py x.ncu-rep [win] [xmin]"""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]; win = int(sys.argv[2]) if len(sys.argv) > 2 else 64; xmin = float(sys.argv[3]) if len(sys.argv) > 3 else 0.25
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
f = lambda r, k: float(r[ix[k]] or 0)
mx = max(f(r, "Instructions Executed") for r in data)
tot = sum(f(r, "# Samples") for r in data)
hot = [(i, r) for i, r in enumerate(data) if f(r, "Instructions Executed") / mx >= xmin]
print(f"hot instructions {len(hot)}, samples in them {100 * sum(f(r, '# Samples') for _, r in hot) / tot:.1f}%")
cum = 0.0
for w in range(0, len(hot), win):
    blk = hot[w:w + win]
    s = sum(f(r, "# Samples") for _, r in blk)
    cum += s
    ops = collections.Counter(re.sub(r"@!?U?P\d\s+", "", r[ix["Source"]]).split()[0].split(".")[0] for _, r in blk)
    fp64 = ops["DFMA"] + ops["DMUL"] + ops["DADD"]
    marks = " ".join(f"{k}{v}" for k, v in ops.items() if k in ("MUFU", "LDS", "STS", "BRA", "BAR", "F2F", "VOTE", "LDG", "STL", "LDL", "DSETP", "BSSY", "SHFL", "ATOM", "RED") and v)
    st = lambda k: 100 * sum(f(r, k) for _, r in blk) / max(s, 1)
    print(f"{blk[0][0]:6d}-{blk[-1][0]:6d} smp={100 * s / tot:5.2f}% cum={100 * cum / tot:5.1f}% per_inst={s / len(blk) / (tot / len(hot)):4.2f} fp64={fp64:3d} "
          f"wait={st('stall_wait'):3.0f} ssb={st('stall_short_sb'):3.0f} math={st('stall_math'):3.0f} sel={st('stall_selected'):3.0f} nsel={st('stall_not_selected'):3.0f} "
          f"br={st('stall_branch_resolving'):3.0f} noi={st('stall_no_inst'):3.0f} lsb={st('stall_long_sb'):3.0f} | {marks}")
