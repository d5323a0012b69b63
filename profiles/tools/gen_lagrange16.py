"""Generates the 16-point (degree 15) Lagrange weight table of ca_device.cuh (CA_LAGRANGE16_INIT): rows for coarsening factors
c = 8, 16, 32, fine offsets r = 1..c-1, columns = coarse nodes at offsets -7..8; exact rational arithmetic, correctly rounded."""
from fractions import Fraction
import sys

P = 16
nodes = list(range(-(P // 2 - 1), P // 2 + 1))
rows = []
for c in (8, 16, 32):
    for r in range(1, c):
        x = Fraction(r, c)
        w = []
        for i in nodes:
            num, den = Fraction(1), Fraction(1)
            for j in nodes:
                if j != i:
                    num *= (x - j); den *= (i - j)
            w.append(num / den)
        assert sum(w) == 1
        rows.append(w)
out = ["#define CA_LAGRANGE16_INIT {{ \\"]
for k, w in enumerate(rows):
    out.append("    {" + ", ".join(repr(float(v)) for v in w) + "}" + ("," if k + 1 < len(rows) else "") + " \\")
out.append("}}")
print("\n".join(out))
print("// rows: %d" % len(rows), file=sys.stderr)
