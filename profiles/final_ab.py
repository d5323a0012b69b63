"""One-call A/B of per-translation-unit compiler variants (profiles/dev_k3.sh / dev_k4.sh builds, e.g. `-Xptxas -regUsageLevel=0`) against the
objects of the full build, run on the GPU box:

    python profiles/final_ab.py gpurun_out/final_ab.json

For every (translation unit, workload) pair below the base library and the variant library are timed alternately (profiles/ab.py child
processes, best kernel time of three launches each, two rounds); a variant WINS when it is > 0.5 % faster and its rho(t) agrees with the base
library's to 1e-9.  The winners' objects are linked with the remaining base objects into variants/cand.so (used by the rest of
profiles/r2_final.sh for the GPU test-suite and the bench lines), and the decision table is written as JSON."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "coldatoms.jl_b200", "csrc", "build")
BASE = os.path.join(ROOT, "coldatoms.jl_b200", "csrc", "libcoldatoms_b200.so")
# (base object, variant name under variants/, AB_WORKLOAD, trajectories, rounds)
CASES = [
    ("l1_9_1", "l1_9_1_ru0", "r1", 151552, 2),       # R1: four whole waves of 148 x 256
    ("l1_15_1", "l1_15_1_ru0", "ns15", 132608, 2),   # four whole waves of 148 x 224
    ("ca_lindblad2", "k4_ru0", "z1", 23680, 2),      # twenty trajectories per warp
    ("l1_9_2", "l1_9_2_ru0", "r2", 75776, 1),
    ("l1_21_0", "l1_21_0_ru0", "ns21", 37888, 1),
]


def run(lib, workload, n, tag):
    out = "/tmp/fab_%s.npy" % tag
    env = dict(os.environ, COLDATOMS_B200_LIB=lib, AB_WORKLOAD=workload)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "ab.py"), "--child", str(n), out], env=env, capture_output=True, text=True, timeout=240)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
    if not line.startswith("kernel_ms"):
        return None, None, (r.stderr or "")[-300:]
    return float(line.split()[1]), np.load(out), line


def link(out, objs):
    cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64", "-shared", "-o", out] + objs + ["-ldl"]
    return subprocess.run(cmd, capture_output=True, text=True)


def objects(replace):   # the objects of the full build with {base object name: variant name} substituted
    objs = []
    for f in sorted(os.listdir(BUILD)):
        if f.endswith(".o"):
            objs.append(os.path.join(ROOT, "variants", replace[f[:-2]] + ".o") if f[:-2] in replace else os.path.join(BUILD, f))
    return objs


table, winners = [], []
for base_obj, var, workload, n, rounds in CASES:
    vlib = os.path.join(ROOT, "variants", var + ".so")
    if not os.path.exists(vlib) and os.path.exists(vlib[:-3] + ".o"):   # only the variant OBJECTS travel to the box (a library is 48 MB)
        link(vlib, objects({base_obj: var}))
    if not os.path.exists(vlib):
        table.append({"unit": base_obj, "variant": var, "error": "variant library missing"})
        continue
    tb, tv, dev, note = [], [], None, ""
    for r in range(rounds):
        t0, rho0, l0 = run(BASE, workload, n, "base")
        t1, rho1, l1 = run(vlib, workload, n, "var")
        if t0 is None or t1 is None:
            note = "FAILED: %s | %s" % (l0, l1)
            break
        tb.append(t0); tv.append(t1)
        dev = float(np.abs(rho0 - rho1).max())
    row = {"unit": base_obj, "variant": var, "workload": workload, "trajectories": n, "base_ms": tb, "variant_ms": tv, "max_drho": dev, "note": note}
    if tb and tv and not note:
        gain = min(tb) / min(tv) - 1.0
        row["gain"] = gain
        row["win"] = bool(gain > 0.005 and dev is not None and dev < 1e-9)
        if row["win"]:
            winners.append((base_obj, var))
    table.append(row)
    print(json.dumps(row), flush=True)

cand = os.path.join(ROOT, "variants", "cand.so")
r = link(cand, objects(dict(winners)))
res = {"table": table, "winners": [v for _, v in winners], "link_rc": r.returncode, "link_err": r.stderr[-400:]}
if r.returncode != 0 or not winners:   # nothing won (or no linker on the box): the candidate IS the base library
    import shutil
    shutil.copyfile(BASE, cand)
with open(sys.argv[1], "w") as fh:
    json.dump(res, fh, indent=1)
print("winners:", res["winners"], "link rc", r.returncode, flush=True)
