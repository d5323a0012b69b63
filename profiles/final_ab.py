"""One-call A/B of per-translation-unit compiler variants (profiles/dev_k3.sh / dev_k4.sh builds, e.g. `-Xptxas -regUsageLevel=0`) against the
objects of the full build, run on the GPU box:

    python profiles/final_ab.py gpurun_out/final_ab.json

For every translation unit below the base library and its variant libraries are timed in turn on the unit's workload (profiles/ab.py child
processes, best kernel time of three launches each, one or two rounds); a variant WINS when it is > 0.5 % faster and its rho(t) agrees with
the base library's to 1e-9 (the best one per unit is taken).  The winners' objects are linked with the remaining base objects into variants/cand.so (used by the rest of
profiles/r2_final.sh for the GPU test-suite and the bench lines), and the decision table is written as JSON."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "coldatoms.jl_b200", "csrc", "build")
BASE = os.path.join(ROOT, "coldatoms.jl_b200", "csrc", "libcoldatoms_b200.so")
# base object: (AB_WORKLOAD, trajectories, rounds, [variant names under variants/]); second pass (the first, `-regUsageLevel=0` for every unit,
# is profiles/r2/final_ab.json: only NS = 15 free flight gained and is now built that way)
CASES = {
    "ca_lindblad2": ("z1", 23680, 1, ["k4_ru10"]),                                # twenty trajectories per warp
    "l1_9_1": ("r1", 151552, 2, ["l1_9_1_ru10", "l1_9_1_noconst"]),               # R1: four whole waves of 148 x 256
    "l1_15_1": ("ns15", 132608, 1, ["l1_15_1_ru10"]),                             # four whole waves of 148 x 224
    "l1_9_2": ("r2", 75776, 1, ["l1_9_2_ru10"]),
    "l1_21_0": ("ns21", 37888, 1, ["l1_21_0_ru10"]),
    "l1_9_3": ("trap", 75776, 1, ["l1_9_3_ru0", "l1_9_3_ru10"]),
    "l1_15_0": ("ns15t", 37888, 1, ["l1_15_0_ru0", "l1_15_0_ru10"]),
}

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
for base_obj, (workload, n, rounds, variants) in CASES.items():
    libs = {}
    for var in variants:   # only the variant OBJECTS travel to the box (a library is 48 MB)
        vlib = os.path.join(ROOT, "variants", var + ".so")
        if os.path.exists(vlib[:-3] + ".o") and link(vlib, objects({base_obj: var})).returncode == 0:
            libs[var] = vlib
        else:
            table.append({"unit": base_obj, "variant": var, "error": "variant object missing or link failed"})
    times = {k: [] for k in ["base"] + list(libs)}
    devs, notes, rho0 = {}, {}, None
    for r in range(rounds):
        for k in times:
            t, rho, line = run(BASE if k == "base" else libs[k], workload, n, k)
            if t is None:
                notes[k] = "FAILED: %s" % line
                continue
            times[k].append(t)
            if k == "base":
                rho0 = rho
            elif rho0 is not None:
                devs[k] = float(np.abs(rho - rho0).max())
    best = None
    for var in libs:
        row = {"unit": base_obj, "variant": var, "workload": workload, "trajectories": n, "base_ms": times["base"], "variant_ms": times[var],
               "max_drho": devs.get(var), "note": notes.get(var, notes.get("base", ""))}
        if times["base"] and times[var] and var in devs:
            row["gain"] = min(times["base"]) / min(times[var]) - 1.0
            row["win"] = bool(row["gain"] > 0.005 and devs[var] < 1e-9)
            if row["win"] and (best is None or row["gain"] > best[1]):
                best = (var, row["gain"])
        table.append(row)
        print(json.dumps(row), flush=True)
    if best:
        winners.append((base_obj, best[0]))

cand = os.path.join(ROOT, "variants", "cand.so")
r = link(cand, objects(dict(winners)))
res = {"table": table, "winners": [v for _, v in winners], "link_rc": r.returncode, "link_err": r.stderr[-400:]}
if r.returncode != 0 or not winners:   # nothing won (or no linker on the box): the candidate IS the base library
    import shutil
    shutil.copyfile(BASE, cand)
with open(sys.argv[1], "w") as fh:
    json.dump(res, fh, indent=1)
print("winners:", res["winners"], "link rc", r.returncode, flush=True)
