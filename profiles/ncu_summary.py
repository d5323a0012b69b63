"""Summarise an .ncu-rep (read here, no GPU needed): key raw metrics incl. register-spill (local-memory) counters, dynamic opcode mix,
stall breakdown.  With NCU_TRAJ=<trajectories of the captured launch> and NCU_KEY=<name> the DRAM bytes per trajectory are merged into
profiles/ncu_traffic.json (what bench.py reports as roofline.traffic).
    python profiles/ncu_summary.py gpurun_out/x.ncu-rep [kernel-index]"""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2 + (int(sys.argv[2]) if len(sys.argv) > 2 else 0)]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__warps_eligible.avg.per_cycle_active", "l1tex__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "sm__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        # occupancy and register spills (local memory)
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "sm__maximum_warps_per_active_cycle_pct", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "smsp__sass_inst_executed_op_local.sum"]
for k in keys:
    if k in d:
        print(f"{k:78s} {d[k][0]:14s} {d[k][1]}")
import json, os
if os.environ.get("NCU_TRAJ") and os.environ.get("NCU_KEY"):
    tf = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traffic.json")
    tj = json.load(open(tf)) if os.path.exists(tf) else {}
    def _bytes(k):
        u, v = d[k]
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    tot_b = _bytes("dram__bytes_read.sum") + _bytes("dram__bytes_write.sum")
    tj[os.environ["NCU_KEY"]] = {"dram_bytes_per_trajectory": tot_b / float(os.environ["NCU_TRAJ"]), "dram_bytes": tot_b,
                                "trajectories": int(os.environ["NCU_TRAJ"]), "source": os.environ.get("NCU_SRC", os.path.basename(rep))}
    json.dump(tj, open(tf, "w"), indent=1)
print("--- stall reasons (warps per issue-active cycle)")
for h in hdr:
    if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h:
        v = float(d[h][1] or 0)
        if v > 0.02:
            print(f"  {h.split('stalled_')[1].split('_per')[0]:22s} {v:.3f}")
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if start:
    hdr2 = rows[start[0]]
    ix = {h: i for i, h in enumerate(hdr2)}
    data = [r for r in rows[start[0] + 1:(start[1] - 1 if len(start) > 1 else None)] if len(r) == len(hdr2)]
    f = lambda r, k: float(r[ix[k]] or 0)
    tot = sum(f(r, "Instructions Executed") for r in data)
    mix = collections.Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
        mix[m.group(2) if m else "?"] += f(r, "Instructions Executed")
    print(f"--- dynamic opcode mix ({tot:.3g} warp instructions, {len(data)} static)")
    print("  " + "  ".join(f"{op} {100 * c / tot:.1f}%" for op, c in mix.most_common(16)))
    fp64 = sum(c for op, c in mix.items() if op in ("DFMA", "DMUL", "DADD", "DSETP"))
    print(f"  fp64 share of issued instructions: {100 * fp64 / tot:.1f}%")
    mx = max(f(r, "Instructions Executed") for r in data)
    hot = [r for r in data if f(r, "Instructions Executed") > 0.3 * mx]
    print(f"  hot loop: {len(hot)} static instructions ({len(hot) * 16 / 1024:.1f} KB), {100 * sum(f(r, 'Instructions Executed') for r in hot) / tot:.1f}% of dynamic")
