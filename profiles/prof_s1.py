"""S1 (BASELINE config 5): detuning x temperature x pulse-time sweep of the 1-atom Rabi model, 100 x 10 x 10 = 10^4 grid points x
10^4 trajectories = 10^8 trajectories, grid points sharded across the ranks (no data-path collective; the per-point results are
gathered at the end).  Each rank issues ONE ca_rydberg1_sweep launch for its points.
    torchrun --nproc-per-node 8 profiles/prof_s1.py [traj_per_point=10000] [n_delta=100]
    python profiles/prof_s1.py 1000            (one GPU, 1/10 of the trajectories per point)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

pkg = g.load_package()
n_pp = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
nd = int(sys.argv[2]) if len(sys.argv) > 2 else 100
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg, _ = pkg.get_default_configs()
cfg.laser_noise = True
T0 = cfg.tspan[20]
nT, nt = 10, 10
δs = cfg.detuning_params[1] + 2 * np.pi * np.linspace(-2, 2, nd)
Ts = np.linspace(10.0, 100.0, nT)
ts = np.linspace(2 * T0 / nt, 2 * T0, nt)
grid = [(δ, T, te) for δ in δs for T in Ts for te in ts]
# interleaved hand-out: every rank gets every pulse time and temperature (the cost of a point scales with its pulse time)
mine = list(range(rank, len(grid), world))
models, psis, tends = [], [], []
for i in mine:
    δ, T, te = grid[i]
    c = cfg.copy()
    c.detuning_params[1] = float(δ)
    c.atom_params[1] = float(T)
    models.append(pkg.model_from_config(c))
    psis.append(pkg.ket_1)
    tends.append(float(te))
ctx = pkg._lib.default_context(local)
kw = dict(f=cfg.f, amp_red=cfg.red_laser_phase_amplitudes, amp_blue=cfg.blue_laser_phase_amplitudes)
for rep in range(2):   # warm-up, then the timed pass
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ev0.record()
    rho, rho2, st = ctx.rydberg1_sweep(models, psis, tends, n_pp, traj_offset=0, seed=7 + rep, **kw)
    ev1.record()
    torch.cuda.synchronize()
    ctx.reset_stream()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda", dtype=torch.float64)
    pr = torch.zeros(len(grid), device="cuda", dtype=torch.float64)
    pr[mine] = torch.as_tensor(rho[:, 2, 2].real.copy(), device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(pr, op=dist.ReduceOp.SUM)   # disjoint supports: a gather of the Rydberg populations of all grid points
    if rank == 0:
        n = len(grid) * n_pp
        print("s1 ranks %d points %d traj/point %d trajectories %.3g max-rank ms %.1f traj/s %.4g  P_r range %.3f..%.3f rhs/traj(rank0) %.0f"
              % (world, len(grid), n_pp, n, ms.item(), n / ms.item() * 1e3, pr.min().item(), pr.max().item(), st["n_rhs"] / (len(mine) * n_pp)), flush=True)
if world > 1:
    dist.destroy_process_group()
