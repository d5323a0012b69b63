"""world_size-2 gloo test (CPU) of the multi-GPU host logic: contiguous trajectory shards keyed by the global
index + one all-reduce of the un-normalised sums reproduce the single-process ensemble.  The per-shard compute
here is the oracle (tests may call it); on GPUs the same shard arithmetic drives ca_rydberg1_run."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    pkg, oracle = g.load_package(), g.load_oracle()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg, _ = pkg.get_default_configs()
    cfg.laser_noise = True
    cfg.tspan = np.linspace(0, 0.05, 3)
    lo, cnt, w = pkg.api._shard(n, True)
    assert w == world
    model = pkg.model_from_config(cfg)
    rho, rho2, _ = oracle.rydberg1_run(model, cfg.tspan, cfg.ψ0, cnt, traj_offset=lo, n_total=n, seed=5, f=cfg.f,
                                       amp_red=cfg.red_laser_phase_amplitudes, amp_blue=cfg.blue_laser_phase_amplitudes,
                                       out_flags=pkg._abi.CA_OUT_SUMS, nthreads=1)
    buf = torch.from_numpy(np.stack([rho, rho2]).view(np.float64))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    if rank == 0:
        q.put((lo, cnt, (buf.numpy().view(np.complex128) / n)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_reproduce_single_process():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    pkg, oracle = g.load_package(), g.load_oracle()
    oracle.build()
    n, world = 13, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    lo, cnt, out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert (lo, cnt) == (0, 6)
    cfg, _ = pkg.get_default_configs()
    cfg.laser_noise = True
    cfg.tspan = np.linspace(0, 0.05, 3)
    model = pkg.model_from_config(cfg)
    rho, rho2, _ = oracle.rydberg1_run(model, cfg.tspan, cfg.ψ0, n, seed=5, f=cfg.f, amp_red=cfg.red_laser_phase_amplitudes,
                                       amp_blue=cfg.blue_laser_phase_amplitudes, nthreads=1)
    assert np.abs(out[0] - rho).max() < 1e-13 and np.abs(out[1] - rho2).max() < 1e-13
