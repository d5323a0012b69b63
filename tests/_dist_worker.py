"""Worker of tests/test_gpu_multi.py: runs under `python -m torch.distributed.run` (one rank per GPU, NCCL) and checks the
distributed=True path of the public API against the single-rank result computed in the same process."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    pkg = g.load_package()
    cfg, cz = pkg.get_default_configs()
    cfg.laser_noise = True
    cfg.tspan = np.linspace(0.0, 0.25, 6)
    cfg.n_samples = 4096 + 37          # ragged over the ranks
    ref, ref2 = pkg.simulation(cfg, seed=21, device=local)                      # this rank alone, whole ensemble
    rho, rho2 = pkg.simulation(cfg, seed=21, device=local, distributed=True)    # sharded + one NCCL all-reduce
    st = pkg.last_stats()
    assert st["n_traj"] == cfg.n_samples and st["n_rhs"] > 0, st
    e1 = max(np.abs(rho - ref).max(), np.abs(rho2 - ref2).max())
    assert e1 < 1e-10, e1
    # a trajectory that hits maxiters on ANY rank raises on EVERY rank (it used to be dropped silently on this path)
    for fn, c in ((pkg.simulation, cfg),):
        try:
            fn(c, seed=21, device=local, distributed=True, maxiters=50)
            raise SystemExit("maxiters=50 did not raise on the distributed path")
        except pkg._lib.ColdAtomsError as e:
            assert e.code == pkg._abi.CA_ERR_MAXITERS, e.code
    # only the LAST rank's shard fails (maxiters is per rank here): every rank must still see the error
    try:
        pkg.simulation(cfg, seed=21, device=local, distributed=True, maxiters=(50 if rank == world - 1 else 10 ** 5))
        raise SystemExit("a failure on one rank was not propagated")
    except pkg._lib.ColdAtomsError as e:
        assert e.code == pkg._abi.CA_ERR_MAXITERS
    # two-atom path
    cz.laser_noise = True
    cz.tspan = np.array([0.0, 0.01, 0.02])
    cz.n_samples = 64 + 5
    zr, zr2 = pkg.simulation_czlp(cz, seed=4, device=local)
    zd, zd2 = pkg.simulation_czlp(cz, seed=4, device=local, distributed=True)
    e2 = max(np.abs(zd - zr).max(), np.abs(zd2 - zr2).max())
    assert e2 < 1e-10, e2
    try:
        pkg.simulation_czlp(cz, seed=4, device=local, distributed=True, maxiters=20)
        raise SystemExit("two-atom maxiters did not raise on the distributed path")
    except pkg._lib.ColdAtomsError as e:
        assert e.code == pkg._abi.CA_ERR_MAXITERS
    # release and recapture: counts all-reduced; harmonic sampler is partition independent, the Metropolis one statistically
    t = np.arange(0.0, 51.0)
    trap, atom = [1000.0, 1.0, np.pi / 0.852], [86.9091835, 50.0]
    p1, _ = pkg.release_recapture(t, trap, atom, 20001, seed=8, device=local)
    pd, _ = pkg.release_recapture(t, trap, atom, 20001, seed=8, device=local, distributed=True)
    assert np.array_equal(p1, pd)
    pm, acc = pkg.release_recapture(t, trap, atom, 20000, harmonic=False, seed=8, device=local, distributed=True, skip=200)
    assert 0.05 < acc < 0.95 and np.abs(pm - p1).max() < 0.05 and pm[0] > 0.99
    if rank == 0:
        print(f"DIST-OK world={world} e1={e1:.2e} e2={e2:.2e}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
