"""Multi-GPU paths on the GPU box: (1) the torch.distributed (NCCL) route of the public API under torchrun, on min(2, #GPUs) ranks
(with one GPU the one-rank group still goes through the device-resident sums + all-reduce + status-tail code); (2) the
single-process route of the C ABI (ca_init_multi + ca_*_multi: ncclCommInitAll inside the library) against one context."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def test_torchrun_distributed_api_matches_single_rank_and_propagates_status():
    n = min(2, _n_gpus())
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert f"DIST-OK world={n}" in r.stdout


@pytest.mark.parametrize("devices", ["one", "all"])
def test_c_abi_multi_gpu_matches_single_context(pkg, ctx, devices):
    n_dev = _n_gpus()
    if devices == "all" and n_dev < 2:
        pytest.skip("needs >= 2 GPUs for the NCCL clique")
    mc = pkg._lib.MultiContext([0] if devices == "one" else None)
    assert mc.size == (1 if devices == "one" else n_dev)
    cfg, cz = pkg.get_default_configs()
    cfg.laser_noise = True
    cfg.tspan = np.linspace(0.0, 0.25, 6)
    model = pkg.model_from_config(cfg)
    kw = dict(f=cfg.f, amp_red=cfg.red_laser_phase_amplitudes, amp_blue=cfg.blue_laser_phase_amplitudes, seed=31)
    n = 3000 + 11
    ref, ref2, rst = ctx.rydberg1_run(model, cfg.tspan, cfg.ψ0, n, **kw)
    rho, rho2, st = mc.rydberg1_run(model, cfg.tspan, cfg.ψ0, n, **kw)
    assert max(np.abs(rho - ref).max(), np.abs(rho2 - ref2).max()) < 1e-10
    assert (st["n_traj"], st["n_rhs"], st["n_accept"], st["n_sample_attempts"]) == (n, rst["n_rhs"], rst["n_accept"], rst["n_sample_attempts"])
    with pytest.raises(pkg._lib.ColdAtomsError) as e:
        mc.rydberg1_run(model, cfg.tspan, cfg.ψ0, n, ode=pkg._abi.ode_opts(maxiters=50), **kw)
    assert e.value.code == pkg._abi.CA_ERR_MAXITERS
    # host-supplied samples are sliced per device
    rng = np.random.default_rng(1)
    smp = rng.normal(size=(n, 6)) * [0.1, 0.1, 0.5, 0.05, 0.05, 0.05]
    a, a2, _ = ctx.rydberg1_run(model, cfg.tspan, cfg.ψ0, n, samples=smp, **kw)
    b, b2, _ = mc.rydberg1_run(model, cfg.tspan, cfg.ψ0, n, samples=smp, **kw)
    assert max(np.abs(a - b).max(), np.abs(a2 - b2).max()) < 1e-10
    # two atoms
    cz.laser_noise = True
    cz.tspan = np.array([0.0, 0.01, 0.02])
    m2 = pkg.model_from_config(cz, two_atom=True)
    kw2 = dict(f=cz.f, amp_red=cz.red_laser_phase_amplitudes, amp_blue=cz.blue_laser_phase_amplitudes, seed=9)
    zr, zr2, _ = ctx.rydberg2_run(m2, cz.tspan, cz.ψ0, 77, **kw2)
    zm, zm2, _ = mc.rydberg2_run(m2, cz.tspan, cz.ψ0, 77, **kw2)
    assert max(np.abs(zm - zr).max(), np.abs(zm2 - zr2).max()) < 1e-10
    # sweep sharded by grid point (no collective): identical to the single-context sweep
    models, psis, tends = [], [], []
    for k in range(7):
        c = cfg.copy()
        c.detuning_params[1] += 2.0 * (k - 3)
        models.append(pkg.model_from_config(c))
        psis.append(pkg.ket_1)
        tends.append(0.05 + 0.03 * k)
    s1, s12, _ = ctx.rydberg1_sweep(models, psis, tends, 40, **kw)
    sm, sm2, sst = mc.rydberg1_sweep(models, psis, tends, 40, **kw)
    assert max(np.abs(sm - s1).max(), np.abs(sm2 - s12).max()) < 1e-12 and sst["n_traj"] == 7 * 40
    # release and recapture
    t = np.arange(0.0, 51.0)
    trap, atom = [1000.0, 1.0, np.pi / 0.852], [86.9091835, 50.0]
    p1, _, _ = ctx.release_recapture(trap, atom, t, 10001, seed=8)
    pm, _, _ = mc.release_recapture(trap, atom, t, 10001, seed=8)
    assert np.array_equal(p1, pm)
    mc.close()
