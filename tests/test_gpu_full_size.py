"""Size-independent properties at BASELINE.json's full ensemble sizes (the oracle cannot follow there): trace, Hermiticity,
positivity of populations, Jensen's inequality between the two outputs, shard additivity, agreement with a small ensemble."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _props(rho, rho2, n):
    tr = np.trace(rho, axis1=1, axis2=2)
    assert np.abs(tr - 1.0).max() < 1e-9                                  # every trajectory conserves the trace
    assert np.abs(rho - rho.conj().transpose(0, 2, 1)).max() < 1e-12      # Hermitian
    assert np.abs(rho2 - rho2.conj().transpose(0, 2, 1)).max() < 1e-12
    p = np.real(np.einsum("tii->ti", rho))
    assert p.min() > -1e-12 and p.max() < 1 + 1e-12
    # second output = mean of ρ·ρ: (ρ²)_ii = Σ_k |ρ_ik|² >= ρ_ii², and the mean of a square dominates the square of the mean
    p2 = np.real(np.einsum("tii->ti", rho2))
    assert np.all(p2 + 1e-12 >= p ** 2)
    assert np.all(np.real(np.trace(rho2, axis1=1, axis2=2)) <= 1 + 1e-9)  # purity <= 1


def test_r1_million_trajectories(pkg, ctx):
    cfg, _ = pkg.get_default_configs()
    cfg.laser_noise = True
    n = 10 ** 6
    model = pkg.model_from_config(cfg)
    kw = dict(f=cfg.f, amp_red=cfg.red_laser_phase_amplitudes, amp_blue=cfg.blue_laser_phase_amplitudes, seed=11)
    rho, rho2, st = ctx.rydberg1_run(model, cfg.tspan, cfg.ψ0, n, **kw)
    assert st["n_traj"] == n and st["n_reject"] >= 0 and st["status"] == 0
    _props(rho, rho2, n)
    # two shards of the same ensemble add up to it (what the multi-GPU path relies on)
    S = pkg._abi.CA_OUT_SUMS
    a, a2, _ = ctx.rydberg1_run(model, cfg.tspan, cfg.ψ0, 400_000, traj_offset=0, n_total=n, out_flags=S, **kw)
    b, b2, _ = ctx.rydberg1_run(model, cfg.tspan, cfg.ψ0, 600_000, traj_offset=400_000, n_total=n, out_flags=S, **kw)
    assert np.abs((a + b) / n - rho).max() < 1e-10 and np.abs((a2 + b2) / n - rho2).max() < 1e-10
    # a 2 % sub-ensemble agrees within its Monte-Carlo error (populations fluctuate by < 0.1 between trajectories)
    small, _, _ = ctx.rydberg1_run(model, cfg.tspan, cfg.ψ0, 20_000, traj_offset=0, n_total=20_000, **kw)
    assert np.abs(np.einsum("tii->ti", small - rho)).max() < 5e-3
    # Rabi contrast survives the ensemble average: the Rydberg population peaks near t = T0/2 = tspan[10]
    assert np.real(rho[10, 2, 2]) > 0.9 and np.real(rho[20, 2, 2]) < 0.1


def test_z1_full_ensemble(pkg, ctx):
    _, cz = pkg.get_default_configs()
    cz.laser_noise = True
    n = 100_000
    model = pkg.model_from_config(cz, two_atom=True)
    kw = dict(f=cz.f, amp_red=cz.red_laser_phase_amplitudes, amp_blue=cz.blue_laser_phase_amplitudes, seed=5, ode=pkg._abi.ode_opts(maxiters=10 ** 7))
    rho, rho2, st = ctx.rydberg2_run(model, cz.tspan, cz.ψ0, n, **kw)
    assert st["status"] == 0 and st["n_traj"] == n
    _props(rho, rho2, n)
    assert np.abs(rho[0] - np.outer(cz.ψ0, np.conj(cz.ψ0))).max() < 1e-12   # first output point is ρ0
    # block structure: no coherence between an l level of one atom and a non-l level of the same atom
    r = rho[-1].reshape(5, 5, 5, 5)   # [a2, a1, b2, b1]
    for lv in range(4):
        assert np.abs(r[:, 4, :, lv]).max() == 0 and np.abs(r[4, :, lv, :]).max() == 0
    # a 5 % sub-ensemble (different hand-out order, same trajectories) agrees within Monte-Carlo error
    small, _, _ = ctx.rydberg2_run(model, cz.tspan, cz.ψ0, 5_000, **kw)
    assert np.abs(np.einsum("tii->ti", small - rho)).max() < 1e-2
