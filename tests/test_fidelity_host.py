"""CPU tests of the host-side fidelity helpers (src/fidelity.jl, src/gates.jl mirrors): no compute calls."""
import numpy as np


def test_gates_and_parity_scan_closed_form(pkg):
    p = pkg
    assert np.allclose(p.Hadamard @ p.Hadamard, np.eye(5)) and np.allclose(p.RZ(0.7) @ p.RZ(-0.7), np.eye(5))
    assert np.allclose(p.X[:2, :2], [[0, 1], [1, 0]]) and np.allclose(p.Y @ p.Y, np.eye(5)) and p.Z[2, 2] == 1
    assert np.allclose(p.RX(np.pi)[:2, :2], -1j * p.X[:2, :2]) and np.allclose(p.RY(np.pi)[:2, :2], [[0, -1], [1, 0]])
    rng = np.random.default_rng(0)
    A = rng.normal(size=(25, 25)) + 1j * rng.normal(size=(25, 25))
    rho = A @ A.conj().T
    rho /= np.trace(rho)
    ph = np.array([0.0, 0.3, 1.7, 4.0, 6.2])
    F = p.parity_scan(rho, ph)
    Had = p.op_tensor(p.Id, p.Hadamard)
    Phi = (p.tensor(p.ket_0, p.ket_0) + p.tensor(p.ket_1, p.ket_1)) / np.sqrt(2)
    for a, f in zip(ph, F):   # literal form of src/fidelity.jl:166
        G = Had @ p.op_tensor(p.RZ(a), p.RZ(a))
        assert abs(f - np.real(np.conj(Phi) @ G @ rho @ G.conj().T @ Phi)) < 1e-14
        assert abs(f - p.get_parity_fidelity_temp(rho, a)[0]) < 1e-14
    # projector on the two-qubit subspace keeps the 4x4 block in tensor order
    blk = p.project_on_qubit(rho, 2)
    idx = [0, 1, 5, 6]
    assert blk.shape == (4, 4) and np.allclose(blk, rho[np.ix_(idx, idx)])
    # a Bell state has unit parity fidelity at the right phase
    bell = np.outer(Phi, Phi.conj())
    rot = Had.conj().T @ bell @ Had   # state that the Hadamard maps onto Φ+
    assert abs(p.parity_scan(rot, [0.0])[0] - 1.0) < 1e-14


def test_fidelity_configs_follow_reference(pkg):
    cfg, _ = pkg.get_default_configs()
    c = pkg.get_rydberg_fidelity_configs(cfg, 7)
    assert list(c) == ["Intermdeiate state decay", "Rydberg state decay", "Laser noise", "Atom motion", "Total"]
    assert c["Laser noise"].free_motion is False and c["Laser noise"].laser_noise and c["Laser noise"].atom_params[1] == 1.0
    assert c["Atom motion"].atom_params[1] == cfg.atom_params[1] and c["Atom motion"].n_samples == 7
    assert c["Rydberg state decay"].n_samples == 1 and c["Rydberg state decay"].spontaneous_decay_rydberg
    assert cfg.atom_params[1] == 100.0   # deepcopy: the caller's config is untouched
    assert len(pkg.basis_fidelity_states) == 6
