"""CPU tests: the C-ABI library loads and exports every symbol the header declares; host-side marshalling and
helpers; the product path fails loudly without a GPU (no CPU fallback)."""
import math
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(pkg):
    hdr = open(os.path.join(ROOT, "include", "coldatoms_b200.h")).read()
    declared = set(re.findall(r"\b(ca_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"ca_ctx"}
    assert declared == set(pkg._lib.EXPORTS)
    lib = pkg._lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ca_version() == 200


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg._lib.ColdAtomsError) as e:
        pkg._lib.Context(0)
    assert e.value.code == pkg._abi.CA_ERR_NODEVICE
    cfg, _ = pkg.get_default_configs()
    with pytest.raises(pkg._lib.ColdAtomsError):
        pkg.simulation(cfg)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "coldatoms.jl_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".jl")):
                src = open(os.path.join(dirpath, fn)).read()
                bad = re.findall(r"import\s+oracle|from\s+oracle|load_oracle|libcoldatoms_oracle|coldatoms_oracle\.h|\bco_[a-z]", src)
                assert not bad, (fn, bad)


def test_struct_layouts_match_header(pkg):
    import ctypes as C
    abi = pkg._abi
    assert C.sizeof(abi.ca_beam) == 6 * 8 + 5 * 4 + 4  # 72 with tail padding
    assert C.sizeof(abi.ca_ode_opts) == 48
    assert C.sizeof(abi.ca_stats) == 5 * 8 + 3 * 8 + 8 + 8
    assert C.sizeof(abi.ca_model) == 5 * 8 + 2 * 72 + 6 * 8 + 12 * 4 + 6 * 8 + 16 + 8 + 2 * 72 + 4 * 8


def test_default_configs_match_reference_numbers(pkg):
    """Derived R1 / Z1 parameters quoted in SURVEY App. C from default.jl."""
    cfg, cz = pkg.get_default_configs()
    assert abs(cfg.red_laser_params[0] - 397.3835) < 1e-3
    assert abs(cfg.tspan[20] - 0.5) < 1e-12 and len(cfg.tspan) == 101 and abs(cfg.tspan[-1] - 2.5) < 1e-12
    assert abs(cfg.trap_params[2] - 3.8642) < 1e-4
    assert abs(cfg.red_laser_params[2] - 9879.22) < 1e-2 and abs(cfg.blue_laser_params[2] - 661.388) < 1e-3
    assert len(cfg.f) == 397 and abs(cfg.f[-1] - 1.0) < 1e-12
    amp = np.asarray(cfg.red_laser_phase_amplitudes)
    assert abs(amp.min() - 5.1e-4) < 1e-5 and abs(amp.max() - 5.39e-2) < 1e-4
    assert abs(math.sqrt(np.sum(amp ** 2) / 2) - 0.0942) < 1e-4
    wr, wz = pkg.trap_frequencies(cfg.atom_params, cfg.trap_params)
    assert abs(wr - 0.61861) < 1e-5 and abs(wz - 0.43742) < 1e-5
    assert abs(cz.tspan[-1] / 2 - 0.341601) < 1e-6
    assert cz.ψ0.shape == (25,) and abs(np.vdot(cz.ψ0, cz.ψ0) - 1) < 1e-15
    assert abs(cz.ψ0[0 + 5 * 1] - 0.5) < 1e-15  # |0>_1 |1>_2 at index a1 + 5 a2
    assert abs(pkg.vconst - 0.09118367518929328) < 1e-17


def test_model_marshalling(pkg):
    abi = pkg._abi
    cfg, cz = pkg.get_default_configs()
    m = pkg.model_from_config(cfg)
    assert m.red.kind == abi.CA_BEAM_GAUSS and abs(m.red.theta - math.pi / 2) < 1e-15 and m.blue.theta == 0.0
    assert m.detuning_sign == abi.CA_SIGN_RYDBERG1 and m.compat == abi.CA_COMPAT_DEFAULT and m.n_noise_nodes == 1001
    m2 = pkg.model_from_config(cz, two_atom=True)
    assert m2.detuning_sign == abi.CA_SIGN_CZ and list(m2.center1) == [-1.0, 0.0, 0.0] and abs(m2.xi - 3.90242) < 1e-12
    b = pkg.beam_from_params([1.0, 2.0, 3.0, "simp_flattop_LG", 4, 0])
    assert b.kind == abi.CA_BEAM_FLATTOP_LG and b.lg_l == 4
    b = pkg.beam_from_params([1.0, 2.0, 3.0, "simp_flattop_HG", 0, 0, 1.0])
    assert b.kind == abi.CA_BEAM_GAUSS  # n = m = 0 -> TEM00 (rydberg_model.jl:66-67)
    with pytest.raises(ValueError):
        pkg.beam_from_params([1.0, 2.0, 3.0, "nonsense", 2, 2])  # the reference returns `nothing` here (App. B9)
    with pytest.raises(TypeError):
        pkg.api._ode({"saveat": 1})


def test_host_helpers(pkg):
    assert abs(pkg.w0_to_z0(1.0, 0.852) - math.pi / 0.852) < 1e-15
    assert abs(pkg.Ω_twophoton(3.0, 4.0, 10.0) - 0.6) < 1e-15 and abs(pkg.δ_twophoton(3.0, 4.0, 10.0) + 7 / 40) < 1e-15
    assert abs(pkg.T_twophoton(3.0, 4.0, 10.0) - 2 * math.pi / 0.6) < 1e-14
    x = np.linspace(-1, 1, 5)
    assert np.allclose(pkg.I(x, 0, 0, 1.0, 3.0), np.exp(-2 * x ** 2))
    e = pkg.E(0.3, 0.1, 2.0, 1.0, 3.0)
    assert abs(abs(e) - pkg.A(0.3, 0.1, 2.0, 1.0, 3.0)) < 1e-15
    rho = np.zeros((3, 5, 5), complex)
    rho[:, 1, 1] = 0.5
    pr = pkg.get_rydberg_probs(rho, rho @ rho)
    assert np.allclose(pr["P1"], 0.5) and np.allclose(pr["S1"], math.sqrt(0.25 - 0.25 + 1e-12) / 3)
