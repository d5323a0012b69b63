"""Physics-level known answers printed by the reference's OWN run: notebooks/modeling.ipynb cells 7-8 fit
0.5 (1 - exp(-p2 t) cos(p1 t)) to the Rydberg population returned by `simulation_blue_intens` (10 thermal samples, flat-top HG / LG
blue beams, parameters of params/default.jl) and print p2/p1:
    HG_0 = 0.002617771802773333   HG_2 = 0.00208206724227503   HG_4 = 0.0020756389407290445   HG_6 = 0.0020758985883022893
    LG_0 = 0.002617771802769474   LG_2 = 0.0020976855137380666 LG_4 = 0.0029374829788899226   LG_6 = 0.006288147252059915
    LG_8 = 0.012576329415650022
The 10 samples of that run (Julia RNG) and the exact temperature / waists of the notebook session are not recoverable, so the thermal
part of these numbers cannot be matched.  What IS parameter-robust: a flat-top beam removes the position dependence, so HG_4 / HG_6 sit
just above the T -> 0 limit, which depends only on Ωr, Ωb, Δ0, Γ and the decay branching of the model -- conventions a factor-2 slip
in any of which would move by 2-4x.  The oracle (CPU) and the CUDA path (GPU) must reproduce that limit, and the ordering with beam order."""
import numpy as np
import pytest
from scipy.optimize import curve_fit

NOTEBOOK = {"HG_2": 0.00208206724227503, "HG_4": 0.0020756389407290445, "HG_6": 0.0020758985883022893,
            "LG_2": 0.0020976855137380666, "LG_4": 0.0029374829788899226, "LG_6": 0.006288147252059915, "LG_8": 0.012576329415650022}


def _setup(pkg):
    """params/default.jl:1-54 and notebooks/modeling.ipynb cell 4."""
    wr, wb = 100.0, 2.0
    zr, zb = pkg.w0_to_z0(wr, 0.795), pkg.w0_to_z0(wb, 0.475)
    atom, trap = [86.9091835, 70.0], [1000.0, 1.1, pkg.w0_to_z0(1.1, 0.82)]
    D0, Ob, Or, G = 2 * np.pi * 1500, 2 * np.pi * 96, 2 * np.pi * 96, 2 * np.pi * 6.0
    det = [D0, pkg.δ_twophoton(Or, Ob, D0)]
    T0 = pkg.T_twophoton(Or, Ob, D0)
    return dict(atom=atom, trap=trap, red=[Or, wr, zr], Ob=Ob, wb=wb, zb=zb, det=det, decay=[G / 4, 3 * G / 4],
                tspan=np.arange(0, 6 * T0 + 1e-12, T0 / 50))


def _blue(a, kind, i):
    s = 1 if i == 0 else i
    if kind == "HG":
        return [a["Ob"], a["wb"], a["zb"], "simp_flattop_HG", i, i, 1.0]          # cell 7
    return [a["Ob"], a["wb"] / s ** 0.5, a["zb"] / s, "simp_flattop_LG", i, i]   # cell 8


def _ratio(tspan, Pr):
    model = lambda x, p1, p2: 0.5 * (1 - np.exp(-p2 * x) * np.cos(p1 * x))   # noqa: E731  (cell 7: `model`, p0 = [18, 1])
    p, _ = curve_fit(model, tspan, Pr, p0=[18.0, 1.0])
    return p[1] / p[0]


def test_oracle_reproduces_the_notebook_flat_top_limit(pkg, oracle):
    a = _setup(pkg)

    def run(kind, i, samples):
        m = pkg.blue_intens_model(a["atom"], a["trap"], a["red"], _blue(a, kind, i), a["det"], a["decay"])
        rho, _, _ = oracle.rydberg1_run(m, a["tspan"], pkg.ket_1, len(samples), samples=samples)
        return _ratio(a["tspan"], rho[:, 2, 2].real)

    cold = run("HG", 6, np.zeros((1, 6)))
    # the reference's flat-top values sit 0.5 % above the T -> 0 limit of the restated model
    assert abs(cold / NOTEBOOK["HG_6"] - 1) < 0.01 and abs(cold / NOTEBOOK["HG_4"] - 1) < 0.01
    assert cold < NOTEBOOK["HG_6"]
    # with thermal samples at the file's 70 uK the ratio rises (Doppler + residual position dependence): the notebook value is bracketed
    warm = [run("HG", 4, oracle.sample_atoms(a["trap"], a["atom"], 10, seed=100 + k)[0]) for k in range(3)]
    assert cold < NOTEBOOK["HG_4"] < max(warm)
    # ordering with the LG order (narrower effective waist wb/sqrt(s), zb/s: position noise grows), as in cell 8
    lg = {i: np.mean([run("LG", i, oracle.sample_atoms(a["trap"], a["atom"], 10, seed=200 + k)[0]) for k in range(3)]) for i in (2, 4, 8)}
    assert lg[2] < lg[4] < lg[8] and NOTEBOOK["LG_2"] < NOTEBOOK["LG_4"] < NOTEBOOK["LG_8"]
    assert abs(lg[2] / NOTEBOOK["LG_2"] - 1) < 0.25


@pytest.mark.gpu
def test_cuda_path_reproduces_the_notebook_flat_top_limit(pkg, ctx):
    a = _setup(pkg)
    for kind, i in (("HG", 4), ("HG", 6)):
        rho, _ = pkg.simulation_blue_intens(a["tspan"], pkg.ket_1, a["atom"], a["trap"], np.zeros((1, 6)), a["red"], _blue(a, kind, i), a["det"],
                                            a["decay"])
        r = _ratio(a["tspan"], rho[:, 2, 2].real)
        assert abs(r / NOTEBOOK["%s_%d" % (kind, i)] - 1) < 0.01
