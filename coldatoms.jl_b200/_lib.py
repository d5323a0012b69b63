"""ctypes binding of csrc/libcoldatoms_b200.so (the C ABI of include/coldatoms_b200.h).

There is deliberately no fallback: if the shared library is missing, or no sm_100 CUDA device is present,
every compute entry point raises.
"""
import ctypes as C
import os
import threading

import numpy as np

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("COLDATOMS_B200_LIB", os.path.join(_HERE, "csrc", "libcoldatoms_b200.so"))  # env override: A/B builds

dp = C.POINTER(C.c_double)
_lib = None
_lock = threading.Lock()

# every symbol include/coldatoms_b200.h declares
EXPORTS = [
    "ca_version", "ca_last_error", "ca_init", "ca_finalize", "ca_set_stream", "ca_reset_stream", "ca_synchronize",
    "ca_collect_stats", "ca_stats_to_device",
    "ca_init_multi", "ca_finalize_multi", "ca_multi_size", "ca_multi_context", "ca_rydberg1_run_multi", "ca_rydberg2_run_multi",
    "ca_rydberg1_sweep_multi", "ca_release_recapture_multi",
    "ca_measure_fp64_peak", "ca_rydberg1_run", "ca_rydberg2_run", "ca_rydberg2_run_ex", "ca_rydberg1_sweep", "ca_rydberg2_sweep", "ca_release_recapture",
    "ca_sample_atoms", "ca_sample_atoms_metropolis", "ca_phase_noise", "ca_eval_beam", "ca_thermal_average",
]


class ColdAtomsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"coldatoms_b200 error {code}: {msg}")
        self.code = code


def load():
    """Load the shared library (no GPU needed for loading; needed for ca_init)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(make -C coldatoms.jl_b200/csrc). There is no CPU fallback.")
            L = C.CDLL(LIB_PATH)
            L.ca_last_error.restype = C.c_char_p
            L.ca_version.restype = C.c_int
            L.ca_multi_context.restype = C.c_void_p
            _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise ColdAtomsError(rc, load().ca_last_error().decode())


def _p(a):
    return None if a is None else a.ctypes.data_as(dp)


def _arr(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _psi(psi0, n):
    v = np.ascontiguousarray(np.asarray(psi0, dtype=np.complex128)).reshape(-1)
    if v.shape[0] != n:
        raise ValueError(f"ψ0 must have {n} components")
    return v.view(np.float64)


class Context:
    """One ca_ctx: a CUDA device, a stream and the library-owned device buffers."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        _check(load().ca_init(int(device), C.byref(self._h)))
        self.device = int(device)

    def close(self):
        if self._h:
            load().ca_finalize(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream_ptr):
        """Run on a caller-owned CUDA stream handle (0 = legacy default stream, e.g. torch's default current stream)."""
        _check(load().ca_set_stream(self._h, C.c_void_p(stream_ptr)))

    def reset_stream(self):
        _check(load().ca_reset_stream(self._h))

    def synchronize(self):
        _check(load().ca_synchronize(self._h))

    def collect_stats(self, check=True):
        """Deferred ca_stats of the most recent ensemble call (synchronises).  A CA_OUT_DEVICE call cannot report a trajectory that
        hit maxiters / went non-finite in its return value: this does, raising ColdAtomsError like the synchronous call would."""
        st = _abi.ca_stats()
        rc = load().ca_collect_stats(self._h, C.byref(st))
        if check:
            _check(rc)
        return st.as_dict()

    def stats_to_device(self, dev_ptr):
        """Enqueue the export of the last call's counters as 8 doubles at device pointer `dev_ptr` (they add across shards)."""
        _check(load().ca_stats_to_device(self._h, C.cast(C.c_void_p(dev_ptr), dp)))

    def measure_fp64_peak(self):
        tf, mhz = C.c_double(), C.c_double()
        _check(load().ca_measure_fp64_peak(self._h, C.byref(tf), C.byref(mhz)))
        return tf.value, mhz.value

    # -- one-atom ensemble -------------------------------------------------------------------------------
    def rydberg1_run(self, model, tspan, psi0, n_traj, traj_offset=0, n_total=0, seed=_abi.DEFAULT_SEED, samples=None,
                     phases_red=None, phases_blue=None, f=None, amp_red=None, amp_blue=None, ode=None, out_flags=0,
                     dev_out=None):
        tspan = _arr(tspan)
        nt = len(tspan)
        samples = _arr(samples, (-1, 6)) if samples is not None else None
        f, amp_red, amp_blue = _arr(f), _arr(amp_red), _arr(amp_blue)
        nf = 0 if f is None else len(f)
        phases_red = _arr(phases_red, (-1, nf)) if phases_red is not None else None
        phases_blue = _arr(phases_blue, (-1, nf)) if phases_blue is not None else None
        for a in (samples, phases_red, phases_blue):
            if a is not None and a.shape[0] != n_traj:
                raise ValueError("host-supplied samples/phases must have n_traj rows")
        st = _abi.ca_stats()
        ode = ode if ode is not None else _abi.ode_opts()
        if out_flags & _abi.CA_OUT_DEVICE:
            prho, prho2 = C.cast(C.c_void_p(dev_out[0]), dp), C.cast(C.c_void_p(dev_out[1]), dp)
            rho = rho2 = None
        else:
            rho = np.zeros((nt, 5, 5), dtype=np.complex128)
            rho2 = np.zeros((nt, 5, 5), dtype=np.complex128)
            prho, prho2 = rho.ctypes.data_as(dp), rho2.ctypes.data_as(dp)
        _check(load().ca_rydberg1_run(self._h, C.byref(model), _p(tspan), C.c_int32(nt), _p(_psi(psi0, 5)), C.c_int64(n_traj),
                                      C.c_int64(traj_offset), C.c_int64(n_total), C.c_uint64(seed), _p(samples), _p(phases_red),
                                      _p(phases_blue), _p(f), _p(amp_red), _p(amp_blue), C.c_int32(nf), C.byref(ode),
                                      C.c_int32(out_flags), prho, prho2, C.byref(st)))
        if rho is None:
            return None, None, st.as_dict()
        # library layout is column-major per matrix; numpy result is [t, row, col]
        return rho.transpose(0, 2, 1).copy(), rho2.transpose(0, 2, 1).copy(), st.as_dict()

    def rydberg1_sweep(self, models, psi0s, t_end, n_traj_per_point, traj_offset=0, n_total=0, seed=_abi.DEFAULT_SEED, f=None,
                       amp_red=None, amp_blue=None, ode=None, out_flags=0, dev_out=None):
        """`models`: a list of ca_model, or an already marshalled ctypes array of them (a caller that repeats a sweep marshals once)."""
        n_points = len(models)
        marr = models if isinstance(models, C.Array) else (_abi.ca_model * n_points)(*models)
        psi = np.ascontiguousarray(np.asarray(psi0s, dtype=np.complex128).reshape(n_points, 5)).view(np.float64)
        t_end = _arr(t_end, (n_points,))
        f, amp_red, amp_blue = _arr(f), _arr(amp_red), _arr(amp_blue)
        nf = 0 if f is None else len(f)
        if out_flags & _abi.CA_OUT_DEVICE:
            rho = rho2 = None
            prho, prho2 = C.cast(C.c_void_p(dev_out[0]), dp), C.cast(C.c_void_p(dev_out[1]), dp)
        else:
            rho = np.zeros((n_points, 5, 5), dtype=np.complex128)
            rho2 = np.zeros((n_points, 5, 5), dtype=np.complex128)
            prho, prho2 = rho.ctypes.data_as(dp), rho2.ctypes.data_as(dp)
        st = _abi.ca_stats()
        ode = ode if ode is not None else _abi.ode_opts()
        _check(load().ca_rydberg1_sweep(self._h, marr, _p(psi), _p(t_end), C.c_int32(n_points), C.c_int64(n_traj_per_point),
                                        C.c_int64(traj_offset), C.c_int64(n_total), C.c_uint64(seed), _p(f), _p(amp_red), _p(amp_blue),
                                        C.c_int32(nf), C.byref(ode), C.c_int32(out_flags), prho, prho2, C.byref(st)))
        if rho is None:
            return None, None, st.as_dict()
        return rho.transpose(0, 2, 1).copy(), rho2.transpose(0, 2, 1).copy(), st.as_dict()

    # -- two-atom ensemble -------------------------------------------------------------------------------
    def rydberg2_run(self, model, tspan, psi0, n_traj, traj_offset=0, n_total=0, seed=_abi.DEFAULT_SEED, samples1=None,
                     samples2=None, phases4=None, f=None, amp_red=None, amp_blue=None, ode=None, out_flags=0, dev_out=None):
        tspan = _arr(tspan)
        nt = len(tspan)
        samples1 = _arr(samples1, (-1, 6)) if samples1 is not None else None
        samples2 = _arr(samples2, (-1, 6)) if samples2 is not None else None
        f, amp_red, amp_blue = _arr(f), _arr(amp_red), _arr(amp_blue)
        nf = 0 if f is None else len(f)
        ph_arr, keep = None, []
        if phases4 is not None:
            keep = [_arr(p, (-1, nf)) for p in phases4]
            ph_arr = (dp * 4)(*[_p(p) for p in keep])
        st = _abi.ca_stats()
        ode = ode if ode is not None else _abi.ode_opts()
        if out_flags & _abi.CA_OUT_DEVICE:
            prho, prho2 = C.cast(C.c_void_p(dev_out[0]), dp), C.cast(C.c_void_p(dev_out[1]), dp)
            rho = rho2 = None
        else:
            rho = np.zeros((nt, 25, 25), dtype=np.complex128)
            rho2 = np.zeros((nt, 25, 25), dtype=np.complex128)
            prho, prho2 = rho.ctypes.data_as(dp), rho2.ctypes.data_as(dp)
        _check(load().ca_rydberg2_run(self._h, C.byref(model), _p(tspan), C.c_int32(nt), _p(_psi(psi0, 25)), C.c_int64(n_traj),
                                      C.c_int64(traj_offset), C.c_int64(n_total), C.c_uint64(seed), _p(samples1), _p(samples2), ph_arr,
                                      _p(f), _p(amp_red), _p(amp_blue), C.c_int32(nf), C.byref(ode), C.c_int32(out_flags), prho, prho2,
                                      C.byref(st)))
        if rho is None:
            return None, None, st.as_dict()
        return rho.transpose(0, 2, 1).copy(), rho2.transpose(0, 2, 1).copy(), st.as_dict()

    def rydberg2_sweep(self, models, psi0s, t_end, n_traj_per_point, traj_offset=0, seed=_abi.DEFAULT_SEED, f=None, amp_red=None,
                       amp_blue=None, ode=None, out_flags=0):
        """ca_rydberg2_sweep: n_points CZ configurations in one launch; `n_traj_per_point` is an int or one count per point.
        Returns (ρ(t_end)[n_points, 25, 25], ρ²(t_end), stats)."""
        n_points = len(models)
        marr = models if isinstance(models, C.Array) else (_abi.ca_model * n_points)(*models)
        psi = np.ascontiguousarray(np.asarray(psi0s, dtype=np.complex128).reshape(n_points, 25)).view(np.float64)
        t_end = _arr(t_end, (n_points,))
        cnt = np.ascontiguousarray(np.broadcast_to(np.asarray(n_traj_per_point, dtype=np.int64), (n_points,)))
        f, amp_red, amp_blue = _arr(f), _arr(amp_red), _arr(amp_blue)
        nf = 0 if f is None else len(f)
        rho = np.zeros((n_points, 25, 25), dtype=np.complex128)
        rho2 = np.zeros((n_points, 25, 25), dtype=np.complex128)
        st = _abi.ca_stats()
        ode = ode if ode is not None else _abi.ode_opts()
        _check(load().ca_rydberg2_sweep(self._h, marr, _p(psi), _p(t_end), C.c_int32(n_points), cnt.ctypes.data_as(C.POINTER(C.c_int64)),
                                        C.c_int64(traj_offset), C.c_uint64(seed), _p(f), _p(amp_red), _p(amp_blue), C.c_int32(nf),
                                        C.byref(ode), C.c_int32(out_flags), rho.ctypes.data_as(dp), rho2.ctypes.data_as(dp), C.byref(st)))
        return rho.transpose(0, 2, 1).copy(), rho2.transpose(0, 2, 1).copy(), st.as_dict()

    def rydberg2_run_ex(self, model, tspan, psi0=None, rho0=None, n_traj=1, nt_seg1=0, traj_offset=0, n_total=0, seed=_abi.DEFAULT_SEED,
                        samples1=None, samples2=None, phases4=None, f=None, amp_red=None, amp_blue=None, ode=None, out_flags=0):
        """ca_rydberg2_run_ex: general initial state (`rho0[25, 25]` as numpy [row, col], or `psi0`), optional two time segments
        (`tspan[:nt_seg1]`, `tspan[nt_seg1:]`), legacy two-atom flags in the model."""
        tspan = _arr(tspan)
        nt = len(tspan)
        samples1 = _arr(samples1, (-1, 6)) if samples1 is not None else None
        samples2 = _arr(samples2, (-1, 6)) if samples2 is not None else None
        f, amp_red, amp_blue = _arr(f), _arr(amp_red), _arr(amp_blue)
        nf = 0 if f is None else len(f)
        ph_arr, keep = None, []
        if phases4 is not None:
            keep = [_arr(p, (-1, nf)) for p in phases4]
            ph_arr = (dp * 4)(*[_p(p) for p in keep])
        r0 = None if rho0 is None else np.ascontiguousarray(np.asarray(rho0, dtype=np.complex128).reshape(25, 25).T).view(np.float64)
        p0 = None if psi0 is None else _psi(psi0, 25)
        st = _abi.ca_stats()
        ode = ode if ode is not None else _abi.ode_opts()
        rho = np.zeros((nt, 25, 25), dtype=np.complex128)
        rho2 = np.zeros((nt, 25, 25), dtype=np.complex128)
        _check(load().ca_rydberg2_run_ex(self._h, C.byref(model), _p(tspan), C.c_int32(nt), C.c_int32(nt_seg1), _p(p0), _p(r0),
                                         C.c_int64(n_traj), C.c_int64(traj_offset), C.c_int64(n_total), C.c_uint64(seed), _p(samples1),
                                         _p(samples2), ph_arr, _p(f), _p(amp_red), _p(amp_blue), C.c_int32(nf), C.byref(ode),
                                         C.c_int32(out_flags), rho.ctypes.data_as(dp), rho2.ctypes.data_as(dp), C.byref(st)))
        return rho.transpose(0, 2, 1).copy(), rho2.transpose(0, 2, 1).copy(), st.as_dict()

    # -- release and recapture ---------------------------------------------------------------------------
    def release_recapture(self, trap, atom, tspan, n, traj_offset=0, n_total=0, eps=1e-3, seed=_abi.DEFAULT_SEED, samples=None,
                          out_flags=0, want_indicators=False, dev_out=None):
        trap, atom, tspan = _arr(trap, (3,)), _arr(atom, (2,)), _arr(tspan)
        samples = _arr(samples, (-1, 6)) if samples is not None else None
        nt = len(tspan)
        ind = np.zeros((n, nt), dtype=np.uint8) if want_indicators else None
        if out_flags & _abi.CA_OUT_DEVICE:
            probs, pp = None, C.cast(C.c_void_p(dev_out), dp)
        else:
            probs = np.zeros(nt)
            pp = _p(probs)
        st = _abi.ca_stats()
        _check(load().ca_release_recapture(self._h, _p(trap), _p(atom), _p(tspan), C.c_int32(nt), C.c_int64(n), C.c_int64(traj_offset),
                                           C.c_int64(n_total), C.c_double(eps), C.c_uint64(seed), _p(samples), C.c_int32(out_flags),
                                           pp, None if ind is None else ind.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(st)))
        return probs, ind, st.as_dict()

    # -- samplers ----------------------------------------------------------------------------------------
    def sample_atoms(self, trap, atom, n, traj_offset=0, seed=_abi.DEFAULT_SEED, stream=0, eps=1e-2):
        trap, atom = _arr(trap, (3,)), _arr(atom, (2,))
        out = np.zeros((n, 6))
        att = C.c_int64()
        _check(load().ca_sample_atoms(self._h, _p(trap), _p(atom), C.c_int64(n), C.c_int64(traj_offset), C.c_uint64(seed),
                                      C.c_int32(stream), C.c_double(eps), _p(out), C.byref(att)))
        return out, att.value

    def sample_atoms_metropolis(self, trap, atom, n, freq=10, skip=1000, harmonic=False, eps=1e-2, n_chains=0, seed=_abi.DEFAULT_SEED, stream=0):
        trap, atom = _arr(trap, (3,)), _arr(atom, (2,))
        out = np.zeros((n, 6))
        acc = C.c_double()
        _check(load().ca_sample_atoms_metropolis(self._h, _p(trap), _p(atom), C.c_int64(n), C.c_int64(freq), C.c_int64(skip),
                                                 C.c_int32(int(bool(harmonic))), C.c_double(eps), C.c_int64(n_chains), C.c_uint64(seed),
                                                 C.c_int32(stream), _p(out), C.byref(acc)))
        return out, acc.value

    def phase_noise(self, tnodes, f, amp, n, traj_offset=0, seed=_abi.DEFAULT_SEED, stream=_abi.STREAM_RED1, phases=None):
        tnodes, f, amp = _arr(tnodes), _arr(f), _arr(amp)
        phases = _arr(phases, (n, len(f))) if phases is not None else None
        out = np.zeros((n, len(tnodes)))
        _check(load().ca_phase_noise(self._h, _p(tnodes), C.c_int32(len(tnodes)), _p(f), _p(amp), C.c_int32(len(f)), C.c_int64(n),
                                     C.c_int64(traj_offset), C.c_uint64(seed), C.c_int32(stream), _p(phases), _p(out)))
        return out

    def thermal_average(self, kind, model, n, traj_offset=0, seed=_abi.DEFAULT_SEED, eps=1e-2, samples1=None, samples2=None):
        """Device-reduced SUMS of the calibrate_two_photon / get_blockade_stark_shift_factor functionals (ca_thermal_average)."""
        samples1 = _arr(samples1, (n, 6)) if samples1 is not None else None
        samples2 = _arr(samples2, (n, 6)) if samples2 is not None else None
        sums = np.zeros(2)
        att = C.c_int64()
        _check(load().ca_thermal_average(self._h, C.c_int32(kind), C.byref(model), C.c_int64(n), C.c_int64(traj_offset), C.c_uint64(seed),
                                         C.c_double(eps), _p(samples1), _p(samples2), _p(sums), C.byref(att)))
        return sums, att.value

    def eval_beam(self, beam, xyz):
        xyz = _arr(xyz, (-1, 3))
        out = np.zeros((len(xyz), 2))
        _check(load().ca_eval_beam(self._h, C.byref(beam), _p(xyz), C.c_int64(len(xyz)), _p(out)))
        return out[:, 0] + 1j * out[:, 1]


class MultiContext:
    """ca_multi: several GPUs driven from this one process (ca_init_multi: one context per device + one NCCL clique).  The ensemble
    methods mirror Context's; trajectories are sharded by contiguous global-index ranges inside the library and combined with one
    ncclAllReduce, so the results equal the single-GPU ones up to summation order."""

    def __init__(self, devices=None):
        self._h = C.c_void_p()
        if devices is None:
            _check(load().ca_init_multi(None, C.c_int32(0), C.byref(self._h)))
        else:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            _check(load().ca_init_multi(arr, C.c_int32(len(devices)), C.byref(self._h)))
        self.size = int(load().ca_multi_size(self._h))

    def close(self):
        if self._h:
            load().ca_finalize_multi(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def rydberg1_run(self, model, tspan, psi0, n_traj, traj_offset=0, n_total=0, seed=_abi.DEFAULT_SEED, samples=None,
                     phases_red=None, phases_blue=None, f=None, amp_red=None, amp_blue=None, ode=None, out_flags=0):
        tspan = _arr(tspan)
        nt = len(tspan)
        samples = _arr(samples, (-1, 6)) if samples is not None else None
        f, amp_red, amp_blue = _arr(f), _arr(amp_red), _arr(amp_blue)
        nf = 0 if f is None else len(f)
        phases_red = _arr(phases_red, (-1, nf)) if phases_red is not None else None
        phases_blue = _arr(phases_blue, (-1, nf)) if phases_blue is not None else None
        st = _abi.ca_stats()
        ode = ode if ode is not None else _abi.ode_opts()
        rho = np.zeros((nt, 5, 5), dtype=np.complex128)
        rho2 = np.zeros((nt, 5, 5), dtype=np.complex128)
        _check(load().ca_rydberg1_run_multi(self._h, C.byref(model), _p(tspan), C.c_int32(nt), _p(_psi(psi0, 5)), C.c_int64(n_traj),
                                            C.c_int64(traj_offset), C.c_int64(n_total), C.c_uint64(seed), _p(samples), _p(phases_red),
                                            _p(phases_blue), _p(f), _p(amp_red), _p(amp_blue), C.c_int32(nf), C.byref(ode),
                                            C.c_int32(out_flags), rho.ctypes.data_as(dp), rho2.ctypes.data_as(dp), C.byref(st)))
        return rho.transpose(0, 2, 1).copy(), rho2.transpose(0, 2, 1).copy(), st.as_dict()

    def rydberg2_run(self, model, tspan, psi0, n_traj, traj_offset=0, n_total=0, seed=_abi.DEFAULT_SEED, samples1=None,
                     samples2=None, phases4=None, f=None, amp_red=None, amp_blue=None, ode=None, out_flags=0):
        tspan = _arr(tspan)
        nt = len(tspan)
        samples1 = _arr(samples1, (-1, 6)) if samples1 is not None else None
        samples2 = _arr(samples2, (-1, 6)) if samples2 is not None else None
        f, amp_red, amp_blue = _arr(f), _arr(amp_red), _arr(amp_blue)
        nf = 0 if f is None else len(f)
        ph_arr, keep = None, []
        if phases4 is not None:
            keep = [_arr(p, (-1, nf)) for p in phases4]
            ph_arr = (dp * 4)(*[_p(p) for p in keep])
        st = _abi.ca_stats()
        ode = ode if ode is not None else _abi.ode_opts()
        rho = np.zeros((nt, 25, 25), dtype=np.complex128)
        rho2 = np.zeros((nt, 25, 25), dtype=np.complex128)
        _check(load().ca_rydberg2_run_multi(self._h, C.byref(model), _p(tspan), C.c_int32(nt), _p(_psi(psi0, 25)), C.c_int64(n_traj),
                                            C.c_int64(traj_offset), C.c_int64(n_total), C.c_uint64(seed), _p(samples1), _p(samples2),
                                            ph_arr, _p(f), _p(amp_red), _p(amp_blue), C.c_int32(nf), C.byref(ode), C.c_int32(out_flags),
                                            rho.ctypes.data_as(dp), rho2.ctypes.data_as(dp), C.byref(st)))
        return rho.transpose(0, 2, 1).copy(), rho2.transpose(0, 2, 1).copy(), st.as_dict()

    def rydberg1_sweep(self, models, psi0s, t_end, n_traj_per_point, traj_offset=0, n_total=0, seed=_abi.DEFAULT_SEED, f=None,
                       amp_red=None, amp_blue=None, ode=None, out_flags=0):
        n_points = len(models)
        marr = (_abi.ca_model * n_points)(*models)
        psi = np.ascontiguousarray(np.asarray(psi0s, dtype=np.complex128).reshape(n_points, 5)).view(np.float64)
        t_end = _arr(t_end, (n_points,))
        f, amp_red, amp_blue = _arr(f), _arr(amp_red), _arr(amp_blue)
        nf = 0 if f is None else len(f)
        rho = np.zeros((n_points, 5, 5), dtype=np.complex128)
        rho2 = np.zeros((n_points, 5, 5), dtype=np.complex128)
        st = _abi.ca_stats()
        ode = ode if ode is not None else _abi.ode_opts()
        _check(load().ca_rydberg1_sweep_multi(self._h, marr, _p(psi), _p(t_end), C.c_int32(n_points), C.c_int64(n_traj_per_point),
                                              C.c_int64(traj_offset), C.c_int64(n_total), C.c_uint64(seed), _p(f), _p(amp_red),
                                              _p(amp_blue), C.c_int32(nf), C.byref(ode), C.c_int32(out_flags), rho.ctypes.data_as(dp),
                                              rho2.ctypes.data_as(dp), C.byref(st)))
        return rho.transpose(0, 2, 1).copy(), rho2.transpose(0, 2, 1).copy(), st.as_dict()

    def release_recapture(self, trap, atom, tspan, n, traj_offset=0, n_total=0, eps=1e-3, seed=_abi.DEFAULT_SEED, samples=None, out_flags=0):
        trap, atom, tspan = _arr(trap, (3,)), _arr(atom, (2,)), _arr(tspan)
        samples = _arr(samples, (-1, 6)) if samples is not None else None
        nt = len(tspan)
        probs = np.zeros(nt)
        st = _abi.ca_stats()
        _check(load().ca_release_recapture_multi(self._h, _p(trap), _p(atom), _p(tspan), C.c_int32(nt), C.c_int64(n), C.c_int64(traj_offset),
                                                 C.c_int64(n_total), C.c_double(eps), C.c_uint64(seed), _p(samples), C.c_int32(out_flags),
                                                 _p(probs), C.byref(st)))
        return probs, None, st.as_dict()


_default_ctx = {}


def default_context(device=None):
    """Process-wide context on `device` (default: LOCAL_RANK or 0)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    ctx = _default_ctx.get(device)
    if ctx is None:
        ctx = _default_ctx[device] = Context(device)
    return ctx
