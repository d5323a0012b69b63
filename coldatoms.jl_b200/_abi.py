"""ctypes mirror of include/coldatoms_b200.h (POD structs and constants of the C ABI).

Kept free of any library loading so that the test-only oracle wrapper (oracle/oracle.py) can share the
struct layouts without importing the product library.
"""
import ctypes as C

CA_OK = 0
CA_ERR_ARG, CA_ERR_CUDA, CA_ERR_NODEVICE, CA_ERR_MAXITERS = -1, -2, -3, -4
CA_ERR_NONFINITE, CA_ERR_UNSUPPORTED, CA_ERR_NOMEM = -5, -6, -7

CA_BEAM_GAUSS, CA_BEAM_FLATTOP_HG, CA_BEAM_FLATTOP_LG, CA_BEAM_UNIFORM = 0, 1, 2, 3
CA_DOPPLER_XZ, CA_DOPPLER_Z = 0, 1
CA_SIGN_RYDBERG1, CA_SIGN_CZ = 0, 1
CA_COMPAT_VZ_FROM_VY = 1
CA_COMPAT_DEFAULT = CA_COMPAT_VZ_FROM_VY
CA_RHO2_MATSQ, CA_RHO2_ABS2 = 0, 1
CA_OUT_MEAN, CA_OUT_SUMS, CA_OUT_DEVICE = 0, 1, 2
CA_AVG_TWOPHOTON, CA_AVG_RM6 = 0, 1
CA_2A_PER_ATOM_LASERS, CA_2A_CONST_VRR, CA_2A_SHIFT_AFTER_MOTION, CA_2A_DIRECT = 1, 2, 4, 8
CA_ERR_NCCL = -8

# Philox stream ids (shared by the device sampler and the oracle)
STREAM_ATOM1, STREAM_ATOM2, STREAM_RED1, STREAM_BLUE1, STREAM_RED2, STREAM_BLUE2 = range(6)

DEFAULT_SEED = 0xC01DA705


class ca_beam(C.Structure):
    _fields_ = [
        ("Omega0", C.c_double), ("w0", C.c_double), ("z0", C.c_double), ("theta", C.c_double),
        ("n", C.c_double), ("sqz", C.c_double),
        ("kind", C.c_int32), ("hg_n", C.c_int32), ("hg_m", C.c_int32), ("lg_l", C.c_int32),
        ("_pad", C.c_int32),
    ]


class ca_model(C.Structure):
    _fields_ = [
        ("atom_m", C.c_double), ("atom_T", C.c_double),
        ("trap_U0", C.c_double), ("trap_w0", C.c_double), ("trap_z0", C.c_double),
        ("red", ca_beam), ("blue", ca_beam),
        ("Delta0", C.c_double), ("delta0", C.c_double),
        ("Gamma0", C.c_double), ("Gamma1", C.c_double), ("Gammal", C.c_double), ("Gammar", C.c_double),
        ("atom_motion", C.c_int32), ("free_motion", C.c_int32), ("laser_noise", C.c_int32),
        ("decay_intermediate", C.c_int32), ("decay_rydberg", C.c_int32),
        ("doppler_kind", C.c_int32), ("parallel", C.c_int32), ("detuning_sign", C.c_int32),
        ("compat", C.c_int32), ("rho2_mode", C.c_int32), ("n_noise_nodes", C.c_int32), ("_pad", C.c_int32),
        ("center1", C.c_double * 3), ("center2", C.c_double * 3),
        ("c6", C.c_double), ("xi", C.c_double),
        ("two_atom_flags", C.c_int32), ("_pad2", C.c_int32),
        ("red2", ca_beam), ("blue2", ca_beam),
        ("Delta0_2", C.c_double), ("delta0_2", C.c_double), ("V_rr", C.c_double), ("seg2_phase", C.c_double),
    ]


class ca_ode_opts(C.Structure):
    _fields_ = [
        ("reltol", C.c_double), ("abstol", C.c_double), ("dt", C.c_double), ("dtmax", C.c_double),
        ("maxiters", C.c_int64), ("adaptive", C.c_int32), ("dense", C.c_int32),
    ]


class ca_stats(C.Structure):
    _fields_ = [
        ("n_traj", C.c_int64), ("n_rhs", C.c_int64), ("n_accept", C.c_int64), ("n_reject", C.c_int64),
        ("n_sample_attempts", C.c_int64),
        ("kernel_ms", C.c_double), ("total_ms", C.c_double), ("noise_ms", C.c_double),
        ("n_launches", C.c_int32), ("status", C.c_int32),
        ("n_exact_steps", C.c_int64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def ode_opts(reltol=1e-6, abstol=1e-8, dt=0.0, dtmax=0.0, maxiters=100000, adaptive=True, dense=True):
    """`ode_kwargs...` of the reference (rydberg_model.jl:197,255) mapped onto ca_ode_opts."""
    return ca_ode_opts(float(reltol), float(abstol), float(dt), float(dtmax), int(maxiters),
                       1 if adaptive else 0, 1 if dense else 0)
