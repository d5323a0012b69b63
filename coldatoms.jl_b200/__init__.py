"""coldatoms_b200 -- B200-native drop-in for the Monte-Carlo master-equation hot path of
mgoloshchapov/ColdAtoms.jl (module NeutralAtoms).

Host-side mirror of the reference's Julia API (same names, argument meaning and return shapes) above the
C ABI of `csrc/libcoldatoms_b200.so` (include/coldatoms_b200.h).  The directory is called
`coldatoms.jl_b200`; because of the dot it is loaded under the module name `coldatoms_b200`
(see `__graft_entry__.load_package`).  There is no CPU fallback: every compute entry point raises if the
CUDA library or a CUDA device is missing.
"""
from . import _abi  # noqa: F401
from .physics import *  # noqa: F401,F403
from .physics import Sϕ, ϕ_amplitudes, Ω_twophoton, T_twophoton, δ_twophoton, Ωr_required  # noqa: F401
from .config import (RydbergConfig, CZLPConfig, get_default_configs, model_from_config, beam_from_params,  # noqa: F401
                     ket_0, ket_1, ket_r, ket_p, ket_l, tensor, nlevelstate)
from .api import *  # noqa: F401,F403
from .gates import get_gate, project_on_qubit, Id, Hadamard, X, Y, Z, RX, RY, RZ, op_tensor  # noqa: F401
from .fidelity import (basis_fidelity_states, get_rydberg_fidelity_configs, get_rydberg_infidelity, get_parity_fidelity,  # noqa: F401
                       get_parity_fidelity_temp, get_cz_infidelity, parity_scan, cz_infidelity_points)
from .legacy import two_atom_simulation, direct_CZ_simulation, legacy_two_atom_model, legacy_direct_model  # noqa: F401
