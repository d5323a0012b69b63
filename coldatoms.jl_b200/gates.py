"""Qubit gates embedded in the 5-level basis (src/gates.jl:1-35): host-side 5x5 matrices acting on {0,1}."""
import numpy as np

from .config import tensor


def get_gate(U):
    """src/gates.jl:1-5: identity on (r,p,l), U on the qubit block."""
    op = np.eye(5, dtype=np.complex128)
    op[:2, :2] = np.asarray(U, dtype=np.complex128)
    return op


def project_on_qubit(ρ, n=1):
    """src/gates.jl:8-12: P ρ P† with P the projector on span{0,1}^{⊗n} (returns the 2^n x 2^n block)."""
    P1 = np.zeros((2, 5), dtype=np.complex128)
    P1[0, 0] = P1[1, 1] = 1.0
    P = P1
    for _ in range(n - 1):
        P = np.kron(P1, P)   # first factor fastest, as in tensor()
    return P @ np.asarray(ρ) @ P.conj().T


Id = np.eye(5, dtype=np.complex128)
Hadamard = get_gate(np.array([[1, 1], [1, -1]]) / np.sqrt(2))
X = get_gate([[0, 1], [1, 0]])
Y = get_gate([[0, -1j], [1j, 0]])
Z = get_gate([[1, 0], [0, -1]])


def RX(θ):
    return get_gate([[np.cos(θ / 2), -1j * np.sin(θ / 2)], [-1j * np.sin(θ / 2), np.cos(θ / 2)]])


def RY(θ):
    return get_gate([[np.cos(θ / 2), -np.sin(θ / 2)], [np.sin(θ / 2), np.cos(θ / 2)]])


def RZ(θ):
    return get_gate([[np.exp(-1j * θ / 2), 0], [0, np.exp(1j * θ / 2)]])


def op_tensor(A, B):
    """Operator `A ⊗ B` in the QuantumOptics convention (first factor = fastest index)."""
    return tensor(A, B)
