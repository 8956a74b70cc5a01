"""Host-side .qsim reader: gate set sanity (no GPU, no reference needed) and, in the
build container, the structure of the reference's Sycamore circuit files."""

import os

import numpy as np
import pytest

from cotengra_b200.circuits import amplitude_network, gate_matrix, read_qsim


def test_gates_are_unitary_and_square_roots():
    X = np.array([[0, 1], [1, 0]], dtype=complex)
    Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
    W = (X + Y) / np.sqrt(2)
    for name, params, target in (("x_1_2", (), X), ("y_1_2", (), Y), ("hz_1_2", (), W)):
        U = gate_matrix(name, params)
        assert np.allclose(U @ U.conj().T, np.eye(2))
        S = U @ U
        phase = S[np.nonzero(np.abs(target) > 0.5)][0] / target[np.nonzero(np.abs(target) > 0.5)][0]
        assert np.allclose(S, phase * target)  # square root up to a global phase
    for t in (0.3, -1.7):
        U = gate_matrix("rz", (t,))
        assert np.allclose(U @ U.conj().T, np.eye(2))
    U = gate_matrix("fs", (1.5157741664069029, 0.5567125777723744))
    assert np.allclose(U @ U.conj().T, np.eye(4))
    with pytest.raises(ValueError):
        gate_matrix("cz", ())


@pytest.mark.reference
def test_sycamore_m10_network_structure():
    path = "/root/reference/examples/circuit_n53_m10_s0_e0_pABCDCDAB.qsim"
    n, gates = read_qsim(path)
    assert n == 53 and len(gates) == 1658
    inputs, output, size_dict, arrays = amplitude_network(path)
    assert len(inputs) == 1658 + 2 * 53 == len(arrays)
    assert output == () and set(size_dict.values()) == {2}
    # every index appears exactly twice (closed network)
    counts = {}
    for t in inputs:
        for ix in t:
            counts[ix] = counts.get(ix, 0) + 1
    assert set(counts.values()) == {2}
    assert all(a.shape == (2,) * len(t) for a, t in zip(arrays, inputs))
