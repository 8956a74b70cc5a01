"""``.qsim`` circuit files -> amplitude tensor networks (host utility).

cotengra's Sycamore examples (``examples/circuit_n53_m*.qsim``; BASELINE configs 3
and 4) are turned into tensor networks by quimb in the reference notebooks; quimb
is not part of the execution path and is absent here, so this is a minimal
reader for the gate set those files use:

    x_1_2 = sqrt(X), y_1_2 = sqrt(Y), hz_1_2 = sqrt(W) with W = (X+Y)/sqrt(2),
    rz(theta) = diag(e^{-i theta/2}, e^{+i theta/2}),
    fs(theta, phi) = fSim.

The amplitude <bits| U |0...0> becomes: one |0> vector per qubit, one tensor per
gate (rank 2 or 4, index order (out..., in...)), one <bit| vector per qubit, no
open index.  Gate conventions only fix global phases of the amplitude; parity
against the oracle holds for any values.
"""

from __future__ import annotations

import numpy as np

from .tree import get_symbol


def gate_matrix(name, params):
    s = 1.0 / np.sqrt(2.0)
    if name == "x_1_2":
        return s * np.array([[1, -1j], [-1j, 1]], dtype=np.complex128)
    if name == "y_1_2":
        return s * np.array([[1, -1], [1, 1]], dtype=np.complex128)
    if name == "hz_1_2":
        return np.array([[s, -(1 + 1j) / 2], [(1 - 1j) / 2, s]], dtype=np.complex128)
    if name == "rz":
        (t,) = params
        return np.diag([np.exp(-0.5j * t), np.exp(0.5j * t)]).astype(np.complex128)
    if name == "fs":
        t, p = params
        c, sn = np.cos(t), -1j * np.sin(t)
        return np.array(
            [[1, 0, 0, 0], [0, c, sn, 0], [0, sn, c, 0], [0, 0, 0, np.exp(-1j * p)]],
            dtype=np.complex128,
        )
    raise ValueError(f"unknown gate {name!r}")


def read_qsim(path):
    """-> (n_qubits, [(name, qubits, params)])"""
    with open(path) as f:
        lines = [ln.split() for ln in f if ln.strip()]
    n = int(lines[0][0])
    gates = []
    for tok in lines[1:]:
        name = tok[1]
        nq = 2 if name == "fs" else 1
        qubits = tuple(int(q) for q in tok[2:2 + nq])
        params = tuple(float(x) for x in tok[2 + nq:])
        gates.append((name, qubits, params))
    return n, gates


def amplitude_network(path, bits=None, dtype="complex128"):
    """Tensor network of one amplitude: ``(inputs, output, size_dict, arrays)``
    with single-character index labels (cotengra convention)."""
    n, gates = read_qsim(path)
    bits = [0] * n if bits is None else [int(b) for b in bits]
    counter = [0]

    def new_ix():
        ix = get_symbol(counter[0])
        counter[0] += 1
        return ix

    wire = [new_ix() for _ in range(n)]
    inputs, arrays = [], []
    zero = np.array([1, 0], dtype=dtype)
    for q in range(n):
        inputs.append((wire[q],))
        arrays.append(zero.copy())
    for name, qubits, params in gates:
        U = gate_matrix(name, params).astype(dtype)
        if len(qubits) == 1:
            (q,) = qubits
            out = new_ix()
            inputs.append((out, wire[q]))
            arrays.append(U)
            wire[q] = out
        else:
            q0, q1 = qubits
            o0, o1 = new_ix(), new_ix()
            inputs.append((o0, o1, wire[q0], wire[q1]))
            arrays.append(U.reshape(2, 2, 2, 2))
            wire[q0], wire[q1] = o0, o1
    for q in range(n):
        v = np.zeros(2, dtype=dtype)
        v[bits[q]] = 1
        inputs.append((wire[q],))
        arrays.append(v)
    size_dict = {ix: 2 for term in inputs for ix in term}
    return inputs, (), size_dict, arrays
