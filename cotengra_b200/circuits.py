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


def rank_simplify(inputs, output, size_dict, arrays, max_rounds=64):
    """Absorb low-rank tensors into their neighbours before the tree search -- the
    ``rank_simplify`` step quimb applies in the reference notebooks (the m10 network drops
    from 1764 to ~160 tensors, ``examples/Quantum Circuit Example Old.ipynb:143``; the
    shipped m20 benchmark JSON with 381 tensors is the product of the same step).

    Two tensors that share an index are merged whenever the result's rank does not exceed
    the larger of the two ranks (vectors into anything, matrices into anything, gates that
    act on the same qubit pair into each other), until nothing changes.  This is network
    *construction* on the host (numpy, tensors of <= 2^8 elements), not part of the
    contraction path.  Returns ``(inputs, output, size_dict, arrays)`` with the surviving
    index labels unchanged."""
    output = tuple(output)
    terms = {i: tuple(t) for i, t in enumerate(inputs)}
    data = {i: np.asarray(a) for i, a in enumerate(arrays)}
    where = {}
    for i, t in terms.items():
        for ix in t:
            where.setdefault(ix, set()).add(i)
    nxt = len(terms)

    def merge(i, j):
        nonlocal nxt
        ti, tj = terms[i], terms[j]
        shared = [ix for ix in ti if ix in tj]
        # an index survives if it is an output index or lives on a third tensor
        gone = [ix for ix in shared if ix not in output and where[ix] <= {i, j}]
        keep = [ix for ix in ti if ix not in gone] + [ix for ix in tj if ix not in gone and ix not in ti]
        return keep, gone

    for _ in range(max_rounds):
        changed = False
        for i in sorted(terms, key=lambda k: len(terms[k])):
            if i not in terms:
                continue
            best = None
            for ix in terms[i]:
                for j in where[ix]:
                    if j == i or j not in terms:
                        continue
                    keep, gone = merge(i, j)
                    if gone and len(keep) <= max(len(terms[i]), len(terms[j])):
                        if best is None or len(keep) < len(best[1]):
                            best = (j, keep)
            if best is None:
                continue
            j, keep = best
            sym = {}
            for ix in terms[i] + terms[j]:
                sym.setdefault(ix, chr(ord("a") + len(sym)) if len(sym) < 26 else chr(ord("A") + len(sym) - 26))
            eq = ("".join(sym[ix] for ix in terms[i]) + "," + "".join(sym[ix] for ix in terms[j]) + "->"
                  + "".join(sym[ix] for ix in keep))
            new = np.einsum(eq, data[i], data[j])
            for k in (i, j):
                for ix in terms[k]:
                    where[ix].discard(k)
                del terms[k], data[k]
            terms[nxt], data[nxt] = tuple(keep), new
            for ix in keep:
                where.setdefault(ix, set()).add(nxt)
            nxt += 1
            changed = True
        if not changed:
            break
    order = sorted(terms)
    new_inputs = [terms[i] for i in order]
    used = {ix for t in new_inputs for ix in t} | set(output)
    return new_inputs, output, {ix: d for ix, d in size_dict.items() if ix in used}, [data[i] for i in order]
