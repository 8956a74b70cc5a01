"""A minimal, standalone description of a (sliced) contraction tree.

cotengra's tree search stays on the host and unchanged; the executor only
needs the *result* of the search.  ``TreeSpec`` is that result in portable form
-- inputs, output, index sizes, the ordered ``(left, right)`` SSA merges and the
ordered sliced indices -- so that the executor also works where cotengra itself
is not importable (e.g. the benchmark box), and so that a tree can be shipped
as JSON.

From it we regenerate the reference's linear contraction IR -- the
``(parent, left, right, tdot, arg, perm)`` records of
``cotengra.contract.extract_contractions`` (contract.py:573-651) -- by
restating the per-node index metadata of cotengra/core.py:

    leaf legs + preprocessing ... core.py:861-904  (compute_leaf_legs)
    node legs ................... core.py:970-999  (get_legs)
    index order ................. core.py:1034-1051 (get_inds)
    can_dot / axes / perm / eq .. core.py:1024-1095
    slice strides / keys ........ core.py:114-122, 3775-3800

``tests/test_tree_ir.py`` pins every record bit-exactly to golden IR produced
by the unmodified reference (and, in the build container, to the live
reference on freshly generated trees).
"""

from __future__ import annotations

import json
import math

_BASE = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"


def get_symbol(i: int) -> str:
    """a-z, A-Z, then unicode from chr(192), skipping surrogates
    (cotengra/utils.py:657-686)."""
    if i < 52:
        return _BASE[i]
    i += 140
    if i >= 55296:
        i += 2048
    return chr(i)


def _unique(seq):
    return tuple(dict.fromkeys(seq))


class TreeSpec:
    """inputs / output / size_dict / ordered SSA merges / ordered sliced indices."""

    def __init__(self, inputs, output, size_dict, path, sliced=(), node_inds=None):
        # explicit index order of intermediate nodes, ``{ssa_id: "inds"}``: set
        # when the tree went through ``sort_contraction_indices``
        # (core.py:3421-3506); default order otherwise (core.py:1034-1051)
        self.node_inds = {int(k): v for k, v in (node_inds or {}).items()}
        self.inputs = tuple(tuple(t) for t in inputs)
        self.output = tuple(output)
        self.size_dict = dict(size_dict)
        self.path = [tuple(p) for p in path]
        # ordered as the reference's ``tree.sliced_inds``
        self.sliced = [
            (ind, int(size), None if project is None else int(project))
            for ind, size, project in sliced
        ]
        self.N = len(self.inputs)
        if self.N > 1 and len(self.path) != self.N - 1:
            raise ValueError(
                f"a complete tree over {self.N} inputs has {self.N - 1} merges, "
                f"got {len(self.path)}"
            )
        self._sliced_set = {s[0] for s in self.sliced}
        self.appearances = {}
        for term in self.inputs:
            for ix in term:
                self.appearances[ix] = self.appearances.get(ix, 0) + 1
        for ix in self.output:
            self.appearances[ix] = self.appearances.get(ix, 0) + 1
        self._ir = None

    # ------------------------------------------------------------ construction
    @classmethod
    def from_cotengra(cls, tree):
        """Capture a live ``cotengra.ContractionTree`` (duck-typed: uses only
        ``inputs, output, size_dict, sliced_inds, gen_leaves, traverse``)."""
        ssas = {leaf: i for i, leaf in enumerate(tree.gen_leaves())}
        nxt = len(ssas)
        path, node_inds = [], {}
        for p, l, r in tree.traverse():
            path.append((ssas.pop(l), ssas.pop(r)))
            ssas[p] = nxt
            node_inds[nxt] = "".join(tree.get_inds(p))
            nxt += 1
        sliced = [(si.ind, si.size, si.project) for si in tree.sliced_inds.values()]
        return cls(tree.inputs, tree.output, tree.size_dict, path, sliced, node_inds)

    @classmethod
    def from_dict(cls, d):
        return cls(d["inputs"], d["output"], d["size_dict"], d["path"], d.get("sliced", ()),
                   d.get("node_inds"))

    def to_dict(self):
        return {
            "inputs": [list(t) for t in self.inputs],
            "output": list(self.output),
            "size_dict": self.size_dict,
            "path": [list(p) for p in self.path],
            "sliced": [list(s) for s in self.sliced],
            "node_inds": {str(k): v for k, v in self.node_inds.items()},
        }

    def to_json(self):
        return json.dumps(self.to_dict())

    # ------------------------------------------------------------ slicing
    @property
    def nslices(self):
        return math.prod(size for _i, size, project in self.sliced if project is None)

    @property
    def sliced_inputs(self):
        return frozenset(
            c for c, term in enumerate(self.inputs)
            if any(ix in self._sliced_set for ix in term)
        )

    def slice_strides(self):
        n = len(self.sliced)
        strides = [1] * n
        for i in range(n - 2, -1, -1):
            strides[i] = strides[i + 1] * self.sliced[i + 1][1]
        return strides

    def slice_key(self, i):
        key = {}
        for (ind, _size, project), stride in zip(self.sliced, self.slice_strides()):
            if project is None:
                key[ind] = i // stride
                i %= stride
            else:
                key[ind] = project
        return key

    def sliced_shapes(self):
        return [
            tuple(self.size_dict[ix] for ix in term if ix not in self._sliced_set)
            for term in self.inputs
        ]

    def shapes(self):
        return [tuple(self.size_dict[ix] for ix in term) for term in self.inputs]

    # ------------------------------------------------------------ index metadata
    def _leaf(self, i):
        """``(legs, preprocessing_eq_or_None)`` of input ``i`` after slicing."""
        term = tuple(ix for ix in self.inputs[i] if ix not in self._sliced_set)
        legs = {}
        for ix in term:
            legs[ix] = legs.get(ix, 0) + 1
        simplify = len(term) != len(legs) or any(
            c == self.appearances[ix] for ix, c in legs.items()
        )
        eq = None
        if simplify:
            legs = {ix: c for ix, c in legs.items() if c != self.appearances[ix]}
            # utils.py:1162-1169 canonicalises lazily: the *output* term is
            # materialised first, so surviving legs get the first symbols
            sym = {}
            for ix in tuple(legs) + term:
                if ix not in sym:
                    sym[ix] = get_symbol(len(sym))
            eq = "".join(sym[ix] for ix in term) + "->" + "".join(sym[ix] for ix in legs)
        return legs, eq

    def contractions(self):
        """The reference's linear IR for this tree (contract.py:573-651)."""
        if self._ir is not None:
            return self._ir
        if self.N == 1:
            term = "".join(ix for ix in self.inputs[0] if ix not in self._sliced_set)
            out = "".join(ix for ix in self.output if ix not in self._sliced_set)
            self._ir = ((1, 0, None, False, f"{term}->{out}", None),)
            self.inds = {0: term, 1: out}
            self.preprocessing = {}
            return self._ir

        legs, inds, pre = {}, {}, {}
        for i in range(self.N):
            lg, eq = self._leaf(i)
            legs[i] = lg
            inds[i] = "".join(lg)
            if eq is not None:
                pre[i] = eq
        root_legs = [ix for ix in self.output if ix not in self._sliced_set]

        records = []
        nxt = self.N
        last = self.N + len(self.path) - 1
        for l, r in self.path:
            p = nxt
            nxt += 1
            if p == last:
                lp = {ix: 0 for ix in root_legs}
                ip = "".join(root_legs)
            else:
                involved = dict(legs[l])
                for ix, c in legs[r].items():
                    involved[ix] = involved.get(ix, 0) + c
                lp = {ix: c for ix, c in involved.items() if c < self.appearances[ix]}
                ip = "".join(_unique(ix for ix in inds[l] + inds[r] if ix in lp))
                if p in self.node_inds:
                    custom = self.node_inds[p]
                    if sorted(custom) != sorted(ip):
                        raise ValueError(f"node_inds[{p}] is not a permutation of the node's legs")
                    ip = custom
            legs[p], inds[p] = lp, ip
            il, ir_ = inds[l], inds[r]
            can_dot = set(lp) == set(legs[l]).symmetric_difference(legs[r])
            if can_dot:
                ax_l, ax_r = [], []
                for i, ix in enumerate(il):
                    j = ir_.find(ix)
                    if j != -1:
                        ax_l.append(i)
                        ax_r.append(j)
                both = il + ir_
                td = "".join(sorted(ip, key=both.find))
                perm = None if td == ip else tuple(td.find(ix) for ix in ip)
                records.append((p, l, r, True, (tuple(ax_l), tuple(ax_r)), perm))
            else:
                sym = {}
                for ix in il + ir_:
                    if ix not in sym:
                        sym[ix] = get_symbol(len(sym))
                eq = (
                    "".join(sym[ix] for ix in il) + ","
                    + "".join(sym[ix] for ix in ir_) + "->"
                    + "".join(sym[ix] for ix in ip)
                )
                records.append((p, l, r, False, eq, None))
        pre_records = tuple((i, None, None, False, eq, None) for i, eq in pre.items())
        self._ir = pre_records + tuple(records)
        self.inds = inds
        self.preprocessing = pre
        return self._ir
