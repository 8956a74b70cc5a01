"""Pin ``oracle/ctg_oracle.py`` to the golden vectors that the UNMODIFIED
reference produced (``oracle/gen_golden.py``): planner tuples bit-exact,
values to rounding."""

import numpy as np
import pytest

from oracle import ctg_oracle as orc
from tests.helpers import (
    decode_ir,
    decode_sliced,
    load_json,
    load_npz,
    make_arrays,
    rel_err,
)

PARSERS = load_json("parsers.json")
PVALS = load_npz("parsers_values.npz")


def _dec_prep(e):
    if isinstance(e, dict):
        return tuple(e["perm"])
    return e


def _tup(x):
    return None if x is None else tuple(x)


def _plan_from_json(p):
    return (
        _dec_prep(p["eq_a"]),
        _dec_prep(p["eq_b"]),
        p["new_shape_a"],
        p["new_shape_b"],
        p["new_shape_ab"],
        _tup(p["perm_ab"]),
        p["pure"],
    )


def _norm_plan(plan):
    eq_a, eq_b, na, nb, nab, perm, pure = plan
    lst = lambda v: None if v is None else list(v)  # noqa: E731
    return (eq_a, eq_b, lst(na), lst(nb), lst(nab), _tup(perm), bool(pure))


def test_pair_planner_bit_exact():
    for n, rec in enumerate(PARSERS["pair"]):
        sa, sb = tuple(rec["shape_a"]), tuple(rec["shape_b"])
        if "error" in rec:
            with pytest.raises(ValueError):
                orc.plan_pair(rec["eq"], sa, sb)
            continue
        got = _norm_plan(orc.plan_pair(rec["eq"], sa, sb))
        want = _norm_plan(_plan_from_json(rec["plan"]))
        assert got == want, (n, rec["eq"], sa, sb)


def test_pair_values():
    checked = 0
    for n, rec in enumerate(PARSERS["pair"]):
        key = f"pair_{n}"
        if key not in PVALS:
            continue
        a, b = make_arrays([rec["shape_a"], rec["shape_b"]], "complex128", seed=n)
        got = orc.einsum(rec["eq"], a, b)
        want = PVALS[key]
        assert got.shape == want.shape, rec
        assert rel_err(got, want) < 1e-13, rec
        checked += 1
    assert checked > 300


def test_single_planner_and_values():
    for n, rec in enumerate(PARSERS["single"]):
        shape = tuple(rec["shape"])
        diag, axes, perm = orc.plan_single(rec["eq"], shape)
        assert (None if diag is None else len(diag)) == rec["n_diag"]
        assert _tup(axes) == _tup(rec["sum_axes"])
        assert _tup(perm) == _tup(rec["perm"])
        (x,) = make_arrays([shape], "complex128", seed=1000 + n)
        got = orc.einsum_single(rec["eq"], x)
        want = PVALS[f"single_{n}"]
        assert got.shape == want.shape
        assert rel_err(got, want) < 1e-13


def test_tensordot_planner_bit_exact():
    for rec in PARSERS["tdot"]:
        axes = (tuple(rec["axes"][0]), tuple(rec["axes"][1]))
        got = _norm_plan(
            orc.plan_tensordot(axes, tuple(rec["shape_a"]), tuple(rec["shape_b"]))
        )
        assert got == _norm_plan(_plan_from_json(rec["plan"]))


TREES = load_json("trees.json")
TVALS = load_npz("trees_values.npz")


@pytest.mark.parametrize("rec", TREES, ids=[r["name"] for r in TREES])
def test_tree_values(rec):
    inputs = [tuple(t) for t in rec["inputs"]]
    sliced = decode_sliced(rec["sliced"])
    ir = decode_ir(rec["contractions"])
    shapes = [tuple(rec["size_dict"][ix] for ix in t) for t in inputs]
    arrays = make_arrays(shapes, rec["dtype"], seed=rec["seed"])

    # slice-id arithmetic is integer work: bit exact
    assert orc.slice_strides(sliced) == rec["slice_strides"]
    for i, key in rec["slice_keys"].items():
        assert orc.slice_key(sliced, int(i)) == key

    name = rec["name"]
    for i in list(rec["slice_keys"])[:3]:
        got = orc.run_contractions(
            ir, orc.slice_arrays(inputs, sliced, arrays, int(i))
        )
        assert rel_err(got, TVALS[f"{name}_slice{i}"]) < 1e-12

    if name in TVALS:
        got = orc.contract_tree(inputs, rec["output"], sliced, ir, arrays)
        want = TVALS[name]
        assert np.shape(got) == want.shape
        assert rel_err(got, want) < 1e-12
        if rec["strip_exponent"]:
            m, e = orc.contract_tree(
                inputs, rec["output"], sliced, ir, arrays, strip_exponent=True
            )
            assert rel_err(m * 10.0**e, want) < 1e-11
            assert rel_err(m, TVALS[name + "_m"]) < 1e-11
            assert abs(e - float(TVALS[name + "_e"])) < 1e-9


def test_equations():
    recs = load_json("equations.json")
    vals = load_npz("equations_values.npz")
    for rec in recs:
        arrays = make_arrays(rec["shapes"], "complex128", seed=rec["seed"])
        want = vals[rec["key"]]
        got = np.einsum(rec["eq"], *arrays)
        # golden == numpy.einsum (how the reference's own tests pin this path)
        assert rel_err(got, want) < 1e-12


def test_sycamore_small_slices():
    recs = {r["name"]: r for r in load_json("sycamore_m20.json")}
    vals = load_npz("sycamore_m20_values.npz")
    rec = recs["sycamore_m20_small"]
    inputs = [tuple(t) for t in rec["inputs"]]
    sliced = decode_sliced(rec["sliced"])
    ir = decode_ir(rec["contractions"])
    shapes = [tuple(rec["size_dict"][ix] for ix in t) for t in inputs]
    arrays = make_arrays(shapes, "complex128", seed=rec["seed"])
    for i in list(rec["slice_keys"])[:3]:
        got = orc.run_contractions(
            ir, orc.slice_arrays(inputs, sliced, arrays, int(i))
        )
        assert rel_err(got, vals[f"sycamore_m20_small_slice{i}"]) < 1e-12
    big = recs["sycamore_m20_appxB"]
    cost, _ = orc.contraction_cost(
        decode_ir(big["contractions"]),
        [
            tuple(2 for ix in t if ix not in {s[0] for s in big["sliced"]})
            for t in big["inputs"]
        ],
    )
    assert cost * big["nslices"] == big["contraction_cost"]
