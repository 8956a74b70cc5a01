"""Stem fusion (cotengra_b200/fusion.py): the re-associated tree is the same multilinear
map -- checked against the reference's golden values through the CPU emulator of the device
addressing -- and moves fewer bytes on the Sycamore stem."""

import math

import numpy as np
import pytest

from cotengra_b200 import ExecPlan, TreeSpec
from cotengra_b200.fusion import fuse_stems, node_time, tree_work
from tests.desc_emulator import emulate_plan
from tests.helpers import decode_sliced, load_json, load_npz, make_arrays, rel_err

TREES = load_json("trees.json")
TVALS = load_npz("trees_values.npz")


def _spec(rec):
    n_in = len(rec["inputs"])
    node_inds = {int(k): v for k, v in rec["inds"].items() if int(k) >= n_in}
    return TreeSpec(rec["inputs"], rec["output"], rec["size_dict"], rec["path"],
                    decode_sliced(rec["sliced"]), node_inds)


def _bytes_only(dtype, B, M, N, K, elems):
    """A model that always prefers fewer bytes: forces fusion on small test trees."""
    return 1e-9 * elems + 1e-12 * B * M * N * K


def _plan(spec, dtype, **kw):
    return ExecPlan(spec.contractions(), spec.inputs, spec.output, spec.size_dict, spec.sliced,
                    dtype=dtype, sm_count=8, **kw)


FUSABLE = [r["name"] for r in TREES if r["name"] in TVALS and len(r["inputs"]) >= 5]


@pytest.mark.parametrize("name", FUSABLE)
def test_fused_tree_matches_golden(name):
    rec = next(r for r in TREES if r["name"] == name)
    spec = _spec(rec)
    new, info = fuse_stems(spec, rec["dtype"], min_big=2, ratio=1.0, min_gain=-1.0, model=_bytes_only)
    arrays = make_arrays(spec.shapes(), rec["dtype"], seed=rec["seed"])
    want = TVALS[name]
    got = emulate_plan(_plan(new, rec["dtype"]), arrays)
    assert got.shape == want.shape
    assert rel_err(got, want) < 1e-10
    if rec["strip_exponent"]:
        m, e = emulate_plan(_plan(new, rec["dtype"], strip_exponent=True), arrays)
        assert rel_err(m * 10.0**e, want) < 1e-10
    # same leaves, same output, same slicing; a complete binary tree
    assert new.inputs == spec.inputs and new.output == spec.output and new.sliced == spec.sliced
    assert len(new.path) == len(spec.path)
    if info["changed"]:
        assert sorted(x for p in new.path for x in p) == list(range(2 * spec.N - 2))


def test_some_golden_trees_really_change():
    changed = 0
    for name in FUSABLE:
        spec = _spec(next(r for r in TREES if r["name"] == name))
        _new, info = fuse_stems(spec, "complex128", min_big=2, ratio=1.0, min_gain=-1.0, model=_bytes_only)
        changed += bool(info["changed"])
    assert changed >= 10


def test_default_thresholds_leave_small_trees_alone():
    for name in FUSABLE[:20]:
        spec = _spec(next(r for r in TREES if r["name"] == name))
        new, info = fuse_stems(spec, "complex128")
        assert new is spec and not info["changed"]


def test_sycamore_stem_moves_fewer_bytes():
    rec = next(r for r in load_json("sycamore_m20.json") if r["name"] == "sycamore_m20_appxB")
    spec = TreeSpec(rec["inputs"], rec["output"], rec["size_dict"], rec["path"], decode_sliced(rec["sliced"]))
    new, info = fuse_stems(spec, "complex128")
    assert info["changed"] and info["root_peel"]["peel_left"] + info["root_peel"]["peel_right"] >= 1
    macs0, inv0, el0 = tree_work(spec)
    macs1, inv1, el1 = tree_work(new)
    assert macs0 + inv0 == rec["contraction_cost"] // rec["nslices"]
    assert el1 < 0.85 * el0            # bytes moved / bytes of the unfused tree
    assert macs1 < 1.1 * macs0         # for a few per cent more flops
    p0, p1 = _plan(spec, "complex128"), _plan(new, "complex128")
    assert p1.elements_per_slice == el1 and p0.elements_per_slice == el0
    assert p1.workspace_bytes <= 1.01 * p0.workspace_bytes
    # the peeled stem tail runs on the dot-stream kernel, the fused pair on the DMMA stream
    from cotengra_b200 import lowering as L

    variants = [nd["plan"].variant for nd in p1.nodes if nd["kind"] == 0 and not nd["invariant"]]
    assert L.VAR_DMMA_32x32 in variants or L.VAR_DOTSTREAM4 in variants
    # estimated time drops by more than 10 %
    def est(p):
        return sum(node_time("complex128", *nd["sizes"], sum(math.prod(x.shape) for x in (nd["a"], nd["b"], nd["c"])))
                   for nd in p.nodes if nd["kind"] == 0 and not nd["invariant"])
    assert est(p1) < 0.9 * est(p0)


def test_sycamore_small_fused_slices_match_reference():
    """The m20 network sliced down to oracle size, fusion forced: golden slice values."""
    recs = {r["name"]: r for r in load_json("sycamore_m20.json")}
    vals = load_npz("sycamore_m20_values.npz")
    rec = recs["sycamore_m20_small"]
    spec = _spec(rec)
    new, info = fuse_stems(spec, "complex128", min_big=16, ratio=2.0, min_gain=-1.0, model=_bytes_only)
    assert info["changed"]
    arrays = make_arrays(spec.shapes(), "complex128", seed=rec["seed"])
    plan = _plan(new, "complex128")
    for i in list(rec["slice_keys"])[:2]:
        if int(i) >= 2**62:
            continue
        got = emulate_plan(plan, arrays, slice_ids=[int(i)])
        assert rel_err(got, vals[f"sycamore_m20_small_slice{i}"]) < 1e-10
