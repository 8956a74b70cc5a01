"""``ExecPlan`` (arenas, slice offsets, invariant hoisting, root views,
exponent stripping) emulated on the CPU against the reference's golden values."""

import numpy as np
import pytest

from cotengra_b200 import ExecPlan, TreeSpec
from tests.desc_emulator import emulate_plan
from tests.helpers import decode_ir, decode_sliced, load_json, load_npz, make_arrays, rel_err

TREES = load_json("trees.json")
TVALS = load_npz("trees_values.npz")


def _plan(rec, ir=None, **kw):
    n_in = len(rec["inputs"])
    node_inds = {int(k): v for k, v in rec["inds"].items() if int(k) >= n_in}
    spec = TreeSpec(rec["inputs"], rec["output"], rec["size_dict"], rec["path"],
                    decode_sliced(rec["sliced"]), node_inds)
    ir = spec.contractions() if ir is None else ir
    return spec, ExecPlan(ir, spec.inputs, spec.output, spec.size_dict, spec.sliced,
                          dtype=rec["dtype"], sm_count=8, **kw)


@pytest.mark.parametrize("rec", TREES, ids=[r["name"] for r in TREES])
def test_plan_values(rec):
    if rec["name"] not in TVALS:
        pytest.skip("no full value recorded")
    spec, plan = _plan(rec)
    arrays = make_arrays(spec.shapes(), rec["dtype"], seed=rec["seed"])
    want = TVALS[rec["name"]]
    got = emulate_plan(plan, arrays)
    assert got.shape == want.shape
    assert rel_err(got, want) < 1e-11
    # the reference's own IR (from the golden file) gives the same plan result
    _, plan2 = _plan(rec, ir=decode_ir(rec["contractions"]), hoist=False)
    assert rel_err(emulate_plan(plan2, arrays), want) < 1e-11
    if rec["strip_exponent"]:
        _, plan3 = _plan(rec, strip_exponent=True)
        m, e = emulate_plan(plan3, arrays)
        assert rel_err(m * 10.0**e, want) < 1e-10
        # same (mantissa, exponent) normalisation as gather_slices: max|m| <= 1 scale
        wm, we = TVALS[rec["name"] + "_m"], float(TVALS[rec["name"] + "_e"])
        assert rel_err(m * 10.0 ** (e - we), wm) < 1e-10


def test_single_slices_and_round_robin():
    rec = next(r for r in TREES if r["name"] == "lattice6x6_d3_sliced")
    spec, plan = _plan(rec)
    arrays = make_arrays(spec.shapes(), rec["dtype"], seed=rec["seed"])
    for i in list(rec["slice_keys"])[:3]:
        got = emulate_plan(plan, arrays, slice_ids=[int(i)])
        assert rel_err(got, TVALS[f"{rec['name']}_slice{i}"]) < 1e-11
    # rank-strided partial sums add up to the full result (core.py:4070)
    world = 3
    parts = [emulate_plan(plan, arrays, slice_ids=range(r, plan.nslices, world)) for r in range(world)]
    assert rel_err(sum(parts), TVALS[rec["name"]]) < 1e-11


def test_hoisting_marks_invariant_nodes():
    rec = next(r for r in load_json("sycamore_m20.json") if r["name"] == "sycamore_m20_appxB")
    spec, plan = _plan(rec)
    inv = sum(1 for nd in plan.nodes if nd["invariant"])
    assert inv == 195  # 195 of 380 nodes are slice-invariant
    assert plan.macs_per_slice + plan.macs_invariant == rec["contraction_cost"] // rec["nslices"]
    assert plan.nslices == 2**36
    # peak live memory stays near the reference's tree.peak_size()
    assert plan.workspace_bytes <= 1.25 * rec["peak_size"] * plan.esize


def test_arena_never_overlaps_live_tensors():
    rec = next(r for r in TREES if r["name"] == "peps8x8_d2")
    spec, plan = _plan(rec)
    live = {}
    for nd in plan.nodes:
        c = nd["c"]
        if c.kind == 1:
            for other in live.values():
                assert c.offset + c.nbytes <= other.offset or other.offset + other.nbytes <= c.offset
            live[id(c)] = c
        for s in (nd["a"], nd["b"]):
            if s is not None and s.kind == 1 and s.last_use == nd["pos"]:
                live.pop(id(s), None)


@pytest.mark.parametrize("name,which", [("lattice6x6_d3_sliced", 0), ("rand_r3_o0_hi0_ho1_root_s666_sliced", 1)])
def test_zero_slices_keep_the_sum_finite(name, which):
    """check_zero (contract.py:819-820): a slice with an all-zero intermediate is
    (0, -inf) and drops out of the exponent-aware sum (core.py:163-170) instead of
    poisoning it with 0/0.  (The zeroed digit belongs to an *inner* sliced index: when every
    slice of one output chunk is zero the reference's adder itself forms 10**(-inf - -inf) = nan,
    core.py:163-170; the device path returns the zeros.)"""
    from oracle import ctg_oracle as orc
    from tests.zero_util import zero_one_digit

    rec = next(r for r in TREES if r["name"] == name)
    spec, plan = _plan(rec, strip_exponent=True)
    arrays, _ = zero_one_digit(spec, make_arrays(spec.shapes(), rec["dtype"], seed=rec["seed"]), which)
    wm, we = orc.contract_tree([tuple(t) for t in spec.inputs], spec.output, spec.sliced,
                               spec.contractions(), arrays, strip_exponent=True, check_zero=True)
    m, e = emulate_plan(plan, arrays)
    assert np.all(np.isfinite(m)) and np.isfinite(e)
    assert rel_err(m * 10.0**e, wm * 10.0**we) < 1e-10
    # every slice zero: mantissa zeros, exponent -inf
    zeros = [np.zeros_like(a) for a in arrays]
    m0, e0 = emulate_plan(plan, zeros)
    assert e0 == -np.inf and not np.any(m0)


SLICED_OUT = [r["name"] for r in TREES if r["name"].endswith("_sliced_out") and r["name"] in TVALS] + ["projected"]


@pytest.mark.parametrize("name", SLICED_OUT)
def test_output_chunks_geometry(name):
    """gen_output_chunks (core.py:3884-3941): chunk o = sum of slice ids [o*step, (o+1)*step)
    of a plan whose output term has no sliced index; stacked by key they give the full result."""
    from cotengra_b200.executor import output_chunking

    rec = next(r for r in TREES if r["name"] == name)
    spec, _ = _plan(rec)
    chunk_out, step, nchunks = output_chunking(spec)
    plan = ExecPlan(spec.contractions(), spec.inputs, chunk_out, spec.size_dict, spec.sliced,
                    dtype=rec["dtype"], sm_count=8)
    arrays = make_arrays(spec.shapes(), rec["dtype"], seed=rec["seed"])
    want = TVALS[name]
    assert step * nchunks == spec.nslices
    for o in range(nchunks):
        got = emulate_plan(plan, arrays, slice_ids=range(o * step, (o + 1) * step))
        key = spec.slice_key(o * step)
        sel = tuple(key[ix] if (ix in key and spec.sliced[[s[0] for s in spec.sliced].index(ix)][2] is None)
                    else slice(None) for ix in spec.output)
        ref = want[sel]
        # projected output indices keep extent 1 in the full result
        ref = ref.reshape(got.shape)
        assert rel_err(got, ref) < 1e-11
