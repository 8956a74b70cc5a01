"""Partial-sum checkpointing of long sliced runs (``contract_checkpointed``)
on the CPU: the block loop, the checkpoint file and the resume logic are host code; the
per-block arithmetic is delegated to a stand-in executor that runs the numpy oracle on the
requested slice range (on the GPU box the same function drives ``TreeExecutor``)."""

import numpy as np
import pytest

import cotengra_b200 as cb
from oracle import ctg_oracle as orc
from tests.helpers import decode_ir, decode_sliced, load_json, load_npz, make_arrays, rel_err

TREES = load_json("trees.json")
TVALS = load_npz("trees_values.npz")


class OracleExecutor:
    """Duck-typed TreeExecutor: spec, dtype, strip_exponent, nslices, contract_host."""

    def __init__(self, rec, strip=False, fail_after=None):
        n_in = len(rec["inputs"])
        node_inds = {int(k): v for k, v in rec["inds"].items() if int(k) >= n_in}
        self.spec = cb.TreeSpec(rec["inputs"], rec["output"], rec["size_dict"], rec["path"],
                                decode_sliced(rec["sliced"]), node_inds)
        self.rec, self.dtype, self.strip_exponent = rec, rec["dtype"], strip
        self.nslices = rec["nslices"]
        self.calls, self.fail_after = [], fail_after

    def contract_host(self, arrays, begin, step, count):
        if self.fail_after is not None and len(self.calls) >= self.fail_after:
            raise KeyboardInterrupt("simulated loss of the job")
        self.calls.append((begin, count))
        rec = self.rec
        return orc.contract_tree([tuple(t) for t in rec["inputs"]], rec["output"], decode_sliced(rec["sliced"]),
                                 decode_ir(rec["contractions"]), arrays, strip_exponent=self.strip_exponent,
                                 slice_ids=range(begin, begin + step * count, step))


def _inner_sliced(rec):
    sliced = decode_sliced(rec["sliced"])
    return bool(sliced) and {s[0] for s in sliced}.isdisjoint(rec["output"]) and rec["nslices"] >= 8


RECS = [r for r in TREES if _inner_sliced(r)][:6]


@pytest.mark.parametrize("strip", [False, True])
@pytest.mark.parametrize("rec", RECS, ids=[r["name"] for r in RECS])
def test_checkpoint_resume_equals_uninterrupted(rec, strip, tmp_path):
    arrays = make_arrays([tuple(rec["size_dict"][ix] for ix in t) for t in rec["inputs"]], rec["dtype"],
                         seed=rec["seed"])
    want = TVALS[rec["name"]]
    ck = str(tmp_path / "run.npz")
    every = max(1, rec["nslices"] // 5)
    # first attempt dies after two blocks: the file holds the sum of exactly those
    ex = OracleExecutor(rec, strip, fail_after=2)
    with pytest.raises(KeyboardInterrupt):
        cb.contract_checkpointed(None, arrays, ck, every=every, executor=ex)
    with np.load(ck) as z:
        assert int(z["next_slice"]) == 2 * every
    # the second attempt resumes behind them and never recomputes a finished block
    ex2 = OracleExecutor(rec, strip)
    seen = []
    res = cb.contract_checkpointed(None, arrays, ck, every=every, executor=ex2,
                                   on_block=lambda done, n: seen.append((done, n)))
    assert ex2.calls[0][0] == 2 * every and sum(c for _b, c in ex2.calls) == rec["nslices"] - 2 * every
    assert seen[-1] == (rec["nslices"], rec["nslices"])
    got = res[0] * 10.0 ** res[1] if strip else res
    assert rel_err(got, want) < 1e-10
    # a finished checkpoint answers without any further work
    ex3 = OracleExecutor(rec, strip)
    res3 = cb.contract_checkpointed(None, arrays, ck, every=every, executor=ex3)
    assert ex3.calls == []
    got3 = res3[0] * 10.0 ** res3[1] if strip else res3
    assert rel_err(got3, want) < 1e-10


def test_checkpoint_of_other_inputs_is_refused(tmp_path):
    rec = RECS[0]
    arrays = make_arrays([tuple(rec["size_dict"][ix] for ix in t) for t in rec["inputs"]], rec["dtype"],
                         seed=rec["seed"])
    ck = str(tmp_path / "run.npz")
    cb.contract_checkpointed(None, arrays, ck, every=4, executor=OracleExecutor(rec))
    other = [a.copy() for a in arrays]
    other[0] = other[0] * 2.0
    with pytest.raises(ValueError):
        cb.contract_checkpointed(None, other, ck, every=4, executor=OracleExecutor(rec))
    with pytest.raises(ValueError):  # same inputs, other mode
        cb.contract_checkpointed(None, arrays, ck, every=4, executor=OracleExecutor(rec, strip=True))
