"""Golden fixtures for BASELINE configs 3 and 4 (Sycamore n53 m10 / m12 amplitudes).

Build container only.  Uses cotengra_b200.circuits to turn the reference's .qsim
files into tensor networks, the UNMODIFIED reference (through oracle/refshim) to
search contraction trees and to compute golden values on the CPU, and writes
tests/golden/circuits.json (+ circuits_values.npz).  Tree search is unseeded, so
the found trees are recorded in the fixture.

    python oracle/gen_circuits.py m10        # unsliced, W <= 2^27
    python oracle/gen_circuits.py m12        # 256 slices
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, os.path.join(HERE, "refshim"), "/root/reference"]

import numpy as np  # noqa: E402

import cotengra as ctg  # noqa: E402

from cotengra_b200 import TreeSpec  # noqa: E402
from cotengra_b200.circuits import amplitude_network  # noqa: E402
from tests.helpers import GOLDEN_DIR  # noqa: E402


def search(inputs, output, size_dict, repeats, max_time, **kw):
    opt = ctg.HyperOptimizer(methods=["greedy"], minimize="combo", max_repeats=repeats,
                             max_time=max_time, reconf_opts={}, parallel=8, optlib="sbplx",
                             progbar=False, **kw)
    return opt.search(inputs, output, size_dict)


def main(which):
    path = os.path.join(GOLDEN_DIR, "circuits.json")
    recs = json.load(open(path)) if os.path.exists(path) else {}
    vpath = os.path.join(GOLDEN_DIR, "circuits_values.npz")
    vals = dict(np.load(vpath)) if os.path.exists(vpath) else {}
    for name in which:
        # "m10s" / "m12s": the rank-simplified network (what the reference notebooks contract:
        # 164 tensors for m10, Quantum Circuit Example Old.ipynb:143)
        simplified = name.endswith("s")
        base = name[:-1] if simplified else name
        qsim = f"/root/reference/examples/circuit_n53_{base}_s0_e0_pABCDCDAB.qsim"
        inputs, output, size_dict, arrays = amplitude_network(qsim)
        if simplified:
            from cotengra_b200.circuits import rank_simplify

            inputs, output, size_dict, arrays = rank_simplify(inputs, output, size_dict, arrays)
            print(name, "rank-simplified to", len(inputs), "tensors", len(size_dict), "indices", flush=True)
        t0 = time.time()
        if base == "m10":
            tree = search(inputs, output, size_dict, 32, 240)
        else:
            tree = search(inputs, output, size_dict, 48, 600)
            tree.slice_(target_slices=256)
        print(name, "search", round(time.time() - t0), "s", tree.contract_stats(), "nslices", tree.nslices,
              "peak", tree.peak_size(), flush=True)
        spec = TreeSpec.from_cotengra(tree)
        rec = {"spec": spec.to_dict(), "qsim": os.path.basename(qsim),
               "contract_stats": {k: int(v) for k, v in tree.contract_stats().items()},
               "nslices": int(tree.nslices), "peak_size": int(tree.peak_size())}
        # golden value(s) from the reference's numpy path
        if base == "m10":
            t0 = time.time()
            val = tree.contract(arrays)
            print(name, "reference contraction", round(time.time() - t0, 1), "s value", val, flush=True)
            vals[f"{name}_amplitude"] = np.asarray(val)
        # a further-sliced copy whose single slices are cheap on the CPU
        small = tree.copy()
        small.slice_(target_size=2**22)
        sspec = TreeSpec.from_cotengra(small)
        rec["small_spec"] = sspec.to_dict()
        for i in (0, 3):
            vals[f"{name}_small_slice{i}"] = np.asarray(small.contract_slice(arrays, i))
        recs[name] = rec
        json.dump(recs, open(path, "w"))
        np.savez_compressed(vpath, **vals)
        # the gate tensors themselves (flattened, input order), so that the fixture is usable
        # where the .qsim files are not (the GPU box)
        apath = os.path.join(GOLDEN_DIR, "circuits_arrays.npz")
        arrs = dict(np.load(apath)) if os.path.exists(apath) else {}
        arrs[f"{name}_arrays_flat"] = np.concatenate([np.asarray(a, dtype=np.complex128).reshape(-1) for a in arrays])
        np.savez_compressed(apath, **arrs)
        print(name, "written", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or ["m10"])
