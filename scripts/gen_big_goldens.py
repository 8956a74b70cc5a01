"""Golden values of single slices of the Sycamore-m20 Appendix-B tree at the widths the GPU is
benchmarked on, computed by the CPU oracle (``oracle/ctg_oracle.py``, the
numpy restatement of the reference's path, pinned to the unmodified reference by
``tests/test_oracle_golden.py``).

    python scripts/gen_big_goldens.py 26 28        # widths (log2 elements of the largest tensor)
    python scripts/gen_big_goldens.py 30           # the bench width itself: ~64 GiB host RAM

W < 30 slices the tree further with ``tests/slicing_util.slice_to_width`` (deterministic), W = 30
is the Appendix-B tree as benchmarked.  Inputs are ``make_arrays(shapes, complex128, seed=0,
scale=0.65)`` -- the bench's operands.  Results are merged into
``tests/golden/big_slices.json``; ``tests/test_gpu_round2.py`` and ``bench.py`` (``parity``
field) check the GPU against them at 1e-10."""

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from oracle import ctg_oracle as orc  # noqa: E402
from tests.helpers import make_arrays  # noqa: E402
from tests.slicing_util import appxB_at_width  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "big_slices.json")
SEED, SCALE = 0, 0.65


def main():
    widths = [int(a) for a in sys.argv[1:]] or [26, 28]
    data = {}
    if os.path.exists(OUT):
        with open(OUT) as f:
            data = json.load(f)
    for w in widths:
        spec = appxB_at_width(w)
        arrays = make_arrays(spec.shapes(), "complex128", seed=SEED, scale=SCALE)
        inputs = [tuple(t) for t in spec.inputs]
        ir = spec.contractions()
        macs, _ = orc.contraction_cost(ir, spec.sliced_shapes())
        for sid in (0, 1):
            t0 = time.perf_counter()
            val = orc.run_contractions(ir, orc.slice_arrays(inputs, spec.sliced, arrays, sid))
            dt = time.perf_counter() - t0
            val = complex(np.asarray(val).reshape(-1)[0])
            key = f"appxB_w{w}_slice{sid}"
            data[key] = {
                "width_log2": w, "slice_id": sid, "seed": SEED, "scale": SCALE, "dtype": "complex128",
                "n_sliced": len(spec.sliced), "macs": int(macs), "re": val.real, "im": val.imag,
                "oracle_seconds": round(dt, 2), "host_cores": os.cpu_count(),
            }
            print(key, val, f"{dt:.1f} s, {8 * macs / dt / 1e9:.1f} GFLOP/s", flush=True)
            with open(OUT, "w") as f:
                json.dump(data, f, indent=1, sort_keys=True)
            # on the GPU box only gpurun_out/ travels back
            scratch = os.path.join(ROOT, "gpurun_out")
            if os.path.isdir(scratch):
                with open(os.path.join(scratch, "big_slices.json"), "w") as f:
                    json.dump(data, f, indent=1, sort_keys=True)
            if w >= 30:
                break  # one slice at the full width is minutes of host time


if __name__ == "__main__":
    main()
