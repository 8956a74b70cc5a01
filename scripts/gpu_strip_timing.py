"""strip_exponent on the benchmarked slice: time of one m20 Appendix-B slice (W = 2^30) with and
without exponent stripping (the fused epilogue must cost < 5 %), and the
stripped (mantissa, exponent) against the unstripped value.

usage: python scripts/gpu_strip_timing.py [dtype]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cotengra_b200 as cb
from tests.helpers import make_arrays
from tests.slicing_util import appxB_at_width

dtype = sys.argv[1] if len(sys.argv) > 1 else "complex128"
spec = appxB_at_width(30)
arrays = make_arrays(spec.shapes(), dtype, seed=0, scale=0.65)
dev = [torch.from_numpy(a).cuda() for a in arrays]
res = {}
vals = {}
for strip in (False, True):
    ex = cb.TreeExecutor(spec, dtype=dtype, strip_exponent=strip)
    for _ in range(2):
        ex.contract_device(dev, begin=0, step=1, count=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = ex.contract_device(dev, begin=1, step=1, count=4)
    e1.record()
    torch.cuda.synchronize()
    res[strip] = e0.elapsed_time(e1) / 4
    one = ex.contract_device(dev, begin=0, step=1, count=1)
    vals[strip] = (complex(one[0].reshape(-1)[0].item()) * 10.0 ** float(one[1].item())) if strip else complex(one.reshape(-1)[0].item())
    res[f"launches_per_slice_{strip}"] = ex.plan.launches_per_slice()
    ex.plan.profile(True)
    ex.contract_device(dev, begin=7, step=1, count=1)
    torch.cuda.synchronize()
    res[f"nodes_{strip}"] = [(nd["sizes"], int(nd["plan"].variant), t) for nd, t in zip(ex.plan.nodes, ex.plan.profile_read())
                             if nd["kind"] == 0 and not nd["invariant"]]
    ex.plan.profile(False)
    del ex
    torch.cuda.empty_cache()
line = {"dtype": dtype, "slice_ms_plain": res[False], "slice_ms_strip_exponent": res[True],
        "overhead": res[True] / res[False] - 1.0,
        "launches_per_slice": [res["launches_per_slice_False"], res["launches_per_slice_True"]],
        "value_plain": [vals[False].real, vals[False].imag], "value_stripped": [vals[True].real, vals[True].imag],
        "rel_diff": abs(vals[True] - vals[False]) / abs(vals[False])}
print(json.dumps(line))
# the nodes that pay most for stripping
rows = sorted(((b[2] - a[2], a[0], a[1], a[2], b[2]) for a, b in zip(res["nodes_False"], res["nodes_True"])), reverse=True)
for d, sizes, var, t0, t1 in rows[:12]:
    print(f"  +{d:6.2f} ms  {t0:7.2f} -> {t1:7.2f}  B,M,N,K={sizes} variant {var}", file=sys.stderr)
