"""Per-node accuracy of a complex64 plan (dev tool): every distinct node of a bench configuration is run
with the variant the plan chose, on random operands, and compared norm-wise with the same node computed in
complex128.  usage: gpu_node_accuracy.py [config] (m20|peps8x8|m10|m10s|m12)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import cotengra_b200 as cb
from cotengra_b200 import _lib, lowering as L

config = sys.argv[1] if len(sys.argv) > 1 else "peps8x8"
spec, _arrays, _ = bench.load_workload(config, "complex64")
from cotengra_b200.fusion import fuse_stems

sm = _lib.device_info()["sm_count"]
spec, _info = fuse_stems(spec, "complex64")  # the tree the executor runs by default
plan = cb.ExecPlan(spec.contractions(), spec.inputs, spec.output, spec.size_dict, spec.sliced, dtype="complex64",
                   sm_count=sm)
only_tc05 = "--tc05" in sys.argv
lib = _lib.load()
seen = set()
for nd in plan.nodes:
    if nd["kind"] != 0:
        continue
    key = (nd["sizes"], int(nd["plan"].variant))
    if key in seen:
        continue
    seen.add(key)
    B, M, N, K = nd["sizes"]
    big = "--big" in sys.argv
    if (B * M * K > 2**29 or B * M * N > 2**29) != big or B * M * N > 2**32:
        continue
    if only_tc05 and int(nd["plan"].variant) not in L.TC05_VARIANTS:
        continue
    g = torch.Generator(device="cuda").manual_seed(B + M + N + K)

    def buf(t):
        n = int(np.prod(t.shape))
        x = torch.empty(n, dtype=torch.complex64, device="cuda")
        torch.view_as_real(x).uniform_(-1, 1, generator=g)
        return x

    a, b = buf(nd["a"]), buf(nd["b"])
    n_c = int(np.prod(nd["c"].shape))
    outs = []
    for dt, tdt in (("complex64", torch.complex64), ("complex128", torch.complex128)):
        pl = L.build_pair_desc(nd["dims"], dt, sm_count=sm, c_dense_elems=n_c)
        c = torch.full((n_c,), float("nan"), dtype=tdt, device="cuda")
        x, y = (a, b) if tdt == torch.complex64 else (a.to(tdt), b.to(tdt))
        pa, pb = (y, x) if pl.swapped != nd["plan"].swapped else (x, y)
        _lib.check(lib.ctgb_contract_pair(pl.words.ctypes.data, pa.data_ptr(), pb.data_ptr(), c.data_ptr(), 0))
        outs.append((c, pl))
        torch.cuda.synchronize()
        del x, y, pa, pb
    torch.cuda.synchronize()
    got, ref = outs[0][0], outs[1][0]
    del a, b
    num = den = dmax = rmax = 0.0
    for o in range(0, n_c, 2**26):  # in pieces: the operands of the big nodes leave no room for a widened copy
        d = got[o:o + 2**26].to(torch.complex128) - ref[o:o + 2**26]
        num += torch.linalg.vector_norm(d).item() ** 2
        den += torch.linalg.vector_norm(ref[o:o + 2**26]).item() ** 2
        dmax, rmax = max(dmax, d.abs().max().item()), max(rmax, ref[o:o + 2**26].abs().max().item())
        del d
    err, mx = (num / den) ** 0.5, dmax / rmax
    W = outs[0][1].words
    del got, ref, outs, c
    torch.cuda.empty_cache()
    print(f"B={B} M={M} N={N} K={K} var={int(W[L.W_VARIANT])} tile=({int(W[L.W_MTA])},{int(W[L.W_NTA])},{int(W[L.W_KTA])}) "
          f"steps_k={int(W[L.W_STEPS_K])} splitk={int(W[L.W_SPLITK])} norm_err={err:.2e} max_err={mx:.2e}", flush=True)
