"""Cross-check every distinct (shape, variant) node of the Sycamore slice plan against
the generic FMA kernel on random operands (dev tool): finds which kernel variant is wrong."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import cotengra_b200 as cb
from cotengra_b200 import _lib, lowering as L
from tests.helpers import decode_sliced, load_json
from tests.slicing_util import slice_to_width

dtype = sys.argv[1] if len(sys.argv) > 1 else "complex128"
wlog = int(sys.argv[2]) if len(sys.argv) > 2 else 26
rec = next(r for r in load_json("sycamore_m20.json") if r["name"] == "sycamore_m20_appxB")
spec = cb.TreeSpec(rec["inputs"], rec["output"], rec["size_dict"], rec["path"], decode_sliced(rec["sliced"]))
if wlog < 30:
    spec = slice_to_width(spec, 2**wlog)
sm = _lib.device_info()["sm_count"]
plan = cb.ExecPlan(spec.contractions(), spec.inputs, spec.output, spec.size_dict, spec.sliced, dtype=dtype, sm_count=sm)
tdt = getattr(torch, dtype)
lib = _lib.load()
seen = set()
bad = 0
for nd in plan.nodes:
    if nd["kind"] != 0 or nd["invariant"] or nd["root"]:
        continue
    key = (nd["sizes"], int(nd["plan"].variant), int(nd["words"][L.W_NGM]), int(nd["words"][L.W_NGN]))
    if key in seen:
        continue
    seen.add(key)
    def buf(t):
        n = int(np.prod(t.shape))
        x = torch.empty(n, dtype=tdt, device="cuda")
        torch.view_as_real(x).uniform_(-1, 1) if x.is_complex() else x.uniform_(-1, 1)
        return x
    a, b = buf(nd["a"]), buf(nd["b"])
    n_c = int(np.prod(nd["c"].shape))
    outs = []
    for variant in (None, L.VAR_SIMT_64x64):
        pl = L.build_pair_desc(nd["dims"], dtype, sm_count=sm, c_dense_elems=n_c, variant=variant)
        c = torch.full((n_c,), float("nan"), dtype=tdt, device="cuda")
        pa, pb = (b, a) if pl.swapped != nd["plan"].swapped else (a, b)
        _lib.check(lib.ctgb_contract_pair(pl.words.ctypes.data, pa.data_ptr(), pb.data_ptr(), c.data_ptr(), 0))
        outs.append(c)
    torch.cuda.synchronize()
    err = (outs[0] - outs[1]).abs().max().item() / max(outs[1].abs().max().item(), 1e-300)
    nbad = int((~torch.isfinite(torch.view_as_real(outs[0]) if outs[0].is_complex() else outs[0])).sum().item())
    flag = "" if err < (1e-6 if dtype in ("complex128", "float64") else 2e-5) and nbad == 0 else "   <<<<<< MISMATCH"
    if flag:
        bad += 1
    B, M, N, K = nd["sizes"]
    W = nd["words"]
    print(f"M=2^{int(np.log2(M))} N={N} K={K} var={key[1]} tiles=({int(W[L.W_TILES_M])},{int(W[L.W_TILES_N])}) ngm={key[2]} flags={int(W[L.W_FLAGS])} relerr={err:.2e} nonfinite={nbad}{flag}", flush=True)
    del a, b, outs
print("XCHECK", "FAIL" if bad else "PASS", bad)
