"""Where a complex64 tree loses accuracy (dev tool): walks the plan of an UNSLICED bench configuration
node by node in complex64 and complex128 and prints, per node, the error the node adds on exact inputs
(local) and the error accumulated so far (chain), both norm-wise against the complex128 walk.
usage: gpu_tree_accuracy.py [config]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import cotengra_b200 as cb
from cotengra_b200 import _lib, lowering as L

config = sys.argv[1] if len(sys.argv) > 1 else "peps8x8"
spec, arrays, _ = bench.load_workload(config, "complex128")
assert spec.nslices == 1
ex = cb.TreeExecutor(spec, dtype="complex64")
plan = ex.plan
sm = _lib.device_info()["sm_count"]
lib = _lib.load()
c64, c128 = torch.complex64, torch.complex128
chain, exact = {}, {}


def get(t):
    if id(t) not in exact:
        assert t.input_index >= 0
        x = torch.from_numpy(np.ascontiguousarray(arrays[t.input_index])).cuda().reshape(-1)
        exact[id(t)] = x
        chain[id(t)] = x.to(c64)
    return chain[id(t)], exact[id(t)]


def nrm(x, ref):
    return (torch.linalg.vector_norm(x.to(c128) - ref) / torch.linalg.vector_norm(ref)).item()


for nd in plan.nodes:
    n_c = int(np.prod(nd["c"].shape)) if nd["c"].shape else 1
    if nd["kind"] != 0:
        a_ch, a_ex = get(nd["a"])
        outs = []
        for src, dt in ((a_ch, c64), (a_ex.to(c64), c64), (a_ex, c128)):
            c = torch.zeros(n_c, dtype=dt, device="cuda")
            w = nd["words"] if dt == c64 else None
            if w is None:
                raise SystemExit("single-operand node: not handled")
            _lib.check(lib.ctgb_reduce_single(np.ascontiguousarray(w).ctypes.data, src.data_ptr(), c.data_ptr(), 0))
            outs.append(c)
        continue
    (a_ch, a_ex), (b_ch, b_ex) = get(nd["a"]), get(nd["b"])
    p64 = nd["plan"]
    p128 = L.build_pair_desc(nd["dims"], "complex128", accumulate=nd["acc"], sm_count=sm, c_dense_elems=nd["dense"])

    def run(pl, a, b, dt):
        c = torch.zeros(n_c, dtype=dt, device="cuda")
        if pl.swapped != p64.swapped:
            a, b = b, a
        _lib.check(lib.ctgb_contract_pair(pl.words.ctypes.data, a.data_ptr(), b.data_ptr(), c.data_ptr(), 0))
        return c

    ref = run(p128, a_ex, b_ex, c128)
    loc = run(p64, a_ex.to(c64), b_ex.to(c64), c64)
    ch = run(p64, a_ch, b_ch, c64)
    torch.cuda.synchronize()
    exact[id(nd["c"])], chain[id(nd["c"])] = ref, ch
    B, M, N, K = nd["sizes"]
    W = p64.words
    # signed shrink: <ref, loc - ref> / <ref, ref> (a negative real part = magnitudes biased towards zero)
    shrink = (torch.vdot(ref, loc.to(c128) - ref) / torch.vdot(ref, ref)).real.item()
    print(f"B={B} M={M} N={N} K={K} var={int(W[L.W_VARIANT])} tile=({int(W[L.W_MTA])},{int(W[L.W_NTA])},{int(W[L.W_KTA])}) "
          f"steps_k={int(W[L.W_STEPS_K])} splitk={int(W[L.W_SPLITK])} local={nrm(loc, ref):.2e} shrink={shrink:+.2e} "
          f"chain={nrm(ch, ref):.2e}", flush=True)
