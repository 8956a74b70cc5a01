"""Walk slice 0 of a sliced bench configuration node by node on its real data (dev tool): every tcgen05
node is run a second time on the generic FMA kernel and the two outputs are compared -- finds the node
that goes wrong inside a tree when the same node on random operands does not.
usage: gpu_tree_xcheck.py [config] [dtype]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import cotengra_b200 as cb
from cotengra_b200 import _lib, lowering as L
from cotengra_b200.fusion import fuse_stems

config = sys.argv[1] if len(sys.argv) > 1 else "m12"
dtype = sys.argv[2] if len(sys.argv) > 2 else "complex64"
tdt = getattr(torch, dtype)
spec, arrays, _ = bench.load_workload(config, dtype)
sm = _lib.device_info()["sm_count"]
spec, _info = fuse_stems(spec, dtype)
plan = cb.ExecPlan(spec.contractions(), spec.inputs, spec.output, spec.size_dict, spec.sliced, dtype=dtype, sm_count=sm)
lib = _lib.load()
uses = {}
for nd in plan.nodes:
    for s in (nd["a"], nd["b"]):
        if s is not None:
            uses[id(s)] = uses.get(id(s), 0) + 1
bufs = {}


def get(t):
    if id(t) not in bufs:
        assert t.input_index >= 0
        # slice 0: every sliced digit is 0, the view starts at the array's first element
        bufs[id(t)] = torch.from_numpy(np.ascontiguousarray(arrays[t.input_index])).to(tdt).cuda().reshape(-1)
    return bufs[id(t)]


def done(t):
    uses[id(t)] -= 1
    if uses[id(t)] == 0:
        del bufs[id(t)]


bad = 0
for pos, nd in enumerate(plan.nodes):
    n_c = int(np.prod(nd["c"].shape)) if nd["c"].shape else 1
    a = get(nd["a"])
    c = torch.zeros(n_c, dtype=tdt, device="cuda") if nd.get("acc") or nd["kind"] != 0 else torch.empty(n_c, dtype=tdt, device="cuda")
    if nd["kind"] != 0:
        w = np.ascontiguousarray(nd["words"])
        _lib.check(lib.ctgb_reduce_single(w.ctypes.data, a.data_ptr(), c.data_ptr(), 0))
        bufs[id(nd["c"])] = c
        done(nd["a"])
        continue
    b = get(nd["b"])
    pl = nd["plan"]
    _lib.check(lib.ctgb_contract_pair(pl.words.ctypes.data, a.data_ptr(), b.data_ptr(), c.data_ptr(), 0))
    B, M, N, K = nd["sizes"]
    W = pl.words
    if int(pl.variant) in L.TC05_VARIANTS:
        p2 = L.build_pair_desc(nd["dims"], dtype, accumulate=nd["acc"], sm_count=sm, c_dense_elems=nd["dense"],
                               variant=L.VAR_SIMT_64x64)
        c2 = torch.zeros(n_c, dtype=tdt, device="cuda")
        x, y = (b, a) if p2.swapped != pl.swapped else (a, b)
        _lib.check(lib.ctgb_contract_pair(p2.words.ctypes.data, x.data_ptr(), y.data_ptr(), c2.data_ptr(), 0))
        torch.cuda.synchronize()
        num = den = 0.0
        for o in range(0, n_c, 2**27):
            d = c[o:o + 2**27] - c2[o:o + 2**27]
            num += torch.linalg.vector_norm(d).item() ** 2
            den += torch.linalg.vector_norm(c2[o:o + 2**27]).item() ** 2
            del d
        err = (num / max(den, 1e-300)) ** 0.5
        flag = "" if err < 1e-4 else "   <<<<<< MISMATCH"
        bad += bool(flag)
        print(f"node {pos}: M={M} N={N} K={K} var={int(W[L.W_VARIANT])} steps_k={int(W[L.W_STEPS_K])} flags={int(W[L.W_FLAGS])} "
              f"c_align={c.data_ptr() % 512} err_vs_fma={err:.2e}{flag}", flush=True)
        del c2
    bufs[id(nd["c"])] = c
    done(nd["a"])
    done(nd["b"])
    del a, b, c
    torch.cuda.empty_cache()
print("TREE XCHECK", "FAIL" if bad else "PASS", bad, "result", bufs[id(plan.nodes[-1]["c"])][:4].cpu().numpy())
