"""Standalone benchmark of the heaviest nodes of one Sycamore-m20 slice (dev tool).

usage: python scripts/gpu_node_bench.py [dtype] [topk] [--ncu] [--rows] [--mnk=M,N,K] [--nofuse] [--variant=ID]
Each selected node is launched alone through ctgb_contract_pair on dummy
operands of the right size, timed with CUDA events; with --ncu one launch per
node is bracketed by cudaProfilerStart/Stop (run under
`ncu --profile-from-start off`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import cotengra_b200 as cb
from cotengra_b200 import _lib, lowering as L
from tests.helpers import decode_sliced, load_json, make_arrays

dtype = sys.argv[1] if len(sys.argv) > 1 else "complex128"
topk = int(sys.argv[2]) if len(sys.argv) > 2 else 8
use_ncu = "--ncu" in sys.argv
rows_only = "--rows" in sys.argv
mnk = next((tuple(int(x) for x in a.split("=")[1].split(",")) for a in sys.argv if a.startswith("--mnk=")), None)
rec = next(r for r in load_json("sycamore_m20.json") if r["name"] == "sycamore_m20_appxB")
spec = cb.TreeSpec(rec["inputs"], rec["output"], rec["size_dict"], rec["path"], decode_sliced(rec["sliced"]))
if "--nofuse" not in sys.argv:
    from cotengra_b200.fusion import fuse_stems

    spec, _info = fuse_stems(spec, dtype)  # the plan the executor runs by default
force = next((int(a.split("=")[1]) for a in sys.argv if a.startswith("--variant=")), None)
plan = cb.ExecPlan(spec.contractions(), spec.inputs, spec.output, spec.size_dict, spec.sliced, dtype=dtype,
                   sm_count=_lib.device_info()["sm_count"], variant=force)
es = plan.esize
nodes = [nd for nd in plan.nodes if nd["kind"] == 0 and not nd["invariant"]]
if rows_only:
    nodes = [nd for nd in nodes if nd["sizes"][2] <= 16 or nd["sizes"][1] == 1]
# rank by ideal cost model: max(flop time at 37 TF, byte time at 6.4 TB/s) is not known a
# priori, so rank by elements moved + flops/6
def weight(nd):
    B, M, N, K = nd["sizes"]
    el = sum(int(np.prod(x.shape)) for x in (nd["a"], nd["b"], nd["c"]))
    return el * es / 6.4e12 + 8 * B * M * N * K / 37e12
nodes.sort(key=lambda nd: -weight(nd))
if mnk:
    nodes = [nd for nd in nodes if tuple(nd["sizes"][1:]) == mnk]
seen, picked = set(), []
for nd in nodes:
    key = (nd["sizes"], int(nd["plan"].variant))
    if key in seen and nd["sizes"][2] > 8:
        continue
    seen.add(key)
    picked.append(nd)
    if len(picked) >= topk:
        break
tdt = getattr(torch, dtype)
lib = _lib.load()
torch.manual_seed(0)
cudart = torch.cuda.cudart()
print(f"{'M':>6} {'N':>4} {'K':>4} var tile            ms     TF/s    GB/s(ideal)")
for nd in picked:
    B, M, N, K = nd["sizes"]
    def buf(t):
        n = int(np.prod(t.shape))
        x = torch.empty(n, dtype=tdt, device="cuda")
        torch.view_as_real(x).normal_() if x.is_complex() else x.normal_()
        return x
    a, b, c = buf(nd["a"]), buf(nd["b"]), buf(nd["c"])
    W = np.array(nd["words"], dtype=np.int64)
    W[L.W_FLAGS] &= ~1  # standalone launch: no accumulate
    W[L.W_CELEMS] = c.numel()
    def launch():
        _lib.check(lib.ctgb_contract_pair(W.ctypes.data, a.data_ptr(), b.data_ptr(), c.data_ptr(), 0))
    for _ in range(2):
        launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        launch()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    el = a.numel() + b.numel() + c.numel()
    print(f"2^{int(np.log2(M)):<4} {N:>4} {K:>4} {int(W[L.W_VARIANT]):>3} ({int(W[9])},{int(W[10])},{int(W[11])})".ljust(36)
          + f"{ms:8.3f} {8*B*M*N*K/ms/1e9:8.2f} {el*es/ms/1e6:9.0f}", flush=True)
    if use_ncu:
        cudart.cudaProfilerStart()
        launch()
        torch.cuda.synchronize()
        cudart.cudaProfilerStop()
    del a, b, c
