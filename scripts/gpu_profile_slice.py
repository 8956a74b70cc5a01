"""Per-node timing of one slice of the Sycamore-m20 Appendix-B tree (dev tool).

usage: python scripts/gpu_profile_slice.py [dtype] [width_log2] [--nodmma] [--nofuse]
Writes gpurun_out/nodes_<dtype>_w<width>.csv and prints a summary."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import cotengra_b200 as cb
from cotengra_b200 import _lib
from tests.helpers import decode_sliced, load_json, make_arrays
from tests.slicing_util import slice_to_width

dtype = sys.argv[1] if len(sys.argv) > 1 else "complex128"
wlog = int(sys.argv[2]) if len(sys.argv) > 2 else 30
nodmma = "--nodmma" in sys.argv
nofuse = "--nofuse" in sys.argv
rec = next(r for r in load_json("sycamore_m20.json") if r["name"] == "sycamore_m20_appxB")
spec = cb.TreeSpec(rec["inputs"], rec["output"], rec["size_dict"], rec["path"], decode_sliced(rec["sliced"]))
if wlog < 30:
    spec = slice_to_width(spec, 2**wlog)
print("peaks", _lib.probe_fp64_peaks(), flush=True)
t0 = time.time()
ex = cb.TreeExecutor(spec, dtype=dtype, allow_dmma=not nodmma, fuse=not nofuse)
print("fusion", {k: v for k, v in ex.fusion.items()}, flush=True)
plan = ex.plan
print(f"plan built in {time.time()-t0:.1f}s  ws={plan.workspace_bytes/2**30:.2f} GiB persistent={plan.persistent_bytes/2**20:.1f} MiB "
      f"macs/slice={plan.macs_per_slice:.4g} elements/slice={plan.elements_per_slice:.4g}", flush=True)
arrays = make_arrays(spec.shapes(), dtype, seed=0, scale=0.65)
dev = [torch.from_numpy(a).cuda() for a in arrays]
for _ in range(2):
    out = ex.contract_device(dev, begin=0, step=1, count=1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 3
out = ex.contract_device(dev, begin=1, step=1, count=reps)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
flops = 8 * ex.reference_work[0]  # the reference tree's work
es = plan.esize
print(f"slice: {ms:.2f} ms  {flops/ms/1e9:.2f} TFLOP/s  ideal-traffic {plan.elements_per_slice*es/ms/1e6:.1f} GB/s  value={out.cpu().numpy()}", flush=True)
plan.profile(True)
ex.contract_device(dev, begin=5, step=1, count=1)
torch.cuda.synchronize()
times = plan.profile_read()
plan.profile(False)
rows = []
for nd, t in zip(plan.nodes, times):
    if nd["kind"] != 0 or nd["invariant"]:
        continue
    B, M, N, K = nd["sizes"]
    el = sum(int(np.prod(x.shape)) for x in (nd["a"], nd["b"], nd["c"]))
    W = nd["words"]
    rows.append(dict(ms=t, M=M, N=N, K=K, B=B, variant=int(nd["plan"].variant), splitk=int(nd["plan"].splitk),
                     tiles=int(nd["plan"].tiles), tflops=8 * B * M * N * K / (t * 1e9) if t > 0 else 0,
                     gbs=el * es / (t * 1e6) if t > 0 else 0, MTa=int(W[9]), NTa=int(W[10]), KTa=int(W[11])))
rows.sort(key=lambda r: -r["ms"])
tot = sum(r["ms"] for r in rows)
os.makedirs("gpurun_out", exist_ok=True)
tag = f"{dtype}_w{wlog}{'_nodmma' if nodmma else ''}{'_nofuse' if nofuse else ''}"
with open(f"gpurun_out/nodes_{tag}.csv", "w") as f:
    f.write("ms,share,M,N,K,variant,splitk,tiles,MTa,NTa,KTa,tflops,gbs\n")
    for r in rows:
        f.write(f"{r['ms']:.4f},{r['ms']/tot:.4f},{r['M']},{r['N']},{r['K']},{r['variant']},{r['splitk']},{r['tiles']},{r['MTa']},{r['NTa']},{r['KTa']},{r['tflops']:.3f},{r['gbs']:.1f}\n")
print(f"sum of node times {tot:.2f} ms over {len(rows)} nodes")
for r in rows[:25]:
    print(f"  {r['ms']:8.3f} ms {r['ms']/tot:6.1%}  M=2^{np.log2(r['M']):.0f} N={r['N']} K={r['K']} var={r['variant']} splitk={r['splitk']} tile=({r['MTa']},{r['NTa']},{r['KTa']})  {r['tflops']:.2f} TF/s  {r['gbs']:.0f} GB/s")
