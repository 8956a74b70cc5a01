"""torchrun --nproc-per-node N scripts/gpu_dist_check.py
N-GPU contract_distributed (NCCL) == golden value of the unmodified reference,
with and without stripped exponents, all-reduce and reduce-to-root."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import cotengra_b200 as cb
from tests.helpers import decode_sliced, load_json, load_npz, make_arrays, rel_err

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
vals = load_npz("trees_values.npz")
ok = True
for name in ("lattice6x6_d3_sliced", "lattice4x4_sliced"):
    rec = next(r for r in load_json("trees.json") if r["name"] == name)
    n_in = len(rec["inputs"])
    node_inds = {int(k): v for k, v in rec["inds"].items() if int(k) >= n_in}
    spec = cb.TreeSpec(rec["inputs"], rec["output"], rec["size_dict"], rec["path"],
                       decode_sliced(rec["sliced"]), node_inds)
    if spec.nslices < world:
        continue
    arrays = make_arrays(spec.shapes(), rec["dtype"], seed=rec["seed"])
    want = vals[name]
    got = cb.contract_distributed(spec, arrays)
    e1 = rel_err(got, want)
    m, e = cb.contract_distributed(spec, arrays, strip_exponent=True)
    e2 = rel_err(m * 10.0**e, want)
    r0 = cb.contract_distributed(spec, arrays, root=0)
    e3 = rel_err(r0, want) if rank == 0 else (0.0 if r0 is None else 1.0)
    if rank == 0:
        print(f"{name}: world={world} allreduce={e1:.1e} stripped={e2:.1e} reduce_root={e3:.1e}")
    ok = ok and max(e1, e2, e3) < 1e-10
# sliced OUTPUT indices sharded over the ranks (beyond contract_mpi, which refuses them):
# every rank scatters its slices into a zeroed full-size output, one all-reduce assembles it
for name in ("rand_r3_o1_hi0_ho1_None_s42_sliced_out", "rand_r2_o1_hi1_ho2_root_s42_sliced_out",
             "rand_r3_o2_hi0_ho1_None_s7_sliced_out"):
    rec = next(r for r in load_json("trees.json") if r["name"] == name)
    n_in = len(rec["inputs"])
    node_inds = {int(k): v for k, v in rec["inds"].items() if int(k) >= n_in}
    spec = cb.TreeSpec(rec["inputs"], rec["output"], rec["size_dict"], rec["path"],
                       decode_sliced(rec["sliced"]), node_inds)
    arrays = make_arrays(spec.shapes(), rec["dtype"], seed=rec["seed"])
    want = vals[name]
    got = cb.contract_distributed(spec, arrays)
    e1 = rel_err(got, want)
    r0 = cb.contract_distributed(spec, arrays, root=world - 1)
    e2 = rel_err(r0, want) if rank == world - 1 else (0.0 if r0 is None else 1.0)
    try:
        cb.contract_distributed(spec, arrays, strip_exponent=True)
        e3 = 1.0
    except NotImplementedError:
        e3 = 0.0
    if rank == 0:
        print(f"{name}: world={world} nslices={spec.nslices} sliced-output allreduce={e1:.1e} "
              f"reduce_root={e2:.1e} stripped_refused={e3 == 0.0}")
    ok = ok and max(e1, e2, e3) < 1e-10
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DIST_CHECK", "PASS" if int(t.item()) == 1 else "FAIL")
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
