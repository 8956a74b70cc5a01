"""One slice of the Sycamore-m20 Appendix-B tree bracketed by cudaProfilerStart/Stop, for

    ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        --clock-control none --csv --log-file out.csv python scripts/gpu_slice_dram.py [dtype] [--nofuse]

(the DRAM traffic of the whole executed plan against the algorithmic bytes of the reference's
unfused tree: VERDICT r1 item 3)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cotengra_b200 as cb
from tests.helpers import make_arrays
from tests.slicing_util import appxB_at_width

dtype = next((a for a in sys.argv[1:] if not a.startswith("--")), "complex128")
spec = appxB_at_width(30)
ex = cb.TreeExecutor(spec, dtype=dtype, fuse="--nofuse" not in sys.argv)
arrays = make_arrays(spec.shapes(), dtype, seed=0, scale=0.65)
dev = [torch.from_numpy(a).cuda() for a in arrays]
for i in range(2):
    ex.contract_device(dev, begin=i, step=1, count=1)
torch.cuda.synchronize()
macs, _inv, elems = ex.reference_work
print(f"reference tree: {macs:.6g} MACs/slice, {elems * ex.plan.esize:.6g} algorithmic bytes/slice; "
      f"executed plan: {ex.plan.elements_per_slice * ex.plan.esize:.6g} bytes/slice", flush=True)
cudart = torch.cuda.cudart()
cudart.cudaProfilerStart()
ex.contract_device(dev, begin=5, step=1, count=1)
torch.cuda.synchronize()
cudart.cudaProfilerStop()
