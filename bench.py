"""bench.py -- Sycamore n53 m20 sliced-contraction throughput (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--dtype complex128] [--impl reference]

Workload (config.workload): the reference's own benchmark structure
``examples/benchmarks/sycamore_n53_m20_s0_e0_pABCDCDAB.json`` (381 tensors, 754
indices) with the SURVEY.md Appendix-B contraction tree: W = 2^30 elements per
slice, 36 sliced indices (2^36 slices), 4.449e11 scalar MACs per slice -- shipped
as ``tests/golden/sycamore_m20.json``.  Synthetic seeded operands.

A *step* contracts ``--slices-per-gpu`` slices on every GPU (slice ids taken
round-robin over ranks exactly like ``ContractionTree.contract_mpi``,
cotengra/core.py:4070), accumulates them on the device, and (N > 1) sums the
partial outputs with one NCCL all-reduce.  Throughput = 8 * C_slice real flops per
slice (complex multiply-add = 8 flops, docs convention) times
slices, divided by device time (CUDA events, max over ranks).  The whole job has
2^36 slices, so -- exactly as ``tree.benchmark()`` (core.py:4143-4158) -- the
number is measured on a slice sample and the total is an extrapolation
(``config.est_total_hours``).

The reference arm (``--impl reference``) times the CPU restatement of the
reference's numpy path (``oracle/``) on the host cores, on a slice of the same
network sliced further until it fits host memory/time.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "sycamore_n53_m20_sliced_contract_tflops"
UNIT = "TFLOP/s"


def load_spec():
    import cotengra_b200 as cb
    from tests.helpers import decode_sliced, load_json

    rec = next(r for r in load_json("sycamore_m20.json") if r["name"] == "sycamore_m20_appxB")
    spec = cb.TreeSpec(rec["inputs"], rec["output"], rec["size_dict"], rec["path"],
                       decode_sliced(rec["sliced"]))
    return spec, rec


def measured_bf16():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        if "bf16_tflops" in d:
            return float(d["bf16_tflops"]), "MEASURED_PEAKS.json dense bf16 (cuBLAS, burst)"
    return 2250.0, "nominal 2.25 PFLOP/s dense bf16 (MEASURED_PEAKS.json absent)"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "MEASURED_PEAKS.json (driver-measured copy bandwidth)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                     "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        sm = sorted(float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples if len(s) >= 7
                          for n, v in zip(names, s[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]),
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle on the host cores
# ---------------------------------------------------------------------------


def cpu_baseline(dtype, width_log2=24, reps=2):
    """Oracle (port of the reference's numpy path) on a bounded sample: one slice
    of the same network, sliced further to W = 2^width_log2 so that it fits host
    memory and ~10-30 s of CPU work.  Returns (tflops, info)."""
    from oracle import ctg_oracle as orc
    from tests.helpers import make_arrays
    from tests.slicing_util import slice_to_width

    spec, _ = load_spec()
    small = slice_to_width(spec, 2 ** width_log2)
    ir = small.contractions()
    inputs = [tuple(t) for t in small.inputs]
    arrays = make_arrays(small.shapes(), dtype, seed=0, scale=0.65)
    macs, _el = orc.contraction_cost(ir, small.sliced_shapes())
    # warm-up (as tree.benchmark, core.py:4143-4144), then timed repetitions
    orc.run_contractions(ir, orc.slice_arrays(inputs, small.sliced, arrays, 0))
    t0 = time.perf_counter()
    for i in range(reps):
        orc.run_contractions(ir, orc.slice_arrays(inputs, small.sliced, arrays, i + 1))
    dt = (time.perf_counter() - t0) / reps
    tflops = 8 * macs / dt / 1e12
    info = {
        "value": tflops, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
        "sample": (f"{reps} slices of the same m20 network sliced to W=2^{width_log2} "
                   f"({macs:.3g} MACs/slice, {dt:.2f} s/slice), oracle/ctg_oracle.py "
                   f"(numpy {np.__version__}, OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS', 'unset')})"),
    }
    return tflops, dt, info


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_all = time.perf_counter()
    vals = []
    info = None
    for _ in range(max(1, min(args.steps, 3))):
        v, dt, info = cpu_baseline(args.dtype, reps=1)
        vals.append(v)
    value = float(np.median(vals))
    info["value"] = value
    spec, rec = load_spec()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (time.perf_counter() - t_all) / max(1, len(vals)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype_tag(args.dtype),
        "data": "synthetic",
        "config": {"workload": "sycamore_n53_m20 AppxB tree; CPU sample: one slice at W=2^24 per step",
                   "note": "reference = the repo's CPU restatement (oracle/) of cotengra's numpy path; "
                           "cotengra itself is pure Python and needs autoray, absent on the box"},
        "cpu_baseline": info,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def dtype_tag(dtype):
    return {"complex128": "c128 (f64 arithmetic)", "complex64": "c64 (f32 arithmetic)"}.get(dtype, dtype)


# ---------------------------------------------------------------------------
# the GPU arm
# ---------------------------------------------------------------------------


def run_gpu(args):
    import torch
    import torch.distributed as dist

    import cotengra_b200 as cb
    from cotengra_b200 import _lib
    from tests.helpers import make_arrays

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    spec, rec = load_spec()
    ex = cb.TreeExecutor(spec, dtype=args.dtype, device=local)
    plan = ex.plan
    S = args.slices_per_gpu
    # scale keeps the amplitude O(1) (381 factors): finite in complex64 too
    arrays = make_arrays(spec.shapes(), args.dtype, seed=0, scale=0.65)
    tensors = [torch.from_numpy(a).to(dev) for a in arrays]
    tdt = getattr(torch, args.dtype)
    out = torch.zeros(plan.out_shape, dtype=tdt, device=dev)
    ex.workspace(host_staging=True)  # allocate once, outside the timed region

    def step(i):
        # slices base, base+1, ... shared round-robin between the ranks (core.py:4070)
        base = i * S * world
        out.zero_()
        ex.contract_device(tensors, begin=base + rank, step=world, count=S, out=out)
        if world > 1:
            dist.all_reduce(torch.view_as_real(out) if out.is_complex() else out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    plan.profile(True)
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    node_ms = plan.profile_read()
    plan.profile(False)
    clocks = sampler.summary() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    flops_slice = 8 * plan.macs_per_slice
    total_slices = S * world * args.steps
    value = flops_slice * total_slices / (ms * 1e-3) / 1e12

    # ---- end to end through the C-ABI host call (rank-local, then max over ranks):
    # pinned host inputs -> H2D -> slices -> D2H of the result, every step
    pinned = []
    for a in arrays:
        t = torch.empty(a.shape, dtype=tdt).pin_memory()
        t.copy_(torch.from_numpy(a))
        pinned.append(t.numpy())
    h2d = int(sum(a.nbytes for a in pinned))
    d2h = int(plan.out_elements * plan.esize)
    ex.contract_host(pinned, begin=rank, step=world, count=1)  # warm
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for i in range(e2e_steps):
        base = (args.warmup + args.steps + i) * S * world
        res = ex.contract_host(pinned, begin=base + rank, step=world, count=S)
        if world > 1:
            r = torch.from_numpy(np.asarray(res)).to(dev)
            dist.all_reduce(torch.view_as_real(r) if r.is_complex() else r)
            res = r.cpu()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = flops_slice * S * world * e2e_steps / e2e_s / 1e12

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (per-node CUDA events of the last slice)
    peaks = _lib.probe_fp64_peaks()
    hbm_peak, hbm_src = measured_peaks()
    pair_nodes = [(nd, t) for nd, t in zip(plan.nodes, node_ms) if nd["kind"] == 0 and t > 0]
    nd, t_ms = max(pair_nodes, key=lambda x: x[1])
    Bn, M, N, K = nd["sizes"]
    el = sum(int(np.prod(x.shape)) for x in (nd["a"], nd["b"], nd["c"]))
    node_flops = 8.0 * Bn * M * N * K
    node_bytes = el * plan.esize
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_top_kernel.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get(args.dtype, {}).get("dram_bytes_per_launch")
    fp64 = args.dtype in ("complex128", "float64")
    achieved_tf = node_flops / (t_ms * 1e-3) / 1e12
    achieved_gbs = node_bytes / (t_ms * 1e-3) / 1e9
    if fp64:
        tensor_peak = peaks["dmma_tflops"]
        tensor_src = ("fp64 DMMA microbenchmark run in this process (ctgb_probe_fp64_peaks); "
                      "MEASURED_PEAKS.json holds no fp64 figure")
    else:
        # complex64 runs as three kind::tf32 passes over the real embedding (8C real flops each):
        # effective peak = dense TF32 peak / 3, dense TF32 = half the measured dense bf16 figure
        bf16, bf16_src = measured_bf16()
        tensor_peak = bf16 / 2.0 / 3.0
        tensor_src = f"{bf16_src} / 2 (tf32) / 3 (3xTF32 passes)"
    frac_tensor, frac_hbm = achieved_tf / tensor_peak, achieved_gbs / hbm_peak
    tensor_bound = frac_tensor >= frac_hbm
    roofline = {
        # the binding roofline of the dominant node: whichever of the two it sits closer to
        "bound": "tensor" if tensor_bound else "hbm",
        "kernel": f"node M={M} N={N} K={K} (variant {int(nd['plan'].variant)})",
        "achieved": achieved_tf if tensor_bound else achieved_gbs,
        "peak": tensor_peak if tensor_bound else hbm_peak,
        "unit": "TFLOP/s" if tensor_bound else "GB/s",
        "peak_source": tensor_src if tensor_bound else hbm_src,
        "frac_tensor": frac_tensor, "frac_hbm": frac_hbm,
        "launch_ms": t_ms,
        "share_of_slice": t_ms / sum(t for _n, t in pair_nodes),
        "algorithmic_bytes": node_bytes,
        "algorithmic_flops": node_flops,
        "traffic": traffic,
    }
    roofline["frac"] = roofline["achieved"] / roofline["peak"]
    slice_ms = ms / (S * args.steps)
    whole = {
        "slice_ms": slice_ms,
        "tflops": flops_slice / (slice_ms * 1e-3) / 1e12,
        "frac_of_fp64_tensor_peak": (flops_slice / (slice_ms * 1e-3) / 1e12) / peaks["dmma_tflops"] if fp64 else None,
        "hbm_achieved_gbs": plan.elements_per_slice * plan.esize / (slice_ms * 1e-3) / 1e9,
        "hbm_peak_gbs": hbm_peak, "hbm_peak_source": hbm_src,
        "hbm_frac": plan.elements_per_slice * plan.esize / (slice_ms * 1e-3) / 1e9 / hbm_peak,
        "fp64_peaks_measured": peaks,
    }

    cpu = None
    if world == 1 and not args.no_cpu:
        _v, _dt, cpu = cpu_baseline(args.dtype)

    nslices = spec.nslices
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": dtype_tag(args.dtype), "data": "synthetic",
        "config": {
            "workload": ("sycamore_n53_m20 amplitude (examples/benchmarks/sycamore_n53_m20_s0_e0_pABCDCDAB.json), "
                         "SURVEY Appendix-B tree: W=2^30, 36 sliced indices (2^36 slices), "
                         f"{plan.macs_per_slice:.4g} MACs/slice; sample of the slice stream"),
            "slices_per_step": S * world, "slices_per_gpu_per_step": S,
            "parallelism": f"slices round-robin over {world} GPU(s), one NCCL all-reduce per step",
            "l2": "inputs larger than L2 (per-slice intermediates of 2-16 GiB stream through HBM)",
            "flop_convention": "8*C real flops per complex MAC (4*C figure = value/2)",
            "est_total_hours": slice_ms * 1e-3 * nslices / world / 3600.0,
            "hoisted_invariant_nodes": sum(1 for n_ in plan.nodes if n_["invariant"]),
            "workspace_gib": plan.total_bytes / 2**30,
        },
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "TreeExecutor.contract_host -> ctgb_plan_execute_host (pinned host inputs)"},
        "gpu_launches": launches,
        "result_finite": bool(torch.isfinite(torch.view_as_real(out)).all().item()),
        "clocks": clocks,
        "roofline": roofline,
        "whole_slice": whole,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="complex128", choices=["complex128", "complex64"])
    ap.add_argument("--slices-per-gpu", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
