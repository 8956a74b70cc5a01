"""Reference-facing host API: the same names, argument meaning and error
behaviour as cotengra's execution path, backed by the sm_100a kernels.

    einsum(eq, a, b=None) ............ cotengra/contract.py:414
    tensordot(a, b, axes) ............ cotengra/contract.py:521
    implementation() ................. the ``(einsum, tensordot)`` pair accepted by
                                       ``implementation=`` (contract.py:775-776; the
                                       order really is einsum first)
    B200Contractor ................... cotengra/contract.py:654 ``Contractor`` /
                                       :840 ``CuQuantumContractor`` (whole-tree)
    contract_tree(tree, arrays) ...... cotengra/core.py:3943 ``ContractionTree.contract``
    contract_distributed(...) ........ cotengra/core.py:4032 ``contract_mpi`` (NCCL)
    gen_output_chunks(tree, arrays) .. cotengra/core.py:3884 ``gen_output_chunks``
    install(tree) .................... seeds ``tree.contraction_cores`` (core.py:3699)

Arrays may be numpy arrays (copied to the GPU and back: the host path) or torch
CUDA tensors (used in place).  torch is only the carrier of device memory and
streams; all arithmetic happens in ``libctgb200.so``.  There is no CPU
fallback: without a CUDA device these functions raise.
"""

from __future__ import annotations

import ctypes as C
import functools
import math

import numpy as np

from . import _lib, lowering
from .executor import ExecPlan, output_chunking
from .lowering import (
    build_pair_desc,
    build_single_desc,
    check_tensordot_shapes,
    classify_pair,
    classify_single,
    dtype_name,
    split_equation,
    tensordot_terms,
)
from .tree import TreeSpec


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError(
            "cotengra_b200 needs a CUDA device (sm_100a); there is no CPU fallback"
        )
    return torch


_NP2T = {"float32": "float32", "float64": "float64", "complex64": "complex64", "complex128": "complex128"}


def _to_device(x, device=None):
    """numpy / torch -> contiguous torch CUDA tensor; returns (tensor, was_numpy)."""
    torch = _torch()
    if isinstance(x, torch.Tensor):
        if not x.is_cuda:
            x = x.cuda(device)
        return x.contiguous(), False
    x = np.asarray(x, order="C")  # (ascontiguousarray would promote 0-d to 1-d)
    dtype_name(x.dtype)
    return torch.from_numpy(x).cuda(device), True


def _from_device(t, as_numpy):
    return t.cpu().numpy() if as_numpy else t


def _stream_ptr():
    torch = _torch()
    return torch.cuda.current_stream().cuda_stream


def _common_dtype(*ts):
    names = {dtype_name(t.dtype) for t in ts}
    if len(names) != 1:
        raise TypeError(f"operands must share one dtype, got {sorted(names)}")
    return names.pop()


# ---------------------------------------------------------------------------
# single nodes (the tuple interface)
# ---------------------------------------------------------------------------


@functools.lru_cache(4096)
def _pair_words(term_a, shape_a, term_b, shape_b, out, dtype, sm_count):
    dims = classify_pair(term_a, shape_a, term_b, shape_b, out)
    plan = build_pair_desc(dims, dtype, sm_count=sm_count,
                           c_dense_elems=math.prod(dims.out_shape))
    return plan, dims.out_shape


@functools.lru_cache(4096)
def _single_words(term, shape, out, dtype):
    odims, sdims, oshape = classify_single(term, shape, out)
    return build_single_desc(odims, sdims, dtype), oshape


def _sm_count():
    return _lib.device_info()["sm_count"]


def _run_pair(term_a, a, term_b, b, out):
    torch = _torch()
    ta, na = _to_device(a)
    tb, nb = _to_device(b, ta.device)
    dtype = _common_dtype(ta, tb)
    plan, oshape = _pair_words(tuple(term_a), tuple(ta.shape), tuple(term_b), tuple(tb.shape),
                               tuple(out), dtype, _sm_count())
    c = torch.empty(oshape, dtype=ta.dtype, device=ta.device)
    if c.numel():
        pa, pb = (tb, ta) if plan.swapped else (ta, tb)
        with torch.cuda.device(ta.device):
            _lib.check(_lib.load().ctgb_contract_pair(
                plan.words.ctypes.data, pa.data_ptr(), pb.data_ptr(), c.data_ptr(), _stream_ptr()))
    return _from_device(c, na and nb)


def _run_single(term, x, out):
    torch = _torch()
    tx, nx = _to_device(x)
    dtype = dtype_name(tx.dtype)
    words, oshape = _single_words(tuple(term), tuple(tx.shape), tuple(out), dtype)
    c = torch.empty(oshape, dtype=tx.dtype, device=tx.device)
    if c.numel():
        with torch.cuda.device(tx.device):
            _lib.check(_lib.load().ctgb_reduce_single(
                words.ctypes.data, tx.data_ptr(), c.data_ptr(), _stream_ptr()))
    return _from_device(c, nx)


def einsum(eq, a, b=None, *, backend=None):
    """Single or pairwise einsum (cotengra/contract.py:414-459), one kernel."""
    terms, out = split_equation(eq)
    if b is None:
        if len(terms) != 1:
            raise ValueError(f"equation {eq!r} needs {len(terms)} operands, got 1")
        return _run_single(terms[0], a, out)
    if len(terms) != 2:
        raise ValueError(f"equation {eq!r} needs {len(terms)} operands, got 2")
    return _run_pair(terms[0], a, terms[1], b, out)


def tensordot(a, b, axes=2, *, backend=None):
    """Tensordot (cotengra/contract.py:521-570), one kernel."""
    na, nb = len(a.shape), len(b.shape)
    try:
        axes = tuple(map(int, axes[0])), tuple(map(int, axes[1]))
    except (IndexError, TypeError):
        n = int(axes)
        axes = tuple(range(na - n, na)), tuple(range(n))
    check_tensordot_shapes(axes, tuple(a.shape), tuple(b.shape))
    ta, tb, to = tensordot_terms(axes, na, nb)
    return _run_pair(ta, a, tb, b, to)


def implementation():
    """The ``(einsum, tensordot)`` pair for cotengra's ``implementation=`` kwarg
    or ``set_default_implementation`` (contract.py:13-31, 775-776)."""
    return (einsum, tensordot)


# ---------------------------------------------------------------------------
# whole-tree contractor
# ---------------------------------------------------------------------------


class TreeExecutor:
    """A compiled sliced contraction: ``ContractionTree.contract`` on the GPU.

    Built from a ``TreeSpec`` or from a live cotengra tree (captured through
    ``TreeSpec.from_cotengra``).  The slice loop, the node loop, the slice
    accumulation and (optionally) exponent stripping all run inside
    ``ctgb_plan_execute``.
    """

    def __init__(self, tree, dtype="complex128", strip_exponent=False, device=None,
                 contractions=None, fuse=True, **plan_opts):
        self.spec = tree if isinstance(tree, TreeSpec) else TreeSpec.from_cotengra(tree)
        # stem fusion (fusion.py): an execution-plan transformation of the tree cotengra found --
        # big stem tensors absorb pre-contracted groups of small tensors in one pass.  ``spec``
        # stays the caller's tree; ``exec_spec`` is what runs.  ``fuse=False`` executes the
        # reference's own node sequence one to one.
        self.exec_spec, self.fusion = self.spec, {"changed": False}
        if fuse and contractions is None:
            from .fusion import fuse_stems

            # (``fuse`` may be a dict of planner options -- min_big, ratio, min_gain, model -- e.g. to
            # force fusion on small trees in tests)
            opts = fuse if isinstance(fuse, dict) else {}
            self.exec_spec, self.fusion = fuse_stems(self.spec, dtype_name(dtype), **opts)
        ir = self.exec_spec.contractions() if contractions is None else contractions
        torch = _torch()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        with torch.cuda.device(self.device):
            self.plan = ExecPlan(ir, self.spec.inputs, self.spec.output, self.spec.size_dict,
                                 self.spec.sliced, dtype=dtype, strip_exponent=strip_exponent,
                                 **plan_opts).create()
        self.dtype = self.plan.dtype
        self.strip_exponent = bool(strip_exponent)
        self._ws = None
        self._ref_work = None

    @property
    def reference_work(self):
        """``(macs_per_slice, macs_invariant, elements_per_slice)`` of the caller's (unfused)
        tree -- the algorithmic work throughput figures are quoted on."""
        if self._ref_work is None:
            if self.fusion.get("changed"):
                from .fusion import tree_work

                self._ref_work = tree_work(self.spec)
            else:
                self._ref_work = (self.plan.macs_per_slice, self.plan.macs_invariant,
                                  self.plan.elements_per_slice)
        return self._ref_work

    @property
    def nslices(self):
        return self.plan.nslices

    def workspace(self, host_staging=False):
        torch = _torch()
        need = self.plan.total_bytes + (self.plan.host_staging_bytes() if host_staging else 0)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def _check_inputs(self, arrays):
        shapes = self.spec.shapes()
        if len(arrays) != len(shapes):
            raise ValueError(f"expected {len(shapes)} arrays, got {len(arrays)}")
        for i, (x, s) in enumerate(zip(arrays, shapes)):
            if tuple(x.shape) != tuple(s):
                raise ValueError(f"array {i} has shape {tuple(x.shape)}, expected {tuple(s)}")

    def contract_device(self, tensors, begin=0, step=1, count=None, out=None, exponent=None):
        """Accumulate slices ``begin, begin+step, ...`` (``count`` of them) of
        already-resident CUDA tensors into ``out`` (zeroed here if not given).
        Returns ``out`` or ``(out, exponent_tensor)``; asynchronous."""
        torch = _torch()
        self._check_inputs(tensors)
        begin, step, count = self._check_slice_range(begin, step, count)
        tdt = getattr(torch, _NP2T[self.dtype])
        with torch.cuda.device(self.device):
            if out is None:
                out = torch.zeros(self.plan.out_shape, dtype=tdt, device=self.device)
            elif (out.device != self.device or not out.is_contiguous() or out.dtype != tdt
                  or tuple(out.shape) != tuple(self.plan.out_shape)):
                raise ValueError("out must be a contiguous tensor of the plan's output shape, dtype and device")
            if self.strip_exponent and exponent is None:
                exponent = torch.full((1,), -math.inf, dtype=torch.float64, device=self.device)
            ws = self.workspace()
            ptrs, keep = [], []
            for i, t in enumerate(tensors):
                if dtype_name(t.dtype) != self.dtype:
                    raise TypeError(f"plan was built for {self.dtype}, got {t.dtype}")
                if t.device != self.device:
                    raise ValueError(f"array {i} lives on {t.device}, the plan on {self.device}")
                if not t.is_contiguous():
                    # the kernels address row-major storage by the plan's own strides
                    t = t.contiguous()
                    keep.append(t)
                ptrs.append(t.data_ptr())
            self.plan.execute(ptrs, out.data_ptr(), exponent.data_ptr() if exponent is not None else None,
                              ws.data_ptr(), ws.numel(), begin, step, count, _stream_ptr())
        return (out, exponent) if self.strip_exponent else out

    def _check_slice_range(self, begin, step, count):
        """Slice ids ``begin, begin+step, ...`` must all lie in ``[0, nslices)``: the device
        decodes digits modulo the radices, so an id past the end would silently wrap and
        count a slice twice."""
        begin, step = int(begin), int(step)
        if step < 1 or begin < 0:
            raise ValueError(f"bad slice range begin={begin} step={step}")
        if count is None:
            count = max(0, -(-(self.nslices - begin) // step))
        count = int(count)
        if count < 0 or (count > 0 and begin + (count - 1) * step >= self.nslices):
            raise ValueError(
                f"slice ids {begin}..{begin + (count - 1) * step} (step {step}) exceed the tree's "
                f"{self.nslices} slices")
        return begin, step, count

    def contract_host(self, arrays, begin=0, step=1, count=None):
        """End-to-end with HOST buffers through ``ctgb_plan_execute_host``:
        H2D of the inputs, all slices, D2H of the result, synchronised."""
        torch = _torch()
        self._check_inputs(arrays)
        begin, step, count = self._check_slice_range(begin, step, count)
        host = [np.asarray(a, dtype=self.dtype, order="C") for a in arrays]
        out = np.zeros(self.plan.out_shape, dtype=self.dtype)
        with torch.cuda.device(self.device):
            ws = self.workspace(host_staging=True)
            e = self.plan.execute_host(host, out, ws.data_ptr(), ws.numel(), begin, step, count,
                                       _stream_ptr())
        return (out, e) if self.strip_exponent else out

    def __call__(self, arrays, **kw):
        return contract_tree(self, arrays, **kw)

    # ------------------------------------------------------------------ output chunks
    def _chunk_plan(self):
        """A second plan over the same program whose output is ONE chunk: the output term
        without its sliced indices, so that every sliced index is summed.  Executed over the
        ``stepsize`` consecutive slice ids of a chunk (core.py:3916-3935)."""
        if getattr(self, "_chunk", None) is None:
            torch = _torch()
            spec = self.exec_spec
            chunk_out, _step, _n = output_chunking(spec)
            with torch.cuda.device(self.device):
                self._chunk = ExecPlan(spec.contractions(), spec.inputs, chunk_out, spec.size_dict,
                                       spec.sliced, dtype=self.dtype,
                                       strip_exponent=self.strip_exponent).create()
        return self._chunk

    def gen_output_chunks(self, arrays, with_key=False):
        """``tree.gen_output_chunks(arrays, with_key)`` (cotengra/core.py:3884-3941): yield
        every output chunk -- one per setting of the sliced *output* indices -- after its
        inner slices have been summed on the device, without ever forming the full output.
        Like the reference this needs the sliced indices ordered output-first (the default
        order, core.py:99-104); unlike it, ``strip_exponent`` chunks are summed with the
        exponent-aware adder and come as ``(mantissa, exponent)``.
        numpy in -> numpy chunks; torch CUDA in -> torch CUDA chunks (a fresh tensor each)."""
        torch = _torch()
        self._check_inputs(arrays)
        spec = self.spec
        _chunk_out, stepsize, nchunks = output_chunking(spec)
        plan = self._chunk_plan()
        all_numpy = all(not isinstance(a, torch.Tensor) for a in arrays)
        tensors = [_to_device(a, self.device)[0] for a in arrays]
        ptrs = [t.data_ptr() for t in tensors]
        tdt = getattr(torch, _NP2T[self.dtype])
        need = plan.total_bytes
        with torch.cuda.device(self.device):
            ws = torch.empty(max(need, 1), dtype=torch.uint8, device=self.device)
        for o in range(nchunks):
            with torch.cuda.device(self.device):
                out = torch.zeros(plan.out_shape, dtype=tdt, device=self.device)
                exp = (torch.full((1,), -math.inf, dtype=torch.float64, device=self.device)
                       if self.strip_exponent else None)
                plan.execute(ptrs, out.data_ptr(), exp.data_ptr() if exp is not None else None,
                             ws.data_ptr(), ws.numel(), o * stepsize, 1, stepsize, _stream_ptr())
            chunk = _from_device(out, all_numpy)
            if self.strip_exponent:
                chunk = (chunk, float(exp.item()))
            if with_key:
                key = {ix: x for ix, x in spec.slice_key(o * stepsize).items() if ix in spec.output}
                yield chunk, key
            else:
                yield chunk


def contract_tree(tree, arrays, strip_exponent=False, check_zero=False, dtype=None,
                  slice_ids=None, **plan_opts):
    """``tree.contract(arrays)`` (cotengra/core.py:3943): takes the *unsliced*
    arrays, handles slicing, contraction and gathering, returns the output in
    ``tree.output`` order -- or ``(mantissa, exponent)`` with ``strip_exponent``.
    numpy in -> numpy out; torch CUDA in -> torch CUDA out."""
    torch = _torch()
    if isinstance(tree, TreeExecutor):
        ex = tree
    else:
        if dtype is None:
            dtype = dtype_name(arrays[0].dtype)
        ex = TreeExecutor(tree, dtype=dtype, strip_exponent=strip_exponent, **plan_opts)
    all_numpy = all(not isinstance(a, torch.Tensor) for a in arrays)
    begin, step, count = (0, 1, None) if slice_ids is None else slice_ids
    if all_numpy:
        res = ex.contract_host(arrays, begin, step, count)
        if ex.strip_exponent:
            m, e = res
            return _finish_stripped(m, e, check_zero)
        return res
    tensors = [_to_device(a, ex.device)[0] for a in arrays]
    res = ex.contract_device(tensors, begin, step, count)
    if ex.strip_exponent:
        m, e = res
        return _finish_stripped(m, float(e.item()), check_zero)
    return res


def gen_output_chunks(tree, arrays, with_key=False, strip_exponent=False, dtype=None, **plan_opts):
    """``tree.gen_output_chunks(arrays, with_key=...)`` (cotengra/core.py:3884-3941) on the
    GPU executor; see ``TreeExecutor.gen_output_chunks``."""
    if isinstance(tree, TreeExecutor):
        ex = tree
    else:
        if dtype is None:
            dtype = dtype_name(arrays[0].dtype)
        ex = TreeExecutor(tree, dtype=dtype, strip_exponent=strip_exponent, **plan_opts)
    yield from ex.gen_output_chunks(arrays, with_key=with_key)


def benchmark(tree, dtype="float64", max_time=60, min_reps=3, max_reps=100, warmup=True,
              executor=None, **plan_opts):
    """``tree.benchmark(dtype, max_time, min_reps, max_reps, warmup)`` (cotengra/core.py:
    4092-4164) on the GPU executor, same protocol and same keys: random inputs, ``warmup``
    untimed slices, then single slices ``i % nslices`` (each one synchronised, as the
    reference's eager numpy calls are) until ``max_time`` seconds or ``max_reps`` repetitions
    are over, but at least ``min_reps``.  Returns ``time_per_slice``, ``est_time_total`` (x
    nslices) and ``est_gigaflops`` with the reference's own flop count
    (``total_flops(dtype)``, core.py:1196-1227: 2 flops per scalar multiply-add for float
    dtypes, 4 for complex ones -- half of the 8-flop convention bench.py reports)."""
    import time

    torch = _torch()
    ex = executor if executor is not None else TreeExecutor(tree, dtype=dtype, **plan_opts)
    tdt = getattr(torch, _NP2T[ex.dtype])
    gen = torch.Generator(device=ex.device)
    gen.manual_seed(0)
    tensors = []
    for shp in ex.spec.shapes():
        t = torch.empty(tuple(shp), dtype=tdt, device=ex.device)
        (torch.view_as_real(t) if t.is_complex() else t).normal_(generator=gen)
        tensors.append(t / max(1.0, float(t.numel()) ** 0.5))
    nslices = int(ex.nslices)
    out = torch.zeros(ex.plan.out_shape, dtype=tdt, device=ex.device)

    def one(i):
        ex.contract_device(tensors, begin=i % nslices, step=1, count=1, out=out)
        torch.cuda.synchronize(ex.device)

    for i in range(int(warmup)):
        one(i)
    t0 = ti = time.time()
    i = 0
    while (ti - t0 < max_time) or (i < min_reps):
        one(i)
        ti = time.time()
        i += 1
        if i >= max_reps:
            break
    time_per_slice = (ti - t0) / i
    est_time_total = time_per_slice * nslices
    per_mac = 4 if "complex" in ex.dtype else 2
    # tree.total_flops(dtype) counts every node of every slice (core.py:1196-1227), the
    # slice-invariant ones included -- each timed single-slice call re-runs them here too
    macs_v, macs_i, _el = ex.reference_work
    total_flops = per_mac * (macs_v + macs_i) * nslices
    return {
        "time_per_slice": time_per_slice,
        "est_time_total": est_time_total,
        "est_gigaflops": total_flops / (1e9 * est_time_total),
    }


def _combine_stripped(m1, e1, m2, e2):
    """``AdderWithMaybeExponentStripped`` (cotengra/core.py:163-170) for two
    (mantissa, exponent) partial sums."""
    if e1 == -math.inf:
        return m2, e2
    if e2 == -math.inf:
        return m1, e1
    e = max(e1, e2)
    return m1 * 10.0 ** (e1 - e) + m2 * 10.0 ** (e2 - e), e


def contract_checkpointed(tree, arrays, checkpoint, every=1024, strip_exponent=False, dtype=None,
                          executor=None, on_block=None, **plan_opts):
    """``tree.contract(arrays)`` for runs too long to lose (partial-sum
    checkpointing; the reference has no equivalent -- ``tree.contract`` restarts at
    slice 0): the slices are contracted in blocks of ``every``, and after each block
    the running sum -- ``(mantissa, exponent)`` with ``strip_exponent``, combined as
    core.py:163-170 -- and the next slice id are written to ``checkpoint`` (.npz,
    atomic replace).  A later call with the same tree, dtype and input values finds
    the file and resumes behind the last completed block; a file written for a
    different tree or different inputs is refused (``ValueError``), never silently
    overwritten.  Slices are independent, so the result equals the uninterrupted
    run up to floating-point summation order.  numpy inputs and outputs (host path).
    ``on_block(next_slice, nslices)`` is called after every saved block."""
    import hashlib
    import os

    if executor is None:
        if dtype is None:
            dtype = dtype_name(arrays[0].dtype)
        executor = TreeExecutor(tree, dtype=dtype, strip_exponent=strip_exponent, **plan_opts)
    ex = executor
    spec = ex.spec
    host = [np.asarray(a, dtype=ex.dtype, order="C") for a in arrays]
    h = hashlib.sha256()
    h.update(spec.to_json().encode())
    h.update(f"|{ex.dtype}|{int(bool(ex.strip_exponent))}|".encode())
    for a in host:
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    tag = h.hexdigest()
    nslices = int(ex.nslices)
    every = max(1, int(every))

    done, total, exponent = 0, None, -math.inf
    if os.path.exists(checkpoint):
        with np.load(checkpoint, allow_pickle=False) as z:
            if str(z["tag"]) != tag:
                raise ValueError(f"{checkpoint} belongs to a different tree, dtype or set of input values")
            done, total, exponent = int(z["next_slice"]), z["partial"], float(z["exponent"])
    while done < nslices:
        count = min(every, nslices - done)
        res = ex.contract_host(host, done, 1, count)
        if ex.strip_exponent:
            m, e = res
            total, exponent = (m, float(e)) if total is None else _combine_stripped(total, exponent, m, float(e))
        else:
            total = res if total is None else total + res
        done += count
        tmp = f"{checkpoint}.tmp.npz"
        np.savez(tmp, tag=np.array(tag), next_slice=np.int64(done), partial=np.asarray(total),
                 exponent=np.float64(exponent))
        os.replace(tmp, checkpoint)
        if on_block is not None:
            on_block(done, nslices)
    return (total, exponent) if ex.strip_exponent else total


def _finish_stripped(m, e, check_zero):
    if check_zero and e == -math.inf:
        # contract.py:819-820
        return 0.0, float("-inf")
    return m, e


class B200Contractor:
    """Drop-in for ``cotengra.contract.Contractor`` (contract.py:654-837): built
    from the reference's contraction records, called with the (already sliced)
    arrays of one slice, returns the output array or ``(mantissa, exponent)``.

    Where the reference walks the records in Python and dispatches three array
    ops per node, this compiles them once per (shapes, dtype) into a ``ctgb_plan``
    and runs the whole node loop in one C call.
    """

    __slots__ = ("contractions", "strip_exponent", "check_zero", "implementation", "backend",
                 "progbar", "_plans", "__weakref__")

    def __init__(self, contractions, strip_exponent=False, check_zero=False,
                 implementation="b200", backend=None, progbar=False):
        self.contractions = tuple(contractions)
        self.strip_exponent = strip_exponent
        self.check_zero = check_zero
        self.implementation = implementation
        self.backend = backend
        self.progbar = progbar
        self._plans = {}

    @classmethod
    def from_tree(cls, tree, **kw):
        """Build from a cotengra tree or a ``TreeSpec`` (records of one slice)."""
        spec = tree if isinstance(tree, TreeSpec) else TreeSpec.from_cotengra(tree)
        return cls(spec.contractions(), **kw)

    def _executor(self, shapes, dtype, strip):
        key = (shapes, dtype, strip, _torch().cuda.current_device())
        ex = self._plans.get(key)
        if ex is None:
            # synthesise a flat (unsliced) network whose inputs are the given arrays
            n_in = len(shapes)
            inputs = [tuple((i, k) for k in range(len(s))) for i, s in enumerate(shapes)]
            size_dict = {(i, k): d for i, s in enumerate(shapes) for k, d in enumerate(s)}
            ex = _FlatExecutor(self.contractions, inputs, size_dict, dtype, strip)
            self._plans[key] = ex
        return ex

    def __call__(self, *arrays, **kwargs):
        kwargs.pop("backend", None)
        kwargs.pop("progbar", None)
        check_zero = kwargs.pop("check_zero", self.check_zero)
        strip_exponent = kwargs.pop("strip_exponent", self.strip_exponent)
        kwargs.pop("implementation", None)
        if kwargs:
            raise TypeError(f"Unknown keyword arguments: {kwargs}.")
        torch = _torch()
        strip = strip_exponent is not False
        devs = [_to_device(a) for a in arrays]
        tensors = [d[0] for d in devs]
        as_numpy = all(d[1] for d in devs)
        dtype = _common_dtype(*tensors)
        with torch.cuda.device(tensors[0].device if tensors else torch.cuda.current_device()):
            ex = self._executor(tuple(tuple(t.shape) for t in tensors), dtype, strip)
            res = ex.run([t.to(ex.device) for t in tensors])
        if strip:
            m, e = res
            e = float(e.item())
            if check_zero and e == -math.inf:
                return 0.0, float("-inf")
            return _from_device(m, as_numpy), e
        return _from_device(res, as_numpy)


class _FlatExecutor:
    """ExecPlan over explicit per-call arrays (no tree-level slicing): the output
    term is whatever the program produces."""

    def __init__(self, contractions, inputs, size_dict, dtype, strip):
        torch = _torch()
        out_shape = _program_output_shape(contractions, [tuple(size_dict[ix] for ix in t) for t in inputs])
        output = tuple(("o", k) for k in range(len(out_shape)))
        sd = dict(size_dict)
        sd.update({("o", k): d for k, d in enumerate(out_shape)})
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.plan = ExecPlan(contractions, inputs, output, sd, (), dtype=dtype,
                             strip_exponent=strip).create()
        self.strip = strip
        self.ws = torch.empty(max(self.plan.total_bytes, 1), dtype=torch.uint8, device=self.device)
        self.tdt = getattr(torch, _NP2T[self.plan.dtype])

    def run(self, tensors):
        torch = _torch()
        with torch.cuda.device(self.device):
            out = torch.zeros(self.plan.out_shape, dtype=self.tdt, device=self.device)
            exp = torch.full((1,), -math.inf, dtype=torch.float64, device=self.device) if self.strip else None
            self.plan.execute([t.data_ptr() for t in tensors], out.data_ptr(),
                              exp.data_ptr() if exp is not None else None, self.ws.data_ptr(),
                              self.ws.numel(), 0, 1, 1, _stream_ptr())
        return (out, exp) if self.strip else out


def _program_output_shape(contractions, shapes):
    """Propagate shapes through the IR (host integer work)."""
    live = {i: tuple(s) for i, s in enumerate(shapes)}
    shp = None
    for p, l, r, tdot, arg, perm in contractions:
        if r is None:
            src = live[p] if l is None else live[l]
            terms, out = split_equation(arg)
            ext = {}
            for ix, d in zip(terms[0], src):
                ext[ix] = d
            shp = tuple(ext[ix] for ix in out)
            live[p] = shp
            continue
        sa, sb = live.pop(l), live.pop(r)
        if tdot:
            axes = (tuple(arg[0]), tuple(arg[1]))
            check_tensordot_shapes(axes, sa, sb)
            ta, tb, to = tensordot_terms(axes, len(sa), len(sb), perm)
        else:
            terms, to = split_equation(arg)
            ta, tb = terms
        shp = classify_pair(ta, sa, tb, sb, to).out_shape
        live[p] = shp
    return shp


def make_contractor(tree, strip_exponent=False, check_zero=False, **_ignored):
    """``cotengra.contract.make_contractor`` for ``implementation="b200"``
    (contract.py:925-1006): the per-slice callable for ``tree``."""
    return B200Contractor.from_tree(tree, strip_exponent=strip_exponent, check_zero=check_zero)


def install(tree, strip_exponent=False, check_zero=False):
    """Route ``tree.contract(...)`` / ``tree.contract_slice(...)`` of a live
    cotengra tree through the B200 contractor by seeding its contractor cache
    (core.py:3699-3711).  Key order: ``(autojit, order, prefer_einsum,
    strip_exponent, check_zero, implementation, progbar)``.  Call after the tree
    is final: slicing/reconfiguration clears the cache (core.py:2040, 2087)."""
    fn = make_contractor(tree, strip_exponent=strip_exponent, check_zero=check_zero)
    key = (False, None, False, bool(strip_exponent), check_zero, None, False)
    tree.contraction_cores[key] = fn
    return fn


# ---------------------------------------------------------------------------
# multi-GPU: slices round-robin over ranks, one reduce at the end
# ---------------------------------------------------------------------------


def contract_distributed(tree, arrays, root=None, group=None, strip_exponent=False,
                         dtype=None, executor=None, **plan_opts):
    """``tree.contract_mpi(arrays, comm, root)`` (cotengra/core.py:4032-4090)
    over ``torch.distributed`` (NCCL on NVLink): rank ``r`` of ``W`` contracts
    slices ``r, r+W, ...`` (core.py:4070), sums them locally on its GPU, then a
    single all-reduce (``root=None``) or reduce (``root=int``) combines the
    partial results.  Refuses fewer slices than ranks as the reference does
    (core.py:4062-4066).  Sliced *output* indices, which ``contract_mpi`` refuses
    (core.py:4051-4055), are sharded too: every rank scatters its
    slices into the chunks of a zeroed full-size output (the executor's root
    strides, core.py:3865-3876), so the same single all-reduce assembles the
    stacked result; only the combination with ``strip_exponent`` stays refused."""
    import torch.distributed as dist

    torch = _torch()
    spec = executor.spec if executor is not None else (
        tree if isinstance(tree, TreeSpec) else TreeSpec.from_cotengra(tree))
    strip = executor.strip_exponent if executor is not None else strip_exponent
    if strip and not {s[0] for s in spec.sliced}.isdisjoint(spec.output):
        raise NotImplementedError(
            "Sliced and output indices overlap - with stripped exponents only a simple "
            "sum of result slices is supported currently."
        )
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    rank_slices(rank, world, spec.nslices)  # raises like core.py:4062-4066
    if executor is None:
        if dtype is None:
            dtype = dtype_name(arrays[0].dtype)
        executor = TreeExecutor(spec, dtype=dtype, strip_exponent=strip_exponent, **plan_opts)
    all_numpy = all(not isinstance(a, torch.Tensor) for a in arrays)
    tensors = [_to_device(a, executor.device)[0] for a in arrays]
    begin, step, count = rank_slices(rank, world, spec.nslices)
    res = executor.contract_device(tensors, begin=begin, step=step, count=count)
    if executor.strip_exponent:
        m, e = reduce_partials(res[0], res[1], root=root, group=group)
        if m is None:
            return None
        return _from_device(m, all_numpy), float(e.item())
    res = reduce_partials(res, None, root=root, group=group)
    if res is None:
        return None
    return _from_device(res, all_numpy)


def rank_slices(rank, world, nslices):
    """Round-robin share of rank ``rank``: slices ``rank, rank+world, ...``
    (core.py:4070) as ``(begin, step, count)``."""
    if nslices < world:
        raise ValueError(
            f"Need to have more slices than MPI processes, but have "
            f"{nslices} and {world} respectively."
        )
    return rank, world, max(0, -(-(nslices - rank) // world))


def reduce_partials(partial, exponent=None, root=None, group=None):
    """Sum the per-rank partial results with ONE collective (core.py:4078-4090):
    all-reduce when ``root is None`` else reduce to ``root`` (other ranks get
    ``None``).  With stripped exponents the pairs are first brought to the
    global maximum exponent (core.py:163-170).  Works on any torch.distributed
    backend (NCCL over NVLink on the GPUs; gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    emax = None
    if exponent is not None:
        emax = exponent.clone()
        dist.all_reduce(emax, op=dist.ReduceOp.MAX, group=group)
        scale = torch.where(torch.isneginf(emax), torch.zeros_like(emax),
                            torch.pow(10.0, exponent - emax))
        rdt = partial.real.dtype if partial.is_complex() else partial.dtype
        partial = partial * scale.to(rdt)
    buf = torch.view_as_real(partial) if partial.is_complex() else partial
    if root is None:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.reduce(buf, dst=root, op=dist.ReduceOp.SUM, group=group)
        if rank != root:
            return None if exponent is None else (None, None)
    return partial if exponent is None else (partial, emax)
