"""Whole-tree execution plans: the reference's linear contraction IR
(cotengra/contract.py:573-651) compiled into one ``ctgb_plan`` whose node loop
(contract.py:791-832) and slice loop (core.py:4015-4030) run in C++/CUDA.

Planning is host-side integer work:

* shapes are propagated through the IR and every node is lowered to a strided
  descriptor (``lowering.py``) -- sliced inputs are *views* of the unsliced
  arrays (base offset per slice + strides of the kept axes, core.py:3811-3817),
  so slicing moves no data;
* nodes whose subtree touches no sliced input are marked slice-invariant and run
  once per execute call into a persistent arena (the reference recontracts them
  for every slice, core.py:4015-4028);
* intermediates are placed in a workspace arena by liveness (children die when
  their parent is formed, contract.py:806-807);
* the root writes straight into the output view selected by the digits of
  sliced *output* indices (gather_slices' stack, core.py:3865-3876) and
  accumulates over inner sliced indices (core.py:3842-3844).
"""

from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib, lowering
from .lowering import (
    DTYPE_CODES,
    DTYPE_SIZES,
    build_pair_desc,
    build_single_desc,
    check_tensordot_shapes,
    classify_pair,
    classify_single,
    dtype_name,
    row_major_strides,
    split_equation,
    tensordot_terms,
)

ALIGN = 256


def _align(x):
    return (x + ALIGN - 1) // ALIGN * ALIGN


class _Arena:
    """Offset allocator over a linear schedule (best fit + coalescing frees)."""

    def __init__(self):
        self.free = []  # sorted (offset, size)
        self.top = 0
        self.peak = 0

    def alloc(self, size):
        size = _align(max(size, 1))
        best = None
        for i, (off, sz) in enumerate(self.free):
            if sz >= size and (best is None or sz < self.free[best][1]):
                best = i
        if best is not None:
            off, sz = self.free.pop(best)
            if sz > size:
                self.free.append((off + size, sz - size))
                self.free.sort()
            return off
        # extend, reusing a trailing free block if there is one
        if self.free and self.free[-1][0] + self.free[-1][1] == self.top:
            off, sz = self.free.pop()
            self.top = off + size
        else:
            off = self.top
            self.top += size
        self.peak = max(self.peak, self.top)
        return off

    def release(self, off, size):
        size = _align(max(size, 1))
        self.free.append((off, size))
        self.free.sort()
        merged = []
        for o, s in self.free:
            if merged and merged[-1][0] + merged[-1][1] == o:
                merged[-1] = (merged[-1][0], merged[-1][1] + s)
            else:
                merged.append((o, s))
        self.free = merged


def output_chunking(spec):
    """Host-side geometry of ``gen_output_chunks`` (cotengra/core.py:3884-3941) for a
    ``TreeSpec``: ``(chunk_output, stepsize, nchunks)`` -- the output term of one chunk (the
    sliced output indices removed), the number of consecutive slice ids summed into a chunk
    (product of the inner sliced extents) and the number of chunks.  Raises ``ValueError``
    when the sliced indices are not ordered output-first (core.py:3912-3913)."""
    inner_seen = False
    for ind, _size, _proj in spec.sliced:
        if ind in spec.output:
            if inner_seen:
                raise ValueError("gen_output_chunks needs the sliced indices sorted output-first "
                                 "(core.py:3912-3913)")
        else:
            inner_seen = True
    sliced = {s[0] for s in spec.sliced}
    chunk_out = tuple(ix for ix in spec.output if ix not in sliced)
    stepsize = math.prod(size for ind, size, proj in spec.sliced
                         if ind not in spec.output and proj is None)
    return chunk_out, stepsize, spec.nslices // stepsize


class _T:
    """A tensor slot while planning."""

    __slots__ = ("shape", "strides", "kind", "input_index", "variant", "slot",
                 "nbytes", "offset", "producer", "last_use", "slice_pos", "slice_stride")

    def __init__(self, shape, strides, kind, variant):
        self.shape = tuple(int(d) for d in shape)
        self.strides = list(strides)
        self.kind = kind
        self.variant = variant
        self.input_index = -1
        self.slot = -1
        self.nbytes = 0
        self.offset = 0
        self.producer = -1
        self.last_use = -1
        self.slice_pos = []
        self.slice_stride = []


class ExecPlan:
    """Compile ``contractions`` (reference IR) for fixed input shapes / dtype.

    Parameters
    ----------
    contractions : the reference IR records ``(p, l, r, tdot, arg, perm)``.
    inputs : the index term of every network input (full, unsliced).
    output : the full output term.
    size_dict : extent of every index.
    sliced : ordered ``[(ind, size, project)]`` as ``tree.sliced_inds``.
    """

    def __init__(self, contractions, inputs, output, size_dict, sliced=(), dtype="complex128",
                 strip_exponent=False, hoist=True, allow_dmma=True, sm_count=None,
                 variant=None):
        self.dtype = dtype_name(dtype)
        self.esize = DTYPE_SIZES[self.dtype]
        self.contractions = tuple(contractions)
        self.inputs = [tuple(t) for t in inputs]
        self.output = tuple(output)
        self.size_dict = dict(size_dict)
        self.sliced = [(i, int(s), None if p is None else int(p)) for i, s, p in sliced]
        self.strip_exponent = bool(strip_exponent)
        self.handle = None
        if sm_count is None:
            try:
                sm_count = _lib.device_info()["sm_count"]
            except Exception:
                sm_count = 148
        self.sm_count = sm_count
        self._keep = []
        self._build(hoist, allow_dmma, variant)

    # ------------------------------------------------------------------ build
    def _build(self, hoist, allow_dmma, variant):
        sliced_pos = {ind: j for j, (ind, _s, _p) in enumerate(self.sliced)}
        nslices = math.prod(s for _i, s, p in self.sliced if p is None)
        self.nslices = nslices
        real_slicing = bool(self.sliced)
        hoist = hoist and real_slicing

        # full output geometry (projected output indices keep extent 1)
        def out_ext(ix):
            if ix in sliced_pos and self.sliced[sliced_pos[ix]][2] is not None:
                return 1
            return self.size_dict[ix]

        self.out_shape = tuple(out_ext(ix) for ix in self.output)
        out_full_strides = row_major_strides(self.out_shape)
        self.out_elements = math.prod(self.out_shape)
        root_term = tuple(ix for ix in self.output if ix not in sliced_pos)
        root_strides = [s for ix, s in zip(self.output, out_full_strides) if ix not in sliced_pos]
        self.root_shape = tuple(self.size_dict[ix] for ix in root_term)
        slice_out_stride = [0] * len(self.sliced)
        for ix, s in zip(self.output, out_full_strides):
            if ix in sliced_pos and self.sliced[sliced_pos[ix]][2] is None:
                slice_out_stride[sliced_pos[ix]] += s

        # network inputs as strided views of the unsliced arrays
        cur = {}
        self.input_nbytes = []
        for i, term in enumerate(self.inputs):
            full_shape = [self.size_dict[ix] for ix in term]
            fs = row_major_strides(full_shape)
            self.input_nbytes.append(math.prod(full_shape) * self.esize)
            keep = [k for k, ix in enumerate(term) if ix not in sliced_pos]
            t = _T([full_shape[k] for k in keep], [fs[k] for k in keep], 0,
                   variant=any(ix in sliced_pos for ix in term))
            t.input_index = i
            t.nbytes = self.input_nbytes[-1]  # the whole (unsliced) array: strip_exponent copies it scaled
            for k, ix in enumerate(term):
                if ix in sliced_pos:
                    t.slice_pos.append(sliced_pos[ix])
                    t.slice_stride.append(fs[k])
            cur[i] = t
        self.sliced_shapes = [cur[i].shape for i in range(len(self.inputs))]

        tensors = list(cur.values())
        nodes = []  # dicts
        n_rec = len(self.contractions)
        for step, (p, l, r, tdot, arg, perm) in enumerate(self.contractions):
            is_last = step == n_rec - 1
            if r is None:
                src = cur[p] if l is None else cur[l]
                terms, out = split_equation(arg)
                if len(terms) != 1:
                    raise ValueError(f"expected a single-term equation, got {arg!r}")
                is_root = l is not None
                if is_root and not is_last:
                    raise ValueError("single-input record must be the only contraction")
                ostr = root_strides if (is_root and not self.strip_exponent) else None
                odims, sdims, oshape = classify_single(terms[0], src.shape, out, out_strides=ostr,
                                                      strides_x=src.strides)
                if is_root:
                    self._check_root_shape(oshape)
                # the root always accumulates into the (zeroed) output, so that
                # slice sums, split-K atomics and plain stores share one path
                acc = is_root and not self.strip_exponent
                words = build_single_desc(odims, sdims, self.dtype, accumulate=acc)
                dst = _T(oshape, row_major_strides(oshape), 1, src.variant)
                nodes.append(dict(kind=1, a=src, b=None, c=dst, words=words, root=is_root))
                tensors.append(dst)
                cur[p] = dst
                continue

            A, Bt = cur.pop(l), cur.pop(r)
            if tdot:
                axes = (tuple(arg[0]), tuple(arg[1]))
                check_tensordot_shapes(axes, A.shape, Bt.shape)
                ta, tb, to = tensordot_terms(axes, len(A.shape), len(Bt.shape), perm)
            else:
                terms, to = split_equation(arg)
                if len(terms) != 2:
                    raise ValueError(f"expected a two-term equation, got {arg!r}")
                ta, tb = terms
            is_root = is_last
            ostr = root_strides if (is_root and not self.strip_exponent) else None
            dims = classify_pair(ta, A.shape, tb, Bt.shape, to, out_strides=ostr,
                                 strides_a=A.strides, strides_b=Bt.strides)
            if is_root:
                self._check_root_shape(dims.out_shape)
            acc = is_root and not self.strip_exponent
            dense = 0 if (is_root and not self.strip_exponent) else math.prod(dims.out_shape)
            plan = build_pair_desc(dims, self.dtype, accumulate=acc, sm_count=self.sm_count,
                                   allow_dmma=allow_dmma, c_dense_elems=dense, variant=variant)
            dst = _T(dims.out_shape, row_major_strides(dims.out_shape), 1, A.variant or Bt.variant)
            a, b = (Bt, A) if plan.swapped else (A, Bt)
            nodes.append(dict(kind=0, a=a, b=b, c=dst, words=plan.words, root=is_root, plan=plan,
                              sizes=plan.sizes, dims=dims, acc=acc, dense=dense))
            tensors.append(dst)
            cur[p] = dst

        if not nodes:
            raise ValueError("empty contraction program")
        root_node = nodes[-1]
        if not self.strip_exponent:
            root_node["c"].kind = 3  # writes the output accumulator directly
        # invariance
        for nd in nodes:
            srcs = [nd["a"]] + ([nd["b"]] if nd["b"] is not None else [])
            nd["invariant"] = bool(hoist and not nd["root"] and not any(s.variant for s in srcs))
            if not nd["invariant"]:
                nd["c"].variant = True
        # schedule: invariant pass then variant pass (the C side runs them so)
        order = [nd for nd in nodes if nd["invariant"]] + [nd for nd in nodes if not nd["invariant"]]
        for pos, nd in enumerate(order):
            nd["pos"] = pos
            nd["c"].producer = pos
            for s in (nd["a"], nd["b"]):
                if s is not None:
                    s.last_use = max(s.last_use, pos)
        # persistent tensors read by the variant pass must survive every slice
        for nd in order:
            if not nd["invariant"]:
                for s in (nd["a"], nd["b"]):
                    if s is not None and s.kind == 1 and s.producer >= 0 and order[s.producer]["invariant"]:
                        s.last_use = 1 << 60
        persistent, scratch = _Arena(), _Arena()
        for pos, nd in enumerate(order):
            c = nd["c"]
            arena = persistent if nd["invariant"] else scratch
            if c.kind == 1:
                c.nbytes = max(math.prod(c.shape), 1) * self.esize
                c.offset = arena.alloc(c.nbytes)
                if nd["invariant"]:
                    c.kind = 2
            elif c.kind == 3:
                c.nbytes = max(math.prod(c.shape), 1) * self.esize
            for s in (nd["a"], nd["b"]):
                if s is None or s.last_use != pos:
                    continue
                if s.kind == 1:
                    scratch.release(s.offset, s.nbytes)
                elif s.kind == 2:
                    persistent.release(s.offset, s.nbytes)
        self.workspace_bytes = _align(scratch.peak)
        self.persistent_bytes = _align(persistent.peak)
        self.nodes = nodes
        self.n_variant_nodes = sum(1 for nd in nodes if not nd["invariant"])

        # cost bookkeeping (scalar MACs and ideal element traffic per slice)
        self.macs_per_slice = 0
        self.macs_invariant = 0
        self.elements_per_slice = 0
        for nd in nodes:
            if nd["kind"] != 0:
                continue
            Bn, M, N, K = nd["sizes"]
            macs = Bn * M * N * K
            el = math.prod(nd["a"].shape) + math.prod(nd["b"].shape) + math.prod(nd["c"].shape)
            if nd["invariant"]:
                self.macs_invariant += macs
            else:
                self.macs_per_slice += macs
                self.elements_per_slice += el

        # ---- marshal for the C-ABI
        for i, t in enumerate(tensors):
            t.slot = i
        n_t = len(tensors)
        ct = (_lib.CtgbTensor * n_t)()
        for i, t in enumerate(tensors):
            ct[i].kind = t.kind
            ct[i].input_index = t.input_index
            ct[i].offset = t.offset
            ct[i].nbytes = t.nbytes
            ct[i].n_sliced = len(t.slice_pos)
            if t.slice_pos:
                pos = (C.c_int32 * len(t.slice_pos))(*t.slice_pos)
                st = (C.c_int64 * len(t.slice_stride))(*t.slice_stride)
                self._keep += [pos, st]
                ct[i].slice_pos = C.cast(pos, C.POINTER(C.c_int32))
                ct[i].slice_stride = C.cast(st, C.POINTER(C.c_int64))
        cn = (_lib.CtgbNode * len(nodes))()
        for i, nd in enumerate(nodes):
            words = np.ascontiguousarray(nd["words"], dtype=np.int64)
            self._keep.append(words)
            cn[i].kind = nd["kind"]
            cn[i].a = nd["a"].slot
            cn[i].b = nd["b"].slot if nd["b"] is not None else -1
            cn[i].c = nd["c"].slot
            cn[i].invariant = int(nd["invariant"])
            cn[i].is_root = int(nd["root"])
            cn[i].desc = words.ctypes.data_as(C.POINTER(C.c_int64))
        ns = len(self.sliced)
        radix = (C.c_int64 * max(ns, 1))(*[s for _i, s, _p in self.sliced])
        proj = (C.c_int64 * max(ns, 1))(*[(-1 if p is None else p) for _i, _s, p in self.sliced])
        ostr = (C.c_int64 * max(ns, 1))(*slice_out_stride)
        pd = _lib.CtgbPlanDesc()
        pd.dtype = DTYPE_CODES[self.dtype]
        pd.n_inputs = len(self.inputs)
        pd.n_tensors = n_t
        pd.tensors = C.cast(ct, C.POINTER(_lib.CtgbTensor))
        pd.n_nodes = len(nodes)
        pd.nodes = C.cast(cn, C.POINTER(_lib.CtgbNode))
        pd.n_sliced = ns
        pd.slice_radix = C.cast(radix, C.POINTER(C.c_int64))
        pd.slice_project = C.cast(proj, C.POINTER(C.c_int64))
        pd.slice_out_stride = C.cast(ostr, C.POINTER(C.c_int64))
        pd.out_elements = self.out_elements
        pd.workspace_bytes = self.workspace_bytes
        pd.persistent_bytes = self.persistent_bytes
        pd.strip_exponent = int(self.strip_exponent)
        self._keep += [ct, cn, radix, proj, ostr]
        self._pd = pd
        if self.strip_exponent:
            # dense root result -> its chunk of the (strided) output
            rs = self.root_shape
            dense = row_major_strides(rs)
            odims = [[e, sx, so] for e, sx, so in zip(rs, dense, root_strides) if e != 1]
            self._chunk_words = build_single_desc(odims, [], self.dtype)
        self.total_bytes = self.workspace_bytes + self.persistent_bytes

    def _check_root_shape(self, shape):
        if tuple(shape) != tuple(self.root_shape):
            raise ValueError(
                f"contraction program produces shape {tuple(shape)}, "
                f"tree output expects {tuple(self.root_shape)}"
            )

    # ------------------------------------------------------------------ device side
    def create(self):
        """Upload the plan to the current CUDA device."""
        if self.handle is not None:
            return self
        lib = _lib.load()
        h = C.c_void_p()
        _lib.check(lib.ctgb_plan_create(C.byref(self._pd), C.byref(h)))
        self.handle = h
        if self.strip_exponent:
            w = self._chunk_words
            _lib.check(lib.ctgb_plan_set_chunk_desc(h, w.ctypes.data_as(C.c_void_p)))
        return self

    def host_staging_bytes(self):
        extra = _align(self.total_bytes) - self.total_bytes
        for n in self.input_nbytes:
            extra += _align(n)
        return extra + _align(self.out_elements * self.esize) + 512

    def execute(self, input_ptrs, out_ptr, exp_ptr, ws_ptr, ws_bytes, begin, step, count, stream=0):
        lib = _lib.load()
        arr = (C.c_void_p * len(input_ptrs))(*input_ptrs)
        _lib.check(lib.ctgb_plan_execute(self.handle, arr, out_ptr, exp_ptr, ws_ptr, ws_bytes,
                                         int(begin), int(step), int(count), stream))

    def execute_host(self, host_arrays, host_out, ws_ptr, ws_bytes, begin, step, count, stream=0):
        lib = _lib.load()
        ptrs = (C.c_void_p * len(host_arrays))(*[a.ctypes.data for a in host_arrays])
        nb = (C.c_int64 * len(host_arrays))(*[a.nbytes for a in host_arrays])
        exp = C.c_double(0.0)
        _lib.check(lib.ctgb_plan_execute_host(self.handle, ptrs, nb, host_out.ctypes.data,
                                              C.byref(exp), ws_ptr, ws_bytes, int(begin), int(step),
                                              int(count), stream))
        return exp.value

    def profile(self, enable=True):
        _lib.check(_lib.load().ctgb_plan_profile(self.handle, int(enable)))

    def profile_read(self):
        """Milliseconds of every node (plan order) for the last executed slice."""
        n = len(self.nodes)
        ms = (C.c_float * n)()
        _lib.check(_lib.load().ctgb_plan_profile_read(self.handle, ms, n))
        return [float(x) for x in ms]

    def launches_per_slice(self):
        return int(_lib.load().ctgb_plan_launches_per_slice(self.handle))

    def destroy(self):
        if self.handle is not None:
            _lib.load().ctgb_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
