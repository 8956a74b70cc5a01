"""Host-side lowering of one contraction-tree node to a kernel descriptor.

The reference lowers a pairwise node to ``transpose -> reshape(copy) -> matmul
-> reshape/transpose`` (cotengra/contract.py:167-329 plans it, :364-411 runs
it).  Here the same index classification is done once on the host --

    bat_inds  (on A, B and the output)      contract.py:226-237
    con_inds  (on A and B, not the output)  contract.py:226-237
    a_keep    (on A and the output)         contract.py:238-239
    b_keep    (on B and the output)         contract.py:241-243

-- but instead of permuting data each class becomes a list of *dims with
strides* in A, B and C.  Repeated indices (diagonals) add their strides,
size-1 dims drop out, broadcast dims get stride 0, indices summed on a single
operand become contracted dims with stride 0 on the other operand
(contract.py:193-216, 256-274).  Adjacent dims that stay adjacent in every
operand are coalesced, each class is split into CTA-tile dims and grid dims,
and everything is packed into the int64 word layout of ``csrc/gett_desc.h``.

All of this is integer work and is tested bit-exactly against the reference's
own planners (tests/test_lowering.py).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# ---- word layout: mirror of csrc/gett_desc.h (checked against the built
# library in _lib.py through ctgb_desc_words()) -----------------------------
MAX_T, MAX_G, MAX_GB, MAX_LD = 12, 40, 12, 24
(W_MAGIC, W_DTYPE, W_NTM, W_NTN, W_NTK, W_NGM, W_NGN, W_NGK, W_NGB, W_MTA,
 W_NTA, W_KTA, W_TILES_M, W_TILES_N, W_TILES_B, W_STEPS_K, W_SPLITK, W_PGM,
 W_MFULL, W_MTEXT, W_MW, W_PGN, W_NFULL, W_NTEXT, W_NW, W_PGK, W_KFULL,
 W_KTEXT, W_KW, W_NLDA, W_NLDB, W_FLAGS, W_VARIANT, W_CELEMS) = range(34)
W_HDR = 40
OFF_TM = W_HDR
OFF_TN = OFF_TM + MAX_T * 3
OFF_TK = OFF_TN + MAX_T * 3
OFF_GM = OFF_TK + MAX_T * 3
OFF_GN = OFF_GM + MAX_G * 4
OFF_GK = OFF_GN + MAX_G * 4
OFF_GB = OFF_GK + MAX_G * 4
OFF_LDA = OFF_GB + MAX_GB * 5
OFF_LDB = OFF_LDA + MAX_LD * 4
DESC_WORDS = OFF_LDB + MAX_LD * 4
DESC_MAGIC = 0x4354474232303031

MAX_S = 40
S_MAGIC, S_DTYPE, S_NO, S_NS, S_OUT_ELEMS, S_SUM_ELEMS, S_FLAGS = range(7)
S_HDR = 8
OFF_SO = S_HDR
OFF_SS = OFF_SO + MAX_S * 3
SDESC_WORDS = OFF_SS + MAX_S * 2
SDESC_MAGIC = 0x4354474253303031

VAR_SIMT_64x64, VAR_KRED, VAR_DMMA_128x64, VAR_DMMA_64x128, VAR_DMMA_256x32 = 0, 1, 2, 3, 4
VAR_DMMA_256x16, VAR_ROW_128x8, VAR_ROW_256x4, VAR_ROWSTREAM = 5, 6, 7, 8
VAR_TC05_128x64, VAR_TC05_128x32, VAR_TC05_128x16 = 9, 10, 11
VAR_DMMA3M_128x32, VAR_DMMA3M_256x16, VAR_DMMASTREAM, VAR_DOTSTREAM, VAR_DOTSTREAM4 = 12, 13, 14, 15, 16
VAR_DMMA_32x32, VAR_ROWSTREAM_K = 18, 19
DMMASTREAM_MAX_N = 16  # the kernel takes N <= 32, but at N = 32 the staged 256x32 policy is faster (31.8 vs 26 TFLOP/s)
TC05_VARIANTS = (VAR_TC05_128x64, VAR_TC05_128x32, VAR_TC05_128x16)
TC05_MAX_K = 16384       # 1024 k-steps (the kernel's k table); beyond 256 in chunks of 256
TC05_CHUNK_STEPS = 16    # full k-steps accumulated in TMEM before a round-to-nearest fold (tc05_chunk_steps in
                         # csrc/tc05_kernel.cuh balances the chunks and shortens them for tiles with fewer than 16 k)
# (MT, NT, KT) of every kernel variant -- must match ctg_b200.cu's dispatch
VARIANT_TILES = {
    VAR_SIMT_64x64: (64, 64, 8),
    VAR_KRED: (1, 1, 512),
    VAR_DMMA_128x64: (128, 64, 16),
    VAR_DMMA_64x128: (64, 128, 16),
    VAR_DMMA_256x32: (256, 32, 8),
    VAR_DMMA_256x16: (256, 16, 8),
    VAR_ROW_128x8: (256, 8, 4),
    VAR_ROW_256x4: (256, 4, 4),
    VAR_ROWSTREAM: (256, 8, 8),
    VAR_TC05_128x64: (128, 64, 16),
    VAR_TC05_128x32: (128, 32, 16),
    VAR_TC05_128x16: (128, 16, 16),
    VAR_DMMA3M_128x32: (128, 32, 16),
    VAR_DMMA3M_256x16: (256, 16, 8),
    VAR_DMMASTREAM: (256, 32, 64),
    VAR_DOTSTREAM: (1, 1, 2048),
    VAR_DOTSTREAM4: (4, 4, 1024),
    VAR_DMMA_32x32: (32, 32, 16),
    VAR_ROWSTREAM_K: (256, 8, 64),
}

DTYPE_CODES = {"float32": 0, "float64": 1, "complex64": 2, "complex128": 3}
DTYPE_SIZES = {"float32": 4, "float64": 8, "complex64": 8, "complex128": 16}


def dtype_name(dtype) -> str:
    if isinstance(dtype, str):
        name = dtype
    elif str(dtype).startswith("torch."):
        name = str(dtype)[len("torch."):]
    else:
        name = str(np.dtype(dtype))
    if name not in DTYPE_CODES:
        raise TypeError(f"unsupported dtype {dtype!r}")
    return name


def row_major_strides(shape):
    strides, acc = [0] * len(shape), 1
    for i in range(len(shape) - 1, -1, -1):
        strides[i] = acc
        acc *= int(shape[i])
    return strides


# ---------------------------------------------------------------------------
# classification
# ---------------------------------------------------------------------------


@dataclass
class PairDims:
    """Index classes of one pairwise node; every dim is ``[ext, sA, sB, sC]``
    (strides in elements, 0 where the operand does not carry the index)."""

    batch: list = field(default_factory=list)
    m: list = field(default_factory=list)
    n: list = field(default_factory=list)
    k: list = field(default_factory=list)
    out_shape: tuple = ()

    def sizes(self):
        pr = lambda ds: math.prod(d[0] for d in ds)  # noqa: E731
        return pr(self.batch), pr(self.m), pr(self.n), pr(self.k)


def classify_pair(term_a, shape_a, term_b, shape_b, out, out_strides=None,
                  strides_a=None, strides_b=None):
    """Classify the indices of ``term_a,term_b->out``.

    ``term_*`` / ``out`` are sequences of hashable labels.  Raises the same
    ``ValueError`` conditions as contract.py:183-186, 200-204, 218-222.
    """
    term_a, term_b, out = tuple(term_a), tuple(term_b), tuple(out)
    shape_a, shape_b = tuple(map(int, shape_a)), tuple(map(int, shape_b))
    if len(term_a) != len(shape_a):
        raise ValueError(f"Term '{term_a}' does not match shape {shape_a}.")
    if len(term_b) != len(shape_b):
        raise ValueError(f"Term '{term_b}' does not match shape {shape_b}.")
    sa = row_major_strides(shape_a) if strides_a is None else list(strides_a)
    sb = row_major_strides(shape_b) if strides_b is None else list(strides_b)

    ext, st_a, st_b = {}, {}, {}
    order = []
    for term, shape, strides, acc in ((term_a, shape_a, sa, st_a),
                                      (term_b, shape_b, sb, st_b)):
        for ix, d, s in zip(term, shape, strides):
            if ix not in ext and ix not in order:
                order.append(ix)
            if d == 1:
                continue
            if ext.setdefault(ix, d) != d:
                raise ValueError(
                    f"Index {ix} has mismatched sizes {ext[ix]} and {d}."
                )
            acc[ix] = acc.get(ix, 0) + s
    for ix in out:
        if ix not in order:
            raise ValueError(f"Output index {ix} does not appear in the inputs.")

    out_shape = tuple(ext.get(ix, 1) for ix in out)
    sc_list = row_major_strides(out_shape) if out_strides is None else list(out_strides)
    st_c = {}
    for ix, s in zip(out, sc_list):
        st_c[ix] = st_c.get(ix, 0) + s

    dims = PairDims(out_shape=out_shape)
    for ix in order:
        if ix not in ext:
            continue  # extent 1 everywhere: no loop at all
        rec = [ext[ix], st_a.get(ix, 0), st_b.get(ix, 0), st_c.get(ix, 0)]
        on_a, on_b = ix in st_a, ix in st_b
        if ix in st_c:
            if on_a and on_b:
                dims.batch.append(rec)
            elif on_a:
                dims.m.append(rec)
            else:
                dims.n.append(rec)
        else:
            dims.k.append(rec)
    return dims


def tensordot_terms(axes, ndim_a, ndim_b, perm=None):
    """Integer labels equivalent to ``tensordot(a, b, axes)`` followed by
    ``transpose(perm)`` (contract.py:472-518 builds the same equation out of
    characters; contract.py:811-812 applies the permutation)."""
    ax_a, ax_b = axes
    if len(ax_a) != len(ax_b):
        raise ValueError(
            f"Axes should have the same length, got {ax_a} and {ax_b}."
        )
    term_a = list(range(ndim_a))
    term_b, out = [], list(term_a)
    nxt = ndim_a
    for j in range(ndim_b):
        if j in ax_b:
            ix = term_a[ax_a[ax_b.index(j)]]
            out.remove(ix)
        else:
            ix = nxt
            nxt += 1
            out.append(ix)
        term_b.append(ix)
    if perm is not None:
        out = [out[p] for p in perm]
    return term_a, term_b, out


def check_tensordot_shapes(axes, shape_a, shape_b):
    for i, j in zip(*axes):
        if shape_a[i] != shape_b[j]:
            raise ValueError(
                f"Dimension mismatch between axes {i} of {tuple(shape_a)} and "
                f"{j} of {tuple(shape_b)}: {shape_a[i]} != {shape_b[j]}."
            )


# ---------------------------------------------------------------------------
# coalescing and tiling
# ---------------------------------------------------------------------------


def coalesce(dims):
    """Merge dims that are adjacent (outer stride == inner stride * inner
    extent) in *every* operand: a rank-30 all-dims-2 Sycamore tensor drops to a
    handful of super-dims (SURVEY.md Appx D.5)."""
    dims = [list(d) for d in dims if d[0] != 1]
    changed = True
    while changed:
        changed = False
        for i in range(len(dims)):
            for j in range(len(dims)):
                if i == j:
                    continue
                inner, outer = dims[i], dims[j]
                if all(outer[t] == inner[t] * inner[0] for t in range(1, len(inner))):
                    inner[0] *= outer[0]
                    del dims[j]
                    changed = True
                    break
            if changed:
                break
    return dims


def _min_stride(d, cols):
    vals = [abs(d[c]) for c in cols if d[c] != 0]
    return min(vals) if vals else 0


def _largest_divisor(n, cap, prod=1, multiple=1):
    """Largest divisor t of n, 2 <= t <= cap, with prod * t a multiple of ``multiple`` (1 if none)."""
    for t in range(min(n, cap), 1, -1):
        if n % t == 0 and (prod * t) % multiple == 0:
            return t
    return 1


def split_tile(dims, cols, limit, order_col, exact=False, multiple=1):
    """Pick the tile dims of one class.

    Greedy by smallest stride in any operand carrying the dim (those are the
    dims whose inclusion makes global accesses contiguous); at most one dim is
    taken partially (blocked).  Returns ``(tile, grid, partial)`` with

      tile : [[text, *strides]]           local order, partial dim last
      grid : [[count, *strides_per_step]] remaining loops (block dim included)
      partial : (grid_index, full_ext, text, weight) or None

    ``exact``: the blocked dim is cut into equal blocks (the largest divisor of its extent
    that fits), so that every tile has the same shape -- for kernels without ragged tiles;
    ``multiple``: ... among the divisors that make the tile's extent a multiple of this (the k of a
    tcgen05 tile is whole UMMA k8 groups of complex numbers: 4).
    """
    cand = sorted(range(len(dims)), key=lambda i: (_min_stride(dims[i], cols), i))
    tile, used, prod, partial_src = [], set(), 1, None
    for i in cand:
        e = dims[i][0]
        if len(tile) >= MAX_T - 1:
            break
        if prod * e <= limit:
            tile.append(list(dims[i]))
            used.add(i)
            prod *= e
        else:
            t = limit // prod
            if exact:
                t = _largest_divisor(e, t, prod, multiple)
            if t >= 2:
                rec = list(dims[i])
                rec[0] = t
                partial_src = (i, rec)
                used.add(i)
                prod *= t
            break
    tile.sort(key=lambda d: (abs(d[order_col]) if d[order_col] else 1 << 62))
    grid = [list(dims[i]) for i in range(len(dims)) if i not in used]
    partial = None
    if partial_src is not None:
        i, rec = partial_src
        full, t = dims[i][0], rec[0]
        weight = math.prod(d[0] for d in tile)
        tile.append(rec)
        blocks = -(-full // t)
        grid.append([blocks] + [s * t for s in dims[i][1:]])
        partial = [len(grid) - 1, full, t, weight]
    return tile, grid, partial


def _order_grid(grid, partial, key_col):
    """Fastest-varying grid dim first = smallest stride of the streamed operand
    (consecutive tiles touch neighbouring memory)."""
    idx = sorted(range(len(grid)),
                 key=lambda i: (abs(grid[i][key_col]) if grid[i][key_col] else 1 << 62, i))
    new = [grid[i] for i in idx]
    if partial is not None:
        partial = [idx.index(partial[0])] + partial[1:]
    return new, partial


def _with_divs(grid):
    out, div = [], 1
    for g in grid:
        out.append([g[0], div] + list(g[1:]))
        div *= g[0]
    return out, div


@dataclass
class PairPlan:
    words: np.ndarray
    variant: int
    sizes: tuple  # (B, M, N, K)
    swapped: bool
    tiles: int
    splitk: int


def choose_variant(dtype, B, M, N, K, allow_dmma=True, allow_stream=True, allow_tc05=True, allow_3m=False):
    if M == 1 and N == 1 and B == 1 and K >= 1 << 20 and allow_stream:
        return VAR_DOTSTREAM
    if M <= 4 and N <= 4 and B == 1 and K >= 1 << 20 and allow_stream:
        return VAR_DOTSTREAM4  # a stem tail peeled over the final inner product (fusion.py)
    if (dtype in ("complex128", "float64") and allow_dmma and M <= 32 and N <= 32 and M * N >= 4 and B == 1
            and K >= 1 << 14):
        # the same with a few more peeled tensors (fusion.py): ONE 32 x 32 fp64 tensor-core tile,
        # the contracted range split over two CTAs per SM
        return VAR_DMMA_32x32
    if M == 1 and N == 1 and B == 1 and K >= 8192:
        return VAR_KRED
    if N <= 8 and K <= 8 and B == 1 and 64 <= M < 1 << 32 and allow_stream:
        return VAR_ROWSTREAM
    # narrow complex128 nodes: DMMA fragments streamed from global memory, no staging
    # (N <= 8 with a contracted space too long for the row-stream kernel included)
    if (allow_dmma and allow_stream and dtype == "complex128" and B == 1 and 4096 <= M < 1 << 32
            and ((N <= DMMASTREAM_MAX_N and K <= 32) or (N <= 8 and 8 < K <= 64))):
        return VAR_DMMASTREAM
    # ... and the narrower element types: the row stream walked in chunks of 8 k
    if (allow_stream and DTYPE_SIZES[dtype] <= 8 and N <= 8 and 8 < K <= 64 and B == 1
            and 4096 <= M < 1 << 32):
        return VAR_ROWSTREAM_K
    if N <= 8 and M >= 64:
        return VAR_ROW_256x4 if N <= 4 else VAR_ROW_128x8
    # complex64 dense nodes with exact power-of-two tiles: tcgen05 (kind::tf32 x3, TMEM)
    # (K > 256 runs in chunks of 256 inside the kernel: every chunk accumulates in TMEM from zero
    # and the epilogue folds it into C with round-to-nearest adds -- the tensor core's own
    # accumulation truncates, which is why a single TMEM accumulation stops at K = 256)
    if (allow_dmma and allow_tc05 and dtype == "complex64" and M >= 128 and K >= 4 and N >= 12
            and K <= TC05_MAX_K and M * N * K >= 1 << 20):
        # (extents need not be powers of two: build_pair_desc cuts every class into EQUAL tiles by
        # divisors -- 108 x 54 x 12 for the bond-6 PEPS GEMMs -- and falls back to the mma.sync
        # policy when that leaves the tensor-core tile too empty)
        if N >= 48:
            return VAR_TC05_128x64
        if N >= 24:
            return VAR_TC05_128x32
        return VAR_TC05_128x16
    # tensor-core tiles: fp64 DMMA for float64/complex128, 3xTF32 for float32/complex64
    if allow_dmma and M * N * K >= 1 << 15 and M * N >= 1024:
        if dtype == "complex128" and allow_3m and N >= 64 and K >= 64:
            # 3M complex product (opt-in): 25 % fewer DMMAs but narrower tiles (accumulator
            # registers).  Measured on B200: 36.6-37.8 vs 33.6 TFLOP/s at N=128 K=64, but 14 vs
            # 24 at N=128 K=16 (per-tile epilogue dominates) and no gain on the whole Sycamore
            # slice (160.5 vs 159 ms), so the default stays the 4-DMMA product.
            return VAR_DMMA3M_128x32
        if N >= 96:
            return VAR_DMMA_64x128
        if N >= 48:
            return VAR_DMMA_128x64
        if N >= 24:
            return VAR_DMMA_256x32
        return VAR_DMMA_256x16
    return VAR_SIMT_64x64


def build_pair_desc(dims: PairDims, dtype, accumulate=False, sm_count=148,
                    variant=None, allow_dmma=True, c_dense_elems=0,
                    force_splitk=None) -> PairPlan:
    """Pack a classified node into descriptor words."""
    dtype = dtype_name(dtype)
    m = coalesce([[d[0], d[1], d[3]] for d in dims.m])          # ext, sA, sC
    n = coalesce([[d[0], d[2], d[3]] for d in dims.n])          # ext, sB, sC
    k = coalesce([[d[0], d[1], d[2]] for d in dims.k])          # ext, sA, sB
    b = coalesce([list(d) for d in dims.batch])                 # ext, sA, sB, sC
    B, M, N, K = dims.sizes()

    # the streamed (large) operand is "A": swap roles when B's kept space is larger
    swapped = N > M
    if swapped:
        m, n = n, m
        k = [[d[0], d[2], d[1]] for d in k]
        b = [[d[0], d[2], d[1], d[3]] for d in b]
        M, N = N, M

    if variant is None:
        variant = choose_variant(dtype, B, M, N, K, allow_dmma)
    if variant in (VAR_DMMA3M_128x32, VAR_DMMA3M_256x16) and dtype != "complex128":
        # the 3M identity is a complex128 kernel: other dtypes take the plain tensor-core tiles
        variant = VAR_DMMA_256x32 if variant == VAR_DMMA3M_128x32 else VAR_DMMA_256x16
    if variant == VAR_DMMASTREAM and not (dtype == "complex128" and N <= 32 and K <= 64 and B == 1 and M < 1 << 32):
        variant = VAR_DMMA_256x16
    if variant == VAR_ROWSTREAM:
        # (a ragged blocked m dim is caught after tiling, below)
        ok = N <= 8 and K <= 8 and B == 1 and M < 1 << 32
        if not ok:
            variant = VAR_ROW_256x4 if N <= 4 else VAR_ROW_128x8
    MT, NT, KT = VARIANT_TILES[variant]
    if variant == VAR_DOTSTREAM4 and DTYPE_SIZES[dtype] < 16:
        KT = 2048  # 8 k per thread for the narrower element types (csrc/dotstream.cuh dot4_u)

    if variant in TC05_VARIANTS:
        # tcgen05: every thread of the epilogue owns a whole row, so the rows of a tile need
        # not be neighbours in C -- pick them for the longest contiguous runs of A instead
        # (and B is re-packed by bprime_kernel anyway: only A's strides matter for k too)
        tm, gm, pm = split_tile(m, (1,), MT, order_col=1, exact=True)
        tk, gk, pk = split_tile(k, (1,), KT, order_col=1, exact=True, multiple=4)
        tn, gn, pn = split_tile(n, (1, 2), NT, order_col=2, exact=True)
    else:
        tm, gm, pm = split_tile(m, (1, 2), MT, order_col=2)
        tk, gk, pk = split_tile(k, (1, 2), KT, order_col=1)
        tn, gn, pn = split_tile(n, (1, 2), NT, order_col=2)
    gm, pm = _order_grid(gm, pm, 1)
    gn, pn = _order_grid(gn, pn, 1)
    gk, pk = _order_grid(gk, pk, 1)
    gb = list(b)
    for name, lst, cap in (("m", gm, MAX_G), ("n", gn, MAX_G), ("k", gk, MAX_G), ("batch", gb, MAX_GB)):
        if len(lst) > cap:
            raise NotImplementedError(
                f"{len(lst)} non-coalescable {name} dims exceed the descriptor capacity {cap}"
            )
    gm, tiles_m = _with_divs(gm)
    gn, tiles_n = _with_divs(gn)
    gk, steps_k = _with_divs(gk)
    gb, tiles_b = _with_divs(gb)

    MTa = math.prod(d[0] for d in tm)
    NTa = math.prod(d[0] for d in tn)
    KTa = math.prod(d[0] for d in tk)

    # local weights (dim 0 fastest; partial dim is last by construction)
    def weights(tile):
        w, acc = [], 1
        for d in tile:
            w.append(acc)
            acc *= d[0]
        return w

    wm, wn, wk = weights(tm), weights(tn), weights(tk)
    # operand load orders: ascending stride in that operand (coalesced gathers)
    lda = [[d[0], d[1], w, 0] for d, w in zip(tm, wm)] + [[d[0], d[1], 0, w] for d, w in zip(tk, wk)]
    ldb = [[d[0], d[2], w, 0] for d, w in zip(tk, wk)] + [[d[0], d[1], 0, w] for d, w in zip(tn, wn)]
    key = lambda r: (abs(r[1]) if r[1] else 1 << 62)  # noqa: E731
    lda.sort(key=key)
    ldb.sort(key=key)

    tiles = tiles_m * tiles_n * tiles_b
    if force_splitk is not None:
        splitk = max(1, min(int(force_splitk), steps_k))
    else:
        splitk = 1
        if tiles < sm_count and steps_k >= 4:
            splitk = min(steps_k, -(-2 * sm_count // tiles))
        if variant == VAR_KRED:
            splitk = min(steps_k, 4 * sm_count)
        if variant == VAR_DMMA_32x32 and tiles == 1:
            splitk = min(steps_k, 2 * sm_count)  # two resident CTAs per SM
    if splitk > 1:
        per = -(-steps_k // splitk)
        splitk = -(-steps_k // per)
    if tiles * splitk >= 1 << 31:
        raise NotImplementedError("node needs more than 2^31 tiles")

    W = np.zeros(DESC_WORDS, dtype=np.int64)
    W[W_MAGIC] = DESC_MAGIC
    W[W_DTYPE] = DTYPE_CODES[dtype]
    W[W_NTM], W[W_NTN], W[W_NTK] = len(tm), len(tn), len(tk)
    W[W_NGM], W[W_NGN], W[W_NGK], W[W_NGB] = len(gm), len(gn), len(gk), len(gb)
    W[W_MTA], W[W_NTA], W[W_KTA] = MTa, NTa, KTa
    W[W_TILES_M], W[W_TILES_N], W[W_TILES_B], W[W_STEPS_K] = tiles_m, tiles_n, tiles_b, steps_k
    W[W_SPLITK] = splitk
    for base, p in ((W_PGM, pm), (W_PGN, pn), (W_PGK, pk)):
        if p is None:
            W[base:base + 4] = (-1, 0, 0, 0)
        else:
            W[base:base + 4] = p
    W[W_NLDA], W[W_NLDB] = len(lda), len(ldb)
    # bit1: columns (2q, 2q+1) of every tile row are adjacent in C and 32-byte
    # aligned -> the kernels may use 256-bit stores (complex128 only)
    expected, dense_n = 1, True
    for d in tn:
        if d[2] != expected:
            dense_n = False
            break
        expected *= d[0]
    pair_ok = (
        dtype == "complex128" and dense_n and pn is None and NTa >= 2 and NTa % 2 == 0
        and all(d[2] % 2 == 0 for d in tm)
        and all(g[3] % 2 == 0 for g in gm) and all(g[3] % 2 == 0 for g in gn)
        and all(g[4] % 2 == 0 for g in gb)
    )
    # bit2: all tile-grid extents are powers of two (Sycamore: always) -> the
    # producers decode tile indices with shifts/masks instead of idiv
    is_p2 = lambda e: e > 0 and (e & (e - 1)) == 0  # noqa: E731
    grid_pow2 = all(is_p2(g[0]) for g in gm + gn + gb)
    m_pow2 = all(is_p2(d[0]) for d in tm) and all(is_p2(g[0]) for g in gm)
    if variant in TC05_VARIANTS:
        # the tcgen05 kernel takes tiles of ONE shape: its native 128 x NT x 16, or smaller with
        # the rest of the tensor-core tile as padding (KTa in steps of 4: whole UMMA k8 groups);
        # below 40 % occupancy the mma.sync policy is the better choice
        # (k padding is free -- the UMMAs of missing k8 groups are not issued -- so only rows and columns count)
        occupancy = (MTa * NTa) / float(MT * NT)
        exact = (MTa <= MT and NTa <= NT and KTa <= KT and KTa % 4 == 0 and dtype == "complex64"
                 and occupancy >= 0.4 and steps_k <= 1024
                 and (KTa >= 8 or steps_k == 1)  # many 4-wide k-steps: per-step overhead, mma.sync is better
                 and all(p is None or p[1] % p[2] == 0 for p in (pm, pn, pk)))
        if not exact:
            fb = choose_variant(dtype, B, M, N, K, allow_dmma, allow_tc05=False)
            return build_pair_desc(dims, dtype, accumulate=accumulate, sm_count=sm_count, variant=fb,
                                   allow_dmma=allow_dmma, c_dense_elems=c_dense_elems,
                                   force_splitk=force_splitk)
    if variant == VAR_DOTSTREAM and (not (M == 1 and N == 1 and B == 1) or len(gk) > 40 or steps_k >= 1 << 31
                                    or (pk is not None and pk[1] % pk[2] != 0)):
        return build_pair_desc(dims, dtype, accumulate=accumulate, sm_count=sm_count, variant=VAR_KRED,
                               allow_dmma=allow_dmma, c_dense_elems=c_dense_elems, force_splitk=force_splitk)
    if variant == VAR_DOTSTREAM4 and (not (M <= 4 and N <= 4 and B == 1) or len(gk) > 40 or steps_k >= 1 << 31
                                     or pm is not None or pn is not None
                                     or (pk is not None and pk[1] % pk[2] != 0)
                                     or not (accumulate or c_dense_elems == M * N)):
        return build_pair_desc(dims, dtype, accumulate=accumulate, sm_count=sm_count, variant=VAR_SIMT_64x64,
                               allow_dmma=allow_dmma, c_dense_elems=c_dense_elems, force_splitk=force_splitk)
    if variant == VAR_DMMASTREAM and (pn is not None or pk is not None or (pm is not None and pm[1] % pm[2] != 0)):
        return build_pair_desc(dims, dtype, accumulate=accumulate, sm_count=sm_count, variant=VAR_DMMA_256x16,
                               allow_dmma=allow_dmma, c_dense_elems=c_dense_elems, force_splitk=force_splitk)
    if variant == VAR_ROWSTREAM_K:
        # exact tiles, and the k offsets must decompose as chunk_base[k // 8] + in_chunk[k % 8]
        def _koff(e, col):
            o = 0
            for d in tk:
                o += (e % d[0]) * d[col]
                e //= d[0]
            return o
        ok8 = all(_koff(e, 1) == _koff(e - e % 8, 1) + _koff(e % 8, 1) for e in range(KTa))
        if (not ok8 or pn is not None or pk is not None or (pm is not None and pm[1] % pm[2] != 0)
                or DTYPE_SIZES[dtype] > 8 or not (N <= 8 and K <= 64 and B == 1 and M < 1 << 32)):
            return build_pair_desc(dims, dtype, accumulate=accumulate, sm_count=sm_count,
                                   variant=VAR_ROW_256x4 if N <= 4 else VAR_ROW_128x8,
                                   allow_dmma=allow_dmma, c_dense_elems=c_dense_elems,
                                   force_splitk=force_splitk)
    if variant == VAR_ROWSTREAM and pm is not None and pm[1] % pm[2] != 0:
        # ragged blocked m dim: fall back to the staged row policy
        return build_pair_desc(dims, dtype, accumulate=accumulate, sm_count=sm_count,
                               variant=VAR_ROW_256x4 if N <= 4 else VAR_ROW_128x8,
                               allow_dmma=allow_dmma, c_dense_elems=c_dense_elems,
                               force_splitk=force_splitk)
    # 8-byte element types: groups of 4 (bit4) / 2 (bit5) columns adjacent in C and
    # 32- / 16-byte aligned -> vector row stores in the streaming row kernel
    def _cols_ok(g):
        if variant in TC05_VARIANTS:
            # exact tiles: only the leading columns of a tile row have to be adjacent
            run = 1
            for d in tn:
                if d[2] != run:
                    break
                run *= d[0]
            dense = run % g == 0
        else:
            dense = dense_n and pn is None
        return (
            DTYPE_SIZES[dtype] == 8 and dense and NTa % g == 0
            and all(d[2] % g == 0 for d in tm)
            and all(x[3] % g == 0 for x in gm) and all(x[3] % g == 0 for x in gn)
            and all(x[4] % g == 0 for x in gb)
        )

    # tcgen05 variants: if the A tile is made of long contiguous runs (dense prefix of
    # the load order), the producers fetch whole runs with TMA bulk copies (bit6)
    run_a, bulk_a = 1, False
    if variant in TC05_VARIANTS:
        for r_ in lda:
            if r_[1] != run_a:
                break
            run_a *= r_[0]
        rest = [r_[1] for r_ in lda if r_[1] >= run_a] + [g[2] for g in gm] + [g[2] for g in gk] + [g[2] for g in gb]
        # (cp.async.bulk: 16-byte aligned source, size a multiple of 16 bytes; one run per
        # producer thread -> at most 128 runs of >= 128 bytes)
        bulk_a = (run_a >= 16 and run_a % 2 == 0 and (MTa * KTa) % run_a == 0
                  and all(x % 2 == 0 for x in rest))
        # chunk-stride padding of the A' images (x16 B): 16 consecutive elements of A's memory
        # order -- the lanes of a half warp in the scatter pass -- should hit 16 different
        # 8-byte bank pairs.  Element (r, kk) sits at (kk//2)*LBO + r*16 + (kk%2)*8 bytes.
        def _conflicts(pad):
            lbo, worst = MT * 16 + 16 * pad, 0
            for half in range(2):
                slots = {}
                for e in range(16 * half, 16 * half + 16):
                    r_ = kk_ = 0
                    x = e
                    for ext, _s, wr, wk_ in lda:
                        r_ += (x % ext) * wr
                        kk_ += (x % ext) * wk_
                        x //= ext
                    slot = (((kk_ // 2) * lbo + r_ * 16 + (kk_ % 2) * 8) >> 3) & 15
                    slots[slot] = slots.get(slot, 0) + 1
                worst += max(slots.values())
            return worst
        W[35] = min((0, 1, 2, 4), key=_conflicts)  # W_LBOPAD
    W[34] = run_a  # W_RUNA
    W[W_FLAGS] = ((1 if accumulate else 0) | (2 if pair_ok else 0) | (4 if grid_pow2 else 0)
                  | (8 if m_pow2 else 0) | (16 if _cols_ok(4) else 0) | (32 if _cols_ok(2) else 0)
                  | (64 if bulk_a else 0))
    W[W_VARIANT] = variant
    W[W_CELEMS] = int(c_dense_elems)

    def put(off, rows, width):
        for i, r in enumerate(rows):
            W[off + i * width: off + (i + 1) * width] = r

    put(OFF_TM, tm, 3)
    put(OFF_TN, tn, 3)
    put(OFF_TK, tk, 3)
    put(OFF_GM, gm, 4)
    put(OFF_GN, gn, 4)
    put(OFF_GK, gk, 4)
    put(OFF_GB, gb, 5)
    put(OFF_LDA, lda, 4)
    put(OFF_LDB, ldb, 4)
    return PairPlan(W, variant, (B, M, N, K), swapped, tiles, splitk)


# ---------------------------------------------------------------------------
# single-operand nodes  (contract.py:61-119, 332-361)
# ---------------------------------------------------------------------------


def classify_single(term, shape, out, out_strides=None, strides_x=None):
    """``out[o] = sum_s X[...]``: output dims ``[ext, sX, sOut]`` and summed
    dims ``[ext, sX]``; repeated labels (diagonals/traces) add their strides."""
    term, out = tuple(term), tuple(out)
    shape = tuple(map(int, shape))
    if len(term) != len(shape):
        raise ValueError(f"Term '{term}' does not match shape {shape}.")
    sx = row_major_strides(shape) if strides_x is None else list(strides_x)
    ext, st_x, order = {}, {}, []
    for ix, d, s in zip(term, shape, sx):
        if ix not in order:
            order.append(ix)
        if ext.setdefault(ix, d) != d:
            raise ValueError(f"Index {ix} has mismatched sizes {ext[ix]} and {d}.")
        st_x[ix] = st_x.get(ix, 0) + s
    for ix in out:
        if ix not in ext:
            raise ValueError(f"Output index {ix} does not appear in the input.")
    out_shape = tuple(ext[ix] for ix in out)
    so = row_major_strides(out_shape) if out_strides is None else list(out_strides)
    odims = [[ext[ix], st_x[ix], s] for ix, s in zip(out, so) if ext[ix] != 1]
    sdims = [[ext[ix], st_x[ix]] for ix in order if ix not in out and ext[ix] != 1]
    return odims, sdims, out_shape


def build_single_desc(odims, sdims, dtype, accumulate=False) -> np.ndarray:
    dtype = dtype_name(dtype)
    odims = coalesce(odims)
    sdims = coalesce(sdims)
    # fastest output dim = smallest output stride (coalesced stores)
    odims.sort(key=lambda d: abs(d[2]) if d[2] else 1 << 62)
    sdims.sort(key=lambda d: abs(d[1]) if d[1] else 1 << 62)
    if len(odims) > MAX_S or len(sdims) > MAX_S:
        raise NotImplementedError("too many dims for a single-operand node")
    W = np.zeros(SDESC_WORDS, dtype=np.int64)
    W[S_MAGIC] = SDESC_MAGIC
    W[S_DTYPE] = DTYPE_CODES[dtype]
    W[S_NO], W[S_NS] = len(odims), len(sdims)
    W[S_OUT_ELEMS] = math.prod(d[0] for d in odims)
    W[S_SUM_ELEMS] = math.prod(d[0] for d in sdims)
    W[S_FLAGS] = 1 if accumulate else 0
    for i, d in enumerate(odims):
        W[OFF_SO + 3 * i: OFF_SO + 3 * i + 3] = d
    for i, d in enumerate(sdims):
        W[OFF_SS + 2 * i: OFF_SS + 2 * i + 2] = d
    return W


def split_equation(eq):
    """``(lhs_terms, out)`` of an explicit or implicit einsum equation
    (contract.py:34-58)."""
    eq = eq.replace(" ", "")
    if "..." in eq:
        raise NotImplementedError("Ellipsis not supported.")
    if "->" in eq:
        lhs, out = eq.split("->")
    else:
        lhs = eq
        flat = lhs.replace(",", "")
        out = "".join(c for c in sorted(set(flat)) if flat.count(c) == 1)
    return lhs.split(","), out
