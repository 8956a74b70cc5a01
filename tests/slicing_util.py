"""Test helper: slice a TreeSpec further (host-side integer work only) so that
its largest intermediate has at most ``width`` elements.  A simple greedy
stand-in for cotengra's SliceFinder (which stays in cotengra): repeatedly slice
the index that appears on the largest intermediate and on most large ones."""

from cotengra_b200 import TreeSpec


def _node_sizes(spec):
    spec._ir = None
    spec.contractions()
    sizes = {}
    for k, inds in spec.inds.items():
        n = 1
        for ix in inds:
            n *= spec.size_dict[ix]
        sizes[k] = (n, inds)
    return sizes


def slice_to_width(spec, width, prefer=None):
    sliced = list(prefer.sliced) if prefer is not None else list(spec.sliced)
    cur = TreeSpec(spec.inputs, spec.output, spec.size_dict, spec.path, sliced)
    while True:
        sizes = _node_sizes(cur)
        big, inds = max(sizes.values(), key=lambda t: t[0])
        if big <= width:
            return cur
        score = {}
        for n, ii in sizes.values():
            if n * 4 >= big:
                for ix in ii:
                    if ix not in cur.output:
                        score[ix] = score.get(ix, 0) + n
        cand = [ix for ix in inds if ix in score]
        best = max(cand, key=lambda ix: (score[ix], ix))
        inner = [s for s in sliced]
        inner.append((best, cur.size_dict[best], None))
        # reference order: output (outer) sliced indices first, then by name
        inner.sort(key=lambda s: (s[0] not in cur.output, s[0]))
        sliced = inner
        cur = TreeSpec(spec.inputs, spec.output, spec.size_dict, spec.path, sliced)


def slice_one_more(spec):
    """Slice exactly one more (inner) index: the one most present on the
    largest intermediates."""
    sizes = _node_sizes(spec)
    big, inds = max(sizes.values(), key=lambda t: t[0])
    best = max((ix for ix in inds if ix not in spec.output), key=lambda ix: ix)
    sliced = list(spec.sliced) + [(best, spec.size_dict[best], None)]
    sliced.sort(key=lambda s: (s[0] not in spec.output, s[0]))
    return TreeSpec(spec.inputs, spec.output, spec.size_dict, spec.path, sliced), best


def slice_id(spec, key):
    """Inverse of ``slice_key``: digits -> slice id (core.py:3775-3800)."""
    i = 0
    for (ind, _size, project), stride in zip(spec.sliced, spec.slice_strides()):
        if project is None:
            i += key[ind] * stride
    return i


def appxB_at_width(width_log2):
    """The Sycamore-m20 Appendix-B tree (W = 2^30 as benchmarked), sliced further to
    W = 2^width_log2 for smaller widths -- the trees of ``tests/golden/big_slices.json``."""
    from tests.helpers import decode_sliced, load_json

    rec = next(r for r in load_json("sycamore_m20.json") if r["name"] == "sycamore_m20_appxB")
    spec = TreeSpec(rec["inputs"], rec["output"], rec["size_dict"], rec["path"], decode_sliced(rec["sliced"]))
    return spec if width_log2 >= 30 else slice_to_width(spec, 2 ** width_log2)
