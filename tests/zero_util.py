"""Test helper: inputs that make some (or all) slices of a sliced tree exactly zero
(``check_zero``, cotengra/contract.py:819-820)."""

import numpy as np


def zero_one_digit(spec, arrays, which=0, digit=1):
    """Zero every element of one input whose ``which``-th sliced index equals ``digit``:
    all slices with that digit contract to an all-zero intermediate."""
    ind = spec.sliced[which][0]
    out = [np.array(a, copy=True) for a in arrays]
    for c, term in enumerate(spec.inputs):
        if ind in term:
            sel = tuple(digit if ix == ind else slice(None) for ix in term)
            out[c][sel] = 0
            return out, ind
    raise ValueError("sliced index not on any input")
