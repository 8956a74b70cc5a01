"""Numpy-only stand-in for the third-party ``autoray`` dispatch layer.

TEST INFRASTRUCTURE ONLY.  The reference (jcmgray/cotengra, pure Python) has one
hard dependency, ``autoray`` (pinned 0.8.10 in its pixi.lock), which is not
installed in the build container and cannot be fetched (no network).  This
module implements the eight names the reference imports
by forwarding straight to numpy, so that the *unmodified* reference under
``/root/reference`` can be imported HERE to (a) validate ``oracle/ctg_oracle.py``
and (b) generate the golden vectors committed under ``tests/golden/``.

It contains no arithmetic of its own and is never imported by the product
(``cotengra_b200``), by ``bench.py`` or by the ``-m gpu`` tests: it is only put
on ``sys.path`` by ``oracle/gen_golden.py`` and ``oracle/refenv.py``.
"""

import contextlib

import numpy as _np


def infer_backend(x):
    return "numpy"


def infer_backend_multi(*xs):
    return "numpy"


def get_namespace(backend=None):
    return _np


def shape(x):
    try:
        return tuple(int(d) for d in x.shape)
    except AttributeError:
        return tuple(int(d) for d in _np.shape(x))


def to_numpy(x):
    return _np.asarray(x)


def do(fn, *args, like=None, **kwargs):
    return getattr(_np, fn)(*args, **kwargs)


@contextlib.contextmanager
def backend_like(backend):
    yield


def autojit(fn=None, **kwargs):
    if fn is None:
        return lambda f: f
    return fn


class _Lazy:
    """Placeholder: constant folding through ``autoray.lazy`` is outside the
    hot path (interface.py:539-560) and is not needed for golden generation."""

    def __getattr__(self, name):
        raise ImportError("autoray.lazy is not provided by the numpy-only stand-in")


lazy = _Lazy()
