"""Time single pairwise contractions given as einsum equations (dev tool).
usage: gpu_pair_shapes.py [dtype] [--only=case] -- prints variant, tile, ms and TFLOP/s (8 flops per complex MAC) per case."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from cotengra_b200 import _lib, lowering as L

dtype = sys.argv[1] if len(sys.argv) > 1 else "complex64"
tdt = getattr(torch, dtype)
lib = _lib.load()
sm = _lib.device_info()["sm_count"]
S6 = (6,) * 6
CASES = [
    ("peps_top", "abcdefgh,bdfgxy->acehxy", (6, 6, 6, 6, 6, 6, 6, 216), (6, 6, 6, 6, 36, 36)),
    ("mat6_k1296", "mk,kn->mn", (46656, 1296), (1296, 1296)),
    ("mat6_k216", "mk,kn->mn", (46656, 216), (216, 1296)),
    ("mat6_k192", "mk,kn->mn", (46656, 192), (192, 1296)),
    ("mat6_k36", "mk,kn->mn", (46656, 36), (36, 1296)),
    ("mat6_kfast_k1296", "km,kn->mn", (1296, 46656), (1296, 1296)),
    ("pow2_k1024", "km,kn->mn", (1024, 32768), (1024, 1024)),
    ("pow2_k256", "km,kn->mn", (256, 32768), (256, 1024)),
    ("pow2_k64", "km,kn->mn", (64, 32768), (64, 1024)),
    ("pow2_mk_k1024", "mk,kn->mn", (32768, 1024), (1024, 1024)),
]
only = next((a.split("=")[1] for a in sys.argv if a.startswith("--only=")), None)
for name, eq, sa, sb in CASES:
    if only and name != only:
        continue
    t, o = L.split_equation(eq)
    dims = L.classify_pair(t[0], sa, t[1], sb, o)
    n_c = int(np.prod(dims.out_shape))
    pl = L.build_pair_desc(dims, dtype, sm_count=sm, c_dense_elems=n_c)
    a = torch.randn(int(np.prod(sa)), dtype=tdt, device="cuda")
    b = torch.randn(int(np.prod(sb)), dtype=tdt, device="cuda")
    c = torch.empty(n_c, dtype=tdt, device="cuda")
    pa, pb = (b, a) if pl.swapped else (a, b)

    def run():
        _lib.check(lib.ctgb_contract_pair(pl.words.ctypes.data, pa.data_ptr(), pb.data_ptr(), c.data_ptr(), 0))

    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    B, M, N, K = pl.sizes
    W = pl.words
    print(f"{name}: M={M} N={N} K={K} var={int(W[L.W_VARIANT])} tile=({int(W[L.W_MTA])},{int(W[L.W_NTA])},{int(W[L.W_KTA])}) "
          f"steps_k={int(W[L.W_STEPS_K])} splitk={int(W[L.W_SPLITK])} swapped={pl.swapped} {ms:.3f} ms "
          f"{8 * B * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
