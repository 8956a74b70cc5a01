"""The N>1 host logic on CPU: world_size-2 gloo processes share the slices
round-robin (``rank_slices`` == core.py:4070), each sums its share, and
``reduce_partials`` combines them with one collective (core.py:4078-4090),
including exponent-stripped pairs.  The per-rank arithmetic is done by the numpy
oracle here (no GPU in this suite); on the GPU box the same two functions sit
behind ``contract_distributed`` over NCCL."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import decode_ir, decode_sliced, load_json, load_npz, make_arrays, rel_err


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, name, strip, root, q):
    import cotengra_b200 as cb
    from oracle import ctg_oracle as orc

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rec = next(r for r in load_json("trees.json") if r["name"] == name)
        inputs = [tuple(t) for t in rec["inputs"]]
        sliced = decode_sliced(rec["sliced"])
        ir = decode_ir(rec["contractions"])
        arrays = make_arrays([tuple(rec["size_dict"][ix] for ix in t) for t in inputs],
                             rec["dtype"], seed=rec["seed"])
        begin, step, count = cb.rank_slices(rank, world, rec["nslices"])
        ids = range(begin, begin + step * count, step)
        part = orc.contract_tree(inputs, rec["output"], sliced, ir, arrays,
                                 strip_exponent=strip, slice_ids=ids)
        if strip:
            m, e = part
            res = cb.reduce_partials(torch.as_tensor(np.asarray(m)).reshape(np.shape(m)),
                                     torch.tensor([float(e)], dtype=torch.float64), root=root)
            out = None if res[0] is None else (res[0].numpy() * 10.0 ** float(res[1].item()))
        else:
            res = cb.reduce_partials(torch.as_tensor(np.asarray(part)), None, root=root)
            out = None if res is None else res.numpy()
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("strip", [False, True])
@pytest.mark.parametrize("root", [None, 0])
def test_two_ranks_gloo(strip, root):
    name = "lattice6x6_d3_sliced"
    want = load_npz("trees_values.npz")[name]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, strip, root, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert rel_err(got[0], want) < 1e-11
    if root is None:
        assert rel_err(got[1], want) < 1e-11
    else:
        assert got[1] is None


def test_rank_slices_and_guards():
    import cotengra_b200 as cb

    # every slice exactly once, round robin
    for world in (1, 2, 3, 8):
        seen = []
        for r in range(world):
            b, s, c = cb.rank_slices(r, world, 37)
            seen += list(range(b, b + s * c, s))
        assert sorted(seen) == list(range(37))
    with pytest.raises(ValueError):
        cb.rank_slices(0, 8, 4)  # fewer slices than ranks (core.py:4062-4066)
