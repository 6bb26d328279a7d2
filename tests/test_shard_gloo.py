"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: pixel sharding with slab
concatenation and transmit partition with an all-reduce. The per-rank DAS is computed by the oracle (the checker)
because the product path has no CPU implementation; what is under test is qups_b200/shard.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import small_problem, oracle_kwargs


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_c
    from qups_b200 import shard
    P = small_problem("FC", nz=12, nx=70, N=6, M=6, T=160)
    kw = oracle_kwargs(P["opts"])
    full = oracle_c.das_spec("DAS", P["Pi"], P["Pr"], P["Pv"], P["Nv"], P["x"], 0.0, P["fs"], P["c"], interp="cubic", **kw)[..., 0, 0, 0]
    if mode == "pixels":
        Pi_s, axis, s0, cnt = shard.pixel_shard(P["Pi"], rank, world, align=32)
        counts = [shard.pixel_shard(P["Pi"], r, world, align=32)[3] for r in range(world)]
        assert sum(counts) == P["Pi"].shape[axis] and axis == 2
        loc = oracle_c.das_spec("DAS", Pi_s, P["Pr"], P["Pv"], P["Nv"], P["x"], 0.0, P["fs"], P["c"], interp="cubic", **kw)[..., 0, 0, 0]
        b = torch.from_numpy(np.ascontiguousarray(loc))
        out = shard.gather_slabs(b, 1, counts)
        ok = np.array_equal(out.numpy(), full)
    else:
        m0, mc = shard.tx_shard(6, rank, world)
        sl = slice(m0, m0 + mc)
        loc = oracle_c.das_spec("DAS", P["Pi"], P["Pr"], P["Pv"][:, sl], P["Nv"][:, sl], np.asfortranarray(P["x"][:, :, sl]),
                                0.0, P["fs"], P["c"], interp="cubic", **kw)[..., 0, 0, 0]
        b = torch.from_numpy(np.ascontiguousarray(loc))
        shard.allreduce_image(b)
        ok = np.max(np.abs(b.numpy() - full)) <= 1e-5 * np.max(np.abs(full))  # summation order changes: tolerance
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["pixels", "transmits"])
def test_two_rank_sharding(mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_slab_partition_properties():
    from qups_b200 import shard
    for n in (1, 7, 128, 1000, 1024):
        for world in (1, 2, 3, 4, 8):
            parts = [shard.slab(n, r, world, 32) for r in range(world)]
            assert sum(c for _, c in parts) == n
            pos = 0
            for s, c in parts:
                assert s == pos and c >= 0
                pos += c
