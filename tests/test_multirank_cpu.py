"""N>1 host logic on CPU (gloo, world_size 2): index partition, unique-id exchange, and the invariant the multi-GPU
design rests on -- per-rank deposits of disjoint particle shards on a replicated mesh sum to the single-rank deposit,
while particle state does not depend on the partition (checked here on the oracle)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from starfish_b200.parallel import exchange_unique_id, shard_bounds


def test_shard_bounds_partition():
    for n in (0, 1, 7, 16, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _f, c in spans) == n
            for (f0, c0), (f1, _c1) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _f, c in spans) - min(c for _f, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch
    from oracle import oracle as O
    from starfish_b200 import synthetic as S
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = exchange_unique_id(lambda: bytes([rank + 1]) * 128)
        assert uid == bytes([1]) * 128  # rank 0's id everywhere
        wl = S.config_b(ni=40, nj=33, bc="open")
        n = 6001
        first, count = shard_bounds(n, rank, world)
        km = O.OracleKM(wl.charge, wl.mass, [wl.mesh])
        km.addParticles(0, wl.particles(first, count), wl.dt, ids=np.arange(first, first + count))
        for _ in range(3):
            km.updateFields(wl.dt)
        dep = torch.from_numpy(km.raw[0].copy())
        dist.all_reduce(dep)  # what ncclAllReduce does on the packed [8][ni][nj] buffer
        sums = torch.from_numpy(km.sums5.copy())
        dist.all_reduce(sums)
        counts = torch.tensor([km.getNp(), km.n_exited])
        dist.all_reduce(counts)
        p = km.sorted_parts(0)
        np.savez(os.path.join(out, f"rank{rank}.npz"), dep=dep.numpy(), sums=sums.numpy(), counts=counts.numpy(), id=p["id"], x=p["x"], u=p["u"])
    finally:
        dist.destroy_process_group()


def test_two_ranks_sum_to_single_rank(tmp_path):
    from oracle import oracle as O
    from starfish_b200 import synthetic as S
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    wl = S.config_b(ni=40, nj=33, bc="open")
    n = 6001
    ref = O.OracleKM(wl.charge, wl.mass, [wl.mesh])
    ref.addParticles(0, wl.particles(0, n), wl.dt)
    for _ in range(3):
        ref.updateFields(wl.dt)
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    assert np.array_equal(r[0]["dep"], r[1]["dep"])  # allreduce result is replicated
    scale = np.abs(ref.raw[0]).max(axis=(1, 2), keepdims=True)
    assert np.all(np.abs(r[0]["dep"] - ref.raw[0]) <= 1e-10 * scale)
    assert np.array_equal(r[0]["dep"][7], ref.raw[0][7])  # cell counts are integers: exact
    assert np.allclose(r[0]["sums"], ref.sums5, rtol=1e-10)
    assert list(r[0]["counts"]) == [ref.getNp(), ref.n_exited]
    # particle state is independent of the partition: bit exact
    ids = np.concatenate([r[0]["id"], r[1]["id"]])
    o = np.argsort(ids)
    full = ref.sorted_parts(0)
    assert np.array_equal(ids[o], full["id"])
    assert np.array_equal(np.concatenate([r[0]["x"], r[1]["x"]])[o], full["x"])
    assert np.array_equal(np.concatenate([r[0]["u"], r[1]["u"]])[o], full["u"])
