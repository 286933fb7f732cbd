"""Two / four / eight GPUs: particles partitioned by index, mesh replicated, NCCL allreduce of the deposit inside sfgpu_step.
Particle state must be bit-identical to the single-population oracle, the summed deposit within 1e-10."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from starfish_b200 import KineticMaterial, Particles, synthetic as S
    from starfish_b200.parallel import attach_communicator, shard_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wl = S.config_b(ni=96, nj=80, bc="open")
        n = 200001
        first, count = shard_bounds(n, rank, world)
        arr = wl.particles(first, count)
        arr["id"] = np.arange(first, first + count, dtype=np.int32)
        with KineticMaterial("O+", wl.charge, wl.mass, [wl.mesh], wl.mesh.domain_type, device=rank) as km:
            km.dt = wl.dt
            attach_communicator(km)
            km.addParticles(wl.mesh, Particles(count, **arr), wl.dt)
            for _ in range(5):
                km.updateFields()
            p = km.getParticles(wl.mesh).sorted_by_id()
            np.savez(os.path.join(out, f"rank{rank}.npz"), dep=km.last_deposit[0], nd=km.getDen(wl.mesh), id=p.id, x=p.x, y=p.y, u=p.u,
                     v=p.v, sums=np.array([km.mass_sum, *km.momentum_sum, km.energy_sum]), np_=km.getNp(), n_exited=km.n_exited)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_gpus_match_single_population_oracle(tmp_path, world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    from oracle import oracle as O
    from starfish_b200 import synthetic as S
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    wl = S.config_b(ni=96, nj=80, bc="open")
    n = 200001
    ok = O.OracleKM(wl.charge, wl.mass, [wl.mesh])
    ok.addParticles(0, wl.particles(0, n), wl.dt)
    for _ in range(5):
        ok.updateFields(wl.dt)
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    for k in range(1, world):
        assert np.array_equal(r[0]["dep"], r[k]["dep"])  # allreduce leaves the same sum on every rank
    scale = np.abs(ok.raw[0]).max(axis=(1, 2), keepdims=True)
    assert np.all(np.abs(r[0]["dep"] - ok.raw[0]) <= 1e-10 * scale)
    assert np.array_equal(r[0]["dep"][7], ok.raw[0][7])
    assert np.allclose(r[0]["nd"], ok.fields[0]["nd"], rtol=1e-10, atol=1e-10 * np.abs(ok.fields[0]["nd"]).max())
    want = np.array([ok.mass_sum, *ok.momentum_sum, ok.energy_sum])
    assert np.allclose(r[0]["sums"], want, rtol=1e-10, atol=1e-10 * abs(want[4]))
    assert sum(int(x["np_"]) for x in r) == ok.getNp()
    assert sum(int(x["n_exited"]) for x in r) == ok.n_exited
    ids = np.concatenate([x["id"] for x in r])
    o = np.argsort(ids)
    full = ok.sorted_parts(0)
    assert np.array_equal(ids[o], full["id"])
    for key in ("x", "y", "u", "v"):
        assert np.array_equal(np.concatenate([x[key] for x in r])[o], full[key]), key


@pytest.mark.parametrize("n_gpus", [1, 2, 8])
def test_single_caller_group_matches_oracle(n_gpus):
    """sfgpu_multi_*: ONE host thread (Starfish's main loop, Starfish.java:77-121) drives all GPUs; the worker threads of the
    group run the per-GPU steps and the NCCL all-reduce concurrently.  Same bars as the one-process-per-GPU layout."""
    import torch
    if torch.cuda.device_count() < n_gpus:
        pytest.skip(f"needs {n_gpus} GPUs")
    from oracle import oracle as O
    from starfish_b200 import Particles, synthetic as S
    from starfish_b200.multi import MultiGpuKineticMaterial
    wl = S.config_b(ni=96, nj=80, bc="open")
    n = 150001
    arr = wl.particles(0, n)
    ok = O.OracleKM(wl.charge, wl.mass, [wl.mesh])
    ok.addParticles(0, arr, wl.dt)
    with MultiGpuKineticMaterial("O+", wl.charge, wl.mass, wl.mesh, n_gpus) as km:
        km.dt = wl.dt
        assert km.n_gpus == n_gpus
        assert km.addParticles(Particles(n, **arr)) == n
        for _ in range(4):
            km.updateFields()
            ok.updateFields(wl.dt)
            assert km.np_alive == ok.getNp() and km.n_exited == ok.n_exited
        dep = km.deposit()
        scale = np.abs(ok.raw[0]).max(axis=(1, 2), keepdims=True)
        assert np.all(np.abs(dep - ok.raw[0]) <= 1e-10 * scale) and np.array_equal(dep[7], ok.raw[0][7])
        assert np.allclose(km.density(), ok.fields[0]["nd"], rtol=1e-10, atol=1e-10 * np.abs(ok.fields[0]["nd"]).max())
        assert np.allclose(km.sums5, ok.sums5, rtol=1e-10, atol=1e-10 * abs(ok.sums5[4]))
        p = km.getParticles().sorted_by_id()
        o = ok.sorted_parts(0)
        assert np.array_equal(p.id, o["id"])  # ids from one counter: unique over the GPUs, the oracle's numbering
        for key in ("x", "y", "u", "v", "w", "li", "lj"):
            assert np.array_equal(getattr(p, key), o[key]), key
