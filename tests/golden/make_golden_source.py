"""Generates tests/golden/source/*.npz: SURVEY 8f-1, particles injected by UniformSource / ColdBeamSource and then moved.

Like make_golden.py, the vectors come from the INDEPENDENT pure-Python restatement of the Java source (tests/pyref.py:
java.util.Random, Spline.randomT, Source.sampleKinetic, KineticMaterial.addParticle, the mover) -- not from the C oracle and
not from the CUDA path -- and pin both (tests/test_golden.py).

    python tests/golden/make_golden_source.py     # rewrites the fixtures (committed; regenerate only on purpose)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import pyref  # noqa: E402
from make_golden import py_mesh  # noqa: E402
from starfish_b200 import synthetic as S  # noqa: E402
from starfish_b200.domain import DomainType  # noqa: E402

CASES = {
    # name: (cold_beam, v_drift, seed, num_mp per step, steps)
    "uniform_xy": (0, 7000.0, 20260117, 150, 4),
    "cold_beam_xy_negative_drift": (1, -6500.0, 42, 131, 4),
}
PTS = [(-0.1, 0.12), (-0.1, 0.03), (-0.09, -0.01), (-0.09, -0.02)]  # part of the inlet lies outside the mesh


def build(name):
    cold, v_drift, seed, num_mp, steps = CASES[name]
    m = S.make_mesh(21, 11, DomainType.XY, 5e-3, "open", x0=(-0.1, 0.0))
    wl = S.Workload(name, m, 1e-7, S.QE, 16 * S.AMU, 5, kick_frac=0.05)
    pts = PTS if v_drift > 0 else PTS[::-1]
    km = pyref.KM(wl.charge, wl.mass, [py_mesh(m)])
    rnd = pyref.JavaRandom(seed)
    added = []
    for it in range(steps):
        added.append(pyref.uniform_source_sample(km, pyref.Spline(pts), v_drift, num_mp, wl.dt, rnd, 1e3, born_it=it, cold_beam=bool(cold)))
        km.updateFields(wl.dt)
    parts = sorted(km.particles[0], key=lambda p: p.id)
    out = dict(
        meta=np.array([cold, seed, num_mp, steps], dtype=np.int64), v_drift=v_drift, dt=wl.dt, charge=wl.charge, mass=wl.mass,
        pts=np.array(pts), efi=m.efi, efj=m.efj, added=np.array(added, dtype=np.int64), rng_state=np.array(rnd.state, dtype=np.uint64),
        id=np.array([p.id for p in parts], dtype=np.int32), born_it=np.array([p.born_it for p in parts], dtype=np.int32),
        x=np.array([p.pos[0] for p in parts]), y=np.array([p.pos[1] for p in parts]), z=np.array([p.pos[2] for p in parts]),
        u=np.array([p.vel[0] for p in parts]), v=np.array([p.vel[1] for p in parts]), w=np.array([p.vel[2] for p in parts]),
        raw=np.array(km.raw[0]), n_exited=np.array(km.n_exited))
    np.savez_compressed(os.path.join(HERE, "source", name + ".npz"), **out)
    print(name, "np", len(parts), "added", added)


if __name__ == "__main__":
    for n in CASES:
        build(n)
