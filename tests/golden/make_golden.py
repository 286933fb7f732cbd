"""Generates tests/golden/*.npz: inputs and expected outputs of the kinetic hot path on small seeded cases.

The reference is Java and no JVM exists in the build image, so these vectors cannot come from the
reference itself.  They are produced by the INDEPENDENT pure-Python restatement of the Java source (tests/pyref.py,
IEEE doubles, no FMA) -- not by the C oracle and not by the CUDA path -- and pin both of them:
tests/test_golden.py checks the C oracle against them on CPU and the CUDA path on the GPU.

    python tests/golden/make_golden.py        # rewrites the fixtures (committed; regenerate only on purpose)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import pyref  # noqa: E402
from starfish_b200 import synthetic as S  # noqa: E402
from starfish_b200.domain import DomainType  # noqa: E402

CASES = {
    # name: (domain, bc, ni, nj, n, steps, vth_cells, kick)
    "xy_periodic": (DomainType.XY, "periodic", 20, 17, 400, 4, 0.8, 0.1),
    "xy_open": (DomainType.XY, "open", 18, 21, 400, 4, 0.9, 0.1),
    "xy_symmetry": (DomainType.XY, "symmetry", 12, 12, 300, 5, 1.7, 0.2),
    "rz_beam": (DomainType.RZ, "beam", 16, 24, 400, 4, 0.6, 0.1),
    "zr_open": (DomainType.ZR, "open", 21, 15, 300, 4, 0.6, 0.1),
}


def py_mesh(m):
    pm = pyref.Mesh(m.ni, m.nj, m.x0, m.dh, int(m.domain_type))
    for f in range(4):
        pm.bc[f] = [int(v) for v in m.bc[f]]
    pm.Efi = pyref.Field(pm, m.efi)
    pm.Efj = pyref.Field(pm, m.efj)
    return pm


def build(name):
    dom, bc, ni, nj, n, steps, vth, kick = CASES[name]
    m = S.make_mesh(ni, nj, dom, 1e-3, bc)
    wl = S.Workload(name, m, 1e-7, S.QE, 16 * S.AMU, 1234, vth_cells=vth, kick_frac=kick)
    arr = wl.particles(0, n)
    km = pyref.KM(wl.charge, wl.mass, [py_mesh(m)])
    for q in range(n):
        km.addParticle(0, pyref.Particle([arr["x"][q], arr["y"][q], arr["z"][q]], [arr["u"][q], arr["v"][q], arr["w"][q]], arr["mpw"][q]), wl.dt)
    for _ in range(steps):
        km.updateFields(wl.dt)
    parts = sorted(km.particles[0], key=lambda p: p.id)
    out = dict(
        meta=np.array([int(dom), ni, nj, n, steps], dtype=np.int64), bc=np.array(bc), dt=wl.dt, charge=wl.charge, mass=wl.mass,
        efi=m.efi, efj=m.efj,
        **{"in_" + k: v for k, v in arr.items()},
        id=np.array([p.id for p in parts], dtype=np.int32),
        x=np.array([p.pos[0] for p in parts]), y=np.array([p.pos[1] for p in parts]), z=np.array([p.pos[2] for p in parts]),
        u=np.array([p.vel[0] for p in parts]), v=np.array([p.vel[1] for p in parts]), w=np.array([p.vel[2] for p in parts]),
        li=np.array([p.lc[0] for p in parts]), lj=np.array([p.lc[1] for p in parts]), dtp=np.array([p.dt for p in parts]),
        raw=np.array(km.raw[0]), sums=np.array(km.sums), n_exited=km.n_exited)
    return out


if __name__ == "__main__":
    for name in CASES:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **build(name))
        print("wrote", name)
