"""Synthetic workloads of BASELINE.json, generated identically for the GPU path and the oracle.

Counter-based RNG: particle k draws from splitmix64(seed, k, stream) so that any shard of the population can
be generated independently (multi-GPU partitioning by particle index) and the CPU oracle sees the same
numbers.  Nothing here is on the measured path: inputs are built once, before the timed region.
"""
from __future__ import annotations

import numpy as np

from .domain import DomainBoundaryType, DomainType, Face, UniformMesh

AMU = 1.660538921e-27  # Constants.java
QE = 1.602176565e-19

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    """Vectorised splitmix64 finaliser over uint64 arrays."""
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def uniform(seed, first, n, stream):
    """n doubles in [0,1) for particle indices first..first+n-1, independent per `stream`."""
    k = np.arange(first, first + n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        key = splitmix64(np.uint64(seed) + np.uint64(stream) * np.uint64(0xD1B54A32D192ED03))
        z = splitmix64(k ^ key)
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def normal(seed, first, n, stream):
    """Box-Muller on two uniform streams."""
    u1 = uniform(seed, first, n, 2 * stream + 100)
    u2 = uniform(seed, first, n, 2 * stream + 101)
    return np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)


def make_mesh(ni, nj, domain_type=DomainType.XY, dh=1e-3, bc="periodic", x0=(0.0, 0.0)):
    """Uniform mesh with the boundary set of the synthetic configs.
    bc: 'periodic' (all faces), 'open' (all faces), 'beam' (RZ config C: LEFT symmetry, others open)."""
    m = UniformMesh(ni, nj, x0, (dh, dh), domain_type)
    if bc == "periodic":
        for f in Face:
            m.setMeshBCType(f, DomainBoundaryType.PERIODIC)
    elif bc == "open":
        pass
    elif bc == "symmetry":
        for f in Face:
            m.setMeshBCType(f, DomainBoundaryType.SYMMETRY)
    elif bc == "beam":
        m.setMeshBCType(Face.LEFT, DomainBoundaryType.SYMMETRY)
    else:
        raise ValueError(bc)
    return m


def set_analytic_field(mesh, e0):
    """efi = E0 sin(2 pi x / Lx), efj = E0 cos(2 pi y / Ly) on the nodes."""
    x = mesh.x0[0] + np.arange(mesh.ni) * mesh.dh[0]
    y = mesh.x0[1] + np.arange(mesh.nj) * mesh.dh[1]
    lx, ly = (mesh.ni - 1) * mesh.dh[0], (mesh.nj - 1) * mesh.dh[1]
    mesh.efi = np.ascontiguousarray(np.broadcast_to((e0 * np.sin(2 * np.pi * (x - mesh.x0[0]) / lx))[:, None], (mesh.ni, mesh.nj)))
    mesh.efj = np.ascontiguousarray(np.broadcast_to((e0 * np.cos(2 * np.pi * (y - mesh.x0[1]) / ly))[None, :], (mesh.ni, mesh.nj)))


class Workload:
    """A synthetic config: mesh, species constants, dt and a particle generator."""

    def __init__(self, name, mesh, dt, charge, mass, seed, vth_cells=0.2, drift_cells=(0.0, 0.0), wth_cells=None,
                 r_frac=1.0, mpw=1e3, kick_frac=0.01):
        self.name, self.mesh, self.dt, self.charge, self.mass, self.seed = name, mesh, dt, charge, mass, seed
        self.mpw = mpw
        dh = mesh.dh[0]
        self.vth = vth_cells * dh / dt
        self.wth = self.vth if wth_cells is None else wth_cells * dh / dt
        self.drift = (drift_cells[0] * dh / dt, drift_cells[1] * dh / dt)
        self.r_frac = r_frac
        # E0 such that the per-step kick is kick_frac * v_th
        qm = charge / mass
        self.e0 = kick_frac * max(self.vth, 1e-30) / (abs(qm) * dt)
        set_analytic_field(mesh, self.e0)

    def particles(self, first, n):
        """SoA arrays for particle indices [first, first+n): uniform positions, Maxwellian velocities."""
        m, s = self.mesh, self.seed
        lx, ly = (m.ni - 1) * m.dh[0], (m.nj - 1) * m.dh[1]
        x = m.x0[0] + uniform(s, first, n, 0) * lx * self.r_frac
        y = m.x0[1] + uniform(s, first, n, 1) * ly
        # keep r > 0 strictly in axisymmetric runs
        if m.domain_type == DomainType.RZ:
            x = np.maximum(x, m.x0[0] + 1e-9 * m.dh[0])
        elif m.domain_type == DomainType.ZR:
            y = np.maximum(y, m.x0[1] + 1e-9 * m.dh[1])
        return dict(x=x, y=y, z=np.zeros(n), u=self.drift[0] + self.vth * normal(s, first, n, 0),
                    v=self.drift[1] + self.vth * normal(s, first, n, 1), w=self.wth * normal(s, first, n, 2),
                    mpw=np.full(n, self.mpw))


def config_b(ni=512, nj=512, bc="periodic", beam=False, seed=20260117):
    """Synthetic B: XY 512x512 nodes, uniform load, O+ ions, v_th*dt/dh = 0.2 (beam: drift 0.5, thermal 0.05)."""
    mesh = make_mesh(ni, nj, DomainType.XY, 1e-3, bc)
    kw = dict(vth_cells=0.05, drift_cells=(0.5, 0.0)) if beam else dict(vth_cells=0.2)
    return Workload("xy%dx%d_%s%s" % (ni, nj, bc, "_beam" if beam else ""), mesh, 1e-7, QE, 16 * AMU, seed, **kw)


def config_c(ni=1024, nj=1024, seed=20260118):
    """Synthetic C: RZ 1024x1024, beam along +z (j) over r in [0, 0.25 Rmax], LEFT symmetry, other faces open."""
    mesh = make_mesh(ni, nj, DomainType.RZ, 1e-3, "beam")
    return Workload("rz%dx%d_beam" % (ni, nj), mesh, 1e-7, QE, 16 * AMU, seed, vth_cells=0.05, drift_cells=(0.0, 0.5),
                    wth_cells=0.05, r_frac=0.25)


def config_e(ni=2048, nj=2048, seed=20260119):
    """Synthetic E: XY 2048x2048, particles loaded as in B, partitioned by index over the ranks."""
    mesh = make_mesh(ni, nj, DomainType.XY, 1e-3, "periodic")
    return Workload("xy%dx%d_periodic" % (ni, nj), mesh, 1e-7, QE, 16 * AMU, seed, vth_cells=0.2)


def config_multi_domain():
    """BASELINE config 4: the mesh layout of dat/examples/multi-domain/domain.xml (RZ, four uniform meshes with different
    spacings, LEFT symmetry), Ar+ with spwt 1e3, dt 5e-8 (materials.xml, starfish.xml); no field solver in that example.
    Returns (meshes, charge, mass, dt)."""
    from .domain import set_mesh_neighbors
    spec = [((0.0, 0.0), (5e-4, 1e-3), (12, 50)), ((0.0055, 0.0), (1e-3, 1e-3), (12, 14)),
            ((0.0055, 0.031), (5e-4, 1e-3), (10, 9)), ((0.0, 0.049), (1e-3, 1e-3), (14, 14))]
    meshes = []
    for k, (x0, dh, (ni, nj)) in enumerate(spec):
        m = UniformMesh(ni, nj, x0, dh, DomainType.RZ, name="mesh%d" % (k + 1))
        m.setMeshBCType(Face.LEFT, DomainBoundaryType.SYMMETRY)
        meshes.append(m)
    set_mesh_neighbors(meshes)
    return meshes, QE, 39.9 * AMU, 5e-8
