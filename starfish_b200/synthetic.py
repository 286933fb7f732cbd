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


class TutorialStep2:
    """BASELINE config 1: dat/examples/tutorial/step2 -- O+ ions streaming past a cylinder at -100 V.

    domain.xml: XY, uniform mesh 81x41 nodes, origin (-0.15, 0), spacing 5e-3; LEFT dirichlet, BOTTOM symmetry, others OPEN.
    boundaries.xml: 20-segment "cylinder" (solid; no surface interaction is loaded for O+ on SS, so Material.java:292-295 removes
    every ion that hits it) and the virtual "inlet" line x = -0.15, y: 0.20 -> 0.  materials.xml: O+ molwt 16, charge +1, spwt 1e3.
    starfish.xml: uniform source on the inlet, mdot 3.72e-13 kg/s, v_drift 7000 m/s; dt 1e-7, 400 iterations.
    The potential solver is outside this path: E is frozen to an analytic sheath, phi = -100 V * exp(-(r - r0) / 1 cm) around
    the cylinder (the example's plasma, n0 = 1e10 m^-3 and Te = 1.5 eV, shields the wall over about a centimetre), sampled on the nodes
    and zero inside the cylinder."""

    CYLINDER = ("0.05, 0 0.0475528, -0.0154508 0.0404508, -0.0293893 0.0293893, -0.0404508 0.0154508, -0.0475528 -9.18E-18, -0.05 "
                "-0.0154508, -0.0475528 -0.0293893, -0.0404508 -0.0404508, -0.0293893 -0.0475528, -0.0154508 -0.05, 6.12E-18 "
                "-0.0475528, 0.0154508 -0.0404508, 0.0293893 -0.0293893, 0.0404508 -0.0154508, 0.0475528 3.06E-18, 0.05 "
                "0.0154508, 0.0475528 0.0293893, 0.0404508 0.0404508, 0.0293893 0.0475528, 0.0154508 0.05, 0")

    def __init__(self, spwt=1e3, wall_kind=0, volts=-100.0):
        from .domain import LinearSpline, SolidBoundary, set_boundaries
        m = UniformMesh(81, 41, (-0.15, 0.0), (5e-3, 5e-3), DomainType.XY, name="mesh1")
        m.setMeshBCType(Face.LEFT, DomainBoundaryType.DIRICHLET)
        m.setMeshBCType(Face.BOTTOM, DomainBoundaryType.SYMMETRY)
        pts = np.array([float(t) for t in self.CYLINDER.replace(",", " ").split()]).reshape(-1, 2)
        self.cylinder = SolidBoundary("cylinder", pts, kind=wall_kind)
        set_boundaries(m, [self.cylinder])
        self.inlet = LinearSpline(np.array([[-0.15, 0.20], [-0.15, 0.0]]))
        x = m.x0[0] + np.arange(m.ni) * m.dh[0]
        y = m.x0[1] + np.arange(m.nj) * m.dh[1]
        X, Y = np.meshgrid(x, y, indexing="ij")
        r = np.sqrt(X * X + Y * Y)
        r0, lam = 0.05, 0.01
        er = np.where(r >= r0 * 0.999, (volts / lam) * np.exp(-(r - r0) / lam), 0.0)  # -d/dr [V0 exp(-(r - r0)/lam)]
        m.efi = np.ascontiguousarray(er * X / np.maximum(r, 1e-9))
        m.efj = np.ascontiguousarray(er * Y / np.maximum(r, 1e-9))
        self.mesh, self.name = m, "tutorial_step2"
        self.charge, self.mass, self.spwt = QE, 16 * AMU, float(spwt)
        self.dt, self.steps = 1e-7, 400
        self.mdot, self.v_drift = 3.72e-13, 7000.0
        self.mp_rem = 0.0

    def num_mp(self):
        """Source.regenerate(), Source.java:121-133: macroparticles of this step, the remainder carried over."""
        mp = (self.mdot * self.dt) / (self.mass * self.spwt) + self.mp_rem
        n = int(mp)
        self.mp_rem = mp - n
        return n
