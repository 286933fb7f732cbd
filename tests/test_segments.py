"""Surface hits on linear segments (SURVEY 8 a10 second half, f-4; BASELINE config 1 = dat/examples/tutorial/step2):
LinearSegment.intersect (LinearSegment.java:113-179), the nearest-hit search with the start-of-step exclusion and the 0.9999
back-off (KM:519-603), ABSORB / keep outcomes.  CPU part: known answers, and the C oracle against the independent Python
restatement on the tutorial geometry.  GPU part (tests/test_gpu_segments.py) compares the CUDA path with the oracle."""
import numpy as np
import pytest

import pyref
from oracle import oracle as O
from starfish_b200 import synthetic as S
from starfish_b200.domain import DomainType, SolidBoundary, UniformMesh, set_boundaries


def py_mesh(m):
    pm = pyref.Mesh(m.ni, m.nj, m.x0, m.dh, int(m.domain_type))
    for f in range(4):
        pm.bc[f] = [int(v) for v in m.bc[f]]
    pm.Efi = pyref.Field(pm, m.efi)
    pm.Efj = pyref.Field(pm, m.efj)
    pm.has_seg = [[int(v) for v in row] for row in m.has_seg]
    sg = m.segments
    pm.segments = [pyref.WallSegment(q, float(sg["x1"][q]), float(sg["y1"][q]), float(sg["x2"][q]), float(sg["y2"][q]), int(sg["kind"][q]), bool(sg["sink"][q]))
                   for q in range(len(sg["x1"]))]
    pm.node_segments = m.node_segments
    return pm


def test_kat_segment_intersect():
    lib = O.load()
    import ctypes as C
    lib.sfo_segment_intersect.restype = None
    seg = O._Segment(0.0, 0.0, 1.0, 0.0, 0, 0)  # along +x
    t = (C.c_double * 2)()

    def hit(p3, p4):
        lib.sfo_segment_intersect(C.byref(seg), (C.c_double * 2)(*p3), (C.c_double * 2)(*p4), t)
        ws = pyref.WallSegment(0, 0.0, 0.0, 1.0, 0.0).intersect(list(p3), list(p4))
        assert [t[0], t[1]] == ws  # C oracle == Python restatement, bit for bit
        return t[0], t[1]
    assert hit((0.25, 1.0), (0.25, -1.0)) == (0.25, 0.5)          # crossing at the quarter point, half way along the path
    assert hit((0.25, 1.0), (0.25, 0.5)) == (-1.0, -1.0)          # stops short
    assert hit((2.0, 1.0), (2.0, -1.0)) == (-1.0, -1.0)           # beside the segment
    assert hit((0.0, 1.0), (1.0, 1.0)) == (-1.0, -1.0)            # parallel: den == 0
    t0, t1 = hit((0.5, 1.0), (0.5, 0.0))                           # ends ON the segment: t_part = 1
    assert (t0, t1) == (0.5, 1.0)
    t0, t1 = hit((1.0 + 5e-8, 1.0), (1.0 + 5e-8, -1.0))            # within FLT_EPS of the end point: clamped to 1
    assert t0 == 1.0 and t1 == 0.5


def test_tutorial_geometry_node_ownership():
    cfg = S.TutorialStep2()
    m = cfg.mesh
    assert len(m.segments["x1"]) == 20 and m.has_seg.sum() > 60
    # every node within one cell of the circle r = 0.05 owns a segment, nodes two cells away from the polygon do not
    x = m.x0[0] + np.arange(m.ni)[:, None] * m.dh[0]
    y = m.x0[1] + np.arange(m.nj)[None, :] * m.dh[1]
    r = np.hypot(x, y)
    assert m.has_seg[(np.abs(r - 0.05) < 0.6 * m.dh[0])].all()
    assert not m.has_seg[np.abs(r - 0.05) > 3.1 * m.dh[0]].any()


@pytest.mark.parametrize("wall_kind", [0, 1], ids=["absorb", "keep"])
def test_oracle_matches_python_restatement_on_tutorial_step2(wall_kind):
    """dat/examples/tutorial/step2 at a coarser specific weight: the inlet source (java.util.Random draws), the frozen sheath
    field, the absorbing cylinder.  Absorbing walls draw no random numbers, so the two restatements must agree bit for bit."""
    cfg = S.TutorialStep2(spwt=2e4, wall_kind=wall_kind)
    m = cfg.mesh
    ok = O.OracleKM(cfg.charge, cfg.mass, [m])
    km = pyref.KM(cfg.charge, cfg.mass, [py_mesh(m)])
    rnd = pyref.JavaRandom(0)
    spl = pyref.Spline(np.array([[-0.15, 0.20], [-0.15, 0.0]]))
    state = O.java_seed(0)
    absorbed = hits = 0
    for it in range(200):
        n_mp = cfg.num_mp()
        n, state = ok.sampleUniformSource(cfg.inlet, cfg.v_drift, n_mp, cfg.dt, state, cfg.spwt, born_it=it)
        assert n == pyref.uniform_source_sample(km, spl, cfg.v_drift, n_mp, cfg.dt, rnd, cfg.spwt, born_it=it)
        ok.updateFields(cfg.dt)
        km.updateFields(cfg.dt)
        assert ok.n_absorbed == km.n_absorbed and ok.n_exited == km.n_exited and not ok.slow and not km.slow
        h = ok.hits[0]
        assert len(h["seg"]) == len(km.hits)
        for q, (sid, ts, vel, mpw, alive) in enumerate(km.hits):  # same hits in the same order (single mover thread)
            assert (int(h["seg"][q]), float(h["t"][q]), float(h["u"][q]), float(h["v"][q]), float(h["mpw"][q]), bool(h["alive"][q])) == (sid, ts, vel[0], vel[1], mpw, alive)
        absorbed += ok.n_absorbed
        hits += len(km.hits)
    assert hits > 20 and (absorbed == hits if wall_kind == 0 else absorbed == 0)
    p = ok.sorted_parts(0)
    q = sorted(km.particles[0], key=lambda a: a.id)
    assert len(q) == len(p["x"]) > 1000
    for key, get in (("x", lambda a: a.pos[0]), ("y", lambda a: a.pos[1]), ("u", lambda a: a.vel[0]), ("v", lambda a: a.vel[1]), ("li", lambda a: a.lc[0]),
                     ("lj", lambda a: a.lc[1]), ("dt", lambda a: a.dt)):
        assert np.array_equal(p[key], np.array([get(a) for a in q])), key
    assert state == rnd.state


def random_walls_case(seed):
    """Mesh with a random polygon (absorbing or specular), a random open polyline whose hits leave the particle alive (unchanged or specular), a random SINK line;
    hot particles."""
    rng = np.random.default_rng(1000 + seed)
    dom = [DomainType.XY, DomainType.RZ, DomainType.ZR][seed % 3]
    bc = ["open", "periodic", "symmetry"][(seed // 3 + seed) % 3]
    ni, nj = int(rng.integers(24, 40)), int(rng.integers(20, 36))
    m = S.make_mesh(ni, nj, dom, 1e-3, bc)
    lx, ly = (ni - 1) * 1e-3, (nj - 1) * 1e-3
    walls = []
    # a closed polygon around a random centre (absorbing) ...
    cx, cy, r = lx * rng.uniform(0.35, 0.65), ly * rng.uniform(0.35, 0.65), min(lx, ly) * rng.uniform(0.08, 0.2)
    nv = int(rng.integers(3, 9))
    ang = np.sort(rng.uniform(0, 2 * np.pi, nv))
    poly = np.stack([cx + r * np.cos(ang), cy + r * np.sin(ang)], axis=1)
    walls.append(SolidBoundary("poly", np.vstack([poly, poly[:1]]), kind=2 if seed % 2 else 0))  # odd seeds: a specular body instead of an absorbing one
    # ... an open polyline whose hits leave the particle alive (it goes on with the rest of its step) ...
    pts = np.stack([lx * rng.uniform(0.05, 0.95, 4), ly * rng.uniform(0.05, 0.95, 4)], axis=1)
    walls.append(SolidBoundary("keep", pts, kind=2 if seed % 3 == 0 else 1))
    # ... and a SINK line
    pts = np.stack([lx * rng.uniform(0.05, 0.95, 2), ly * rng.uniform(0.05, 0.95, 2)], axis=1)
    walls.append(SolidBoundary("sink", pts, kind=1, sink=True))
    set_boundaries(m, walls)
    wl = S.Workload("t", m, 1e-7, S.QE, 16 * S.AMU, 77 + seed, vth_cells=1.7, kick_frac=0.2)
    arr = wl.particles(0, 3000)
    return m, wl, arr


@pytest.mark.parametrize("seed", range(6))
def test_oracle_matches_python_restatement_on_random_walls(seed):
    """Differential test beyond the tutorial geometry: random closed and open polylines of absorbing, surviving and SINK segments, hot particles
    that cross several cells (and several segments) per step, every domain type and face type: two independent restatements, bit for bit."""
    m, wl, arr = random_walls_case(seed)
    ok = O.OracleKM(wl.charge, wl.mass, [m])
    km = pyref.KM(wl.charge, wl.mass, [py_mesh(m)])
    ok.addParticles(0, arr, wl.dt)
    for q in range(len(arr["x"])):
        km.addParticle(0, pyref.Particle([arr["x"][q], arr["y"][q], arr["z"][q]], [arr["u"][q], arr["v"][q], arr["w"][q]], arr["mpw"][q]), wl.dt)
    total_hits = 0
    for _ in range(6):
        ok.updateFields(wl.dt)
        km.updateFields(wl.dt)
        assert (ok.n_absorbed, ok.n_exited, len(ok.slow)) == (km.n_absorbed, km.n_exited, len(km.slow))
        h = ok.hits[0]
        assert len(h["seg"]) == len(km.hits)
        for q, (sid, ts, vel, mpw, alive) in enumerate(km.hits):
            assert (int(h["seg"][q]), float(h["t"][q]), float(h["u"][q]), float(h["v"][q]), float(h["w"][q]), float(h["mpw"][q]), bool(h["alive"][q])) == \
                   (sid, ts, vel[0], vel[1], vel[2], mpw, alive)
        total_hits += len(km.hits)
        p = ok.sorted_parts(0)
        q = sorted(km.particles[0], key=lambda a: a.id)
        assert len(q) == len(p["x"])
        for key, get in (("x", lambda a: a.pos[0]), ("y", lambda a: a.pos[1]), ("z", lambda a: a.pos[2]), ("u", lambda a: a.vel[0]), ("v", lambda a: a.vel[1]),
                         ("w", lambda a: a.vel[2]), ("li", lambda a: a.lc[0]), ("lj", lambda a: a.lc[1]), ("dt", lambda a: a.dt)):
            assert np.array_equal(p[key], np.array([get(a) for a in q])), key
    assert total_hits > 30


def handoff_walls_case(kind):
    """Two RZ meshes stacked in z (MESH faces between them), a two-segment wall on either side of the interface, a SINK line further up; a drifting population."""
    from starfish_b200.domain import DomainBoundaryType as BC, Face
    a = UniformMesh(13, 11, (0.0, 0.0), (1e-3, 1e-3), DomainType.RZ)
    b = UniformMesh(7, 13, (0.0, 10e-3), (2e-3, 0.5e-3), DomainType.RZ)
    for m in (a, b):
        m.setMeshBCType(Face.LEFT, BC.SYMMETRY)
    for i in range(a.ni):
        a.setNeighbor(Face.TOP, i, 0, 1)
    for i in range(b.ni):
        b.setNeighbor(Face.BOTTOM, i, 0, 0)
    set_boundaries(a, [SolidBoundary("wa", np.array([[2e-3, 8.4e-3], [6e-3, 9.3e-3], [9e-3, 7.7e-3]]), kind=kind)])
    set_boundaries(b, [SolidBoundary("wb", np.array([[1e-3, 11.2e-3], [5e-3, 10.6e-3], [10e-3, 12.1e-3]]), kind=kind),
                       SolidBoundary("sink", np.array([[0.5e-3, 14e-3], [11e-3, 14.4e-3]]), kind=1, sink=True)])
    wl = S.Workload("t", a, 1e-7, S.QE, 16 * S.AMU, 19, vth_cells=0.5, drift_cells=(0.0, 1.3), kick_frac=0.0)
    arr = wl.particles(0, 1500)
    arr["x"] = np.abs(arr["x"]) * 0.9 + 1e-5
    return [a, b], wl, arr


@pytest.mark.parametrize("kind", [0, 1], ids=["absorb", "keep"])
def test_oracle_matches_python_restatement_walls_next_to_a_mesh_handoff(kind):
    """Two RZ meshes stacked in z with different spacings (MESH faces, KM:708-722) and a wall on either side of the interface: a particle can hit a
    surviving segment, finish its step across the MESH face and hit the neighbour's wall in the transfer sweep (KM:131-142) -- both restatements, bit for bit."""
    (a, b), wl, arr = handoff_walls_case(kind)

    def pm(m):
        q = py_mesh(m)
        for f in range(4):
            q.nbr[f] = [[None if v < 0 else int(v) for v in row] for row in m.nbr[f]]
        return q
    ok = O.OracleKM(wl.charge, wl.mass, [a, b])
    km = pyref.KM(wl.charge, wl.mass, [pm(a), pm(b)])
    ok.addParticles(0, arr, wl.dt)
    for q in range(len(arr["x"])):
        km.addParticle(0, pyref.Particle([arr["x"][q], arr["y"][q], arr["z"][q]], [arr["u"][q], arr["v"][q], arr["w"][q]], arr["mpw"][q]), wl.dt)
    hits = [0, 0]
    for _ in range(14):
        ok.updateFields(wl.dt)
        km.updateFields(wl.dt)
        assert (ok.n_absorbed, ok.n_exited, len(ok.slow)) == (km.n_absorbed, km.n_exited, len(km.slow))
        got = sorted((int(h["mesh"][q]) if "mesh" in h else k, int(h["seg"][q]), float(h["t"][q]), float(h["u"][q]), float(h["v"][q]), float(h["mpw"][q]), bool(h["alive"][q]))
                     for k, h in enumerate(ok.hits) for q in range(len(h["seg"])))
        for k, h in enumerate(ok.hits):
            hits[k] += len(h["seg"])
        assert len(got) == len(km.hits)
        for k in range(2):
            p = ok.sorted_parts(k)
            q = sorted(km.particles[k], key=lambda a_: a_.id)
            assert len(q) == len(p["x"])
            for key, get in (("x", lambda a_: a_.pos[0]), ("y", lambda a_: a_.pos[1]), ("u", lambda a_: a_.vel[0]), ("v", lambda a_: a_.vel[1]),
                             ("w", lambda a_: a_.vel[2]), ("li", lambda a_: a_.lc[0]), ("lj", lambda a_: a_.lc[1]), ("dt", lambda a_: a_.dt)):
                assert np.array_equal(p[key], np.array([get(a_) for a_ in q])), (k, key)
    assert hits[0] > 10 and hits[1] > 10 and ok.getNp(1) > 0


def test_kat_specular_as_the_reference_writes_it():
    """SurfaceImpactSpecular (SurfaceInteraction.java:104-149) adds normal * |vel_xy| * sqrt(2) to the in-plane velocity; LinearSegment.normal = (-dy, dx) / length
    (LinearSegment.java:26-44).  One particle flying straight down onto a horizontal wall that runs along +x (normal +y): u unchanged, v = -|v| + |v| sqrt 2."""
    m = S.make_mesh(11, 11, DomainType.XY, 1e-3, "open")
    m.efi[:] = 0.0
    m.efj[:] = 0.0
    set_boundaries(m, [SolidBoundary("floor", np.array([[1e-3, 4.25e-3], [9e-3, 4.25e-3]]), kind=2)])
    ok = O.OracleKM(S.QE, 16 * S.AMU, [m])
    v0 = -7000.0
    arr = dict(x=np.array([5.2e-3]), y=np.array([4.6e-3]), z=np.zeros(1), u=np.zeros(1), v=np.array([v0]), w=np.array([300.0]), mpw=np.array([2.0]))
    ok.addParticles(0, arr, 1e-7)
    ok.updateFields(1e-7)  # 0.7 mm of travel: crosses y = 4.25 mm at half of the step
    h = ok.hits[0]
    assert len(h["seg"]) == 1 and bool(h["alive"][0]) and ok.n_absorbed == 0
    mag = np.sqrt(0.0 * 0.0 + v0 * v0) * np.sqrt(2.0)
    want_v = v0 + 1.0 * mag
    p = ok.sorted_parts(0)
    assert p["u"][0] == 0.0 and p["v"][0] == want_v and p["w"][0] == 300.0 and h["v"][0] == want_v
    t_hit = (4.6e-3 - 4.25e-3) / 7e-4 * 0.9999
    assert abs(p["y"][0] - (4.6e-3 + v0 * 1e-7 * t_hit + want_v * 1e-7 * (1 - t_hit))) < 1e-12  # the rest of the step runs with the new velocity
