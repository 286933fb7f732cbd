"""Parity of the CUDA path (through the C ABI of libstarfish_gpu.so) against the CPU oracle.

Bars (BASELINE.json north_star): particle positions/velocities bit exact in XY (the theta coordinate pos[2]
of axisymmetric runs goes through asin/acos: 1e-12 relative), cell indices and particle counts bit exact,
deposited fields within 1e-10 relative (atomic summation order differs).
"""
import numpy as np
import pytest

from oracle import oracle as O
from starfish_b200 import KineticMaterial, Particles, synthetic as S
from starfish_b200 import _lib
from starfish_b200.domain import DomainBoundaryType as BC, DomainType, Face, UniformMesh

pytestmark = pytest.mark.gpu

PATHS = [pytest.param(_lib.STEP_GENERIC, id="generic"), pytest.param(_lib.STEP_INPLACE, id="tiled"), pytest.param(_lib.STEP_STREAM, id="stream")]


def to_particles(arr):
    return Particles(len(arr["x"]), **arr)


def field_close(got, want, rtol=1e-10):
    scale = np.abs(want).max()
    return np.allclose(got, want, rtol=rtol, atol=rtol * scale)


def compare_state(km, ok, theta_tol=1e-12):
    """Particles by id, per mesh."""
    for k, m in enumerate(km.meshes):
        g = km.getParticles(m).sorted_by_id()
        o = ok.sorted_parts(k)
        assert g.n == len(o["x"]), f"mesh {k}: np {g.n} vs oracle {len(o['x'])}"
        assert np.array_equal(g.id, o["id"])
        for key in ("x", "y", "u", "v", "w", "mpw", "li", "lj", "dt"):
            a, b = getattr(g, key), o[key]
            assert np.array_equal(a, b, equal_nan=True), f"mesh {k} field {key}: {np.sum(a != b)} of {g.n} differ"
        if m.domain_type == DomainType.XY:
            assert np.array_equal(g.z, o["z"])
        else:
            assert np.allclose(g.z, o["z"], rtol=theta_tol, atol=1e-300)
        assert np.array_equal(g.li.astype(np.int64), o["li"].astype(np.int64))
        assert np.array_equal(g.lj.astype(np.int64), o["lj"].astype(np.int64))


def compare_fields(km, ok):
    for k in range(len(km.meshes)):
        dep, raw = km.last_deposit[k], ok.raw[k]
        for f in range(7):
            assert field_close(dep[f], raw[f]), f"mesh {k} raw field {_lib.FIELD_NAMES[f]}"
        assert np.array_equal(dep[7], raw[7]), "mpc (cell counts) must be exact"
        for name in ("nd", "u", "v", "w", "count-sum", "u-sum", "uu-sum", "ww-sum", "mpc-sum"):
            assert field_close(km.fields[k][name], ok.fields[k][name]), name
    assert np.allclose([km.mass_sum, *km.momentum_sum, km.energy_sum], [ok.mass_sum, *ok.momentum_sum, ok.energy_sum], rtol=1e-10,
                       atol=1e-10 * abs(ok.energy_sum))
    assert km.n_exited == ok.n_exited
    assert km.getNp() == ok.getNp()


def make_pair(meshes, wl, arrays, flags, dom=None):
    dom = meshes[0].domain_type if dom is None else dom
    km = KineticMaterial("ion", wl.charge, wl.mass, meshes, dom, step_flags=flags)
    ok = O.OracleKM(wl.charge, wl.mass, meshes)
    km.dt = wl.dt
    for k, arr in enumerate(arrays):
        if arr is None:
            continue
        assert km.addParticles(meshes[k], to_particles(arr), wl.dt) == ok.addParticles(k, arr, wl.dt)
    return km, ok


@pytest.mark.parametrize("flags", PATHS)
@pytest.mark.parametrize("dom,bc", [(DomainType.XY, "periodic"), (DomainType.XY, "open"), (DomainType.XY, "symmetry"),
                                    (DomainType.RZ, "beam"), (DomainType.RZ, "symmetry"), (DomainType.ZR, "open")])
def test_move_and_deposit_match_oracle(dom, bc, flags):
    m = S.make_mesh(67, 45, dom, 1e-3, bc)
    wl = S.Workload("t", m, 1e-7, S.QE, 16 * S.AMU, 42, vth_cells=0.7, kick_frac=0.1)
    arr = wl.particles(0, 20000)
    km, ok = make_pair([m], wl, [arr], flags)
    with km:
        compare_state(km, ok)  # injection: XtoL + -0.5dt rewind
        for _ in range(5):
            km.updateFields()
            ok.updateFields(wl.dt)
            compare_state(km, ok)
            compare_fields(km, ok)
        assert km.num_samples == ok.num_samples == 5
        km.clearSamples()  # KM:1509-1528
        assert km.num_samples == 0 and not km.fields[0]["count-sum"].any() and not km.fields[0]["mpc-sum"].any()


@pytest.mark.parametrize("knob", ["SFGPU_STREAM_SORT", "SFGPU_HYBRID"])
def test_streaming_resort_and_hybrid_schedule(knob, monkeypatch):
    """Default (tiled) path with its periodic re-sort done by the streaming pass k_stream_sort, and with the step in which a
    re-sort is due run by the streaming step kernel instead: order-only changes, same bars."""
    monkeypatch.setenv(knob, "1")
    monkeypatch.setenv("SFGPU_STREAM_CHECK", "1")
    m = S.make_mesh(70, 50, DomainType.XY, 1e-3, "open")
    wl = S.Workload("t", m, 1e-7, S.QE, 16 * S.AMU, 5, vth_cells=0.6, kick_frac=0.1)
    arr = wl.particles(0, 30000)
    km, ok = make_pair([m], wl, [arr], 0)
    with km:
        km.setSortInterval(2)
        for _ in range(7):
            km.updateFields()
            ok.updateFields(wl.dt)
        compare_state(km, ok)
        compare_fields(km, ok)


@pytest.mark.parametrize("path", ["stream", "tiled"])
def test_step_kernels_alternate_from_step_to_step(path, monkeypatch):
    """The step flags are interchangeable from one step to the next (sfgpu.h): a tiled in-place step right after a streaming
    step must see work items of the slab the streaming step wrote (round-1 advisor finding: under SFGPU_PATH=stream the items
    still described the previous slab), with injection and exits in between."""
    monkeypatch.setenv("SFGPU_PATH", path)
    monkeypatch.setenv("SFGPU_STREAM_CHECK", "1")
    m = S.make_mesh(70, 50, DomainType.XY, 1e-3, "open")
    wl = S.Workload("t", m, 1e-7, S.QE, 16 * S.AMU, 6, vth_cells=0.6, kick_frac=0.1)
    km, ok = make_pair([m], wl, [wl.particles(0, 30000)], 0)
    seq = [_lib.STEP_STREAM, _lib.STEP_INPLACE, _lib.STEP_INPLACE, _lib.STEP_STREAM, 0, _lib.STEP_INPLACE, _lib.STEP_STREAM, _lib.STEP_STREAM, _lib.STEP_INPLACE]
    with km:
        km.setSortInterval(5)
        for it, fl in enumerate(seq):
            if it in (2, 6):
                arr = wl.particles(100000 * (it + 1), 3000)
                assert km.addParticles(m, to_particles(arr), wl.dt) == ok.addParticles(0, arr, wl.dt)
            km.step_flags = fl
            km.updateFields()
            ok.updateFields(wl.dt)
            compare_state(km, ok)
            compare_fields(km, ok)


@pytest.mark.parametrize("flags", PATHS)
def test_fast_particles_many_bounces_and_residual_dt(flags):
    """CFL >> 1 on a tiny symmetric box: >10 bounces leaves dt > 0 that is added to the next step (KM:333, :360)."""
    m = S.make_mesh(5, 4, DomainType.XY, 1e-3, "symmetry")
    wl = S.Workload("t", m, 1e-7, S.QE, 16 * S.AMU, 9, vth_cells=25.0, kick_frac=0.05)
    arr = wl.particles(0, 3000)
    km, ok = make_pair([m], wl, [arr], flags)
    with km:
        for _ in range(4):
            km.updateFields()
            ok.updateFields(wl.dt)
            compare_state(km, ok)
            compare_fields(km, ok)
        assert (ok.parts[0]["dt"] > 0).any(), "case must exercise the residual dt"


@pytest.mark.parametrize("flags", PATHS)
def test_boris_rotation(flags):
    m = S.make_mesh(33, 33, DomainType.XY, 1e-3, "periodic")
    wl = S.Workload("t", m, 1e-9, -S.QE, 9.109e-31, 5, vth_cells=0.4, kick_frac=0.05)
    m.bfi = np.full((33, 33), 0.02)
    m.bfj = np.linspace(-0.01, 0.03, 33 * 33).reshape(33, 33)
    arr = wl.particles(0, 5000)
    km, ok = make_pair([m], wl, [arr], flags)
    with km:
        for _ in range(3):
            km.updateFields()
            ok.updateFields(wl.dt)
            compare_state(km, ok)
            compare_fields(km, ok)


@pytest.mark.parametrize("flags", PATHS)
def test_mesh_handoff(flags):
    """Two RZ meshes joined by MESH faces, different spacings: transfer sweeps (KM:131-142, :708-722)."""
    a = UniformMesh(33, 33, (0.0, 0.0), (1e-3, 1e-3), DomainType.RZ)
    b = UniformMesh(17, 41, (0.0, 32e-3), (2e-3, 0.5e-3), DomainType.RZ)
    for mm in (a, b):
        mm.setMeshBCType(Face.LEFT, BC.SYMMETRY)
    for i in range(a.ni):
        a.setNeighbor(Face.TOP, i, 0, 1)
    for i in range(b.ni):
        b.setNeighbor(Face.BOTTOM, i, 0, 0)
    wl = S.Workload("t", a, 1e-7, S.QE, 16 * S.AMU, 11, vth_cells=0.3, drift_cells=(0.0, 0.9), kick_frac=0.0)
    arr = wl.particles(0, 8000)
    km, ok = make_pair([a, b], wl, [arr, None], flags)
    with km:
        for _ in range(30):
            km.updateFields()
            ok.updateFields(wl.dt)
        compare_state(km, ok)
        compare_fields(km, ok)
        assert km.getNp(b) == ok.getNp(1) > 0


@pytest.mark.parametrize("flags", PATHS)
def test_multi_domain_example_layout(flags):
    """BASELINE config 4: the four-mesh RZ layout of dat/examples/multi-domain (different spacings, MESH faces found by
    Mesh.setMeshNeighbors, LEFT symmetry, Ar+): particles loaded in every mesh wander across the hand-off faces and out."""
    meshes, charge, mass, dt = S.config_multi_domain()
    assert sum(int((np.asarray(mm.bc[int(f)]) == int(BC.MESH)).any()) for mm in meshes for f in Face) >= 5  # the layout is connected
    arrays, oks = [], None
    for k, mm in enumerate(meshes):
        wl = S.Workload("t%d" % k, mm, dt, charge, mass, 100 + k, vth_cells=0.35, drift_cells=(0.1, 0.25), kick_frac=0.0)
        arrays.append(wl.particles(0, 3000))
    km, ok = make_pair(meshes, wl, arrays, flags)
    with km:
        for _ in range(40):
            km.updateFields()
            ok.updateFields(dt)
        compare_state(km, ok)
        compare_fields(km, ok)
        assert ok.n_exited > 0 and all(ok.getNp(k) > 0 for k in range(len(meshes)))


@pytest.mark.parametrize("flags", PATHS)
def test_slow_path_classification(flags):
    """Particles whose substep bounding box touches a segment node are handed to the host untouched (KM:504-518)."""
    m = S.make_mesh(40, 40, DomainType.XY, 1e-3, "open")
    m.has_seg[18:22, 18:22] = 1
    wl = S.Workload("t", m, 1e-7, S.QE, 16 * S.AMU, 21, vth_cells=0.5, kick_frac=0.05)
    arr = wl.particles(0, 10000)
    km, ok = make_pair([m], wl, [arr], flags)
    taken = {}
    km.slow_path_handler = lambda k_, slow, extra: (taken.update(slow=slow, extra=extra), [])[1]
    with km:
        km.updateFields()
        ok.updateFields(wl.dt)
        compare_state(km, ok)
        compare_fields(km, ok)
        assert ok.slow and taken["slow"].n == sum(len(s[1]["x"]) for s in ok.slow)
        g = taken["slow"]
        order = np.argsort(g.id)
        _mid, op, oaux = ok.slow[0]
        oo = np.argsort(op["id"])
        for key in ("x", "y", "z", "u", "v", "w", "li", "lj", "dt"):
            assert np.array_equal(getattr(g, key)[order], op[key][oo]), key
        for key in ("old_x", "old_y", "old_li", "old_lj", "bounces"):
            assert np.array_equal(taken["extra"][key][order], oaux[key][oo]), key


@pytest.mark.parametrize("flags", PATHS)
def test_slow_path_round_trip_with_survivors(flags):
    """Full slow-path protocol: sfgpu_step(DEFER_FINISH) -> sfgpu_take_slowpath -> the host runs ProcessBoundary and the
    remaining sub-steps (here: the Python restatement of the Java, with segment nodes that carry no real segment) ->
    sfgpu_inject(DEPOSIT_NOW) -> sfgpu_finish_step.  With fictitious segments the end state must equal a run on a
    mesh without any segment node, bit for bit."""
    import pyref
    m = S.make_mesh(36, 30, DomainType.XY, 1e-3, "symmetry")
    m.has_seg[10:14, 8:12] = 1
    m.has_seg[25, :] = 1
    wl = S.Workload("t", m, 1e-7, S.QE, 16 * S.AMU, 77, vth_cells=1.3, kick_frac=0.1)
    arr = wl.particles(0, 8000)
    plain = S.make_mesh(36, 30, DomainType.XY, 1e-3, "symmetry")  # same mesh, no segment nodes
    plain.efi, plain.efj = m.efi, m.efj
    km = KineticMaterial("ion", wl.charge, wl.mass, [m], DomainType.XY, step_flags=flags)
    ok = O.OracleKM(wl.charge, wl.mass, [plain])
    km.dt = wl.dt
    pm = pyref.Mesh(m.ni, m.nj, m.x0, m.dh, 0)
    for f in range(4):
        pm.bc[f] = [int(v) for v in m.bc[f]]
    host = pyref.KM(wl.charge, wl.mass, [pm])  # has_seg all zero: no real segments anywhere
    handed = []

    def handler(_km, slow, extra):
        keep = []
        for q in range(slow.n):
            part = pyref.Particle([slow.x[q], slow.y[q], slow.z[q]], [slow.u[q], slow.v[q], slow.w[q]], slow.mpw[q], int(slow.id[q]))
            part.lc, part.dt = [slow.li[q], slow.lj[q]], slow.dt[q]
            if host.finish_slow(0, part, [extra["old_x"][q], extra["old_y"][q]], [extra["old_li"][q], extra["old_lj"][q]], int(extra["bounces"][q])):
                keep.append((part, int(slow.born_it[q])))
        handed.append(slow.n)
        n = len(keep)
        col = lambda f: np.array([f(p) for p, _b in keep], dtype=np.float64)
        surv = Particles(n, x=col(lambda p: p.pos[0]), y=col(lambda p: p.pos[1]), z=col(lambda p: p.pos[2]), u=col(lambda p: p.vel[0]),
                         v=col(lambda p: p.vel[1]), w=col(lambda p: p.vel[2]), mpw=col(lambda p: p.mpw), li=col(lambda p: p.lc[0]),
                         lj=col(lambda p: p.lc[1]), dt=col(lambda p: p.dt), id=np.array([p.id for p, _b in keep], np.int32),
                         born_it=np.array([b for _p, b in keep], np.int32))
        return [(0, surv)]

    km.slow_path_handler = handler
    with km:
        assert km.addParticles(m, to_particles(arr), wl.dt) == ok.addParticles(0, arr, wl.dt)
        for _ in range(4):
            host.n_exited = 0
            km.updateFields()
            ok.updateFields(wl.dt)
            assert km.n_exited + host.n_exited == ok.n_exited
            km.n_exited = ok.n_exited  # exits seen by the host are the host's to count
            compare_state(km, ok)
            compare_fields(km, ok)
        assert sum(handed) > 100, "the case must send particles through the slow path"


@pytest.mark.parametrize("flags", PATHS)
def test_injection_every_step_zero_weight_and_edges(flags):
    m = S.make_mesh(30, 20, DomainType.XY, 1e-3, "open")
    wl = S.Workload("t", m, 1e-7, S.QE, 16 * S.AMU, 33, vth_cells=0.8, kick_frac=0.1)
    km, ok = make_pair([m], wl, [None], flags)
    with km:
        km.updateFields()  # empty store
        ok.updateFields(wl.dt)
        assert km.getNp() == 0
        first = 0
        for step in range(6):
            arr = wl.particles(first, 1500)
            first += 1500
            if step == 2:
                arr["mpw"][::7] = 0.0  # removed at the next move (KM:322)
                arr["x"][:5] = m.x0[0] + (m.ni - 1) * m.dh[0]  # exactly on the plus edge: gather_safe, no deposit
                arr["x"][5:8] = m.x0[0] + (m.ni + 2.5) * m.dh[0]  # beyond: lc clamp of KM:770-773
            assert km.addParticles(m, to_particles(arr), wl.dt) == ok.addParticles(0, arr, wl.dt)
            km.updateFields()
            ok.updateFields(wl.dt)
            compare_state(km, ok)
            compare_fields(km, ok)


@pytest.mark.parametrize("every", [1, 3])
@pytest.mark.parametrize("bc", ["open", "periodic"])
def test_fused_counting_pass_of_the_sort(every, bc, monkeypatch):
    """SFGPU_FUSE_COUNT=1: the tiled step before a cell sort writes the sort's keys, ranks and histogram itself; deferred particles (exits, wraps) and the
    injected tail are keyed by k_sort_count_fix when the sort runs.  Order only: same results."""
    monkeypatch.setenv("SFGPU_FUSE_COUNT", "1")
    m = S.make_mesh(61, 47, DomainType.XY, 1e-3, bc)
    wl = S.Workload("t", m, 1e-7, S.QE, 16 * S.AMU, 5, vth_cells=0.6, kick_frac=0.1)
    arr = wl.particles(0, 40000)
    km, ok = make_pair([m], wl, [arr], _lib.STEP_INPLACE)
    with km:
        km.setSortInterval(every)
        first = 40000
        for step in range(8):
            if step % 2:
                extra = wl.particles(first, 700)
                first += 700
                assert km.addParticles(m, to_particles(extra), wl.dt) == ok.addParticles(0, extra, wl.dt)
            km.updateFields()
            ok.updateFields(wl.dt)
            compare_state(km, ok)
            compare_fields(km, ok)


@pytest.mark.parametrize("flags", PATHS)
def test_explicit_lc_injection_goes_through_records(flags):
    """addParticle(md, part) with a caller-supplied lc and residual dt (restart load, KM:953-1000): full records."""
    m = S.make_mesh(40, 33, DomainType.XY, 1e-3, "symmetry")
    wl = S.Workload("t", m, 1e-7, S.QE, 16 * S.AMU, 8, vth_cells=0.6, kick_frac=0.1)
    arr = wl.particles(0, 6000)
    arr["li"] = (arr["x"] - m.x0[0]) / m.dh[0]
    arr["lj"] = (arr["y"] - m.x0[1]) / m.dh[1]
    arr["li"][::3] = np.floor(arr["li"][::3])  # stale lc: differs from XtoL(pos)
    arr["dt"] = np.zeros(6000)
    km = KineticMaterial("ion", wl.charge, wl.mass, [m], DomainType.XY, step_flags=flags)
    ok = O.OracleKM(wl.charge, wl.mass, [m])
    km.dt = wl.dt
    with km:
        assert km.addParticles(m, to_particles(arr), wl.dt) == ok.addParticles(0, arr, wl.dt)
        compare_state(km, ok)
        for _ in range(4):
            km.updateFields()
            ok.updateFields(wl.dt)
            compare_state(km, ok)
            compare_fields(km, ok)


BOTH = [pytest.param(0, id="default"), pytest.param(_lib.STEP_STREAM, id="stream")]


@pytest.mark.parametrize("flags", BOTH)
@pytest.mark.parametrize("every", [1, 3, 100])
def test_sort_interval_does_not_change_results(every, flags):
    m = S.make_mesh(70, 50, DomainType.XY, 1e-3, "periodic")
    wl = S.Workload("t", m, 1e-7, S.QE, 16 * S.AMU, 4, vth_cells=0.9, kick_frac=0.1)
    arr = wl.particles(0, 30000)
    km, ok = make_pair([m], wl, [arr], flags)
    with km:
        km.setSortInterval(every)
        for _ in range(7):
            km.updateFields()
            ok.updateFields(wl.dt)
        compare_state(km, ok)
        compare_fields(km, ok)


@pytest.mark.parametrize("halo,vth", [(1, 0.15), (1, 0.9), (2, 0.15), (2, 0.9), (0, 0.15), (0, 0.9)])
def test_tile_halo_is_a_speed_choice_only(halo, vth):
    """sfgpu_set_tile_halo: deposits that miss the narrow accumulation tile are made by the deferred kernel -- same results for either
    halo and any thermal spread; the automatic mode widens the tile for good once a step leaves > 0.5 % of its deposits to it."""
    m = S.make_mesh(90, 75, DomainType.XY, 1e-3, "periodic")
    wl = S.Workload("t", m, 1e-7, S.QE, 16 * S.AMU, 11, vth_cells=vth, kick_frac=0.1)
    arr = wl.particles(0, 60000)
    km, ok = make_pair([m], wl, [arr], _lib.STEP_INPLACE)
    with km:
        km.setTileHalo(halo)
        assert km.tileHalo() == (halo if halo else 1, halo == 0)
        missed = 0
        for _ in range(7):
            km.updateFields()
            ok.updateFields(wl.dt)
            missed = max(missed, km.lastStepFallback())
            compare_state(km, ok)
            compare_fields(km, ok)
        if halo:
            assert km.tileHalo() == (halo, False)
        elif vth > 0.5:
            assert km.tileHalo() == (2, True) and missed * 200 > 60000
        else:
            assert km.tileHalo() == (1, True) and missed * 200 <= 60000


@pytest.mark.parametrize("cold,v_drift", [(False, 7000.0), (True, -7000.0)])
@pytest.mark.parametrize("flags", PATHS)
def test_device_side_uniform_source(flags, cold, v_drift):
    """UniformSource sampled on the device with java.util.Random's draws (LCG jump-ahead per particle) against the
    oracle's sequential loop: same particles in the same meshes with the same ids, same RNG state afterwards, and the same
    simulation when sampling and stepping alternate (inlet partly outside every mesh, two overlapping meshes)."""
    from starfish_b200.domain import LinearSpline
    a = S.make_mesh(21, 11, DomainType.XY, 5e-3, "open", x0=(-0.1, 0.0))
    b = S.make_mesh(11, 9, DomainType.XY, 5e-3, "open", x0=(-0.1, 0.05))
    wl = S.Workload("t", a, 1e-7, S.QE, 16 * S.AMU, 9, kick_frac=0.05)
    pts = [(-0.1, 0.12), (-0.1, 0.03), (-0.09, -0.01), (-0.09, -0.02)]
    sp = LinearSpline(pts if v_drift > 0 else pts[::-1])  # a negative drift along the flipped normal: -0.0 vs +0.0 in vel[2]
    km, ok = make_pair([a, b], wl, [None, None], flags, dom=DomainType.XY)
    state_g = state_o = O.java_seed(2026)
    with km:
        for it in range(6):
            n_g, state_g = km.sampleUniformSource(sp, v_drift, 3001, state_g, dt=wl.dt, mpw=1e3, born_it=it, cold_beam=cold)
            n_o, state_o = ok.sampleUniformSource(sp, v_drift, 3001, wl.dt, state_o, 1e3, born_it=it, cold_beam=cold)
            assert n_g == n_o and state_g == state_o and 0 < n_g < 3001
            compare_state(km, ok)
            km.updateFields()
            ok.updateFields(wl.dt)
            compare_state(km, ok)
            compare_fields(km, ok)
        g = km.getParticles(a)
        assert sorted(set(g.born_it.tolist())) == list(range(6))


@pytest.mark.parametrize("flags", BOTH)
@pytest.mark.parametrize("dom", [DomainType.RZ, DomainType.ZR])
def test_device_side_uniform_source_axisymmetric(dom, flags):
    """Spline.randomT in RZ / ZR (secant search over the frustum area, Spline.java:594-637) on the device against the oracle."""
    from starfish_b200.domain import LinearSpline
    m = S.make_mesh(21, 31, dom, 1e-3, "open")
    wl = S.Workload("t", m, 1e-7, S.QE, 16 * S.AMU, 9, kick_frac=0.05)
    pts = [(0.0021, 1e-9), (0.0102, 1e-9), (0.0198, 0.0007)] if dom == DomainType.RZ else [(1e-9, 0.0198), (1e-9, 0.0102), (0.0007, 0.0021)]
    sp = LinearSpline(pts, dom)
    km, ok = make_pair([m], wl, [None], flags, dom=dom)
    state_g = state_o = O.java_seed(99)
    with km:
        for it in range(5):
            n_g, state_g = km.sampleUniformSource(sp, 5000.0, 2500, state_g, dt=wl.dt, mpw=1e3, born_it=it)
            n_o, state_o = ok.sampleUniformSource(sp, 5000.0, 2500, wl.dt, state_o, 1e3, born_it=it)
            assert n_g == n_o > 0 and state_g == state_o
            compare_state(km, ok)
            km.updateFields()
            ok.updateFields(wl.dt)
            compare_state(km, ok)
            compare_fields(km, ok)


@pytest.mark.parametrize("flags", BOTH)
def test_restart_records_round_trip(flags):
    """restart.bin particle section (KM:904-1000): the saved bytes are the DataOutputStream layout (checked with Python's
    big-endian struct), and loading them goes through addParticle like the reference (rewind re-applied, ids renumbered)."""
    import struct
    m = S.make_mesh(6, 5, DomainType.XY, 1e-3, "symmetry")
    wl = S.Workload("t", m, 1e-7, S.QE, 16 * S.AMU, 55, vth_cells=25.0, kick_frac=0.05)  # >10 bounces: records with residual dt
    arr = wl.particles(0, 5000)
    km, ok = make_pair([m], wl, [arr], flags)
    with km:
        for _ in range(3):
            km.updateFields()
            ok.updateFields(wl.dt)
        data = km.saveRestartParticles(m)
        p = km.getParticles(m)  # same store order as the stream
        (np_,) = struct.unpack(">q", data[:8])
        assert np_ == p.n == ok.getNp() and len(data) == 8 + 96 * p.n
        rec = np.frombuffer(data, dtype=np.dtype([("d", ">f8", 11), ("born", ">i4"), ("id", ">i4")]), offset=8)
        for col, key in enumerate(("x", "u", "y", "v", "z", "w", "li", "lj", "dt", "mpw")):
            assert np.array_equal(rec["d"][:, col], getattr(p, key)), key
        assert np.all(rec["d"][:, 10] == wl.mass) and np.array_equal(rec["id"], p.id) and np.array_equal(rec["born"], p.born_it)
        assert (p.dt > 0).any() or (p.li != (p.x - m.x0[0]) / m.dh[0]).any(), "case should include exceptional records"
        # load into a fresh material and into a fresh oracle through addParticle(md, part)
        km2, ok2 = make_pair([m], wl, [None], flags)
        with km2:
            used, added = km2.loadRestartParticles(m, data + b"FIELDS...", wl.dt)
            assert used == len(data)
            src = {k: np.ascontiguousarray(getattr(p, k)) for k in ("x", "y", "z", "u", "v", "w", "mpw", "li", "lj", "dt")}
            assert added == ok2.addParticles(0, src, wl.dt, rewind=True, born_it=p.born_it)
            compare_state(km2, ok2)
            km2.updateFields()
            ok2.updateFields(wl.dt)
            compare_state(km2, ok2)
            compare_fields(km2, ok2)


@pytest.mark.parametrize("flags", BOTH)
def test_download_upload_round_trip(flags):
    """The particle-store view of MCC / DSMC / output: download, mutate on the host, upload, keep stepping."""
    m = S.make_mesh(16, 16, DomainType.XY, 1e-3, "periodic")
    wl = S.Workload("t", m, 1e-7, S.QE, 16 * S.AMU, 1)
    arr = wl.particles(0, 1000)
    km, ok = make_pair([m], wl, [arr], flags)
    with km:
        km.updateFields()
        ok.updateFields(wl.dt)
        p = km.getParticles(m)
        p.u[:] *= 2.0
        km.setParticles(m, p)
        q = km.getParticles(m)
        assert np.array_equal(p.u, q.u) and np.array_equal(p.id, q.id)
        km.updateFields()  # the in-place edit invalidated the cell histogram: the streaming step must re-establish it
        assert km.getNp() == 1000
        assert km.last_deposit[0][7].sum() == 1000


@pytest.mark.parametrize("flags", BOTH)
@pytest.mark.parametrize("n", [1 << 24])
def test_full_size_properties_config_b(n, flags):
    """BASELINE config B at full size (512x512, 16M): size-independent properties instead of the oracle."""
    wl = S.config_b()
    m = wl.mesh
    km = KineticMaterial("O+", wl.charge, wl.mass, [m], DomainType.XY, capacity_hint=n, step_flags=flags)
    km.dt = wl.dt
    with km:
        chunk = 1 << 22
        for first in range(0, n, chunk):
            km.addParticles(m, to_particles(wl.particles(first, chunk)), wl.dt)
        assert km.getNp() == n
        for _ in range(3):
            km.updateFields()
        dep = km.last_deposit[0]
        assert km.getNp() == n and km.n_exited == 0  # periodic box: nothing leaves
        assert dep[7].sum() == n  # every particle counted once in mpc
        w_total = n * wl.mpw
        assert abs(dep[0].sum() - w_total) <= 1e-10 * w_total  # bilinear weights sum to 1
        assert abs(km.mass_sum / wl.mass - w_total) <= 1e-10 * w_total
        for f, s in ((1, km.momentum_sum[0]), (2, km.momentum_sum[1]), (3, km.momentum_sum[2])):
            scale = wl.mpw * n * wl.vth
            assert abs(dep[f].sum() - s / wl.mass) <= 1e-9 * scale  # checksum of checksums: sum_nodes U = sum_p mpw*u
        p = km.getParticles(m)
        lx = (m.ni - 1) * m.dh[0]
        assert p.x.min() >= 0 and p.x.max() <= lx and p.y.min() >= 0 and p.y.max() <= lx
        assert np.array_equal(np.sort(p.id), np.arange(n, dtype=np.int32))  # a permutation: nobody lost or duplicated


def _full_size_vs_oracle(wl, n, steps, flags, inject_every=0, theta_tol=1e-12):
    """Whole BASELINE configuration against the oracle (threads = host cores for the mover, serial deposit): ids, cell
    indices and x/y/u/v/w bit exact, deposit within 1e-10, cell counts exact."""
    import os
    m = wl.mesh
    km = KineticMaterial("O+", wl.charge, wl.mass, [m], m.domain_type, capacity_hint=n + (n // 8 if inject_every else 0), step_flags=flags)
    ok = O.OracleKM(wl.charge, wl.mass, [m], threads=min(os.cpu_count() or 1, 32))
    km.dt = wl.dt
    with km:
        chunk = 1 << 22
        for first in range(0, n, chunk):
            c = min(chunk, n - first)
            arr = wl.particles(first, c)
            assert km.addParticles(m, to_particles(arr), wl.dt) == ok.addParticles(0, arr, wl.dt)
        for it in range(steps):
            if inject_every and it % inject_every == 0:  # beam injection plane: the first half cell above z = 0
                c = n // 64
                arr = wl.particles(n + it * c, c)
                arr["y"] = m.x0[1] + (arr["y"] - m.x0[1]) * (0.5 / (m.nj - 1))
                assert km.addParticles(m, to_particles(arr), wl.dt) == ok.addParticles(0, arr, wl.dt)
            km.updateFields()
            ok.updateFields(wl.dt)
            assert km.getNp() == ok.getNp() and km.n_exited == ok.n_exited
        compare_fields(km, ok)
        g = km.getParticles(m).sorted_by_id()
        o = ok.sorted_parts(0)
        assert np.array_equal(g.id, o["id"])
        for key in ("x", "y", "u", "v", "w", "mpw", "li", "lj", "dt"):
            assert np.array_equal(getattr(g, key), o[key], equal_nan=True), key
        if m.domain_type == DomainType.XY:
            assert np.array_equal(g.z, o["z"])
        else:
            assert np.allclose(g.z, o["z"], rtol=theta_tol, atol=1e-300)
        return km.n_exited


@pytest.mark.parametrize("flags", BOTH)
def test_full_size_config_b_matches_oracle(flags):
    """BASELINE config B at full size (XY 512x512, 16,777,216 particles, periodic), 3 steps: every work-item split, tile
    boundary and >2048-particles-per-tile path of the tiled kernel against the oracle, not only by conservation."""
    _full_size_vs_oracle(S.config_b(), 1 << 24, 3, flags)


@pytest.mark.parametrize("flags", BOTH)
def test_config_c_beam_with_injection_and_exits_matches_oracle(flags):
    """BASELINE config C geometry (RZ 1024x1024, beam over r < 0.25 Rmax, LEFT symmetry, open exits) at 2^24 particles with
    beam injection every step; the rotation, the Ruyten weights, the symmetry axis and compaction at full mesh size."""
    exited = _full_size_vs_oracle(S.config_c(), 1 << 24, 4, flags, inject_every=1)
    assert exited > 0


def test_cell_lists_for_collision_consumers():
    """sfgpu_cell_lists sorts the store and tells where every cell's particles sit in download order."""
    m = S.make_mesh(37, 29, DomainType.XY, 1e-3, "symmetry")
    wl = S.Workload("t", m, 1e-7, S.QE, 16 * S.AMU, 9, vth_cells=0.8, kick_frac=0.1)
    with KineticMaterial("ion", wl.charge, wl.mass, [m], m.domain_type) as km:
        km.dt = wl.dt
        km.addParticles(m, to_particles(wl.particles(0, 40000)), wl.dt)
        for _ in range(4):
            km.updateFields()
        first, count, n_sorted = km.cellLists(m)
        p = km.getParticles(m)
        assert count.sum() == n_sorted <= p.n == km.getNp()
        ci, cj = p.li.astype(np.int64), p.lj.astype(np.int64)
        for i, j in [(0, 0), (5, 7), (35, 27), (17, 0), (20, 13)]:
            sl = slice(first[i, j], first[i, j] + count[i, j])
            assert count[i, j] == np.sum((ci[:n_sorted] == i) & (cj[:n_sorted] == j))
            assert np.all(ci[sl] == i) and np.all(cj[sl] == j)
        # the lists do not disturb the physics: one more step still matches the oracle
        ok = O.OracleKM(wl.charge, wl.mass, [m])
        ok.addParticles(0, wl.particles(0, 40000), wl.dt)
        for _ in range(5):
            ok.updateFields(wl.dt)
        km.updateFields()
        compare_state(km, ok)


def test_one_deferred_step_per_context():
    """The step counters (mover sums included) belong to the context: while one species has a deferred step open, stepping or
    injecting into another species of the same context is refused instead of silently clobbering them (round-1 advisor finding)."""
    import ctypes as C
    from starfish_b200.kinetic_material import SfgpuError
    m = S.make_mesh(20, 20, DomainType.XY, 1e-3, "open")
    wl = S.Workload("t", m, 1e-7, S.QE, 16 * S.AMU, 3, vth_cells=0.3)
    with KineticMaterial("a", wl.charge, wl.mass, [m], m.domain_type) as km:
        km.dt = wl.dt
        sp2 = C.c_int32(-1)
        km._check(km.lib.sfgpu_species_add(km._ctx, wl.charge, 2 * wl.mass, 0, C.byref(sp2)))
        km.addParticles(m, to_particles(wl.particles(0, 500)), wl.dt)
        km._check(km.lib.sfgpu_step(km._ctx, km._sp, wl.dt, _lib.STEP_DEFER_FINISH))
        rc = km.lib.sfgpu_step(km._ctx, sp2.value, wl.dt, 0)
        assert rc == -6 and b"deferred step open" in km.lib.sfgpu_last_error(km._ctx)
        km._check(km.lib.sfgpu_finish_step(km._ctx, km._sp))
        km._check(km.lib.sfgpu_step(km._ctx, sp2.value, wl.dt, 0))


@pytest.mark.parametrize("flags", PATHS)
def test_differential_cuda_vs_oracle_random_cases(flags):
    """hypothesis-drawn meshes (size, spacing, origin), per-face boundary types (OPEN / SYMMETRY / PERIODIC / DIRICHLET), time steps,
    charge sign, field amplitudes, thermal speeds from 0.05 to 2.5 cells per step, particles exactly on nodes / faces / the plus
    edge and with zero or negative weight: the CUDA path must match the oracle (state bit exact, deposit 1e-10) after every step."""
    pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st, HealthCheck
    bcs = [int(BC.OPEN), int(BC.SYMMETRY), int(BC.PERIODIC), int(BC.DIRICHLET)]

    @settings(max_examples=16, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
    @given(dom=st.sampled_from([DomainType.XY, DomainType.RZ, DomainType.ZR]), ni=st.integers(3, 40), nj=st.integers(3, 40),
           dhx=st.sampled_from([1e-3, 0.5e-3, 3.3e-4, 0.1]), dhy=st.sampled_from([1e-3, 2e-3, 7e-4, 0.25]),
           x0=st.sampled_from([0.0, -0.15, 0.013]), faces=st.lists(st.sampled_from(bcs), min_size=4, max_size=4),
           dt=st.sampled_from([1e-7, 3e-8, 1e-6]), neg=st.booleans(), seed=st.integers(0, 10**6), e_amp=st.sampled_from([0.0, 1e2, 1e5]),
           vth=st.sampled_from([0.05, 0.6, 2.5]))
    def run(dom, ni, nj, dhx, dhy, x0, faces, dt, neg, seed, e_amp, vth):
        ox = x0 if dom != DomainType.RZ else abs(x0)
        m = UniformMesh(ni, nj, (ox, 0.0), (dhx, dhy), dom)
        for f, t in zip(Face, faces):
            if t == int(BC.PERIODIC) and dom != DomainType.XY:
                t = int(BC.OPEN)
            m.setMeshBCType(f, BC(t))
        rng = np.random.default_rng(seed)
        m.efi = e_amp * rng.standard_normal((ni, nj))
        m.efj = e_amp * rng.standard_normal((ni, nj))
        n = 3000
        li, lj = rng.uniform(0, ni - 1, n), rng.uniform(0, nj - 1, n)
        li[:6] = np.array([0.0, ni - 1.0, 1.0, float(ni // 2), 0.0, ni - 1.0])
        lj[:6] = np.array([0.0, nj - 1.0, float(nj // 2), 1.0, nj - 1.0, 0.0])
        x, y = ox + li * dhx, lj * dhy
        if dom == DomainType.RZ:
            x = np.maximum(x, ox + 1e-9 * dhx)
        if dom == DomainType.ZR:
            y = np.maximum(y, 1e-9 * dhy)
        mpw = np.full(n, 1e3)
        mpw[6], mpw[7] = 0.0, -1.0
        arr = dict(x=x, y=y, z=np.zeros(n), u=vth * dhx / dt * rng.standard_normal(n), v=vth * dhy / dt * rng.standard_normal(n),
                   w=vth * dhx / dt * rng.standard_normal(n), mpw=mpw)

        class W:
            charge, mass = (-S.QE if neg else S.QE), 16 * S.AMU
        W.dt = dt
        km, ok = make_pair([m], W, [arr], flags)
        with km:
            km.setSortInterval(2)
            for _ in range(4):
                km.updateFields()
                ok.updateFields(dt)
                compare_state(km, ok)
                compare_fields(km, ok)

    run()
