"""BASELINE config 1 (dat/examples/tutorial/step2: ions streaming past an absorbing 20-segment cylinder) on the GPU.

1. device path: the segment part of ProcessBoundary runs inside the step kernels, nothing is handed to the host;
2. host path: the device only classifies, the host finishes those particles with the unchanged reference logic
   (here: the pure-Python restatement tests/pyref.py standing in for the Java ProcessBoundary).
Both against the oracle: particle state and counts bit exact, surface hits identical as a set, deposit within 1e-10."""
import numpy as np
import pytest

import pyref
from oracle import oracle as O
from starfish_b200 import KineticMaterial, Particles, _lib, synthetic as S
from test_gpu_parity import compare_fields, compare_state, to_particles
from test_segments import handoff_walls_case, py_mesh, random_walls_case

pytestmark = pytest.mark.gpu

PATHS = [pytest.param(_lib.STEP_GENERIC, id="generic"), pytest.param(_lib.STEP_INPLACE, id="tiled"), pytest.param(_lib.STEP_STREAM, id="stream")]


def _hit_key(h):
    o = np.lexsort((h["mpw"], h["v"], h["u"], h["t"], h["seg"]))
    return [(int(h["seg"][q]), float(h["t"][q]), float(h["u"][q]), float(h["v"][q]), float(h["w"][q]), float(h["mpw"][q]), int(h["alive"][q])) for q in o]


@pytest.mark.parametrize("flags", PATHS)
@pytest.mark.parametrize("wall_kind", [0, 1], ids=["absorb", "keep"])
def test_tutorial_step2_device_segments_match_oracle(flags, wall_kind):
    cfg = S.TutorialStep2(wall_kind=wall_kind)
    m = cfg.mesh
    ok = O.OracleKM(cfg.charge, cfg.mass, [m])
    steps = cfg.steps if wall_kind == 0 else 250
    with KineticMaterial("O+", cfg.charge, cfg.mass, [m], m.domain_type, spwt=cfg.spwt, step_flags=flags) as km:
        km.dt = cfg.dt
        state_o = state_g = O.java_seed(0)
        absorbed = hits = 0
        for it in range(steps):
            n_mp = cfg.num_mp()
            no, state_o = ok.sampleUniformSource(cfg.inlet, cfg.v_drift, n_mp, cfg.dt, state_o, cfg.spwt, born_it=it)
            ng, state_g = km.sampleUniformSource(cfg.inlet, cfg.v_drift, n_mp, state_g, born_it=it)
            assert (ng, state_g) == (no, state_o)
            km.updateFields()
            ok.updateFields(cfg.dt)
            h = km.takeSurfaceHits()
            assert km.n_slow == 0 and not ok.slow  # nothing goes back to the host
            assert km.n_absorbed == ok.n_absorbed and km.n_exited == ok.n_exited and km.getNp() == ok.getNp(), it
            assert _hit_key(h) == _hit_key(ok.hits[0]), it
            absorbed += km.n_absorbed
            hits += len(h["seg"])
            if it % 100 == 99 or it == steps - 1:
                compare_state(km, ok)
                compare_fields(km, ok)
        assert hits > 1000 and (absorbed == hits if wall_kind == 0 else absorbed == 0)


@pytest.mark.parametrize("flags", PATHS)
def test_tutorial_step2_host_slow_path_matches_oracle(flags):
    """Without the segment table the device hands every particle whose sub-step touches a segment node back UNMOVED; the host runs
    ProcessBoundary + the remaining sub-steps and re-injects the survivors into the still-open step."""
    cfg = S.TutorialStep2(spwt=4e3)
    m = cfg.mesh
    ok = O.OracleKM(cfg.charge, cfg.mass, [m])
    host = pyref.KM(cfg.charge, cfg.mass, [py_mesh(m)])  # the "unchanged Java": only its process_boundary / finish_slow are used
    absorbed = {"n": 0}

    def handler(km, slow, extra):
        keep = []
        absorbed["slow"] = absorbed.get("slow", 0) + slow.n
        host.n_absorbed, host.n_exited, host.hits = 0, 0, []
        for q in range(slow.n):
            part = pyref.Particle([slow.x[q], slow.y[q], slow.z[q]], [slow.u[q], slow.v[q], slow.w[q]], slow.mpw[q], int(slow.id[q]))
            part.lc, part.dt, part.born_it = [slow.li[q], slow.lj[q]], float(slow.dt[q]), int(slow.born_it[q])
            if host.finish_slow(int(extra["mesh"][q]), part, [extra["old_x"][q], extra["old_y"][q]], [extra["old_li"][q], extra["old_lj"][q]], int(extra["bounces"][q])):
                keep.append(part)
        absorbed["n"] = host.n_absorbed
        absorbed["exited"] = host.n_exited
        arr = dict(x=[p.pos[0] for p in keep], y=[p.pos[1] for p in keep], z=[p.pos[2] for p in keep], u=[p.vel[0] for p in keep], v=[p.vel[1] for p in keep],
                   w=[p.vel[2] for p in keep], mpw=[p.mpw for p in keep], li=[p.lc[0] for p in keep], lj=[p.lc[1] for p in keep], dt=[p.dt for p in keep])
        yield 0, Particles(len(keep), **{k: np.array(v, dtype=np.float64) for k, v in arr.items()}, id=np.array([p.id for p in keep], np.int32),
                           born_it=np.array([p.born_it for p in keep], np.int32))

    with KineticMaterial("O+", cfg.charge, cfg.mass, [m], m.domain_type, spwt=cfg.spwt, step_flags=flags, device_segments=False) as km:
        km.dt = cfg.dt
        km.slow_path_handler = handler
        state_o = state_g = O.java_seed(0)
        total_abs = 0
        for it in range(300):
            n_mp = cfg.num_mp()
            no, state_o = ok.sampleUniformSource(cfg.inlet, cfg.v_drift, n_mp, cfg.dt, state_o, cfg.spwt, born_it=it)
            ng, state_g = km.sampleUniformSource(cfg.inlet, cfg.v_drift, n_mp, state_g, born_it=it)
            assert (ng, state_g) == (no, state_o)
            absorbed["n"] = absorbed["exited"] = 0
            km.updateFields()
            ok.updateFields(cfg.dt)
            total_abs += absorbed["n"]
            assert absorbed["n"] == ok.n_absorbed and km.getNp() == ok.getNp(), it
            assert km.n_exited + absorbed["exited"] == ok.n_exited
        assert absorbed["slow"] > 1000 and total_abs > 100
        compare_state(km, ok)
        for f in range(7):
            scale = np.abs(ok.raw[0][f]).max()
            assert np.allclose(km.last_deposit[0][f], ok.raw[0][f], rtol=1e-10, atol=1e-10 * scale), f
        assert np.array_equal(km.last_deposit[0][7], ok.raw[0][7])


@pytest.mark.parametrize("flags", PATHS)
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_walls_device_segments_match_oracle(seed, flags):
    """Random absorbing / surviving / SINK polylines in XY, RZ and ZR domains with open, periodic and symmetry faces; particles that cross several
    cells and segments per step (tests/test_segments.py runs the same cases oracle against Python restatement)."""
    m, wl, arr = random_walls_case(seed)
    ok = O.OracleKM(wl.charge, wl.mass, [m])
    with KineticMaterial("ion", wl.charge, wl.mass, [m], m.domain_type, step_flags=flags) as km:
        km.dt = wl.dt
        assert km.addParticles(m, to_particles(arr), wl.dt) == ok.addParticles(0, arr, wl.dt)
        hits = 0
        for it in range(6):
            km.updateFields()
            ok.updateFields(wl.dt)
            h = km.takeSurfaceHits()
            assert km.n_slow == 0 and not ok.slow
            assert km.n_absorbed == ok.n_absorbed and km.n_exited == ok.n_exited and km.getNp() == ok.getNp(), it
            assert _hit_key(h) == _hit_key(ok.hits[0]), it
            hits += len(h["seg"])
            compare_state(km, ok)
            compare_fields(km, ok)
        assert hits > 30


@pytest.mark.parametrize("flags", PATHS)
@pytest.mark.parametrize("kind", [0, 1], ids=["absorb", "keep"])
def test_walls_next_to_a_mesh_handoff_match_oracle(kind, flags):
    """Walls on both sides of a MESH face: hits in the main pass and in the transfer sweeps of the neighbour mesh (KM:131-142), hit lists per mesh."""
    meshes, wl, arr = handoff_walls_case(kind)
    ok = O.OracleKM(wl.charge, wl.mass, meshes)
    with KineticMaterial("ion", wl.charge, wl.mass, meshes, meshes[0].domain_type, step_flags=flags) as km:
        km.dt = wl.dt
        assert km.addParticles(meshes[0], to_particles(arr), wl.dt) == ok.addParticles(0, arr, wl.dt)
        hits = [0, 0]
        for it in range(14):
            km.updateFields()
            ok.updateFields(wl.dt)
            h = km.takeSurfaceHits()
            assert km.n_slow == 0 and not ok.slow
            assert km.n_absorbed == ok.n_absorbed and km.n_exited == ok.n_exited and km.getNp() == ok.getNp(), it
            for k in range(2):
                sel = h["mesh"] == k
                assert _hit_key({key: v[sel] for key, v in h.items()}) == _hit_key(ok.hits[k]), (it, k)
                hits[k] += int(sel.sum())
            compare_state(km, ok)
            compare_fields(km, ok)
        assert hits[0] > 10 and hits[1] > 10
