"""Golden fixtures (tests/golden/*.npz, produced by the independent Python restatement of the Java source, see
tests/golden/make_golden.py): the C oracle must reproduce them on CPU, the CUDA path on the GPU."""
import glob
import os

import numpy as np
import pytest

from starfish_b200 import synthetic as S
from starfish_b200.domain import DomainType

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def _case(path):
    g = np.load(path)
    dom, ni, nj, n, steps = [int(v) for v in g["meta"]]
    m = S.make_mesh(ni, nj, DomainType(dom), 1e-3, str(g["bc"]))
    m.efi, m.efj = g["efi"], g["efj"]
    arr = {k[3:]: g[k] for k in g.files if k.startswith("in_")}
    return g, m, arr, steps


def _check(g, got, raw, sums, n_exited, theta_exact):
    assert np.array_equal(got["id"], g["id"])
    for key, ref in (("x", "x"), ("y", "y"), ("u", "u"), ("v", "v"), ("w", "w"), ("li", "li"), ("lj", "lj"), ("dt", "dtp")):
        assert np.array_equal(got[key], g[ref]), key
    if theta_exact:
        assert np.array_equal(got["z"], g["z"])
    else:
        assert np.allclose(got["z"], g["z"], rtol=1e-12, atol=1e-300)
    scale = np.abs(g["raw"]).max(axis=(1, 2), keepdims=True)
    assert np.all(np.abs(raw - g["raw"]) <= 1e-10 * scale)
    assert np.array_equal(raw[7], g["raw"][7])
    assert np.allclose(sums, g["sums"], rtol=1e-10, atol=1e-10 * abs(g["sums"][4]))
    assert n_exited == int(g["n_exited"])


def test_fixtures_exist():
    assert len(FIXTURES) >= 5


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_oracle_reproduces_golden(path):
    from oracle import oracle as O
    g, m, arr, steps = _case(path)
    ok = O.OracleKM(float(g["charge"]), float(g["mass"]), [m])
    ok.addParticles(0, arr, float(g["dt"]))
    for _ in range(steps):
        ok.updateFields(float(g["dt"]))
    p = ok.sorted_parts(0)
    # the oracle is plain C on the same libm: theta (asin/acos) is bit identical too
    _check(g, p, ok.raw[0], ok.sums5, ok.n_exited, theta_exact=True)


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_cuda_reproduces_golden(path):
    from starfish_b200 import KineticMaterial, Particles
    g, m, arr, steps = _case(path)
    with KineticMaterial("ion", float(g["charge"]), float(g["mass"]), [m], m.domain_type) as km:
        km.dt = float(g["dt"])
        km.addParticles(m, Particles(len(arr["x"]), **arr), km.dt)
        for _ in range(steps):
            km.updateFields()
        p = km.getParticles(m).sorted_by_id()
        got = {k: getattr(p, k) for k in ("id", "x", "y", "z", "u", "v", "w", "li", "lj", "dt")}
        sums = np.array([km.mass_sum, *km.momentum_sum, km.energy_sum]) / km.mass
        _check(g, got, km.last_deposit[0], sums, km.n_exited, theta_exact=m.domain_type == DomainType.XY)


# ---------------------------------------------------------------- UniformSource / ColdBeamSource fixtures
SRC_FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "source", "*.npz")))


def _src_case(path):
    from starfish_b200.domain import LinearSpline
    g = np.load(path)
    cold, seed, num_mp, steps = [int(v) for v in g["meta"]]
    m = S.make_mesh(21, 11, DomainType.XY, 5e-3, "open", x0=(-0.1, 0.0))
    m.efi, m.efj = g["efi"], g["efj"]
    return g, m, LinearSpline(g["pts"]), bool(cold), seed, num_mp, steps


def _src_check(g, got, raw, n_exited, state):
    assert state == int(g["rng_state"])
    assert np.array_equal(got["id"], g["id"]) and np.array_equal(got["born_it"], g["born_it"])
    for key in ("x", "y", "z", "u", "v", "w"):
        assert np.array_equal(got[key], g[key]) and np.array_equal(np.signbit(got[key]), np.signbit(g[key])), key
    scale = np.abs(g["raw"]).max(axis=(1, 2), keepdims=True)
    assert np.all(np.abs(raw - g["raw"]) <= 1e-10 * scale) and np.array_equal(raw[7], g["raw"][7])
    assert n_exited == int(g["n_exited"])


def test_source_fixtures_exist():
    assert len(SRC_FIXTURES) >= 2


@pytest.mark.parametrize("path", SRC_FIXTURES, ids=[os.path.basename(p)[:-4] for p in SRC_FIXTURES])
def test_oracle_reproduces_source_golden(path):
    from oracle import oracle as O
    g, m, spline, cold, seed, num_mp, steps = _src_case(path)
    ok = O.OracleKM(float(g["charge"]), float(g["mass"]), [m])
    state = O.java_seed(seed)
    for it in range(steps):
        n, state = ok.sampleUniformSource(spline, float(g["v_drift"]), num_mp, float(g["dt"]), state, 1e3, born_it=it, cold_beam=cold)
        assert n == int(g["added"][it])
        ok.updateFields(float(g["dt"]))
    _src_check(g, ok.sorted_parts(0), ok.raw[0], ok.n_exited, state)


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [0, 8], ids=["default", "stream"])
@pytest.mark.parametrize("path", SRC_FIXTURES, ids=[os.path.basename(p)[:-4] for p in SRC_FIXTURES])
def test_cuda_reproduces_source_golden(path, flags):
    from starfish_b200 import KineticMaterial
    g, m, spline, cold, seed, num_mp, steps = _src_case(path)
    with KineticMaterial("ion", float(g["charge"]), float(g["mass"]), [m], m.domain_type, step_flags=flags) as km:
        km.dt = float(g["dt"])
        state = (seed ^ 0x5DEECE66D) & ((1 << 48) - 1)  # java.util.Random.setSeed scrambling
        for it in range(steps):
            n, state = km.sampleUniformSource(spline, float(g["v_drift"]), num_mp, state, mpw=1e3, born_it=it, cold_beam=cold)
            assert n == int(g["added"][it])
            km.updateFields()
        p = km.getParticles(m).sorted_by_id()
        got = {k: getattr(p, k) for k in ("id", "born_it", "x", "y", "z", "u", "v", "w")}
        _src_check(g, got, km.last_deposit[0], km.n_exited, state)
