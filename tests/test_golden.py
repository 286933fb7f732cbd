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
