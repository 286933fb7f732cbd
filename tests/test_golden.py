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


# ---------------------------------------------------------------- pin against the REAL reference (tools/java/ParityDump.java)
def parse_java_dump(path):
    """Output of tools/java/ParityDump.java: the raw bits of every double the real KineticMaterial produced."""
    unhex = lambda s: np.array([int(s, 16)], dtype=np.uint64).view(np.float64)[0]
    parts, fields, sums = [], {}, None
    for line in open(path):
        t = line.split()
        if not t:
            continue
        if t[0] == "P":
            parts.append([int(t[1])] + [unhex(s) for s in t[2:12]])
        elif t[0] == "F":
            fields[t[1]] = np.array([unhex(s) for s in t[2:]])
        elif t[0] == "sums":
            sums = np.array([unhex(s) for s in t[1:6]])
    parts.sort(key=lambda r: r[0])
    cols = list(zip(*parts)) if parts else [[]] * 11
    keys = ("id", "x", "y", "z", "u", "v", "w", "li", "lj", "dt", "mpw")
    got = {k: np.array(c, dtype=np.int32 if k == "id" else np.float64) for k, c in zip(keys, cols)}
    return got, fields, sums


JAVA_DUMPS = sorted(glob.glob(os.path.join(HERE, "golden", "java", "*.txt")))


def test_java_dump_parser_round_trip(tmp_path):
    """The text format ParityDump writes (hex of Double.doubleToRawLongBits) parses back bit for bit."""
    vals = np.array([0.0, -0.0, 1.5, -2.25e-300, np.pi, np.inf])
    hx = lambda a: " ".join("%x" % v for v in np.asarray(a, np.float64).view(np.uint64))
    p = tmp_path / "d.txt"
    p.write_text("java 11 np 1 steps 1\nsums " + hx(vals[:5]) + "\nP 7 " + hx(np.arange(10) * 0.1) + "\nF nd " + hx(vals) + "\n")
    got, fields, sums = parse_java_dump(str(p))
    assert got["id"][0] == 7 and np.array_equal(got["mpw"], [0.9]) and np.array_equal(got["x"], [0.0])
    assert np.array_equal(fields["nd"].view(np.uint64), vals.view(np.uint64)) and np.array_equal(sums, vals[:5])


@pytest.mark.parametrize("path", JAVA_DUMPS or [None], ids=[os.path.basename(p)[:-4] for p in JAVA_DUMPS] or ["absent"])
def test_oracle_matches_java_reference(path):
    """Bit-for-bit against the real KineticMaterial (run tools/java/run_parity_dump.sh on a machine with a JDK and commit
    tests/golden/java/*.txt).  Skipped while no dump is committed: parity stays "unpinned" (DESIGN.md section 2)."""
    if path is None:
        pytest.skip("no Java dumps committed: no JDK in the build image; see tools/java/run_parity_dump.sh")
    from oracle import oracle as O
    g, m, arr, steps = _case(os.path.join(HERE, "golden", os.path.basename(path)[:-4] + ".npz"))
    got, fields, sums = parse_java_dump(path)
    ok = O.OracleKM(float(g["charge"]), float(g["mass"]), [m])
    ok.addParticles(0, arr, float(g["dt"]))
    for _ in range(steps):
        ok.updateFields(float(g["dt"]))
    p = ok.sorted_parts(0)
    assert np.array_equal(p["id"], got["id"])
    for key in ("x", "y", "u", "v", "w", "li", "lj", "dt", "mpw"):
        assert np.array_equal(p[key], got[key]), key
    assert np.allclose(p["z"], got["z"], rtol=1e-12, atol=1e-300)  # StrictMath vs libm asin/acos
    for name in ("nd", "u", "v", "w", "count-sum", "u-sum", "uu-sum", "ww-sum", "mpc-sum"):
        want = fields[name].reshape(m.ni, m.nj)
        assert np.allclose(ok.fields[0][name], want, rtol=1e-10, atol=1e-10 * np.abs(want).max()), name
    assert np.allclose(np.array([ok.mass_sum, *ok.momentum_sum, ok.energy_sum]), sums, rtol=1e-10, atol=1e-10 * abs(sums[4]))
