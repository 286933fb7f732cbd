"""sf_div_exact (reciprocal multiply + two FMA residual corrections) must equal the IEEE quotient bit for bit:
tools/div_check.c compares it with a/b on random dividends and on neighbours of rounding midpoints."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reciprocal_division_is_exact(tmp_path):
    exe = str(tmp_path / "div_check")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-fopenmp", os.path.join(ROOT, "tools", "div_check.c"), "-lm", "-o", exe], check=True)
    out = subprocess.run([exe, "2000000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "0 mismatches" in out.stdout
