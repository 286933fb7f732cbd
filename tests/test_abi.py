"""The C-ABI library loads and exports every symbol include/sfgpu.h declares (no compute calls: no GPU needed),
the ctypes table matches the header, and the product never reaches into oracle/."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "sfgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sfgpu_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_functions():
    fns = header_functions()
    assert "sfgpu_step" in fns and "sfgpu_create" in fns and len(fns) >= 20


def test_library_exports_every_declared_symbol():
    from starfish_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build with __graft_entry__.build()"
    lib = C.CDLL(_lib.LIB_PATH)
    for name in header_functions():
        assert hasattr(lib, name), f"{name} declared in sfgpu.h but not exported"
    lib.sfgpu_abi_version.restype = C.c_int
    assert lib.sfgpu_abi_version() >= 1


def test_ctypes_table_matches_header():
    from starfish_b200 import _lib
    assert sorted(_lib.EXPORTS) == header_functions()
    _lib.load()


def test_create_without_device_fails_loudly():
    """No CPU fallback: on a box without CUDA sfgpu_create returns an error and a message."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from starfish_b200 import _lib
    lib = _lib.load()
    ctx = C.c_void_p()
    rc = lib.sfgpu_create(0, 0, C.byref(ctx))
    assert rc != 0 and not ctx.value
    assert b"no CPU fallback" in lib.sfgpu_last_error(None)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "starfish_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "sf_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f


def _build_c_harness(tmp_path):
    """gcc compiles tests/c/abi_harness.c against include/sfgpu.h (-Wall -Werror): a prototype that drifted from the
    implementation's argument list fails here or at link time, not in a ctypes table."""
    import shutil
    import subprocess
    from starfish_b200 import _lib
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    exe = str(tmp_path / "abi_harness")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_harness.c"),
                    "-o", exe, "-L", libdir, "-l:" + os.path.basename(_lib.LIB_PATH), "-lm", "-Wl,-rpath," + libdir], check=True)
    return exe


def test_c_harness_compiles_and_links(tmp_path):
    import subprocess
    exe = _build_c_harness(tmp_path)
    out = subprocess.run([exe, "--link"], check=True, capture_output=True, text=True).stdout
    assert "link OK" in out


@pytest.mark.gpu
def test_c_harness_runs_the_hot_path(tmp_path):
    """create -> mesh -> fields -> species -> inject -> 3 steps -> sums / deposit / moments / download from plain C."""
    import subprocess
    exe = _build_c_harness(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "abi_harness OK" in r.stdout


def test_jni_adapter_matches_the_header():
    """integration/jni/sfgpu_jni.c (the reference-side binding a maintainer builds against a real JDK) is type-checked against
    include/sfgpu.h with a stand-in jni.h: every sfgpu_* call in it has the header's argument list."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "tests", "c", "jni_stub"), "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "integration", "jni", "sfgpu_jni.c")], check=True)
    # and every native method of SfgpuJni.java has its C function
    java = open(os.path.join(ROOT, "integration", "java", "starfish", "core", "materials", "SfgpuJni.java")).read()
    csrc = open(os.path.join(ROOT, "integration", "jni", "sfgpu_jni.c")).read()
    natives = re.findall(r"static native [\w\[\]<>]+ (\w+)\(", java)
    assert len(natives) >= 20
    for name in natives:
        assert re.search(r"FN\(%s\)\(" % name, csrc), f"SfgpuJni.{name} has no JNI function"
