"""The reference's OWN test suites, unmodified, against this repo's reference-named flavour
libraries (SURVEY.md section 8(f)-1's definition of done).

baseline/fetch_ref.sh copies the reference's ctypes wrappers and tests into the git-ignored
baseline/_ref/ (it ships to the GPU box).  Each test assembles a scratch directory the way a
user of the reference would -- the wrapper package with the shared library next to it
(scalar/examples/python_ctypes/src/pyopengjk/opengjk.py:25-41 and
gpu/examples/python/src/pyopengjk_gpu/opengjk_gpu.py:76-116 look there) -- and launches the
reference's own runner in a subprocess:
  * scalar: `pytest test_min_distance.py` through libopengjk_ce.so (each call = a batch of one
    on the GPU);
  * batched: `python test_examples.py` (9 scenarios) through libopenGJK_GPU.so.
On a CPU-only box the scalar suite runs against the reference's own library (oracle/_ref),
which pins the expected pass count."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
PKG = os.path.join(ROOT, "opengjk_b200")
SCALAR_TESTS = 54  # scalar/examples/python_ctypes/test/test_min_distance.py

needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "scalar", "test", "test_min_distance.py")),
                               reason="baseline/_ref not fetched (run baseline/fetch_ref.sh with the reference mounted)")


def _scalar_tree(tmp_path, libs):
    """<tmp>/pyopengjk/{__init__,opengjk}.py + the library files; <tmp>/test_min_distance.py"""
    pkg = tmp_path / "pyopengjk"
    shutil.copytree(os.path.join(REF, "scalar", "pyopengjk"), pkg)
    for src, name in libs:
        shutil.copy(src, pkg / name)
    shutil.copy(os.path.join(REF, "scalar", "test", "test_min_distance.py"), tmp_path / "test_min_distance.py")
    return tmp_path


def _run(cmd, cwd, extra_path):
    env = dict(os.environ)
    env["PYTHONPATH"] = str(extra_path) + os.pathsep + env.get("PYTHONPATH", "")
    return subprocess.run(cmd, cwd=cwd, env=env, capture_output=True, text=True, timeout=1500)


@needs_ref
def test_reference_scalar_suite_against_the_reference_library(tmp_path):
    """CPU: the harness itself, with the unmodified reference compiled into oracle/_ref standing
    where libopengjk_ce.so goes.  Pins the pass count the GPU test must reproduce."""
    ref_lib = os.path.join(ROOT, "oracle", "_ref", "libopengjk_ref_f64.so")
    if not os.path.exists(ref_lib):
        pytest.skip("oracle/_ref not built")
    t = _scalar_tree(tmp_path, [(ref_lib, "libopengjk_ce.so")])
    r = _run([sys.executable, "-m", "pytest", "test_min_distance.py", "-q", "-p", "no:cacheprovider"], t, t)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert re.search(rf"\b{SCALAR_TESTS} passed\b", r.stdout), r.stdout[-1000:]


@needs_ref
@pytest.mark.gpu
def test_reference_scalar_suite_unmodified(tmp_path):
    t = _scalar_tree(tmp_path, [(os.path.join(PKG, "libopengjk_ce.so"), "libopengjk_ce.so"),
                                (os.path.join(PKG, "libopengjk_b200.so"), "libopengjk_b200.so")])
    r = _run([sys.executable, "-m", "pytest", "test_min_distance.py", "-q", "-p", "no:cacheprovider"], t, t)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert re.search(rf"\b{SCALAR_TESTS} passed\b", r.stdout), r.stdout[-1000:]
    assert "failed" not in r.stdout.splitlines()[-1]


@needs_ref
@pytest.mark.gpu
def test_reference_gpu_examples_unmodified(tmp_path):
    pkg = tmp_path / "pyopengjk_gpu"
    shutil.copytree(os.path.join(REF, "gpu", "pyopengjk_gpu"), pkg)
    shutil.copy(os.path.join(PKG, "libopenGJK_GPU.so"), pkg / "libopenGJK_GPU.so")   # fp32: the wrapper's USE_32BITS default
    shutil.copy(os.path.join(PKG, "libopengjk_b200.so"), pkg / "libopengjk_b200.so")
    shutil.copy(os.path.join(REF, "gpu", "test_examples.py"), tmp_path / "test_examples.py")
    r = _run([sys.executable, "test_examples.py"], tmp_path, tmp_path)
    out = re.sub(r"\x1b\[[0-9;]*m", "", r.stdout)
    assert r.returncode == 0, out[-3000:] + r.stderr[-2000:]
    assert "All tests completed successfully!" in out, out[-3000:] + r.stderr[-2000:]
    fails = [ln for ln in out.splitlines() if "FAIL:" in ln]
    assert not fails, fails
    assert "Error:" not in out and "Traceback" not in r.stderr, out[-2000:] + r.stderr[-2000:]
    passes = [ln for ln in out.splitlines() if "PASS:" in ln]
    assert len(passes) >= 9, out[-3000:]     # the 9 scenarios of test_examples.py:135-662
