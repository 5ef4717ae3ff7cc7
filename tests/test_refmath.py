"""cl-rrt_b200/csrc/refmath.cuh (the float libm restatement the kernels use) against the C library, bit for bit.
The header is compiled for the host (REFMATH_HOST); the device build uses the same code with IEEE fma / div / sqrt.
The sweep with stride 1 (every float of each domain, ~2.5 min) was run when the header was written: 0 mismatches
of 2.2e9 (sinf, cosf), 2.1e9 (acosf, asinf), 4.3e9 (atanf), 4e8 (atan2f); here a strided sweep keeps it fast."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_refmath_matches_glibc(tmp_path):
    exe = str(tmp_path / "refmath_sweep")
    subprocess.check_call(["g++", "-O2", "-mfma", "-ffp-contract=off", "-DREFMATH_HOST", "-o", exe,
                           os.path.join(ROOT, "tests", "native", "refmath_sweep.cpp"), "-lm"])
    out = subprocess.run([exe, "257"], capture_output=True, text=True)
    print(out.stdout)
    assert out.returncode == 0, out.stdout
    for name in ("sinf", "cosf", "acosf", "asinf", "atanf", "atan2f"):
        assert f"{name}: 0 mismatches" in out.stdout


def test_refmath64_matches_glibc(tmp_path):
    """cl-rrt_b200/csrc/refmath64.cuh (double sin / cos / tan / exp of the rollout kernels) against the C library, bit for bit:
    2 M random arguments per range here; 1.05e9 (sin, cos) and 7.5e8 (tan) when the header was written: 0 mismatches."""
    exe = str(tmp_path / "refmath64_sweep")
    subprocess.check_call(["g++", "-O2", "-mfma", "-ffp-contract=off", "-DREFMATH_HOST", "-o", exe,
                           os.path.join(ROOT, "tests", "native", "refmath64_sweep.cpp"), "-lm"])
    out = subprocess.run([exe, "2"], capture_output=True, text=True)
    print(out.stdout)
    assert out.returncode == 0, out.stdout
    for name in ("sin", "cos", "sincos", "tan", "special", "exp"):
        assert f"{name}: 0 mismatches" in out.stdout


def test_refmath64_tables_are_the_installed_libms():
    """The lookup tables compiled into the kernels (csrc/refmath64_tables.inc) are those of this host's glibc: the generator
    re-reads them from libm.so.6 and compares.  (On another glibc build the generator's own address checks fail first.)"""
    import sys
    if not os.path.exists("/lib/x86_64-linux-gnu/libm.so.6"):
        import pytest
        pytest.skip("no x86-64 glibc libm at the expected path")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gen_refmath64_tables.py"), "--check"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
