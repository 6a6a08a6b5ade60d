"""The table-driven FP64 log / exp of the fused seeds kernel (mamba.jl_b200/csrc/fasttab_fn.cuh + fasttab.cuh): the committed tables are what the
generator produces, and a host build of the very same header stays within 1.5 ulp (log) / 1 ulp (exp) of long double."""
import os
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mamba.jl_b200", "csrc")


def test_tables_match_generator(tmp_path, monkeypatch):
    pytest.importorskip("mpmath")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import gen_fastmath_tables as gen
    finally:
        sys.path.pop(0)
    out = tmp_path / "fasttab.cuh"
    monkeypatch.setattr(gen, "OUT", str(out))
    gen.main()
    assert out.read_text() == open(os.path.join(CSRC, "fasttab.cuh")).read()


def test_host_build_accuracy(tmp_path):
    exe = tmp_path / "check_fasttab"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-I", CSRC,
                           os.path.join(ROOT, "tools", "check_fasttab.cpp"), "-o", str(exe)])
    res = subprocess.run([str(exe), "1500000"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "tlog: max error" in res.stdout and "texp: max error" in res.stdout
