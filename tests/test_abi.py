"""CPU tests of the boundary: libmambacuda.so builds for sm_100a, loads, exports every symbol include/mambacuda.h
declares, fails loudly without a GPU, and the product never touches the oracle."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mambacuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mcu_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for s in ["mcu_create", "mcu_set_data", "mcu_set_scheme", "mcu_set_inits", "mcu_run", "mcu_logpdf", "mcu_gradlogpdf",
              "mcu_get_state", "mcu_set_state", "mcu_moments", "mcu_gelman", "mcu_summarystats", "mcu_set_rng_mode", "mcu_last_error", "mcu_destroy"]:
        assert s in syms


def test_library_exports_every_declared_symbol(mcu_built):
    L = ctypes.CDLL(mcu_built)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, missing
    from mambacuda import _lib
    assert sorted(_lib.SYMBOLS) == declared_symbols()
    assert L.mcu_abi_version() == 1


def test_library_contains_sm100a_code(mcu_built):
    out = subprocess.run(["cuobjdump", "-lelf", mcu_built], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_block_desc_layout_matches_the_header():
    from mambacuda._lib import BlockDesc
    # int32 x 2, int32[8], int32 x 8, (pad), double x 4, pointer  — as laid out by a C compiler for mcu_block_desc
    assert ctypes.sizeof(BlockDesc) == 4 * 2 + 4 * 8 + 4 * 8 + 8 * 4 + 8
    assert BlockDesc.target.offset == 72 and BlockDesc.scale.offset == 104
    src = f'#include <stddef.h>\n#include <stdio.h>\n#include "{ROOT}/include/mambacuda.h"\n' \
          'int main(){printf("%zu %zu %zu", sizeof(mcu_block_desc), offsetof(mcu_block_desc, target), offsetof(mcu_block_desc, scale));return 0;}'
    exe = os.path.join(ROOT, "mamba.jl_b200", "build", "abi_probe")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["gcc", "-x", "c", "-", "-o", exe], input=src, text=True, check=True)
    assert subprocess.run([exe], capture_output=True, text=True).stdout.split() == ["112", "72", "104"]


def test_no_cpu_fallback_without_a_device(mcu_built):
    from mambacuda import _lib
    from mambacuda.engine import Engine, MambaCudaError
    L = _lib.lib()
    if L.mcu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(MambaCudaError, match="no usable CUDA device"):
        Engine("seeds", 4)


def test_argument_validation_does_not_need_a_device(mcu_built):
    from mambacuda import _lib
    L = _lib.lib()
    h = ctypes.c_void_p()
    assert L.mcu_create(99, 4, 0, 0, 1, ctypes.byref(h)) == _lib.ERR_ARG
    assert b"unknown template" in L.mcu_last_error(None)
    assert L.mcu_create(1, 0, 0, 0, 1, ctypes.byref(h)) == _lib.ERR_ARG
    assert L.mcu_kept(0, 2000, 1000, 10) == 100 and L.mcu_kept(0, 50, 11, 4) == 9 and L.mcu_kept(130, 170, 120, 3) == 57


def test_product_never_touches_the_oracle():
    # the oracle is test infrastructure: nothing under mamba.jl_b200/ or include/ may import, link or name it
    bad = []
    for base in ("mamba.jl_b200", "include"):
        for root, _, files in os.walk(os.path.join(ROOT, base)):
            if os.sep + "build" in root:
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".jl")):
                    txt = open(os.path.join(root, f), errors="ignore").read()
                    if re.search(r"pyoracle|liboracle|oracle/|orc_[a-z]", txt):
                        bad.append(os.path.join(root, f))
    assert not bad, bad


def test_block_descriptor_layout_is_the_same_in_all_three_bindings():
    """mcu_block_desc (include/mambacuda.h), the ctypes mirror (mambacuda/_lib.py, pyoracle.py) and the Julia `immutable BlockDesc`
    (julia/MambaCUDA.jl) must list the same fields in the same order with the same widths: the struct crosses the ABI by pointer."""
    import ctypes as C
    import re
    sys.path.insert(0, os.path.join(ROOT, "mamba.jl_b200"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from mambacuda import _lib
    import pyoracle
    h = open(os.path.join(ROOT, "include", "mambacuda.h")).read()
    body = re.search(r"typedef struct mcu_block_desc \{(.*?)\} mcu_block_desc;", h, re.S).group(1)
    fields = []
    for line in body.splitlines():
        m = re.match(r"\s*(const double\*|int32_t|double)\s+(\w+)(\[\w+\])?;", line)
        if m:
            fields.append((m.group(2), m.group(1), bool(m.group(3))))
    assert [f[0] for f in fields] == ["kind", "n_nodes", "nodes", "transform", "adapt", "batchsize", "proposal", "L", "grad", "max_depth", "n_scale",
                                      "target", "epsilon", "beta", "amm_scale", "scale"]
    width = {"int32_t": 4, "double": 8, "const double*": 8}
    for cls in (_lib.BlockDesc, pyoracle.BlockDesc):
        assert [n for n, _ in cls._fields_] == [f[0] for f in fields]
        for (n, t), (_, ctype, is_arr) in zip(cls._fields_, fields):
            assert C.sizeof(t) == width[ctype] * (8 if is_arr else 1), n
        assert C.sizeof(cls) == 112                      # 11 x int32 + 8 x int32 nodes = 76 → padded to 80, + 4 doubles + 1 pointer
    jl = open(os.path.join(ROOT, "mamba.jl_b200", "julia", "MambaCUDA.jl")).read()
    jbody = re.search(r"immutable BlockDesc\n(.*?)\nend", jl, re.S).group(1)
    jfields = [tuple(x.strip() for x in ln.split("::")) for ln in jbody.splitlines() if "::" in ln]
    assert [f[0] for f in jfields] == [f[0] for f in fields]
    jt = {"int32_t": "Int32", "double": "Float64", "const double*": "Ptr{Float64}"}
    for (n, t), (_, ctype, is_arr) in zip(jfields, fields):
        assert t == ("NTuple{8, Int32}" if is_arr else jt[ctype]), (n, t)
    assert "MCU_MAX_BLOCK_NODES = 8" in jl and re.search(r"#define MCU_MAX_BLOCK_NODES\s+8", h)


def _split_top(s):
    """split on commas that are not inside (), {} or []"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def test_julia_and_ctypes_bindings_pass_as_many_arguments_as_the_header_declares():
    """Every ccall of the Julia shim and every argtypes list of the ctypes stub has the arity of the prototype in include/mambacuda.h
    (a call with a missing trailing pointer would still link and then read garbage)."""
    import re
    hdr = re.sub(r"/\*.*?\*/", " ", open(os.path.join(ROOT, "include", "mambacuda.h")).read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|double|int64_t|const char\*)\s+(mcu_\w+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len(_split_top(args))
    assert len(protos) > 50
    jl = open(os.path.join(ROOT, "mamba.jl_b200", "julia", "MambaCUDA.jl")).read()
    seen = 0
    for m in re.finditer(r"ccall\(\(:(mcu_\w+),\s*libmambacuda\),\s*[\w{}]+,\s*\(", jl):
        name, i = m.group(1), m.end()
        depth, j = 1, i
        while depth:
            depth += {"(": 1, ")": -1}.get(jl[j], 0); j += 1
        types = _split_top(jl[i:j - 1])
        assert name in protos, name
        assert len(types) == protos[name], f"{name}: Julia passes {len(types)} arguments, the header declares {protos[name]}"
        seen += 1
    assert seen >= 25
    sys.path.insert(0, os.path.join(ROOT, "mamba.jl_b200"))
    from mambacuda import _lib
    src = open(_lib.__file__).read()
    for m in re.finditer(r"L\.(mcu_\w+)\.argtypes\s*=\s*\[([^\]]*)\]", src):
        name, n = m.group(1), len(_split_top(m.group(2)))
        assert name in protos, name
        assert n == protos[name], f"{name}: ctypes declares {n} arguments, the header {protos[name]}"


def test_measurement_helpers_compile():
    """tools/*.py (variant timing, ncu drivers, table generator) byte-compile: they are run by hand on the GPU box, so nothing else would notice a typo."""
    import glob
    import py_compile
    files = sorted(glob.glob(os.path.join(ROOT, "tools", "*.py")))
    assert len(files) >= 8
    for f in files:
        py_compile.compile(f, doraise=True)
