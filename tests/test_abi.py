"""CPU: the C-ABI library builds, loads and exports every symbol include/csvb200.h declares
(no compute calls -- there is no GPU here and no CPU fallback to call instead)."""
import ctypes as C
import os
import re

import pytest

from csv_simd_b200 import _lib, build as cbuild
from tests.conftest import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "csvb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(csvb200_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_for_sm100a():
    so = cbuild.build()
    assert os.path.exists(so)
    flags = " ".join(cbuild.NVCC_FLAGS)
    assert "arch=compute_100a,code=sm_100a" in flags and "-lineinfo" in flags


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in csvb200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_version_and_status_strings():
    lib = _lib.load()
    assert lib.csvb200_version() == 100
    assert lib.csvb200_status_string(3).decode() == "Unsupported csv structure: likely variable number of fields"
    assert lib.csvb200_status_string(2).decode() == "Invalid state"
    assert lib.csvb200_status_string(4).decode() == "Missing a value"


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.csvb200_ctx_create(0, C.byref(h)) == 6  # CSVB200_ERR_CUDA
    assert not h.value
    import csv_simd_b200 as cs
    with pytest.raises(cs.GpuError):
        cs.Context(0)
    with pytest.raises(cs.GpuError):
        cs.reader.read(b"a,b\n" * 32)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "csv_simd_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower().replace("no cpu fallback", ""), f"{f} mentions the oracle"


def test_header_is_plain_c_and_links():
    """include/csvb200.h is the drop-in boundary: it must compile as C99 (not just C++), and a C program that
    references every declared function must link against libcsvb200.so (no compute call is made)."""
    import subprocess
    import tempfile
    so = cbuild.build()
    names = declared_symbols()
    src = '#include "csvb200.h"\n#include <stdio.h>\nint main(void) {\n    const void* fns[] = {\n'
    src += "".join(f"        (const void*)&{n},\n" for n in names)
    src += ('    };\n    printf("%d %s %u\\n", csvb200_version(), csvb200_status_string(CSVB200_ERR_INVALID_CSV_FORMAT),'
            ' (unsigned)(sizeof(fns) / sizeof(fns[0])));\n    return 0;\n}\n')
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-Wno-pedantic",
                               "-I", os.path.join(ROOT, "include"), c, "-o", exe, "-L", os.path.dirname(so), "-lcsvb200",
                               "-Wl,-rpath," + os.path.dirname(so)])
        out = subprocess.run([exe], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        v, *msg, cnt = out.stdout.split()
        assert int(v) == 100 and int(cnt) == len(names) and "Unsupported" in out.stdout


def test_rust_sys_crate_is_in_sync_with_the_header():
    """rust/csv-simd-b200-sys/src/lib.rs is generated from include/csvb200.h (there is no rustc in the image, so the
    check is textual): the committed file is what the generator emits now, and it declares every exported symbol."""
    import subprocess
    import sys as _sys
    subprocess.check_call([_sys.executable, os.path.join(ROOT, "tools", "gen_rust_sys.py"), "--check"])
    text = open(os.path.join(ROOT, "rust", "csv-simd-b200-sys", "src", "lib.rs")).read()
    for name in declared_symbols():
        assert f"pub fn {name}(" in text, name
    safe = "".join(open(os.path.join(ROOT, "rust", "csv-simd-b200", "src", f)).read()
                   for f in ("lib.rs", "gpu.rs", "reader.rs", "record_source.rs"))
    import re
    for used in set(re.findall(r"sys::(csvb200_\w+)\(", safe)):
        assert f"pub fn {used}(" in text, used
