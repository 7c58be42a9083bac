"""CPU-only checks of the boundary: the shared library loads, exports every symbol include/smoke_b200.h
declares, the ctypes table matches the header, and the host-side layout arithmetic is right.
No compute calls are made here (there is no GPU in the build container)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "smoke_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"SMK_API\s+([\w\s\*]+?)\s*\b(smk_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = [a.strip() for a in m.group(3).split(",")]
        out[m.group(2)] = 0 if args == ["void"] else len(args)
    return out


@pytest.fixture(scope="module")
def so_path():
    from smokephysai_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(so_path):
    decl = header_functions()
    assert len(decl) >= 15
    syms = subprocess.check_output(["nm", "-D", "--defined-only", so_path]).decode()
    exported = set(re.findall(r" T (smk_\w+)", syms))
    assert set(decl) == exported, "header vs exports: %s" % (set(decl) ^ exported)


def test_ctypes_table_matches_header(so_path):
    from smokephysai_b200 import _lib
    decl = header_functions()
    table = dict(_lib.SIGNATURES)
    table["smk_last_error_string"] = []
    assert set(table) == set(decl)
    for name, args in table.items():
        assert len(args) == decl[name], "%s: %d ctypes args vs %d in the header" % (name, len(args), decl[name])
    lib = _lib.load()
    assert lib.smk_version() == 4
    assert isinstance(lib.smk_last_error_string(), bytes)


def test_struct_sizes_match_c(so_path, tmp_path):
    """sizeof/offsetof of the ABI structs as gcc sees them == the ctypes mirrors."""
    from smokephysai_b200 import _lib
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(){printf("%%zu %%zu %%zu %%zu %%zu %%zu\\n",'
                    'sizeof(smk_grid_t),sizeof(smk_source_t),sizeof(smk_state_t),sizeof(smk_params_t),'
                    'offsetof(smk_grid_t,stride_u),offsetof(smk_state_t,cur_u));return 0;}\n' % HEADER)
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-o", str(exe), str(prog)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(_lib.Grid), C.sizeof(_lib.Source), C.sizeof(_lib.State), C.sizeof(_lib.Params),
            _lib.Grid.stride_u.offset, _lib.State.cur_u.offset]
    assert got == want


def test_field_layout():
    from smokephysai_b200 import FieldLayout
    L = FieldLayout(128, 128, 4)
    assert (L.pitch_u, L.pitch_v, L.pitch_c) == (128, 132, 128)
    assert L.stride_u == 129 * 128 and L.stride_v == 128 * 132 and L.stride_c == 128 * 128
    offs = [L.offset[n] for n in L.FIELDS]
    assert offs == sorted(offs) and all(o % 4 == 0 for o in offs)
    assert L.total == 4 * (2 * L.stride_u + 2 * L.stride_v + 6 * L.stride_c)
    R = FieldLayout(17, 29)
    assert (R.pitch_u, R.pitch_v, R.pitch_c) == (32, 32, 32)
    assert R.shape_of("v1") == (17, 30, 32) and R.shape_of("u0") == (18, 29, 32) and R.shape_of("div") == (17, 29, 32)
    with pytest.raises(ValueError):
        FieldLayout(0, 4)


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from smokephysai_b200 import SmokeSimulator
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        SmokeSimulator((32, 32), device="cuda")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        SmokeSimulator((32, 32), device="cpu")


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under smokephysai_b200/ or src/ may reference it."""
    bad = []
    for base in ("smokephysai_b200", "src"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".py", ".cu", ".cuh", ".h")):
                    txt = open(os.path.join(dp, fn)).read()
                    if re.search(r"\boracle\b", txt):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_no_packed_fma_in_the_library():
    """The arithmetic contract is 'every fp32 operation rounded once' (DESIGN.md s2).  ptxas 12.9 contracts
    mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under --fmad=false; the kernels are written so that no packed add
    consumes a packed product, and this test keeps it that way: no FFMA2 may appear anywhere in the SASS."""
    import shutil
    import subprocess
    from smokephysai_b200 import _lib
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([tool, "-sass", _lib.SO_PATH], capture_output=True, text=True, check=True).stdout
    assert "FADD2" in sass and "FMUL2" in sass, "expected the f32x2 kernels in the library"
    assert "FFMA2" not in sass, "a packed multiply-add was contracted into FFMA2: results would not be rounded-once"
