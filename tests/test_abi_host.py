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
    want = int(re.search(r"#define SMK_ABI_VERSION (\d+)", open(HEADER).read()).group(1))
    assert lib.smk_version() == want == 7
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


@pytest.mark.parametrize("nsims,nsteps,piece", [(256, 20, 35), (5, 7, 3), (5, 7, 9), (4, 6, 1), (7, 4, 28), (300, 1, 3), (149, 100, 101), (1, 50, 7)])
def test_fused_time_sliced_plan(so_path, nsims, nsteps, piece):
    """smk_fused_plan (host only): the items cover every simulation-step once, an item that continues a simulation comes after
    the item that ends where it begins, and a machine that hands the items out in index order to `sms` workers as they free
    up -- what the GPU's CTA dispatch does -- finishes within the piece length plus the waits the plan itself contains."""
    import ctypes as C
    from smokephysai_b200 import _lib
    lib = _lib.load()
    cap = nsims * nsteps
    buf = (C.c_int32 * (4 * cap))()
    n = C.c_int32(0)
    assert lib.smk_fused_plan(nsims, nsteps, piece, buf, cap, C.byref(n)) == 0
    items = [tuple(buf[4 * k: 4 * k + 3]) for k in range(n.value)]
    seen = set()
    ends = {}
    for k, (b, t0, t1) in enumerate(items):
        assert 0 <= b < nsims and 0 <= t0 < t1 <= nsteps and t1 - t0 <= piece
        if t0 > 0:
            assert (b, t0) in ends and ends[(b, t0)] < k, "item %d continues a simulation nobody has advanced to step %d" % (k, t0)
        ends[(b, t1)] = k
        for t in range(t0, t1):
            assert (b, t) not in seen
            seen.add((b, t))
    assert len(seen) == nsims * nsteps
    npieces = -(-nsims * nsteps // piece)
    assert n.value <= nsims + npieces
    # in-order dispatch onto as many workers as there are pieces: event simulation in units of one step
    import heapq
    free = [(0, w) for w in range(npieces)]
    heapq.heapify(free)
    done_at = {}
    makespan = 0
    for b, t0, t1 in items:
        t_free, w = heapq.heappop(free)
        start = max(t_free, done_at.get((b, t0), 0))
        end = start + (t1 - t0)
        done_at[(b, t1)] = end
        makespan = max(makespan, end)
        heapq.heappush(free, (end, w))
    assert makespan <= max(piece, nsteps), "makespan %d steps for pieces of %d" % (makespan, piece)
    # too small a buffer is an error, not an overrun
    assert lib.smk_fused_plan(nsims, nsteps, piece, buf, 0, C.byref(n)) != 0 and n.value == len(items)
