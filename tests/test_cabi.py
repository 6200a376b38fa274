"""The C-ABI library loads and exports every symbol include/vpint_b200.h declares.
No compute call is made here (no GPU in the build container); the host-only
entry points (create / workspace sizing / argument validation) are exercised."""

import ctypes as C
import os
import re

import pytest

from vpint_b200 import _native as N
from vpint_b200.errors import VPintError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vpint_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vpint_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = N.load_library()
    syms = declared_symbols()
    assert len(syms) >= 20
    for name in syms:
        assert hasattr(lib, name), "missing export: " + name
    assert sorted(N.SIGNATURES) == syms, "ctypes prototypes and header disagree"
    assert lib.vpint_abi_version() == N.ABI_VERSION


def test_row_pitch():
    lib = N.load_library()
    assert lib.vpint_row_pitch(1) == 64
    assert lib.vpint_row_pitch(64) == 64
    assert lib.vpint_row_pitch(65) == 128
    assert lib.vpint_row_pitch(10980) == 11008


def make_desc(**kw):
    d = N.Desc()
    d.ndim, d.nprob, d.H, d.W, d.T = 2, 1, 16, 16, 1
    d.storage, d.weight_mode = N.F64, N.W_PLANES
    for k, v in kw.items():
        setattr(d, k, v)
    return d


def test_create_and_workspace_size_host_only():
    lib = N.load_library()
    h = C.c_void_p()
    assert lib.vpint_solver_create(C.byref(make_desc(nprob=13, H=256, W=256)), C.byref(h)) == N.OK
    nbytes = lib.vpint_solver_workspace_bytes(h)
    # 2 state buffers + 4 weight planes of 13 x 256 x 256 fp64, plus bookkeeping
    dense = 13 * 256 * 256 * 8 * 6
    assert dense < nbytes < dense * 1.2
    assert lib.vpint_solver_total_tiles(h) == 13 * 2 * 8
    assert lib.vpint_solver_k_block(h) >= 1
    lib.vpint_solver_destroy(h)
    # fp32 storage halves the big planes
    assert lib.vpint_solver_create(C.byref(make_desc(nprob=13, H=256, W=256, storage=N.F32)), C.byref(h)) == N.OK
    assert lib.vpint_solver_workspace_bytes(h) < 0.6 * nbytes
    lib.vpint_solver_destroy(h)


@pytest.mark.parametrize("bad", [dict(ndim=4), dict(nprob=0), dict(T=3), dict(storage=7), dict(weight_mode=5),
                                 dict(k_block=99), dict(nprob=6, nprob_inner=4),
                                 dict(global_H=8, row_offset=0, own_row0=0, own_rows=4)])
def test_create_rejects_bad_descriptors(bad):
    lib = N.load_library()
    h = C.c_void_p()
    rc = lib.vpint_solver_create(C.byref(make_desc(**bad)), C.byref(h))
    assert rc in (N.ERR_INVALID, N.ERR_UNSUPPORTED)
    assert lib.vpint_last_error()


def test_unbound_solver_fails_loudly():
    lib = N.load_library()
    h = C.c_void_p()
    assert lib.vpint_solver_create(C.byref(make_desc()), C.byref(h)) == N.OK
    prm = N.RunParams()
    prm.max_iter = 3
    assert lib.vpint_solver_run(h, C.byref(prm), None) == N.ERR_WORKSPACE
    with pytest.raises(VPintError):
        N.check(lib, lib.vpint_solver_step_begin(h, None), "step")
    lib.vpint_solver_destroy(h)


def test_missing_library_is_an_error(tmp_path):
    with pytest.raises(VPintError):
        N.load_library(str(tmp_path / "nope.so"))


def test_built_library_is_sm100a_and_uses_the_blackwell_paths():
    """The shipped .so holds sm_100a code only, and the kernels that claim DSMEM st.async + mbarriers (resident), the
    copy engine (scene stencil) and TMA tensor loads (temporally blocked kernel) really contain those instructions."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    lib = os.path.join(ROOT, "vpint_b200", "libvpint_b200.so")
    elfs = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True, check=True).stdout
    archs = set(re.findall(r"\.(sm_\w+)\.cubin", elfs))
    assert archs == {"sm_100a"}, archs
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    per_kernel, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = per_kernel.setdefault(m.group(1), set())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur.add(m.group(1))

    def ops(fragment):
        found = [v for k, v in per_kernel.items() if fragment in k]
        assert found, "kernel not in the library: " + fragment
        return set.union(*found)

    assert {"STAS", "SYNCS"} <= ops("resident2d_kernel")            # st.async into distributed shared memory, mbarriers
    assert {"UBLKCP", "SYNCS"} <= ops("jacobi2d_ratio_bulk_kernel")  # cp.async.bulk + mbarrier pipeline
    assert {"UTMALDG", "SYNCS"} <= ops("jacobi2d_tb_kernel")         # cp.async.bulk.tensor
