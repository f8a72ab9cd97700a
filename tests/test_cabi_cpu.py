"""CPU-only: the C-ABI library loads, exports every symbol include/ezflow_b200.h declares, and its
argument validation / layout logic (no compute, no GPU)."""

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from ezflow_b200 import _cabi, build

    build.build()
    return _cabi.lib()


def test_exports_match_header(lib):
    from ezflow_b200 import _cabi

    text = open(os.path.join(ROOT, "include", "ezflow_b200.h")).read()
    declared = set(re.findall(r"\b(?:int|int64_t)\s+(ezb_[a-z0-9_]+)\s*\(", text))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_cabi.EXPORTED_SYMBOLS), declared ^ set(_cabi.EXPORTED_SYMBOLS)


def test_version_and_error_string(lib):
    from ezflow_b200 import _cabi

    assert lib.ezb_version() == 100
    rc = lib.ezb_corr_pyramid_layout(0, 4, 1, 0, None, None, None, None, None, None, None)
    assert rc == -1 and "bad pyramid shape" in _cabi.last_error()
    with pytest.raises(ValueError):
        _cabi.check(rc)


def test_pyramid_layout(lib):
    from ezflow_b200 import _cabi
    from ezflow_b200.similarity.pairwise import PyramidLayout, grad_layout

    # 1080p: floor pooling 135x240 -> 67x120 -> 33x60 -> 16x30, 8x8-pixel subtiles
    lay = PyramidLayout(135, 240, 4, _cabi.EZB_F16)
    assert lay.levels == [(135, 240), (67, 120), (33, 60), (16, 30)]
    assert lay.sub_base == [0, 17 * 30, 17 * 30 + 9 * 15, 17 * 30 + 9 * 15 + 5 * 8, 17 * 30 + 9 * 15 + 5 * 8 + 2 * 4]
    assert (lay.n_tiles, lay.n_groups) == (174, 1013)
    assert lay.group_bytes == 16384 and lay.pair_bytes == 174 * 1013 * 16384
    lay32 = PyramidLayout(46, 62, 4, _cabi.EZB_F32)
    assert lay32.group_bytes == 32768 and lay32.n_groups == 90
    assert lay32.n_tiles == -(-(6 * 8 + 3 * 4 + 2 * 2 + 1 * 1) // 4)
    with pytest.raises(ValueError):
        PyramidLayout(4, 4, 4, 0)  # level 3 would be empty
    with pytest.raises(NotImplementedError):
        PyramidLayout(64, 64, 5, 0)
    # every record column maps to a distinct, in-range, element-aligned offset of the block
    for dt, es, gb in ((_cabi.EZB_F16, 2, 16384), (_cabi.EZB_F32, 4, 32768)):
        seen = set()
        for lane in range(32):
            for n in range(256):
                off = lib.ezb_corr_pyramid_record_offset(dt, lane, n)
                assert 0 <= off < gb and off % es == 0 and off not in seen
                seen.add(off)
        assert len(seen) == 32 * 256
        # an 8x8 subtile (64 consecutive columns) is exactly one (fp16) / two (fp32) 128-byte lines
        for lane in (0, 5, 31):
            for sub in range(4):
                offs = [lib.ezb_corr_pyramid_record_offset(dt, lane, sub * 64 + k) for k in range(64)]
                assert len({o // 128 for o in offs}) == es // 2
    assert lib.ezb_corr_pyramid_record_offset(1, 32, 0) == -1
    # every pixel of every level maps to a distinct (tile, column)
    seen = set()
    for l, (h, w) in enumerate(lay32.levels):
        t, c = lay32.level_map(l)
        assert tuple(t.shape) == (h, w)
        assert int(t.max()) < lay32.n_tiles and int(c.max()) < 256 and int(c.min()) >= 0
        pairs = set(zip(t.reshape(-1).tolist(), c.reshape(-1).tolist()))
        assert len(pairs) == h * w and not (pairs & seen)
        seen |= pairs
    g = grad_layout(46, 62, 4)
    assert [(h, w, p, qs) for h, w, p, qs in g] == [(46, 62, 62, 2852), (23, 31, 31, 716), (11, 15, 15, 168), (5, 7, 7, 36)]


def test_argument_validation_without_gpu(lib):
    from ezflow_b200 import _cabi

    assert lib.ezb_corr_pyramid_fwd(None, None, 1, 8, 8, 8, 1, 0, None, None, None, 0, None) == -1
    assert lib.ezb_corr_lookup_fwd(None, None, None, 1, 8, 8, 1, 4, 0, None, None) == -1
    one = ctypes.c_void_p(16)
    # even patch size -> unsupported, same message as the reference (sampler.py:93-94)
    rc = lib.ezb_local_corr_fwd(one, one, 1, 2, 4, 4, 4, 4, 1, 1, 0, 0, 1, 1, 1.0, 1.0, one, 0, None, 0, None)
    assert rc == -2 and "odd patch sizes" in _cabi.last_error()
    with pytest.raises(NotImplementedError):
        _cabi.check(rc)
    n = ctypes.c_size_t(0)
    assert lib.ezb_corr_pyramid_workspace_bytes(1, 256, 135, 240, 4, ctypes.byref(n)) == 0
    assert n.value >= 2 * 32400 * 256 * 2


def test_sass_is_blackwell_native():
    """The GEMM must be tcgen05 + TMA (SASS: UTCHMMA / LDTM / UTMALDG), not mma.sync."""
    import shutil
    import subprocess

    from ezflow_b200 import build

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", build.LIBPATH], capture_output=True, text=True).stdout
    for mnem in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnem in sass, f"{mnem} missing from SASS"
    assert "HMMA." not in sass.replace("UTCHMMA", "")


def test_header_is_c99_and_library_links_from_plain_c(lib, tmp_path):
    """The boundary is a C ABI: a C99 program includes the header, links the .so and calls host-only entry points."""
    import shutil
    import subprocess

    from ezflow_b200 import _cabi

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    exe = str(tmp_path / "cabi_link")
    libdir = os.path.dirname(_cabi.LIBPATH)
    cmd = [gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c", "cabi_link.c"), "-o", exe, "-L", libdir, "-lezflow_b200", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "cabi_link ok" in r.stdout, (r.returncode, r.stdout, r.stderr)
