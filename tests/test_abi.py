"""The C-ABI library: it loads, exports every symbol include/lrc.h declares, and -- with no GPU -- fails loudly."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    import __graft_entry__
    __graft_entry__.build()
    from lrc_b200 import _native
    return _native


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "lrc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lrc_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = _header_symbols()
    for must in ("lrc_create", "lrc_destroy", "lrc_last_error", "lrc_set_mesh", "lrc_cast_rays", "lrc_rays_intersect",
                 "lrc_scan_rays", "lrc_scan_single_axis", "lrc_scan_dual_axis", "lrc_gen_rays_single_axis",
                 "lrc_gen_rays_dual_axis", "lrc_counters"):
        assert must in syms


def test_library_exports_every_declared_symbol(native):
    lib = native.load()
    declared = _header_symbols()
    assert sorted(native.SYMBOLS) == declared          # the ctypes table and the header agree
    for name in declared:
        assert hasattr(lib, name), f"liblrc.so does not export {name}"
    assert lib.lrc_abi_version() == 1


def test_no_cpu_fallback(native):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = native.load()
    h = ctypes.c_void_p()
    rc = lib.lrc_create(0, ctypes.byref(h))
    assert rc == -5 and not h.value                     # LRC_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.lrc_last_error(None)
    import lrc_b200 as lrc
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lrc.RaycastEngineGPU()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lrc.IndoorLidar(lrc.Indoor8LineLidarIntrinsics(), __import__("numpy").eye(4)).get_rays()


def test_product_never_imports_the_oracle():
    """Nothing under the product package may import, link or dlopen anything from oracle/."""
    pkg = os.path.join(ROOT, "indoor-point-cloud-datasets-controllable-generation-method-for-mobile-robots-3d-scene-perception_b200")
    bad = re.compile(r"^\s*(from|import)\s+\.*oracle\b|liblrc_oracle|\borc_[a-z_]+\s*\(|import_module\([^)]*oracle", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not bad.search(src), f"{f} references the oracle"
