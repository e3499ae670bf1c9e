"""The drop-in boundary on a machine without a GPU: libhdiff_b200.so loads, exports every entry point that
include/hdiff_b200.h declares, and the ctypes table (hdiff_b200/_lib.py) binds exactly that set.  No compute calls."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "hdiff_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(?:int|int64_t|const char\s*\*)\s+(hd_\w+)\s*\(", src))


@pytest.fixture(scope="module")
def lib():
    from hdiff_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    return ctypes.CDLL(_lib.LIB_PATH)


def test_header_declares_the_hot_path():
    names = _declared()
    assert len(names) >= 40
    for must in ("hd_conv_tc", "hd_wgrad_tc", "hd_attn_fwd_tc", "hd_attn_bwd_tc", "hd_gn_apply", "hd_gn_bwd_reduce", "hd_gn_bwd_apply",
                 "hd_q_sample", "hd_sampler_step", "hd_adamw_flat"):
        assert must in names, must


def test_library_exports_every_declared_symbol(lib):
    missing = [n for n in sorted(_declared()) if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_table_matches_the_header(lib):
    from hdiff_b200 import _lib
    declared = _declared()
    bound = set(_lib.PROTOTYPES) | {"hd_last_error", "hd_abi_version"}
    assert bound <= declared, sorted(bound - declared)
    unbound = declared - bound
    assert not unbound, f"declared in the header but not bound in _lib.PROTOTYPES: {sorted(unbound)}"
    _lib.load()          # resolves every prototype; raises on a missing symbol


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(ROOT, "hybrid-diffusion-underwater-atmopheric-image-enhancement_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(d, f)
