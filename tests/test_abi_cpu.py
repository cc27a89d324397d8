"""The drop-in boundary without a GPU: libhmgan.so builds (nvcc cross-compiles for sm_100a), loads through ctypes and
exports every entry point that include/hmgan.h declares; the Python prototype table covers the same set; argument
errors come back as status codes with a message, never as crashes.  No kernel is launched here."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gan-heightmaps_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)
import _lib   # noqa: E402


def _declared():
    src = open(os.path.join(ROOT, "include", "hmgan.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"^\s*(?:int|long long|const char\s*\*)\s+(hm_[a-z0-9_]+)\s*\(", src, flags=re.M)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(PKG, "csrc"), "-j", "8"])
    return C.CDLL(_lib.LIB_PATH)


def test_header_declares_the_documented_surface():
    names = _declared()
    assert len(names) >= 40
    for must in ("hm_tc_conv", "hm_tc_wgrad", "hm_conv_gather", "hm_conv_wgrad", "hm_c1s2_conv", "hm_c1s2_bwd",
                 "hm_bn_finalize", "hm_maxpool2_bwd_scaled", "hm_adv_loss_pair", "hm_rmsprop", "hm_adam", "hm_version",
                 "hm_last_error_string"):
        assert must in names, must


def test_library_exports_every_declared_symbol(lib):
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, "include/hmgan.h declares symbols the library does not export: %s" % missing


def test_python_prototypes_cover_the_header():
    protos = set(_lib.exported_symbols())
    declared = set(_declared())
    assert declared <= protos, sorted(declared - protos)
    assert protos <= declared, sorted(protos - declared)


def test_bad_arguments_are_reported_not_fatal(lib):
    lib.hm_last_error_string.restype = C.c_char_p
    assert lib.hm_version() > 0
    # null pointers / impossible sizes: negative status + a message (checked before anything touches the device)
    rc = lib.hm_col_sum(None, 1, C.c_longlong(8), 8, None, None)
    assert rc < 0 and b"hm_col_sum" in lib.hm_last_error_string()
    rc = lib.hm_pack_conv_weight(None, None, 0, 1, 1, 1, 1, 0, 0, 1, None)
    assert rc < 0 and b"hm_pack_conv_weight" in lib.hm_last_error_string()
    rc = lib.hm_c1s2_conv(None, None, None, None, None, 1, 8, 8, 64, 0, C.c_float(0.0), None)
    assert rc < 0 and b"hm_c1s2_conv" in lib.hm_last_error_string()
    assert lib.hm_tc_conv_supported(None) == 0


def test_pack_element_counts_agree_between_python_and_the_library(lib):
    """_lib.pack_count (used to fill the HmPackJob table of hm_pack_conv_weight_multi) == hm_pack_conv_weight_count."""
    lib.hm_pack_conv_weight_count.restype = C.c_longlong
    for mode in (0, 1, 2, 3, 4, 5, 6, 7, 8, 11, 12, 14, 15, 16, 17, 18, 19, 20, 21, 22):
        for (cout, cin, kh, kw) in ((64, 1, 5, 5), (128, 64, 5, 5), (3, 128, 2, 2), (64, 4, 3, 3), (1, 64, 5, 5)):
            assert _lib.pack_count(mode, cout, cin, kh, kw) == lib.hm_pack_conv_weight_count(mode, cout, cin, kh, kw), \
                (mode, cout, cin, kh, kw)
    raw, max_n = _lib.pack_job_table([(0x1000, 0x2000, 5, 128, 64, 5, 5, 0, 0, 1), (0x3000, 0x4000, 15, 64, 1, 5, 5, 0, 0, 1)])
    assert len(raw) == 2 * 56 and max_n == 128 * 64 * 25
