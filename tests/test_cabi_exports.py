"""The nvcc-built library loads without a GPU and exports every symbol include/cistgcn_b200.h declares."""
import ctypes
import os
import re
import shutil

import pytest

import __graft_entry__ as entry
from cistgcn_b200 import _cabi
from cistgcn_b200.pack import F, DEFINES

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "cistgcn_b200.h")


def _declared():
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(cistgcn_[a-z0-9_]+)\s*\(", src)))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.isfile(_cabi.LIB_PATH), reason="no nvcc and no prebuilt library")
def test_library_exports_every_declared_symbol():
    entry.build()
    L = ctypes.CDLL(_cabi.LIB_PATH)
    names = _declared()
    assert set(names) == set(_cabi.EXPORTS)
    for n in names:
        assert hasattr(L, n), n
    assert _cabi.lib().cistgcn_abi_version() == DEFINES["CISTGCN_ABI_VERSION"]


def test_error_reporting_without_gpu():
    if not os.path.isfile(_cabi.LIB_PATH):
        pytest.skip("library not built")
    L = _cabi.lib()
    bad = (ctypes.c_int32 * F["CP_HEADER_COUNT"])()
    assert L.cistgcn_workspace_bytes(bad, 4) == 0            # ABI field is 0 -> rejected
    assert b"ABI" in L.cistgcn_last_error()
    assert L.cistgcn_mpjpe_f32(None, None, 4, 25, 22, None, None, None) < 0
    assert b"NULL" in L.cistgcn_last_error()


def test_enum_parser_matches_header_order():
    assert F["CB_CI"] == 0 and F["CF_CIN"] == 0 and F["CT_TIN"] == 0 and F["CP_ABI"] == 0
    assert F["CB_COUNT"] > F["CB_RS_B"] and F["CT_COUNT"] > F["CT_SE2_WT"]
    assert F["CB_TC3_WT_T"] == F["CB_TC3_WT_S"] + 1 and F["CB_P_A_T"] == F["CB_P_A_S"] + 1
