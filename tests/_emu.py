"""Drives the SIMT-emulator build of the kernel sources on CPU tensors (tests only)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import build_emu  # noqa: E402
from cistgcn_b200 import _cabi  # noqa: E402
from cistgcn_b200.pack import pack_state_dict  # noqa: E402

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = _cabi.bind(build_emu.build())
    return _LIB


def forward(sd, geom, x, target=None, want_taps=True):
    """Full forward through cistgcn_forward_f32 of the emulator library.  Returns (pred, sums, taps)."""
    L = lib()
    pk = pack_state_dict(sd, geom, "cpu")
    B = x.shape[0]
    x = x.contiguous()
    pred = torch.empty(B, geom.output_n, geom.joints, 3)
    ws = torch.empty(L.cistgcn_workspace_bytes(pk.plan_c, B), dtype=torch.uint8)
    sums = torch.zeros(geom.output_n, dtype=torch.float64) if target is not None else None
    taps_struct, holders = _cabi.make_taps(geom, B, "cpu") if want_taps else (None, {})
    rc = L.cistgcn_forward_f32(pk.plan_c, len(pk.plan), pk.blob.data_ptr(), x.data_ptr(), pred.data_ptr(),
                               target.contiguous().data_ptr() if target is not None else None,
                               sums.data_ptr() if sums is not None else None,
                               ws.data_ptr(), ws.numel(), B, taps_struct, None)
    _cabi.check(rc, "cistgcn_forward_f32[emu]", L)
    return pred, sums, holders
