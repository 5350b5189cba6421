"""Loader for the UNMODIFIED reference model (tests / golden generation only): thin front of
oracle/ref_loader.py.  /root/reference only exists in the build container; on the GPU box
the vendored copy under oracle/_ref/ (git-ignored, shipped by gpurun) is used when present -- callers must
check ``available()`` and skip otherwise."""
import importlib

from oracle import ref_loader as _R

available = _R.available
load_opt = _R.load_opt
build = _R.build
_modules = _R.modules


def ref_mpjpe():
    """losses.mpjpe of the reference when the checkout is mounted (losses.py needs utils/ that are not vendored)."""
    _R.modules()
    L = importlib.import_module("human_motion_prediction.losses.losses")
    return L.mpjpe
