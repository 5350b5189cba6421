"""Loader for the UNMODIFIED reference model (tests / golden generation only).

Follows SURVEY.md App. D: registers a stub parent package so hmp/__init__.py (which imports
matplotlib, fvcore, tensorboardX) is skipped.  /root/reference only exists in the build
container, never on the GPU box -- callers must check ``available()`` and skip otherwise.
"""
import copy
import importlib
import os
import sys
import types
from unittest.mock import MagicMock

REF = os.environ.get("CISTGCN_REFERENCE", "/root/reference/human_motion_prediction")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "models", "CISTGCN", "CISTGCN.py"))


def _modules():
    if "human_motion_prediction" not in sys.modules:
        pkg = types.ModuleType("human_motion_prediction")
        pkg.__path__ = [REF]
        sys.modules["human_motion_prediction"] = pkg
        for m in ("matplotlib", "matplotlib.pyplot", "tensorboardX", "fvcore", "fvcore.nn"):
            sys.modules.setdefault(m, MagicMock())
    M = importlib.import_module("human_motion_prediction.models.CISTGCN.CISTGCN")
    yu = importlib.import_module("human_motion_prediction.utils.yaml_utils")
    return M, yu


def load_opt(dataset: str = "h36m"):
    _, yu = _modules()
    return yu.load_yaml(os.path.join(REF, "config", "CISTGCN", f"train_{dataset}.yaml"), class_mode=True)


def build(embed: int, joints: int, seed: int = 0, interpretable: bool = True):
    """Reference CISTGCN in eval mode; config deep-copied because the ctor mutates it."""
    import torch
    M, _ = _modules()
    opt = copy.deepcopy(load_opt("h36m" if joints == 22 else "amass"))
    mp = opt.architecture_config.model_params
    mp.input_gcn.model_complexity = [embed] * 4
    mp.joints = joints
    if not interpretable:
        mp.input_gcn.interpretable = [False] * 5
        mp.output_gcn.interpretable = [False]
    torch.manual_seed(seed)
    return M.CISTGCN(opt.architecture_config, opt.learning_config).eval()


def ref_mpjpe():
    _modules()
    L = importlib.import_module("human_motion_prediction.losses.losses")
    return L.mpjpe
