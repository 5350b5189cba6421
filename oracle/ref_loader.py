"""Loader / vendoring recipe for the UNMODIFIED reference model  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The reference (QualityMinds/cistgcn) is a Python package; its model imports cleanly in isolation
(SURVEY.md App. D).  Two places can hold it:

  /root/reference/human_motion_prediction      the read-only checkout (build container only)
  oracle/_ref/human_motion_prediction          a copy of the 7 files App. D lists, made by ``vendor()``
                                               (called from ``__graft_entry__.build()``); git-ignored, so it
                                               never enters the history, but it travels to the GPU box with
                                               the gpurun snapshot, where ``bench.py --impl reference`` times it

Only tests/, __graft_entry__ and bench.py's reference / cpu_baseline legs import this module.
"""
import copy
import importlib
import os
import shutil
import sys
import types
from unittest.mock import MagicMock

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/human_motion_prediction"
VENDORED = os.path.join(_HERE, "_ref", "human_motion_prediction")
# minimal file set: the model, its SE layers, the yaml loader and the two shipped configs
FILES = ["models/CISTGCN/CISTGCN.py", "models/CISTGCN/__init__.py", "models/layers/SE.py", "models/layers/__init__.py",
         "utils/yaml_utils.py", "config/CISTGCN/train_h36m.yaml", "config/CISTGCN/train_amass.yaml"]


def vendor() -> bool:
    """Copy FILES from the read-only checkout into oracle/_ref/ (outputs only there).  Returns True when the
    vendored copy is complete afterwards."""
    if os.path.isdir(SRC):
        for f in FILES:
            dst = os.path.join(VENDORED, f)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            src = os.path.join(SRC, f)
            if not os.path.isfile(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
                shutil.copyfile(src, dst)
    return all(os.path.isfile(os.path.join(VENDORED, f)) for f in FILES)


def root() -> str:
    env = os.environ.get("CISTGCN_REFERENCE")
    if env:
        return env
    if os.path.isfile(os.path.join(SRC, FILES[0])):
        return SRC
    return VENDORED


def available() -> bool:
    return os.path.isfile(os.path.join(root(), FILES[0]))


def which() -> str:
    r = root()
    return "checkout" if r == SRC else ("vendored" if r == VENDORED else "env")


def modules():
    """(CISTGCN module, yaml_utils module) of the reference, imported without the package __init__
    (which pulls matplotlib, fvcore, tensorboardX)."""
    if "human_motion_prediction" not in sys.modules:
        pkg = types.ModuleType("human_motion_prediction")
        pkg.__path__ = [root()]
        sys.modules["human_motion_prediction"] = pkg
        for m in ("matplotlib", "matplotlib.pyplot", "tensorboardX", "fvcore", "fvcore.nn"):
            sys.modules.setdefault(m, MagicMock())
    M = importlib.import_module("human_motion_prediction.models.CISTGCN.CISTGCN")
    yu = importlib.import_module("human_motion_prediction.utils.yaml_utils")
    return M, yu


def load_opt(dataset: str = "h36m"):
    _, yu = modules()
    return yu.load_yaml(os.path.join(root(), "config", "CISTGCN", f"train_{dataset}.yaml"), class_mode=True)


def build(embed: int, joints: int, seed: int = 0, interpretable: bool = True, dropout=None):
    """The reference CISTGCN (eval mode); the config is deep-copied because the constructor mutates it
    (CISTGCN.py:516-517, 548)."""
    import torch
    M, _ = modules()
    opt = copy.deepcopy(load_opt("h36m" if joints == 22 else "amass"))
    mp = opt.architecture_config.model_params
    mp.input_gcn.model_complexity = [embed] * 4
    mp.joints = joints
    if not interpretable:
        mp.input_gcn.interpretable = [False] * 5
        mp.output_gcn.interpretable = [False]
    if dropout is not None:
        opt.learning_config.dropout = dropout
    torch.manual_seed(seed)
    return M.CISTGCN(opt.architecture_config, opt.learning_config).eval()
