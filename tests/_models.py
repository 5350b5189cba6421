"""Shared model / input builders for the tests."""
import types

import torch

from cistgcn_b200 import CISTGCN
from oracle import cistgcn_oracle as O


def make_opt(E=8, V=22, interp=True):
    ns = types.SimpleNamespace
    mp = ns(input_n=10, output_n=25, joints=V, n_txcnn_layers=4, txc_kernel_size=3, reduction=8, hidden_dim=64,
            input_gcn=ns(model_complexity=[E] * 4, interpretable=[interp] * 5),
            output_gcn=ns(model_complexity=[3], interpretable=[interp]), clipping=15)
    return ns(architecture_config=ns(model="CISTGCN_0", model_params=mp), learning_config=ns(dropout=0.1))


def build(E, V, weights="W1", interp=True, seed=0):
    """(model in eval mode on CPU, state_dict clone, OracleConfig)."""
    opt = make_opt(E, V, interp)
    torch.manual_seed(seed)
    m = CISTGCN(opt.architecture_config, opt.learning_config).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    if weights == "W2":
        O.stress_init_(sd, seed=7)
        m.load_state_dict(sd)
    elif weights != "W1":
        raise ValueError(weights)
    return m, sd, O.OracleConfig(joints=V, input_gcn=[E] * 4)
