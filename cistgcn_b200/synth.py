"""Synthetic workloads for benchmarks and examples (datasets and checkpoints are unavailable offline).

Shapes follow the reference's data (H36M: 22 joints, AMASS: 18 joints; 10 input / 25 output frames,
config/CISTGCN/train_{h36m,amass}.yaml); the generator is the one BASELINE.md section 2 fixes:
x = N(0,1)[B,1,V,3] + cumsum_t N(0, 0.03^2)[B,T,V,3], target = x[:, -1:] + N(0, 0.1^2)."""
from __future__ import annotations

import types
from typing import Tuple

import torch


def make_opt(embed: int = 32, joints: int = 22, input_n: int = 10, output_n: int = 25, interpretable: bool = True,
             dropout: float = 0.1):
    """Config object with the fields CISTGCN(arch, learn) reads (architecture_config.model_params of train_h36m.yaml:3-28)."""
    ns = types.SimpleNamespace
    mp = ns(input_n=input_n, output_n=output_n, joints=joints, n_txcnn_layers=4, txc_kernel_size=3, reduction=8,
            hidden_dim=64, input_gcn=ns(model_complexity=[embed] * 4, interpretable=[interpretable] * 5),
            output_gcn=ns(model_complexity=[3], interpretable=[interpretable]), clipping=15)
    return ns(architecture_config=ns(model="CISTGCN_0", model_params=mp), learning_config=ns(dropout=dropout))


def synth_inputs(batch: int, joints: int = 22, input_n: int = 10, output_n: int = 25, seed: int = 123,
                 scale: str = "unit") -> Tuple[torch.Tensor, torch.Tensor]:
    """(x (B, input_n, V, 3), target (B, output_n, V, 3)) float32 on the CPU.  scale "mm": 50 + 350 * unit."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(batch, 1, joints, 3, generator=g)
    steps = 0.03 * torch.randn(batch, input_n, joints, 3, generator=g)
    x = base + steps.cumsum(1)
    tgt = x[:, -1:] + 0.1 * torch.randn(batch, output_n, joints, 3, generator=g)
    if scale == "mm":
        x, tgt = 50.0 + 350.0 * x, 50.0 + 350.0 * tgt
    elif scale != "unit":
        raise ValueError(scale)
    return x.contiguous(), tgt.contiguous()
