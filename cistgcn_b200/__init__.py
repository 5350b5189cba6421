"""cistgcn_b200 -- B200-native (sm_100a) implementation of the CIST-GCN forward hot path.

Public surface mirrors the reference for this path only:
  CISTGCN(arch, learn).forward(x) -> (pred,)     models/CISTGCN/CISTGCN.py:478-597
  choose_net(architecture, opt)                   models/choose_net.py:4-11
  mpjpe(predicted, target, reduce_axis=[])        losses/losses.py:50-61
"""
from .model import CISTGCN, choose_net
from .losses import mpjpe

CISTGCN_0 = CISTGCN        # registry aliases of models/__init__.py:1-2
CISTGCN_eval = CISTGCN

__all__ = ["CISTGCN", "CISTGCN_0", "CISTGCN_eval", "choose_net", "mpjpe"]
