"""CPU oracle for the CIST-GCN hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this file.  The product package
(``cistgcn_b200``) never imports it and has no CPU fallback.

What it is
----------
A functional restatement (plain functions over a raw ``state_dict``; no nn.Module
classes) of the eval-mode and train-mode forward pass of the reference model
``human_motion_prediction/models/CISTGCN/CISTGCN.py`` and of ``losses.mpjpe``.  Every
function cites the reference file:line it follows.  The arithmetic is executed by
PyTorch's CPU ATen kernels -- the same library calls the reference itself makes -- in
fp32 (or fp64 when the state_dict / input are cast to double), which also makes it the
honest "reference on host cores" timing arm for bench.py (kind = "port").

Pinning
-------
The reference holds no golden vectors or tests.  The oracle is
pinned against outputs of the *reference module itself*, executed in the build
container by ``tests/golden/make_golden.py`` (which imports /root/reference) and
committed as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them.
When /root/reference is present the same test also compares live, on more configs.

Conventions: x is (B, T_in, V, 3); activations inside are (B, C, T, V) like the
reference.  ``sd`` maps reference parameter names to tensors.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # torch.nn.BatchNorm default, never overridden by the reference
BN_MOMENTUM = 0.1


@dataclass
class OracleConfig:
    """The subset of architecture_config.model_params the path reads (CISTGCN.py:491-503)."""
    input_n: int = 10
    output_n: int = 25
    joints: int = 22
    n_txcnn_layers: int = 4
    txc_kernel_size: int = 3
    reduction: int = 8
    hidden_dim: int = 64
    input_gcn: List[int] = field(default_factory=lambda: [32, 32, 32, 32])
    output_gcn: List[int] = field(default_factory=lambda: [3])
    in_ch: int = 10  # hard-coded at CISTGCN.py:512

    @property
    def input_chain(self) -> List[int]:
        # CISTGCN.py:516-517: insert(0, in_ch); append(in_ch)
        return [self.in_ch] + list(self.input_gcn) + [self.in_ch]

    @property
    def output_chain(self) -> List[int]:
        # CISTGCN.py:548: insert(0, 3)
        return [3] + list(self.output_gcn)


class _Ctx:
    """Carries the state dict, mode flags and the taps collected during one forward."""

    def __init__(self, sd: Dict[str, torch.Tensor], train: bool, taps: Optional[dict],
                 bn_updates: Optional[dict]):
        self.sd = sd
        self.train = train
        self.taps = taps
        self.bn_updates = bn_updates

    def tap(self, name: str, t: torch.Tensor):
        if self.taps is not None:
            self.taps[name] = t.detach().clone()


# ----------------------------------------------------------------------------------
# primitive layers
# ----------------------------------------------------------------------------------
def _bn(c: _Ctx, x: torch.Tensor, p: str) -> torch.Tensor:
    """nn.BatchNorm{1,2}d.  eval: running stats; train: batch stats (biased var for the
    normalisation, unbiased for the running update, momentum 0.1)."""
    sd = c.sd
    if not c.train:
        return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                            sd[p + ".weight"], sd[p + ".bias"], False, 0.0, BN_EPS)
    rm = sd[p + ".running_mean"].detach().clone()
    rv = sd[p + ".running_var"].detach().clone()
    y = F.batch_norm(x, rm, rv, sd[p + ".weight"], sd[p + ".bias"], True, BN_MOMENTUM, BN_EPS)
    if c.bn_updates is not None:
        c.bn_updates[p + ".running_mean"] = rm
        c.bn_updates[p + ".running_var"] = rv
    return y


def _prelu(c: _Ctx, x: torch.Tensor, p: str) -> torch.Tensor:
    return F.prelu(x, c.sd[p + ".weight"])


def _conv(c: _Ctx, x: torch.Tensor, p: str, padding=0, dilation=1) -> torch.Tensor:
    return F.conv2d(x, c.sd[p + ".weight"], c.sd.get(p + ".bias"), padding=padding, dilation=dilation)


def _linear(c: _Ctx, x: torch.Tensor, p: str) -> torch.Tensor:
    return F.linear(x, c.sd[p + ".weight"], c.sd.get(p + ".bias"))


def _se(c: _Ctx, x: torch.Tensor, p: str) -> torch.Tensor:
    """SELayer1d / SELayer2d (models/layers/SE.py:16-20, 37-41): x * sigmoid(W2 relu(W1 mean(x)))."""
    b, ch = x.shape[0], x.shape[1]
    y = x.reshape(b, ch, -1).mean(-1)
    y = torch.sigmoid(_linear(c, torch.relu(_linear(c, y, p + ".excitation.0")), p + ".excitation.2"))
    return x * y.reshape(b, ch, *([1] * (x.dim() - 2)))


# ----------------------------------------------------------------------------------
# Map2Adj  (CISTGCN.py:127-189)
# ----------------------------------------------------------------------------------
def map2adj(c: _Ctx, x: torch.Tensor, p: str, domain: str) -> torch.Tensor:
    # time_compress: conv1x1, BN, PReLU, conv (T,1), BN, [dropout], conv1x1 -> (B, T, 1, V)   :138-145
    a = _prelu(c, _bn(c, _conv(c, x, p + ".time_compress.0"), p + ".time_compress.1"), p + ".time_compress.2")
    a = _bn(c, _conv(c, a, p + ".time_compress.3"), p + ".time_compress.4")
    dim_seq = _conv(c, a, p + ".time_compress.6")
    # joint_compress: conv1x1, BN, PReLU, conv (1,V), BN, [dropout], conv1x1 -> (B, V, T, 1)  :146-153
    g = _prelu(c, _bn(c, _conv(c, x, p + ".joint_compress.0"), p + ".joint_compress.1"), p + ".joint_compress.2")
    g = _bn(c, _conv(c, g, p + ".joint_compress.3"), p + ".joint_compress.4")
    dim_space = _conv(c, g, p + ".joint_compress.6")
    if domain == "space":   # :155-158, 187  -> (B, V, T, T): o[v,t,q] = dsp[v,t] * dseq[q,v]
        o = torch.matmul(dim_space, dim_seq.permute(0, 3, 2, 1))
    else:                   # :159-162, 187  -> (B, T, V, V): o[t,v,w] = dsp[v,t] * dseq[t,w]
        o = torch.matmul(dim_space.permute(0, 2, 1, 3), dim_seq)
    # expansor: conv1x1, BN, [dropout], PReLU, conv1x1 over the leading (V or T) axis        :165-170, 188
    e = _prelu(c, _bn(c, _conv(c, o, p + ".expansor.0"), p + ".expansor.1"), p + ".expansor.3")
    return _conv(c, e, p + ".expansor.4")


# ----------------------------------------------------------------------------------
# Domain_GCNN_layer  (CISTGCN.py:192-269) + ConvTemporalGraphical (:86-124)
# ----------------------------------------------------------------------------------
def domain_layer(c: _Ctx, x: torch.Tensor, p: str, domain: str, interpretable: bool) -> torch.Tensor:
    has_res = (p + ".residual.0.weight") in c.sd
    res = _bn(c, _conv(c, x, p + ".residual.0"), p + ".residual.1") if has_res else x      # :239-247, 261
    if interpretable:
        adj = map2adj(c, x, p + ".map_to_adj", domain)                                      # :262
        c.tap(p + ".Adj", adj)
        eq = "nctv,nvtq->ncqv" if domain == "space" else "nctv,ntvw->nctw"                  # :110, 117
    else:
        adj = c.sd[p + ".gcn.A"]                                                            # :106-115
        eq = "nctv,vtq->ncqv" if domain == "space" else "nctv,tvw->nctw"
    g = torch.einsum(eq, x, adj)                                                            # :123
    y = _bn(c, _conv(c, g, p + ".tcn.0"), p + ".tcn.1")                                     # :229-237, 266
    return _prelu(c, y + res, p + ".prelu")                                                 # :267-268


# ----------------------------------------------------------------------------------
# DSTD_GC  (CISTGCN.py:273-390)
# ----------------------------------------------------------------------------------
def get_stats(x: torch.Tensor) -> torch.Tensor:
    """CISTGCN.py:360-371 -- (B, 2 + 2T); torch.std is Bessel-corrected."""
    return torch.cat((x.mean((3, 2)).mean(1, keepdim=True),
                      x.mean(3).mean(1),
                      x.std((3, 2)).std(1, keepdim=True),
                      x.std(3).std(1)), dim=1)


def _gate(c: _Ctx, xn: torch.Tensor, stats: torch.Tensor, conv_p: str, map_p: str) -> torch.Tensor:
    b = xn.shape[0]
    h = _prelu(c, _bn(c, _conv(c, xn, conv_p + ".0"), conv_p + ".1"), conv_p + ".3")        # :323-326
    h = _prelu(c, _bn(c, _conv(c, h, conv_p + ".4"), conv_p + ".5"), conv_p + ".7")         # :327-330
    w = torch.cat((h.reshape(b, -1), stats), dim=1)                                         # :378, 380
    z = _prelu(c, _bn(c, _linear(c, w, map_p + ".0"), map_p + ".1"), map_p + ".3")          # :341-344
    return _linear(c, z, map_p + ".4")                                                      # :345


def dstd_gc(c: _Ctx, x: torch.Tensor, p: str, interpretable: bool = True) -> torch.Tensor:
    xn = _bn(c, x, p + ".global_norm")                                                      # :375
    stats = get_stats(xn)                                                                   # :377, 379
    w1 = _gate(c, xn, stats, p + ".conv_s", p + ".map_s")                                   # :378, 381
    w2 = _gate(c, xn, stats, p + ".conv_t", p + ".map_t")                                   # :380, 382
    c.tap(p + ".w1", w1)
    c.tap(p + ".w2", w2)
    x1 = domain_layer(c, xn, p + ".dsgn", "space", interpretable)                           # :386
    x2 = domain_layer(c, xn, p + ".tsgn", "time", interpretable)                            # :387
    u1 = _prelu(c, _bn(c, w1[..., None, None] * x1, p + ".prelu1.0"), p + ".prelu1.1")      # :388
    u2 = _prelu(c, _bn(c, w2[..., None, None] * x2, p + ".prelu2.0"), p + ".prelu2.1")
    u = torch.cat((u1, u2), dim=1)
    cc = _prelu(c, _bn(c, _conv(c, u, p + ".compressor.0"), p + ".compressor.1"), p + ".compressor.2")
    cc = _se(c, cc, p + ".compressor.3")                                                    # :305-309, 389
    has_res = (p + ".residual.0.weight") in c.sd
    res = _bn(c, _conv(c, xn, p + ".residual.0"), p + ".residual.1") if has_res else xn     # :310-318
    return cc + res                                                                         # :390


# ----------------------------------------------------------------------------------
# FPN  (CISTGCN.py:38-79)
# ----------------------------------------------------------------------------------
def fpn(c: _Ctx, x: torch.Tensor, p: str, k: int) -> torch.Tensor:
    pad = (k - 1) // 2                                                                      # :47
    outs = []
    for i in (1, 2, 3):                                                                     # :48-68: pad = dil = i*pad
        bp = f"{p}.block{i}"
        d = 1 + (i - 1) * pad
        y = F.conv2d(x, c.sd[bp + ".0.weight"], c.sd[bp + ".0.bias"], padding=i * pad, dilation=d)
        outs.append(_prelu(c, _bn(c, y, bp + ".1"), bp + ".3"))
    outs.append(x.mean((2, 3), keepdim=True).expand(-1, -1, x.shape[2], x.shape[3]))        # :69, 76
    return F.conv2d(torch.cat(outs, dim=1), c.sd[p + ".compress.weight"], c.sd[p + ".compress.bias"])  # :77-78


# ----------------------------------------------------------------------------------
# ContextLayer  (CISTGCN.py:393-475)
# ----------------------------------------------------------------------------------
def context_layer(c: _Ctx, z: torch.Tensor, p: str, n_out: int, n_joints: int) -> torch.Tensor:
    b, _, _, jd = z.shape
    cv = lambda i: _prelu(c, _bn(c, _conv(c, z, f"{p}.context_conv{i}.0"), f"{p}.context_conv{i}.1"),
                          f"{p}.context_conv{i}.2")
    y1 = cv(1).amax(dim=(2, 3))                                                             # :465
    y2 = cv(2).reshape(b, -1, jd).amax(dim=-1)                                              # :466
    ym = cv(3).mean((2, 3))                                                                 # :467
    y = torch.cat((_prelu(c, _linear(c, y1, p + ".map1.0"), p + ".map1.2"),
                   _prelu(c, _linear(c, y2, p + ".map2.0"), p + ".map2.2"),
                   _prelu(c, _linear(c, ym, p + ".map3.0"), p + ".map3.2")), dim=1)         # :468
    joints = _bn(c, _linear(c, y, p + ".fmap_s.0"), p + ".fmap_s.1")                        # :469
    disp = _bn(c, _linear(c, y, p + ".fmap_t.0"), p + ".fmap_t.1")                          # :470
    sj = torch.bmm(disp.unsqueeze(2), joints.unsqueeze(1))                                  # :471  (B, n_out, V)
    n = F.conv1d(sj, c.sd[p + ".norm_map.0.weight"])                                        # :443-451
    n = _prelu(c, _bn(c, n, p + ".norm_map.1"), p + ".norm_map.3")
    n = _se(c, n, p + ".norm_map.4")
    n = F.conv1d(n, c.sd[p + ".norm_map.5.weight"])
    sjn = _prelu(c, _bn(c, n, p + ".norm_map.6"), p + ".norm_map.8")                        # :472
    f = sjn.reshape(b, 1, n_out, n_joints)
    f = _prelu(c, _bn(c, _conv(c, f, p + ".fconv.0"), p + ".fconv.1"), p + ".fconv.2")      # :454-460
    sjd = _prelu(c, _bn(c, _conv(c, f, p + ".fconv.3"), p + ".fconv.4"), p + ".fconv.5")    # :473
    c.tap(p + ".joints", joints)
    c.tap(p + ".displacements", disp)
    c.tap(p + ".seq_joints_n", sjn)
    c.tap(p + ".seq_joints_dims", sjd)
    return _se(c, sjd.permute(0, 2, 3, 1), p + ".SE")                                       # :474


# ----------------------------------------------------------------------------------
# CISTGCN.forward  (CISTGCN.py:567-597)
# ----------------------------------------------------------------------------------
def build_features(x: torch.Tensor) -> torch.Tensor:
    """CISTGCN.py:568-577.  (B,T,V,3) -> (B,10,T,V), channel order [x, acc, vel, |vel|]."""
    vel = torch.zeros_like(x)
    vel[:, :-1] = x[:, 1:] - x[:, :-1]
    vel[:, -1] = x[:, -1]
    acc = torch.zeros_like(x)
    acc[:, :-1] = vel[:, 1:] - vel[:, :-1]
    acc[:, -1] = vel[:, -1]
    speed = torch.linalg.vector_norm(vel, dim=-1, keepdim=True)
    return torch.cat((x, acc, vel, speed), dim=-1).permute(0, 3, 1, 2)


def fpn_stack(c: _Ctx, x5: torch.Tensor, cfg: OracleConfig) -> torch.Tensor:
    """CISTGCN.py:584-589: the FPN layers with the caller's PReLU / residual, dim_conversor, cumsum.
    x5 (B, input_n, 10, V) -> x7 (B, output_n, V, 3)."""
    x6 = _prelu(c, fpn(c, x5, "txcnns.0", cfg.txc_kernel_size), "prelus.0")                 # :584
    for i in range(1, cfg.n_txcnn_layers):                                                  # :585-586
        x6 = _prelu(c, fpn(c, x6, f"txcnns.{i}", cfg.txc_kernel_size), f"prelus.{i}") + x6
    c.tap("x6", x6)
    d = x6.permute(0, 2, 1, 3)                                                              # :588, 541-545
    d = _prelu(c, _bn(c, _conv(c, d, "dim_conversor.0"), "dim_conversor.1"), "dim_conversor.2")
    d = _prelu(c, _conv(c, d, "dim_conversor.3"), "dim_conversor.4")
    return d.permute(0, 2, 3, 1).cumsum(1)                                                  # :588-589


def fpn_stack_eval(sd: Dict[str, torch.Tensor], cfg: OracleConfig, x5: torch.Tensor) -> torch.Tensor:
    """Eval-mode fpn_stack on a raw state_dict (entry point for the FPN-only parity tests)."""
    with torch.no_grad():
        return fpn_stack(_Ctx(sd, False, None, None), x5, cfg)


def forward(sd: Dict[str, torch.Tensor], cfg: OracleConfig, x: torch.Tensor, *, train: bool = False,
            taps: Optional[dict] = None, bn_updates: Optional[dict] = None,
            interpretable_in: Optional[List[bool]] = None,
            interpretable_out: Optional[List[bool]] = None) -> torch.Tensor:
    """Full forward; returns pred (B, output_n, V, 3) (the reference returns the 1-tuple (pred,))."""
    c = _Ctx(sd, train, taps, bn_updates)
    b, _, joints, _ = x.shape
    h = build_features(x)
    chain = cfg.input_chain
    for i in range(len(chain) - 1):                                                         # :579-580
        itp = True if interpretable_in is None else interpretable_in[i]
        h = dstd_gc(c, h, f"st_gcnns.{i}", itp)
        c.tap(f"st_gcnns.{i}.out", h)
    x5 = h.permute(0, 2, 1, 3)                                                              # :582
    x7 = fpn_stack(c, x5, cfg)                                                              # :584-589
    c.tap("x7", x7)
    act = context_layer(c, x7.reshape(b, 1, cfg.output_n, joints * 3), "context_layer",
                        cfg.output_n, joints)                                               # :591
    c.tap("act", act)
    x8 = x7.permute(0, 3, 2, 1)                                                             # :592
    ochain = cfg.output_chain
    for i in range(len(ochain) - 1):                                                        # :593-594
        itp = True if interpretable_out is None else interpretable_out[i]
        x8 = dstd_gc(c, x8, f"st_gcnns_o.{i}", itp)
    x9 = x8.permute(0, 3, 2, 1) + act                                                       # :595
    return x[:, -1:] + x9                                                                   # :597


# ----------------------------------------------------------------------------------
# losses.mpjpe  (losses/losses.py:50-61)
# ----------------------------------------------------------------------------------
def mpjpe(pred: torch.Tensor, target: torch.Tensor, reduce_axis=()) -> torch.Tensor:
    """reduce_axis: () / [] -> scalar mean over everything (train), (0, 2) -> per frame (eval),
    None -> no reduction (B, T, V)."""
    assert pred.shape == target.shape
    err = torch.linalg.vector_norm(pred - target, ord=2, dim=-1)
    if reduce_axis is None:
        return err
    if isinstance(reduce_axis, int):
        reduce_axis = (reduce_axis,)
    if len(reduce_axis) == 0:
        return err.mean()
    return err.mean(tuple(reduce_axis))


# ----------------------------------------------------------------------------------
# synthetic workloads shared by tests and bench
# ----------------------------------------------------------------------------------
def synth_inputs(batch: int, cfg: OracleConfig, seed: int = 123, scale: str = "unit",
                 dtype=torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
    """X1: unit-scale smooth motion  base~N(0,1)(B,1,V,3) + cumsum_t N(0,0.03^2); X2: 50 + 350*X1.
    target = x_last + N(0, 0.1^2) broadcast over the output frames (scaled likewise)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(batch, 1, cfg.joints, 3, generator=g)
    steps = 0.03 * torch.randn(batch, cfg.input_n, cfg.joints, 3, generator=g)
    x = base + steps.cumsum(1)
    tgt = x[:, -1:] + 0.1 * torch.randn(batch, cfg.output_n, cfg.joints, 3, generator=g)
    if scale == "mm":
        x, tgt = 50.0 + 350.0 * x, 50.0 + 350.0 * tgt
    elif scale != "unit":
        raise ValueError(scale)
    return x.to(dtype).contiguous(), tgt.to(dtype).contiguous()


def stress_init_(sd: Dict[str, torch.Tensor], seed: int = 7) -> Dict[str, torch.Tensor]:
    """W2 stress weights: conv/linear ~ N(0, 1/fan_in), BN gamma in [.75,1.25],
    beta ~ N(0,.1), running_mean ~ N(0,.1), running_var in [.75,1.25], PReLU slopes in [.05,.5].
    Deterministic in key order; modifies sd in place and returns it."""
    g = torch.Generator().manual_seed(seed)
    names = list(sd.keys())
    for k in names:
        t = sd[k]
        if k.endswith("num_batches_tracked"):
            continue
        if k.endswith("running_mean"):
            t.copy_(0.1 * torch.randn(t.shape, generator=g))
        elif k.endswith("running_var"):
            t.copy_(0.75 + 0.5 * torch.rand(t.shape, generator=g))
        elif k.endswith(".weight") and t.dim() == 1:
            base = k[: -len(".weight")]
            if base + ".running_mean" in sd:            # BatchNorm gamma
                t.copy_(0.75 + 0.5 * torch.rand(t.shape, generator=g))
            else:                                       # PReLU slope(s)
                t.copy_(0.05 + 0.45 * torch.rand(t.shape, generator=g))
        elif k.endswith(".bias"):
            base = k[: -len(".bias")]
            std = 0.1
            t.copy_(std * torch.randn(t.shape, generator=g))
            del base
        elif k.endswith(".weight") or k.endswith(".A"):
            fan_in = t[0].numel() if t.dim() > 1 else t.numel()
            t.copy_(torch.randn(t.shape, generator=g) / (fan_in ** 0.5))
        else:
            raise KeyError(f"stress_init_: unclassified entry {k} {tuple(t.shape)}")
    return sd
