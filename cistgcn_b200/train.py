"""Differentiable CIST-GCN path: train-mode forward, backward, fused Adam, data-parallel step.

Reference semantics: environment/train.py:54-107 (model.train(); outputs = model(inputs); loss = mpjpe(target, out)
with reduce_axis=[]; optimizer.zero_grad(); loss.backward(); optimizer.step()), environment/utils.py:53-57 (Adam, L2
weight decay in the gradient), and -- with BatchNorm frozen -- the eval-mode input-gradient path of
environment/adversarial_attacks.py:184, 422, 495-511.

Train mode needs whole-batch BatchNorm statistics between any two layers (per replica: the reference has no SyncBN),
so this path runs layer by layer.  ``DiffGraph.forward`` replays the reference's layer sequence (CISTGCN.py:567-597 and
the sub-module forwards it calls) through the hand-written kernels of csrc/train_ops.cu (C-ABI:
include/cistgcn_b200_train.h) and records a tape; ``DiffGraph.backward`` walks the tape in reverse.  PyTorch only owns
the memory (torch.empty) and the streams; no ATen arithmetic is on the path.

Parameters and gradients live in two flat fp32 buffers (``FlatParams``): the module's Parameters are views into the
first, the backward kernels write every parameter's gradient straight into the second, so the data-parallel step is
ONE all-reduce over one buffer (NCCL on NVLink / NVSwitch) followed by ONE fused Adam launch.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional

import torch

from . import _cabi

_f = ctypes.c_float
_i64 = ctypes.c_int64


class ConvShape(ctypes.Structure):
    _fields_ = [("B", ctypes.c_int64)] + [(n, ctypes.c_int32) for n in
                                           ("Ci", "H", "W", "Co", "kh", "kw", "ph", "pw", "dh", "dw")]


def _strides(shape):
    s, acc = [], 1
    for d in reversed(shape):
        s.append(acc)
        acc *= d
    return list(reversed(s))


def _arr4(vals):
    return (ctypes.c_int64 * 4)(*[int(v) for v in vals])


class Var:
    """A tensor on the tape.  ``g`` is its gradient (None until a consumer produced one)."""
    __slots__ = ("t", "g", "shared", "needs")

    def __init__(self, t: torch.Tensor, needs: bool = True):
        self.t = t
        self.g: Optional[torch.Tensor] = None
        self.shared = False          # g aliases a tensor owned by someone else: copy before accumulating in place
        self.needs = needs


class FlatParams:
    """All parameters of a module as views into ONE flat fp32 buffer, and a flat gradient buffer of the same layout."""

    def __init__(self, module: torch.nn.Module):
        params = [(n, p) for n, p in module.named_parameters()]
        dev = params[0][1].device
        total = sum(p.numel() for _, p in params)
        self.flat = torch.empty(total, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(total, device=dev, dtype=torch.float32)
        self.offsets: Dict[str, tuple] = {}
        off = 0
        for n, p in params:
            k = p.numel()
            view = self.flat[off: off + k].view(p.shape)
            view.copy_(p.data)
            p.data = view                                 # the Parameter object stays, its storage moves into the buffer
            self.offsets[n] = (off, k, tuple(p.shape))
            off += k
        self.numel = total
        self.device = dev

    def grad_view(self, name: str) -> torch.Tensor:
        off, k, shape = self.offsets[name]
        return self.grad[off: off + k].view(shape)

    def param_view(self, name: str) -> torch.Tensor:
        off, k, shape = self.offsets[name]
        return self.flat[off: off + k].view(shape)


class DiffGraph:
    """One differentiable forward / backward of the CISTGCN module tree ``model`` (cistgcn_b200.model.CISTGCN).

    training=True : BatchNorm uses batch statistics and updates the running ones (momentum 0.1), dropout active (p from
                    learning_config.dropout), parameter gradients are produced.
    training=False: BatchNorm frozen (running statistics), no dropout; with ``input_grad`` the gradient w.r.t. x is produced
                    (and parameter gradients only if ``param_grads``).
    """

    BN_MOMENTUM = 0.1
    BN_EPS = 1e-5

    def __init__(self, model, flat: Optional[FlatParams] = None, lib=None):
        self.model = model
        self.lib = lib or _cabi.lib()
        _cabi.bind_train(self.lib)
        self.flat = flat
        self.params = dict(model.named_parameters())
        self.buffers = dict(model.named_buffers())
        self.tape: List = []
        self.training = True
        self.param_grads = True
        self.dropout_p = float(getattr(model, "dropout", 0.0))
        self.seed = 0x5EED
        self._drop_counter = 0
        self.grads: Dict[str, torch.Tensor] = {}
        self.launches = 0
        self._step_dev: Optional[torch.Tensor] = None      # device-side step counter mixed into every dropout seed (graph replays)
        self.grad_target: Optional[torch.Tensor] = None    # flat buffer the parameter gradients are written into (default: flat.grad)
        self._scratch: Optional[torch.Tensor] = None       # partial sums of the reductions (one stream: reused by every launch)

    # ------------------------------------------------------------------------------------------ plumbing
    def _stream(self, t):
        return torch.cuda.current_stream(t.device).cuda_stream if t.is_cuda else None

    def _ck(self, rc, what):
        self.launches += 1
        _cabi.check(rc, what, self.lib)

    def new(self, *shape, like: torch.Tensor, dtype=torch.float32):
        return torch.empty(*shape, device=like.device, dtype=dtype)

    def scratch(self, floats: int, like: torch.Tensor) -> torch.Tensor:
        if self._scratch is None or self._scratch.device != like.device or self._scratch.numel() < floats:
            self._scratch = torch.empty(max(int(floats), 1 << 20), device=like.device, dtype=torch.float32)
        return self._scratch

    def _gview(self, name):
        """Where the gradient of parameter ``name`` is written (flat gradient buffer when there is one)."""
        if not self.param_grads:
            return None
        if self.flat is not None:
            off, k, shape = self.flat.offsets[name]
            tgt = self.grad_target if self.grad_target is not None else self.flat.grad
            g = tgt[off: off + k].view(shape)
        else:
            g = torch.empty_like(self.params[name])
        self.grads[name] = g
        return g

    def acc(self, v: Var, g: torch.Tensor, fresh: bool = True):
        """Accumulate gradient g into v.  fresh: g is a new tensor owned by the caller (may be adopted and updated in place)."""
        if not v.needs:
            return
        if v.g is None:
            v.g, v.shared = g, not fresh
            return
        if v.shared:                                       # copy-on-write
            own = torch.empty_like(v.g)
            self._ck(self.lib.cistgcn_axpby(_f(1.0), v.g.data_ptr(), _f(0.0), own.data_ptr(), own.numel(), self._stream(own)), "axpby")
            v.g, v.shared = own, False
        self._ck(self.lib.cistgcn_axpby(_f(1.0), g.data_ptr(), _f(1.0), v.g.data_ptr(), g.numel(), self._stream(g)), "axpby")

    def copy4d(self, dst, dst_strides, src, src_strides, sizes, accumulate=False):
        self._ck(self.lib.cistgcn_copy4d(dst.data_ptr(), _arr4(dst_strides), src.data_ptr(), _arr4(src_strides), _arr4(sizes),
                                         int(accumulate), self._stream(dst)), "copy4d")

    # ------------------------------------------------------------------------------------------ layers
    def conv(self, x: Var, wname: str, bname: Optional[str] = None, pad=(0, 0), dil=(1, 1)) -> Var:
        w = self.params[wname]
        b = self.params[bname] if bname else None
        B, Ci, H, W = x.t.shape
        Co, _, kh, kw = (w.shape if w.dim() == 4 else (w.shape[0], w.shape[1], 1, 1))
        sh = ConvShape(B, Ci, H, W, Co, kh, kw, pad[0], pad[1], dil[0], dil[1])
        Ho, Wo = H + 2 * pad[0] - dil[0] * (kh - 1), W + 2 * pad[1] - dil[1] * (kw - 1)
        y = Var(self.new(B, Co, Ho, Wo, like=x.t))
        st = self._stream(x.t)
        self._ck(self.lib.cistgcn_conv2d_fwd(ctypes.byref(sh), x.t.data_ptr(), w.data_ptr(), b.data_ptr() if b is not None else None,
                                             y.t.data_ptr(), st), "conv2d_fwd")

        def bwd():
            if y.g is None:
                return
            if self.param_grads:
                gw = self._gview(wname)
                gb = self._gview(bname) if bname else None
                scr = self.scratch(self.lib.cistgcn_conv2d_bwd_weight_scratch_floats(ctypes.byref(sh)), x.t)
                self._ck(self.lib.cistgcn_conv2d_bwd_weight(ctypes.byref(sh), x.t.data_ptr(), y.g.data_ptr(), gw.data_ptr(),
                                                            gb.data_ptr() if gb is not None else None, scr.data_ptr(), st), "conv2d_bwd_weight")
            if x.needs:
                dx = torch.empty_like(x.t)
                self._ck(self.lib.cistgcn_conv2d_bwd_input(ctypes.byref(sh), y.g.data_ptr(), w.data_ptr(), dx.data_ptr(), st), "conv2d_bwd_input")
                self.acc(x, dx)
        self.tape.append(bwd)
        return y

    def linear(self, x: Var, wname: str) -> Var:
        """nn.Linear(bias=False) on (B, K, 1, 1)."""
        return self.conv(x, wname)

    def bn(self, x: Var, p: str) -> Var:
        B, C = x.t.shape[0], x.t.shape[1]
        HW = x.t.numel() // (B * C)
        gamma, beta = self.params[p + ".weight"], self.params[p + ".bias"]
        rm, rv = self.buffers[p + ".running_mean"], self.buffers[p + ".running_var"]
        y = Var(torch.empty_like(x.t))
        sm, si = self.new(C, like=x.t), self.new(C, like=x.t)
        st = self._stream(x.t)
        tr = int(self.training)
        scr = self.scratch(self.lib.cistgcn_bn_scratch_floats(C), x.t)
        self._ck(self.lib.cistgcn_bn_fwd(x.t.data_ptr(), gamma.data_ptr(), beta.data_ptr(), rm.data_ptr(), rv.data_ptr(), y.t.data_ptr(),
                                         sm.data_ptr(), si.data_ptr(), scr.data_ptr(), B, C, HW, tr, _f(self.BN_MOMENTUM), _f(self.BN_EPS), st),
                 "bn_fwd")
        if self.training:
            nbt = self.buffers.get(p + ".num_batches_tracked")
            if nbt is not None:
                self._nbt.append(nbt)

        def bwd():
            if y.g is None:
                return
            gg = self._gview(p + ".weight")
            gb = self._gview(p + ".bias")
            dx = torch.empty_like(x.t)
            self._ck(self.lib.cistgcn_bn_bwd(x.t.data_ptr(), y.g.data_ptr(), gamma.data_ptr(), sm.data_ptr(), si.data_ptr(), dx.data_ptr(),
                                             gg.data_ptr() if gg is not None else None, gb.data_ptr() if gb is not None else None,
                                             self.scratch(self.lib.cistgcn_bn_scratch_floats(C), x.t).data_ptr(), B, C, HW, tr, st), "bn_bwd")
            self.acc(x, dx)
        self.tape.append(bwd)
        return y

    def prelu(self, x: Var, p: str) -> Var:
        a = self.params[p + ".weight"]
        B, C = x.t.shape[0], x.t.shape[1]
        HW = x.t.numel() // (B * C)
        ns = a.numel()
        y = Var(torch.empty_like(x.t))
        st = self._stream(x.t)
        self._ck(self.lib.cistgcn_prelu_fwd(x.t.data_ptr(), a.data_ptr(), y.t.data_ptr(), B, C, HW, ns, st), "prelu_fwd")

        def bwd():
            if y.g is None:
                return
            ga = self._gview(p + ".weight")
            scratch = self.new(ns * 64, like=x.t, dtype=torch.float64) if ga is not None else None
            dx = torch.empty_like(x.t)
            self._ck(self.lib.cistgcn_prelu_bwd(x.t.data_ptr(), y.g.data_ptr(), a.data_ptr(), dx.data_ptr(),
                                                ga.data_ptr() if ga is not None else None,
                                                scratch.data_ptr() if scratch is not None else None, B, C, HW, ns, st), "prelu_bwd")
            self.acc(x, dx)
        self.tape.append(bwd)
        return y

    def act(self, x: Var, kind: int) -> Var:
        y = Var(torch.empty_like(x.t))
        st = self._stream(x.t)
        self._ck(self.lib.cistgcn_act_fwd(x.t.data_ptr(), y.t.data_ptr(), x.t.numel(), kind, st), "act_fwd")

        def bwd():
            if y.g is None:
                return
            dx = torch.empty_like(x.t)
            self._ck(self.lib.cistgcn_act_bwd(y.t.data_ptr(), y.g.data_ptr(), dx.data_ptr(), x.t.numel(), kind, st), "act_bwd")
            self.acc(x, dx)
        self.tape.append(bwd)
        return y

    def dropout(self, x: Var, p: Optional[float] = None) -> Var:
        p = self.dropout_p if p is None else p
        if not self.training or p <= 0.0:
            return x
        self._drop_counter += 1
        seed = (self.seed * 0x9E3779B1 + self._drop_counter) & 0xFFFFFFFFFFFFFFFF      # per layer; the step enters on the device
        y = Var(torch.empty_like(x.t))
        st = self._stream(x.t)
        step = self._step_dev.data_ptr() if self._step_dev is not None else None
        self._ck(self.lib.cistgcn_dropout(x.t.data_ptr(), y.t.data_ptr(), x.t.numel(), _f(p), ctypes.c_uint64(seed), step, st), "dropout")

        def bwd():
            if y.g is None:
                return
            dx = torch.empty_like(x.t)
            self._ck(self.lib.cistgcn_dropout(y.g.data_ptr(), dx.data_ptr(), x.t.numel(), _f(p), ctypes.c_uint64(seed), step, st), "dropout")
            self.acc(x, dx)
        self.tape.append(bwd)
        return y

    def view(self, x: Var, *shape) -> Var:
        y = Var(x.t.view(*shape), needs=x.needs)

        def bwd():
            if y.g is not None:
                self.acc(x, y.g.view(x.t.shape), fresh=False)
        self.tape.append(bwd)
        return y

    def permute(self, x: Var, dims) -> Var:
        src = x.t
        out_shape = [src.shape[d] for d in dims]
        y = Var(self.new(*out_shape, like=src), needs=x.needs)
        ss = _strides(src.shape)
        self.copy4d(y.t, _strides(out_shape), src, [ss[d] for d in dims], out_shape)

        def bwd():
            if y.g is None:
                return
            dx = torch.empty_like(src)
            # dx[src index] = dy[permuted index]: iterate over the output index space, scatter with the source strides
            self.copy4d(dx, [ss[d] for d in dims], y.g, _strides(out_shape), out_shape)
            self.acc(x, dx)
        self.tape.append(bwd)
        return y

    def cat(self, xs: List[Var]) -> Var:
        """torch.cat over the channel axis of (B, C_i, H, W) tensors."""
        B, _, H, W = xs[0].t.shape
        Cs = [v.t.shape[1] for v in xs]
        Ct = sum(Cs)
        y = Var(self.new(B, Ct, H, W, like=xs[0].t))
        off = 0
        for v, c in zip(xs, Cs):
            self.copy4d(y.t[:, off: off + c], [Ct * H * W, H * W, W, 1], v.t, _strides(v.t.shape), [B, c, H, W])
            off += c

        def bwd():
            if y.g is None:
                return
            o = 0
            for v, c in zip(xs, Cs):
                if v.needs:
                    dx = torch.empty_like(v.t)
                    self.copy4d(dx, _strides(v.t.shape), y.g[:, o: o + c], [Ct * H * W, H * W, W, 1], [B, c, H, W])
                    self.acc(v, dx)
                o += c
        self.tape.append(bwd)
        return y

    def add(self, a: Var, b: Var) -> Var:
        y = Var(torch.empty_like(a.t))
        st = self._stream(a.t)
        self._ck(self.lib.cistgcn_axpby(_f(1.0), a.t.data_ptr(), _f(0.0), y.t.data_ptr(), y.t.numel(), st), "axpby")
        self._ck(self.lib.cistgcn_axpby(_f(1.0), b.t.data_ptr(), _f(1.0), y.t.data_ptr(), y.t.numel(), st), "axpby")

        def bwd():
            if y.g is not None:
                self.acc(a, y.g, fresh=False)
                self.acc(b, y.g, fresh=False)
        self.tape.append(bwd)
        return y

    def gcn(self, x: Var, A: Var, domain: int, batched: bool, aname: Optional[str] = None) -> Var:
        B, C, T, V = x.t.shape
        y = Var(torch.empty_like(x.t))
        st = self._stream(x.t)
        self._ck(self.lib.cistgcn_gcn_fwd(x.t.data_ptr(), A.t.data_ptr(), y.t.data_ptr(), B, C, T, V, domain, int(batched), st), "gcn_fwd")

        def bwd():
            if y.g is None:
                return
            dx = torch.empty_like(x.t) if x.needs else None
            if batched:
                dA = torch.empty_like(A.t) if A.needs else None
            else:
                dA = self._gview(aname)
            self._ck(self.lib.cistgcn_gcn_bwd(x.t.data_ptr(), A.t.data_ptr(), y.g.data_ptr(), dx.data_ptr() if dx is not None else None,
                                              dA.data_ptr() if dA is not None else None, B, C, T, V, domain, int(batched), st), "gcn_bwd")
            if dx is not None:
                self.acc(x, dx)
            if batched and dA is not None:
                self.acc(A, dA)
        self.tape.append(bwd)
        return y

    def outer(self, dsp: Var, dseq: Var, T: int, V: int, domain: int) -> Var:
        B = dsp.t.shape[0]
        shape = (B, V, T, T) if domain == 0 else (B, T, V, V)
        y = Var(self.new(*shape, like=dsp.t))
        st = self._stream(dsp.t)
        self._ck(self.lib.cistgcn_outer_fwd(dsp.t.data_ptr(), dseq.t.data_ptr(), y.t.data_ptr(), B, T, V, domain, st), "outer_fwd")

        def bwd():
            if y.g is None:
                return
            d1, d2 = torch.empty_like(dsp.t), torch.empty_like(dseq.t)
            self._ck(self.lib.cistgcn_outer_bwd(dsp.t.data_ptr(), dseq.t.data_ptr(), y.g.data_ptr(), d1.data_ptr(), d2.data_ptr(),
                                                B, T, V, domain, st), "outer_bwd")
            self.acc(dsp, d1)
            self.acc(dseq, d2)
        self.tape.append(bwd)
        return y

    def stats(self, x: Var) -> Var:
        B, C, T, V = x.t.shape
        y = Var(self.new(B, 2 + 2 * T, 1, 1, like=x.t))
        st = self._stream(x.t)
        self._ck(self.lib.cistgcn_stats_fwd(x.t.data_ptr(), y.t.data_ptr(), B, C, T, V, st), "stats_fwd")

        def bwd():
            if y.g is None or not x.needs:
                return
            dx = torch.zeros_like(x.t)
            self._ck(self.lib.cistgcn_stats_bwd(x.t.data_ptr(), y.t.data_ptr(), y.g.data_ptr(), dx.data_ptr(), B, C, T, V, st), "stats_bwd")
            self.acc(x, dx)
        self.tape.append(bwd)
        return y

    def spatial_mean(self, x: Var) -> Var:
        B, C = x.t.shape[0], x.t.shape[1]
        HW = x.t.numel() // (B * C)
        y = Var(self.new(B, C, 1, 1, like=x.t))
        st = self._stream(x.t)
        self._ck(self.lib.cistgcn_spatial_mean_fwd(x.t.data_ptr(), y.t.data_ptr(), B, C, HW, st), "spatial_mean_fwd")

        def bwd():
            if y.g is None or not x.needs:
                return
            dx = torch.empty_like(x.t)
            self._ck(self.lib.cistgcn_spatial_mean_bwd(y.g.data_ptr(), dx.data_ptr(), B, C, HW, 0, st), "spatial_mean_bwd")
            self.acc(x, dx)
        self.tape.append(bwd)
        return y

    def broadcast_hw(self, m: Var, H: int, W: int) -> Var:
        """(B, C, 1, 1) -> (B, C, H, W)  (F.interpolate of a pooled 1x1 map, CISTGCN.py:76)."""
        B, C = m.t.shape[0], m.t.shape[1]
        y = Var(self.new(B, C, H, W, like=m.t))
        self.copy4d(y.t, _strides(y.t.shape), m.t, [C, 1, 0, 0], [B, C, H, W])
        st = self._stream(m.t)

        def bwd():
            if y.g is None:
                return
            dm = torch.empty_like(m.t)
            self._ck(self.lib.cistgcn_spatial_mean_fwd(y.g.data_ptr(), dm.data_ptr(), B, C, H * W, st), "spatial_mean_fwd")
            self._ck(self.lib.cistgcn_axpby(_f(float(H * W)), dm.data_ptr(), _f(0.0), dm.data_ptr(), dm.numel(), st), "axpby")
            self.acc(m, dm)
        self.tape.append(bwd)
        return y

    def scale(self, x: Var, g: Var) -> Var:
        """y = x * g[b, c] broadcast over the spatial axes (SE scaling; w[..., None, None] * x of CISTGCN.py:388)."""
        B, C = x.t.shape[0], x.t.shape[1]
        HW = x.t.numel() // (B * C)
        y = Var(torch.empty_like(x.t))
        st = self._stream(x.t)
        self._ck(self.lib.cistgcn_scale_fwd(x.t.data_ptr(), g.t.data_ptr(), y.t.data_ptr(), B, C, HW, st), "scale_fwd")

        def bwd():
            if y.g is None:
                return
            dx, dg = torch.empty_like(x.t), torch.empty_like(g.t)
            self._ck(self.lib.cistgcn_scale_bwd(x.t.data_ptr(), g.t.data_ptr(), y.g.data_ptr(), dx.data_ptr(), dg.data_ptr(), B, C, HW, st), "scale_bwd")
            self.acc(x, dx)
            self.acc(g, dg)
        self.tape.append(bwd)
        return y

    def rowmax(self, x: Var, R: int, N: int) -> Var:
        y = Var(self.new(R, like=x.t))
        idx = self.new(R, like=x.t, dtype=torch.int32)
        st = self._stream(x.t)
        self._ck(self.lib.cistgcn_rowmax_fwd(x.t.data_ptr(), y.t.data_ptr(), idx.data_ptr(), R, N, st), "rowmax_fwd")

        def bwd():
            if y.g is None or not x.needs:
                return
            dx = torch.zeros_like(x.t)
            self._ck(self.lib.cistgcn_rowmax_bwd(y.g.data_ptr(), idx.data_ptr(), dx.data_ptr(), R, N, st), "rowmax_bwd")
            self.acc(x, dx)
        self.tape.append(bwd)
        return y

    def cumsum1(self, x: Var) -> Var:
        B, L = x.t.shape[0], x.t.shape[1]
        N = x.t.numel() // (B * L)
        y = Var(torch.empty_like(x.t))
        st = self._stream(x.t)
        self._ck(self.lib.cistgcn_cumsum(x.t.data_ptr(), y.t.data_ptr(), B, L, N, 0, st), "cumsum")

        def bwd():
            if y.g is None:
                return
            dx = torch.empty_like(x.t)
            self._ck(self.lib.cistgcn_cumsum(y.g.data_ptr(), dx.data_ptr(), B, L, N, 1, st), "cumsum")
            self.acc(x, dx)
        self.tape.append(bwd)
        return y

    def sum_axis1(self, x: Var) -> Var:
        """(B, L, N) -> (B, N) sum over axis 1 (only as a backward helper's forward; see bcast_axis1)."""
        raise NotImplementedError

    def bcast_axis1(self, m: Var, L: int) -> Var:
        """(B, N) -> (B, L, N), repeated along a new axis 1.  Backward: sum over that axis."""
        B, N = m.t.shape[0], m.t.numel() // m.t.shape[0]
        y = Var(self.new(B, L, N, like=m.t), needs=m.needs)
        self.copy4d(y.t, [L * N, N, 1, 0], m.t, [N, 0, 1, 0], [B, L, N, 1])
        st = self._stream(m.t)

        def bwd():
            if y.g is None or not m.needs:
                return
            tmp = self.new(B, N, L, like=m.t)                                     # (B, N, L): reduce the last axis
            self.copy4d(tmp, [N * L, L, 1, 0], y.g, [L * N, 1, N, 0], [B, N, L, 1])
            dm = self.new(*m.t.shape, like=m.t)
            self._ck(self.lib.cistgcn_spatial_mean_fwd(tmp.data_ptr(), dm.data_ptr(), B, N, L, st), "spatial_mean_fwd")
            self._ck(self.lib.cistgcn_axpby(_f(float(L)), dm.data_ptr(), _f(0.0), dm.data_ptr(), dm.numel(), st), "axpby")
            self.acc(m, dm)
        self.tape.append(bwd)
        return y

    def features(self, x: Var) -> Var:
        B, T, V, _ = x.t.shape
        y = Var(self.new(B, 10, T, V, like=x.t))
        st = self._stream(x.t)
        self._ck(self.lib.cistgcn_features_fwd(x.t.data_ptr(), y.t.data_ptr(), B, T, V, st), "features_fwd")

        def bwd():
            if y.g is None or not x.needs:
                return
            dx = torch.empty_like(x.t)
            self._ck(self.lib.cistgcn_features_bwd(x.t.data_ptr(), y.g.data_ptr(), dx.data_ptr(), B, T, V, st), "features_bwd")
            self.acc(x, dx)
        self.tape.append(bwd)
        return y

    # ------------------------------------------------------------------------------------------ sub-modules
    def se(self, x: Var, p: str) -> Var:
        """SELayer1d / SELayer2d (models/layers/SE.py:16-20, 37-41)."""
        m = self.spatial_mean(x)
        h = self.act(self.linear(m, p + ".excitation.0.weight"), 0)
        s = self.act(self.linear(h, p + ".excitation.2.weight"), 1)
        return self.scale(x, s)

    def map2adj(self, x: Var, p: str, T: int, V: int, domain: int) -> Var:
        """Map2Adj.forward (CISTGCN.py:183-189)."""
        B = x.t.shape[0]
        a = self.prelu(self.bn(self.conv(x, p + ".time_compress.0.weight"), p + ".time_compress.1"), p + ".time_compress.2")
        a = self.dropout(self.bn(self.conv(a, p + ".time_compress.3.weight"), p + ".time_compress.4"))
        dim_seq = self.conv(a, p + ".time_compress.6.weight")                      # (B, T, 1, V)
        g = self.prelu(self.bn(self.conv(x, p + ".joint_compress.0.weight"), p + ".joint_compress.1"), p + ".joint_compress.2")
        g = self.dropout(self.bn(self.conv(g, p + ".joint_compress.3.weight"), p + ".joint_compress.4"))
        dim_space = self.conv(g, p + ".joint_compress.6.weight")                   # (B, V, T, 1)
        o = self.outer(self.view(dim_space, B, V, T), self.view(dim_seq, B, T, V), T, V, domain)
        e = self.prelu(self.dropout(self.bn(self.conv(o, p + ".expansor.0.weight"), p + ".expansor.1")), p + ".expansor.3")
        return self.conv(e, p + ".expansor.4.weight")

    def domain_layer(self, x: Var, p: str, T: int, V: int, domain: int, interp: bool, has_res: bool) -> Var:
        """Domain_GCNN_layer.forward (CISTGCN.py:259-269)."""
        res = self.bn(self.conv(x, p + ".residual.0.weight", p + ".residual.0.bias"), p + ".residual.1") if has_res else x
        if interp:
            adj = self.map2adj(x, p + ".map_to_adj", T, V, domain)
            self.taps[p + ".Adj"] = adj.t
            g = self.gcn(x, adj, domain, True)
        else:
            A = Var(self.params[p + ".gcn.A"], needs=False)
            g = self.gcn(x, A, domain, False, aname=p + ".gcn.A")
        y = self.dropout(self.bn(self.conv(g, p + ".tcn.0.weight", p + ".tcn.0.bias"), p + ".tcn.1"))
        return self.prelu(self.add(y, res), p + ".prelu")

    def dstd_gc(self, x: Var, p: str, ci: int, co: int, T: int, V: int, interp: bool) -> Var:
        """DSTD_GC.forward (CISTGCN.py:373-390)."""
        B = x.t.shape[0]
        has_res = ci != co
        xn = self.bn(x, p + ".global_norm")
        st = self.stats(xn)                                                       # computed twice in the reference, same values
        ws = []
        for ck, mk in (("conv_s", "map_s"), ("conv_t", "map_t")):
            h = self.conv(xn, f"{p}.{ck}.0.weight")
            h = self.prelu(self.dropout(self.bn(h, f"{p}.{ck}.1")), f"{p}.{ck}.3")
            h = self.conv(h, f"{p}.{ck}.4.weight")
            h = self.prelu(self.dropout(self.bn(h, f"{p}.{ck}.5")), f"{p}.{ck}.7")
            w = self.cat([self.view(h, B, co, 1, 1), st])
            z = self.prelu(self.dropout(self.bn(self.linear(w, f"{p}.{mk}.0.weight"), f"{p}.{mk}.1")), f"{p}.{mk}.3")
            ws.append(self.linear(z, f"{p}.{mk}.4.weight"))
        self.taps[p + ".w1"], self.taps[p + ".w2"] = ws[0].t, ws[1].t
        x1 = self.domain_layer(xn, p + ".dsgn", T, V, 0, interp, has_res)
        x2 = self.domain_layer(xn, p + ".tsgn", T, V, 1, interp, has_res)
        u1 = self.prelu(self.bn(self.scale(x1, ws[0]), p + ".prelu1.0"), p + ".prelu1.1")
        u2 = self.prelu(self.bn(self.scale(x2, ws[1]), p + ".prelu2.0"), p + ".prelu2.1")
        c = self.conv(self.cat([u1, u2]), p + ".compressor.0.weight")
        c = self.prelu(self.bn(c, p + ".compressor.1"), p + ".compressor.2")
        c = self.se(c, p + ".compressor.3")
        res = self.bn(self.conv(xn, p + ".residual.0.weight", p + ".residual.0.bias"), p + ".residual.1") if has_res else xn
        return self.add(c, res)

    def fpn(self, x: Var, p: str) -> Var:
        """FPN.forward (CISTGCN.py:74-79); dropout p = 0 there (:533)."""
        _, _, H, W = x.t.shape
        outs = []
        for i in (1, 2, 3):
            y = self.conv(x, f"{p}.block{i}.0.weight", f"{p}.block{i}.0.bias", pad=(i, i), dil=(i, i))
            outs.append(self.prelu(self.bn(y, f"{p}.block{i}.1"), f"{p}.block{i}.3"))
        outs.append(self.broadcast_hw(self.spatial_mean(x), H, W))
        return self.conv(self.cat(outs), p + ".compress.weight", p + ".compress.bias")

    def context_layer(self, z: Var, To: int, V: int) -> Var:
        """ContextLayer.forward (CISTGCN.py:463-475) on z (B, 1, To, 3V)."""
        p = "context_layer"
        B = z.t.shape[0]
        H = self.params[p + ".context_conv1.0.weight"].shape[0]
        cv = lambda i: self.prelu(self.bn(self.conv(z, f"{p}.context_conv{i}.0.weight"), f"{p}.context_conv{i}.1"), f"{p}.context_conv{i}.2")
        c1 = cv(1)
        y1 = self.rowmax(c1, B * H, To * 3 * V)
        c2 = cv(2)                                                                 # (B, H, 1, 3V)
        y2 = self.rowmax(c2, B * H, 3 * V)
        ym = self.spatial_mean(cv(3))
        ms = []
        for i, y in ((1, self.view(y1, B, H, 1, 1)), (2, self.view(y2, B, H, 1, 1)), (3, ym)):
            ms.append(self.prelu(self.dropout(self.linear(y, f"{p}.map{i}.0.weight")), f"{p}.map{i}.2"))
        y = self.cat(ms)                                                           # (B, 3*To, 1, 1)
        joints = self.dropout(self.bn(self.linear(y, p + ".fmap_s.0.weight"), p + ".fmap_s.1"))      # (B, V, 1, 1)
        disp = self.dropout(self.bn(self.linear(y, p + ".fmap_t.0.weight"), p + ".fmap_t.1"))        # (B, To, 1, 1)
        self.taps[p + ".joints"], self.taps[p + ".displacements"] = joints.t.view(B, V), disp.t.view(B, To)
        jb = self.view(self.bcast_axis1(self.view(joints, B, V), To), B, To, V, 1)                   # joints repeated over the frames
        sj = self.scale(jb, disp)                                                  # bmm(disp[:, :, None], joints[:, None, :])  (:471)
        n = self.conv(sj, p + ".norm_map.0.weight")                                # Conv1d k=1 over the frame axis
        n = self.prelu(self.dropout(self.bn(n, p + ".norm_map.1")), p + ".norm_map.3")
        n = self.se(n, p + ".norm_map.4")
        n = self.conv(n, p + ".norm_map.5.weight")
        sjn = self.prelu(self.dropout(self.bn(n, p + ".norm_map.6")), p + ".norm_map.8")             # (B, To, V, 1)
        self.taps[p + ".seq_joints_n"] = sjn.t.view(B, To, V)
        f = self.view(sjn, B, 1, To, V)
        f = self.prelu(self.bn(self.conv(f, p + ".fconv.0.weight"), p + ".fconv.1"), p + ".fconv.2")
        sjd = self.prelu(self.bn(self.conv(f, p + ".fconv.3.weight"), p + ".fconv.4"), p + ".fconv.5")   # (B, 3, To, V)
        self.taps[p + ".seq_joints_dims"] = sjd.t
        return self.se(self.permute(sjd, (0, 2, 3, 1)), p + ".SE")                 # (B, To, V, 3)

    # ------------------------------------------------------------------------------------------ whole model
    def forward(self, x: torch.Tensor, training: bool = True, input_grad: bool = False, param_grads: Optional[bool] = None) -> torch.Tensor:
        """pred (B, output_n, V, 3).  Records the tape for ``backward``."""
        m = self.model
        self.tape, self.grads, self.taps, self._nbt = [], {}, {}, []
        self.training = bool(training)
        self.param_grads = bool(training) if param_grads is None else bool(param_grads)
        self._drop_counter = 0
        self.launches = 0
        if self.training and self.dropout_p > 0.0:
            # fresh masks every forward: a device-side counter (bumped by a kernel, so a captured graph advances it too)
            if self._step_dev is None or self._step_dev.device != x.device:
                self._step_dev = torch.zeros(1, device=x.device, dtype=torch.int64)
            self._ck(self.lib.cistgcn_counter_bump(self._step_dev.data_ptr(), self._stream(x)), "counter_bump")
        B, T, V, To = x.shape[0], m.n_input, m.n_joints, m.n_output
        xv = Var(x.contiguous(), needs=input_grad)
        self._x = xv
        h = self.features(xv)
        chain, interp = m._in_chain, m._in_interp
        for i in range(len(chain) - 1):
            h = self.dstd_gc(h, f"st_gcnns.{i}", chain[i], chain[i + 1], T, V, interp[i])
        x6 = self.permute(h, (0, 2, 1, 3))                                         # (B, T, 10, V)  (:582)
        x6 = self.prelu(self.fpn(x6, "txcnns.0"), "prelus.0")
        for i in range(1, m.n_txcnn_layers):
            x6 = self.add(self.prelu(self.fpn(x6, f"txcnns.{i}"), f"prelus.{i}"), x6)
        d = self.permute(x6, (0, 2, 1, 3))                                         # (B, 10, To, V)
        d = self.prelu(self.bn(self.conv(d, "dim_conversor.0.weight"), "dim_conversor.1"), "dim_conversor.2")
        d = self.prelu(self.conv(d, "dim_conversor.3.weight"), "dim_conversor.4")  # (B, 3, To, V), 3 slopes
        x7 = self.cumsum1(self.permute(d, (0, 2, 3, 1)))                           # (B, To, V, 3)  (:588-589)
        act = self.context_layer(self.view(x7, B, 1, To, 3 * V), To, V)
        x8 = self.permute(x7, (0, 3, 2, 1))                                        # (B, 3, V, To)  (:592)
        ochain, ointerp = m._out_chain, m._out_interp
        for i in range(len(ochain) - 1):
            x8 = self.dstd_gc(x8, f"st_gcnns_o.{i}", ochain[i], ochain[i + 1], V, To, ointerp[i])
        x9 = self.add(self.permute(x8, (0, 3, 2, 1)), act)                         # (:595)
        last = Var(xv.t[:, -1].contiguous(), needs=input_grad)                     # x[:, -1:]  (:597)
        self._last = last if input_grad else None                                  # d pred / d x[:, -1] = sum over the output frames
        lastb = self.bcast_axis1(self.view(last, B, V * 3), To)
        pred = self.add(x9, self.view(lastb, B, To, V, 3))
        self._pred = pred
        if self.training and self._nbt:
            torch._foreach_add_(self._nbt, 1)                                      # num_batches_tracked (bookkeeping only, one launch)
        return pred.t

    def backward(self, dpred: torch.Tensor) -> Optional[torch.Tensor]:
        """Runs the tape in reverse from d(loss)/d(pred).  Parameter gradients land in ``self.grads`` (views of the flat
        gradient buffer when the graph owns one); returns d(loss)/d(x) when the forward asked for it."""
        self._pred.g, self._pred.shared = dpred.contiguous(), True
        for fn in reversed(self.tape):
            fn()
        self.tape = []
        if self._x.needs:
            dx = self._x.g if self._x.g is not None else torch.zeros_like(self._x.t)
            if getattr(self, "_last", None) is not None and self._last.g is not None:
                if self._x.shared:
                    dx = dx.clone()
                B, T = dx.shape[0], dx.shape[1]
                n = dx[0, -1].numel()
                self.copy4d(dx[:, -1], [T * n, 1, 0, 0], self._last.g, [n, 1, 0, 0], [B, n, 1, 1], accumulate=True)
            return dx
        return None

    # ------------------------------------------------------------------------------------------ loss
    def mpjpe_loss(self, pred: torch.Tensor, target: torch.Tensor):
        """losses.mpjpe(target, pred) with reduce_axis=[] (environment/train.py:76): returns (loss 0-dim float32 tensor on
        the device, d loss / d pred)."""
        B, To, V, _ = pred.shape
        sums = torch.zeros(To, device=pred.device, dtype=torch.float64)
        st = self._stream(pred)
        self._ck(self.lib.cistgcn_mpjpe_f32(pred.data_ptr(), target.data_ptr(), B, To, V, None, sums.data_ptr(), st), "mpjpe_f32")
        n = B * To * V
        dpred = torch.empty_like(pred)
        self._ck(self.lib.cistgcn_mpjpe_bwd(pred.data_ptr(), target.data_ptr(), dpred.data_ptr(), n, _f(1.0 / n), st), "mpjpe_bwd")
        return sums, dpred


class Trainer:
    """The reference's training step (environment/train.py:54-107 with environment/utils.py:53-57) on one GPU, or
    data-parallel over the ranks of ``torch.distributed`` (local BatchNorm statistics like the reference, ONE all-reduce of
    the flat gradient buffer per step, identical Adam on every replica)."""

    def __init__(self, model, lr: float = 0.01, weight_decay: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, lib=None,
                 cuda_graph: bool = False):
        self.model = model
        self.flat = FlatParams(model)
        self.graph = DiffGraph(model, self.flat, lib)
        self.lib = self.graph.lib
        self.lr, self.wd, self.betas, self.eps = float(lr), float(weight_decay), betas, float(eps)
        self.m = torch.zeros_like(self.flat.flat)
        self.v = torch.zeros_like(self.flat.flat)
        self.step_count = 0
        self.allreduce_ms = None
        self.use_cuda_graph = bool(cuda_graph)
        self._cg = None                                    # (torch.cuda.CUDAGraph, static x, static target, static sums, launches)
        self._eager_steps = 0

    def _fwd_bwd(self, x, target):
        g = self.graph
        pred = g.forward(x, training=True)
        sums, dpred = g.mpjpe_loss(pred, target)
        g.backward(dpred)
        return sums

    def _fwd_bwd_graphed(self, x, target):
        """The ~1 900 launches of forward + loss + backward as ONE CUDA graph (train_h36m.yaml's batch of 128 is launch-bound:
        19 us of host work per launch against 2-3 us of kernel).  The first step runs eagerly (it also serves as the warm-up
        the capture needs), the second is captured and replayed, later ones are replays.  Inputs are copied into static
        buffers; dropout masks change per replay through the device-side step counter; the gradient buffer, the running
        statistics and num_batches_tracked are written in place by the graph's kernels."""
        if self._cg is not None and (self._cg[1].shape != x.shape or self._cg[1].device != x.device):
            self._cg, self._eager_steps = None, 0          # new batch shape: capture again
        if self._cg is None:
            if self._eager_steps < 1:
                self._eager_steps += 1
                return self._fwd_bwd(x, target)
            sx, st = torch.empty_like(x), torch.empty_like(target)
            sx.copy_(x)
            st.copy_(target)
            torch.cuda.synchronize(x.device)
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg):
                ssums = self._fwd_bwd(sx, st)
            self._cg = (cg, sx, st, ssums, self.graph.launches)
            cg.replay()
            return ssums.clone()
        cg, sx, st, ssums, launches = self._cg
        sx.copy_(x, non_blocking=True)
        st.copy_(target, non_blocking=True)
        cg.replay()
        self.graph.launches = launches
        return ssums.clone()

    def step(self, x: torch.Tensor, target: torch.Tensor, group=None, timing=None):
        """One step; returns the per-frame error sums of this rank's batch (float64 (output_n,), on the device):
        loss = sums.sum() / (B * T * V), read it lazily to avoid a sync per step."""
        import torch.distributed as dist
        g = self.graph
        self.model.train()
        sums = self._fwd_bwd_graphed(x, target) if (self.use_cuda_graph and x.is_cuda) else self._fwd_bwd(x, target)
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        if timing is not None:
            timing[0].record()
        if world > 1:
            dist.all_reduce(self.flat.grad, op=dist.ReduceOp.SUM, group=group)     # the only collective: one flat buffer
        if timing is not None:
            timing[1].record()
        self.step_count += 1
        st = g._stream(self.flat.flat)
        _cabi.check(self.lib.cistgcn_adam_step(self.flat.flat.data_ptr(), self.flat.grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                               self.flat.numel, _f(self.lr), _f(self.betas[0]), _f(self.betas[1]), _f(self.eps),
                                               _f(self.wd), self.step_count, _f(1.0 / world), st), "adam_step", self.lib)
        self.model._invalidate_pack()                      # the eval-mode packed weights are stale now
        return sums
