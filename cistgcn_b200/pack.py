"""Weight packing for the C-ABI: eval-mode BatchNorm folding, k-major layouts, descriptor arrays.

The field indices of every descriptor are parsed from ``include/cistgcn_b200.h`` so that the
Python side and the kernels cannot drift apart.  Algebra follows SURVEY.md App. A:
``fold(bn) -> (s, b)`` with ``s = gamma / sqrt(var + eps)``, ``b = beta - mean * s``.
"""
from __future__ import annotations

import ctypes
import os
import re
from dataclasses import dataclass
from typing import Dict, List

import torch

_HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "cistgcn_b200.h")
BN_EPS = 1e-5


def _parse_header(path: str = _HEADER):
    src = open(path).read()
    src_nc = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    enums: Dict[str, int] = {}
    for body in re.findall(r"enum\s+\w+\s*\{(.*?)\}", src_nc, flags=re.S):
        nxt = 0
        for tok in body.split(","):
            tok = tok.strip()
            if not tok:
                continue
            if "=" in tok:
                name, val = (t.strip() for t in tok.split("="))
                nxt = int(val, 0)
            else:
                name = tok
            enums[name] = nxt
            nxt += 1
    defines = {k: int(v, 0) for k, v in re.findall(r"#define\s+(CISTGCN_\w+)\s+(\d+)", src_nc)}
    return enums, defines


F, DEFINES = _parse_header()          # F["CB_CI"] == 0 ...
MPAD = DEFINES["CISTGCN_MPAD"]
ABI_VERSION = DEFINES["CISTGCN_ABI_VERSION"]


def pad_up(n: int, m: int = MPAD) -> int:
    return (n + m - 1) // m * m


@dataclass
class ModelGeometry:
    input_n: int
    output_n: int
    joints: int
    in_chain: List[int]
    out_chain: List[int]
    in_interp: List[bool]
    out_interp: List[bool]
    n_fpn: int
    hidden_dim: int
    reduction: int
    feat_ch: int = 10

    @property
    def cmax(self) -> int:
        return max(self.in_chain + self.out_chain)


@dataclass
class PackedModel:
    blob: torch.Tensor            # float32, on the target device
    plan: List[int]
    plan_c: "ctypes.Array"        # (c_int32 * len(plan)), host memory
    geometry: ModelGeometry
    offsets: Dict[str, int]       # debug: entry name -> float offset
    fpn_tc: bool = True           # every FPN layer carries its tcgen05 operand image (else: FP32-FMA kernel)
    fpn_tc_reason: str = ""       # why not, when fpn_tc is False

    def block_desc(self, which: str, i: int) -> "ctypes.Array":
        g = self.geometry
        n_in = len(g.in_chain) - 1
        base = F["CP_HEADER_COUNT"]
        if which == "in":
            off = base + i * F["CB_COUNT"]
        else:
            off = base + n_in * F["CB_COUNT"] + g.n_fpn * F["CF_COUNT"] + F["CT_COUNT"] + i * F["CB_COUNT"]
        return (ctypes.c_int32 * F["CB_COUNT"])(*self.plan[off: off + F["CB_COUNT"]])

    def fpn_descs(self) -> "ctypes.Array":
        g = self.geometry
        off = F["CP_HEADER_COUNT"] + (len(g.in_chain) - 1) * F["CB_COUNT"]
        n = g.n_fpn * F["CF_COUNT"]
        return (ctypes.c_int32 * n)(*self.plan[off: off + n])

    def tail_desc(self) -> "ctypes.Array":
        g = self.geometry
        off = F["CP_HEADER_COUNT"] + (len(g.in_chain) - 1) * F["CB_COUNT"] + g.n_fpn * F["CF_COUNT"]
        return (ctypes.c_int32 * F["CT_COUNT"])(*self.plan[off: off + F["CT_COUNT"]])


class _Blob:
    def __init__(self):
        self.chunks: List[torch.Tensor] = []
        self.size = 0
        self.offsets: Dict[str, int] = {}

    def add_raw(self, name: str, words: torch.Tensor) -> int:
        """Appends raw 32-bit words (bf16 pairs of the tensor-core images); bit patterns are preserved."""
        words = words.detach().reshape(-1).cpu()
        assert words.dtype == torch.int32 and words.numel() % 4 == 0
        off = self.size
        self.chunks.append(words)
        self.size += words.numel()
        self.offsets[name] = off
        return off

    def add(self, name: str, t: torch.Tensor) -> int:
        t = t.detach().to(torch.float64).reshape(-1).cpu()
        off = self.size
        n = t.numel()
        padded = pad_up(max(n, 1), 4)                    # every entry starts 16-byte aligned
        if padded != n:
            t = torch.cat([t, torch.zeros(padded - n, dtype=torch.float64)])
        self.chunks.append(t)
        self.size += padded
        self.offsets[name] = off
        return off

    def finish(self, device) -> torch.Tensor:
        words = [c if c.dtype == torch.int32 else c.to(torch.float32).view(torch.int32) for c in self.chunks]
        return torch.cat(words).view(torch.float32).to(device).contiguous()


def host_state(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """state_dict -> fp64 host tensors with ONE device-to-host transfer per source device: the floating-point
    entries are flattened into a single buffer on their device (one concatenation kernel), copied once and
    split into views on the host.  (Per-tensor `.double().cpu()` cost ~1.2 K cast kernels + ~1.2 K synchronous
    copies per pack.)  `num_batches_tracked` counters are not needed for packing and are skipped."""
    out: Dict[str, torch.Tensor] = {}
    by_dev: Dict[torch.device, List[str]] = {}
    for k, t in sd.items():
        if not t.is_floating_point():
            continue
        by_dev.setdefault(t.device, []).append(k)
    for dev, keys in by_dev.items():
        ts = [sd[k].detach() for k in keys]
        if len({t.dtype for t in ts}) != 1:
            ts = [t.to(torch.float32) for t in ts]
        flat = torch.cat([t.reshape(-1) for t in ts]).cpu().to(torch.float64)
        off = 0
        for k, t in zip(keys, ts):
            n = t.numel()
            out[k] = flat[off: off + n].view(t.shape)
            off += n
    return out


def _fold(sd, p):
    g, b = sd[p + ".weight"], sd[p + ".bias"]
    m, v = sd[p + ".running_mean"], sd[p + ".running_var"]
    s = g / torch.sqrt(v + BN_EPS)
    return s, b - m * s


def _w(sd, k):
    return sd[k]


def _kmajor(W: torch.Tensor, pad: int = MPAD) -> torch.Tensor:
    """(M, K) -> [K][pad_up(M, pad)] zero padded."""
    M, K = W.shape
    out = torch.zeros(K, pad_up(M, pad), dtype=torch.float64)
    out[:, :M] = W.t()
    return out


def _padvec(v: torch.Tensor, pad: int) -> torch.Tensor:
    out = torch.zeros(pad_up(v.numel(), pad), dtype=torch.float64)
    out[: v.numel()] = v.reshape(-1)
    return out


def _pack_block(blob: _Blob, sd, p: str, ci: int, co: int, T: int, V: int, interp: bool, reduction: int,
                in_mode: int, in_strides, out_strides) -> List[int]:
    d = [0] * F["CB_COUNT"]
    ch, cg, hs = ci // 2, max(co // 2, 1), max(co // reduction, 1)
    has_res = int(ci != co)
    d[F["CB_CI"]], d[F["CB_CO"]], d[F["CB_T"]], d[F["CB_V"]] = ci, co, T, V
    d[F["CB_CH"]], d[F["CB_CG"]], d[F["CB_HS"]] = ch, cg, hs
    d[F["CB_HAS_RES"]], d[F["CB_INTERP"]], d[F["CB_IN_MODE"]] = has_res, int(interp), in_mode
    for i, n in enumerate(("SB", "SC", "ST", "SV")):
        d[F["CB_IN_" + n]] = int(in_strides[i])
        d[F["CB_OUT_" + n]] = int(out_strides[i])
    if interp and ch < 1:
        raise ValueError("Map2Adj needs at least 2 input channels")

    def put(field, name, t):
        d[F[field]] = blob.add(f"{p}:{name}", t)

    tc_mats = {}          # field -> (matrix (outputs, inputs), input row blocks) of the tensor-core channel mixes
    s, b = _fold(sd, p + ".global_norm")
    put("CB_GN_S", "gn_s", s)
    put("CB_GN_B", "gn_b", b)
    # ---- gates (CISTGCN.py:323-352)
    w0, b0, a0, w4, b4, a4, m0, mb0, ma0, m4 = [], [], [], [], [], [], [], [], [], []
    for k, mk in (("conv_s", "map_s"), ("conv_t", "map_t")):
        s, b = _fold(sd, f"{p}.{k}.1")
        w0.append(s[:, None] * _w(sd, f"{p}.{k}.0.weight").reshape(cg, ci * T))
        b0.append(b)
        a0.append(_w(sd, f"{p}.{k}.3.weight"))
        s, b = _fold(sd, f"{p}.{k}.5")
        w4.append(_kmajor(s[:, None] * _w(sd, f"{p}.{k}.4.weight").reshape(co, cg * V)))
        b4.append(b)
        a4.append(_w(sd, f"{p}.{k}.7.weight"))
        s, b = _fold(sd, f"{p}.{mk}.1")
        m0.append(_kmajor(s[:, None] * _w(sd, f"{p}.{mk}.0.weight")))
        mb0.append(b)
        ma0.append(_w(sd, f"{p}.{mk}.3.weight"))
        m4.append(_kmajor(_w(sd, f"{p}.{mk}.4.weight")))
    put("CB_G0_WT", "g0_wt", _kmajor(torch.cat(w0, 0)))
    put("CB_G0_B", "g0_b", torch.cat(b0))
    put("CB_R_G0_WT", "r_g0_wt", _kmajor(torch.cat(w0, 0), 32))     # reduce stage: lane = output channel
    put("CB_R_G0_B", "r_g0_b", _padvec(torch.cat(b0), 32))
    put("CB_G0_A", "g0_a", torch.cat(a0))
    put("CB_G4_WT", "g4_wt", torch.stack(w4))
    put("CB_G4_B", "g4_b", torch.stack(b4))
    put("CB_G4_A", "g4_a", torch.cat(a4))
    put("CB_M0_WT", "m0_wt", torch.stack(m0))
    put("CB_M0_B", "m0_b", torch.stack(mb0))
    put("CB_M0_A", "m0_a", torch.cat(ma0))
    put("CB_M4_WT", "m4_wt", torch.stack(m4))
    # ---- Map2Adj (CISTGCN.py:127-189) or static adjacency (:106-120)
    if interp:
        aw, ab, aa = [], [], []
        tc3w, tc3b, jc3w, jc3b = [], [], [], []
        for li, Lname in enumerate(("dsgn", "tsgn")):
            q = f"{p}.{Lname}.map_to_adj"
            sfx = "_S" if li == 0 else "_T"
            for br in ("time_compress", "joint_compress"):
                s, b = _fold(sd, f"{q}.{br}.1")
                aw.append(s[:, None] * _w(sd, f"{q}.{br}.0.weight").reshape(ch, ci))
                ab.append(b)
                aa.append(_w(sd, f"{q}.{br}.2.weight"))
            s, b = _fold(sd, f"{q}.time_compress.4")
            put("CB_TC3_WT" + sfx, Lname + ".tc3_wt", _kmajor(s[:, None] * _w(sd, f"{q}.time_compress.3.weight").reshape(ch, ch * T)))
            put("CB_TC3_B" + sfx, Lname + ".tc3_b", b)
            tc3w.append(s[:, None] * _w(sd, f"{q}.time_compress.3.weight").reshape(ch, ch * T))
            tc3b.append(b)
            put("CB_TC6_WT" + sfx, Lname + ".tc6_wt", _kmajor(_w(sd, f"{q}.time_compress.6.weight").reshape(T, ch)))
            s, b = _fold(sd, f"{q}.joint_compress.4")
            put("CB_JC3_WT" + sfx, Lname + ".jc3_wt", _kmajor(s[:, None] * _w(sd, f"{q}.joint_compress.3.weight").reshape(ch, ch * V)))
            put("CB_JC3_B" + sfx, Lname + ".jc3_b", b)
            jc3w.append(s[:, None] * _w(sd, f"{q}.joint_compress.3.weight").reshape(ch, ch * V))
            jc3b.append(b)
            put("CB_JC6_WT" + sfx, Lname + ".jc6_wt", _kmajor(_w(sd, f"{q}.joint_compress.6.weight").reshape(V, ch)))
            n = V if li == 0 else T
            s, b = _fold(sd, f"{q}.expansor.1")
            put("CB_E0_WT" + sfx, Lname + ".e0_wt", _kmajor(s[:, None] * _w(sd, f"{q}.expansor.0.weight").reshape(n, n)))
            put("CB_E0_B" + sfx, Lname + ".e0_b", b)
            put("CB_E0_A" + sfx, Lname + ".e0_a", _w(sd, f"{q}.expansor.3.weight"))
            put("CB_E4_WT" + sfx, Lname + ".e4_wt", _kmajor(_w(sd, f"{q}.expansor.4.weight").reshape(n, n)))
            put("CB_E4N_WT" + sfx, Lname + ".e4n_wt", _kmajor(_w(sd, f"{q}.expansor.4.weight").reshape(n, n).t()))
        put("CB_A0_WT", "a0_wt", _kmajor(torch.cat(aw, 0)))
        tc_mats["CB_TC_A0"] = (torch.cat(aw, 0), [ci])
        put("CB_A0_B", "a0_b", torch.cat(ab))
        put("CB_A0_A", "a0_a", torch.cat(aa))
        put("CB_R_A0_WT", "r_a0_wt", _kmajor(torch.cat(aw, 0), 32))
        put("CB_R_A0_B", "r_a0_b", _padvec(torch.cat(ab), 32))
        put("CB_R_TC3_WT", "r_tc3_wt", _kmajor(torch.cat(tc3w, 0), 32))
        put("CB_R_TC3_B", "r_tc3_b", _padvec(torch.cat(tc3b), 32))
        put("CB_R_JC3_WT", "r_jc3_wt", _kmajor(torch.cat(jc3w, 0), 32))
        put("CB_R_JC3_B", "r_jc3_b", _padvec(torch.cat(jc3b), 32))
    else:
        put("CB_ADJ_S", "adj_s", _w(sd, f"{p}.dsgn.gcn.A"))
        put("CB_ADJ_T", "adj_t", _w(sd, f"{p}.tsgn.gcn.A"))
    # ---- Domain_GCNN_layer channel mix + the prelu{1,2} gating stage (CISTGCN.py:229-247, 353-358)
    for li, (Lname, pk) in enumerate((("dsgn", "prelu1"), ("tsgn", "prelu2"))):
        sfx = "_S" if li == 0 else "_T"
        q = f"{p}.{Lname}"
        s, b = _fold(sd, f"{q}.tcn.1")
        W = s[:, None] * _w(sd, f"{q}.tcn.0.weight").reshape(co, ci)
        bias = s * _w(sd, f"{q}.tcn.0.bias") + b
        if has_res:
            sr, br = _fold(sd, f"{q}.residual.1")
            W = torch.cat([W, sr[:, None] * _w(sd, f"{q}.residual.0.weight").reshape(co, ci)], 1)
            bias = bias + sr * _w(sd, f"{q}.residual.0.bias") + br
        put("CB_TCN_WT" + sfx, Lname + ".tcn_wt", _kmajor(W))
        tc_mats["CB_TC_TCN" + sfx] = (W, [ci, ci] if has_res else [ci])
        put("CB_TCN_B" + sfx, Lname + ".tcn_b", bias)
        put("CB_TCN_A" + sfx, Lname + ".tcn_a", _w(sd, f"{q}.prelu.weight"))
        s, b = _fold(sd, f"{p}.{pk}.0")
        put("CB_P_S" + sfx, pk + ".s", s)
        put("CB_P_B" + sfx, pk + ".b", b)
        put("CB_P_A" + sfx, pk + ".a", _w(sd, f"{p}.{pk}.1.weight"))
    # ---- compressor + SE (CISTGCN.py:305-309, SE.py:24-41)
    s, b = _fold(sd, f"{p}.compressor.1")
    put("CB_CP_WT", "cp_wt", _kmajor(s[:, None] * _w(sd, f"{p}.compressor.0.weight").reshape(co, 2 * co)))
    tc_mats["CB_TC_CP"] = (s[:, None] * _w(sd, f"{p}.compressor.0.weight").reshape(co, 2 * co), [co, co])
    put("CB_CP_B", "cp_b", b)
    put("CB_CP_A", "cp_a", _w(sd, f"{p}.compressor.2.weight"))
    put("CB_SE1_WT", "se1_wt", _kmajor(_w(sd, f"{p}.compressor.3.excitation.0.weight")))
    put("CB_SE2_WT", "se2_wt", _kmajor(_w(sd, f"{p}.compressor.3.excitation.2.weight")))
    if has_res:
        s, b = _fold(sd, f"{p}.residual.1")
        put("CB_RS_WT", "rs_wt", _kmajor(s[:, None] * _w(sd, f"{p}.residual.0.weight").reshape(co, ci)))
        put("CB_RS_B", "rs_b", s * _w(sd, f"{p}.residual.0.bias") + b)
        tc_mats["CB_TC_RS"] = (s[:, None] * _w(sd, f"{p}.residual.0.weight").reshape(co, ci), [ci])
    # tensor-core images: only where K is worth a k-step and the fp16 copies cannot overflow
    if ci >= 16 and max(4 * ch, co) <= 64 and all(float(m.abs().max()) < TC_FP16_MAX for m, _ in tc_mats.values()):
        for field, (mat, segs) in tc_mats.items():
            d[F[field]] = blob.add_raw(f"{p}:{field.lower()}", tc_image(mat, segs))
    return d


def bf16_split3(w: torch.Tensor):
    """w (any float dtype) -> three bf16 tensors with fp32(w) == b1 + b2 + b3 up to 2^-24 relative."""
    w32 = w.to(torch.float32)
    b1 = w32.to(torch.bfloat16)
    r = w32 - b1.float()
    b2 = r.to(torch.bfloat16)
    b3 = (r - b2.float()).to(torch.bfloat16)
    return b1, b2, b3


TC_FP16_MAX = 6.0e4


def tc_slice(Wm: torch.Tensor, kc: int) -> torch.Tensor:
    """(cout <= 32, cin <= 8*kc) matrix -> int32 words of the K-major tcgen05 operand image
    [kc k-chunks][128 rows][8 x 16 bit] (csrc/fpn_tc.cuh): rows 32*term + cout hold the three bf16 terms of
    the weights (term = 0, 1, 2), rows 96 + cout their fp16 rounding (the partner of the activations' fp16
    remainder)."""
    co, ci = Wm.shape
    assert co <= 32 and ci <= 8 * kc
    full = torch.zeros(32, kc * 8, dtype=torch.float64)
    full[:co, :ci] = Wm
    img = torch.zeros(kc, 128, 8, dtype=torch.int16)
    for j, part in enumerate(bf16_split3(full)):
        img[:, 32 * j: 32 * j + 32, :] = part.view(torch.int16).reshape(32, kc, 8).permute(1, 0, 2)
    img[:, 96:, :] = full.to(torch.float32).to(torch.float16).view(torch.int16).reshape(32, kc, 8).permute(1, 0, 2)
    return img.contiguous().view(torch.int32).reshape(-1)


def tc_slice_bf16(Wm: torch.Tensor, kc: int) -> torch.Tensor:
    """(cout <= 32, cin <= 8*kc) matrix -> int32 words of the single-term image [kc k-chunks][32 rows][8 x bf16]."""
    co, ci = Wm.shape
    assert co <= 32 and ci <= 8 * kc
    full = torch.zeros(32, kc * 8, dtype=torch.float64)
    full[:co, :ci] = Wm
    img = full.to(torch.float32).to(torch.bfloat16).view(torch.int16).reshape(32, kc, 8).permute(1, 0, 2)
    return img.contiguous().view(torch.int32).reshape(-1)


def tc_image(Wm: torch.Tensor, segs) -> torch.Tensor:
    """(M outputs, sum(segs) inputs) matrix -> int32 words of the tcgen05 B-operand image used by the DSTD-GC channel
    mixes: [K/8 chunks][4*Np rows][8 x 16 bit], Np = pad16(M); rows j*Np + m = bf16 term j (j < 3), rows 3*Np + m = fp16(w).
    `segs` are the row blocks of the activation operand (e.g. [Co, Co] for cat(u1, u2)); each is padded to 16 channels
    (one k-step), because the kernel converts and multiplies one row block at a time."""
    M = Wm.shape[0]
    Np = pad_up(M, 16)
    kp = [pad_up(k, 16) for k in segs]
    K = sum(kp)
    full = torch.zeros(Np, K, dtype=torch.float64)
    src = dst = 0
    for k, kpad in zip(segs, kp):
        full[:M, dst: dst + k] = Wm[:, src: src + k]
        src += k
        dst += kpad
    kc = K // 8
    img = torch.zeros(kc, 4 * Np, 8, dtype=torch.int16)
    for j, part in enumerate(bf16_split3(full)):
        img[:, j * Np: (j + 1) * Np, :] = part.view(torch.int16).reshape(Np, kc, 8).permute(1, 0, 2)
    img[:, 3 * Np:, :] = full.to(torch.float32).to(torch.float16).view(torch.int16).reshape(Np, kc, 8).permute(1, 0, 2)
    return img.contiguous().view(torch.int32).reshape(-1)


TC_PRM_BIAS, TC_PRM_SLOPE, TC_PRM_OUT_A, TC_PRM_CPB, TC_PRM_WAVG, TC_PRM_FLOATS = 0, 96, 99, 100, 132, 132 + 32 * 32


def _pack_fpn_tc(blob: _Blob, sd, p: str, prelu_key: str, cin: int, cout: int, d: List[int]) -> str:
    """Tensor-core image of one FPN layer (only when it fits the kernel's 32-wide tiles).
    Returns "" or the reason the layer has no image (the chain then runs on the FP32-FMA kernel)."""
    if cout > 32 or cin > 32:
        return f"{p}: {cin}->{cout} channels exceed the 32-wide tensor-core tiles"
    folded = [_fold(sd, f"{p}.block{i}.1")[0][:, None, None, None] * _w(sd, f"{p}.block{i}.0.weight") for i in (1, 2, 3)]
    if max(float(t.abs().max()) for t in folded + [_w(sd, f"{p}.compress.weight")]) >= TC_FP16_MAX:
        return f"{p}: |weight| >= {TC_FP16_MAX:g} would overflow the fp16 operand copy"
    kc = pad_up(cin, 16) // 8
    words, words16, prm = [], [], torch.zeros(TC_PRM_FLOATS, dtype=torch.float64)
    Wc = _w(sd, f"{p}.compress.weight").reshape(cout, 3 * cout + cin)
    for i in (1, 2, 3):
        s, b = _fold(sd, f"{p}.block{i}.1")
        W = s[:, None, None, None] * _w(sd, f"{p}.block{i}.0.weight")                   # (cout, cin, 3, 3)
        for kh in range(3):
            for kw in range(3):
                words.append(tc_slice(W[:, :, kh, kw], kc))
                words16.append(tc_slice_bf16(W[:, :, kh, kw], kc))
        words.append(tc_slice(Wc[:, (i - 1) * cout: i * cout], 4))
        words16.append(tc_slice_bf16(Wc[:, (i - 1) * cout: i * cout], 4))
        prm[TC_PRM_BIAS + 32 * (i - 1): TC_PRM_BIAS + 32 * (i - 1) + cout] = s * _w(sd, f"{p}.block{i}.0.bias") + b
        prm[TC_PRM_SLOPE + i - 1] = _w(sd, f"{p}.block{i}.3.weight").reshape(-1)[0]
    prm[TC_PRM_OUT_A] = _w(sd, prelu_key + ".weight").reshape(-1)[0]
    prm[TC_PRM_CPB: TC_PRM_CPB + cout] = _w(sd, f"{p}.compress.bias")
    wavg = torch.zeros(32, 32, dtype=torch.float64)
    wavg[:cin, :cout] = Wc[:, 3 * cout:].t()
    prm[TC_PRM_WAVG:] = wavg.reshape(-1)
    d[F["CF_TC_KC"]] = kc
    d[F["CF_TC_W"]] = blob.add_raw(f"{p}:tc_w", torch.cat(words))
    d[F["CF_TC_W16"]] = blob.add_raw(f"{p}:tc_w16", torch.cat(words16))
    d[F["CF_TC_PRM"]] = blob.add(f"{p}:tc_prm", prm)
    return ""


def _pack_fpn(blob: _Blob, sd, p: str, prelu_key: str, cin: int, cout: int, resid: bool, notes: List[str]) -> List[int]:
    if cout % 5 != 0:
        raise ValueError("cistgcn_b200 FPN kernel needs output_n to be a multiple of 5")
    d = [0] * F["CF_COUNT"]
    d[F["CF_CIN"]], d[F["CF_COUT"]], d[F["CF_RESID"]] = cin, cout, int(resid)
    nt = cout // 5
    for i in (1, 2, 3):
        s, b = _fold(sd, f"{p}.block{i}.1")
        W = s[:, None, None, None] * _w(sd, f"{p}.block{i}.0.weight")                   # (cout, cin, 3, 3)
        lay = torch.zeros(cin, 3, nt, 16, dtype=torch.float64)
        # lay[c][kh][ot][kw*5 + m] = W[ot*5 + m][c][kh][kw]
        lay[..., :15] = W.reshape(nt, 5, cin, 3, 3).permute(2, 3, 0, 4, 1).reshape(cin, 3, nt, 15)
        d[F[f"CF_W_D{i}"]] = blob.add(f"{p}:w_d{i}", lay)
        d[F[f"CF_B_D{i}"]] = blob.add(f"{p}:b_d{i}", s * _w(sd, f"{p}.block{i}.0.bias") + b)
        d[F[f"CF_A_D{i}"]] = blob.add(f"{p}:a_d{i}", _w(sd, f"{p}.block{i}.3.weight"))
    Wc = _w(sd, f"{p}.compress.weight").reshape(cout, 3 * cout + cin)
    d[F["CF_CP_WT"]] = blob.add(f"{p}:cp_wt", _kmajor(Wc[:, : 3 * cout]))
    d[F["CF_CP_AVG_WT"]] = blob.add(f"{p}:cp_avg_wt", _kmajor(Wc[:, 3 * cout:]))
    d[F["CF_CP_B"]] = blob.add(f"{p}:cp_b", _w(sd, f"{p}.compress.bias"))
    d[F["CF_OUT_A"]] = blob.add(f"{p}:out_a", _w(sd, prelu_key + ".weight"))
    why = _pack_fpn_tc(blob, sd, p, prelu_key, cin, cout, d)
    if why:
        notes.append(why)
    return d


def _pack_tail(blob: _Blob, sd, g: ModelGeometry) -> List[int]:
    d = [0] * F["CT_COUNT"]
    To, V, hid = g.output_n, g.joints, g.hidden_dim
    seh1, seh2 = To // g.reduction, max(To // g.reduction, 1)
    if seh1 < 1:
        raise ValueError("SELayer1d hidden width would be 0 (output_n < reduction)")
    for k, v in (("CT_TIN", g.input_n), ("CT_TOUT", To), ("CT_V", V), ("CT_F", g.feat_ch), ("CT_HID", hid),
                 ("CT_SEH1", seh1), ("CT_SEH2", seh2)):
        d[F[k]] = v

    def put(field, name, t):
        d[F[field]] = blob.add("tail:" + name, t)

    s, b = _fold(sd, "dim_conversor.1")
    put("CT_DC0_WT", "dc0_wt", _kmajor(s[:, None] * _w(sd, "dim_conversor.0.weight").reshape(3, g.feat_ch)))
    put("CT_DC0_B", "dc0_b", b)
    put("CT_DC0_A", "dc0_a", _w(sd, "dim_conversor.2.weight"))
    put("CT_DC3_WT", "dc3_wt", _kmajor(_w(sd, "dim_conversor.3.weight").reshape(3, 3)))
    put("CT_DC3_A", "dc3_a", _w(sd, "dim_conversor.4.weight"))
    p = "context_layer"
    for i, tag in ((1, "C1"), (3, "C3")):
        s, b = _fold(sd, f"{p}.context_conv{i}.1")
        put(f"CT_{tag}_S", f"c{i}_s", s * _w(sd, f"{p}.context_conv{i}.0.weight").reshape(hid))
        put(f"CT_{tag}_B", f"c{i}_b", b)
        put(f"CT_{tag}_A", f"c{i}_a", _w(sd, f"{p}.context_conv{i}.2.weight"))
    s, b = _fold(sd, f"{p}.context_conv2.1")
    put("CT_C2_WT", "c2_wt", _kmajor(s[:, None] * _w(sd, f"{p}.context_conv2.0.weight").reshape(hid, To)))
    put("CT_C2_B", "c2_b", b)
    put("CT_C2_A", "c2_a", _w(sd, f"{p}.context_conv2.2.weight"))
    put("CT_MAP_WT", "map_wt", torch.stack([_kmajor(_w(sd, f"{p}.map{i}.0.weight")) for i in (1, 2, 3)]))
    put("CT_MAP_A", "map_a", torch.cat([_w(sd, f"{p}.map{i}.2.weight") for i in (1, 2, 3)]))
    s, b = _fold(sd, f"{p}.fmap_s.1")
    put("CT_FS_WT", "fs_wt", _kmajor(s[:, None] * _w(sd, f"{p}.fmap_s.0.weight")))
    put("CT_FS_B", "fs_b", b)
    s, b = _fold(sd, f"{p}.fmap_t.1")
    put("CT_FT_WT", "ft_wt", _kmajor(s[:, None] * _w(sd, f"{p}.fmap_t.0.weight")))
    put("CT_FT_B", "ft_b", b)
    s, b = _fold(sd, f"{p}.norm_map.1")
    put("CT_N0_WT", "n0_wt", _kmajor(s[:, None] * _w(sd, f"{p}.norm_map.0.weight").reshape(To, To)))
    put("CT_N0_B", "n0_b", b)
    put("CT_N0_A", "n0_a", _w(sd, f"{p}.norm_map.3.weight"))
    put("CT_NSE1_WT", "nse1_wt", _kmajor(_w(sd, f"{p}.norm_map.4.excitation.0.weight")))
    put("CT_NSE2_WT", "nse2_wt", _kmajor(_w(sd, f"{p}.norm_map.4.excitation.2.weight")))
    s, b = _fold(sd, f"{p}.norm_map.6")
    put("CT_N5_WT", "n5_wt", _kmajor(s[:, None] * _w(sd, f"{p}.norm_map.5.weight").reshape(To, To)))
    put("CT_N5_B", "n5_b", b)
    put("CT_N5_A", "n5_a", _w(sd, f"{p}.norm_map.8.weight"))
    s, b = _fold(sd, f"{p}.fconv.1")
    put("CT_FC0_S", "fc0_s", s * _w(sd, f"{p}.fconv.0.weight").reshape(3))
    put("CT_FC0_B", "fc0_b", b)
    put("CT_FC0_A", "fc0_a", _w(sd, f"{p}.fconv.2.weight"))
    s, b = _fold(sd, f"{p}.fconv.4")
    put("CT_FC3_WT", "fc3_wt", _kmajor(s[:, None] * _w(sd, f"{p}.fconv.3.weight").reshape(3, 3)))
    put("CT_FC3_B", "fc3_b", b)
    put("CT_FC3_A", "fc3_a", _w(sd, f"{p}.fconv.5.weight"))
    put("CT_SE1_WT", "se1_wt", _kmajor(_w(sd, f"{p}.SE.excitation.0.weight")))
    put("CT_SE2_WT", "se2_wt", _kmajor(_w(sd, f"{p}.SE.excitation.2.weight")))
    return d


def pack_state_dict(sd: Dict[str, torch.Tensor], g: ModelGeometry, device) -> PackedModel:
    """Reference-named state_dict -> (device blob, plan).  Called by CISTGCN.pack()."""
    sd = host_state(sd)                      # fp64 on the host, one transfer
    blob = _Blob()
    T, V, To, Fc = g.input_n, g.joints, g.output_n, g.feat_ch
    n_in, n_out = len(g.in_chain) - 1, len(g.out_chain) - 1
    if g.in_chain[-1] != Fc or g.in_chain[0] != Fc or g.out_chain[0] != 3 or g.out_chain[-1] != 3:
        raise ValueError("unsupported channel chain")
    if n_in > DEFINES["CISTGCN_MAX_BLOCKS"] or n_out > DEFINES["CISTGCN_MAX_BLOCKS"] or g.n_fpn > DEFINES["CISTGCN_MAX_FPN"]:
        raise ValueError("too many blocks for CISTGCN_MAX_BLOCKS / CISTGCN_MAX_FPN")
    plan: List[int] = [0] * F["CP_HEADER_COUNT"]
    TV = T * V
    for i in range(n_in):
        ci, co = g.in_chain[i], g.in_chain[i + 1]
        in_mode = 1 if i == 0 else 0
        in_str = (TV * 3, 0, 0, 0) if in_mode else (ci * TV, TV, V, 1)
        # last input block writes (T, F, V): frames become the FPN's channel axis (CISTGCN.py:582)
        out_str = (T * co * V, V, co * V, 1) if i == n_in - 1 else (co * TV, TV, V, 1)
        plan += _pack_block(blob, sd, f"st_gcnns.{i}", ci, co, T, V, g.in_interp[i], g.reduction,
                            in_mode, in_str, out_str)
    fpn_notes: List[str] = []
    for j in range(g.n_fpn):
        plan += _pack_fpn(blob, sd, f"txcnns.{j}", f"prelus.{j}", T if j == 0 else To, To, j > 0, fpn_notes)
    plan += _pack_tail(blob, sd, g)
    for i in range(n_out):
        ci, co = g.out_chain[i], g.out_chain[i + 1]
        # x7 / x8 live as (To, V, 3); the output block sees (C=3, "T"=V, "V"=To)  (CISTGCN.py:550-553, 592)
        perm = (To * V * 3, 1, 3, 3 * V)
        in_str = perm if i == 0 else (ci * V * To, V * To, To, 1)
        out_str = perm if i == n_out - 1 else (co * V * To, V * To, To, 1)
        plan += _pack_block(blob, sd, f"st_gcnns_o.{i}", ci, co, V, To, g.out_interp[i], g.reduction,
                            0, in_str, out_str)
    plan[F["CP_ABI"]] = ABI_VERSION
    plan[F["CP_TIN"]], plan[F["CP_TOUT"]], plan[F["CP_V"]] = T, To, V
    plan[F["CP_N_IN_BLOCKS"]], plan[F["CP_N_FPN"]], plan[F["CP_N_OUT_BLOCKS"]] = n_in, g.n_fpn, n_out
    plan[F["CP_CMAX"]] = g.cmax
    plan[F["CP_WEIGHT_FLOATS"]] = blob.size
    dev_blob = blob.finish(device)
    plan_c = (ctypes.c_int32 * len(plan))(*plan)
    if not fpn_notes and (V not in (22, 18) or To > 25 or T > 16 or g.n_fpn > 4):
        fpn_notes.append(f"shape (input_n={T}, output_n={To}, joints={V}, {g.n_fpn} layers) has no tcgen05 instantiation")
    return PackedModel(blob=dev_blob, plan=plan, plan_c=plan_c, geometry=g, offsets=blob.offsets,
                       fpn_tc=not fpn_notes, fpn_tc_reason="; ".join(fpn_notes))
