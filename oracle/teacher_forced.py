"""Teacher-forced check of one UNet_Nested training step.  TEST INFRASTRUCTURE ONLY (see oracle/unetpp_oracle.py).

A network whose activations and activation gradients are stored in bf16 is chaotic at the level of single
roundings: a 1e-7 difference in the accumulation order of one convolution flips a handful of bf16 roundings,
each flip perturbs ~100 values of the next layer by a fraction of an ulp, and after four layers a quarter of
all stored values differ by one ulp (measured: oracle/bf16_emulation.py in fp32 against itself in fp64).  An
end-to-end comparison of parameter gradients therefore cannot be much tighter than the bf16 noise floor,
whoever computes them — which hides wiring errors of the same size.

This module removes the chaos instead of bounding it: every tensor the CUDA step STORED (its ``TrainState``:
pre-BatchNorm outputs, activations, pooled copies, upsampled tensors, every activation gradient, BatchNorm
statistics, heat maps) is recomputed here from the step's OWN stored inputs of that one operation — the
reference arithmetic of models/unet.py:121-300 and of its autograd backward, restated with
torch.nn.functional / torch.nn.grad in fp64 — and rounded once.  A stored bf16 tensor must then agree to one
unit in the last place (fp32 accumulation order can flip a rounding, nothing more), and each of the 74
fp32 parameter gradients (sums of exact bf16 x bf16 products) to ~1e-4.  A wrong slice, tap, mask, scale or
consumer anywhere in the DAG shows up as an O(1) mismatch of exactly the tensor that kernel wrote.

Only the default constructor flags (is_deconv=True, is_batchnorm=True) are restated here.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F
from torch.nn.grad import conv2d_input, conv2d_weight

from . import unetpp_oracle as O

DEC = {"up_concat01": ("X10", ("X00",), 0), "up_concat11": ("X20", ("X10",), 1), "up_concat21": ("X30", ("X20",), 2),
       "up_concat02": ("X11", ("X00", "X01"), 0), "up_concat12": ("X21", ("X10", "X11"), 1), "up_concat03": ("X12", ("X00", "X01", "X02"), 0)}
HEAD_NODE = {"final_1": "X01", "final_2": "X02", "final_3": "X03"}


def _r(t: torch.Tensor) -> torch.Tensor:
    """One bf16 store (round to nearest even)."""
    return t.to(torch.bfloat16).to(t.dtype)


class Report:
    def __init__(self):
        self.rows: List[dict] = []

    def bf16(self, name: str, got: torch.Tensor, ref: torch.Tensor):
        """Stored bf16 tensor ``got`` against the fp64 recomputation ``ref`` (not yet rounded)."""
        got, ref = got.detach(), ref.detach()
        refr = _r(ref)
        scale = float(ref.abs().max()) + 1e-30
        d = (got - refr).abs()
        # one unit in the last place of a bf16 value v is at most 2^-7 |v|; values that cancel to (almost) nothing get an
        # absolute allowance of 2^-16 of the tensor's range (fp32 accumulation noise of the producing kernel)
        ulp = (2.0 ** -7) * torch.maximum(refr.abs(), got.abs()) + (2.0 ** -16) * scale
        self.rows.append(dict(name=name, kind="bf16", max_rel=float(d.max()) / scale, frac_diff=float((d > 0).double().mean()),
                              frac_gt_ulp=float((d > ulp).double().mean()), scale=scale))

    def f32(self, name: str, got: torch.Tensor, ref: torch.Tensor):
        got, ref = got.detach(), ref.detach()
        scale = float(ref.abs().max()) + 1e-30
        d = (got.to(ref.dtype) - ref).abs()
        self.rows.append(dict(name=name, kind="f32", max_rel=float(d.max()) / scale, scale=scale))

    def zero(self, name: str, got: torch.Tensor, scale: float):
        self.rows.append(dict(name=name, kind="zero", max_rel=float(got.abs().max()) / (scale + 1e-30), scale=scale))

    def worst(self, kind: str):
        rows = [r for r in self.rows if r["kind"] == kind]
        return max(rows, key=lambda r: r["max_rel"]) if rows else None

    def check(self, bf16_max_rel: float = 2.0 ** -7, bf16_frac_gt_ulp: float = 1e-3, f32_max_rel: float = 1e-3):
        """The stated teacher-forced bounds: a stored bf16 tensor differs from the rounded fp64 recomputation by at most
        one ulp of the tensor's largest value anywhere and by more than one ulp of the ELEMENT in at most 0.1 % of its
        elements; fp32 results (statistics, heat maps, parameter gradients) agree to 1e-3 of the tensor's largest value."""
        bad = []
        for r in self.rows:
            if r["kind"] == "bf16" and (r["max_rel"] > bf16_max_rel or r["frac_gt_ulp"] > bf16_frac_gt_ulp):
                bad.append(r)
            if r["kind"] == "f32" and r["max_rel"] > f32_max_rel:
                bad.append(r)
            if r["kind"] == "zero" and r["max_rel"] > 1e-2:  # (bf16 dz: its exact sum is ~2^-9 of the sum of magnitudes / sqrt(count))
                bad.append(r)
        return bad


def verify_step(t: Dict[str, torch.Tensor], heats: Sequence[torch.Tensor], grads: Dict[str, torch.Tensor], sd: Dict[str, torch.Tensor], x: torch.Tensor,
                target: Optional[torch.Tensor] = None, dheats: Optional[Sequence[Optional[torch.Tensor]]] = None, masks: Optional[Sequence[torch.Tensor]] = None,
                p_drop: float = 0.0, loss: str = "mse", focal_gamma: float = 3.0, dtype=torch.float64) -> Report:
    """``t``: the step's stored tensors by their TrainState name (NHWC bf16 activations / gradients, fp32 statistics);
    ``heats``: its three heat maps; ``grads``: its 74 parameter gradients by state_dict name; ``sd``: the fp32 weights it
    ran with; ``x`` / ``target`` (or ``dheats``: the upstream heat-map gradients) / ``masks`` ([B,16,H,W] keep masks): its inputs.
    Everything is moved to ``x.device`` and computed in ``dtype``."""
    dev = x.device
    rep = Report()
    B, _, H, W = x.shape

    def act(name):  # stored NHWC bf16 -> NCHW
        return t[name].to(dev).permute(0, 3, 1, 2).to(dtype)

    P = {k: v.to(dev).to(dtype) for k, v in sd.items() if v.dtype.is_floating_point}
    Wr = lambda k: _r(P[k])  # the packed bf16 operand
    eps = 1e-5
    ncls = P["final_1.weight"].shape[0]

    # ------------------------------------------------------------------ forward
    x16 = act("x16")[:, :x.shape[1]]
    rep.bf16("x16", x16, x.to(dtype))
    src = x16
    for lvl, name in enumerate(O.ENCODER):
        M = B * (H >> lvl) * (W >> lvl)
        for n, outname in ((1, f"{name}.a"), (2, f"X{lvl}0")):
            p = f"{name}.conv{n}"
            z = act(f"{name}.z{n}")
            rep.bf16(f"{name}.z{n}", z, F.conv2d(src, Wr(f"{p}.0.weight"), P[f"{p}.0.bias"], padding=1))
            mean = z.mean((0, 2, 3))
            var = z.var((0, 2, 3), unbiased=False)
            istd = 1.0 / torch.sqrt(var + eps)
            rep.f32(f"{name}.bn{n}.mean", t[f"{name}.bn{n}.mean"].to(dev), mean)
            rep.f32(f"{name}.bn{n}.istd", t[f"{name}.bn{n}.istd"].to(dev), istd)
            scale = P[f"{p}.1.weight"] * istd
            shift = P[f"{p}.1.bias"] - mean * scale
            y = F.relu(z * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
            got = act(outname)
            rep.bf16(outname, got, y)
            src = got
        if lvl < 3:
            pooled = act(f"P{lvl}0")
            rep.bf16(f"P{lvl}0", pooled, F.max_pool2d(src, 2))
            src = pooled
    for name in O.DECODER:
        high, lows, lvl = DEC[name]
        tag = name[-2:]
        U = act(f"U{tag}")
        rep.bf16(f"U{tag}", U, F.conv_transpose2d(act(high), Wr(f"{name}.up.weight"), P[f"{name}.up.bias"], stride=2))
        cat = torch.cat([U] + [act(l) for l in lows], 1)
        a = act(f"{name}.a")
        rep.bf16(f"{name}.a", a, F.relu(F.conv2d(cat, Wr(f"{name}.conv.conv1.0.weight"), P[f"{name}.conv.conv1.0.bias"], padding=1)))
        y = F.relu(F.conv2d(a, Wr(f"{name}.conv.conv2.0.weight"), P[f"{name}.conv.conv2.0.bias"], padding=1))
        rep.bf16(f"X{tag}", act(f"X{tag}"), y)
        for k, (h, node) in enumerate(HEAD_NODE.items()):
            if node == f"X{tag}":  # the fused head reads the fp32 value of X in the epilogue that produced it
                keep = (masks[k].to(dev).to(dtype) / (1.0 - p_drop)) if masks is not None else 1.0
                rep.f32(f"heat{k}", heats[k].to(dev), torch.sigmoid(F.conv2d(y * keep, P[f"{h}.weight"], P[f"{h}.bias"])))

    # ------------------------------------------------------------------ backward
    def G(name, ref):
        rep.f32("grad " + name, grads[name].to(dev), ref)

    dXh = {}
    for k, (h, node) in enumerate(HEAD_NODE.items()):
        p = heats[k].to(dev).to(dtype)
        if dheats is not None:
            dh = torch.zeros_like(p) if dheats[k] is None else dheats[k].to(dev).to(dtype)
        elif loss == "mse":
            dh = (2.0 / (3.0 * p.numel())) * (p - target.to(dev).to(dtype))
        else:  # FocalLoss_BCE_2d (tools/losses/focal_loss.py:264-301), mean of the three heads (trainer.py:125-134)
            pp = p.detach().clone().requires_grad_(True)
            (O.focal_loss_bce_2d(pp, target.to(dev).to(dtype), focal_gamma) / 3.0).backward()
            dh = pp.grad
        dl = dh * p * (1.0 - p)
        X = act(node)
        keep = (masks[k].to(dev).to(dtype) / (1.0 - p_drop)) if masks is not None else torch.ones_like(X)
        G(f"{h}.weight", torch.einsum("bkhw,bchw->kc", dl, X * keep).reshape(ncls, -1, 1, 1))
        G(f"{h}.bias", dl.sum((0, 2, 3)))
        d = torch.einsum("bkhw,kc->bchw", dl, P[f"{h}.weight"].reshape(ncls, -1)) * keep * (X > 0)
        dXh[node] = act(f"dXh{k}")
        rep.bf16(f"dXh{k}", dXh[node], d)

    def consumers(node):
        res = []
        for name in O.DECODER:
            _, lows, lvl = DEC[name]
            c = (16, 32, 64)[lvl]
            for j, l in enumerate(lows):
                if l == node:
                    res.append((name, c * (j + 1)))
        return res

    def gather(node, shape):
        """Sum over the same-resolution consumers of ``node`` of the dgrad of their first conv's slice (the transpose of unet.py:199-201)."""
        tot = torch.zeros(shape, dtype=dtype, device=dev)
        for cname, begin in consumers(node):
            w = Wr(f"{cname}.conv.conv1.0.weight")
            c = shape[1]
            tot = tot + conv2d_input(shape, w[:, begin:begin + c].contiguous(), act(f"dZ1{cname[-2:]}"), padding=1)
        return tot

    def deconv_dgrad(dname):
        return F.conv2d(act(f"dU{dname[-2:]}"), Wr(f"{dname}.up.weight"), stride=2)

    def decoder_node(name):
        high, lows, lvl = DEC[name]
        tag = name[-2:]
        dZ2, a = act(f"dZ2{tag}"), act(f"{name}.a")
        pre = f"{name}.conv"
        G(f"{pre}.conv2.0.weight", conv2d_weight(a, P[f"{pre}.conv2.0.weight"].shape, dZ2, padding=1))
        G(f"{pre}.conv2.0.bias", dZ2.sum((0, 2, 3)))
        dZ1 = act(f"dZ1{tag}")
        rep.bf16(f"dZ1{tag}", dZ1, conv2d_input(a.shape, Wr(f"{pre}.conv2.0.weight"), dZ2, padding=1) * (a > 0))
        U = act(f"U{tag}")
        cat = torch.cat([U] + [act(l) for l in lows], 1)
        G(f"{pre}.conv1.0.weight", conv2d_weight(cat, P[f"{pre}.conv1.0.weight"].shape, dZ1, padding=1))
        G(f"{pre}.conv1.0.bias", dZ1.sum((0, 2, 3)))
        c = U.shape[1]
        dU = act(f"dU{tag}")
        rep.bf16(f"dU{tag}", dU, conv2d_input(U.shape, Wr(f"{pre}.conv1.0.weight")[:, :c].contiguous(), dZ1, padding=1))
        G(f"{name}.up.bias", dU.sum((0, 2, 3)))
        G(f"{name}.up.weight", conv2d_weight(dU, P[f"{name}.up.weight"].shape, act(high), stride=2))

    def x_node(node, dname, addend=None, deconv_from=None):
        """dZ2 of decoder ``dname`` (the gradient at the pre-ReLU output of the conv that produced ``node``)."""
        X = act(node)
        if consumers(node):
            add = addend
            if deconv_from is not None:
                add = act("tmpX11")
                rep.bf16("tmpX11", add, deconv_dgrad(deconv_from))
            ref = (gather(node, X.shape) + (add if add is not None else 0.0)) * (X > 0)
        else:
            ref = deconv_dgrad(deconv_from) * (X > 0)
        rep.bf16(f"dZ2{dname[-2:]}", act(f"dZ2{dname[-2:]}"), ref)
        decoder_node(dname)

    decoder_node("up_concat03")  # X03's only consumer is head 3: dZ2_03 IS dXh2 (checked above)
    x_node("X02", "up_concat02", addend=dXh["X02"])
    x_node("X12", "up_concat12", deconv_from="up_concat03")
    x_node("X01", "up_concat01", addend=dXh["X01"])
    x_node("X11", "up_concat11", deconv_from="up_concat02")
    x_node("X21", "up_concat21", deconv_from="up_concat12")

    deconv_into = {3: "up_concat21", 2: "up_concat11", 1: "up_concat01"}
    for lvl in (3, 2, 1, 0):
        name, node = O.ENCODER[lvl], f"X{lvl}0"
        X = act(node)
        M = float(B * (H >> lvl) * (W >> lvl))
        addend = act(f"dpool{lvl}") if lvl < 3 else None
        if lvl == 0:
            ref = gather(node, X.shape) + addend
        elif consumers(node):
            tmp = act(f"tmp{lvl}")
            rep.bf16(f"tmp{lvl}", tmp, deconv_dgrad(deconv_into[lvl]) + addend)
            ref = gather(node, X.shape) + tmp
        else:
            ref = deconv_dgrad(deconv_into[lvl])
        dyh = act(f"{name}.dyh2")
        rep.bf16(f"{name}.dyh2", dyh, ref * (X > 0))
        src_of = {2: f"{name}.a", 1: "x16" if lvl == 0 else f"P{lvl - 1}0"}
        for n in (2, 1):
            p = f"{name}.conv{n}"
            z = act(f"{name}.z{n}")
            mean, istd = t[f"{name}.bn{n}.mean"].to(dev).to(dtype).view(1, -1, 1, 1), t[f"{name}.bn{n}.istd"].to(dev).to(dtype).view(1, -1, 1, 1)
            xhat = (z - mean) * istd
            s1, s2 = dyh.sum((0, 2, 3)), (dyh * xhat).sum((0, 2, 3))
            G(f"{p}.1.bias", s1)
            G(f"{p}.1.weight", s2)
            dz = act(f"{name}.dz{n}")
            # a conv bias in front of a BatchNorm has an analytically zero gradient (sum of dz = 0): the CUDA step holds an exact
            # zero, autograd holds rounding noise; either must vanish against the per-channel sum of |dz|
            rep.zero(f"grad {p}.0.bias", grads[f"{p}.0.bias"].to(dev), float(dz.abs().sum((0, 2, 3)).max()))
            rep.bf16(f"{name}.dz{n}", dz, P[f"{p}.1.weight"].view(1, -1, 1, 1) * istd * (dyh - s1.view(1, -1, 1, 1) / M - xhat * s2.view(1, -1, 1, 1) / M))
            inp = act(src_of[n])
            if n == 1 and lvl == 0:
                inp = inp[:, :x.shape[1]]
            G(f"{p}.0.weight", conv2d_weight(inp, P[f"{p}.0.weight"].shape, dz, padding=1))
            if n == 2:
                a = inp
                dyh = act(f"{name}.dyh1")
                rep.bf16(f"{name}.dyh1", dyh, conv2d_input(a.shape, Wr(f"{p}.0.weight"), dz, padding=1) * (a > 0))
            elif lvl > 0:
                dP = act(f"dP{lvl - 1}0")
                rep.bf16(f"dP{lvl - 1}0", dP, conv2d_input(inp.shape, Wr(f"{p}.0.weight"), dz, padding=1))
                # MaxPool2d(2) backward: the gradient goes to the first maximum of each window (ATen's rule)
                Xp = act(f"X{lvl - 1}0")
                _, idx = F.max_pool2d(Xp, 2, return_indices=True)
                dpool = torch.zeros_like(Xp).flatten(2).scatter_(2, idx.flatten(2), dP.flatten(2)).view_as(Xp)
                rep.bf16(f"dpool{lvl - 1}", act(f"dpool{lvl - 1}"), dpool)
    return rep
