"""bf16-storage emulation of the UNet_Nested training step.  TEST INFRASTRUCTURE ONLY (see oracle/unetpp_oracle.py).

The reference computes ``UNet_Nested.forward`` (models/unet.py:255-300) and its autograd backward
(trainer/trainer.py:135) in fp32.  The sm_100a path keeps every activation and every activation
gradient in **bf16** between its fused kernels (fp32 accumulation inside a kernel, fp32 parameter
gradients).  Compared with the fp32 oracle its parameter gradients therefore carry the rounding
noise of ~60 bf16 stores, which hides real errors below it.  This module restates the SAME arithmetic
as the fp32 oracle with a round-to-nearest-even bf16 rounding inserted at exactly the points where
the kernels store a tensor, so that the CUDA path can be checked against it with bounds two orders
of magnitude tighter than against the fp32 reference:

  forward  (training.forward_train)
    * input image, every conv / transposed-conv weight: rounded (the packed operands are bf16);
      biases, BatchNorm parameters and the 1x1 heads stay fp32;
    * encoder: z = bf16(conv + bias); batch statistics of the ROUNDED z; a = bf16(relu(bn(z)));
    * decoder: U = bf16(deconv + bias); a = bf16(relu(conv(cat) + bias)); X = bf16(relu(conv + bias));
      the fused head reads the fp32 (un-rounded) X in the same epilogue, its weight gradient later reads
      the stored bf16 X;
  backward (training.backward_train) — every stored gradient tensor is one rounding of the fp32 sum
  the producing kernel holds in TMEM / registers:
    * d(head input) (ReLU-masked), dZ2 / dZ1 / dU of every decoder node, dyh / dz of every encoder
      conv, the gradient of each pooled tensor;
    * where a node has a transposed-conv (or pool) consumer AND same-resolution skip consumers, the
      transposed-conv (+ pool) part is stored first (``tmp``) and the gather over the skip consumers
      adds it: two roundings.

Everything between two rounding points is torch autograd in the dtype of the inputs (fp32, or fp64
when the caller passes double tensors).  ``scale_grad`` lets a test inject a deliberate error into one
gradient edge to prove that the parity bound would catch it.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as F

from . import unetpp_oracle as O


def _round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(t.dtype)


class _RoundFwd(torch.autograd.Function):
    """bf16 rounding of a stored value; the gradient passes unchanged."""

    @staticmethod
    def forward(ctx, t):
        return _round(t)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundGrad(torch.autograd.Function):
    """Identity whose backward rounds the (already accumulated) gradient of the node to bf16: a stored gradient tensor.
    ``sink`` = (dict, name, mask | None): the stored tensor (times the ReLU mask the producing kernel applies) is recorded."""

    @staticmethod
    def forward(ctx, t, scale, sink):
        ctx.scale, ctx.sink = scale, sink
        return t.view_as(t)

    @staticmethod
    def backward(ctx, g):
        if ctx.scale != 1.0:
            g = g * ctx.scale
        g = _round(g)
        if ctx.sink is not None:
            d, name, mask = ctx.sink
            d[name] = _nhwc(g if mask is None else g * mask)
        return g, None, None


def _nhwc(t: torch.Tensor) -> torch.Tensor:
    """NCHW value -> the NHWC bf16 tensor the CUDA step would hold."""
    return t.detach().permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def Rf(t):
    return _RoundFwd.apply(t)


def Rg(t, scale: float = 1.0, sink=None):
    return _RoundGrad.apply(t, scale, sink)


DEC = {"up_concat01": ("X10", ("X00",)), "up_concat11": ("X20", ("X10",)), "up_concat21": ("X30", ("X20",)),
       "up_concat02": ("X11", ("X00", "X01")), "up_concat12": ("X21", ("X10", "X11")), "up_concat03": ("X12", ("X00", "X01", "X02"))}
# name of the 'tmp' tensor (gradient through the pool / upsample consumers, stored before the gather adds it) per node
TMP_NAME = {"X10": "tmp1", "X20": "tmp2", "X11": "tmpX11"}


def forward_bf16(sd: Dict[str, torch.Tensor], x: torch.Tensor, dropout_masks: Optional[Sequence[torch.Tensor]] = None, p_drop: float = 0.4,
                 new_stats: Optional[dict] = None, scale_grad: Optional[Dict[str, float]] = None, capture: Optional[dict] = None):
    """Training-mode forward with the storage points of unet_nested4tiny_objects_keypoints_b200/training.py.
    Tensor names are those of training.TrainState ('conv00.z1', 'X00', 'U01', 'up_concat01.a', 'dZ101', 'conv00.dyh2', 'tmp1',
    'dP00', 'dpool0', 'dXh0', ...).  ``scale_grad``: {gradient tensor name: factor} multiplies the gradient stored under that
    name (error injection).  ``capture``: a dict that receives every stored tensor — activations now, gradients during
    backward() — in the CUDA step's own format (NHWC bf16; BatchNorm mean / istd fp32), for oracle/teacher_forced.py."""
    sg = scale_grad or {}
    bn = "conv00.conv1.1.weight" in sd
    deconv = "up_concat01.up.weight" in sd
    W = lambda k: Rf(sd[k])  # bf16 packed operand of a conv / transposed conv

    def stored(t, name, gname, relu_out=False):
        """A tensor stored in bf16 whose total gradient is stored in bf16 too (after the ReLU mask when it is a ReLU output)."""
        v = Rf(t)
        if capture is not None and name is not None:
            capture[name] = _nhwc(v)
        sink = None if capture is None or gname is None else (capture, gname, (v.detach() > 0).to(v.dtype) if relu_out else None)
        return Rg(v, sg.get(gname, 1.0), sink)

    def grad_only(t, gname, mask_of=None):
        sink = None if capture is None or gname is None else (capture, gname, None if mask_of is None else (mask_of.detach() > 0).to(t.dtype))
        return Rg(t, sg.get(gname, 1.0), sink)

    def enc_conv(src, name, n, outname):
        p = f"{name}.conv{n}"
        z = F.conv2d(src, W(f"{p}.0.weight"), sd[f"{p}.0.bias"], padding=1)
        if not bn:  # unet.py:137-143
            return stored(F.relu(z), outname, f"{name}.dz{n}", relu_out=True)
        z = stored(z, f"{name}.z{n}", f"{name}.dz{n}")
        if capture is not None:
            zd = z.detach()
            capture[f"{name}.bn{n}.mean"] = zd.mean((0, 2, 3)).float()
            capture[f"{name}.bn{n}.istd"] = (1.0 / torch.sqrt(zd.var((0, 2, 3), unbiased=False) + 1e-5)).float()
        rm, rv = sd[f"{p}.1.running_mean"].clone(), sd[f"{p}.1.running_var"].clone()
        y = F.batch_norm(z, rm, rv, sd[f"{p}.1.weight"], sd[f"{p}.1.bias"], training=True, momentum=0.1, eps=1e-5)
        if new_stats is not None:
            new_stats[f"{p}.1.running_mean"], new_stats[f"{p}.1.running_var"] = rm, rv
        return stored(F.relu(y), outname, f"{name}.dyh{n}", relu_out=True)

    X: Dict[str, torch.Tensor] = {}   # node -> tensor the same-resolution (skip) consumers read
    Xa: Dict[str, torch.Tensor] = {}  # node -> tensor the pool / upsample consumers read (their gradient is stored separately: 'tmp')
    src = Rf(x)
    if capture is not None:
        capture["x16"] = _nhwc(F.pad(src, (0, 0, 0, 0, 0, 16 - src.shape[1])))
    for lvl, name in enumerate(O.ENCODER):
        a = enc_conv(src, name, 1, f"{name}.a")
        node = f"X{lvl}0"
        X[node] = enc_conv(a, name, 2, node)
        Xa[node] = grad_only(X[node], TMP_NAME.get(node))
        if lvl < 3:
            pooled = F.max_pool2d(grad_only(Xa[node], f"dpool{lvl}"), 2)
            if capture is not None:
                capture[f"P{lvl}0"] = _nhwc(pooled)
            src = grad_only(pooled, f"dP{lvl}0")
    fp32_out: Dict[str, torch.Tensor] = {}
    for name in O.DECODER:
        high, lows = DEC[name]
        tag = name[-2:]
        if deconv:
            U = stored(F.conv_transpose2d(Xa[high], W(f"{name}.up.weight"), sd[f"{name}.up.bias"], stride=2), f"U{tag}", f"dU{tag}")
        else:  # unet.py:189-191; the kernels run the 1x1 conv on the low-resolution tensor first (both are linear, the bilinear weights sum to one)
            V = stored(F.conv2d(Xa[high], W(f"{name}.up.1.weight"), sd[f"{name}.up.1.bias"]), f"V{tag}", f"dV{tag}")
            U = stored(F.interpolate(V, scale_factor=2, mode="bilinear", align_corners=True), f"U{tag}", f"dU{tag}")
        cat = torch.cat([U] + [X[l] for l in lows], 1)
        a = stored(F.relu(F.conv2d(cat, W(f"{name}.conv.conv1.0.weight"), sd[f"{name}.conv.conv1.0.bias"], padding=1)), f"{name}.a", f"dZ1{tag}", relu_out=True)
        y = F.relu(F.conv2d(a, W(f"{name}.conv.conv2.0.weight"), sd[f"{name}.conv.conv2.0.bias"], padding=1))
        node = f"X{tag}"
        fp32_out[node] = y
        X[node] = stored(y, node, f"dZ2{tag}", relu_out=True)
        Xa[node] = grad_only(X[node], TMP_NAME.get(node))
    outs = []
    for i, (h, node) in enumerate(zip(O.HEADS, ("X01", "X02", "X03"))):
        hr = grad_only(X[node], f"dXh{i}", mask_of=X[node])  # gradient of the head input: stored in bf16, ReLU-masked (dXh)
        hf = fp32_out[node].detach()                         # the fused head reads the un-rounded fp32 value in the epilogue that produced it
        if dropout_masks is not None:
            keep = dropout_masks[i].to(hr.dtype) / (1.0 - p_drop)
            hr, hf = hr * keep, hf * keep
        w, b = sd[f"{h}.weight"], sd[f"{h}.bias"]
        logits = F.conv2d(hr, w, b)  # carries the gradients: dW from the stored bf16 X, dX through the fp32 head weights
        logits = logits + (F.conv2d(hf, w.detach(), b.detach()) - logits.detach())  # value: the head of the fp32 X
        outs.append(torch.sigmoid(logits))
    return tuple(outs)


def train_step_grads_bf16(sd: Dict[str, torch.Tensor], x, target, dropout_masks=None, loss: str = "mse", scale_grad=None, dheats=None, p_drop: float = 0.4,
                          capture: Optional[dict] = None, loss_fn=None):
    """Like unetpp_oracle.train_step_grads with the bf16 storage points of the CUDA path.
    ``dheats``: optional upstream gradients of the three heat maps (then ``target`` / ``loss`` are ignored); ``loss_fn``: optional
    callable mapping the three heat maps to a scalar loss (takes precedence)."""
    params = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in sd.items() if v.dtype.is_floating_point and "running_" not in k)
    full = dict(sd)
    full.update(params)
    new_stats: dict = {}
    outs = forward_bf16(full, x, dropout_masks=dropout_masks, p_drop=p_drop, new_stats=new_stats, scale_grad=scale_grad, capture=capture)
    if loss_fn is not None:  # any differentiable function of the three heat maps (the trainer builds its loss on them, trainer.py:125-135)
        L = loss_fn(outs)
    elif dheats is not None:
        L = sum((o * d).sum() for o, d in zip(outs, dheats) if d is not None)
    elif loss == "mse":
        L = O.mse_heatmap_loss(outs, target)
    else:
        L = sum(O.focal_loss_bce_2d(o, target) for o in outs) / len(outs)
    L.backward()
    grads = OrderedDict((k, (p.grad.detach() if p.grad is not None else torch.zeros_like(p))) for k, p in params.items())
    return L.detach(), tuple(o.detach() for o in outs), grads, new_stats
