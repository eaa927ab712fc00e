"""CPU oracle for the UNet_Nested (UNet++) hot path.  TEST INFRASTRUCTURE ONLY.

This is a plain restatement, in functional torch-CPU / numpy arithmetic, of what the reference
repository computes on the path BASELINE.json's ``north_star`` names.  It is the checker for
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs; nothing under ``unet_nested4tiny_objects_keypoints_b200/`` imports it and the product path
never routes through it.

Parity status: **pinned**.  The reference's own tests hold no golden vectors for this path
(SURVEY.md §4, §8c), so ``oracle/make_golden.py`` imports the real reference module from
``/root/reference`` in the build container, checks every function below against it and commits the
resulting vectors under ``tests/golden/``; ``tests/test_oracle_golden.py`` replays them wherever
the tests run (the reference itself does not travel to the GPU box).

All ``file:line`` citations are into ``/root/reference``.
The arithmetic of Conv2d / BatchNorm2d / ConvTranspose2d / MaxPool2d / Dropout / sigmoid lives in
PyTorch (third-party; the reference pins only "Pytorch 1.0", README.md:20; installed here: torch
2.11.0, CPU backend) — their semantics are restated through ``torch.nn.functional`` calls in fp32
(or fp64 when the caller passes double tensors).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

ENCODER = ("conv00", "conv10", "conv20", "conv30")
DECODER = ("up_concat01", "up_concat11", "up_concat21", "up_concat02", "up_concat12", "up_concat03")
HEADS = ("final_1", "final_2", "final_3")


def filters(feature_scale: int = 2) -> List[int]:
    """models/unet.py:214-216 — ``[32,64,128,256,512] / feature_scale`` (the 5th is unused)."""
    return [int(x / feature_scale) for x in (32, 64, 128, 256, 512)]


# --------------------------------------------------------------------------------------------
# state_dict layout (models/unet.py:206-254)
# --------------------------------------------------------------------------------------------
def state_dict_spec(in_channels: int = 3, n_classes: int = 4, feature_scale: int = 2, is_deconv: bool = True,
                    is_batchnorm: bool = True) -> "OrderedDict[str, Tuple[Tuple[int, ...], torch.dtype]]":
    """Key -> (shape, dtype) in the exact registration order of the reference module.
    ``is_batchnorm=False`` drops the BatchNorm entries of the encoder (unet.py:129-143: the Sequential is then conv, ReLU);
    ``is_deconv=False`` replaces ``up.weight/up.bias`` by ``up.1.weight [Cout,Cin,1,1] / up.1.bias`` (unet.py:189-191:
    Sequential(UpsamplingBilinear2d(2), Conv2d 1x1))."""
    f = filters(feature_scale)
    spec: "OrderedDict[str, Tuple[Tuple[int, ...], torch.dtype]]" = OrderedDict()
    cin = in_channels
    for name, cout in zip(ENCODER, f[:4]):  # unetConv2 with BN (unet.py:129-136)
        c = cin
        for n in (1, 2):
            p = f"{name}.conv{n}"
            spec[f"{p}.0.weight"] = ((cout, c, 3, 3), torch.float32)
            spec[f"{p}.0.bias"] = ((cout,), torch.float32)
            if is_batchnorm:
                spec[f"{p}.1.weight"] = ((cout,), torch.float32)
                spec[f"{p}.1.bias"] = ((cout,), torch.float32)
                spec[f"{p}.1.running_mean"] = ((cout,), torch.float32)
                spec[f"{p}.1.running_var"] = ((cout,), torch.float32)
                spec[f"{p}.1.num_batches_tracked"] = ((), torch.int64)
            c = cout
        cin = cout
    # unetUp(in_size, out_size, is_deconv, n_concat) — unet.py:182-187, 226-236
    dec = {"up_concat01": (f[1], f[0], 2), "up_concat11": (f[2], f[1], 2), "up_concat21": (f[3], f[2], 2),
           "up_concat02": (f[1], f[0], 3), "up_concat12": (f[2], f[1], 3), "up_concat03": (f[1], f[0], 4)}
    for name in DECODER:
        cin_, cout, ncat = dec[name]
        c = cin_ + (ncat - 2) * cout
        for n in (1, 2):  # `conv` is registered before `up` (unet.py:185-187)
            spec[f"{name}.conv.conv{n}.0.weight"] = ((cout, c, 3, 3), torch.float32)
            spec[f"{name}.conv.conv{n}.0.bias"] = ((cout,), torch.float32)
            c = cout
        if is_deconv:
            spec[f"{name}.up.weight"] = ((cin_, cout, 2, 2), torch.float32)
            spec[f"{name}.up.bias"] = ((cout,), torch.float32)
        else:
            spec[f"{name}.up.1.weight"] = ((cout, cin_, 1, 1), torch.float32)
            spec[f"{name}.up.1.bias"] = ((cout,), torch.float32)
    for h in HEADS:
        spec[f"{h}.weight"] = ((n_classes, f[0], 1, 1), torch.float32)
        spec[f"{h}.bias"] = ((n_classes,), torch.float32)
    return spec


def synth_state_dict(seed: int, dtype=torch.float32, **kw) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic synthetic weights with the reference layout (NOT the reference init):
    conv weights ~ N(0, sqrt(2/fan_in)), biases ~ U(-0.1, 0.1), BN gamma ~ N(1, 0.1), beta ~
    N(0, 0.1), running_mean ~ N(0, 0.2), running_var ~ U(0.5, 1.5).  Generated from a CPU
    ``torch.Generator`` so that tests and the fixture generator agree without storing 2.2 MB."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for k, (shape, dt) in state_dict_spec(**kw).items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.int64)
        elif k.endswith("running_var"):
            sd[k] = (torch.rand(shape, generator=g) + 0.5).to(dtype)
        elif k.endswith("running_mean"):
            sd[k] = (torch.randn(shape, generator=g) * 0.2).to(dtype)
        elif ".1.weight" in k and ".up." not in k:
            sd[k] = (1 + 0.1 * torch.randn(shape, generator=g)).to(dtype)
        elif ".1.bias" in k and ".up." not in k:
            sd[k] = (0.1 * torch.randn(shape, generator=g)).to(dtype)
        elif k.endswith("bias"):
            sd[k] = (torch.rand(shape, generator=g) * 0.2 - 0.1).to(dtype)
        else:  # conv / deconv weights
            fan_in = int(np.prod(shape[1:])) if not k.endswith("up.weight") else shape[0] * 4
            sd[k] = (torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)).to(dtype)
    return sd


# --------------------------------------------------------------------------------------------
# forward (models/unet.py:255-300)
# --------------------------------------------------------------------------------------------
def _unet_conv2(sd, prefix: str, x, bn: bool, training: bool, new_stats: Optional[dict], inter: Optional[dict]):
    """unetConv2.forward (unet.py:150-156): n=2 x (Conv2d 3x3 s1 p1 + bias -> [BN] -> ReLU)."""
    for n in (1, 2):
        p = f"{prefix}.conv{n}"
        x = F.conv2d(x, sd[f"{p}.0.weight"], sd[f"{p}.0.bias"], stride=1, padding=1)  # unet.py:132,140
        if bn:  # unet.py:133 — BatchNorm2d(eps=1e-5, momentum=0.1)
            rm, rv = sd[f"{p}.1.running_mean"], sd[f"{p}.1.running_var"]
            if training:
                rm, rv = rm.clone(), rv.clone()
            if inter is not None:
                inter[f"{p}.z"] = x
            x = F.batch_norm(x, rm, rv, sd[f"{p}.1.weight"], sd[f"{p}.1.bias"], training=training, momentum=0.1, eps=1e-5)
            if training and new_stats is not None:
                new_stats[f"{p}.1.running_mean"] = rm
                new_stats[f"{p}.1.running_var"] = rv
                new_stats[f"{p}.1.num_batches_tracked"] = sd[f"{p}.1.num_batches_tracked"] + 1
        x = F.relu(x)  # unet.py:134,141
        if inter is not None:
            inter[f"{p}.y"] = x
    return x


def _unet_up(sd, prefix: str, high, lows: Sequence[torch.Tensor], inter: Optional[dict]):
    """unetUp.forward (unet.py:198-202): cat([up(high), *lows], 1) -> unetConv2 (no BN); ``up`` is ConvTranspose2d(k2,s2)
    (unet.py:187, is_deconv=True) or UpsamplingBilinear2d(2) [= bilinear, align_corners=True] + Conv2d 1x1 (unet.py:189-191);
    which one is read off the state_dict keys."""
    if f"{prefix}.up.weight" in sd:
        up = F.conv_transpose2d(high, sd[f"{prefix}.up.weight"], sd[f"{prefix}.up.bias"], stride=2)  # unet.py:187
    else:
        up = F.interpolate(high, scale_factor=2, mode="bilinear", align_corners=True)  # nn.UpsamplingBilinear2d(scale_factor=2)
        up = F.conv2d(up, sd[f"{prefix}.up.1.weight"], sd[f"{prefix}.up.1.bias"])
    if inter is not None:
        inter[f"{prefix}.up"] = up
    cat = torch.cat([up, *lows], 1)  # unet.py:200-201: order = [up, low_1, low_2, ...]
    return _unet_conv2(sd, f"{prefix}.conv", cat, False, False, None, inter)


def forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, training: bool = False,
            dropout_masks: Optional[Sequence[torch.Tensor]] = None, p_drop: float = 0.4,
            new_stats: Optional[dict] = None, inter: Optional[dict] = None):
    """UNet_Nested.forward (unet.py:255-300).  Returns (final_1, final_2, final_3).

    ``dropout_masks``: in training mode, three 0/1 keep-masks of shape [B,16,H,W] applied as
    ``x * mask / (1 - p)`` (nn.Dropout(p=0.4), unet.py:254,283-286).  ``None`` in training mode
    means "dropout disabled" (the documented parity policy: torch's RNG stream is not reproducible
    from a custom kernel, so parity runs pass explicit masks or none).  Eval: identity."""
    pool = lambda t: F.max_pool2d(t, 2)  # unet.py:219
    bn = "conv00.conv1.1.weight" in sd  # is_batchnorm (unet.py:129-143), read off the state_dict keys
    X00 = _unet_conv2(sd, "conv00", x, bn, training, new_stats, inter)
    X10 = _unet_conv2(sd, "conv10", pool(X00), bn, training, new_stats, inter)
    X20 = _unet_conv2(sd, "conv20", pool(X10), bn, training, new_stats, inter)
    X30 = _unet_conv2(sd, "conv30", pool(X20), bn, training, new_stats, inter)
    X01 = _unet_up(sd, "up_concat01", X10, [X00], inter)
    X11 = _unet_up(sd, "up_concat11", X20, [X10], inter)
    X21 = _unet_up(sd, "up_concat21", X30, [X20], inter)
    X02 = _unet_up(sd, "up_concat02", X11, [X00, X01], inter)
    X12 = _unet_up(sd, "up_concat12", X21, [X10, X11], inter)
    X03 = _unet_up(sd, "up_concat03", X12, [X00, X01, X02], inter)
    outs = []
    for i, (h, X) in enumerate(zip(HEADS, (X01, X02, X03))):
        if training and dropout_masks is not None:
            X = X * dropout_masks[i].to(X.dtype) / (1.0 - p_drop)
        outs.append(torch.sigmoid(F.conv2d(X, sd[f"{h}.weight"], sd[f"{h}.bias"])))  # unet.py:283-286
    if inter is not None:
        inter.update(X00=X00, X10=X10, X20=X20, X30=X30, X01=X01, X11=X11, X21=X21, X02=X02, X12=X12, X03=X03)
    return tuple(outs)


# --------------------------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------------------------
def mse_heatmap_loss(outputs: Sequence[torch.Tensor], target: torch.Tensor) -> torch.Tensor:
    """nn.MSELoss() (trainer.py:427) per head, averaged over the heads like trainer.py:125-135:
    L = (1/3) sum_k mean((final_k - T)^2)."""
    return sum(F.mse_loss(o, target) for o in outputs) / len(outputs)


def focal_loss_bce_2d(inp: torch.Tensor, target: torch.Tensor, gamma: float = 3.0) -> torch.Tensor:
    """FocalLoss_BCE_2d.forward with size_average=False (focal_loss.py:264-301):
    e = 1 - |p - t| + 1e-20; loss = sum(-(1-e)^gamma * log e) / (B*C)."""
    samples = inp.shape[0] * inp.shape[1]  # focal_loss.py:277-282 (view(-1,H,W) then shape[0])
    e = 1 - torch.abs(inp - target) + 1e-20
    return (-1 * (1 - e) ** gamma * torch.log(e)).sum() / samples


def train_step_grads(sd: Dict[str, torch.Tensor], x, target, dropout_masks=None, loss: str = "mse"):
    """One forward+backward in training mode through torch-CPU autograd.
    Returns (loss, outputs, grads{name: tensor}, new_stats{...})."""
    params = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in sd.items() if v.dtype.is_floating_point
                         and "running_" not in k)
    full = dict(sd)
    full.update(params)
    new_stats: dict = {}
    outs = forward(full, x, training=True, dropout_masks=dropout_masks, new_stats=new_stats)
    if loss == "mse":
        L = mse_heatmap_loss(outs, target)
    else:
        L = sum(focal_loss_bce_2d(o, target) for o in outs) / len(outs)
    L.backward()
    grads = OrderedDict((k, p.grad.detach()) for k, p in params.items())
    return L.detach(), tuple(o.detach() for o in outs), grads, new_stats


# --------------------------------------------------------------------------------------------
# optimizer (tools/optimizers/adamw.py:38-100)
# --------------------------------------------------------------------------------------------
def adamw_reference_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: int, lr: float = 1e-3,
                         betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
    """The reference's AdamW: NOT torch.optim.AdamW.  Decay is ``weight_decay * p_old`` — *not*
    multiplied by lr — subtracted after the Adam update (adamw.py:92-96); denom = sqrt(v)+eps with
    the bias corrections folded into step_size (adamw.py:86-90).  ``step`` is the 1-based count
    after increment (adamw.py:73).  Returns new (p, m, v)."""
    b1, b2 = betas
    m = m * b1 + (1 - b1) * g  # adamw.py:76
    v = v * b2 + (1 - b2) * g * g  # adamw.py:77
    denom = v.sqrt() + eps  # adamw.py:84
    step_size = lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step)  # adamw.py:86-88
    decayed = p * weight_decay  # adamw.py:91 (pre-update p)
    p = p - step_size * m / denom  # adamw.py:92/96
    if weight_decay != 0:
        p = p - decayed  # adamw.py:93
    return p, m, v


def sgdw_reference_step(p: torch.Tensor, g: torch.Tensor, buf: Optional[torch.Tensor], lr: float, momentum: float = 0.0, dampening: float = 0.0,
                        weight_decay: float = 0.0):
    """The reference's SGDW.step AS SHIPPED (tools/optimizers/sgdw.py:77-110): the momentum buffer is maintained
    (sgdw.py:95-102: first step ``buf = g``, later ``buf = momentum*buf + (1-dampening)*g``) but the descent direction
    ``d_p`` is never applied to the parameter — the only write to ``p`` is the decoupled decay ``p -= weight_decay * p``
    (sgdw.py:107-108), NOT scaled by ``lr``.  Returns new (p, buf); ``buf`` stays None when momentum == 0."""
    if momentum != 0:
        buf = g.clone() if buf is None else buf * momentum + (1 - dampening) * g
    if weight_decay != 0:
        p = p - weight_decay * p
    return p, buf


def adabound_reference_step(p, g, m, v, step: int, lr: float = 1e-3, betas=(0.9, 0.999), final_lr: float = 0.1, gamma: float = 1e-3,
                            eps: float = 1e-8, weight_decay: float = 0.0, base_lr: Optional[float] = None):
    """AdaBound.step without amsbound (tools/optimizers/adabound.py:57-122): L2 decay folded into the gradient (101-102),
    Adam moments (105-106), per-element step ``clamp(step_size / (sqrt(v)+eps), lower, upper) * m`` with the bounds of
    lines 119-121 (``final_lr`` scaled by lr/base_lr, base_lr = the lr at construction, adabound.py:49).  ``step`` is 1-based."""
    b1, b2 = betas
    base_lr = lr if base_lr is None else base_lr
    if weight_decay != 0:
        g = g + weight_decay * p
    m = m * b1 + (1 - b1) * g
    v = v * b2 + (1 - b2) * g * g
    denom = v.sqrt() + eps
    step_size = lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step)
    flr = final_lr * lr / base_lr
    lower = flr * (1 - 1 / (gamma * step + 1))
    upper = flr * (1 + 1 / (gamma * step))
    upd = (torch.full_like(denom, step_size) / denom).clamp(lower, upper) * m
    return p - upd, m, v


def sgd_reference_step(p, g, buf, lr: float, momentum: float = 0.0, weight_decay: float = 0.0):
    """torch.optim.SGD as the trainer builds it (trainer/trainer.py:345-349: lr, momentum, weight_decay; dampening 0, no
    Nesterov): g += wd*p; buf = g (first step) or momentum*buf + g; p -= lr*buf."""
    if weight_decay != 0:
        g = g + weight_decay * p
    if momentum != 0:
        buf = g.clone() if buf is None else buf * momentum + g
        g = buf
    return p - lr * g, buf


def adam_reference_step(p, g, m, v, step: int, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
    """torch.optim.Adam as the trainer builds it (trainer/trainer.py:351-355: L2 weight decay folded into the gradient):
    denom = sqrt(v)/sqrt(1-b2^t) + eps; p -= lr/(1-b1^t) * m/denom."""
    b1, b2 = betas
    if weight_decay != 0:
        g = g + weight_decay * p
    m = m * b1 + (1 - b1) * g
    v = v * b2 + (1 - b2) * g * g
    denom = v.sqrt() / math.sqrt(1 - b2 ** step) + eps
    return p - (lr / (1 - b1 ** step)) * m / denom, m, v


# --------------------------------------------------------------------------------------------
# peak extraction (tools/misc/heatmap.py:173-178) and target synthesis (tools/misc/helper.py:87-172)
# --------------------------------------------------------------------------------------------
def argmax_keypoints(heat: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Per (b, c) plane: ``yx = np.where(h == h.max())`` and the FIRST hit in row-major order
    (heatmap.py:173-176), returned as [x, y] (x first, heatmap.py:178) with the identity scale
    ``origin/size`` (heatmap.py:151-152,175-176).  Returns (xy int32 [B,C,2], peak value [B,C])."""
    heat = np.asarray(heat)
    B, C, H, W = heat.shape
    xy = np.zeros((B, C, 2), dtype=np.int32)
    val = np.zeros((B, C), dtype=heat.dtype)
    for b in range(B):
        for c in range(C):
            h = heat[b, c]
            yx = np.where(h == h.max())
            y, x = int(yx[0][0]), int(yx[1][0])
            xy[b, c] = (int(W * x / W), int(H * y / H))
            val[b, c] = h[y, x]
    return xy, val


def topk_peaks(heat: np.ndarray, num: int, threshold: float = 0.5):
    """The multi-point form of Heatmap.extract_points_ (heatmap.py:148-208) with the OpenCV watershed replaced by its effect on
    separated blobs: threshold the plane (heatmap.py:158), take the strict local maxima of the 8-neighbourhood under the order
    (value descending, index ascending) as the regions' maxima (heatmap.py:165-166: ``np.max`` per region, first index, 173-176),
    sort brightest first (heatmap.py:167) and keep ``num`` (heatmap.py:170); one retry at 0.9 x threshold (heatmap.py:187-190).
    Returns (xy int32 [B,C,num,2] as [x,y], -1 padded; value [B,C,num]; count [B,C]).  tests/golden/unetpp_r2.* pins it against the
    real ``Heatmap.extract_points_`` (cv2) on synthetic targets."""
    heat = np.asarray(heat)
    B, C, H, W = heat.shape
    xy = -np.ones((B, C, num, 2), dtype=np.int32)
    val = np.zeros((B, C, num), dtype=heat.dtype)
    cnt = np.zeros((B, C), dtype=np.int32)
    for b in range(B):
        for c in range(C):
            h = heat[b, c]
            for thr in (threshold, np.float32(threshold) * np.float32(0.9)):
                cands = []
                ys, xs = np.where(h >= thr)
                for y, x in zip(ys.tolist(), xs.tolist()):
                    v, p, ok = h[y, x], y * W + x, True
                    for dy in (-1, 0, 1):
                        for dx in (-1, 0, 1):
                            yy, xx = y + dy, x + dx
                            if (dy or dx) and 0 <= yy < H and 0 <= xx < W:
                                u, q = h[yy, xx], yy * W + xx
                                if u > v or (u == v and q < p):
                                    ok = False
                    if ok:
                        cands.append((-float(v), p))
                if cands:
                    cands.sort()
                    for r, (nv, p) in enumerate(cands[:num]):
                        xy[b, c, r] = (p % W, p // W)
                        val[b, c, r] = h[p // W, p % W]
                    cnt[b, c] = min(num, len(cands))
                    break
    return xy, val, cnt


def create_heatmap(target: np.ndarray, image_height: int, image_width: int) -> np.ndarray:
    """helper.create_heatmap (helper.py:87-172): 7 keypoints -> 4 channels, groups {0},{1,2,3},{4},{5,6};
    per point ``exp(-0.5 * dist / 3)`` with dist the Euclidean DISTANCE (not squared); channels 1 and 3
    are summed then divided by their max (also when channel 3 holds a single point, i.e. 6 key points)."""
    target = np.asarray(target, dtype=np.float64)
    N, C, _ = target.shape
    out = np.zeros((N, 4, image_height, image_width), dtype=np.float32)
    xs = np.arange(image_width, dtype=np.float64)[None, :]
    ys = np.arange(image_height, dtype=np.float64)[:, None]
    groups = ([0], [1, 2, 3], [4], list(range(5, C)))
    for n in range(N):
        for ch, pts in enumerate(groups):
            acc = np.zeros((image_height, image_width), dtype=np.float32)
            for pt in pts:
                d = np.sqrt((xs - target[n, pt, 0]) ** 2 + (ys - target[n, pt, 1]) ** 2)
                g = np.exp(-0.5 * d / 3)
                if ch in (0, 2):
                    acc = g.astype(np.float32)  # helper.py:106,142: plain assignment
                else:
                    acc = acc + g  # helper.py:122,158: float32 += float64 -> float32
                    acc = acc.astype(np.float32)
            if ch in (1, 3):
                acc = acc / np.max(acc)  # helper.py:123,159: planes 1 and 3 are always divided by their maximum
            out[n, ch] = acc
    return out
