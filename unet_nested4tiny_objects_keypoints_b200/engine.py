"""Host-side engine: drives libunpp.so (C ABI, include/unpp.h) for one UNet_Nested on one device.

The reference executes ``UNet_Nested.forward`` (models/unet.py:255-300) as ~80 ATen/cuDNN launches
on NCHW fp32 tensors.  Here the same DAG runs as 25 launches of hand-written sm_100a kernels on
NHWC bf16 tensors with fp32 accumulation:

  * the fp32 NCHW input becomes a 4-channel NHWC bf16 tensor (8 B/pixel) that the first conv reads through the
    first-layer MMA mode of ``unpp_conv_tc`` (include/unpp.h, block2x2 = 2);
  * every 3x3 conv (+bias / folded BatchNorm, + ReLU) is one ``unpp_conv_tc`` call whose K loop walks
    the list of source tensors, so ``torch.cat`` (unet.py:199-201) is never materialised;
  * ``ConvTranspose2d(k2,s2)`` (unet.py:187) is a pointwise tensor-core GEMM with a scatter epilogue;
  * the 1x1 heads + sigmoid (unet.py:242-244,283-286) ride in the epilogue of the node's second conv;
  * eval-mode BatchNorm (unet.py:133) is folded into the packed bf16 weights and the fp32 bias;
  * ``is_deconv=False`` (unet.py:189-191): the 1x1 conv of the bilinear branch runs on the low-resolution tensor (pointwise
    tensor-core GEMM), ``unpp_bilinear_up2x`` upsamples its output; ``is_batchnorm=False``: nothing to fold.

PyTorch owns all device memory (activation arena, packed weights); this module only passes
``data_ptr()``s and the current CUDA stream across the C ABI.  There is no fallback path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from .ops import MODE_DECONV, pick_n_tile

ENCODER = ("conv00", "conv10", "conv20", "conv30")
# name -> (high-resolution source node, low sources in concat order, level)   (unet.py:268-277)
DECODER = {
    "up_concat01": ("X10", ("X00",), 0),
    "up_concat11": ("X20", ("X10",), 1),
    "up_concat21": ("X30", ("X20",), 2),
    "up_concat02": ("X11", ("X00", "X01"), 0),
    "up_concat12": ("X21", ("X10", "X11"), 1),
    "up_concat03": ("X12", ("X00", "X01", "X02"), 0),
}
DECODER_ORDER = ("up_concat01", "up_concat11", "up_concat21", "up_concat02", "up_concat12", "up_concat03")
HEAD_OF = {"up_concat01": "final_1", "up_concat02": "final_2", "up_concat03": "final_3"}


def _named_params_walk(model):
    for prefix, mod in model.named_modules():
        src = mod._former_parameters if getattr(mod, "_is_replica", False) else mod._parameters
        for k, v in src.items():
            if v is not None:
                yield (prefix + "." + k if prefix else k), v


def named_params(model):
    """``model.named_parameters()`` that also works on nn.DataParallel replicas: a replica's parameters are broadcast copies
    kept as plain (non-leaf) tensors in ``_former_parameters`` and ``parameters()`` is empty there (torch/nn/parallel/replicate.py).
    The walk over the module tree (~40 modules, three times per eager training step) is cached on the module: parameters are updated
    in place (optimizers, load_state_dict, .to()), so the list stays valid as long as the module keeps its Parameter OBJECTS — checked
    on the first and last one; replicas are fresh objects at every forward and are walked every time."""
    if getattr(model, "_is_replica", False):
        return list(_named_params_walk(model))
    cache = model.__dict__.get("_unpp_named_params")
    if cache is not None:
        first_owner, first_key, last_owner, last_key, lst = cache
        if first_owner._parameters.get(first_key) is lst[0][1] and last_owner._parameters.get(last_key) is lst[-1][1]:
            return lst
    lst = list(_named_params_walk(model))
    if lst:
        def owner(name):
            mod = model
            parts = name.split(".")
            for part in parts[:-1]:
                mod = getattr(mod, part)
            return mod, parts[-1]
        fo, fk = owner(lst[0][0])
        lo, lk = owner(lst[-1][0])
        model.__dict__["_unpp_named_params"] = (fo, fk, lo, lk, lst)
    return lst


class Engine:
    def __init__(self, model, device: torch.device):
        if device.type != "cuda":
            raise RuntimeError("the B200-native UNet_Nested engine needs a CUDA device")
        f = [int(c / model.feature_scale) for c in (32, 64, 128, 256)]
        if f != [16, 32, 64, 128] or model.n_classes > 8 or model.in_channels > 16:
            raise ValueError("the sm_100a engine supports feature_scale=2, n_classes<=8, in_channels<=16")
        self.model = model
        self.device = device
        self.filters = f
        ops.lib()  # fail now if libunpp.so is missing: there is no other execution path
        self._packed_eval = None
        self._packed_key = None
        self._arena: Dict[Tuple, Dict[str, torch.Tensor]] = {}
        self._keep: List = []
        self.first_layer_c4 = True  # inference: conv00.conv1 reads a 4-channel NHWC input (8 B/pixel) through the first-layer MMA mode
        self.fuse_deconv = True  # inference: fold the k2s2 transposed conv of the full-resolution nodes into the consuming conv

    # ------------------------------------------------------------------------------ weights
    def _param_key(self):
        return tuple((p.data_ptr(), p._version) for p in [q for _, q in named_params(self.model)] + list(self.model.buffers()))

    def _conv_seq(self, name: str, n: int):
        m = self.model
        for part in name.split("."):
            m = getattr(m, part)
        return getattr(m, "conv%d" % n)

    def packed_eval(self):
        """bf16 UMMA-layout weights with eval-mode BatchNorm folded in (unet.py:133), cached until a
        parameter or buffer changes."""
        # An nn.DataParallel replica holds fresh broadcast copies at every forward: their versions are always 0 and the caching
        # allocator tends to hand out the same addresses again, so (data_ptr, _version) cannot tell new weights from old.
        # Replicas are re-packed at every eval forward; the master module (and every single-device use) is cached.
        replica = bool(getattr(self.model, "_is_replica", False))
        key = None if replica else self._param_key()
        if key is not None and self._packed_eval is not None and self._packed_key == key:
            return self._packed_eval
        P: Dict[str, Dict] = {}
        with torch.no_grad(), torch.cuda.device(self.device), ops.pack_arena(self.device):
            for name in ENCODER:
                for n in (1, 2):
                    seq = self._conv_seq(name, n)
                    conv = seq[0]
                    if self.model.is_batchnorm:
                        scale, bias = ops.bn_fold(seq[1], conv.bias)
                    else:  # unet.py:137-143: conv + ReLU only
                        scale, bias = None, conv.bias.detach().float().contiguous()
                    if name == "conv00" and n == 1 and self.first_layer_c4 and conv.weight.shape[1] <= 4:
                        # first layer: 4-channel (8-byte) input pixels instead of the 16-channel zero-padded tensor
                        P[f"{name}.c{n}"] = dict(w=ops.pack_weights_c4(conv.weight.detach().float(), scale=scale), bias=bias, n_total=16,
                                                 n_tile=ops.NTile(16, b2=2), c4=True)
                        continue
                    P[f"{name}.c{n}"] = self._pack_fwd_conv(conv.weight, scale, bias)
            for name in DECODER_ORDER:
                up = getattr(self.model, name)
                for n in (1, 2):
                    conv = self._conv_seq(name + ".conv", n)[0]
                    P[f"{name}.c{n}"] = self._pack_fwd_conv(conv.weight, None, conv.bias.detach().float().contiguous())
                if not self.model.is_deconv:
                    # unet.py:189-191: UpsamplingBilinear2d(2) + Conv2d 1x1.  Both are linear and the bilinear weights sum to one,
                    # so the 1x1 conv (bias included) runs first, on the low-resolution tensor: a pointwise tensor-core GEMM
                    pw = up.up[1]
                    cout, cin = pw.weight.shape[0], pw.weight.shape[1]
                    nt = pick_n_tile(cout, cin, 1)
                    P[f"{name}.up"] = dict(w=ops.pack_weights(pw.weight.detach().float(), 0, 1, cout, nt, cin), bias=pw.bias.detach().float().contiguous(),
                                           n_total=cout, n_tile=nt, cout=cout)
                    continue
                conv1 = self._conv_seq(name + ".conv", 1)[0]
                cu = up.up.weight.shape[1]
                if cu == 16 and up.up.weight.shape[0] == 32 and conv1.weight.shape[1] - cu <= 48:
                    # full-resolution node: the transposed conv is folded into the consuming conv (no U tensor, no deconv launch)
                    comp, table = ops.compose_deconv_conv(conv1.weight, cu, up.up.weight, up.up.bias, conv1.bias)
                    klow = conv1.weight.shape[1] - cu
                    P[f"{name}.c1.fused"] = dict(w=ops.pack_weights_b2(conv1.weight.detach().float(), False, klow, k_begin=cu), bias=table,
                                                 low_w=ops.pack_weights(comp, 6, 9, 64, 64, 32), n_total=16, n_tile=ops.NTile(16, b2=True))
                cin, cout = up.up.weight.shape[0], up.up.weight.shape[1]
                nt = pick_n_tile(4 * cout, cin, 1, deconv=True)
                P[f"{name}.up"] = dict(w=ops.pack_weights(up.up.weight.float(), 2, 1, 4 * cout, nt, cin), bias=up.up.bias.detach().float().contiguous(),
                                       n_total=4 * cout, n_tile=nt, cout=cout)
            for h in ("final_1", "final_2", "final_3"):
                m = getattr(self.model, h)
                P[h] = dict(w=m.weight.detach().float().reshape(m.weight.shape[0], -1).contiguous(), b=m.bias.detach().float().contiguous())
        self._packed_eval, self._packed_key = P, key
        return P

    def _pack_fwd_conv(self, weight: torch.Tensor, scale, bias):
        w = weight.detach().float()
        cout, cin = w.shape[0], w.shape[1]
        if cin % 16:  # first layer: 3 input channels live in a 16-channel zero-padded NHWC tensor (the pack kernel zero-fills)
            cin = 16 * ((cin + 15) // 16)
        if cout == 16 and cin % 16 == 0 and cin <= 64:
            # full-resolution level (every source has 16 channels): 2x2 output-blocked kernel path
            return dict(w=ops.pack_weights_b2(w, False, cin, scale=scale), bias=bias, n_total=cout, n_tile=ops.NTile(16, b2=True))
        nt = pick_n_tile(cout, cin, 9)
        return dict(w=ops.pack_weights(w, 0, 9, cout, nt, cin, scale=scale), bias=bias, n_total=cout, n_tile=nt)

    # ------------------------------------------------------------------------------ activations
    def arena(self, B: int, H: int, W: int, kind: str) -> Dict[str, torch.Tensor]:
        key = (B, H, W, kind)
        a = self._arena.get(key)
        if a is None:
            f = self.filters
            bf = dict(dtype=torch.bfloat16, device=self.device)
            a = {"x16": torch.empty(B, H, W, 16, **bf)} if kind != "eval4" else {"x4": torch.empty(B, H, W, 4, **bf)}
            for lvl, (node, c) in enumerate(zip(ENCODER, f)):
                h, w = H >> lvl, W >> lvl
                a[f"{node}.a"] = torch.empty(B, h, w, c, **bf)
                a[f"X{lvl}0"] = torch.empty(B, h, w, c, **bf)
                if lvl < 3:
                    a[f"P{lvl}0"] = torch.empty(B, h // 2, w // 2, c, **bf)
            for name in DECODER_ORDER:
                _, _, lvl = DECODER[name]
                h, w, c = H >> lvl, W >> lvl, f[lvl]
                tag = name[-2:]
                a[f"U{tag}"] = torch.empty(B, h, w, c, **bf)
                if not self.model.is_deconv:
                    a[f"V{tag}"] = torch.empty(B, h // 2, w // 2, c, **bf)  # 1x1 conv of the low-resolution source, before the x2 upsample
                a[f"{name}.a"] = torch.empty(B, h, w, c, **bf)
                a[f"X{tag}"] = torch.empty(B, h, w, c, **bf)
            if len(self._arena) >= 4:
                self._arena.pop(next(iter(self._arena)))
            self._arena[key] = a
        return a

    # ------------------------------------------------------------------------------ forward
    def _check_input(self, x: torch.Tensor, channels_last: bool = False):
        if x.dtype != torch.float32 and not (x.dtype == torch.uint8 and not self.model.training):
            raise ValueError("UNet_Nested expects float32 input (the reference runs fp32 NCHW); eval mode also takes uint8 images (scaled by 1/255 "
                             "like torchvision's ToTensor)")
        if channels_last and x.dtype != torch.uint8:
            raise ValueError("channels_last input is the uint8 image layout [B,H,W,C]; float32 input is NCHW like the reference's")
        if channels_last:
            B, H, W, Cin = x.shape
        else:
            B, Cin, H, W = x.shape
        if Cin != self.model.in_channels:
            raise ValueError(f"expected {self.model.in_channels} input channels, got {Cin}")
        if H % 8 or W % 8 or H < 8 or W < 8:
            # the reference's torch.cat throws for sizes not divisible by 8 (no pad logic in unetUp, unet.py:198-202)
            raise ValueError("UNet_Nested needs H and W divisible by 8")
        if x.device != self.device:
            raise RuntimeError("input is on a different device than the engine")
        return B, H, W

    def forward(self, x: torch.Tensor):
        if self.model.training:
            from .training import run_autograd
            return run_autograd(self, x)
        if getattr(self.model, "precision", "bf16") == "fp32":
            return self.forward_eval_fp32(x)
        # eval mode (trainer/trainer.py:200-210 runs it under set_grad_enabled(False)): inference kernels, no graph
        return self.forward_eval(x)

    def forward_eval(self, x: torch.Tensor, heads: Sequence[int] = (0, 1, 2), channels_last: bool = False):
        """Inference forward (BN folded, dropout off).  Returns the three sigmoid heat maps
        (fp32 NCHW) — ``None`` for heads not requested.  The requested heat maps are slices of ONE tensor
        ``self.last_heats`` [len(heads), B, classes, H, W] (one arg-max launch covers them all).
        ``x``: float32 [B,C,H,W] like the reference, or uint8 images ([B,C,H,W], or [B,H,W,C] with ``channels_last``) that are
        scaled by 1/255 on the device — torchvision's ToTensor (datasets/datasets_base.py:71-72) fused into the layout change."""
        B, H, W = self._check_input(x, channels_last)
        if B == 0:  # an empty batch: the reference's layers return empty tensors in eval mode
            self.last_heats = torch.empty(len(set(heads)), 0, self.model.n_classes, H, W, dtype=torch.float32, device=self.device)
            want = sorted(set(int(k) for k in heads))
            return tuple(self.last_heats[want.index(k)] if k in want else None for k in range(3))
        x = x.contiguous()
        # the kernels, their TMA descriptors and the per-device shared-memory opt-in belong to the ENGINE's device, whatever the
        # calling thread's current device is (the reference works regardless of it)
        with torch.cuda.device(self.device):
            return self._forward_eval(x, B, H, W, heads, channels_last)

    def _forward_eval(self, x: torch.Tensor, B: int, H: int, W: int, heads: Sequence[int], channels_last: bool = False):
        P = self.packed_eval()
        c4 = bool(P["conv00.c1"].get("c4"))
        A = self.arena(B, H, W, "eval4" if c4 else "eval")
        ncls = self.model.n_classes
        src = A["x4"] if c4 else A["x16"]
        ops.tag("input")
        if x.dtype == torch.uint8:
            ops.u8_to_nhwc(x, src, channels_last)
        elif c4:
            ops.nchw_to_nhwc4(x, src)
        else:
            ops.nchw_to_nhwc16(x, src)
        for lvl, name in enumerate(ENCODER):
            h, w = H >> lvl, W >> lvl
            p1, p2 = P[f"{name}.c1"], P[f"{name}.c2"]
            ops.tag(f"{name}.c1")
            ops.conv([src], B, h, w, p1["w"], p1["n_total"], p1["n_tile"], 9, bias=p1["bias"], relu=True, out=A[f"{name}.a"])
            # MaxPool2d(2) (unet.py:219,258,260,262) is written by the conv's own epilogue
            ops.tag(f"{name}.c2")
            ops.conv([A[f"{name}.a"]], B, h, w, p2["w"], p2["n_total"], p2["n_tile"], 9, bias=p2["bias"], relu=True, out=A[f"X{lvl}0"],
                     pooled=A[f"P{lvl}0"] if lvl < 3 else None)
            if lvl < 3:
                src = A[f"P{lvl}0"]
        heats: List[Optional[torch.Tensor]] = [None, None, None]
        want = sorted(set(int(k) for k in heads))
        self.last_heats = torch.empty(len(want), B, ncls, H, W, dtype=torch.float32, device=self.device)
        for i, k in enumerate(want):
            heats[k] = self.last_heats[i]
        for name in DECODER_ORDER:
            high, lows, lvl = DECODER[name]
            tag = name[-2:]
            h, w = H >> lvl, W >> lvl
            pu, p1, p2 = P[f"{name}.up"], P[f"{name}.c1"], P[f"{name}.c2"]
            pf = P.get(f"{name}.c1.fused")
            ops.tag(f"up{tag}.c1")
            if pf is not None and self.fuse_deconv:
                ops.conv([A[l] for l in lows], B, h, w, pf["w"], 16, pf["n_tile"], 9, bias=pf["bias"], bias_classes=9, relu=True, out=A[f"{name}.a"],
                         lowres=(A[high], pf["low_w"]))
            else:
                if self.model.is_deconv:
                    ops.conv([A[high]], B, h // 2, w // 2, pu["w"], pu["n_total"], pu["n_tile"], 1, bias=pu["bias"], mode=MODE_DECONV, out=A[f"U{tag}"])
                else:
                    ops.conv([A[high]], B, h // 2, w // 2, pu["w"], pu["n_total"], pu["n_tile"], 1, bias=pu["bias"], out=A[f"V{tag}"])
                    ops.bilinear_up2x(A[f"V{tag}"], A[f"U{tag}"])
                ops.conv([A[f"U{tag}"]] + [A[l] for l in lows], B, h, w, p1["w"], p1["n_total"], p1["n_tile"], 9, bias=p1["bias"], relu=True,
                          out=A[f"{name}.a"])
            head = None
            out = A[f"X{tag}"]
            ops.tag(f"up{tag}.c2")
            if name in HEAD_OF:
                k = int(HEAD_OF[name][-1]) - 1
                if heats[k] is not None:
                    ph = P[HEAD_OF[name]]
                    head = (ph["w"], ph["b"], heats[k], None, None, 1.0)
                if name == "up_concat03":
                    out = None  # X03 has no consumer besides its head
                    if head is None:
                        continue
            ops.conv([A[f"{name}.a"]], B, h, w, p2["w"], p2["n_total"], p2["n_tile"], 9, bias=p2["bias"], relu=True, out=out, head=head)
        ops.tag("")
        return tuple(heats)

    @torch.no_grad()
    def forward_eval_fp32(self, x: torch.Tensor):
        """The fp32 validation mode (``model.precision = "fp32"``, csrc/ref_kernels.cu): the same fused plan — virtual concat,
        eval-mode BatchNorm as per-channel scale / bias, ReLU, k2s2 transposed conv, 2x2 max pool, 1x1 heads + sigmoid — on NCHW
        fp32 tensors with fp32 FMAs, straight from the state_dict weights.  BASELINE's tolerance for this mode is <= 1e-3 relative
        against the reference (tests: ~1e-6).  ~30x slower than the tensor-core path; it exists to tell wiring from rounding."""
        B, H, W = self._check_input(x)
        if x.dtype != torch.float32:
            raise ValueError("the fp32 validation mode takes float32 input")
        if not self.model.is_deconv:
            raise ValueError("the fp32 validation mode covers the default is_deconv=True graph")
        m = self.model
        with torch.cuda.device(self.device):
            def block(srcs, seq1, seq2):
                for seq in (seq1, seq2):
                    conv = seq[0]
                    if m.is_batchnorm and len(seq) == 3:  # conv, BatchNorm, ReLU (unet.py:132-134)
                        scale, bias = ops.bn_fold(seq[1], conv.bias)
                    else:
                        scale, bias = None, conv.bias.detach().float().contiguous()
                    srcs = [ops.ref_conv(srcs, conv.weight, bias, scale=scale, relu=True)]
                return srcs[0]

            X = {}
            src = x.contiguous()
            for lvl, name in enumerate(ENCODER):
                mod = getattr(m, name)
                X[f"X{lvl}0"] = block([src], mod.conv1, mod.conv2)
                if lvl < 3:
                    src = ops.ref_maxpool2x2(X[f"X{lvl}0"])
            for name in DECODER_ORDER:
                high, lows, lvl = DECODER[name]
                up = getattr(m, name)
                U = ops.ref_deconv2x2(X[high], up.up.weight, up.up.bias)
                X[f"X{name[-2:]}"] = block([U] + [X[l] for l in lows], up.conv.conv1, up.conv.conv2)
            outs = []
            for hname, node in (("final_1", "X01"), ("final_2", "X02"), ("final_3", "X03")):
                hm = getattr(m, hname)
                outs.append(ops.ref_conv([X[node]], hm.weight, hm.bias.detach().float().contiguous(), sigmoid=True))
        return tuple(outs)

    @torch.no_grad()
    def predict_keypoints(self, x: torch.Tensor, head: int = 2):
        if head not in (0, 1, 2):
            raise ValueError("head must be 0, 1 or 2")
        heats = self.forward_eval(x, heads=(head,))
        if x.shape[0] == 0:
            return (torch.empty(0, self.model.n_classes, 2, dtype=torch.int32, device=self.device),
                    torch.empty(0, self.model.n_classes, dtype=torch.float32, device=self.device), heats)
        with torch.cuda.device(self.device):
            xy, val = ops.argmax_peaks(heats[head])
        return xy, val, heats
