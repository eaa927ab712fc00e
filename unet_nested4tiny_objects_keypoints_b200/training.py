"""Training step of UNet_Nested on the sm_100a kernels: forward with BatchNorm batch statistics and
head dropout, the full backward pass, and the glue that exposes both to ``torch.autograd``.

What the reference does (file:line into the reference repository):
  * forward in train mode — models/unet.py:255-300 with nn.BatchNorm2d batch statistics (unet.py:133)
    and nn.Dropout(0.4) before each 1x1 head (unet.py:254,283-286);
  * backward — PyTorch autograd over that graph, entered from ``avgloss.backward()``
    (trainer/trainer.py:135) with one upstream gradient per head.

How it runs here (all NHWC bf16 activations / gradients, fp32 accumulation, fp32 parameter grads):
  * encoder convs write the pre-BN tensor z and per-CTA (sum, sum^2) partials from the tensor-core
    epilogue; ``unpp_bn_finalize`` turns them into mean/istd (+ running-stat update) and
    ``unpp_bn_relu`` applies the affine + ReLU (+ the 2x2 max-pool copy);
  * every activation's gradient is ONE gather: a dgrad implicit GEMM whose K loop walks the dZ
    tensors of all same-resolution consumers (the transpose of the virtual concat), with the
    ReLU mask, the addend from the pool / transposed-conv / head branch and the per-channel sums for
    bias / BatchNorm gradients fused into its epilogue — fan-out accumulation is a fixed-order sum
    inside one kernel, hence deterministic;
  * weight gradients are per-CTA partials of ``unpp_wgrad`` reduced in fixed order straight into a
    flat fp32 gradient buffer laid out like ``model.parameters()`` (the buffer the data-parallel
    all-reduce and the fused AdamW work on).
"""
from __future__ import annotations

import os
import weakref
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from .engine import DECODER, DECODER_ORDER, ENCODER, HEAD_OF, Engine, named_params
from .ops import MODE_DECONV, pick_n_tile

PQ = [(0, 0), (0, 1), (1, 0), (1, 1)]
# Side streams of the training step: weight gradients (see backward_train), weight re-packing and dropout masks (fused.FusedTrainStep).
# bench.py switches them off for its per-kernel trace so that CUDA events bracket one kernel at a time.
WGRAD_SIDE_STREAM = os.environ.get("UNPP_WGRAD_STREAM", "1") != "0"


# ---------------------------------------------------------------------------------------------- flat parameter layout
def flat_layout(model) -> Tuple[Dict[str, Tuple[int, int]], int]:
    """name -> (offset, numel) in ``model.named_parameters()`` order (74 tensors, 553 260 elements)."""
    lay, off = {}, 0
    for name, p in named_params(model):
        lay[name] = (off, p.numel())
        off += p.numel()
    return lay, off


class TrainState:
    """Everything one (B, H, W) training shape needs on the device: saved activations, gradient
    tensors, statistics, partial-sum scratch and packed weights.  Pointers stay fixed for the life
    of the object, so a whole step can be captured in a CUDA graph."""

    def __init__(self, eng: Engine, B: int, H: int, W: int):
        self.eng, self.B, self.H, self.W = eng, B, H, W
        dev = eng.device
        f = eng.filters
        ncls = eng.model.n_classes
        bf = dict(dtype=torch.bfloat16, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        t: Dict[str, torch.Tensor] = {}
        self.t = t

        def act(name, lvl, c=None):
            t[name] = torch.empty(B, H >> lvl, W >> lvl, f[lvl] if c is None else c, **bf)

        t["x16"] = torch.empty(B, H, W, 16, **bf)
        for lvl, node in enumerate(ENCODER):
            for nm in ("z1", "a", "z2"):
                act(f"{node}.{nm}", lvl)
            act(f"X{lvl}0", lvl)
            if lvl < 3:
                t[f"P{lvl}0"] = torch.empty(B, H >> (lvl + 1), W >> (lvl + 1), f[lvl], **bf)
            for nm in ("dyh2", "dz2", "dyh1", "dz1"):
                act(f"{node}.{nm}", lvl)
            if lvl > 0:
                t[f"dP{lvl - 1}0"] = torch.empty(B, H >> lvl, W >> lvl, f[lvl - 1], **bf)  # grad of the pooled input of this level
            if lvl < 3:
                act(f"dpool{lvl}", lvl)  # un-pooled gradient flowing into X_{lvl}0
                act(f"tmp{lvl}", lvl)
            for n in (1, 2):
                for nm in ("mean", "istd", "scale", "shift"):
                    t[f"{node}.bn{n}.{nm}"] = torch.empty(f[lvl], **f32)
                t[f"{node}.bn{n}.sums"] = torch.empty(2 * f[lvl], **f32)
        for name in DECODER_ORDER:
            _, _, lvl = DECODER[name]
            tag = name[-2:]
            for nm in (f"U{tag}", f"{name}.a", f"X{tag}", f"dZ2{tag}", f"dZ1{tag}", f"dU{tag}"):
                act(nm, lvl)
            if not eng.model.is_deconv:  # unet.py:189-191: the 1x1 conv runs on the low-resolution tensor, before the x2 bilinear upsample
                for nm in (f"V{tag}", f"dV{tag}"):
                    t[nm] = torch.empty(B, H >> (lvl + 1), W >> (lvl + 1), f[lvl], **bf)
        act("tmpX11", 1)
        for k in range(3):
            t[f"mask{k}"] = torch.full((B, H, W), -1, dtype=torch.int16, device=dev)  # dropout keep bits, one 16-bit word per pixel
            # head 3 is X03's only consumer: its (already ReLU-masked) input gradient IS dZ2 of up_concat03
            t[f"dXh{k}"] = t["dZ203"] if k == 2 else torch.empty(B, H, W, 16, **bf)
        self.head_grid = ops.head_bwd_grid(B, H, W)
        self.head_nacc = ncls * 16 + ncls + 1 + 16
        t["head_partial"] = torch.empty(3, self.head_grid, self.head_nacc, **f32)
        t["head_red"] = torch.empty(3, self.head_nacc, **f32)
        self.heats: List[Optional[torch.Tensor]] = [None, None, None]
        self.scratch: Dict[str, torch.Tensor] = {}
        self.packed: Dict[str, torch.Tensor] = {}
        self.lay, self.nflat = flat_layout(eng.model)
        self.drop_scale = 1.0
        self.use_masks = False
        # Which forward the saved activations belong to: forward_train bumps `generation`; an autograd node remembers the value it
        # ran at and refuses to run its backward on activations a later forward overwrote.  `owner` is a weak reference to the
        # token of the autograd node whose backward is still outstanding (dead once the node is freed or its backward has run).
        self.generation = 0
        self.owner = None

    def busy(self) -> bool:
        tok = self.owner() if self.owner is not None else None
        return tok is not None and not tok.done

    def wgrad_stream(self) -> torch.cuda.Stream:
        s = getattr(self, "_wgrad_stream", None)
        if s is None:
            s = self._wgrad_stream = torch.cuda.Stream(self.eng.device)
        return s

    def scratch_f32(self, key: str, numel: int) -> torch.Tensor:
        s = self.scratch.get(key)
        if s is None or s.numel() < numel:
            s = self.scratch[key] = torch.empty(numel, dtype=torch.float32, device=self.eng.device)
        return s


# ---------------------------------------------------------------------------------------------- weights
def _mod(model, path: str):
    m = model
    for part in path.split("."):
        m = getattr(m, part)
    return m


def _w1(model, name):  # first conv of a decoder node / encoder level
    return _mod(model, (name + ".conv" if name.startswith("up_") else name) + ".conv1")[0]


def _w2(model, name):
    return _mod(model, (name + ".conv" if name.startswith("up_") else name) + ".conv2")[0]


def consumers_of(node: str) -> List[Tuple[str, int]]:
    """Same-resolution conv consumers of activation ``node`` (as a low source of a decoder c1):
    [(decoder name, first input channel of the slice)] — the transpose of unet.py:199-201."""
    res = []
    for name in DECODER_ORDER:
        _, lows, lvl = DECODER[name]
        c = (16, 32, 64)[lvl]
        for j, l in enumerate(lows):
            if l == node:
                res.append((name, c * (j + 1)))
    return res


def pack_train(ts: TrainState) -> None:
    """(Re)pack every weight the step needs from the current fp32 parameters: forward operands,
    flipped/transposed dgrad operands, and the gather operands of the fan-out activations.
    The ~60 pack jobs are recorded once into a device-resident table (source pointers are the
    parameters' storage, which stays put) and replayed with ONE launch per step."""
    m = ts.eng.model
    key = tuple(p.data_ptr() for _, p in named_params(m))
    if getattr(ts, "_pack_key", None) == key:
        ops.pack_batched(ts._pack_table, ts._pack_n)
        return
    ops.begin_pack_record()
    try:
        with ops.pack_arena(ts.eng.device, 8 << 20):
            _pack_train_jobs(ts)
    finally:
        jobs = ops.end_pack_record()
    ts._pack_jobs = jobs  # keeps the source views alive
    ts._pack_table, ts._pack_n, ts._pack_key = ops.make_pack_table(jobs, ts.eng.device), len(jobs), key
    ops.pack_batched(ts._pack_table, ts._pack_n)


def _pack_train_jobs(ts: TrainState) -> None:
    m, P = ts.eng.model, ts.packed
    dev = ts.eng.device

    def put(key, src, kind, taps, n_total, k_count, **kw):
        nt = kw.pop("n_tile")
        dst = P.get(key)
        k8_total = kw.get("k8_total", k_count // 8)
        b2 = kind in (0, 1) and taps == 9 and n_total == 16 and k_count % 16 == 0 and src.shape[0] == 16 and k8_total <= 8
        if b2:  # 16-channel full-resolution level: 2x2 output-blocked kernel path (unpp.h kinds 4 / 5)
            if dst is None:
                dst = P[key] = ops._zeros_bf16(64 * 16 * k8_total * 8, dev)
                P[key + ".nt"] = ops.NTile(16, b2=True)
            kw.pop("k8_total", None)
            ops.pack_weights_b2(src, kind == 1, k_count, dst=dst, k8_total=k8_total, **kw)
            return
        if dst is None:
            dst = P[key] = ops._zeros_bf16(n_total * taps * k8_total * 8, dev)
            P[key + ".nt"] = nt
        ops.pack_weights(src, kind, taps, n_total, nt, k_count, dst=dst, **kw)

    with torch.no_grad():
        for name in list(ENCODER) + list(DECODER_ORDER):
            for n, conv in ((1, _w1(m, name)), (2, _w2(m, name))):
                w = conv.weight.detach()
                cout, cin = w.shape[0], w.shape[1]
                if cin % 16:
                    cin = 16 * ((cin + 15) // 16)  # first layer: the pack kernel zero-fills input channels >= 3
                put(f"{name}.c{n}.fwd", w, 0, 9, cout, cin, n_tile=pick_n_tile(cout, cin, 9))
            w2 = _w2(m, name).weight.detach()
            c = w2.shape[0]
            put(f"{name}.c2.dgrad", w2, 1, 9, c, c, n_tile=pick_n_tile(c, c, 9))
        for lvl, name in enumerate(ENCODER):
            if lvl == 0:
                continue
            w1 = _w1(m, name).weight.detach()  # [cout, cin(prev level), 3, 3]: dgrad toward the pooled input
            cout, cin = w1.shape[0], w1.shape[1]
            put(f"{name}.c1.dgrad", w1, 1, 9, cin, cout, n_tile=pick_n_tile(cin, cout, 9))
        for name in DECODER_ORDER:
            up = getattr(m, name).up
            if m.is_deconv:
                wd = up.weight.detach()
                cin, cout = wd.shape[0], wd.shape[1]
                put(f"{name}.up.fwd", wd, 2, 1, 4 * cout, cin, n_tile=pick_n_tile(4 * cout, cin, 1, deconv=True))
                put(f"{name}.up.dgrad", wd, 3, 4, cin, cout, n_tile=pick_n_tile(cin, 4 * cout, 1))
            else:  # UpsamplingBilinear2d + Conv2d 1x1 (unet.py:189-191): pointwise GEMMs on the low-resolution grid
                wp = up[1].weight.detach()  # [cout, cin, 1, 1]
                cout, cin = wp.shape[0], wp.shape[1]
                put(f"{name}.up.fwd", wp, 0, 1, cout, cin, n_tile=pick_n_tile(cout, cin, 1))
                put(f"{name}.up.dgrad", wp, 1, 1, cin, cout, n_tile=pick_n_tile(cin, cout, 1))
            w1 = _w1(m, name).weight.detach()
            put(f"{name}.c1.dgradU", w1, 1, 9, cout, cout, n_tile=pick_n_tile(cout, cout, 9), n_begin=0)
        for node in ("X00", "X10", "X20", "X01", "X11", "X02"):
            cons = consumers_of(node)
            lvl = int(node[1])
            c = ts.eng.filters[lvl]
            ktot = c * len(cons)  # every consumer's dZ has as many channels as the node itself
            nt = pick_n_tile(c, ktot, 9)
            for j, (cname, begin) in enumerate(cons):
                put(f"{node}.gather", _w1(m, cname).weight.detach(), 1, 9, c, c, n_tile=nt, n_begin=begin, k8_total=ktot // 8, k_dst8=j * c // 8)


# ---------------------------------------------------------------------------------------------- forward (train mode)
def forward_train(ts: TrainState, x: torch.Tensor, update_running_stats: bool = True, before_decoder=None, after_input=None,
                  channels_last: bool = False) -> Tuple[torch.Tensor, ...]:
    eng, t, P, m = ts.eng, ts.t, ts.packed, ts.eng.model
    B, H, W = ts.B, ts.H, ts.W
    ncls = m.n_classes
    ts.generation += 1  # the saved activations now belong to this forward
    ops.tag("fwd")
    if update_running_stats:
        eng._packed_key = None  # running statistics change below without a torch version bump: drop the eval-mode fold cache
    if x.dtype == torch.uint8:  # 8-bit images ([B,C,H,W] or, with channels_last, [B,H,W,C]): ToTensor's 1/255 fused into the layout change
        ops.u8_to_nhwc(x, t["x16"], channels_last)
    else:
        ops.nchw_to_nhwc16(x, t["x16"])
    if after_input is not None:
        after_input()  # join point for the weight re-packing the caller put on a side stream
    src = t["x16"]
    tracked: List[torch.Tensor] = []
    stat_buffers: List[torch.Tensor] = []
    for lvl, name in enumerate(ENCODER):
        h, w, c = H >> lvl, W >> lvl, eng.filters[lvl]
        count = B * h * w
        for n in (1, 2):
            seq = _mod(m, f"{name}.conv{n}")
            conv = seq[0]
            key = f"{name}.c{n}.fwd"
            nt = P[key + ".nt"]
            if not m.is_batchnorm:  # unet.py:137-143: conv + ReLU (+ the 2x2 max pool of the level's output, written by the same epilogue)
                if n == 1:
                    ops.conv([src], B, h, w, P[key], c, nt, 9, bias=conv.bias, relu=True, out=t[f"{name}.a"])
                    src = t[f"{name}.a"]
                else:
                    ops.conv([src], B, h, w, P[key], c, nt, 9, bias=conv.bias, relu=True, out=t[f"X{lvl}0"], pooled=t[f"P{lvl}0"] if lvl < 3 else None)
                    src = t[f"P{lvl}0"] if lvl < 3 else None
                continue
            bn = seq[1]
            z = t[f"{name}.z{n}"]
            cin = src.shape[-1]
            g = ops.conv_grid([cin], B, h, w, c, nt, 9)
            part = ts.scratch_f32("stats", g * 2 * c)
            ops.conv([src], B, h, w, P[key], c, nt, 9, bias=conv.bias, out=z, stats_partial=part)
            pre = f"{name}.bn{n}"
            mom = 0.1 if bn.momentum is None else bn.momentum
            ops.bn_finalize(part, g, c, count, bn.weight, bn.bias, bn.running_mean if update_running_stats else None,
                            bn.running_var if update_running_stats else None, mom, bn.eps, t[pre + ".mean"], t[pre + ".istd"], t[pre + ".scale"],
                            t[pre + ".shift"])
            if update_running_stats:
                tracked.append(bn.num_batches_tracked)
                stat_buffers += [bn.running_mean, bn.running_var]
            if n == 1:
                ops.bn_relu(z, t[pre + ".scale"], t[pre + ".shift"], t[f"{name}.a"])
                src = t[f"{name}.a"]
            else:
                ops.bn_relu(z, t[pre + ".scale"], t[pre + ".shift"], t[f"X{lvl}0"], t[f"P{lvl}0"] if lvl < 3 else None)
                src = t[f"P{lvl}0"] if lvl < 3 else None
    if tracked:
        torch._foreach_add_(tracked, 1)  # the eight int64 num_batches_tracked counters: one launch
        if not torch.cuda.is_current_stream_capturing():  # (a captured step bumps the versions once per replay, fused.FusedTrainStep.step_device)
            torch._C._increment_version(stat_buffers)  # running_mean / running_var were written through raw pointers
    if before_decoder is not None:
        before_decoder()  # join point for work the caller put on a side stream (the head dropout masks)
    heats: List[torch.Tensor] = [None, None, None]
    for name in DECODER_ORDER:
        high, lows, lvl = DECODER[name]
        tag = name[-2:]
        h, w, c = H >> lvl, W >> lvl, eng.filters[lvl]
        node = getattr(m, name)
        ku, k1, k2 = f"{name}.up.fwd", f"{name}.c1.fwd", f"{name}.c2.fwd"
        if m.is_deconv:
            ops.conv([t[high]], B, h // 2, w // 2, P[ku], 4 * c, P[ku + ".nt"], 1, bias=node.up.bias, mode=MODE_DECONV, out=t[f"U{tag}"])
        else:
            ops.conv([t[high]], B, h // 2, w // 2, P[ku], c, P[ku + ".nt"], 1, bias=node.up[1].bias, out=t[f"V{tag}"])
            ops.bilinear_up2x(t[f"V{tag}"], t[f"U{tag}"])
        ops.conv([t[f"U{tag}"]] + [t[l] for l in lows], B, h, w, P[k1], c, P[k1 + ".nt"], 9, bias=_w1(m, name).bias, relu=True, out=t[f"{name}.a"])
        head = None
        if name in HEAD_OF:
            k = int(HEAD_OF[name][-1]) - 1
            hm = getattr(m, HEAD_OF[name])
            heats[k] = torch.empty(B, ncls, H, W, dtype=torch.float32, device=eng.device)
            head = (hm.weight.view(ncls, -1), hm.bias, heats[k], None, t[f"mask{k}"] if ts.use_masks else None, ts.drop_scale)
        ops.conv([t[f"{name}.a"]], B, h, w, P[k2], c, P[k2 + ".nt"], 9, bias=_w2(m, name).bias, relu=True, out=t[f"X{tag}"], head=head)
    ts.heats = heats
    return tuple(heats)


# ---------------------------------------------------------------------------------------------- backward
def backward_train(ts: TrainState, G: torch.Tensor, dheats: Optional[Sequence[Optional[torch.Tensor]]] = None, target: Optional[torch.Tensor] = None,
                   coef: float = 0.0, loss_kind: int = 0, gamma: float = 3.0) -> None:
    """Fills the flat fp32 gradient buffer ``G`` (layout ``ts.lay``).  Either ``dheats`` (upstream
    gradients of the three heat maps, fp32 NCHW; ``None`` entries mean zero) or ``target`` + ``coef``
    (fused MSE: d heat = coef * (heat - target); the summed squared error lands in
    ``ts.t['head_red'][k, ncls*17]``)."""
    eng, t, P, m = ts.eng, ts.t, ts.packed, ts.eng.model
    B, H, W = ts.B, ts.H, ts.W
    ncls = m.n_classes
    lay = ts.lay

    def goff(pname):
        return lay[pname][0]

    # Reductions whose result only the optimizer reads (weight / bias gradients) are queued and run as ONE batched
    # launch at the end of the backward pass; each therefore owns its partial buffer (keyed by the parameter name).
    ops.begin_reduce_queue()
    ops.tag("bwd")

    def stats_buf(srcs_C, n_total, nt, taps, h, w, key="stats"):
        g = ops.conv_grid(srcs_C, B, h, w, n_total, nt, taps)
        return g, ts.scratch_f32(key, g * 2 * n_total)

    def bias_from_stats(part, g, c, pname):
        ops.reduce_partials(part, g, 2 * c, c, G, out_offset=goff(pname), defer=True)

    # Weight gradients only feed the optimizer: they run on a side stream (a parallel branch of the captured step), each one
    # ordered after the kernel that produced its dZ.  The one-CTA-per-SM kernels of the two streams cannot share an SM, but the
    # CTAs of a queued weight-gradient kernel take over the SMs a finishing dgrad kernel frees, instead of the whole GPU draining
    # and re-filling between every pair of dependent launches.
    cur = torch.cuda.current_stream(eng.device)
    side = ts.wgrad_stream() if WGRAD_SIDE_STREAM else None
    if side is not None:
        side.wait_stream(cur)

    class _on_side:
        def __enter__(self_):
            if side is not None:
                ev = torch.cuda.Event()
                ev.record(cur)
                side.wait_event(ev)
                self_.ctx = torch.cuda.stream(side)
                self_.ctx.__enter__()

        def __exit__(self_, *exc):
            if side is not None:
                self_.ctx.__exit__(*exc)

    def wgrad_conv(srcs, dz, h, w, pname, ci_count=None):
        cins = [s.shape[-1] for s in srcs]
        cin, cout = sum(cins), dz.shape[-1]
        g = ops.wgrad_grid(cins, B, h, w, cout, 9)
        part = ts.scratch_f32("wgrad:" + pname, g * 9 * cin * cout)
        with _on_side():
            ops.wgrad(srcs, B, h, w, dz, cout, 9, part)
        real = cin if ci_count is None else ci_count
        ops.wgrad_reduce(part, g, 9, cin, cout, G, 0, real, real * 9, 9, 1, dst_offset=goff(pname), defer=True)

    def wgrad_deconv(xhigh, dU, h2, w2, pname):  # xhigh [B,h2,w2,cin], dU [B,2h2,2w2,cout]; the four taps in one launch
        cin, cout = xhigh.shape[-1], dU.shape[-1]
        g = ops.wgrad_grid([cin], B, h2, w2, cout, 1, dz_view="all4")
        part = ts.scratch_f32("wgrad:" + pname, 4 * g * cin * cout)
        with _on_side():
            ops.wgrad([xhigh], B, h2, w2, dU, cout, 1, part, dz_view="all4")
        for pq in range(4):
            ops.wgrad_reduce(part, g, 1, cin, cout, G, 0, cin, 4, cout * 4, 0, dst_offset=goff(pname) + pq, partial_offset=pq * g * cin * cout, defer=True)

    def wgrad_up_bilinear(xhigh, tag, h2, w2, pname):
        """is_deconv=False (unet.py:189-191): dV = bilinear_x2^T(dU) on the low-resolution grid, then the 1x1 conv's weight gradient."""
        dV = t[f"dV{tag}"]
        ops.bilinear_up2x_bwd(t[f"dU{tag}"], dV)
        cin, cout = xhigh.shape[-1], dV.shape[-1]
        g = ops.wgrad_grid([cin], B, h2, w2, cout, 1)
        part = ts.scratch_f32("wgrad:" + pname, g * cin * cout)
        with _on_side():
            ops.wgrad([xhigh], B, h2, w2, dV, cout, 1, part)
        ops.wgrad_reduce(part, g, 1, cin, cout, G, 0, cin, cin, 1, 0, dst_offset=goff(pname), defer=True)

    def up_dgrad_srcs_C(dname):
        """Source channel list of the gradient GEMM toward the low-resolution input of decoder ``dname``'s upsample."""
        c = t[f"dU{dname[-2:]}"].shape[-1]
        return [c] * 4 if m.is_deconv else [c]

    def deconv_dgrad(dname, h2, w2, out, addend=None, mask=None, stats=None):
        """grad wrt the LOW-resolution input of decoder ``dname``'s upsample (transposed conv, or bilinear + 1x1 conv)."""
        key = f"{dname}.up.dgrad"
        cin = out.shape[-1]
        if m.is_deconv:
            ops.conv([t[f"dU{dname[-2:]}"]] * 4, B, h2, w2, P[key], cin, P[key + ".nt"], 1, out=out, addend=addend, relu_mask_src=mask, strided=PQ, **(stats or {}))
        else:
            ops.conv([t[f"dV{dname[-2:]}"]], B, h2, w2, P[key], cin, P[key + ".nt"], 1, out=out, addend=addend, relu_mask_src=mask, **(stats or {}))

    # ---- heads: sigmoid' + 1x1 dgrad/wgrad + dropout mask (+ fused MSE); output already masked by X_0k > 0
    for name, hname in HEAD_OF.items():
        k = int(hname[-1]) - 1
        hm = getattr(m, hname)
        part = t["head_partial"][k]
        dh = None if dheats is None else dheats[k]
        if target is None and dh is None:  # this head does not contribute to the loss
            dh = torch.zeros_like(ts.heats[k])
        if True:
            ops.head_bwd(ts.heats[k], dh, target if dh is None else None, coef, t[f"X{name[-2:]}"], t[f"mask{k}"] if ts.use_masks else None,
                         ts.drop_scale, hm.weight.view(ncls, -1), t[f"dXh{k}"], part, loss_kind=loss_kind, gamma=gamma)
            with _on_side():  # (only the optimizer and the loss read-out consume it)
                ops.reduce_partials(part, ts.head_grid, ts.head_nacc, ts.head_nacc, t["head_red"][k])
        ops.reduce_partials(t["head_red"], 1, 0, ncls * 16, G, out_offset=goff(hname + ".weight"), partial_offset=k * ts.head_nacc, defer=True)
        ops.reduce_partials(t["head_red"], 1, 0, ncls, G, out_offset=goff(hname + ".bias"), partial_offset=k * ts.head_nacc + ncls * 16, defer=True)

    def decoder_node_backward(name, dZ2_ready_bias_done):
        """Given dZ2 (grad at the pre-ReLU output of the node's second conv), produce dZ1, dU and all
        parameter gradients of the node."""
        high, lows, lvl = DECODER[name]
        tag = name[-2:]
        h, w, c = H >> lvl, W >> lvl, eng.filters[lvl]
        dZ2, dZ1, dU = t[f"dZ2{tag}"], t[f"dZ1{tag}"], t[f"dU{tag}"]
        pre = f"{name}.conv"
        wgrad_conv([t[f"{name}.a"]], dZ2, h, w, f"{pre}.conv2.0.weight")
        key = f"{name}.c2.dgrad"
        g, part = stats_buf([c], c, P[key + ".nt"], 9, h, w, key=f"stats:{pre}.conv1.0.bias")
        ops.conv([dZ2], B, h, w, P[key], c, P[key + ".nt"], 9, out=dZ1, relu_mask_src=t[f"{name}.a"], stats_partial=part)
        bias_from_stats(part, g, c, f"{pre}.conv1.0.bias")
        wgrad_conv([t[f"U{tag}"]] + [t[l] for l in lows], dZ1, h, w, f"{pre}.conv1.0.weight")
        key = f"{name}.c1.dgradU"
        up_bias = f"{name}.up.bias" if m.is_deconv else f"{name}.up.1.bias"
        g, part = stats_buf([c], c, P[key + ".nt"], 9, h, w, key=f"stats:{up_bias}")
        ops.conv([dZ1], B, h, w, P[key], c, P[key + ".nt"], 9, out=dU, stats_partial=part)
        bias_from_stats(part, g, c, up_bias)  # bilinear weights sum to one: sum(dV) == sum(dU)
        if m.is_deconv:
            wgrad_deconv(t[high], dU, h // 2, w // 2, f"{name}.up.weight")
        else:
            wgrad_up_bilinear(t[high], tag, h // 2, w // 2, f"{name}.up.1.weight")

    def gather(node, h, w, out, addend, mask, stats):
        cons = consumers_of(node)
        key = f"{node}.gather"
        c = out.shape[-1]
        ops.conv([t[f"dZ1{cn[-2:]}"] for cn, _ in cons], B, h, w, P[key], c, P[key + ".nt"], 9, out=out, addend=addend, relu_mask_src=mask, **stats)

    # ---- decoder, deepest nesting first.  X03's only consumer is head 3: dZ2_03 = dXh2 (already masked)
    c0 = eng.filters[0]
    ops.reduce_partials(t["head_red"], 1, 0, c0, G, out_offset=goff("up_concat03.conv.conv2.0.bias"), partial_offset=2 * ts.head_nacc + ncls * 17 + 1, defer=True)
    decoder_node_backward("up_concat03", True)

    def x_node_decoder(node, dname, addend, extra_deconv_from=None):
        """dZ2 of decoder node ``dname`` whose output activation is ``node``."""
        lvl = int(node[1])
        h, w, c = H >> lvl, W >> lvl, eng.filters[lvl]
        out = t[f"dZ2{dname[-2:]}"]
        cons = consumers_of(node)
        key_for_stats = f"{node}.gather" if cons else f"{extra_deconv_from}.up.dgrad"
        if cons:
            add = addend
            if extra_deconv_from is not None:
                tmp = t["tmpX11"]
                deconv_dgrad(extra_deconv_from, h, w, tmp, addend=addend)
                add = tmp
            g, part = stats_buf([c] * len(cons), c, P[key_for_stats + ".nt"], 9, h, w, key=f"stats:{dname}.conv.conv2.0.bias")
            gather(node, h, w, out, add, t[node], dict(stats_partial=part))
        else:
            g, part = stats_buf(up_dgrad_srcs_C(extra_deconv_from), c, P[key_for_stats + ".nt"], 1, h, w, key=f"stats:{dname}.conv.conv2.0.bias")
            deconv_dgrad(extra_deconv_from, h, w, out, addend=addend, mask=t[node], stats=dict(stats_partial=part))
        bias_from_stats(part, g, c, f"{dname}.conv.conv2.0.bias")
        decoder_node_backward(dname, True)

    x_node_decoder("X02", "up_concat02", t["dXh1"])
    x_node_decoder("X12", "up_concat12", None, extra_deconv_from="up_concat03")
    x_node_decoder("X01", "up_concat01", t["dXh0"])
    x_node_decoder("X11", "up_concat11", None, extra_deconv_from="up_concat02")
    x_node_decoder("X21", "up_concat21", None, extra_deconv_from="up_concat12")

    # ---- encoder, deepest level first
    deconv_into = {3: "up_concat21", 2: "up_concat11", 1: "up_concat01"}
    for lvl in (3, 2, 1, 0):
        name = ENCODER[lvl]
        node = f"X{lvl}0"
        h, w, c = H >> lvl, W >> lvl, eng.filters[lvl]
        count = B * h * w
        bn2, bn1 = f"{name}.bn2", f"{name}.bn1"
        seq1, seq2 = _mod(m, f"{name}.conv1"), _mod(m, f"{name}.conv2")
        has_bn = m.is_batchnorm
        # with BatchNorm the gather produces dyh2 (gradient at the BN output, ReLU-masked) plus the two BN-backward sums;
        # without it (unet.py:137-143) the same launch produces dz2 directly and its per-channel sum is the conv bias gradient
        aux2 = dict(stats_aux=t[f"{name}.z2"], aux_mean=t[bn2 + ".mean"], aux_istd=t[bn2 + ".istd"]) if has_bn else {}
        addend = t[f"dpool{lvl}"] if lvl < 3 else None
        cons = consumers_of(node)
        dz2 = t[f"{name}.dz2"]
        dyh2 = t[f"{name}.dyh2"] if has_bn else dz2
        skey = "stats" if has_bn else f"stats:{name}.conv2.0.bias"
        if lvl == 0:
            key = f"{node}.gather"
            g, part = stats_buf([c] * len(cons), c, P[key + ".nt"], 9, h, w, key=skey)
            gather(node, h, w, dyh2, addend, t[node], dict(stats_partial=part, **aux2))
        elif cons:
            tmp = t[f"tmp{lvl}"]
            deconv_dgrad(deconv_into[lvl], h, w, tmp, addend=addend)
            key = f"{node}.gather"
            g, part = stats_buf([c] * len(cons), c, P[key + ".nt"], 9, h, w, key=skey)
            gather(node, h, w, dyh2, tmp, t[node], dict(stats_partial=part, **aux2))
        else:  # X30: the upsample of up_concat21 is its only consumer
            key = f"{deconv_into[lvl]}.up.dgrad"
            g, part = stats_buf(up_dgrad_srcs_C(deconv_into[lvl]), c, P[key + ".nt"], 1, h, w, key=skey)
            deconv_dgrad(deconv_into[lvl], h, w, dyh2, addend=addend, mask=t[node], stats=dict(stats_partial=part, **aux2))
        if has_bn:
            # BatchNorm 2 backward: sums = (dbeta, dgamma); dz = gamma*istd*(dyh - s1/M - xhat*s2/M)
            ops.reduce_partials(part, g, 2 * c, 2 * c, t[bn2 + ".sums"])
            ops.reduce_partials(t[bn2 + ".sums"], 1, 0, c, G, out_offset=goff(f"{name}.conv2.1.bias"), defer=True)
            ops.reduce_partials(t[bn2 + ".sums"], 1, 0, c, G, out_offset=goff(f"{name}.conv2.1.weight"), partial_offset=c, defer=True)
            # {name}.conv2.0.bias: a bias in front of BatchNorm has an exactly zero gradient; G is zero-initialised and never written there
            ops.bn_bwd_apply(dyh2, t[f"{name}.z2"], t[bn2 + ".mean"], t[bn2 + ".istd"], seq2[1].weight, t[bn2 + ".sums"], count, dz2)
        else:
            bias_from_stats(part, g, c, f"{name}.conv2.0.bias")
        wgrad_conv([t[f"{name}.a"]], dz2, h, w, f"{name}.conv2.0.weight")
        # first conv of the level
        key = f"{name}.c2.dgrad"
        dz1 = t[f"{name}.dz1"]
        if has_bn:
            g, part = stats_buf([c], c, P[key + ".nt"], 9, h, w)
            dyh1 = t[f"{name}.dyh1"]
            ops.conv([dz2], B, h, w, P[key], c, P[key + ".nt"], 9, out=dyh1, relu_mask_src=t[f"{name}.a"], stats_partial=part, stats_aux=t[f"{name}.z1"],
                     aux_mean=t[bn1 + ".mean"], aux_istd=t[bn1 + ".istd"])
            ops.reduce_partials(part, g, 2 * c, 2 * c, t[bn1 + ".sums"])
            ops.reduce_partials(t[bn1 + ".sums"], 1, 0, c, G, out_offset=goff(f"{name}.conv1.1.bias"), defer=True)
            ops.reduce_partials(t[bn1 + ".sums"], 1, 0, c, G, out_offset=goff(f"{name}.conv1.1.weight"), partial_offset=c, defer=True)
            ops.bn_bwd_apply(dyh1, t[f"{name}.z1"], t[bn1 + ".mean"], t[bn1 + ".istd"], seq1[1].weight, t[bn1 + ".sums"], count, dz1)
        else:
            g, part = stats_buf([c], c, P[key + ".nt"], 9, h, w, key=f"stats:{name}.conv1.0.bias")
            ops.conv([dz2], B, h, w, P[key], c, P[key + ".nt"], 9, out=dz1, relu_mask_src=t[f"{name}.a"], stats_partial=part)
            bias_from_stats(part, g, c, f"{name}.conv1.0.bias")
        if lvl == 0:
            wgrad_conv([t["x16"]], dz1, h, w, f"{name}.conv1.0.weight", ci_count=m.in_channels)
        else:
            pin = t[f"P{lvl - 1}0"]
            wgrad_conv([pin], dz1, h, w, f"{name}.conv1.0.weight")
            key = f"{name}.c1.dgrad"
            cprev = eng.filters[lvl - 1]
            ops.conv([dz1], B, h, w, P[key], cprev, P[key + ".nt"], 9, out=t[f"dP{lvl - 1}0"])
            ops.maxpool_bwd(t[f"X{lvl - 1}0"], t[f"dP{lvl - 1}0"], t[f"dpool{lvl - 1}"])
    if side is not None:
        cur.wait_stream(side)
    ops.flush_reduce_queue(ts.__dict__.setdefault("_reduce_tables", {}), eng.device)


# ---------------------------------------------------------------------------------------------- autograd boundary
def _train_state(eng: Engine, B: int, H: int, W: int) -> TrainState:
    """The cached state of this shape — or, while an earlier forward of the same shape still waits for its backward
    (``o1 = model(x1); o2 = model(x2); (l1 + l2).backward()``, which autograd supports), a fresh one that lives as long as
    the autograd node that owns it."""
    cache = eng.__dict__.setdefault("_train_states", {})
    ts = cache.get((B, H, W))
    if ts is not None and ts.busy():
        return TrainState(eng, B, H, W)
    if ts is None:
        if len(cache) >= 2:
            cache.pop(next(iter(cache)))
        ts = cache[(B, H, W)] = TrainState(eng, B, H, W)
    return ts


class _Token:
    """Lives exactly as long as the autograd node of one forward (ctx holds the only strong reference)."""
    __slots__ = ("done", "__weakref__")

    def __init__(self):
        self.done = False


def _prepare_dropout(ts: TrainState) -> None:
    m = ts.eng.model
    p = float(m.drop_out.p)
    forced = getattr(m, "_forced_dropout_masks", None)
    if forced is not None:  # parity runs: externally supplied keep-masks [B,16,H,W] (see tests)
        for k in range(3):
            ts.t[f"mask{k}"].copy_(ops.pack_keep_mask(forced[k].to(ts.eng.device)))
        ts.use_masks, ts.drop_scale = True, 1.0 / (1.0 - p)
    elif p > 0.0:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())  # torch's CPU generator: torch.manual_seed() reproduces the run
        for k in range(3):
            ops.dropout_mask(ts.t[f"mask{k}"].view(-1), p, seed + k)
        ts.use_masks, ts.drop_scale = True, 1.0 / (1.0 - p)
    else:
        ts.use_masks, ts.drop_scale = False, 1.0


class _UNetNestedFn(torch.autograd.Function):
    """forward: train-mode UNet_Nested on libunpp.so; backward: gradients for the 74 parameters from
    the three upstream heat-map gradients (trainer/trainer.py:127-135 builds its loss on CPU copies
    of the outputs, so arbitrary upstream gradients arrive here)."""

    @staticmethod
    def forward(ctx, eng, x, *params):
        B, H, W = eng._check_input(x)
        ts = _train_state(eng, B, H, W)
        x = x.contiguous()
        with torch.cuda.device(eng.device):
            pack_train(ts)
            _prepare_dropout(ts)
            heats = forward_train(ts, x)
        ctx.eng, ctx.ts, ctx.generation, ctx.token = eng, ts, ts.generation, _Token()
        ts.owner = weakref.ref(ctx.token)
        return heats

    @staticmethod
    def backward(ctx, *dheats):
        eng, ts = ctx.eng, ctx.ts
        if ts.generation != ctx.generation:
            raise RuntimeError("UNet_Nested backward: the activations saved by this forward were overwritten by a later train-mode forward of the "
                               "same shape (a second backward through a retained graph after another forward is not supported)")
        ctx.token.done = True
        G = torch.zeros(ts.nflat, dtype=torch.float32, device=eng.device)  # fresh buffer: p.grad may alias views of it
        dh = [None if d is None else d.contiguous().float() for d in dheats]
        with torch.cuda.device(eng.device):
            backward_train(ts, G, dheats=dh)
        grads = []
        for name, p in named_params(eng.model):
            off, n = ts.lay[name]
            grads.append(G[off:off + n].view_as(p) if p.requires_grad else None)
        return (None, None, *grads)


def run_autograd(eng: Engine, x: torch.Tensor):
    if eng.model.is_batchnorm and x.dim() == 4 and x.shape[0] * (x.shape[2] >> 3) * (x.shape[3] >> 3) <= 1:
        # nn.BatchNorm2d refuses batch statistics of a single value (torch/nn/functional.py:_verify_batch_size): the deepest level
        # (models/unet.py:262-263) sees B x H/8 x W/8 values per channel — mirror the reference's error instead of dividing by zero variance
        raise ValueError(f"Expected more than 1 value per channel when training, got input size {[x.shape[0], eng.filters[3], x.shape[2] >> 3, x.shape[3] >> 3]}")
    params = [p for _, p in named_params(eng.model)]
    if torch.is_grad_enabled() and any(p.requires_grad for p in params):
        return _UNetNestedFn.apply(eng, x, *params)
    # train mode without a graph (e.g. a no_grad warm-up): still batch-stat BN + dropout
    B, H, W = eng._check_input(x)
    ts = _train_state(eng, B, H, W)
    with torch.cuda.device(eng.device):
        pack_train(ts)
        _prepare_dropout(ts)
        return forward_train(ts, x.contiguous())


# ---------------------------------------------------------------------------------------------- smoke
def smoke_train_step(pkg, O, verify_step) -> None:
    """One small forward+backward on cuda:0 (used by __graft_entry__.smoke, which passes in the oracle modules: the product
    package never imports them).  Checked two ways: the loss against the oracle's fp32 autograd, and EVERY tensor the step
    stored plus all 74 parameter gradients against the teacher-forced fp64 recomputation (oracle/teacher_forced.py):
    bf16 tensors to one ulp, fp32 results to 1e-3."""
    sd = O.synth_state_dict(seed=12)
    model = pkg.UNet_Nested()
    model.load_state_dict(sd)
    model = model.to("cuda:0").train()
    model.drop_out.p = 0.0
    g = torch.Generator().manual_seed(6)
    B, H, W = 2, 64, 64
    x = torch.randn(B, 3, H, W, generator=g)
    target = torch.rand(B, 4, H, W, generator=g)
    outs = model(x.cuda())
    loss = sum(torch.nn.functional.mse_loss(o, target.cuda()) for o in outs) / 3
    loss.backward()
    torch.cuda.synchronize()
    rl, _, _, _ = O.train_step_grads(sd, x, target, dropout_masks=None)
    loss = loss.detach()
    assert abs(float(loss) - float(rl)) <= 1e-2 * abs(float(rl)), (float(loss), float(rl))
    ts = model._engine(torch.device("cuda", 0))._train_states[(B, H, W)]
    rep = verify_step(ts.t, ts.heats, {k: p.grad for k, p in model.named_parameters()}, sd, x.cuda(), target=target)
    bad = rep.check()
    assert not bad, f"teacher-forced mismatch: {bad[:3]}"
    wb, wf = rep.worst("bf16"), rep.worst("f32")
    print(f"smoke OK: train step loss {float(loss):.6f} (oracle {float(rl):.6f}); {len(rep.rows)} stored tensors / gradients teacher-forced: "
          f"worst bf16 tensor {wb['name']} {wb['max_rel']:.2e} of its range (0 elements beyond one ulp), worst fp32 result {wf['name']} {wf['max_rel']:.2e}")
