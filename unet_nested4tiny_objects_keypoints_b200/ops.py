"""Thin tensor-level wrappers over the C ABI of libunpp.so (include/unpp.h).

Every function takes CUDA torch tensors (PyTorch owns the memory), passes raw pointers and the
current CUDA stream across the boundary, and raises ``UnppError`` on a non-zero return code.
Nothing here computes anything in PyTorch: a missing library or a failing kernel is an exception,
never a fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import ConvArgs, PackArgs, ReduceJob, WgradArgs, MODE_CONV, MODE_DECONV  # noqa: F401

# Packed weights of one n_tile kept resident in shared memory.  148 KB admits n_tile = 64 for the 128-channel layers (147 KB): the kernel
# then stages 32-channel K chunks (three 21 KB stages) and issues N = 64 MMAs — 1.5-1.6x faster than two N = 32 slices (UNPP_WBUDGET: knob).
W_SMEM_BUDGET = int(os.environ.get("UNPP_WBUDGET", 148 * 1024))


# Per-thread recording state (nn.DataParallel drives one host thread per GPU: module globals would be shared between replicas):
#   _tls.pack_record  — when a list: pack_weights() appends its validated PackArgs instead of launching (see pack_batched)
#   _tls.reduce_queue — when a list: reductions called with defer=True are queued for ONE unpp_reduce_batched launch
_tls = threading.local()
launch_count = 0   # kernels of libunpp.so enqueued through this module (bench.py reads it for "gpu_launches")
trace = None       # when a list: every wrapper appends (label, start_event, end_event, bytes the launch moves, flops, plan-row tag)
trace_tag = ""     # row of the fused plan the following launches belong to, e.g. "up01.c1" (set by the engine; read by bench.py)


def _stream() -> int:
    # (the raw handle of the calling thread's current stream on its current device: ~10x cheaper than torch.cuda.current_stream(),
    # which every launch of an eager step used to pay — 130 launches per training step)
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def _count(n: int = 1) -> None:
    global launch_count
    launch_count += n


class _Traced:
    """Brackets one launch with CUDA events on the current stream when tracing is on."""

    def __init__(self, label, nbytes, flops):
        self.label, self.nbytes, self.flops = label, nbytes, flops

    def __enter__(self):
        if trace is not None:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if trace is not None:
            self.e1.record()
            trace.append((self.label, self.e0, self.e1, self.nbytes, self.flops, trace_tag))


def tag(name: str) -> None:
    """Name the plan row of the launches that follow (only read while tracing; a plain global store otherwise)."""
    global trace_tag
    trace_tag = name


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def lib():
    return _lib.load()


class NTile(int):
    """n_tile of a conv launch; ``NTile(16, b2=True)`` selects the 2x2 output-blocked kernel path
    (weights packed with pack_weights_b2) wherever an ``n_tile`` is passed to conv() / conv_grid()."""

    def __new__(cls, value, b2=False):
        obj = super().__new__(cls, value)
        obj.b2 = b2
        return obj


def pack_weights_b2(src: torch.Tensor, dgrad: bool, k_count: int, *, scale=None, n_begin: int = 0, k_begin: int = 0, dst=None, k8_total=None,
                    k_dst8: int = 0) -> torch.Tensor:
    """Weights of a 16-output-channel 3x3 conv for the 2x2 output-blocked kernel (unpp.h kinds 4 / 5)."""
    return pack_weights(src, 5 if dgrad else 4, 16, 64, 64, k_count, scale=scale, n_begin=n_begin, k_begin=k_begin, dst=dst, k8_total=k8_total,
                        k_dst8=k_dst8)


def pick_n_tile(n_total: int, k_total: int, taps: int, deconv: bool = False) -> int:
    """Largest multiple of 16 dividing n_total whose packed weights fit the shared-memory budget."""
    best = 16
    for nt in range(16, min(n_total, 256) + 1, 16):
        if n_total % nt:
            continue
        if taps * (k_total // 8) * nt * 16 > W_SMEM_BUDGET:
            continue
        if deconv and nt > 64:
            continue
        best = nt
    return best


def pack_weights(src: torch.Tensor, kind: int, taps: int, n_total: int, n_tile: int, k_count: int, *, scale=None, n_begin: int = 0,
                 k_begin: int = 0, dst: Optional[torch.Tensor] = None, k8_total: Optional[int] = None, k_dst8: int = 0) -> torch.Tensor:
    """fp32 weights -> bf16 UMMA B-operand layout [n_total/n_tile][taps][k8_total][n_tile][8] (see unpp.h)."""
    src = src.detach()
    if not src.is_contiguous():
        src = src.contiguous()
    assert src.dtype == torch.float32 and src.is_cuda
    k8_total = k_count // 8 if k8_total is None else k8_total
    if dst is None:
        dst = _zeros_bf16(n_total * taps * k8_total * 8, src.device)
    a = PackArgs()
    a.src, a.dst, a.scale = src.data_ptr(), dst.data_ptr(), _ptr(scale)
    a.kind, a.src_O, a.src_I, a.taps = kind, src.shape[0], src.shape[1], taps
    a.n_total, a.n_tile, a.n_begin = n_total, n_tile, n_begin
    a.k_begin, a.k_count, a.k8_total, a.k_dst8 = k_begin, k_count, k8_total, k_dst8
    rec = getattr(_tls, "pack_record", None)
    if rec is not None:
        rec.append((a, src))  # keep the source tensor alive with its job
        return dst
    _count()
    with _Traced("pack_weights", 0, 0):
        _lib.check(lib().unpp_pack_weights(C.byref(a), _stream()), "unpp_pack_weights")
    return dst


def _zeros_bf16(n: int, device) -> torch.Tensor:
    """Zero-filled bf16 buffer for one packed weight: a slice of the thread's pack arena when one is open (ONE fill launch for all the
    packed weights of a model instead of one per tensor), else its own allocation."""
    arena = getattr(_tls, "pack_arena", None)
    if arena is not None:
        buf, used = arena
        n_al = (n + 127) // 128 * 128  # 256-byte alignment of every slice (TMA bulk copies need 16)
        if used + n_al <= buf.numel() and buf.device == torch.device(device):
            arena[1] = used + n_al
            return buf[used:used + n]
    return torch.zeros(n, dtype=torch.bfloat16, device=device)


class pack_arena:
    """``with ops.pack_arena(device, elems): ...`` — packed-weight buffers allocated inside are slices of one zero-filled tensor."""

    def __init__(self, device, elems: int = 4 << 20):
        self.device, self.elems = device, elems

    def __enter__(self):
        self.prev = getattr(_tls, "pack_arena", None)
        _tls.pack_arena = [torch.zeros(self.elems, dtype=torch.bfloat16, device=self.device), 0]
        return self

    def __exit__(self, *exc):
        _tls.pack_arena = self.prev


def begin_pack_record() -> None:
    _tls.pack_record = []


def end_pack_record():
    jobs, _tls.pack_record = getattr(_tls, "pack_record", None), None
    return jobs


def make_pack_table(jobs, device) -> torch.Tensor:
    """Device-resident table of PackArgs (for pack_batched) from the jobs recorded through ``pack_record``."""
    raw = b"".join(bytes(a) for a, _ in jobs)
    return torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)


def pack_batched(table: torch.Tensor, n: int) -> None:
    _count()
    with _Traced("pack_weights_batched", 0, 0):
        _lib.check(lib().unpp_pack_weights_batched(table.data_ptr(), n, _stream()), "unpp_pack_weights_batched")


def _conv_args(srcs, N, H, W, n_total, n_tile, taps, strided=None, b2=False) -> ConvArgs:
    a = ConvArgs()
    a.N, a.H, a.W, a.nsrc = N, H, W, len(srcs)
    for i, s in enumerate(srcs):
        if isinstance(s, torch.Tensor):
            a.src[i] = s.data_ptr()
            a.src_C[i] = s.shape[-1]
        else:  # geometry-only query: channel count
            a.src[i] = 16
            a.src_C[i] = int(s)
        if strided is not None:
            a.src_step[i], a.src_oy[i], a.src_ox[i] = 2, strided[i][0], strided[i][1]
    a.taps, a.n_total, a.n_tile = taps, n_total, int(n_tile)
    a.block2x2 = int(b2 or getattr(n_tile, "b2", False))
    return a


def conv(srcs: Sequence[torch.Tensor], N: int, H: int, W: int, wpacked: torch.Tensor, n_total: int, n_tile: int, taps: int, *, bias=None,
         relu=False, mode=MODE_CONV, out=None, head=None, addend=None, relu_mask_src=None, stats_partial=None, stats_aux=None, aux_mean=None,
         aux_istd=None, strided=None, b2=False, lowres=None, bias_classes=0, pooled=None) -> None:
    """One unpp_conv_tc launch.  ``srcs``: NHWC bf16 tensors (virtual concat along K).
    ``head`` = (w fp32 [cls,16], b fp32 [cls], heat fp32 NCHW, logit|None, drop_mask int16 [N,H,W] keep bits|None, drop_scale).
    ``strided`` = [(oy, ox), ...]: every source is a [N,2H,2W,C] tensor read at (2y+oy, 2x+ox)."""
    a = _conv_args(srcs, N, H, W, n_total, n_tile, taps, strided, b2)
    a.wpacked, a.bias = wpacked.data_ptr(), _ptr(bias)
    a.mode, a.relu = mode, int(relu)
    a.out = _ptr(out)
    if head is not None:
        hw, hb, heat, logit, dmask, dscale = head
        a.head_w, a.head_b, a.heat, a.logit = hw.data_ptr(), hb.data_ptr(), heat.data_ptr(), _ptr(logit)
        a.head_classes = hw.shape[0]
        a.drop_mask, a.drop_scale = _ptr(dmask), float(dscale)
    if lowres is not None:  # (low-resolution tensor [N,H/2,W/2,C], composed packed weights): fused transposed conv
        a.lowres_src, a.lowres_wpacked, a.lowres_C = lowres[0].data_ptr(), lowres[1].data_ptr(), lowres[0].shape[-1]
    a.bias_classes = bias_classes
    a.pooled = _ptr(pooled)  # fused MaxPool2d(2) of the output (inference epilogue)
    a.addend, a.relu_mask_src = _ptr(addend), _ptr(relu_mask_src)
    a.stats_partial, a.stats_aux, a.aux_mean, a.aux_istd = _ptr(stats_partial), _ptr(stats_aux), _ptr(aux_mean), _ptr(aux_istd)
    _count()
    if trace is None:
        _lib.check(lib().unpp_conv_tc(C.byref(a), _stream()), "unpp_conv_tc")
        return
    # algorithmic traffic: every source once, the output once, weights once; heat maps fp32
    k_total = sum(s.shape[-1] for s in srcs)
    px = N * H * W
    # (2x2-blocked layouts store only the non-zero blocks: 36 of 64 per 16 input channels, 16 of 36 for the fused transposed conv)
    nbytes = px * k_total * 2 + ((k_total // 16) * 36 * 512 if a.block2x2 else wpacked.numel() * 2)
    if lowres is not None:
        nbytes += lowres[0].numel() * 2 + (lowres[0].shape[-1] // 16) * 16 * 512
    if out is not None:
        nbytes += px * n_total * 2
    if pooled is not None:
        nbytes += px * n_total * 2 // 4
    if head is not None:
        nbytes += px * head[0].shape[0] * 4 + (px * 16 if head[4] is not None else 0)
    for extra in (addend, relu_mask_src, stats_aux):
        if extra is not None:
            nbytes += extra.numel() * 2
    flops = 2 * px * k_total * n_total * taps
    if lowres is not None:  # count the unfused arithmetic it replaces: the k2s2 transposed conv + its 3x3 taps
        flops += 2 * px * lowres[0].shape[-1] * n_total + 2 * px * n_total * n_total * taps
    label = "conv_tc %s taps%d K%d N%d %dx%d" % ("deconv" if mode == MODE_DECONV else ("conv2x2" if a.block2x2 else "conv"), taps, k_total, n_total, H, W)
    flags = [n for n, v in (("st", stats_partial), ("aux", stats_aux), ("mask", relu_mask_src), ("add", addend), ("head", head), ("s2", strided), ("low", lowres)) if v is not None]
    if flags:
        label += " +" + "+".join(flags)
    with _Traced(label, nbytes, flops):
        _lib.check(lib().unpp_conv_tc(C.byref(a), _stream()), "unpp_conv_tc")


def conv_grid(srcs_C: Sequence[int], N: int, H: int, W: int, n_total: int, n_tile: int, taps: int, b2: bool = False) -> int:
    a = _conv_args(list(srcs_C), N, H, W, n_total, n_tile, taps, None, b2)
    a.stats_partial = 16  # the grid is asked for to size a statistics buffer: plan like the training-epilogue launch that will fill it
    g = lib().unpp_conv_grid(C.byref(a))
    if g < 0:
        _lib.check(g, "unpp_conv_grid")
    return g


def nchw_to_nhwc16(x: torch.Tensor, out: torch.Tensor) -> None:
    B, Cin, H, W = x.shape
    _count()
    with _Traced("nchw_to_nhwc", 0, 0):
        _lib.check(lib().unpp_nchw_to_nhwc(x.data_ptr(), out.data_ptr(), B, Cin, H, W, 16, _stream()), "unpp_nchw_to_nhwc")


def nchw_to_nhwc4(x: torch.Tensor, out: torch.Tensor) -> None:
    """fp32 NCHW (C <= 4) -> bf16 NHWC with 4 channels per pixel (8 B): the input of the first-layer conv mode (NTile(16, b2=2))."""
    B, Cin, H, W = x.shape
    assert tuple(out.shape) == (B, H, W, 4) and out.dtype == torch.bfloat16 and out.is_contiguous()
    _count()
    with _Traced("nchw_to_nhwc4", x.numel() * 4 + out.numel() * 2, 0):
        _lib.check(lib().unpp_nchw_to_nhwc(x.data_ptr(), out.data_ptr(), B, Cin, H, W, 4, _stream()), "unpp_nchw_to_nhwc")


def bn_fold(bn, conv_bias: torch.Tensor):
    """Eval-mode nn.BatchNorm2d folded into the conv in front of it: (scale fp32 [C] for pack_weights, bias fp32 [C])."""
    c = conv_bias.numel()
    out = torch.empty(2, c, dtype=torch.float32, device=conv_bias.device)
    _count()
    _lib.check(lib().unpp_bn_fold(bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(), bn.running_var.data_ptr(), conv_bias.data_ptr(), float(bn.eps), c,
                                  out[0].data_ptr(), out[1].data_ptr(), _stream()), "unpp_bn_fold")
    return out[0], out[1]


def compose_deconv_conv(w_conv: torch.Tensor, cu: int, w_up: torch.Tensor, b_up: torch.Tensor, b_conv: torch.Tensor):
    """conv3x3(ConvTranspose2d_k2s2(x)) as one 3x3 conv over the low-resolution x (unpp_compose_deconv_conv): returns the composed
    weight fp32 [4*Co, Ci, 3, 3] and the (row class, column class) bias table fp32 [9, Co]."""
    w_conv, w_up, b_up, b_conv = (t.detach().float().contiguous() for t in (w_conv, w_up, b_up, b_conv))
    co, ctot, ci = w_conv.shape[0], w_conv.shape[1], w_up.shape[0]
    assert w_up.shape[1] == cu and b_up.numel() == cu and b_conv.numel() == co and w_conv.is_cuda
    comp = torch.empty(4 * co, ci, 3, 3, dtype=torch.float32, device=w_conv.device)
    table = torch.empty(9, co, dtype=torch.float32, device=w_conv.device)
    _count()
    _lib.check(lib().unpp_compose_deconv_conv(w_conv.data_ptr(), ctot, cu, w_up.data_ptr(), b_up.data_ptr(), b_conv.data_ptr(), co, ci, comp.data_ptr(),
                                              table.data_ptr(), _stream()), "unpp_compose_deconv_conv")
    return comp, table


def u8_to_nhwc(x: torch.Tensor, out: torch.Tensor, channels_last: bool) -> None:
    """uint8 images ([B,H,W,C] when ``channels_last`` else [B,C,H,W]) -> bf16 NHWC [B,H,W,Cpad] (Cpad = out.shape[-1] in {4,16}) scaled by 1/255
    like torchvision's ToTensor (datasets/datasets_base.py:71-72)."""
    assert x.dtype == torch.uint8 and x.is_cuda and x.is_contiguous() and out.dtype == torch.bfloat16 and out.is_contiguous()
    if channels_last:
        B, H, W, Cin = x.shape
    else:
        B, Cin, H, W = x.shape
    cpad = out.shape[-1]
    assert tuple(out.shape) == (B, H, W, cpad)
    _count()
    with _Traced("u8_to_nhwc%d" % cpad, x.numel() + out.numel() * 2, 0):
        _lib.check(lib().unpp_u8_to_nhwc(x.data_ptr(), out.data_ptr(), B, Cin, H, W, cpad, int(channels_last), _stream()), "unpp_u8_to_nhwc")


def pack_weights_c4(src: torch.Tensor, scale=None) -> torch.Tensor:
    """Weights [16][Cin <= 4][3][3] of the network's first conv for the 4-channel first-layer mode (unpp.h kind 7)."""
    return pack_weights(src, 7, 4, 64, 64, 32, scale=scale)


def maxpool(x: torch.Tensor, out: torch.Tensor) -> None:
    B, H, W, Cc = x.shape
    _count()
    with _Traced("maxpool2x2", 0, 0):
        _lib.check(lib().unpp_maxpool2x2(x.data_ptr(), out.data_ptr(), B, H, W, Cc, _stream()), "unpp_maxpool2x2")


def argmax_peaks(heat: torch.Tensor):
    """fp32 [B,C,H,W] -> (xy int32 [B,C,2] as [x,y], peak fp32 [B,C]); reference tools/misc/heatmap.py:173-178."""
    if heat.dtype != torch.float32 or not heat.is_cuda or heat.dim() != 4:
        raise ValueError("argmax_peaks expects an fp32 CUDA tensor [B,C,H,W]")
    heat = heat.contiguous()
    B, Cc, H, W = heat.shape
    xy = torch.empty(B, Cc, 2, dtype=torch.int32, device=heat.device)
    val = torch.empty(B, Cc, dtype=torch.float32, device=heat.device)
    splits = lib().unpp_argmax_splits(B * Cc, H, W)  # few large planes (1024x1024 at batch 16): several CTAs per plane + a fold launch
    nbytes = heat.numel() * 4
    if splits > 1:
        ws = torch.empty(B * Cc * splits * 2, dtype=torch.float32, device=heat.device)
        _count(2)
        with _Traced("argmax_peaks_split", nbytes, 0):
            _lib.check(lib().unpp_argmax_peaks_split(heat.data_ptr(), B * Cc, H, W, xy.data_ptr(), val.data_ptr(), ws.data_ptr(), splits, _stream()),
                       "unpp_argmax_peaks_split")
        return xy, val
    _count()
    with _Traced("argmax_peaks", nbytes, 0):
        _lib.check(lib().unpp_argmax_peaks(heat.data_ptr(), B * Cc, H, W, xy.data_ptr(), val.data_ptr(), _stream()), "unpp_argmax_peaks")
    return xy, val


def topk_peaks(heat: torch.Tensor, num: int, threshold: float = 0.5):
    """fp32 [B,C,H,W] -> (xy int32 [B,C,num,2] as [x,y] (-1 where a plane has fewer peaks), peak fp32 [B,C,num], count int32 [B,C]):
    the ``num`` brightest local maxima above ``threshold`` per plane — Heatmap.extract_points_(pred, num) of tools/misc/heatmap.py:148-208."""
    if heat.dtype != torch.float32 or not heat.is_cuda or heat.dim() != 4:
        raise ValueError("topk_peaks expects an fp32 CUDA tensor [B,C,H,W]")
    heat = heat.contiguous()
    B, Cc, H, W = heat.shape
    xy = torch.empty(B, Cc, num, 2, dtype=torch.int32, device=heat.device)
    val = torch.empty(B, Cc, num, dtype=torch.float32, device=heat.device)
    cnt = torch.empty(B, Cc, dtype=torch.int32, device=heat.device)
    _count()
    with _Traced("topk_peaks", heat.numel() * 4, 0):
        _lib.check(lib().unpp_topk_peaks(heat.data_ptr(), B * Cc, H, W, int(num), float(threshold), xy.data_ptr(), val.data_ptr(), cnt.data_ptr(), _stream()),
                   "unpp_topk_peaks")
    return xy, val, cnt


# ---------------------------------------------------------------------------------------------- training
def _wgrad_args(srcs, N, H, W, dz, cout, taps, dz_view=None) -> WgradArgs:
    a = WgradArgs()
    a.N, a.H, a.W, a.nsrc = N, H, W, len(srcs)
    for i, s in enumerate(srcs):
        if isinstance(s, torch.Tensor):
            a.src[i], a.src_C[i] = s.data_ptr(), s.shape[-1]
        else:
            a.src[i], a.src_C[i] = 16, int(s)
    a.dz = 16 if dz is None else dz.data_ptr()
    a.cout, a.taps = cout, taps
    if dz_view is None:
        a.dz_step, a.dz_oy, a.dz_ox = 1, 0, 0
    elif dz_view == "all4":  # the four taps of a k2s2 transposed conv in one launch: partial [4][grid][1][cin][cout]
        a.dz_step, a.dz_oy, a.dz_ox = 2, -1, -1
    else:
        a.dz_step, a.dz_oy, a.dz_ox = 2, dz_view[0], dz_view[1]
    return a


def wgrad_grid(srcs_C: Sequence[int], N: int, H: int, W: int, cout: int, taps: int, dz_view=None) -> int:
    a = _wgrad_args(list(srcs_C), N, H, W, None, cout, taps, dz_view)
    g = lib().unpp_wgrad_grid(C.byref(a))
    if g < 0:
        _lib.check(g, "unpp_wgrad_grid")
    return g


def wgrad(srcs: Sequence[torch.Tensor], N: int, H: int, W: int, dz: torch.Tensor, cout: int, taps: int, partial: torch.Tensor, dz_view=None) -> None:
    a = _wgrad_args(srcs, N, H, W, dz, cout, taps, dz_view)
    a.partial = partial.data_ptr()
    _count()
    k_total = sum(s.shape[-1] for s in srcs)
    px = N * H * W
    nz = 4 if dz_view == "all4" else 1
    with _Traced("wgrad taps%d K%d N%d %dx%d%s" % (taps, k_total, cout, H, W, " x4" if nz == 4 else ""), px * (k_total + nz * cout) * 2,
                 2 * px * k_total * cout * taps * nz):
        _lib.check(lib().unpp_wgrad(C.byref(a), _stream()), "unpp_wgrad")


def _queue_reduce(partial_ptr, dst_ptr, stride, s_co, s_ci, s_tap, nparts, taps, cin_total, cout, ci_begin, ci_count, scale, keep) -> None:
    j = ReduceJob()
    j.partial, j.dst, j.stride, j.s_co, j.s_ci, j.s_tap = partial_ptr, dst_ptr, stride, s_co, s_ci, s_tap
    j.nparts, j.taps, j.cin_total, j.cout, j.ci_begin, j.ci_count, j.scale = nparts, taps, cin_total, cout, ci_begin, ci_count, scale
    _tls.reduce_queue.append((j, keep))


def begin_reduce_queue() -> None:
    """From here until flush_reduce_queue() reductions called with defer=True on THIS thread are queued, not launched."""
    _tls.reduce_queue = []


def reduce_queue_active() -> bool:
    return getattr(_tls, "reduce_queue", None) is not None


def flush_reduce_queue(cache: dict, device) -> None:
    """Launch every queued reduction as one unpp_reduce_batched and switch queueing off.  The device-resident job
    table is cached by content (pointers are stable across steps), so a captured step replays without host work."""
    jobs, _tls.reduce_queue = getattr(_tls, "reduce_queue", None), None
    if not jobs:
        return
    blocks = 0
    for j, _ in jobs:
        blocks += (j.taps * j.ci_count * j.cout + 127) // 128
        j.block_end = blocks
    raw = b"".join(bytes(j) for j, _ in jobs)
    table = cache.get(raw)
    if table is None:
        if len(cache) >= 8:  # the autograd path may see a new gradient buffer address per call
            cache.clear()
        table = cache[raw] = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)
    _count()
    with _Traced("reduce_batched", 0, 0):
        _lib.check(lib().unpp_reduce_batched(table.data_ptr(), len(jobs), blocks, _stream()), "unpp_reduce_batched")


def wgrad_reduce(partial: torch.Tensor, nparts: int, taps: int, cin_total: int, cout: int, dst: torch.Tensor, ci_begin: int, ci_count: int,
                 s_co: int, s_ci: int, s_tap: int, scale: float = 1.0, dst_offset: int = 0, partial_offset: int = 0, defer: bool = False) -> None:
    if defer and reduce_queue_active():
        _queue_reduce(partial.data_ptr() + 4 * partial_offset, dst.data_ptr() + 4 * dst_offset, taps * cin_total * cout, s_co, s_ci, s_tap, nparts, taps,
                      cin_total, cout, ci_begin, ci_count, scale, (partial, dst))
        return
    _count()
    with _Traced("wgrad_reduce", 0, 0):
        _lib.check(lib().unpp_wgrad_reduce(partial.data_ptr() + 4 * partial_offset, nparts, taps, cin_total, cout, dst.data_ptr() + 4 * dst_offset, ci_begin, ci_count,
                                       s_co, s_ci, s_tap, scale, _stream()), "unpp_wgrad_reduce")


def reduce_partials(partial: torch.Tensor, nparts: int, stride: int, n: int, out: torch.Tensor, scale: float = 1.0, accumulate: bool = False,
                    partial_offset: int = 0, out_offset: int = 0, defer: bool = False) -> None:
    if defer and reduce_queue_active() and not accumulate:
        _queue_reduce(partial.data_ptr() + 4 * partial_offset, out.data_ptr() + 4 * out_offset, stride, 1, 0, 0, nparts, 1, 1, n, 0, 1, scale, (partial, out))
        return
    _count()
    with _Traced("reduce_partials", 0, 0):
        _lib.check(lib().unpp_reduce_partials(partial.data_ptr() + 4 * partial_offset, nparts, stride, n, scale, out.data_ptr() + 4 * out_offset,
                                          int(accumulate), _stream()), "unpp_reduce_partials")


def bn_finalize(partial, nparts, Cc, count, gamma, beta, running_mean, running_var, momentum, eps, mean, istd, scale, shift) -> None:
    _count()
    with _Traced("bn_finalize", 0, 0):
        _lib.check(lib().unpp_bn_finalize(partial.data_ptr(), nparts, Cc, float(count), gamma.data_ptr(), beta.data_ptr(), _ptr(running_mean),
                                      _ptr(running_var), float(momentum), float(eps), mean.data_ptr(), istd.data_ptr(), scale.data_ptr(),
                                      shift.data_ptr(), _stream()), "unpp_bn_finalize")


def bn_relu(z, scale, shift, y, pooled=None) -> None:
    N, H, W, Cc = z.shape
    _count()
    with _Traced("bn_relu", 0, 0):
        _lib.check(lib().unpp_bn_relu(z.data_ptr(), scale.data_ptr(), shift.data_ptr(), y.data_ptr(), _ptr(pooled), N, H, W, Cc, _stream()), "unpp_bn_relu")


def maxpool_bwd(x, dpooled, dx) -> None:
    N, H, W, Cc = x.shape
    _count()
    with _Traced("maxpool2x2_bwd", 0, 0):
        _lib.check(lib().unpp_maxpool2x2_bwd(x.data_ptr(), dpooled.data_ptr(), dx.data_ptr(), N, H, W, Cc, _stream()), "unpp_maxpool2x2_bwd")


def bn_bwd_apply(dyh, z, mean, istd, gamma, sums, count, dz) -> None:
    N, H, W, Cc = z.shape
    _count()
    with _Traced("bn_bwd_apply", 0, 0):
        _lib.check(lib().unpp_bn_bwd_apply(dyh.data_ptr(), z.data_ptr(), mean.data_ptr(), istd.data_ptr(), gamma.data_ptr(), sums.data_ptr(),
                                       float(count), dz.data_ptr(), N, H, W, Cc, _stream()), "unpp_bn_bwd_apply")


def head_bwd_grid(N: int, H: int, W: int) -> int:
    return lib().unpp_head_bwd_grid(N, H, W)


def head_bwd(heat, dheat, target, coef, x, drop_mask, drop_scale, head_w, dx, partial, loss_kind: int = 0, gamma: float = 3.0) -> None:
    N, ncls, H, W = heat.shape
    _count()
    with _Traced("head_bwd", 0, 0):
        _lib.check(lib().unpp_head_bwd(heat.data_ptr(), _ptr(dheat), _ptr(target), int(loss_kind), float(gamma), float(coef), x.data_ptr(), _ptr(drop_mask), float(drop_scale),
                                   head_w.data_ptr(), ncls, dx.data_ptr(), partial.data_ptr(), N, H, W, _stream()), "unpp_head_bwd")


def adamw(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0) -> None:
    """tools/optimizers/adamw.py:38-100 on flat fp32 buffers (in place)."""
    for t in (p, g, m, v):
        assert t.dtype == torch.float32 and t.is_cuda and t.is_contiguous() and t.numel() == p.numel()
    _count()
    with _Traced("adamw", 0, 0):
        _lib.check(lib().unpp_adamw(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), float(lr), float(beta1), float(beta2), float(eps),
                                float(weight_decay), int(step), float(grad_scale), _stream()), "unpp_adamw")


def adamw_dev(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step_counter, step_size_scratch, grad_scale=1.0) -> None:
    """Same update with the step counter (int64 [1]) and step size (fp32 [1]) living on the device."""
    for t in (p, g, m, v):
        assert t.dtype == torch.float32 and t.is_cuda and t.is_contiguous() and t.numel() == p.numel()
    assert step_counter.dtype == torch.int64 and step_size_scratch.dtype == torch.float32
    _count(2)
    _lib.check(lib().unpp_adamw_dev(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), float(lr), float(beta1), float(beta2),
                                    float(eps), float(weight_decay), step_counter.data_ptr(), step_size_scratch.data_ptr(), float(grad_scale),
                                    _stream()), "unpp_adamw_dev")


def dropout_mask(mask: torch.Tensor, p_drop: float, seed: int, step_counter: Optional[torch.Tensor] = None) -> None:
    """Fill ``mask`` (int16, one word per pixel: bit c = keep channel c of the 16-channel head input) for nn.Dropout(p_drop)."""
    assert mask.dtype == torch.int16 and mask.is_cuda and mask.is_contiguous()
    _count()
    with _Traced("dropout_mask", 0, 0):
        _lib.check(lib().unpp_dropout_mask(mask.data_ptr(), mask.numel(), float(p_drop), int(seed) & (2**64 - 1), _ptr(step_counter), _stream()),
                   "unpp_dropout_mask")


def pack_keep_mask(keep: torch.Tensor) -> torch.Tensor:
    """Boolean / 0-1 keep-mask [N,16,H,W] (the layout nn.Dropout sees) -> the kernels' int16 [N,H,W] bit words."""
    assert keep.dim() == 4 and keep.shape[1] == 16
    w = torch.zeros(keep.shape[0], keep.shape[2], keep.shape[3], dtype=torch.int32, device=keep.device)
    for c in range(16):
        w |= (keep[:, c] != 0).to(torch.int32) << c
    return (w - ((w >> 15) << 16)).to(torch.int16)  # two's-complement wrap of bit 15


def unpack_keep_mask(words: torch.Tensor) -> torch.Tensor:
    """int16 [N,H,W] bit words -> bool [N,16,H,W]."""
    w = words.to(torch.int32) & 0xFFFF
    return torch.stack([((w >> c) & 1).bool() for c in range(16)], 1)


def create_heatmap(keypoints: torch.Tensor, H: int, W: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Target heat maps on the device (reference tools/misc/helper.py:87-172): keypoints fp32 [N, 7, 2] as (x, y) -> fp32 [N, 4, H, W]."""
    if keypoints.dtype != torch.float32 or not keypoints.is_cuda or keypoints.dim() != 3 or keypoints.shape[2] != 2:
        raise ValueError("create_heatmap expects an fp32 CUDA tensor [N, points, 2]")
    keypoints = keypoints.contiguous()
    N, npts = keypoints.shape[0], keypoints.shape[1]
    if out is None:
        out = torch.empty(N, 4, H, W, dtype=torch.float32, device=keypoints.device)
    _count()
    _lib.check(lib().unpp_create_heatmap(keypoints.data_ptr(), N, npts, H, W, out.data_ptr(), _stream()), "unpp_create_heatmap")
    return out


# ---------------------------------------------------------------------------------------------- non-default constructor flags / optimizers
OPTIMIZER_KINDS = {"adamw": _lib.OPT_ADAMW, "adam": _lib.OPT_ADAM, "adabound": _lib.OPT_ADABOUND, "sgd": _lib.OPT_SGD, "sgdw": _lib.OPT_SGDW}


def bilinear_up2x(x: torch.Tensor, out: torch.Tensor) -> None:
    """nn.UpsamplingBilinear2d(scale_factor=2) (align_corners=True, models/unet.py:190) on NHWC bf16: [N,H,W,C] -> [N,2H,2W,C]."""
    N, H, W, Cc = x.shape
    assert x.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and tuple(out.shape) == (N, 2 * H, 2 * W, Cc) and x.is_contiguous() and out.is_contiguous()
    _count()
    with _Traced("bilinear_up2x %dx%d C%d" % (H, W, Cc), x.numel() * 2 + out.numel() * 2, 0):
        _lib.check(lib().unpp_bilinear_up2x(x.data_ptr(), out.data_ptr(), N, H, W, Cc, _stream()), "unpp_bilinear_up2x")


def bilinear_up2x_bwd(dy: torch.Tensor, dx: torch.Tensor) -> None:
    """Adjoint of bilinear_up2x: dy [N,2H,2W,C] -> dx [N,H,W,C]."""
    N, H, W, Cc = dx.shape
    assert dy.dtype == torch.bfloat16 and dx.dtype == torch.bfloat16 and tuple(dy.shape) == (N, 2 * H, 2 * W, Cc) and dy.is_contiguous() and dx.is_contiguous()
    _count()
    with _Traced("bilinear_up2x_bwd %dx%d C%d" % (H, W, Cc), dy.numel() * 2 + dx.numel() * 2, 0):
        _lib.check(lib().unpp_bilinear_up2x_bwd(dy.data_ptr(), dx.data_ptr(), N, H, W, Cc, _stream()), "unpp_bilinear_up2x_bwd")


def optim_step(kind: str, p, g, state1, state2, *, lr, beta1=0.0, beta2=0.0, eps=1e-8, weight_decay=0.0, final_lr=0.1, gamma=1e-3, base_lr=None,
               grad_scale=1.0, step: int = 0, step_counter: Optional[torch.Tensor] = None, lr_dev: Optional[torch.Tensor] = None,
               scalars: Optional[torch.Tensor] = None) -> None:
    """One flat-buffer update of any optimizer trainer/trainer.py:344-376 can select (see UNPP_OPT_* in unpp.h).
    Either ``step`` (1-based host count) or ``step_counter`` (int64 [1] on the device, incremented by the call; then ``scalars`` is
    an fp32 [4] device scratch and ``lr_dev`` an optional fp32 [1] device learning rate that overrides ``lr``)."""
    for t in (p, g, state1, state2):
        assert t is None or (t.dtype == torch.float32 and t.is_cuda and t.is_contiguous() and t.numel() == p.numel())
    a = _lib.OptimArgs()
    a.kind, a.lr, a.beta1, a.beta2, a.eps, a.weight_decay = OPTIMIZER_KINDS[kind], float(lr), float(beta1), float(beta2), float(eps), float(weight_decay)
    a.final_lr, a.gamma, a.base_lr, a.grad_scale = float(final_lr), float(gamma), float(lr if base_lr is None else base_lr), float(grad_scale)
    _count(2 if step_counter is not None else 1)
    with _Traced("optim_step " + kind, 0, 0):
        _lib.check(lib().unpp_optim_step(p.data_ptr(), g.data_ptr(), _ptr(state1), _ptr(state2), p.numel(), C.byref(a), int(step), _ptr(step_counter),
                                         _ptr(lr_dev), _ptr(scalars), _stream()), "unpp_optim_step")


def optim_step_multi(kind: str, tensors, *, step: int, lr, beta1=0.0, beta2=0.0, eps=1e-8, weight_decay=0.0, final_lr=0.1, gamma=1e-3, base_lr=None,
                     grad_scale=1.0) -> None:
    """One launch updating many tensors: ``tensors`` = [(p, g, state1 | None, state2 | None), ...], contiguous fp32 CUDA tensors on one device,
    all at the same (1-based) ``step``.  The 3 KB pointer table travels host -> device with the call."""
    raw, blocks = [], 0
    for p, g, s1, s2 in tensors:
        for t in (p, g, s1, s2):
            assert t is None or (t.dtype == torch.float32 and t.is_cuda and t.is_contiguous() and t.numel() == p.numel())
        e = _lib.OptimTensor()
        e.p, e.g, e.state1, e.state2, e.n = p.data_ptr(), g.data_ptr(), _ptr(s1), _ptr(s2), p.numel()
        blocks += (p.numel() + 1023) // 1024
        e.block_end = blocks
        raw.append(bytes(e))
    table = torch.frombuffer(bytearray(b"".join(raw)), dtype=torch.uint8).to(tensors[0][0].device)
    a = _lib.OptimArgs()
    a.kind, a.lr, a.beta1, a.beta2, a.eps, a.weight_decay = OPTIMIZER_KINDS[kind], float(lr), float(beta1), float(beta2), float(eps), float(weight_decay)
    a.final_lr, a.gamma, a.base_lr, a.grad_scale = float(final_lr), float(gamma), float(lr if base_lr is None else base_lr), float(grad_scale)
    _count()
    with _Traced("optim_step_multi " + kind, 0, 0):
        _lib.check(lib().unpp_optim_step_multi(table.data_ptr(), len(tensors), blocks, C.byref(a), int(step), _stream()), "unpp_optim_step_multi")


# ---------------------------------------------------------------------------------------------- fp32 validation mode (csrc/ref_kernels.cu)
def ref_conv(srcs: Sequence[torch.Tensor], weight: torch.Tensor, bias, *, scale=None, relu=False, sigmoid=False) -> torch.Tensor:
    """conv (3x3 pad 1, or 1x1) of the channel-concatenation of NCHW fp32 ``srcs`` with the OIHW fp32 ``weight``, then
    ``* scale + bias`` per output channel, ReLU and / or sigmoid: fp32 FMAs on the CUDA cores."""
    N, _, H, W = srcs[0].shape
    cout, ctot, k, _ = weight.shape
    assert sum(s.shape[1] for s in srcs) == ctot and k in (1, 3)
    out = torch.empty(N, cout, H, W, dtype=torch.float32, device=srcs[0].device)
    a = _lib.RefConvArgs()
    a.N, a.H, a.W, a.nsrc = N, H, W, len(srcs)
    for i, s in enumerate(srcs):
        assert s.dtype == torch.float32 and s.is_contiguous() and tuple(s.shape[2:]) == (H, W)
        a.src[i], a.src_C[i] = s.data_ptr(), s.shape[1]
    weight = weight.detach().float().contiguous()
    a.weight, a.scale, a.bias = weight.data_ptr(), _ptr(scale), _ptr(bias)
    a.cout, a.taps, a.relu, a.sigmoid, a.out = cout, k * k, int(relu), int(sigmoid), out.data_ptr()
    _count()
    _lib.check(lib().unpp_ref_conv(C.byref(a), _stream()), "unpp_ref_conv")
    return out


def ref_deconv2x2(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    N, cin, H, W = x.shape
    cout = weight.shape[1]
    out = torch.empty(N, cout, 2 * H, 2 * W, dtype=torch.float32, device=x.device)
    weight, bias = weight.detach().float().contiguous(), bias.detach().float().contiguous()
    _count()
    _lib.check(lib().unpp_ref_deconv2x2(x.data_ptr(), weight.data_ptr(), bias.data_ptr(), out.data_ptr(), N, cin, cout, H, W, _stream()), "unpp_ref_deconv2x2")
    return out


def ref_maxpool2x2(x: torch.Tensor) -> torch.Tensor:
    N, Cc, H, W = x.shape
    out = torch.empty(N, Cc, H // 2, W // 2, dtype=torch.float32, device=x.device)
    _count()
    _lib.check(lib().unpp_ref_maxpool2x2(x.data_ptr(), out.data_ptr(), N, Cc, H, W, _stream()), "unpp_ref_maxpool2x2")
    return out
