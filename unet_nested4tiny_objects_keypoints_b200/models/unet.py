"""Drop-in ``UNet_Nested`` (UNet++) whose forward/backward run on hand-written sm_100a kernels.

API contract kept from the reference (file:line into the reference repository):
  * constructor ``UNet_Nested(in_channels=3, n_classes=4, feature_scale=2, is_deconv=True,
    is_batchnorm=True, is_ds=True)`` — models/unet.py:206; built with no arguments by the trainer
    (trainer/trainer.py:337,340);
  * ``forward(inputs[B,3,H,W] fp32 NCHW) -> (final_1, final_2, final_3)``, each ``[B,4,H,W]`` fp32
    in (0,1) — models/unet.py:255-300;
  * the parameter tree, hence the 98-key / 2 216 944-byte ``state_dict`` — models/unet.py:121-254.
    The holder modules below are ordinary ``torch.nn`` containers created in the reference's order
    and re-initialised by the same three passes, so ``torch.manual_seed(s); UNet_Nested()`` draws
    the same random stream as the reference does and ``load_state_dict`` of a reference checkpoint
    round-trips (tests/test_dropin_api.py pins both against golden hashes).

Only the parameter *holders* are torch modules; none of their ``forward`` methods is used on the
product path.  ``UNet_Nested.forward`` hands the whole network to ``engine.Engine``, which calls
libunpp.so through its C ABI.  There is no CPU or cuDNN fallback: a CPU tensor, a missing
``libunpp.so`` or an unsupported option raises.
"""
from __future__ import annotations

import threading

import torch
import torch.nn as nn
from torch.nn import init


def weights_init_kaiming(m: nn.Module) -> None:
    """Class-name driven initialiser (reference models/unet.py:165-174): convs/linears get
    kaiming-normal (fan_in, a=0) weights, batch norms get gamma ~ N(1, 0.02), beta = 0."""
    kind = type(m).__name__
    if "Conv" in kind or "Linear" in kind:
        init.kaiming_normal_(m.weight.data, a=0, mode="fan_in")
    elif "BatchNorm" in kind:
        init.normal_(m.weight.data, 1.0, 0.02)
        init.constant_(m.bias.data, 0.0)


def init_weights(net: nn.Module, init_type: str = "normal") -> None:
    """reference models/unet.py:158-163 — only 'kaiming' exists; anything else raises."""
    if init_type != "kaiming":
        raise NotImplementedError("initialization method [%s] is not implemented" % init_type)
    net.apply(weights_init_kaiming)


def count_param(model: nn.Module) -> int:
    """reference models/unet.py:176-180."""
    return sum(p.numel() for p in model.parameters())


class unetConv2(nn.Module):
    """Parameter holder for n x (Conv2d 3x3 s1 p1 [+ BatchNorm2d] + ReLU) — models/unet.py:121-148.
    Attribute names (conv1, conv2, ...) and Sequential indices (0 conv, 1 bn) define state_dict keys."""

    def __init__(self, in_size, out_size, is_batchnorm, n=2, ks=3, stride=1, padding=1):
        super().__init__()
        self.n, self.ks, self.stride, self.padding = n, ks, stride, padding
        width = in_size
        for i in range(1, n + 1):
            layers = [nn.Conv2d(width, out_size, ks, stride, padding)]
            if is_batchnorm:
                layers.append(nn.BatchNorm2d(out_size))
            layers.append(nn.ReLU(inplace=True))
            setattr(self, "conv%d" % i, nn.Sequential(*layers))
            width = out_size
        for child in self.children():
            init_weights(child, init_type="kaiming")

    def forward(self, inputs):  # pragma: no cover - holders are never run
        raise RuntimeError("unetConv2 is a parameter holder; run the enclosing UNet_Nested")


class unetUp(nn.Module):
    """Parameter holder for ConvTranspose2d(k2,s2) + concat + unetConv2(no BN) — models/unet.py:182-196.
    ``conv`` is registered before ``up`` (state_dict order)."""

    def __init__(self, in_size, out_size, is_deconv, n_concat=2):
        super().__init__()
        self.conv = unetConv2(in_size + (n_concat - 2) * out_size, out_size, False)
        if is_deconv:
            self.up = nn.ConvTranspose2d(in_size, out_size, kernel_size=2, stride=2, padding=0)
        else:
            self.up = nn.Sequential(nn.UpsamplingBilinear2d(scale_factor=2), nn.Conv2d(in_size, out_size, 1))
        for child in self.children():
            if "unetConv2" in type(child).__name__:
                continue
            init_weights(child, init_type="kaiming")

    def forward(self, high_feature, *low_feature):  # pragma: no cover
        raise RuntimeError("unetUp is a parameter holder; run the enclosing UNet_Nested")


class UNet_Nested(nn.Module):
    def __init__(self, in_channels=3, n_classes=4, feature_scale=2, is_deconv=True, is_batchnorm=True, is_ds=True):
        super().__init__()
        self.in_channels = in_channels
        self.n_classes = n_classes
        self.feature_scale = feature_scale
        self.is_deconv = is_deconv
        self.is_batchnorm = is_batchnorm
        self.is_ds = is_ds  # stored, unused — as in the reference (models/unet.py:212)

        f = [int(c / feature_scale) for c in (32, 64, 128, 256, 512)]
        self.maxpool = nn.MaxPool2d(kernel_size=2)
        self.conv00 = unetConv2(in_channels, f[0], is_batchnorm)
        self.conv10 = unetConv2(f[0], f[1], is_batchnorm)
        self.conv20 = unetConv2(f[1], f[2], is_batchnorm)
        self.conv30 = unetConv2(f[2], f[3], is_batchnorm)
        self.up_concat01 = unetUp(f[1], f[0], is_deconv)
        self.up_concat11 = unetUp(f[2], f[1], is_deconv)
        self.up_concat21 = unetUp(f[3], f[2], is_deconv)
        self.up_concat02 = unetUp(f[1], f[0], is_deconv, 3)
        self.up_concat12 = unetUp(f[2], f[1], is_deconv, 3)
        self.up_concat03 = unetUp(f[1], f[0], is_deconv, 4)
        self.final_1 = nn.Conv2d(f[0], n_classes, 1)
        self.final_2 = nn.Conv2d(f[0], n_classes, 1)
        self.final_3 = nn.Conv2d(f[0], n_classes, 1)
        # third init pass of the reference (models/unet.py:248-252): every nn.Conv2d and
        # nn.BatchNorm2d again; ConvTranspose2d is not an nn.Conv2d instance and is left alone.
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.BatchNorm2d)):
                init_weights(m, init_type="kaiming")
        self.drop_out = nn.Dropout(p=0.4)

        # "bf16" = the tensor-core path (bf16 storage, fp32 accumulation); "fp32" = the validation mode of the eval forward
        # (csrc/ref_kernels.cu: fp32 storage and arithmetic on the CUDA cores, <= 1e-3 of the reference by BASELINE's tolerance)
        self.precision = "bf16"
        self._engines = {}
        self._engines_lock = threading.Lock()

    # ------------------------------------------------------------------------------ engine
    def _engine(self, device: torch.device):
        from ..engine import Engine
        key = (device.type, device.index)
        with self._engines_lock:
            eng = self._engines.get(key)
            if eng is None:
                eng = self._engines[key] = Engine(self, device)
            # nn.DataParallel re-creates its replicas (sharing this dict) at every forward: the cached engine of a device
            # must read THIS module's parameters, not those of the replica it was built from
            eng.model = self
        return eng

    def __getstate__(self):  # engines hold device buffers and a lock: never pickled / deep-copied
        state = self.__dict__.copy()
        state["_engines"] = {}
        state["_engines_lock"] = None
        state.pop("_unpp_named_params", None)  # (engine.named_params' cache: rebuilt on demand)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._engines = {}
        self._engines_lock = threading.Lock()

    def forward(self, inputs):
        if not isinstance(inputs, torch.Tensor) or inputs.dim() != 4:
            raise ValueError("UNet_Nested expects a [B, C, H, W] tensor")
        if not inputs.is_cuda:
            raise RuntimeError("UNet_Nested (B200-native) has no CPU path: move the module and its input to a CUDA device")
        return self._engine(inputs.device).forward(inputs)

    @torch.no_grad()
    def predict_keypoints(self, inputs, head: int = 2):
        """Inference + fused-epilogue heat maps + warp-reduction arg-max: returns
        ``(xy int32 [B, n_classes, 2] as [x, y], peak fp32 [B, n_classes], heatmaps tuple)``
        for head index ``head`` (0..2; default the deepest, final_3).  The arg-max follows
        tools/misc/heatmap.py:173-178 (first maximum in row-major order, x first)."""
        return self._engine(inputs.device).predict_keypoints(inputs, head)

    @staticmethod
    @torch.no_grad()
    def extract_points(heatmaps, num: int, threshold: float = 0.5):
        """The multi-point form of ``Heatmap.extract_points_(pred, num)`` (tools/misc/heatmap.py:148-208), for all planes of a heat-map
        tensor ``[B, C, H, W]`` at once: the ``num`` brightest regions above ``threshold`` (one retry at 0.9 x threshold), brightest
        first, as ``(xy int32 [B, C, num, 2] as [x, y], -1 where a plane has fewer; peak fp32 [B, C, num]; count int32 [B, C])``.
        Regions are strict local maxima instead of the OpenCV watershed: the same points on separated blobs (tests/golden/unetpp_r2.*)."""
        from .. import ops
        if not heatmaps.is_cuda:
            raise RuntimeError("extract_points runs on the device: pass the CUDA heat maps the model returned")
        with torch.cuda.device(heatmaps.device):
            return ops.topk_peaks(heatmaps.float(), num, threshold)
