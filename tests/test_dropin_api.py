"""CPU: the drop-in boundary — constructor, state_dict layout, seeded init stream,
error behaviour, and the exported C ABI."""
import ctypes
import hashlib
import os
import re

import pytest
import torch

import unet_nested4tiny_objects_keypoints_b200 as pkg
from unet_nested4tiny_objects_keypoints_b200 import _lib, models

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sha(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode() + v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def test_constructed_with_no_arguments_like_the_trainer():
    m = getattr(models, "UNet_Nested")()  # trainer/trainer.py:337,340
    assert (m.in_channels, m.n_classes, m.feature_scale, m.is_deconv, m.is_batchnorm, m.is_ds) == (3, 4, 2, True, True, True)
    assert models.count_param(m) == 553260


def test_state_dict_layout_and_seeded_init_equal_the_reference(golden):
    _, meta = golden
    torch.manual_seed(0)
    m = pkg.UNet_Nested()
    sd = m.state_dict()
    assert [[k, list(v.shape), str(v.dtype)] for k, v in sd.items()] == meta["state_dict_keys"]
    assert sum(v.numel() * v.element_size() for v in sd.values()) == 2216944
    # same parameter-holder tree + same three init passes => same RNG stream as the reference
    assert _sha(sd) == meta["init_seed0_sha256"]


def test_state_dict_round_trip():
    from oracle import unetpp_oracle as O
    m = pkg.UNet_Nested()
    sd = O.synth_state_dict(seed=5)
    missing, unexpected = m.load_state_dict(sd)
    assert not missing and not unexpected
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k]), k


def test_dataparallel_wrapper_exposes_module_state_dict():
    m = pkg.UNet_Nested()
    dp = torch.nn.DataParallel(m)  # trainer/trainer.py:338; .module.state_dict() at 240
    assert list(dp.module.state_dict().keys()) == list(m.state_dict().keys())


def test_no_cpu_fallback():
    m = pkg.UNet_Nested().eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 3, 32, 32))
    with pytest.raises(ValueError):
        m(torch.zeros(3, 32, 32))


def test_unknown_init_type_raises_like_the_reference():
    with pytest.raises(NotImplementedError):
        models.init_weights(torch.nn.Conv2d(1, 1, 1), init_type="normal")  # unet.py:163


def test_library_exports_every_symbol_declared_in_the_header():
    hdr = open(os.path.join(ROOT, "include", "unpp.h")).read()
    declared = set(re.findall(r"\b(unpp_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found in include/unpp.h"
    assert declared == set(_lib.exported_symbols())
    lib = ctypes.CDLL(_lib.LIB_PATH)  # loading must work without a GPU
    for name in declared:
        assert hasattr(lib, name), name
    loaded = _lib.load()
    assert loaded.unpp_version() >= 1
    assert isinstance(loaded.unpp_last_error(), bytes)


def test_struct_layouts_match_the_header():
    # sizes computed by the C compiler are embedded in the library (unpp_sizeof_*)
    lib = _lib.load()
    assert lib.unpp_sizeof_conv_args() == ctypes.sizeof(_lib.ConvArgs)
    assert lib.unpp_sizeof_pack_args() == ctypes.sizeof(_lib.PackArgs)
    assert lib.unpp_sizeof_wgrad_args() == ctypes.sizeof(_lib.WgradArgs)
    assert lib.unpp_sizeof_reduce_job() == ctypes.sizeof(_lib.ReduceJob)
    assert lib.unpp_sizeof_optim_args() == ctypes.sizeof(_lib.OptimArgs)


@pytest.mark.parametrize("kw", [dict(is_deconv=False), dict(is_batchnorm=False), dict(is_deconv=False, is_batchnorm=False)])
def test_constructor_flag_variants_keep_the_reference_state_dict_layout(variants_golden, kw):
    """is_deconv=False -> up.1.weight [Cout,Cin,1,1] / up.1.bias (unet.py:189-191); is_batchnorm=False -> no BN entries (unet.py:137-143)."""
    _, meta = variants_golden
    tag = {(False, True): "bilinear", (True, False): "nobn", (False, False): "bilinear_nobn"}[(kw.get("is_deconv", True), kw.get("is_batchnorm", True))]
    m = pkg.UNet_Nested(**kw)
    assert [[k, list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()] == meta[tag]["state_dict_keys"]


def test_optimizer_drop_ins_keep_the_reference_constructors():
    from unet_nested4tiny_objects_keypoints_b200 import optimizers
    p = [torch.nn.Parameter(torch.zeros(3))]
    o = optimizers.SGDW(p, lr=0.1, weight_decay=1e-4)            # trainer/trainer.py:364-368
    assert o.defaults == dict(lr=0.1, momentum=0, dampening=0, weight_decay=1e-4, nesterov=False)
    o = optimizers.AdaBound(p, lr=1e-3, weight_decay=1e-4)        # trainer/trainer.py:371-375
    assert o.defaults["final_lr"] == 0.1 and o.defaults["gamma"] == 1e-3 and o.base_lrs == [1e-3]
    with pytest.raises(ValueError):
        optimizers.SGDW(p, lr=0.1, nesterov=True)                 # sgdw.py:63-64
    with pytest.raises(ValueError):
        optimizers.AdaBound(p, gamma=1.5)                         # adabound.py:43-44
    p[0].grad = torch.zeros(3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        o.step()


def test_keep_mask_bit_words_round_trip_on_cpu():
    """The head dropout keep-mask travels as one 16-bit word per pixel (bit c = keep channel c): pack/unpack are exact inverses."""
    import torch
    from unet_nested4tiny_objects_keypoints_b200 import ops
    keep = torch.rand(2, 16, 5, 7, generator=torch.Generator().manual_seed(3)) >= 0.4
    words = ops.pack_keep_mask(keep)
    assert words.dtype == torch.int16 and words.shape == (2, 5, 7)
    assert torch.equal(ops.unpack_keep_mask(words), keep)
    assert int(ops.pack_keep_mask(torch.ones(1, 16, 1, 1, dtype=torch.bool)).item()) == -1  # all sixteen bits set


def test_recording_state_is_per_thread():
    """nn.DataParallel drives one host thread per replica: the deferred-reduction queue of one thread must not leak into another."""
    import threading
    from unet_nested4tiny_objects_keypoints_b200 import ops
    ops.begin_reduce_queue()
    seen = []
    th = threading.Thread(target=lambda: seen.append(ops.reduce_queue_active()))
    th.start()
    th.join()
    assert ops.reduce_queue_active() and seen == [False]
    ops._tls.reduce_queue = None
