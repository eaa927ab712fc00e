"""Pins the oracle's constructor-flag variants and extra optimizers against the REAL reference and writes
tests/golden/unetpp_variants.{npz,json}.  TEST INFRASTRUCTURE ONLY — run in the build container
(reference mounted read-only at /root/reference):

    python oracle/make_golden_variants.py

Covers SURVEY.md §8f rows N3 and N4:
  * ``UNet_Nested(is_deconv=False)`` — UpsamplingBilinear2d(2) + Conv2d 1x1 instead of ConvTranspose2d (models/unet.py:189-191),
    ``UNet_Nested(is_batchnorm=False)`` (models/unet.py:137-143) and both: state_dict layout, eval forward, one training
    step (masked dropout, MSE mean over the three heads like trainer/trainer.py:125-135) with every parameter gradient;
  * tools/optimizers/sgdw.py SGDW, tools/optimizers/adabound.py AdaBound, torch.optim.SGD and torch.optim.Adam with the
    arguments trainer/trainer.py:344-376 passes: three steps on two tensors.
The REFERENCE outputs are stored; tests/test_oracle_golden.py replays the oracle against them anywhere.
"""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("UNPP_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from oracle import unetpp_oracle as O  # noqa: E402
from oracle.make_golden import GOLD, _MaskDropout, digest  # noqa: E402

VARIANTS = {"bilinear": dict(is_deconv=False, is_batchnorm=True), "nobn": dict(is_deconv=True, is_batchnorm=False),
            "bilinear_nobn": dict(is_deconv=False, is_batchnorm=False)}


def main():
    torch.set_num_threads(4)
    from models.unet import UNet_Nested as RefNet
    meta, arrays = {}, {}
    g = torch.Generator().manual_seed(21)
    x = torch.randn(2, 3, 32, 32, generator=g)
    xt = torch.randn(3, 3, 32, 32, generator=g)
    target = torch.rand(3, 4, 32, 32, generator=g)
    masks = [(torch.rand(3, 16, 32, 32, generator=g) >= 0.4).to(torch.uint8) for _ in range(3)]
    arrays["eval_x"], arrays["train_x"], arrays["train_target"] = x.numpy(), xt.numpy(), target.numpy()
    for i, m in enumerate(masks):
        arrays[f"train_mask{i}"] = np.packbits(m.numpy().reshape(-1))
    for tag, kw in VARIANTS.items():
        torch.manual_seed(0)
        ref = RefNet(**kw)
        sd = ref.state_dict()
        spec = O.state_dict_spec(**kw)
        assert list(sd.keys()) == list(spec.keys()), (tag, "state_dict key order differs")
        for k, v in sd.items():
            assert tuple(v.shape) == spec[k][0] and v.dtype == spec[k][1], (tag, k)
        vm = meta[tag] = {"state_dict_keys": [[k, list(v.shape), str(v.dtype)] for k, v in sd.items()]}
        wsd = O.synth_state_dict(seed=31, **kw)
        ref.load_state_dict(wsd)
        ref.eval()
        with torch.no_grad():
            r_out = ref(x)
        o_out = O.forward(wsd, x)
        for i, (a, b) in enumerate(zip(r_out, o_out)):
            assert torch.allclose(a, b, rtol=0, atol=2e-6), (tag, float((a - b).abs().max()))
            arrays[f"{tag}.eval_out{i}"] = a.numpy()
        ref.load_state_dict(wsd)
        ref.train()
        ref.drop_out = _MaskDropout(masks)
        ref.zero_grad()
        outs = ref(xt)
        crit = torch.nn.MSELoss()
        loss = sum(crit(o, target) for o in outs) / len(outs)
        loss.backward()
        o_loss, o_outs, o_grads, o_stats = O.train_step_grads(wsd, xt, target, dropout_masks=masks)
        assert abs(float(loss.detach()) - float(o_loss)) < 1e-7, tag
        ref_grads = {k: p.grad for k, p in ref.named_parameters()}
        assert set(ref_grads) == set(o_grads), tag
        for k, a in ref_grads.items():
            b = o_grads[k]
            assert torch.allclose(a, b, rtol=1e-4, atol=1e-7 + 1e-5 * float(a.abs().max())), (tag, k, float((a - b).abs().max()))
        vm["train_loss"] = float(loss.detach())
        vm["train_grad_digest"] = {k: digest(v) for k, v in ref_grads.items()}
        arrays[f"{tag}.train_out2"] = outs[2].detach().numpy()
        for k in ("up_concat01.up.1.weight", "up_concat21.up.1.weight", "conv10.conv1.0.bias", "conv00.conv2.0.weight"):
            if k in ref_grads:
                arrays[f"{tag}.train_grad_{k}"] = ref_grads[k].numpy()

    # ---- optimizers (trainer/trainer.py:344-376)
    from tools.optimizers.adabound import AdaBound as RefAdaBound
    from tools.optimizers.sgdw import SGDW as RefSGDW
    p0 = [torch.randn(7, 5, generator=g), torch.randn(11, generator=g)]
    grads = [[torch.randn(7, 5, generator=g), torch.randn(11, generator=g)] for _ in range(3)]
    for j in range(2):
        arrays[f"opt_p0_{j}"] = p0[j].numpy()
        for s in range(3):
            arrays[f"opt_g{s}_{j}"] = grads[s][j].numpy()
    cases = {
        "sgdw": (RefSGDW, dict(lr=1e-2, weight_decay=1e-2)),                       # as the trainer builds it (momentum 0)
        "sgdw_momentum": (RefSGDW, dict(lr=1e-2, momentum=0.9, weight_decay=1e-2)),
        "adabound": (RefAdaBound, dict(lr=1e-2, weight_decay=1e-2)),
        "sgd": (torch.optim.SGD, dict(lr=1e-2, momentum=0.9, weight_decay=1e-2)),
        "adam": (torch.optim.Adam, dict(lr=1e-2, weight_decay=1e-2)),
    }
    meta["optimizers"] = {}
    for tag, (cls, hyper) in cases.items():
        params = [torch.nn.Parameter(t.clone()) for t in p0]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            opt = cls(params, **hyper)
            for step_g in grads:
                for p, gr in zip(params, step_g):
                    p.grad = gr.clone()
                opt.step()
        for j in range(2):
            p, a, b = p0[j].clone(), None, None
            if tag in ("adabound", "adam"):
                a, b = torch.zeros_like(p), torch.zeros_like(p)
            for s, step_g in enumerate(grads, 1):
                if tag.startswith("sgdw"):
                    p, a = O.sgdw_reference_step(p, step_g[j], a, **hyper)
                elif tag == "sgd":
                    p, a = O.sgd_reference_step(p, step_g[j], a, **hyper)
                elif tag == "adabound":
                    p, a, b = O.adabound_reference_step(p, step_g[j], a, b, s, **hyper)
                else:
                    p, a, b = O.adam_reference_step(p, step_g[j], a, b, s, **hyper)
            assert torch.allclose(p, params[j].detach(), rtol=1e-6, atol=1e-7), (tag, j, float((p - params[j].detach()).abs().max()))
            arrays[f"opt_{tag}_p3_{j}"] = params[j].detach().numpy()
            st = opt.state[params[j]]
            if "momentum_buffer" in st and st["momentum_buffer"] is not None:
                assert torch.allclose(a, st["momentum_buffer"], rtol=1e-6, atol=1e-7), (tag, j)
                arrays[f"opt_{tag}_buf3_{j}"] = st["momentum_buffer"].numpy()
        meta["optimizers"][tag] = hyper

    np.savez_compressed(os.path.join(GOLD, "unetpp_variants.npz"), **arrays)
    meta["torch_version"] = torch.__version__
    meta["generator"] = "oracle/make_golden_variants.py (reference imported from /root/reference)"
    with open(os.path.join(GOLD, "unetpp_variants.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("variant fixtures written; oracle pinned against the reference: OK")


if __name__ == "__main__":
    main()
