"""GPU: SURVEY.md §8f rows N3 / N4 — the non-default constructor flags of ``UNet_Nested`` (``is_deconv=False``: bilinear x2 +
1x1 conv, models/unet.py:189-191; ``is_batchnorm=False``, unet.py:137-143) and every optimizer the reference trainer can
select (trainer/trainer.py:344-376), against the reference outputs in tests/golden/unetpp_variants.* and the oracle.

Bounds: the bf16 bounds of tests/test_forward_gpu.py / tests/test_training_gpu.py (heat maps max |err| <= 3e-2, mean <= 3e-3; loss
1e-2 relative; parameter gradients per group).  Without BatchNorm the conv biases of the encoder have real gradients and are
checked like every other encoder parameter.  Optimizer steps are fp32 element-wise arithmetic: rtol 2e-6."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import unet_nested4tiny_objects_keypoints_b200 as pkg  # noqa: E402
from oracle import unetpp_oracle as O  # noqa: E402
from unet_nested4tiny_objects_keypoints_b200 import fused, ops, optimizers  # noqa: E402
from test_training_gpu import emulation_bounds  # noqa: E402

DEV = "cuda"
VARIANTS = {"bilinear": dict(is_deconv=False, is_batchnorm=True), "nobn": dict(is_deconv=True, is_batchnorm=False),
            "bilinear_nobn": dict(is_deconv=False, is_batchnorm=False)}


def bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(DEV)


def nchw(t):
    return t.float().cpu().permute(0, 3, 1, 2).contiguous()


def close(got, ref, rel, what=""):
    ref = ref.double()
    err = float((got.double() - ref).abs().max())
    scale = float(ref.abs().max()) + 1e-30
    assert err <= rel * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e} (rel {err / scale:.3e} > {rel})"


# ---------------------------------------------------------------------------------------------- kernels
@pytest.mark.parametrize("N,H,W,C", [(2, 8, 8, 16), (1, 16, 24, 32), (3, 4, 4, 64), (1, 1, 1, 16), (1, 1, 5, 32), (2, 32, 32, 16), (1, 64, 48, 64)])
def test_bilinear_up2x_forward_and_adjoint(N, H, W, C):
    g = torch.Generator().manual_seed(H * 100 + W)
    x = bf(torch.randn(N, C, H, W, generator=g))
    ref = F.interpolate(x.double(), scale_factor=2, mode="bilinear", align_corners=True)  # nn.UpsamplingBilinear2d(scale_factor=2)
    y = torch.empty(N, 2 * H, 2 * W, C, dtype=torch.bfloat16, device=DEV)
    ops.bilinear_up2x(nhwc(x), y)
    close(nchw(y), ref, 6e-3, "bilinear_up2x")
    dy = bf(torch.randn(N, C, 2 * H, 2 * W, generator=g))
    xd = x.double().requires_grad_(True)
    F.interpolate(xd, scale_factor=2, mode="bilinear", align_corners=True).backward(dy.double())
    dx = torch.empty(N, H, W, C, dtype=torch.bfloat16, device=DEV)
    ops.bilinear_up2x_bwd(nhwc(dy), dx)
    close(nchw(dx), xd.grad, 6e-3, "bilinear_up2x_bwd")
    # determinism
    dx2 = torch.empty_like(dx)
    ops.bilinear_up2x_bwd(nhwc(dy), dx2)
    assert torch.equal(dx, dx2)


@pytest.mark.parametrize("N,H,W,cin,cout", [(2, 16, 16, 32, 16), (1, 8, 24, 64, 32), (2, 4, 4, 128, 64), (1, 32, 32, 16, 32)])
def test_conv1x1_on_tensor_cores(N, H, W, cin, cout):
    """The pointwise GEMM the bilinear path runs on the low-resolution grid: forward (pack kind 0, taps 1) and dgrad (kind 1)."""
    g = torch.Generator().manual_seed(cin + cout)
    x = bf(torch.randn(N, cin, H, W, generator=g))
    w = bf(torch.randn(cout, cin, 1, 1, generator=g) * (2.0 / cin) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    nt = ops.pick_n_tile(cout, cin, 1)
    out = torch.empty(N, H, W, cout, dtype=torch.bfloat16, device=DEV)
    ops.conv([nhwc(x)], N, H, W, ops.pack_weights(w.to(DEV), 0, 1, cout, nt, cin), cout, nt, 1, bias=b.to(DEV), out=out)
    close(nchw(out), F.conv2d(x.double(), w.double(), b.double()), 6e-3, "conv1x1 fwd")
    dz = bf(torch.randn(N, cout, H, W, generator=g))
    ntd = ops.pick_n_tile(cin, cout, 1)
    dx = torch.empty(N, H, W, cin, dtype=torch.bfloat16, device=DEV)
    ops.conv([nhwc(dz)], N, H, W, ops.pack_weights(w.to(DEV), 1, 1, cin, ntd, cout), cin, ntd, 1, out=dx)
    close(nchw(dx), F.conv_transpose2d(dz.double(), w.double()), 6e-3, "conv1x1 dgrad")
    # weight gradient of the 1x1 conv
    grid = ops.wgrad_grid([cin], N, H, W, cout, 1)
    part = torch.empty(grid * cin * cout, dtype=torch.float32, device=DEV)
    ops.wgrad([nhwc(x)], N, H, W, nhwc(dz), cout, 1, part)
    dw = torch.zeros(cout * cin, dtype=torch.float32, device=DEV)
    ops.wgrad_reduce(part, grid, 1, cin, cout, dw, 0, cin, cin, 1, 0)
    ref = torch.einsum("nihw,nohw->oi", x.double(), dz.double())
    close(dw.cpu().view(cout, cin), ref, 2e-3, "conv1x1 wgrad")


# ---------------------------------------------------------------------------------------------- optimizers
def _opt_kwargs(tag, hyper):
    if tag.startswith("sgdw") or tag == "sgd":
        return dict(lr=hyper["lr"], beta1=hyper.get("momentum", 0.0), weight_decay=hyper["weight_decay"])
    return dict(lr=hyper["lr"], beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=hyper["weight_decay"])


@pytest.mark.parametrize("device_counter", [False, True])
def test_optim_step_matches_reference(variants_golden, device_counter):
    arr, meta = variants_golden
    for tag, hyper in meta["optimizers"].items():
        kind = "sgdw" if tag.startswith("sgdw") else tag
        for j in range(2):
            p = torch.from_numpy(arr[f"opt_p0_{j}"]).reshape(-1).to(DEV)
            s1, s2 = torch.zeros_like(p), torch.zeros_like(p)
            counter = torch.zeros(1, dtype=torch.int64, device=DEV)
            scal = torch.zeros(4, dtype=torch.float32, device=DEV)
            for s in range(3):
                g = torch.from_numpy(arr[f"opt_g{s}_{j}"]).reshape(-1).to(DEV)
                if device_counter:
                    ops.optim_step(kind, p, g, s1, s2, step_counter=counter, scalars=scal, **_opt_kwargs(tag, hyper))
                else:
                    ops.optim_step(kind, p, g, s1, s2, step=s + 1, **_opt_kwargs(tag, hyper))
            assert np.allclose(p.cpu().numpy(), arr[f"opt_{tag}_p3_{j}"].reshape(-1), rtol=2e-6, atol=1e-7), (tag, j)
            if f"opt_{tag}_buf3_{j}" in arr:
                assert np.allclose(s1.cpu().numpy(), arr[f"opt_{tag}_buf3_{j}"].reshape(-1), rtol=2e-6, atol=1e-7), (tag, j)
            if device_counter:
                assert int(counter) == 3


def test_adamw_through_optim_step_equals_unpp_adamw(golden):
    arr, meta = golden
    h = meta["adamw_hyper"]
    p = torch.from_numpy(arr["adamw_p0_0"]).reshape(-1).to(DEV)
    q = p.clone()
    m1, v1, m2, v2 = (torch.zeros_like(p) for _ in range(4))
    for s in range(3):
        g = torch.from_numpy(arr[f"adamw_g{s}_0"]).reshape(-1).to(DEV)
        ops.adamw(p, g, m1, v1, h["lr"], h["betas"][0], h["betas"][1], h["eps"], h["weight_decay"], s + 1)
        ops.optim_step("adamw", q, g, m2, v2, lr=h["lr"], beta1=h["betas"][0], beta2=h["betas"][1], eps=h["eps"], weight_decay=h["weight_decay"], step=s + 1)
    for a, b in ((p, q), (m1, m2), (v1, v2)):  # same arithmetic, the compiler may contract the multiply-adds differently
        assert torch.allclose(a, b, rtol=2e-6, atol=1e-7)
    assert np.allclose(q.cpu().numpy(), arr["adamw_p3_0"].reshape(-1), rtol=2e-6, atol=1e-7)


def test_sgdw_and_adabound_drop_in_classes(variants_golden):
    """Same constructors as tools/optimizers/sgdw.py / adabound.py, driven like trainer.py:115-136 (p.grad set, step())."""
    arr, meta = variants_golden
    for tag, cls in (("sgdw", optimizers.SGDW), ("sgdw_momentum", optimizers.SGDW), ("adabound", optimizers.AdaBound)):
        params = [torch.nn.Parameter(torch.from_numpy(arr[f"opt_p0_{j}"]).to(DEV)) for j in range(2)]
        opt = cls(params, **meta["optimizers"][tag])
        for s in range(3):
            for j, p in enumerate(params):
                p.grad = torch.from_numpy(arr[f"opt_g{s}_{j}"]).to(DEV)
            opt.step()
        for j, p in enumerate(params):
            assert np.allclose(p.detach().cpu().numpy(), arr[f"opt_{tag}_p3_{j}"], rtol=2e-6, atol=1e-7), (tag, j)
    # an lr scheduler acts on AdaBound's bounds through lr / base_lr (adabound.py:117-121)
    p = torch.nn.Parameter(torch.ones(64, device=DEV))
    opt = optimizers.AdaBound([p], lr=1e-2)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, 0.5)  # trainer.py:386-388
    ref_p, m, v = torch.ones(64), torch.zeros(64), torch.zeros(64)
    for s in range(1, 4):
        p.grad = torch.full((64,), 0.25 * s, device=DEV)
        opt.step()
        ref_p, m, v = O.adabound_reference_step(ref_p, torch.full((64,), 0.25 * s), m, v, s, lr=1e-2 * 0.5 ** (s - 1), base_lr=1e-2)
        sched.step()
    assert np.allclose(p.detach().cpu().numpy(), ref_p.numpy(), rtol=2e-6, atol=1e-7)


# ---------------------------------------------------------------------------------------------- constructor flags
def bf16_noise_floor(sd, x):
    """Error of a torch-CPU emulation of bf16 storage (weights, input, every activation rounded to bf16, fp32 arithmetic) against the
    fp32 oracle: (max, mean) over the three heat maps.  The synthetic ``up.1.weight`` of the bilinear variants has twice the gain of
    the transposed-conv weights (fan-in Cin instead of 4*Cin), so their decoder activations — and the bf16 noise on the logits — are
    larger than the default model's; the kernels are held to 1.5x this floor where it exceeds the default bound."""
    ref = O.forward(sd, x)
    sdb = {k: (bf(v) if (v.dtype.is_floating_point and v.dim() == 4) else v) for k, v in sd.items()}
    relu, interp = O.F.relu, O.F.interpolate
    try:
        O.F.relu = lambda t: bf(relu(t))
        O.F.interpolate = lambda *a, **k: bf(interp(*a, **k))
        out = O.forward(sdb, bf(x))
    finally:
        O.F.relu, O.F.interpolate = relu, interp
    return max(float((a - b).abs().max()) for a, b in zip(ref, out)), max(float((a - b).abs().mean()) for a, b in zip(ref, out))


def make(tag, seed=31, train=False):
    kw = VARIANTS[tag]
    m = pkg.UNet_Nested(**kw)
    sd = O.synth_state_dict(seed=seed, **kw)
    m.load_state_dict(sd)
    m = m.to(DEV)
    return (m.train() if train else m.eval()), sd


def check_grads(model, ref, has_bn, bounds):
    """``bounds``: test_training_gpu.emulation_bounds — 3x the deviation of the exact bf16-storage emulation of this variant."""
    for k, p in model.named_parameters():
        r, g = ref[k].double(), p.grad.detach().cpu().double()
        assert g.shape == r.shape, k
        if has_bn and k.startswith("conv") and k.endswith(".0.bias"):
            assert float(g.abs().max()) <= 1e-6 + 10 * float(r.abs().max()), k  # bias in front of BatchNorm: exactly zero here
            continue
        rel = float((g - r).abs().max()) / (float(r.abs().max()) + 1e-30)
        cos = float((g * r).sum() / (g.norm() * r.norm() + 1e-300))
        max_rel, min_cos = bounds(k)
        assert rel <= max_rel and cos >= min_cos, f"{k}: err/max {rel:.3e} (bound {max_rel:.3e}), cosine {cos:.5f} (bound {min_cos:.5f})"


@pytest.mark.parametrize("tag", list(VARIANTS))
def test_state_dict_layout_of_variants(variants_golden, tag):
    _, meta = variants_golden
    m = pkg.UNet_Nested(**VARIANTS[tag])
    assert [[k, list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()] == meta[tag]["state_dict_keys"]


@pytest.mark.parametrize("tag", list(VARIANTS))
def test_variant_eval_forward_matches_reference_golden(variants_golden, tag):
    arr, _ = variants_golden
    m, _ = make(tag)
    with torch.no_grad():
        outs = m(torch.from_numpy(arr["eval_x"]).to(DEV))
    for i, o in enumerate(outs):
        err = np.abs(o.cpu().numpy() - arr[f"{tag}.eval_out{i}"])
        assert err.max() <= 3e-2 and err.mean() <= 3e-3, (tag, i, err.max(), err.mean())


@pytest.mark.parametrize("tag", list(VARIANTS))
@pytest.mark.parametrize("B,H,W", [(1, 64, 96), (2, 8, 8), (2, 128, 128)])
def test_variant_eval_forward_matches_oracle(tag, B, H, W):
    m, sd = make(tag, seed=33)
    x = torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(B + H))
    ref = O.forward(sd, x)
    with torch.no_grad():
        outs = m(x.to(DEV))
    floor_max, floor_mean = bf16_noise_floor(sd, x)
    for r, o in zip(ref, outs):
        err = (o.cpu() - r).abs()
        assert float(err.max()) <= max(3e-2, 1.5 * floor_max) and float(err.mean()) <= max(3e-3, 1.5 * floor_mean), (tag, float(err.max()), floor_max)
    xy, val, heats = m.predict_keypoints(x.to(DEV))
    rxy, _ = O.argmax_keypoints(heats[2].cpu().numpy())
    assert np.array_equal(xy.cpu().numpy(), rxy)  # bit-exact on identical heat maps


@pytest.mark.parametrize("tag", list(VARIANTS))
def test_variant_train_step_matches_reference_golden(variants_golden, tag):
    arr, meta = variants_golden
    m, sd = make(tag, train=True)
    x, target = torch.from_numpy(arr["train_x"]), torch.from_numpy(arr["train_target"])
    masks = [torch.from_numpy(np.unpackbits(arr[f"train_mask{i}"]).reshape(3, 16, 32, 32)) for i in range(3)]
    m._forced_dropout_masks = masks
    outs = m(x.to(DEV))
    loss = sum(F.mse_loss(o, target.to(DEV)) for o in outs) / len(outs)  # trainer.py:125-134 with nn.MSELoss (427)
    loss.backward()
    torch.cuda.synchronize()
    err = np.abs(outs[2].detach().cpu().numpy() - arr[f"{tag}.train_out2"])
    assert err.max() <= 3e-2 and err.mean() <= 3e-3
    assert abs(float(loss) - meta[tag]["train_loss"]) <= 1e-2 * meta[tag]["train_loss"]
    named = dict(m.named_parameters())
    _, _, rg, _ = O.train_step_grads(sd, x, target, dropout_masks=masks)
    bounds = emulation_bounds(sd, x, target, masks, 0.4, rg)
    for k in ("up_concat01.up.1.weight", "up_concat21.up.1.weight", "conv10.conv1.0.bias", "conv00.conv2.0.weight"):
        if f"{tag}.train_grad_{k}" in arr and not (VARIANTS[tag]["is_batchnorm"] and k.endswith(".0.bias")):
            r = arr[f"{tag}.train_grad_{k}"]
            assert np.abs(named[k].grad.cpu().numpy() - r).max() <= bounds(k)[0] * np.abs(r).max(), (tag, k)
    check_grads(m, rg, VARIANTS[tag]["is_batchnorm"], bounds)


@pytest.mark.parametrize("tag", list(VARIANTS))
@pytest.mark.parametrize("B,H,W", [(2, 64, 64), (1, 32, 48)])
def test_variant_train_step_matches_oracle(tag, B, H, W):
    m, sd = make(tag, seed=35, train=True)
    m.drop_out.p = 0.0
    g = torch.Generator().manual_seed(H + W)
    x, target = torch.randn(B, 3, H, W, generator=g), torch.rand(B, 4, H, W, generator=g)
    outs = m(x.to(DEV))
    loss = sum(F.mse_loss(o, target.to(DEV)) for o in outs) / len(outs)
    loss.backward()
    rl, routs, rg, _ = O.train_step_grads(sd, x, target, dropout_masks=None)
    for o, r in zip(outs, routs):
        assert float((o.detach().cpu() - r).abs().max()) <= 3e-2
    assert abs(float(loss) - float(rl)) <= 1e-2 * float(rl)
    check_grads(m, rg, VARIANTS[tag]["is_batchnorm"], emulation_bounds(sd, x, target, None, 0.0, rg))


@pytest.mark.parametrize("opt", ["sgd", "adam", "adabound", "sgdw", "adamw"])
def test_fused_train_step_with_every_trainer_optimizer(opt):
    """FusedTrainStep(optimizer=...) in a CUDA graph: two steps equal the oracle's optimizer applied to the step's own gradients;
    set_lr (what MultiStepLR / ExponentialLR do, trainer.py:383-388) takes effect without re-capturing."""
    B, H, W = 2, 32, 32
    m = pkg.UNet_Nested()
    m.load_state_dict(O.synth_state_dict(seed=41))
    m = m.to(DEV).train()
    m.drop_out.p = 0.0
    lr, wd = 1e-3, 1e-2
    step = fused.FusedTrainStep(m, B, H, W, lr=lr, weight_decay=wd, optimizer=opt, momentum=0.9)
    g = torch.Generator().manual_seed(5)
    p, s1, s2 = step.flat_p.cpu().clone(), None, None
    if opt in ("adam", "adabound", "adamw"):
        s1, s2 = torch.zeros_like(p), torch.zeros_like(p)
    for t in (1, 2):
        if t == 2:
            lr = 5e-4
            step.set_lr(lr)
        x, target = torch.randn(B, 3, H, W, generator=g), torch.rand(B, 4, H, W, generator=g)
        step.step(x, target)
        torch.cuda.synchronize()
        grad = step.flat_g.cpu()
        if opt == "sgd":
            p, s1 = O.sgd_reference_step(p, grad, s1, lr, momentum=0.9, weight_decay=wd)
        elif opt == "sgdw":
            p, s1 = O.sgdw_reference_step(p, grad, s1, lr, momentum=0.9, weight_decay=wd)
        elif opt == "adam":
            p, s1, s2 = O.adam_reference_step(p, grad, s1, s2, t, lr=lr, weight_decay=wd)
        elif opt == "adabound":
            p, s1, s2 = O.adabound_reference_step(p, grad, s1, s2, t, lr=lr, weight_decay=wd, base_lr=1e-3)
        else:
            p, s1, s2 = O.adamw_reference_step(p, grad, s1, s2, t, lr=lr, weight_decay=wd)
        assert np.allclose(step.flat_p.cpu().numpy(), p.numpy(), rtol=1e-5, atol=1e-7), (opt, t)


@pytest.mark.parametrize("tag", ["bilinear", "nobn"])
def test_fused_train_step_on_variants(tag):
    """The captured fused step (fwd + bwd + MSE + AdamW) on the constructor-flag variants: loss and gradients vs the oracle."""
    B, H, W = 2, 32, 32
    m, sd = make(tag, seed=43, train=True)
    m.drop_out.p = 0.0
    step = fused.FusedTrainStep(m, B, H, W, loss="mse")
    g = torch.Generator().manual_seed(9)
    x, target = torch.randn(B, 3, H, W, generator=g), torch.rand(B, 4, H, W, generator=g)
    loss = step.step(x, target)
    torch.cuda.synchronize()
    rl, _, rg, _ = O.train_step_grads(sd, x, target, dropout_masks=None)
    assert abs(float(loss) - float(rl)) <= 1e-2 * float(rl)
    bounds = emulation_bounds(sd, x, target, None, 0.0, rg)
    for k, (off, n) in step.ts.lay.items():
        r = rg[k].double().reshape(-1)
        got = step.flat_g[off:off + n].cpu().double()
        if VARIANTS[tag]["is_batchnorm"] and k.startswith("conv") and k.endswith(".0.bias"):
            continue
        rel = float((got - r).abs().max()) / (float(r.abs().max()) + 1e-30)
        assert rel <= bounds(k)[0], (tag, k, rel)


# ---------------------------------------------------------------------------------------------- first-layer mode (4-channel input pixels)
@pytest.mark.parametrize("N,H,W,cin", [(2, 32, 32, 3), (1, 8, 8, 3), (1, 24, 40, 3), (3, 64, 96, 1), (1, 16, 136, 4), (2, 256, 256, 3), (1, 72, 8, 2)])
def test_first_layer_conv_on_4_channel_pixels(N, H, W, cin):
    """conv00.conv1 (models/unet.py:132) in the first-layer mode: fp32 NCHW -> NHWC4 bf16 (8 B/pixel), 3x3 conv + folded-BN scale + bias +
    ReLU through overlapping unswizzled K-major descriptors (block2x2 = 2, pack kind 7) vs torch fp64 on the same bf16-rounded operands."""
    g = torch.Generator().manual_seed(N * 1000 + H + W + cin)
    x = torch.randn(N, cin, H, W, generator=g)
    w = bf(torch.randn(16, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5)
    b = torch.randn(16, generator=g) * 0.1
    x4 = torch.empty(N, H, W, 4, dtype=torch.bfloat16, device=DEV)
    ops.nchw_to_nhwc4(x.to(DEV), x4)
    ref4 = torch.zeros(N, 4, H, W)
    ref4[:, :cin] = bf(x)
    assert torch.equal(nchw(x4), ref4)
    out = torch.empty(N, H, W, 16, dtype=torch.bfloat16, device=DEV)
    ops.conv([x4], N, H, W, ops.pack_weights_c4(w.to(DEV)), 16, ops.NTile(16, b2=2), 9, bias=b.to(DEV), relu=True, out=out)
    ref = F.relu(F.conv2d(bf(x).double(), w.double(), b.double(), padding=1))
    close(nchw(out), ref, 6e-3, "first-layer conv")
    # folded BatchNorm scale (eval mode) and no ReLU
    scale = torch.rand(16, generator=g) + 0.5
    ops.conv([x4], N, H, W, ops.pack_weights_c4(w.to(DEV), scale=scale.to(DEV)), 16, ops.NTile(16, b2=2), 9, bias=b.to(DEV), out=out)
    ref = F.conv2d(bf(x).double(), bf(w * scale.view(16, 1, 1, 1)).double(), b.double(), padding=1)
    close(nchw(out), ref, 6e-3, "first-layer conv, folded scale")


def test_first_layer_mode_equals_the_16_channel_path():
    """Same bf16 operands, same fp32 accumulation, different MMA shapes: the two inference paths agree to output rounding."""
    m = pkg.UNet_Nested()
    m.load_state_dict(O.synth_state_dict(seed=3))
    m = m.to(DEV).eval()
    x = torch.randn(2, 3, 64, 96, generator=torch.Generator().manual_seed(1)).to(DEV)
    with torch.no_grad():
        a = [h.clone() for h in m(x)]
        eng = m._engine(x.device)
        eng.first_layer_c4, eng._packed_key = False, None
        b = m(x)
    for u, v in zip(a, b):
        assert float((u - v).abs().max()) <= 2e-2


# ---------------------------------------------------------------------------------------------- other constructor values (models/unet.py:206,242-244)
@pytest.mark.parametrize("in_channels,n_classes", [(1, 1), (4, 8), (5, 2), (3, 3)])
def test_other_in_channels_and_n_classes_inference_and_training(in_channels, n_classes):
    """``UNet_Nested(in_channels=..., n_classes=...)``: the first conv reads 1..16 input channels (<= 4: the 8-byte-pixel first-layer mode,
    above: the 16-channel padded tensor), the heads produce 1..8 classes.  Inference against the oracle, the training step through the
    teacher-forced check (every stored tensor to one bf16 ulp, every gradient to 1e-3) and the key points bit-exact."""
    from oracle import teacher_forced as T
    kw = dict(in_channels=in_channels, n_classes=n_classes)
    sd = O.synth_state_dict(seed=91, **kw)
    m = pkg.UNet_Nested(**kw)
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, tuple(v.shape)) for k, v in sd.items()]
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    g = torch.Generator().manual_seed(in_channels * 10 + n_classes)
    B, H, W = 2, 32, 48
    x = torch.randn(B, in_channels, H, W, generator=g)
    with torch.no_grad():
        outs = m(x.to(DEV))
    ref = O.forward(sd, x)
    for o, r in zip(outs, ref):
        assert tuple(o.shape) == (B, n_classes, H, W)
        assert float((o.cpu() - r).abs().max()) <= 3e-2
    xy, val, heats = m.predict_keypoints(x.to(DEV))
    rxy, _ = O.argmax_keypoints(heats[2].cpu().numpy())
    assert np.array_equal(xy.cpu().numpy(), rxy)
    m.train()
    m.drop_out.p = 0.0
    target = torch.rand(B, n_classes, H, W, generator=g)
    outs = m(x.to(DEV))
    (sum(F.mse_loss(o, target.to(DEV)) for o in outs) / 3).backward()
    torch.cuda.synchronize()
    ts = m._engine(torch.device("cuda", torch.cuda.current_device()))._train_states[(B, H, W)]
    rep = T.verify_step(ts.t, ts.heats, {k: p.grad for k, p in m.named_parameters()}, sd, x.to(DEV), target=target)
    assert not rep.check(), rep.check()


def test_feature_scale_other_than_two_is_rejected_with_a_reason():
    """feature_scale changes every channel count of the network (8- or 256-channel levels, a 32-channel head input): outside the kernels'
    16..128-channel tiling.  The reference trainer never passes it (trainer.py:337,340 build the model with no arguments)."""
    m = pkg.UNet_Nested(feature_scale=4).to(DEV).eval()
    with pytest.raises(ValueError, match="feature_scale"):
        m(torch.randn(1, 3, 32, 32, device=DEV))
