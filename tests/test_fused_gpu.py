"""GPU: the CUDA-graph sessions (fused.py) — inference + keypoints, and the fused training step
(forward, MSE fused into the head backward, backward, reference AdamW on the flat buffer)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import unet_nested4tiny_objects_keypoints_b200 as pkg  # noqa: E402
from unet_nested4tiny_objects_keypoints_b200 import fused  # noqa: E402
from oracle import unetpp_oracle as O  # noqa: E402
from test_training_gpu import check_grads, emulation_bounds  # noqa: E402


def test_inference_session_graph_equals_eager_and_oracle():
    sd = O.synth_state_dict(seed=31)
    m = pkg.UNet_Nested()
    m.load_state_dict(sd)
    m = m.cuda().eval()
    x = torch.randn(3, 3, 64, 64, generator=torch.Generator().manual_seed(2))
    sess = fused.InferenceSession(m, 3, 64, 64, head=2)
    assert sess.graph is not None
    xy, val = sess.run(x.pin_memory())
    torch.cuda.synchronize()
    xy2, val2, heats = m.predict_keypoints(x.cuda(), head=2)
    assert torch.equal(xy, xy2) and torch.equal(val, val2) and torch.equal(sess.heat, heats[2])
    ref = O.forward(sd, x)[2]
    assert float((sess.heat.cpu() - ref).abs().max()) <= 3e-2
    rxy, _ = O.argmax_keypoints(sess.heat.cpu().numpy())
    assert np.array_equal(xy.cpu().numpy(), rxy)
    # a second batch through the same graph
    x2 = torch.randn(3, 3, 64, 64, generator=torch.Generator().manual_seed(3))
    sess.run(x2.pin_memory())
    torch.cuda.synchronize()
    assert float((sess.heat.cpu() - O.forward(sd, x2)[2]).abs().max()) <= 3e-2


@pytest.mark.parametrize("use_graph", [False, True])
def test_fused_train_step_loss_grads_and_adamw_update(use_graph):
    sd = O.synth_state_dict(seed=32)
    m = pkg.UNet_Nested()
    m.load_state_dict(sd)
    m = m.cuda().train()
    m.drop_out.p = 0.0
    B, H, W = 2, 32, 32
    hyper = dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    step = fused.FusedTrainStep(m, B, H, W, use_graph=use_graph, loss="mse", **hyper)
    # construction (warm-up + capture) must leave weights, buffers and optimizer state untouched
    for k, v in m.state_dict().items():
        assert torch.equal(v.cpu(), sd[k]), k
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, 3, H, W, generator=g)
    target = torch.rand(B, 4, H, W, generator=g)
    p_old = step.flat_p.clone()
    loss = step.step(x.pin_memory(), target.pin_memory())
    torch.cuda.synchronize()
    rl, _, rg, rstats = O.train_step_grads(sd, x, target, dropout_masks=None)
    assert abs(float(loss) - float(rl)) <= 1e-2 * float(rl)
    grads = {k: step.flat_g[off:off + n].view_as(p) for (k, p), (off, n) in zip(m.named_parameters(), step.ts.lay.values())}
    check_grads(grads, rg, emulation_bounds(sd, x, target, None, 0.0, rg, loss="focal"))
    # the update is the reference AdamW (decay = wd * p_old, not scaled by lr) applied to OUR gradient
    p_exp, _, _ = O.adamw_reference_step(p_old.cpu(), step.flat_g.cpu(), torch.zeros_like(p_old.cpu()), torch.zeros_like(p_old.cpu()), 1, **hyper)
    assert torch.allclose(step.flat_p.cpu(), p_exp, rtol=1e-5, atol=1e-7)
    assert int(step.step_counter.item()) == 1
    # parameters are views of the flat buffer: the module sees the update, and so does state_dict()
    assert torch.equal(m.state_dict()["final_1.weight"].reshape(-1), step.flat_p[step.ts.lay["final_1.weight"][0]:][:64])
    for k, v in rstats.items():
        got = m.state_dict()[k].cpu()
        if k.endswith("num_batches_tracked"):
            assert int(got) == 1
        else:
            assert float((got - v).abs().max()) <= 1e-2 * float(v.abs().max()), k
    # second step: Adam moments and the device step counter advance
    step.step(x.pin_memory(), target.pin_memory())
    torch.cuda.synchronize()
    assert int(step.step_counter.item()) == 2 and float(step.flat_v.abs().max()) > 0


def test_fused_train_step_dropout_masks_change_every_replay():
    m = pkg.UNet_Nested().cuda().train()
    step = fused.FusedTrainStep(m, 1, 32, 32)
    x = torch.randn(1, 3, 32, 32).pin_memory()
    t = torch.rand(1, 4, 32, 32).pin_memory()
    step.step(x, t)
    m1 = step.ts.t["mask0"].clone()
    step.step(x, t)
    m2 = step.ts.t["mask0"].clone()
    assert not torch.equal(m1, m2)
    from unet_nested4tiny_objects_keypoints_b200 import ops
    assert abs(float(ops.unpack_keep_mask(m1).float().mean()) - 0.6) < 0.05


def test_fused_train_step_with_the_shipped_focal_criterion():
    """trainer.py:426 ships FocalLoss_BCE_2d(gamma=3); fused into the head backward like the MSE."""
    sd = O.synth_state_dict(seed=33)
    m = pkg.UNet_Nested()
    m.load_state_dict(sd)
    m = m.cuda().train()
    m.drop_out.p = 0.0
    B, H, W = 2, 32, 32
    step = fused.FusedTrainStep(m, B, H, W, loss="focal", lr=1e-4)
    g = torch.Generator().manual_seed(8)
    x = torch.randn(B, 3, H, W, generator=g)
    target = torch.rand(B, 4, H, W, generator=g)
    loss = step.step(x.pin_memory(), target.pin_memory())
    torch.cuda.synchronize()
    rl, _, rg, _ = O.train_step_grads(sd, x, target, dropout_masks=None, loss="focal")
    assert abs(float(loss) - float(rl)) <= 2e-2 * abs(float(rl))
    grads = {k: step.flat_g[off:off + n].view_as(p) for (k, p), (off, n) in zip(m.named_parameters(), step.ts.lay.values())}
    check_grads(grads, rg, emulation_bounds(sd, x, target, None, 0.0, rg, loss="focal"))


def test_reference_adamw_dropin_optimizer_matches_reference_semantics(golden):
    from unet_nested4tiny_objects_keypoints_b200.optimizers import AdamW
    arr, meta = golden
    h = meta["adamw_hyper"]
    params = [torch.nn.Parameter(torch.from_numpy(arr[f"adamw_p0_{j}"]).cuda()) for j in range(2)]
    opt = AdamW(params, lr=h["lr"], betas=tuple(h["betas"]), eps=h["eps"], weight_decay=h["weight_decay"])
    for s in range(3):
        for j, p in enumerate(params):
            p.grad = torch.from_numpy(arr[f"adamw_g{s}_{j}"]).cuda()
        opt.step()
    for j, p in enumerate(params):
        assert np.allclose(p.detach().cpu().numpy(), arr[f"adamw_p3_{j}"], rtol=2e-6, atol=1e-7)
    with pytest.raises(NotImplementedError):
        AdamW(params, amsgrad=True)


def test_fused_train_step_from_keypoints_builds_targets_on_the_device():
    sd = O.synth_state_dict(seed=34)
    m = pkg.UNet_Nested()
    m.load_state_dict(sd)
    m = m.cuda().train()
    m.drop_out.p = 0.0
    B, H, W = 2, 32, 32
    step = fused.FusedTrainStep(m, B, H, W, loss="mse")
    g = torch.Generator().manual_seed(9)
    x = torch.randn(B, 3, H, W, generator=g)
    kp = (torch.rand(B, 7, 2, generator=g) * 28 + 2).float()
    loss = step.step_keypoints(x.pin_memory(), kp.pin_memory())
    torch.cuda.synchronize()
    target = torch.from_numpy(O.create_heatmap(kp.numpy(), H, W))  # trainer.py:122-123
    assert float((step.target.cpu() - target).abs().max()) <= 2e-6
    rl, _, _, _ = O.train_step_grads(sd, x, target, dropout_masks=None)
    assert abs(float(loss) - float(rl)) <= 1e-2 * float(rl)


def test_pipelined_sessions_equal_the_sequential_calls():
    """run_many / step_many overlap the H2D copy of batch k+1 with the kernels of batch k: same results, bit for bit."""
    sd = O.synth_state_dict(seed=35)
    g = torch.Generator().manual_seed(10)
    B, H, W, K = 2, 32, 32, 5
    xs = [torch.randn(B, 3, H, W, generator=g).pin_memory() for _ in range(K)]
    ts = [torch.rand(B, 4, H, W, generator=g).pin_memory() for _ in range(K)]
    # inference
    m = pkg.UNet_Nested()
    m.load_state_dict(sd)
    m = m.cuda().eval()
    sess = fused.InferenceSession(m, B, H, W, head=2)
    seq = []
    for x in xs:
        xy, val = sess.run(x)
        torch.cuda.synchronize()
        seq.append((xy.cpu().clone(), val.cpu().clone()))
    xy_h = [torch.empty(B, 4, 2, dtype=torch.int32).pin_memory() for _ in range(K)]
    val_h = [torch.empty(B, 4, dtype=torch.float32).pin_memory() for _ in range(K)]
    assert sess.run_many(xs, xy_h, val_h) == K
    torch.cuda.synchronize()
    for k in range(K):
        assert torch.equal(xy_h[k], seq[k][0]) and torch.equal(val_h[k], seq[k][1]), k
    # training: two identical models, one stepped sequentially, one through the pipeline
    losses = []
    for mode in ("seq", "pipe"):
        mt = pkg.UNet_Nested()
        mt.load_state_dict(sd)
        mt = mt.cuda().train()
        mt.drop_out.p = 0.0
        step = fused.FusedTrainStep(mt, B, H, W, lr=1e-3)
        if mode == "seq":
            out = []
            for x, t in zip(xs, ts):
                out.append(float(step.step(x, t)))
        else:
            lh = [torch.empty(1).pin_memory() for _ in range(K)]
            assert step.step_many(xs, ts, lh) == K
            torch.cuda.synchronize()
            out = [float(l) for l in lh]
        losses.append((out, step.flat_p.clone()))
    assert losses[0][0] == losses[1][0]
    assert torch.equal(losses[0][1], losses[1][1])


@pytest.mark.parametrize("opt", ["adamw", "sgd"])
def test_tar_checkpoint_resume_round_trip(tmp_path, opt):
    """trainer/trainer.py:402-413: a ``.tar`` with model_state_dict / optimizer_state_dict / epoch.  Three uninterrupted fused steps ==
    two steps, save, a NEW model + step restored from the file, one more step (bit for bit: weights, moments, step count, dropout
    stream); and the optimizer state loads into the matching torch-style optimizer (the format the reference trainer expects)."""
    sd = O.synth_state_dict(seed=71)
    g = torch.Generator().manual_seed(3)
    B, H, W = 2, 32, 32
    xs = [torch.randn(B, 3, H, W, generator=g).pin_memory() for _ in range(3)]
    ts = [torch.rand(B, 4, H, W, generator=g).pin_memory() for _ in range(3)]

    def fresh():
        m = pkg.UNet_Nested()
        m.load_state_dict(sd)
        m = m.cuda().train()
        return m, fused.FusedTrainStep(m, B, H, W, lr=1e-3, weight_decay=1e-2, optimizer=opt, momentum=0.9, loss="focal", seed=5)

    m_a, st_a = fresh()
    for k in range(3):
        st_a.step(xs[k], ts[k])
    m_b, st_b = fresh()
    for k in range(2):
        st_b.step(xs[k], ts[k])
    path = str(tmp_path / "ck.tar")
    st_b.save_checkpoint(path, epoch=7)
    m_c, st_c = fresh()
    assert st_c.load_checkpoint(path) == 8
    st_c.step(xs[2], ts[2])
    torch.cuda.synchronize()
    assert torch.equal(st_c.flat_p, st_a.flat_p) and torch.equal(st_c.flat_m, st_a.flat_m) and torch.equal(st_c.flat_v, st_a.flat_v)
    assert int(st_c.step_counter) == 3
    for (k, a), b in zip(m_a.state_dict().items(), m_c.state_dict().values()):
        assert torch.equal(a, b), k
    # the file is what the reference trainer resumes from: plain keys, and the optimizer state loads into a torch-style optimizer
    ck = torch.load(path)
    assert set(ck) == {"epoch", "model_state_dict", "optimizer_state_dict"} and list(ck["model_state_dict"]) == list(sd)
    from unet_nested4tiny_objects_keypoints_b200 import optimizers
    m_d = pkg.UNet_Nested().cuda()
    o = optimizers.AdamW(m_d.parameters(), lr=1e-3, weight_decay=1e-2) if opt == "adamw" else torch.optim.SGD(m_d.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-2)
    o.load_state_dict(ck["optimizer_state_dict"])
    st0 = o.state[next(iter(m_d.parameters()))]
    if opt == "adamw":
        assert st0["step"] == 2 and torch.equal(st0["exp_avg"].reshape(-1), st_b.flat_m[:st0["exp_avg"].numel()])
    else:
        assert torch.equal(st0["momentum_buffer"].reshape(-1), st_b.flat_m[:st0["momentum_buffer"].numel()])
    # ... and the other way round: a state saved by that optimizer resumes in the fused step
    st_e = fresh()[1]
    st_e.load_optimizer_state_dict(o.state_dict())
    assert torch.equal(st_e.flat_m, st_b.flat_m) and int(st_e.step_counter) >= 1


def test_fused_step_from_uint8_images_and_keypoints_like_the_trainers_loader():
    """The trainer's loader delivers (inputs, labels) = images + key points (trainer.py:106-109).  step_many with 8-bit images and key
    points: ToTensor's 1/255 and helper.create_heatmap run on the device; the result equals the step fed with the fp32 tensors the
    reference would have built on the CPU (same x16 after the bf16 rounding, targets to 2e-6 -> same loss to 1e-6)."""
    sd = O.synth_state_dict(seed=72)
    g = torch.Generator().manual_seed(6)
    B, H, W = 2, 32, 32
    u8 = [torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8).pin_memory() for _ in range(2)]
    kps = [(torch.rand(B, 7, 2, generator=g) * 24 + 4).pin_memory() for _ in range(2)]
    res = []
    for mode in ("u8", "f32"):
        m = pkg.UNet_Nested()
        m.load_state_dict(sd)
        m = m.cuda().train()
        m.drop_out.p = 0.0
        st = fused.FusedTrainStep(m, B, H, W, lr=1e-3, loss="mse", input="uint8_nhwc" if mode == "u8" else "float32")
        lh = [torch.empty(1).pin_memory() for _ in range(2)]
        if mode == "u8":
            st.step_many(u8, kps, lh)
        else:
            xs = [(u.permute(0, 3, 1, 2).float() / 255).contiguous().pin_memory() for u in u8]          # torchvision ToTensor
            ts = [torch.from_numpy(O.create_heatmap(k.numpy(), H, W)).pin_memory() for k in kps]       # trainer.py:122-123
            st.step_many(xs, ts, lh)
        torch.cuda.synchronize()
        res.append(([float(l) for l in lh], st.ts.t["x16"].clone(), st.flat_g.clone()))
    assert torch.equal(res[0][1], res[1][1])  # identical bf16 input tensor
    for a, b in zip(res[0][0], res[1][0]):
        assert abs(a - b) <= 1e-4 * abs(b)
    assert float((res[0][2] - res[1][2]).abs().max()) <= 2e-2 * float(res[1][2].abs().max())  # (the second step starts from parameters that differ in the last bits)
