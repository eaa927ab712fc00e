"""GPU: the BASELINE.json configurations at their FULL sizes, checked through size-independent properties plus the
oracle on a bounded sample (the CPU oracle needs ~0.1 s per 256x256 image, ~7 s for a batch-32 training step).

  * configs[1]  batch 128 inference at 256x256 + arg-max key points
  * configs[2]  batch 32 training step at 256x256
  * configs[4]  batch 16 inference at 1024x1024

Properties: run-to-run determinism (bit-identical), batch independence in eval mode (an image gives bit-identical heat maps
whatever batch it travels in: tiles are scheduled differently, the per-pixel arithmetic is not), exact linearity of the
backward pass in its upstream gradient under a power-of-two scale, arg-max key points bit-exact against numpy on the heat
maps the GPU produced, and the stated bf16 bound against the oracle on the sample."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import unet_nested4tiny_objects_keypoints_b200 as pkg  # noqa: E402
from unet_nested4tiny_objects_keypoints_b200 import fused, ops  # noqa: E402
from oracle import unetpp_oracle as O  # noqa: E402
from oracle import teacher_forced as T  # noqa: E402
from test_training_gpu import check_grads, stored_tensors, table_bounds  # noqa: E402


def _model(seed, train=False):
    m = pkg.UNet_Nested()
    m.load_state_dict(O.synth_state_dict(seed=seed))
    m = m.cuda()
    return m.train() if train else m.eval()


@pytest.mark.parametrize("B,S,sample", [(128, 256, (0, 77, 127)), (16, 1024, (5,))])
def test_full_size_inference_properties(B, S, sample):
    sd = O.synth_state_dict(seed=41)
    m = _model(41)
    x = torch.randn(B, 3, S, S, generator=torch.Generator().manual_seed(B))
    sess = fused.InferenceSession(m, B, S, S, head=2)
    xy, val = sess.run(x.pin_memory())
    torch.cuda.synchronize()
    heat = sess.heat.clone()
    xy, val = xy.clone(), val.clone()
    # determinism
    xy2, val2 = sess.run(x.pin_memory())
    torch.cuda.synchronize()
    assert torch.equal(sess.heat, heat) and torch.equal(xy2, xy) and torch.equal(val2, val)
    # arg-max key points: bit-exact against numpy on identical heat maps, every plane of the batch
    rxy, rval = O.argmax_keypoints(heat.cpu().numpy())
    assert np.array_equal(xy.cpu().numpy(), rxy) and np.array_equal(val.cpu().numpy(), rval)
    # batch independence + the oracle on the sample
    idx = list(sample)
    with torch.no_grad():
        small = m(x[idx].cuda())[2]
    assert torch.equal(small, heat[idx])
    ref = O.forward(sd, x[idx])[2]
    err = (heat[idx].cpu() - ref).abs()
    assert float(err.max()) <= 3e-2 and float(err.mean()) <= 3e-3


def test_full_size_training_step_properties():
    B, S = 32, 256
    sd = O.synth_state_dict(seed=42)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, 3, S, S, generator=g)
    target = torch.rand(B, 4, S, S, generator=g)

    def grads_for(scale):
        m = _model(42, train=True)
        m.drop_out.p = 0.0
        outs = m(x.cuda())
        loss = sum(torch.nn.functional.mse_loss(o, target.cuda()) for o in outs) / 3
        (loss * scale).backward()
        torch.cuda.synchronize()
        return float(loss.detach()), {k: p.grad.clone() for k, p in m.named_parameters()}, m

    loss1, g1, m1 = grads_for(1.0)
    # teacher-forced parity at the full size (oracle/teacher_forced.py, fp64 on the GPU): every stored tensor of THIS step to
    # one bf16 ulp, every one of the 74 parameter gradients to 1e-3
    t, heats = stored_tensors(m1, B, S, S)
    rep = T.verify_step(t, heats, g1, sd, x.cuda(), target=target)
    assert not rep.check(), rep.check()
    wb, wf = rep.worst("bf16"), rep.worst("f32")
    print(f"teacher-forced B={B} {S}x{S}: worst bf16 tensor {wb['name']} {wb['max_rel']:.2e} (> 1 ulp: {wb['frac_gt_ulp']:.1e}); worst fp32 {wf['name']} {wf['max_rel']:.2e}")
    del rep, t, heats
    torch.cuda.empty_cache()
    loss1b, g1b, _ = grads_for(1.0)
    assert loss1 == loss1b
    for k in g1:  # determinism: fixed-order reductions everywhere
        assert torch.equal(g1[k], g1b[k]), k
    _, g4, _ = grads_for(4.0)
    for k in g1:  # a power-of-two scale of the upstream gradient passes through every bf16 rounding exactly
        assert torch.equal(g4[k], 4.0 * g1[k]), k
    # the oracle on the full batch (BatchNorm couples the images, so there is no smaller sample): loss, gradients, statistics
    rl, _, rg, rstats = O.train_step_grads(sd, x, target, dropout_masks=None)
    assert abs(loss1 - float(rl)) <= 1e-2 * float(rl)
    check_grads(g1, rg, table_bounds("b32_256"))  # 2x the committed per-tensor table profiles/r02_grad_errors_b32_256.md (same seeds)
    new_sd = m1.state_dict()
    for k, v in rstats.items():
        if not k.endswith("num_batches_tracked"):
            assert float((new_sd[k].cpu() - v).abs().max()) <= 1e-2 * float(v.abs().max()), k


@pytest.mark.parametrize("cins,cout,H,taps", [([16, 16, 16, 16, 16], 16, 256, 9), ([32, 32, 32], 32, 128, 9), ([128], 128, 32, 9), ([32], 16, 128, 1)])
def test_full_size_weight_gradient_linearity_and_reference(cins, cout, H, taps):
    """wgrad at batch 32: dW(X, 2*dZ) == 2*dW(X, dZ) bit for bit, and dW against torch's fp64 weight gradient."""
    N = 32
    gen = torch.Generator().manual_seed(H + cout)
    xs = [torch.randn(N, H, H, c, generator=gen).to(torch.bfloat16).cuda() for c in cins]
    if taps == 9:
        dz = (torch.randn(N, H, H, cout, generator=gen) * 0.5).to(torch.bfloat16).cuda()
        view = None
    else:  # the four taps of a k2s2 transposed conv: dz is the [N, 2H, 2W, cout] gradient of its output
        dz = (torch.randn(N, 2 * H, 2 * H, cout, generator=gen) * 0.5).to(torch.bfloat16).cuda()
        view = "all4"
    cin = sum(cins)
    grid = ops.wgrad_grid(cins, N, H, H, cout, taps, dz_view=view)
    nparts = grid * (4 if view else 1)

    def run(d):
        part = torch.empty(nparts, taps, cin, cout, device="cuda")
        ops.wgrad(xs, N, H, H, d, cout, taps, part, dz_view=view)
        return part

    p1, p2 = run(dz), run(dz * 2)
    assert torch.equal(p2, 2 * p1)
    assert torch.equal(run(dz), p1)
    if taps == 9:
        got = p1.double().sum(0).permute(2, 1, 0).reshape(cout, cin, 3, 3)  # [tap][ci][co] -> [co][ci][r][s]
        xcat = torch.cat([t.double().permute(0, 3, 1, 2) for t in xs], 1)
        ref = torch.nn.grad.conv2d_weight(xcat, (cout, cin, 3, 3), dz.double().permute(0, 3, 1, 2), padding=1)
    else:
        got = p1.double().view(4, grid, cin, cout).sum(1).permute(1, 2, 0).reshape(cin, cout, 2, 2)
        w = torch.zeros(cin, cout, 2, 2, dtype=torch.double, device="cuda", requires_grad=True)
        torch.nn.functional.conv_transpose2d(xs[0].double().permute(0, 3, 1, 2), w, stride=2).backward(dz.double().permute(0, 3, 1, 2))
        ref = w.grad
    err = float((got - ref).abs().max()) / float(ref.abs().max())
    assert err <= 2e-3, err
