"""GPU: the cases where the reference (plain autograd / cuDNN) is robust by construction and a cached, raw-pointer engine
must be made so — interleaved forwards, weights that change under a captured graph, optimizers that write through raw
pointers, replicas, a current device other than the module's."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import unet_nested4tiny_objects_keypoints_b200 as pkg  # noqa: E402
from unet_nested4tiny_objects_keypoints_b200 import fused, ops, optimizers  # noqa: E402
from oracle import unetpp_oracle as O  # noqa: E402


def _model(seed, train, dev="cuda"):
    m = pkg.UNet_Nested()
    m.load_state_dict(O.synth_state_dict(seed=seed))
    m = m.to(dev)
    return m.train() if train else m.eval()


def test_two_forwards_before_one_backward_like_autograd():
    """o1 = model(x1); o2 = model(x2); (l1 + l2).backward(): every forward keeps its own saved activations."""
    g = torch.Generator().manual_seed(3)
    x1, x2 = torch.randn(2, 3, 32, 32, generator=g).cuda(), torch.randn(2, 3, 32, 32, generator=g).cuda()
    t = torch.rand(2, 4, 32, 32, generator=g).cuda()

    def loss_of(m, x):
        return sum(F.mse_loss(o, t) for o in m(x)) / 3

    m = _model(61, True)
    m.drop_out.p = 0.0
    (loss_of(m, x1) + loss_of(m, x2)).backward()
    both = torch.cat([p.grad.reshape(-1) for p in m.parameters()]).clone()
    sep = []
    for x in (x1, x2):
        ms = _model(61, True)
        ms.drop_out.p = 0.0
        loss_of(ms, x).backward()
        sep.append(torch.cat([p.grad.reshape(-1) for p in ms.parameters()]))
    assert torch.equal(both, sep[0] + sep[1])
    # running statistics saw both batches, in order
    ms = _model(61, True)
    ms.drop_out.p = 0.0
    with torch.no_grad():
        ms(x1), ms(x2)
    for (k, a), b in zip(m.state_dict().items(), ms.state_dict().values()):
        assert torch.equal(a, b), k


def test_backward_after_the_activations_were_overwritten_raises():
    m = _model(62, True)
    x = torch.randn(1, 3, 32, 32).cuda()
    out = m(x)
    out[0].sum().backward(retain_graph=True)
    m(x)  # same shape: the state is free again (its backward ran) and is reused
    with pytest.raises(RuntimeError, match="overwritten"):
        out[0].sum().backward()


def test_inference_session_follows_weight_changes_and_survives_arena_eviction():
    m = _model(63, False)
    x = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(1))
    sess = fused.InferenceSession(m, 2, 32, 32, head=2)
    sess.run(x.pin_memory())
    torch.cuda.synchronize()
    assert float((sess.heat.cpu() - O.forward(O.synth_state_dict(seed=63), x)[2]).abs().max()) <= 3e-2
    # five other shapes evict this shape's arena from the engine's cache; the session keeps its own alive
    with torch.no_grad():
        for s in (16, 24, 40, 48, 56):
            m(torch.randn(1, 3, s, s).cuda())
    first = sess.heat.clone()
    sess.run(x.pin_memory())
    torch.cuda.synchronize()
    assert torch.equal(sess.heat, first)
    # new weights under the captured graph: the next run re-folds and re-captures
    sd2 = O.synth_state_dict(seed=64)
    m.load_state_dict(sd2)
    sess.run(x.pin_memory())
    torch.cuda.synchronize()
    assert float((sess.heat.cpu() - O.forward(sd2, x)[2]).abs().max()) <= 3e-2
    assert not torch.equal(sess.heat, first)


@pytest.mark.parametrize("opt", ["adamw", "sgdw", "adabound"])
def test_eval_after_a_dropin_optimizer_step_uses_the_new_weights(opt):
    """The drop-in optimizers write through raw pointers; the eval-mode fold cache is keyed on tensor versions."""
    m = _model(65, False)
    x = torch.randn(1, 3, 32, 32, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        before = m(x.cuda())[2].clone()
    o = {"adamw": lambda: optimizers.AdamW(m.parameters(), lr=5e-2, weight_decay=1e-2), "sgdw": lambda: optimizers.SGDW(m.parameters(), lr=0.1, weight_decay=0.2),
         "adabound": lambda: optimizers.AdaBound(m.parameters(), lr=5e-2)}[opt]()
    for p in m.parameters():
        p.grad = torch.ones_like(p) * 0.1
    o.step()
    with torch.no_grad():
        after = m(x.cuda())[2]
    ref = O.forward({k: v.cpu() for k, v in m.state_dict().items()}, x)[2]
    assert float((after.cpu() - ref).abs().max()) <= 3e-2
    assert float((after - before).abs().max()) > 1e-3


def test_eval_after_a_fused_train_step_uses_the_new_weights_and_statistics():
    m = _model(66, True)
    step = fused.FusedTrainStep(m, 2, 32, 32, lr=1e-2, loss="mse")
    g = torch.Generator().manual_seed(4)
    x, t = torch.randn(2, 3, 32, 32, generator=g), torch.rand(2, 4, 32, 32, generator=g)
    m.eval()
    with torch.no_grad():
        before = m(x.cuda())[2].clone()
    m.train()
    step.step(x.pin_memory(), t.pin_memory())
    m.eval()
    with torch.no_grad():
        after = m(x.cuda())[2]
    ref = O.forward({k: v.cpu() for k, v in m.state_dict().items()}, x)[2]
    assert float((after.cpu() - ref).abs().max()) <= 3e-2 and not torch.equal(after, before)


def test_create_heatmap_with_six_points_normalises_plane_three():
    """helper.py:148-159: planes 1 and 3 are ALWAYS divided by their maximum, also when plane 3 holds one point."""
    kp = (torch.rand(3, 6, 2, generator=torch.Generator().manual_seed(6)) * 30 + 1.3).float()
    got = ops.create_heatmap(kp.cuda(), 32, 40).cpu().numpy()
    ref = O.create_heatmap(kp.numpy(), 32, 40)
    assert np.abs(got - ref).max() <= 2e-6
    assert np.allclose(got[:, 3].reshape(3, -1).max(1), 1.0) and got[:, 2].max() < 1.0


@pytest.mark.parametrize("planes,H,W", [(3, 1024, 1024), (8, 512, 768), (64, 1024, 1024), (5, 250, 250)])
def test_split_argmax_is_bit_identical_to_numpy(planes, H, W):
    g = torch.Generator().manual_seed(planes)
    heat = torch.rand(1, planes, H, W, generator=g)
    heat[0, 0] = 0.5                       # a constant plane: the first index wins
    heat[0, 1, H - 1, W - 1] = 2.0         # maximum in the very last element (last segment)
    if planes > 2:
        heat[0, 2, 7, 9] = heat[0, 2, H // 2, 3] = 3.0  # a tie across segments: the smaller index wins
    assert ops.lib().unpp_argmax_splits(planes, H, W) > 1 or planes * H * W < 1 << 20
    xy, val = ops.argmax_peaks(heat.cuda())
    rxy, rval = O.argmax_keypoints(heat.numpy())
    assert np.array_equal(xy.cpu().numpy(), rxy) and np.array_equal(val.cpu().numpy(), rval)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_module_on_device_one_while_the_current_device_is_zero():
    torch.cuda.set_device(0)
    sd = O.synth_state_dict(seed=67)
    m = _model(67, False, "cuda:1")
    x = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        out = m(x.to("cuda:1"))[2]
    xy, val, _ = m.predict_keypoints(x.to("cuda:1"))
    torch.cuda.synchronize(1)
    assert out.device.index == 1 and float((out.cpu() - O.forward(sd, x)[2]).abs().max()) <= 3e-2
    assert np.array_equal(xy.cpu().numpy(), O.argmax_keypoints(out.cpu().numpy())[0])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_dataparallel_eval_sees_a_load_state_dict_between_two_forwards():
    """Replicas are fresh broadcast copies (version 0, recycled addresses): they must never be served cached folded weights."""
    m = _model(68, False)
    dp = torch.nn.DataParallel(m, device_ids=[0, 1])
    x = torch.randn(4, 3, 32, 32, generator=torch.Generator().manual_seed(8))
    with torch.no_grad():
        a = dp(x.cuda())[2].cpu()
    assert float((a - O.forward(O.synth_state_dict(seed=68), x)[2]).abs().max()) <= 3e-2
    sd2 = O.synth_state_dict(seed=69)
    m.load_state_dict(sd2)
    with torch.no_grad():
        b = dp(x.cuda())[2].cpu()
    assert float((b - O.forward(sd2, x)[2]).abs().max()) <= 3e-2


def test_topk_peaks_kernel_matches_the_oracle():
    """unpp_topk_peaks (the multi-point form of Heatmap.extract_points_, tools/misc/heatmap.py:148-208) bit for bit against the oracle
    restatement, which tests/test_oracle_golden.py pins against the real reference: fixture planes, noisy planes with hundreds of
    local maxima, plateaus, empty planes, the retry threshold."""
    import json, os
    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    arr = dict(np.load(os.path.join(g, "unetpp_r2.npz")))
    cases = [(torch.from_numpy(arr["peaks_planes"]), 3, 0.5), (torch.from_numpy(arr["peaks_low"]), 2, 0.5)]
    gen = torch.Generator().manual_seed(4)
    noisy = torch.rand(2, 3, 40, 56, generator=gen)          # ~ 1/9 of the pixels above 0.5 are local maxima
    noisy[0, 0] = 0.25                                         # nothing above either threshold
    noisy[0, 1] = 0.75                                         # one plateau: a single peak at index 0
    noisy[1, 2, 5:9, 7:12] = 2.0                               # a rectangular plateau inside noise
    cases += [(noisy, 5, 0.5), (noisy, 1, 0.9), (torch.rand(1, 2, 1024, 1024, generator=gen), 4, 0.99)]
    for heat, num, thr in cases:
        xy, val, cnt = ops.topk_peaks(heat.cuda(), num, thr)
        rxy, rval, rcnt = O.topk_peaks(heat.numpy(), num, thr) if heat.shape[-1] <= 64 else (None, None, None)
        if rxy is None:  # large planes: the python oracle is too slow; check against a torch restatement of the same definition
            h = heat.cuda()
            pad = torch.nn.functional.pad(h, (1, 1, 1, 1), value=-1.0)
            idx = torch.arange(h.shape[2] * h.shape[3], device="cuda").view(1, 1, *h.shape[2:]).expand_as(h)
            pidx = torch.nn.functional.pad(idx.float(), (1, 1, 1, 1), value=1e12)
            ok = h >= thr
            for dy in range(3):
                for dx in range(3):
                    if dy == 1 and dx == 1:
                        continue
                    u, q = pad[:, :, dy:dy + h.shape[2], dx:dx + h.shape[3]], pidx[:, :, dy:dy + h.shape[2], dx:dx + h.shape[3]]
                    ok &= ~((u > h) | ((u == h) & (q < idx)))
            for b in range(h.shape[0]):
                for c in range(h.shape[1]):
                    cand = torch.nonzero(ok[b, c].flatten()).flatten()                     # candidates in index order
                    _, o = torch.sort(h[b, c].flatten()[cand], descending=True, stable=True)  # value desc, ties keep the index order
                    order = cand[o][:num]
                    n = order.numel()
                    assert int(cnt[b, c]) == n
                    assert torch.equal(xy[b, c, :n, 0].long(), order % h.shape[3]) and torch.equal(xy[b, c, :n, 1].long(), order // h.shape[3])
            continue
        assert np.array_equal(cnt.cpu().numpy(), rcnt) and np.array_equal(xy.cpu().numpy(), rxy) and np.array_equal(val.cpu().numpy(), rval)


def test_extract_points_api_on_model_outputs():
    m = _model(70, False)
    x = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        heats = m(x.cuda())
    xy, val, cnt = m.extract_points(heats[2], 2, threshold=0.3)
    rxy, rval, rcnt = O.topk_peaks(heats[2].cpu().numpy(), 2, 0.3)
    assert np.array_equal(xy.cpu().numpy(), rxy) and np.array_equal(val.cpu().numpy(), rval) and np.array_equal(cnt.cpu().numpy(), rcnt)


def test_edge_shapes_empty_batch_smallest_and_ragged():
    """Edge cases of the reference's layers: an empty batch gives empty outputs in eval mode; the smallest legal image is 8x8 (one pixel at
    the deepest level); extreme aspect ratios; a single value per channel at the deepest level is an error in train mode, like
    nn.BatchNorm2d's ("Expected more than 1 value per channel when training")."""
    sd = O.synth_state_dict(seed=73)
    m = _model(73, False)
    with torch.no_grad():
        outs = m(torch.empty(0, 3, 32, 32, device="cuda"))
    assert all(tuple(o.shape) == (0, 4, 32, 32) for o in outs)
    xy, val, _ = m.predict_keypoints(torch.empty(0, 3, 32, 32, device="cuda"))
    assert tuple(xy.shape) == (0, 4, 2) and tuple(val.shape) == (0, 4)
    for shape in [(1, 3, 8, 8), (3, 3, 8, 8), (1, 3, 8, 2048), (1, 3, 1024, 8), (2, 3, 24, 8)]:
        x = torch.randn(*shape, generator=torch.Generator().manual_seed(shape[3]))
        with torch.no_grad():
            outs = m(x.cuda())
        for o, r in zip(outs, O.forward(sd, x)):
            assert float((o.cpu() - r).abs().max()) <= 3e-2, shape
        xy, _, heats = m.predict_keypoints(x.cuda())
        assert np.array_equal(xy.cpu().numpy(), O.argmax_keypoints(heats[2].cpu().numpy())[0])
    m.train()
    with pytest.raises(ValueError, match="more than 1 value per channel"):
        m(torch.randn(1, 3, 8, 8, device="cuda"))
    with pytest.raises(ValueError, match="divisible by 8"):
        m(torch.randn(1, 3, 12, 16, device="cuda"))
    out = m(torch.randn(2, 3, 8, 8, device="cuda"))  # two values per channel at the deepest level: legal
    sum(o.sum() for o in out).backward()
    assert all(torch.isfinite(p.grad).all() for p in m.parameters())
