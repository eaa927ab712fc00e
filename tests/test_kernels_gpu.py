"""GPU: each libunpp.so kernel, called through the C ABI (ops.py -> ctypes), against a plain
torch-CPU fp32/fp64 restatement of the same op on the same bf16-rounded inputs.

Tolerances: the kernels read bf16, accumulate in fp32 and store bf16, so against an fp64 reference
on identical bf16 inputs the error is the output rounding (2^-9 relative) plus fp32 accumulation
noise: |err| <= 6e-3 * max|ref| for bf16 outputs, 2e-3 for fp32 outputs.  Integer outputs (arg-max)
are bit-exact.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from unet_nested4tiny_objects_keypoints_b200 import ops  # noqa: E402

DEV = "cuda"


def bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


def nhwc(t):  # NCHW fp32 cpu -> NHWC bf16 cuda
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(DEV)


def nchw(t):  # NHWC bf16 cuda -> NCHW fp32 cpu
    return t.float().cpu().permute(0, 3, 1, 2).contiguous()


def close(got, ref, rel, what=""):
    ref = ref.double()
    err = float((got.double() - ref).abs().max())
    scale = float(ref.abs().max()) + 1e-30
    assert err <= rel * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e} (rel {err / scale:.3e} > {rel})"


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


@pytest.mark.parametrize("N,H,W,cins,cout", [
    (2, 24, 40, [16], 16),
    (1, 16, 16, [64], 64),
    (2, 8, 8, [128], 128),
    (1, 40, 72, [32], 32),
    (1, 32, 32, [16, 16, 16, 16], 16),
    (2, 16, 24, [32, 32, 32], 32),
    (1, 16, 16, [64, 64], 64),
    (1, 8, 8, [64], 128),
    (3, 64, 64, [16, 16], 16),
    (1, 128, 128, [16], 16),
])
def test_conv3x3_bias_relu_virtual_concat(N, H, W, cins, cout):
    cin = sum(cins)
    xs = [bf(rnd(N, c, H, W, seed=i + 1)) for i, c in enumerate(cins)]
    w = bf(rnd(cout, cin, 3, 3, seed=50, scale=(2.0 / (9 * cin)) ** 0.5))
    b = rnd(cout, seed=51, scale=0.1)
    ref = F.relu(F.conv2d(torch.cat(xs, 1).double(), w.double(), b.double(), padding=1))
    nt = ops.pick_n_tile(cout, cin, 9)
    wp = ops.pack_weights(w.to(DEV), 0, 9, cout, nt, cin)
    out = torch.empty(N, H, W, cout, dtype=torch.bfloat16, device=DEV)
    ops.conv([nhwc(x) for x in xs], N, H, W, wp, cout, nt, 9, bias=b.to(DEV), relu=True, out=out)
    torch.cuda.synchronize()
    close(nchw(out), ref, 6e-3, "conv3x3")


@pytest.mark.parametrize("n_tile", [16, 32, 64])
def test_conv3x3_n_tile_split(n_tile):
    N, H, W, cin, cout = 1, 16, 24, 32, 64
    x = bf(rnd(N, cin, H, W, seed=3))
    w = bf(rnd(cout, cin, 3, 3, seed=4, scale=0.1))
    ref = F.conv2d(x.double(), w.double(), padding=1)
    wp = ops.pack_weights(w.to(DEV), 0, 9, cout, n_tile, cin)
    out = torch.empty(N, H, W, cout, dtype=torch.bfloat16, device=DEV)
    ops.conv([nhwc(x)], N, H, W, wp, cout, n_tile, 9, out=out)
    close(nchw(out), ref, 6e-3, "conv n_tile")


def test_bn_fold_scale_in_pack():
    N, H, W, cin, cout = 1, 16, 16, 16, 32
    x = bf(rnd(N, cin, H, W, seed=5))
    w = rnd(cout, cin, 3, 3, seed=6, scale=0.1)
    scale = torch.rand(cout) + 0.5
    ref = F.conv2d(x.double(), bf(w * scale[:, None, None, None]).double(), padding=1)
    wp = ops.pack_weights(w.to(DEV), 0, 9, cout, 32, cin, scale=scale.to(DEV))
    out = torch.empty(N, H, W, cout, dtype=torch.bfloat16, device=DEV)
    ops.conv([nhwc(x)], N, H, W, wp, cout, 32, 9, out=out)
    close(nchw(out), ref, 6e-3, "bn fold")


@pytest.mark.parametrize("N,H,W,cin,cout", [(2, 8, 12, 32, 16), (1, 16, 16, 64, 32), (1, 4, 4, 128, 64), (1, 64, 64, 32, 16)])
def test_deconv_k2s2(N, H, W, cin, cout):
    x = bf(rnd(N, cin, H, W, seed=7))
    w = bf(rnd(cin, cout, 2, 2, seed=8, scale=0.2))
    b = rnd(cout, seed=9, scale=0.1)
    ref = F.conv_transpose2d(x.double(), w.double(), b.double(), stride=2)
    nt = ops.pick_n_tile(4 * cout, cin, 1, deconv=True)
    wp = ops.pack_weights(w.to(DEV), 2, 1, 4 * cout, nt, cin)
    out = torch.empty(N, 2 * H, 2 * W, cout, dtype=torch.bfloat16, device=DEV)
    ops.conv([nhwc(x)], N, H, W, wp, 4 * cout, nt, 1, bias=b.to(DEV), mode=ops.MODE_DECONV, out=out)
    close(nchw(out), ref, 6e-3, "deconv")


@pytest.mark.parametrize("with_mask", [False, True])
def test_fused_head_sigmoid_dropout(with_mask):
    N, H, W = 2, 16, 24
    x = bf(rnd(N, 16, H, W, seed=10))
    w = bf(rnd(16, 16, 3, 3, seed=11, scale=0.12))
    b = rnd(16, seed=12, scale=0.1)
    hw, hb = rnd(4, 16, seed=13, scale=0.5), rnd(4, seed=14, scale=0.2)
    y = F.relu(F.conv2d(x.double(), w.double(), b.double(), padding=1))
    mask = (torch.rand(N, 16, H, W, generator=torch.Generator().manual_seed(15)) >= 0.4)
    yd = y * mask.double() / 0.6 if with_mask else y
    logit = F.conv2d(yd, hw.double()[:, :, None, None], hb.double())
    wp = ops.pack_weights(w.to(DEV), 0, 9, 16, 16, 16)
    out = torch.empty(N, H, W, 16, dtype=torch.bfloat16, device=DEV)
    heat = torch.empty(N, 4, H, W, device=DEV)
    lg = torch.empty(N, 4, H, W, device=DEV)
    m8 = ops.pack_keep_mask(mask.to(DEV)) if with_mask else None
    ops.conv([nhwc(x)], N, H, W, wp, 16, 16, 9, bias=b.to(DEV), relu=True, out=out,
             head=(hw.to(DEV), hb.to(DEV), heat, lg, m8, 1 / 0.6 if with_mask else 1.0))
    close(nchw(out), y, 6e-3, "head conv out")
    close(lg.cpu(), logit, 2e-3, "logit")
    close(heat.cpu(), torch.sigmoid(logit), 2e-3, "heat")


def test_input_layout_and_maxpool():
    x = rnd(2, 3, 16, 24, seed=16)
    out = torch.empty(2, 16, 24, 16, dtype=torch.bfloat16, device=DEV)
    ops.nchw_to_nhwc16(x.to(DEV), out)
    got = nchw(out)
    assert torch.equal(got[:, :3], bf(x)) and float(got[:, 3:].abs().max()) == 0.0
    y = bf(rnd(2, 32, 16, 24, seed=17))
    p = torch.empty(2, 8, 12, 32, dtype=torch.bfloat16, device=DEV)
    ops.maxpool(nhwc(y), p)
    assert torch.equal(nchw(p), F.max_pool2d(y, 2))


def test_argmax_bit_exact(golden):
    from oracle import unetpp_oracle as O
    arr, _ = golden
    for planes in (arr["peaks_planes"], arr["tie_plane"], np.random.default_rng(0).random((3, 4, 40, 56), dtype=np.float32)):
        xy, val = ops.argmax_peaks(torch.from_numpy(planes).to(DEV))
        rxy, rval = O.argmax_keypoints(planes)
        assert np.array_equal(xy.cpu().numpy(), rxy)
        assert np.array_equal(val.cpu().numpy(), rval)
    assert np.array_equal(ops.argmax_peaks(torch.from_numpy(arr["peaks_planes"]).to(DEV))[0].cpu().numpy(), arr["peaks_xy"])
    # plateau / constant / last-pixel / odd-size planes
    z = np.zeros((1, 3, 9, 13), dtype=np.float32)
    z[0, 1, 8, 12] = 1.0
    z[0, 2, 3:6, 4:9] = 0.5
    xy, _ = ops.argmax_peaks(torch.from_numpy(z).to(DEV))
    assert xy.cpu().tolist() == [[[0, 0], [12, 8], [4, 3]]]


# ---------------------------------------------------------------------------------------------- training kernels
@pytest.mark.parametrize("N,H,W,cins,cout", [
    (2, 16, 40, [16], 16),
    (1, 24, 32, [16, 16, 16], 16),
    (2, 16, 16, [32, 32], 32),
    (1, 16, 16, [64], 64),
    (2, 8, 8, [128], 128),
    (1, 8, 8, [64, 64], 64),
    (1, 8, 8, [64], 128),
    (3, 64, 64, [16], 16),
    # tcgen05 path (16/32-channel sources and outputs): ragged tiles, mixed widths, every accumulator split
    (1, 8, 8, [16], 16),
    (2, 40, 72, [16, 16], 16),
    (1, 48, 96, [16, 16, 16, 16, 16], 16),
    (2, 24, 40, [16], 32),
    (1, 32, 64, [32], 32),
    (2, 24, 56, [32, 32, 32], 32),
    (1, 16, 32, [32, 16], 16),
    (5, 128, 128, [16, 16, 16, 16], 16),
    # wide layers: (<= 64-channel chunk, <= 64-channel output slice) jobs, two MMAs per K step for 64-channel chunks
    (1, 24, 40, [32], 64),
    (2, 16, 48, [64, 128], 64),
    (3, 32, 32, [128], 128),
    (2, 16, 16, [64], 32),
    (32, 32, 32, [128], 128),
])
def test_wgrad_conv3x3(N, H, W, cins, cout):
    cin = sum(cins)
    xs = [bf(rnd(N, c, H, W, seed=i + 20)) for i, c in enumerate(cins)]
    dz = bf(rnd(N, cout, H, W, seed=30, scale=0.5))
    ref = torch.nn.grad.conv2d_weight(torch.cat(xs, 1).double(), (cout, cin, 3, 3), dz.double(), padding=1)
    g = ops.wgrad_grid(cins, N, H, W, cout, 9)
    partial = torch.full((g, 9, cin, cout), float("nan"), device=DEV)
    ops.wgrad([nhwc(x) for x in xs], N, H, W, nhwc(dz), cout, 9, partial)
    dst = torch.zeros(cout, cin, 3, 3, device=DEV)
    ops.wgrad_reduce(partial, g, 9, cin, cout, dst, 0, cin, cin * 9, 9, 1)
    close(dst.cpu(), ref, 2e-3, "wgrad")


@pytest.mark.parametrize("N,H,W,cin,cout", [(2, 8, 12, 32, 16), (1, 24, 40, 32, 16), (2, 16, 16, 64, 32), (3, 8, 8, 128, 64), (2, 40, 72, 64, 32),
                                           (32, 32, 32, 128, 64)])
def test_wgrad_deconv_strided_dz(N, H, W, cin, cout):
    x = bf(rnd(N, cin, H, W, seed=31))
    du = bf(rnd(N, cout, 2 * H, 2 * W, seed=32))
    w = torch.zeros(cin, cout, 2, 2, dtype=torch.double, requires_grad=True)
    F.conv_transpose2d(x.double(), w, stride=2).backward(du.double())
    dst = torch.zeros(cin, cout, 2, 2, device=DEV)
    g = ops.wgrad_grid([cin], N, H, W, cout, 1)
    partial = torch.empty(g, 1, cin, cout, device=DEV)
    du_d = nhwc(du)
    for pq in range(4):
        ops.wgrad([nhwc(x)], N, H, W, du_d, cout, 1, partial, dz_view=(pq >> 1, pq & 1))
        ops.wgrad_reduce(partial, g, 1, cin, cout, dst, 0, cin, 4, cout * 4, 0, dst_offset=pq)
    close(dst.cpu(), w.grad, 2e-3, "deconv wgrad")
    # the four taps in one launch, reduced through the deferred (batched) reduction queue
    g4 = ops.wgrad_grid([cin], N, H, W, cout, 1, dz_view="all4")
    partial4 = torch.full((4, g4, 1, cin, cout), float("nan"), device=DEV)
    ops.wgrad([nhwc(x)], N, H, W, du_d, cout, 1, partial4, dz_view="all4")
    dst4 = torch.zeros(cin, cout, 2, 2, device=DEV)
    bias_part = torch.randn(7, 2 * cout, device=DEV)
    bias_out = torch.zeros(cout, device=DEV)
    ops.begin_reduce_queue()
    for pq in range(4):
        ops.wgrad_reduce(partial4, g4, 1, cin, cout, dst4, 0, cin, 4, cout * 4, 0, dst_offset=pq, partial_offset=pq * g4 * cin * cout, defer=True)
    ops.reduce_partials(bias_part, 7, 2 * cout, cout, bias_out, scale=0.5, defer=True)
    assert float(dst4.abs().max()) == 0.0  # nothing ran yet
    ops.flush_reduce_queue({}, DEV)
    assert not ops.reduce_queue_active()
    close(dst4.cpu(), w.grad, 2e-3, "deconv wgrad (4 taps, batched reduce)")
    close(bias_out.cpu(), 0.5 * bias_part[:, :cout].double().sum(0).cpu(), 1e-5, "batched flat reduce")


@pytest.mark.parametrize("cins_consumers,c_t", [([16], 16), ([16, 16, 16], 16), ([32, 32], 32), ([64], 32), ([128], 64)])
def test_dgrad_gather_with_relu_mask_and_stats(cins_consumers, c_t):
    """dT = sum_k conv_transpose(dZ_k, W_k[:, slice]) (+ addend), masked by T > 0, with per-channel sums."""
    N, H, W = 2, 16, 24
    t_act = bf(F.relu(rnd(N, c_t, H, W, seed=40)))
    dzs = [bf(rnd(N, c, H, W, seed=41 + i)) for i, c in enumerate(cins_consumers)]
    ws = [bf(rnd(c, c_t + 16, 3, 3, seed=45 + i, scale=0.1)) for i, c in enumerate(cins_consumers)]  # consumer weights [co, ci_total, 3, 3]
    addend = bf(rnd(N, c_t, H, W, seed=49, scale=0.3))
    n_begin = 16  # T occupies input channels [16, 16 + c_t) of every consumer
    ref = addend.double().clone()
    for dz, w in zip(dzs, ws):
        ref += F.conv_transpose2d(dz.double(), w.double()[:, n_begin:n_begin + c_t], padding=1)
    ref = ref * (t_act > 0)
    ktot = sum(cins_consumers)
    nt = ops.pick_n_tile(c_t, ktot, 9)
    wp = torch.zeros(c_t * 9 * ktot, dtype=torch.bfloat16, device=DEV)
    k8 = 0
    for c, w in zip(cins_consumers, ws):
        ops.pack_weights(w.to(DEV), 1, 9, c_t, nt, c, n_begin=n_begin, dst=wp, k8_total=ktot // 8, k_dst8=k8)
        k8 += c // 8
    out = torch.empty(N, H, W, c_t, dtype=torch.bfloat16, device=DEV)
    g = ops.conv_grid(cins_consumers, N, H, W, c_t, nt, 9)
    stats = torch.full((g, 2, c_t), float("nan"), device=DEV)
    ops.conv([nhwc(d) for d in dzs], N, H, W, wp, c_t, nt, 9, out=out, addend=nhwc(addend), relu_mask_src=nhwc(t_act), stats_partial=stats)
    got = nchw(out)
    close(got, ref, 6e-3, "dgrad gather")
    sums = torch.empty(2 * c_t, device=DEV)
    ops.reduce_partials(stats, g, 2 * c_t, 2 * c_t, sums)
    close(sums[:c_t].cpu(), got.double().sum((0, 2, 3)), 1e-3, "sum v")
    close(sums[c_t:].cpu(), (got.double() ** 2).sum((0, 2, 3)), 1e-3, "sum v^2")


def test_deconv_dgrad_strided_sources():
    N, H, W, cin, cout = 2, 8, 12, 32, 16  # high: [N,cin,H,W]; up: [N,cout,2H,2W]
    du = bf(rnd(N, cout, 2 * H, 2 * W, seed=60))
    w = bf(rnd(cin, cout, 2, 2, seed=61, scale=0.2))
    x = torch.zeros(N, cin, H, W, dtype=torch.double, requires_grad=True)
    F.conv_transpose2d(x, w.double(), stride=2).backward(du.double())
    nt = ops.pick_n_tile(cin, 4 * cout, 1)
    wp = ops.pack_weights(w.to(DEV), 3, 4, cin, nt, cout)  # [nt][tap=pq][k8][n][8]
    # the kernel sees 4 sources (one per tap) and taps=1: K index = pq*cout + co, so repack tap-major into K
    wp = wp.view(cin // nt, 4, cout // 8, nt, 8).reshape(cin // nt, 1, 4 * cout // 8, nt, 8).contiguous()
    du_d = nhwc(du)
    out = torch.empty(N, H, W, cin, dtype=torch.bfloat16, device=DEV)
    ops.conv([du_d] * 4, N, H, W, wp, cin, nt, 1, out=out, strided=[(0, 0), (0, 1), (1, 0), (1, 1)])
    close(nchw(out), x.grad, 6e-3, "deconv dgrad")


def test_bn_train_forward_and_backward_kernels():
    N, H, W, Cc = 2, 16, 24, 32
    z = bf(rnd(N, Cc, H, W, seed=70) * 1.5 + 0.3)
    gamma, beta = torch.rand(Cc) + 0.5, rnd(Cc, seed=71, scale=0.2)
    rm, rv = rnd(Cc, seed=72, scale=0.1), torch.rand(Cc) + 0.5
    dy = bf(rnd(N, Cc, H, W, seed=73))
    zz = z.double().requires_grad_(True)
    g64, b64 = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    rm_ref, rv_ref = rm.double().clone(), rv.double().clone()
    y_ref = F.relu(F.batch_norm(zz, rm_ref, rv_ref, g64, b64, training=True, momentum=0.1, eps=1e-5))
    y_ref.backward(dy.double())
    # statistics partials as the conv epilogue would emit them (two fake CTAs)
    zd = nhwc(z)
    half = z[:1].double(), z[1:].double()
    partial = torch.stack([torch.stack([h.sum((0, 2, 3)), (h * h).sum((0, 2, 3))]) for h in half]).float().to(DEV)
    count = N * H * W
    mean, istd, scale, shift = (torch.empty(Cc, device=DEV) for _ in range(4))
    rm_d, rv_d = rm.to(DEV), rv.to(DEV)
    ops.bn_finalize(partial, 2, Cc, count, gamma.to(DEV), beta.to(DEV), rm_d, rv_d, 0.1, 1e-5, mean, istd, scale, shift)
    close(rm_d.cpu(), rm_ref, 1e-5, "running_mean")
    close(rv_d.cpu(), rv_ref, 1e-5, "running_var")
    y = torch.empty_like(zd)
    pooled = torch.empty(N, H // 2, W // 2, Cc, dtype=torch.bfloat16, device=DEV)
    ops.bn_relu(zd, scale, shift, y, pooled)
    close(nchw(y), y_ref.detach(), 6e-3, "bn relu")
    assert torch.equal(nchw(pooled), F.max_pool2d(nchw(y), 2))
    y2 = torch.empty_like(zd)
    ops.bn_relu(zd, scale, shift, y2, None)
    assert torch.equal(y2, y)
    # backward: dyh = dy * (y > 0); sums = (sum dyh, sum dyh*xhat)
    dyh = dy.double() * (y_ref.detach() > 0)
    xhat = (z.double() - z.double().mean((0, 2, 3), keepdim=True)) / torch.sqrt(z.double().var((0, 2, 3), unbiased=False, keepdim=True) + 1e-5)
    sums = torch.cat([dyh.sum((0, 2, 3)), (dyh * xhat).sum((0, 2, 3))]).float().to(DEV)
    close(sums[:Cc].cpu(), b64.grad, 1e-4, "dbeta")
    close(sums[Cc:].cpu(), g64.grad, 1e-4, "dgamma")
    dz = torch.empty_like(zd)
    ops.bn_bwd_apply(nhwc(dyh.float()), zd, mean, istd, gamma.to(DEV), sums, count, dz)
    close(nchw(dz), zz.grad, 8e-3, "bn dz")


def test_maxpool_backward_first_max_wins():
    x = bf(F.relu(rnd(2, 16, 8, 12, seed=80)))  # many exact ties at 0
    dp = bf(rnd(2, 16, 4, 6, seed=81))
    xx = x.double().requires_grad_(True)
    F.max_pool2d(xx, 2).backward(dp.double())
    dx = torch.empty(2, 8, 12, 16, dtype=torch.bfloat16, device=DEV)
    ops.maxpool_bwd(nhwc(x), nhwc(dp), dx)
    assert torch.equal(nchw(dx).double(), xx.grad)


@pytest.mark.parametrize("mode", ["upstream", "mse", "focal"])
def test_head_backward(mode):
    N, H, W = 2, 16, 24
    x = bf(F.relu(rnd(N, 16, H, W, seed=90)))  # output of a ReLU conv: the kernel also applies that ReLU's mask
    hw, hb = rnd(4, 16, seed=91, scale=0.5), rnd(4, seed=92, scale=0.2)
    mask = (torch.rand(N, 16, H, W, generator=torch.Generator().manual_seed(93)) >= 0.4)
    xx = x.double().requires_grad_(True)
    w64, b64 = hw.double().requires_grad_(True), hb.double().requires_grad_(True)
    heat_ref = torch.sigmoid(F.conv2d(xx * mask.double() / 0.6, w64[:, :, None, None], b64))
    heat = heat_ref.detach().float()
    target = torch.rand(N, 4, H, W, generator=torch.Generator().manual_seed(94))
    coef = 2.0 / (3 * heat.numel())
    if mode == "mse":
        (((heat_ref - target.double()) ** 2).sum() * coef / 2).backward()
        dheat = None
    elif mode == "focal":  # FocalLoss_BCE_2d(gamma=3), mean over 3 heads (oracle restatement pinned to the reference)
        from oracle import unetpp_oracle as O
        coef = 1.0 / (3 * N * 4)
        (O.focal_loss_bce_2d(heat_ref, target.double(), gamma=3.0) / 3).backward()
        dheat = None
    else:
        dheat = rnd(N, 4, H, W, seed=95)
        heat_ref.backward(dheat.double())
    g = ops.head_bwd_grid(N, H, W)
    partial = torch.full((g, 4 * 16 + 4 + 1 + 16), float("nan"), device=DEV)
    dx = torch.empty(N, H, W, 16, dtype=torch.bfloat16, device=DEV)
    m8 = ops.pack_keep_mask(mask.to(DEV))
    ops.head_bwd(heat.to(DEV), None if dheat is None else dheat.to(DEV), target.to(DEV) if mode != "upstream" else None, coef, nhwc(x), m8, 1 / 0.6,
                 hw.to(DEV), dx, partial, loss_kind=1 if mode == "focal" else 0, gamma=3.0)
    dx_ref = xx.grad * (x > 0)
    close(nchw(dx), dx_ref, 6e-3, "head dx")
    red = torch.empty(85, device=DEV)
    ops.reduce_partials(partial, g, 85, 85, red)
    close(red[69:].cpu(), nchw(dx).double().sum((0, 2, 3)), 1e-4, "dx channel sums")
    close(red[:64].cpu().view(4, 16), w64.grad, 2e-3, "head dW")
    close(red[64:68].cpu(), b64.grad, 2e-3, "head db")
    if mode == "focal":
        a = (heat.double() - target.double()).abs()
        close(red[68:69].cpu(), (-(a ** 3) * torch.log(1 - a + 1e-20)).sum().reshape(1), 1e-4, "focal loss sum")
    if mode == "mse":
        close(red[68:69].cpu(), ((heat.double() - target.double()) ** 2).sum().reshape(1), 1e-4, "loss sum")


def test_adamw_matches_reference_semantics(golden):
    arr, meta = golden
    h = meta["adamw_hyper"]
    for j in range(2):
        p = torch.from_numpy(arr[f"adamw_p0_{j}"]).reshape(-1).to(DEV)
        m, v = torch.zeros_like(p), torch.zeros_like(p)
        for s in range(3):
            g = torch.from_numpy(arr[f"adamw_g{s}_{j}"]).reshape(-1).to(DEV)
            ops.adamw(p, g, m, v, h["lr"], h["betas"][0], h["betas"][1], h["eps"], h["weight_decay"], s + 1)
        assert np.allclose(p.cpu().numpy(), arr[f"adamw_p3_{j}"].reshape(-1), rtol=2e-6, atol=1e-7)


def test_dropout_mask_rate_and_determinism():
    m1 = torch.empty(4, 128, 128, dtype=torch.int16, device=DEV)  # one 16-bit keep word per pixel
    m2 = torch.empty_like(m1)
    ops.dropout_mask(m1, 0.4, 1234)
    ops.dropout_mask(m2, 0.4, 1234)
    assert torch.equal(m1, m2)
    keep = ops.unpack_keep_mask(m1)  # [N,16,H,W] bool
    assert abs(float(keep.float().mean()) - 0.6) < 5e-3
    per_channel = keep.float().mean((0, 2, 3))
    assert float((per_channel - 0.6).abs().max()) < 2e-2  # every bit position is used and unbiased
    assert torch.equal(ops.pack_keep_mask(keep), m1)
    ops.dropout_mask(m2, 0.4, 99)
    assert not torch.equal(m1, m2)


def test_bad_arguments_return_errors_not_crashes():
    from unet_nested4tiny_objects_keypoints_b200._lib import UnppError
    x = torch.zeros(1, 8, 8, 24, dtype=torch.bfloat16, device=DEV)  # 24 channels: unsupported
    wp = torch.zeros(16 * 9 * 24, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(UnppError, match="16/32/64/128"):
        ops.conv([x], 1, 8, 8, wp, 16, 16, 9, out=torch.empty(1, 8, 8, 16, dtype=torch.bfloat16, device=DEV))
    with pytest.raises(UnppError):
        ops.maxpool(torch.zeros(1, 7, 8, 16, dtype=torch.bfloat16, device=DEV), torch.zeros(1, 3, 4, 16, dtype=torch.bfloat16, device=DEV))


# ---------------------------------------------------------------------------------------------- 2x2 output-blocked conv (16-channel levels)
@pytest.mark.parametrize("N,H,W,nsrc", [(2, 32, 32, 1), (1, 24, 40, 2), (3, 64, 64, 3), (1, 8, 8, 4), (2, 16, 48, 4), (1, 128, 96, 1)])
def test_conv3x3_block2x2_forward(N, H, W, nsrc):
    cin = 16 * nsrc
    xs = [bf(rnd(N, 16, H, W, seed=i + 101)) for i in range(nsrc)]
    w = bf(rnd(16, cin, 3, 3, seed=150, scale=(2.0 / (9 * cin)) ** 0.5))
    b = rnd(16, seed=151, scale=0.1)
    ref = F.relu(F.conv2d(torch.cat(xs, 1).double(), w.double(), b.double(), padding=1))
    wp = ops.pack_weights_b2(w.to(DEV), False, cin)
    out = torch.empty(N, H, W, 16, dtype=torch.bfloat16, device=DEV)
    ops.conv([nhwc(x) for x in xs], N, H, W, wp, 16, 16, 9, bias=b.to(DEV), relu=True, out=out, b2=True)
    torch.cuda.synchronize()
    close(nchw(out), ref, 6e-3, "conv3x3 2x2-blocked")


def test_block2x2_fused_head_and_bn_fold():
    N, H, W = 2, 32, 48
    x = bf(rnd(N, 16, H, W, seed=160))
    w = rnd(16, 16, 3, 3, seed=161, scale=0.12)
    scale = torch.rand(16) + 0.5
    b = rnd(16, seed=162, scale=0.1)
    hw, hb = rnd(4, 16, seed=163, scale=0.5), rnd(4, seed=164, scale=0.2)
    y = F.relu(F.conv2d(x.double(), bf(w * scale[:, None, None, None]).double(), b.double(), padding=1))
    logit = F.conv2d(y, hw.double()[:, :, None, None], hb.double())
    wp = ops.pack_weights_b2(w.to(DEV), False, 16, scale=scale.to(DEV))
    out = torch.empty(N, H, W, 16, dtype=torch.bfloat16, device=DEV)
    heat = torch.empty(N, 4, H, W, device=DEV)
    ops.conv([nhwc(x)], N, H, W, wp, 16, 16, 9, bias=b.to(DEV), relu=True, out=out, head=(hw.to(DEV), hb.to(DEV), heat, None, None, 1.0), b2=True)
    close(nchw(out), y, 6e-3, "b2 head conv out")
    close(heat.cpu(), torch.sigmoid(logit), 2e-3, "b2 heat")


@pytest.mark.parametrize("ncons", [1, 3])
def test_block2x2_dgrad_gather_mask_stats(ncons):
    N, H, W = 2, 32, 40
    t_act = bf(F.relu(rnd(N, 16, H, W, seed=170)))
    dzs = [bf(rnd(N, 16, H, W, seed=171 + i)) for i in range(ncons)]
    ws = [bf(rnd(16, 48, 3, 3, seed=175 + i, scale=0.1)) for i in range(ncons)]
    addend = bf(rnd(N, 16, H, W, seed=179, scale=0.3))
    aux = bf(rnd(N, 16, H, W, seed=180))
    mean, istd = rnd(16, seed=181, scale=0.2), torch.rand(16) + 0.5
    n_begin = 16
    ref = addend.double().clone()
    for dz, w in zip(dzs, ws):
        ref += F.conv_transpose2d(dz.double(), w.double()[:, n_begin:n_begin + 16], padding=1)
    ref = ref * (t_act > 0)
    ktot = 16 * ncons
    wp = torch.zeros(64 * 16 * ktot, dtype=torch.bfloat16, device=DEV)
    for i, w in enumerate(ws):
        ops.pack_weights_b2(w.to(DEV), True, 16, n_begin=n_begin, dst=wp, k8_total=ktot // 8, k_dst8=2 * i)
    out = torch.empty(N, H, W, 16, dtype=torch.bfloat16, device=DEV)
    g = ops.conv_grid([16] * ncons, N, H, W, 16, 16, 9, b2=True)
    stats = torch.full((g, 2, 16), float("nan"), device=DEV)
    ops.conv([nhwc(d) for d in dzs], N, H, W, wp, 16, 16, 9, out=out, addend=nhwc(addend), relu_mask_src=nhwc(t_act), stats_partial=stats,
             stats_aux=nhwc(aux), aux_mean=mean.to(DEV), aux_istd=istd.to(DEV), b2=True)
    got = nchw(out)
    close(got, ref, 6e-3, "b2 dgrad gather")
    sums = torch.empty(32, device=DEV)
    ops.reduce_partials(stats, g, 32, 32, sums)
    close(sums[:16].cpu(), got.double().sum((0, 2, 3)), 1e-3, "sum v")
    xhat = (aux.double() - mean.double()[None, :, None, None]) * istd.double()[None, :, None, None]
    close(sums[16:].cpu(), (got.double() * xhat).sum((0, 2, 3)), 2e-3, "sum v*xhat")


@pytest.mark.parametrize("N,H,W,cin,cout,b2", [(2, 64, 96, 16, 16, True), (1, 24, 40, 16, 16, True), (2, 32, 48, 32, 32, False), (1, 16, 24, 64, 64, False),
                                                (3, 8, 8, 16, 32, False), (1, 40, 72, 16, 16, False)])
def test_conv_with_fused_maxpool(N, H, W, cin, cout, b2):
    """conv3x3 + bias + ReLU with nn.MaxPool2d(2) of the result written by the same epilogue (unet.py:258-262): the pooled
    tensor equals the pool of the stored bf16 output bit for bit."""
    x = bf(rnd(N, cin, H, W, seed=300))
    w = bf(rnd(cout, cin, 3, 3, seed=301, scale=(2.0 / (9 * cin)) ** 0.5))
    b = rnd(cout, seed=302, scale=0.1)
    if b2:
        wp, nt = ops.pack_weights_b2(w.to(DEV), False, cin), ops.NTile(16, b2=True)
    else:
        nt = ops.pick_n_tile(cout, cin, 9)
        wp = ops.pack_weights(w.to(DEV), 0, 9, cout, nt, cin)
    out = torch.empty(N, H, W, cout, dtype=torch.bfloat16, device=DEV)
    pooled = torch.full((N, H // 2, W // 2, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.conv([nhwc(x)], N, H, W, wp, cout, nt, 9, bias=b.to(DEV), relu=True, out=out, pooled=pooled)
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(x.double(), w.double(), b.double(), padding=1))
    close(nchw(out), ref, 6e-3, "conv + relu")
    assert torch.equal(nchw(pooled), F.max_pool2d(nchw(out), 2))


def compose_deconv_conv_fp64(w_conv_u: torch.Tensor, w_up: torch.Tensor, b_up: torch.Tensor, b_conv: torch.Tensor):
    """conv3x3(ConvTranspose2d_k2s2(x)) as ONE 3x3 conv over the low-resolution x (the k2s2 upsample never overlaps).

    w_conv_u [Co, Cu, 3, 3]: the slice of the consuming conv's weight that multiplies the upsampled tensor (unet.py:199-201:
    the first Cu input channels); w_up [Ci, Cu, 2, 2], b_up [Cu]: the transposed conv (unet.py:187); b_conv [Co].
    Returns (composed weight [4*Co, Ci, 3, 3] whose output channel is (2*jy+jx)*Co + co for pixel (jy, jx) of the 2x2 output
    block, and whose taps are offsets -1..1 on the low-res grid; bias table [9, Co] indexed by 3*rowclass + colclass,
    class 0 = first row/column, 1 = interior, 2 = last: zero padding is applied AFTER the upsample, so border pixels see
    fewer taps of the upsample bias)."""
    co, cu = w_conv_u.shape[0], w_conv_u.shape[1]
    ci = w_up.shape[0]
    wc, wd = w_conv_u.double(), w_up.double()
    comp = torch.zeros(4 * co, ci, 3, 3, dtype=torch.float64, device=wc.device)
    for jy in range(2):
        for jx in range(2):
            for r in range(3):
                for s in range(3):
                    uy, ux = jy + r - 1, jx + s - 1          # position in the upsampled grid relative to the block origin
                    dyl, p_ = uy // 2, uy % 2                # low-res row offset (-1, 0, 1) and row parity of that upsampled pixel
                    dxl, q_ = ux // 2, ux % 2
                    blk = (2 * jy + jx) * co
                    comp[blk:blk + co, :, dyl + 1, dxl + 1] += wc[:, :, r, s] @ wd[:, :, p_, q_].t()
    t = torch.einsum("ocrs,c->rso", wc, b_up.double())       # contribution of the upsample bias through tap (r, s)
    valid = {0: (1, 2), 1: (0, 1, 2), 2: (0, 1)}              # taps that stay inside the image for first / interior / last row
    table = torch.zeros(9, co, dtype=torch.float64, device=wc.device)
    for rc in range(3):
        for cc in range(3):
            table[3 * rc + cc] = b_conv.double() + sum(t[r, s] for r in valid[rc] for s in valid[cc])
    return comp.float().contiguous(), table.float().contiguous()


def test_compose_kernel_matches_the_fp64_composition():
    """unpp_compose_deconv_conv (fp32, one thread per element) against the fp64 torch restatement above."""
    w = rnd(16, 48, 3, 3, seed=301, scale=0.1)
    w_up, b_up, b = rnd(32, 16, 2, 2, seed=302, scale=0.2), rnd(16, seed=303, scale=0.3), rnd(16, seed=304, scale=0.1)
    rc, rt = compose_deconv_conv_fp64(w[:, :16], w_up, b_up, b)
    comp, table = ops.compose_deconv_conv(w.to(DEV), 16, w_up.to(DEV), b_up.to(DEV), b.to(DEV))
    assert comp.shape == rc.shape and table.shape == rt.shape
    assert float((comp.cpu() - rc).abs().max()) <= 1e-6 * float(rc.abs().max()) + 1e-7
    assert float((table.cpu() - rt).abs().max()) <= 1e-6 * float(rt.abs().max()) + 1e-7


@pytest.mark.parametrize("N,H,W,nlows", [(2, 32, 32, 1), (1, 24, 40, 2), (2, 64, 64, 3), (1, 8, 8, 1), (1, 128, 64, 2)])
def test_fused_transposed_conv_into_block2x2_conv(N, H, W, nlows):
    """conv3x3(cat[ConvTranspose2d_k2s2(x_low), lows...]) + bias + ReLU in ONE launch, the upsampled tensor never exists."""
    xlow = bf(rnd(N, 32, H // 2, W // 2, seed=200))
    lows = [bf(rnd(N, 16, H, W, seed=201 + i)) for i in range(nlows)]
    w_up = rnd(32, 16, 2, 2, seed=210, scale=0.2)
    b_up = rnd(16, seed=211, scale=0.3)
    cin = 16 * (1 + nlows)
    w = rnd(16, cin, 3, 3, seed=212, scale=(2.0 / (9 * cin)) ** 0.5)
    b = rnd(16, seed=213, scale=0.1)
    up = F.conv_transpose2d(xlow.double(), w_up.double(), b_up.double(), stride=2)
    ref = F.relu(F.conv2d(torch.cat([up] + [l.double() for l in lows], 1), w.double(), b.double(), padding=1))
    comp, table = ops.compose_deconv_conv(w.to(DEV), 16, w_up.to(DEV), b_up.to(DEV), b.to(DEV))
    wp = ops.pack_weights_b2(w.to(DEV), False, 16 * nlows, k_begin=16)
    lw = ops.pack_weights(comp, 6, 9, 64, 64, 32)  # kind 6: only the non-zero (tap, pixel) blocks (csrc/b2_blocks.h)
    out = torch.empty(N, H, W, 16, dtype=torch.bfloat16, device=DEV)
    ops.conv([nhwc(l) for l in lows], N, H, W, wp, 16, ops.NTile(16, b2=True), 9, bias=table, bias_classes=9, relu=True, out=out,
             lowres=(nhwc(xlow), lw))
    torch.cuda.synchronize()
    close(nchw(out), ref, 8e-3, "fused transposed conv")


def test_target_heatmap_synthesis_matches_reference(golden):
    """tools/misc/helper.py:87-172 on the device: against the reference output in the fixture and the oracle on new points."""
    from oracle import unetpp_oracle as O
    arr, _ = golden
    got = ops.create_heatmap(torch.from_numpy(arr["hm_keypoints"]).to(DEV), 48, 48)
    assert np.allclose(got.cpu().numpy(), arr["hm_target"], rtol=0, atol=2e-6)
    kp = (torch.rand(3, 7, 2, generator=torch.Generator().manual_seed(5)) * torch.tensor([70.0, 38.0]) + 1).float()
    got = ops.create_heatmap(kp.to(DEV), 40, 72).cpu().numpy()
    ref = O.create_heatmap(kp.numpy(), 40, 72)
    assert got.shape == ref.shape == (3, 4, 40, 72) and np.allclose(got, ref, rtol=0, atol=2e-6)
    assert np.allclose(got[:, [1, 3]].max(axis=(2, 3)), 1.0)  # multi-point planes are max-normalised
