"""Pins the oracle (oracle/unetpp_oracle.py) against the REAL reference and writes the golden
fixtures under tests/golden/.  TEST INFRASTRUCTURE ONLY — run in the build container, where the
reference is mounted read-only at /root/reference:

    python oracle/make_golden.py

For every function of the oracle the script (1) runs the reference's own code (models/unet.py
UNet_Nested, tools/optimizers/adamw.py AdamW, tools/misc/helper.py create_heatmap,
tools/misc/heatmap.py Heatmap.extract_points_/create_heatmap, tools/losses/focal_loss.py
FocalLoss_BCE_2d, torch.nn.MSELoss as used at trainer/trainer.py:427) on seeded inputs, (2) asserts
that the oracle restatement agrees, and (3) stores the REFERENCE outputs (or digests of them, for
the 2.2 MB gradient set) so that tests/test_oracle_golden.py can replay the check on machines
where /root/reference does not exist.
"""
from __future__ import annotations

import hashlib
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

GOLD = os.path.join(ROOT, "tests", "golden")


def digest(t: torch.Tensor):
    t = t.detach().double().reshape(-1)
    return [float(t.sum()), float(t.abs().sum()), float((t * t).sum())] + [float(v) for v in t[:6]]


class _MaskDropout(torch.nn.Module):
    """Stands in for nn.Dropout(p=0.4) (unet.py:254) with externally supplied keep-masks, consumed in
    call order (final_1, final_2, final_3 — unet.py:283-286)."""

    def __init__(self, masks, p=0.4):
        super().__init__()
        self.masks, self.p, self.i = masks, p, 0

    def forward(self, x):
        m = self.masks[self.i % len(self.masks)]
        self.i += 1
        return x * m.to(x.dtype) / (1.0 - self.p)


def main():
    torch.set_num_threads(4)
    torch.use_deterministic_algorithms(False)
    from models.unet import UNet_Nested as RefNet  # the real reference
    meta = {}
    arrays = {}

    # ---- 1. constructor: key order, shapes, dtypes, and the seeded init stream (unet.py:206-254)
    torch.manual_seed(0)
    ref = RefNet()
    sd = ref.state_dict()
    spec = O.state_dict_spec()
    assert list(sd.keys()) == list(spec.keys()), "state_dict key order differs"
    for k, v in sd.items():
        assert tuple(v.shape) == spec[k][0] and v.dtype == spec[k][1], k
    h = hashlib.sha256()
    nbytes = 0
    for k, v in sd.items():
        b = v.detach().cpu().contiguous().numpy().tobytes()
        h.update(k.encode() + b)
        nbytes += len(b)
    meta["init_seed0_sha256"] = h.hexdigest()
    meta["state_dict_bytes"] = nbytes
    meta["state_dict_keys"] = [[k, list(v.shape), str(v.dtype)] for k, v in sd.items()]
    meta["n_params"] = sum(p.numel() for p in ref.parameters())
    assert nbytes == 2216944 and meta["n_params"] == 553260

    # ---- 2. eval forward (unet.py:255-300) with synthetic weights
    wsd = O.synth_state_dict(seed=1)
    ref.load_state_dict(wsd)
    ref.eval()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 3, 32, 32, generator=g)
    with torch.no_grad():
        r_out = ref(x)
    inter = {}
    o_out = O.forward(wsd, x, training=False, inter=inter)
    assert isinstance(r_out, tuple) and len(r_out) == 3
    for a, b in zip(r_out, o_out):
        assert torch.allclose(a, b, rtol=0, atol=2e-6), float((a - b).abs().max())
    arrays["eval_x"] = x.numpy()
    for i, a in enumerate(r_out):
        arrays[f"eval_out{i}"] = a.numpy()
    # non-square eval case (H != W, both divisible by 8)
    x2 = torch.randn(1, 3, 16, 40, generator=g)
    with torch.no_grad():
        r2 = ref(x2)
    o2 = O.forward(wsd, x2)
    for a, b in zip(r2, o2):
        assert torch.allclose(a, b, rtol=0, atol=2e-6)
    arrays["eval2_x"] = x2.numpy()
    arrays["eval2_out2"] = r2[2].numpy()

    # ---- 3. training step: BN batch stats, masked dropout, MSE (trainer.py:125-135,427), grads
    ref.load_state_dict(wsd)
    ref.train()
    xt = torch.randn(3, 3, 32, 32, generator=g)
    target = torch.rand(3, 4, 32, 32, generator=g)
    masks = [(torch.rand(3, 16, 32, 32, generator=g) >= 0.4).to(torch.uint8) for _ in range(3)]
    ref.drop_out = _MaskDropout(masks)
    ref.zero_grad()
    outs = ref(xt)
    crit = torch.nn.MSELoss()
    loss = sum(crit(o, target) for o in outs) / len(outs)  # trainer.py:125-134 (mean of the 3 head losses)
    loss.backward()
    o_loss, o_outs, o_grads, o_stats = O.train_step_grads(wsd, xt, target, dropout_masks=masks, loss="mse")
    assert abs(float(loss.detach()) - float(o_loss)) < 1e-7
    ref_grads = {k: p.grad for k, p in ref.named_parameters()}
    assert set(ref_grads) == set(o_grads)
    for k in ref_grads:
        a, b = ref_grads[k], o_grads[k]
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-7 + 1e-5 * float(a.abs().max())), (k, float((a - b).abs().max()))
    new_sd = ref.state_dict()
    for k, v in o_stats.items():
        assert torch.allclose(new_sd[k].to(v.dtype), v, rtol=1e-5, atol=1e-6), k
    arrays["train_x"] = xt.numpy()
    arrays["train_target"] = target.numpy()
    for i, m in enumerate(masks):
        arrays[f"train_mask{i}"] = np.packbits(m.numpy().reshape(-1))
    for i, o in enumerate(outs):
        arrays[f"train_out{i}"] = o.detach().numpy()
    meta["train_loss"] = float(loss.detach())
    meta["train_grad_digest"] = {k: digest(v) for k, v in ref_grads.items()}
    meta["train_new_stats_digest"] = {k: digest(new_sd[k].float()) for k in o_stats}
    # two small full gradients for element-wise replay
    arrays["train_grad_final_3.weight"] = ref_grads["final_3.weight"].numpy()
    arrays["train_grad_conv00.conv1.0.weight"] = ref_grads["conv00.conv1.0.weight"].numpy()
    arrays["train_grad_up_concat01.up.weight"] = ref_grads["up_concat01.up.weight"].numpy()

    # ---- 3b. the criterion the trainer ships: FocalLoss_BCE_2d (focal_loss.py:255-301, trainer.py:426)
    from tools.losses.focal_loss import FocalLoss_BCE_2d
    fl = FocalLoss_BCE_2d(gamma=3, size_average=False)
    p_ = torch.rand(2, 4, 8, 8, generator=g)
    t_ = torch.rand(2, 4, 8, 8, generator=g)
    r_fl = fl(p_, t_)
    o_fl = O.focal_loss_bce_2d(p_, t_, gamma=3.0)
    assert abs(float(r_fl) - float(o_fl)) < 1e-5 * max(1.0, abs(float(r_fl)))
    arrays["focal_p"], arrays["focal_t"] = p_.numpy(), t_.numpy()
    meta["focal_loss"] = float(r_fl)

    # ---- 4. AdamW (tools/optimizers/adamw.py:38-100): 3 steps on two tensors
    from tools.optimizers.adamw import AdamW as RefAdamW
    p0 = [torch.randn(7, 5, generator=g), torch.randn(11, generator=g)]
    grads = [[torch.randn(7, 5, generator=g), torch.randn(11, generator=g)] for _ in range(3)]
    params = [torch.nn.Parameter(t.clone()) for t in p0]
    hyper = dict(lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        opt = RefAdamW(params, **hyper)
        for step_g in grads:
            for p, gr in zip(params, step_g):
                p.grad = gr.clone()
            opt.step()
    for j in range(2):
        p, m, v = p0[j].clone(), torch.zeros_like(p0[j]), torch.zeros_like(p0[j])
        for s, step_g in enumerate(grads, 1):
            p, m, v = O.adamw_reference_step(p, step_g[j], m, v, s, **hyper)
        assert torch.allclose(p, params[j].detach(), rtol=1e-6, atol=1e-7), j
        arrays[f"adamw_p0_{j}"] = p0[j].numpy()
        arrays[f"adamw_p3_{j}"] = params[j].detach().numpy()
        for s in range(3):
            arrays[f"adamw_g{s}_{j}"] = grads[s][j].numpy()
    meta["adamw_hyper"] = dict(lr=1e-2, betas=[0.9, 0.999], eps=1e-8, weight_decay=1e-2)

    # ---- 5. target synthesis (helper.py:87-172) and peak extraction (heatmap.py:148-208)
    from tools.misc import helper as RH
    from tools.misc.heatmap import Heatmap
    kp = (torch.rand(2, 7, 2, generator=g) * 40 + 4).numpy().astype(np.float32)
    r_hm = RH.create_heatmap(kp, 48, 48)
    o_hm = O.create_heatmap(kp, 48, 48)
    assert r_hm.dtype == np.float32 and np.allclose(r_hm, o_hm, rtol=0, atol=1e-6)
    arrays["hm_keypoints"] = kp
    arrays["hm_target"] = r_hm
    hmaper = Heatmap([[0], [1, 2, 3], [4], [5, 6]], 48, 48)
    # single-blob planes: the watershed keeps one region, so extract_points_ reduces to its arg-max core
    planes = r_hm[:, [0, 2]].copy()  # channels 0 and 2 hold exactly one point each
    ref_pts = np.zeros((2, 2, 2), dtype=np.int32)
    for b in range(2):
        for c in range(2):
            pts = hmaper.extract_points_(planes[b, c], 1)
            assert len(pts) == 1, pts
            ref_pts[b, c] = pts[0]
    o_xy, o_val = O.argmax_keypoints(planes)
    assert np.array_equal(ref_pts, o_xy), (ref_pts, o_xy)
    arrays["peaks_planes"] = planes
    arrays["peaks_xy"] = ref_pts
    # tie-breaking: two equal maxima inside one blob (first in row-major order wins, heatmap.py:173-176)
    tie = planes[0:1, 0:1].copy()
    y0, x0 = int(ref_pts[0, 0, 1]), int(ref_pts[0, 0, 0])
    tie[0, 0, y0 + 1, x0 - 1] = 2.0
    tie[0, 0, y0, x0 + 1] = 2.0
    pts = hmaper.extract_points_(tie[0, 0], 1)
    o_tie, _ = O.argmax_keypoints(tie)
    assert pts[0] == [x0 + 1, y0] and pts[0] == list(o_tie[0, 0]), (pts, o_tie)
    arrays["tie_plane"], arrays["tie_xy"] = tie, np.asarray(pts[0], dtype=np.int32)

    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, "unetpp_golden.npz"), **arrays)
    meta["torch_version"] = torch.__version__
    meta["generator"] = "oracle/make_golden.py (reference imported from /root/reference)"
    with open(os.path.join(GOLD, "unetpp_golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("golden fixtures written:", {k: v.shape for k, v in arrays.items() if v.size > 64})
    print("oracle pinned against the reference: OK")


if __name__ == "__main__":
    main()
