"""How often does bf16 storage move a key point?  (arg-max is bit-exact on identical heat maps, but the product's purpose
is the key points, and bf16 heat maps differ from fp32 ones by up to ~1e-2.)

A UNet_Nested is trained for a few hundred fused steps on a synthetic task whose targets are the reference's own heat maps
(helper.create_heatmap of 7 random key points; the image shows the same blobs under noise: channel 0 = plane 0 minus plane 2,
channel 1 = plane 1, channel 2 = plane 3 — every plane can be recovered, so a short training run suffices),
so that its outputs are peaky like a trained model's.  Fresh images then go through (a) the tensor-core bf16 path and (b) the fp32
validation mode (csrc/ref_kernels.cu, ~1e-6 of the reference's fp32 arithmetic); key points are extracted from both exactly as
the reference does (arg-max for the single-point planes 0 and 2, heatmap.py:173-178; the brightest 3 / 2 regions for planes 1 and 3,
heatmap.py:148-208) and compared.  Writes a markdown report.   usage: python scripts/keypoint_agreement.py out.md [steps] [side]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unet_nested4tiny_objects_keypoints_b200 as pkg  # noqa: E402
from unet_nested4tiny_objects_keypoints_b200 import fused, ops  # noqa: E402


def batch(B, S, gen):
    kp = (torch.rand(B, 7, 2, generator=gen, device="cuda") * (S - 24) + 12).float()
    t = ops.create_heatmap(kp, S, S)
    x = torch.stack([t[:, 0] - t[:, 2], t[:, 1], t[:, 3]], 1)
    x = x + 0.05 * torch.randn(x.shape, generator=gen, device="cuda")
    return x.contiguous(), t, kp


def main():
    out = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
    S = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    B = 32
    torch.manual_seed(0)
    gen = torch.Generator(device="cuda").manual_seed(1)
    m = pkg.UNet_Nested().cuda().train()
    loss_kind = sys.argv[4] if len(sys.argv) > 4 else "mse"
    st = fused.FusedTrainStep(m, B, S, S, lr=1e-3, weight_decay=0.0, loss=loss_kind, seed=3)
    losses = []
    for k in range(steps):
        x, t, _ = batch(B, S, gen)
        st.x.copy_(x)
        st.target.copy_(t)
        st.step_device()
        if k == steps // 2:
            st.set_lr(2e-4)  # one MultiStepLR-style drop (trainer.py:383-388)
        if k % 250 == 0 or k == steps - 1:
            losses.append((k, float(st.loss)))
    m.eval()
    ident = {c: 0 for c in range(4)}
    within1 = {c: 0 for c in range(4)}
    total = {c: 0 for c in range(4)}
    maxd = {c: 0.0 for c in range(4)}
    gt_err, gt_err16, heat_err, peak = [], [], [], []
    nums = {0: 1, 1: 3, 2: 1, 3: 2}
    for _ in range(8):
        x, t, kp = batch(B, S, gen)
        with torch.no_grad():
            m.precision = "bf16"
            hb = m(x)[2]
            m.precision = "fp32"
            hf = m(x)[2]
        m.precision = "bf16"
        heat_err.append(float((hb - hf).abs().max()))
        peak.append(float(hf.amax((2, 3)).mean()))
        for c in range(4):
            n = nums[c]
            if n == 1:
                xb, _ = ops.argmax_peaks(hb[:, c:c + 1].contiguous())
                xf, _ = ops.argmax_peaks(hf[:, c:c + 1].contiguous())
                xb, xf = xb.view(B, 1, 2).float(), xf.view(B, 1, 2).float()
            else:
                xb, _, _ = ops.topk_peaks(hb[:, c:c + 1].contiguous(), n)
                xf, _, _ = ops.topk_peaks(hf[:, c:c + 1].contiguous(), n)
                xb, xf = xb.view(B, n, 2).float(), xf.view(B, n, 2).float()
            d = torch.cdist(xb, xf).min(2).values  # every bf16 point against its nearest fp32 point (order may swap between equally bright blobs)
            ident[c] += int((d == 0).sum())
            within1[c] += int((d <= 1.5).sum())
            total[c] += d.numel()
            maxd[c] = max(maxd[c], float(d.max()))
        for c, pt in ((0, 0), (2, 4)):  # the single-point planes: distance of the arg-max from the drawn point
            a0, _ = ops.argmax_peaks(hf[:, c:c + 1].contiguous())
            b0, _ = ops.argmax_peaks(hb[:, c:c + 1].contiguous())
            gt_err.append(float((a0.view(B, 2).float() - kp[:, pt]).norm(dim=1).mean()))
            gt_err16.append(float((b0.view(B, 2).float() - kp[:, pt]).norm(dim=1).mean()))
    with open(out, "w") as f:
        f.write(f"# Key-point agreement of the bf16 tensor-core path with the fp32 validation mode ({S}x{S}, {steps} fused training steps at batch {B})\n\n")
        f.write(__doc__.split("usage")[0].strip() + "\n\n")
        f.write(f"training loss ({loss_kind}, mean of three heads): " + ", ".join(f"step {k}: {v:.5f}" for k, v in losses) + "\n\n")
        f.write(f"mean distance of the key points of the single-point planes 0 and 2 from the drawn (sub-pixel) points: fp32 mode {sum(gt_err) / len(gt_err):.2f} px, bf16 path {sum(gt_err16) / len(gt_err16):.2f} px\n\n")
        f.write(f"max |heat_bf16 - heat_fp32| over the evaluation batches: {max(heat_err):.3e}; mean plane maximum of the fp32 heat maps: {sum(peak) / len(peak):.3f}\n\n")
        f.write("| plane | points per plane | key points compared | identical | within 1 px (incl. diagonal) | largest distance (px) |\n|---|---|---|---|---|---|\n")
        for c in range(4):
            f.write(f"| {c} | {nums[c]} | {total[c]} | {100.0 * ident[c] / total[c]:.2f} % | {100.0 * within1[c] / total[c]:.2f} % | {maxd[c]:.1f} |\n")
    print(open(out).read())


if __name__ == "__main__":
    main()
