"""Per-parameter gradient error table of one training step (all 74 tensors), written as markdown.

Columns (err = max|g - g_ref| / max|g_ref|, cos = cosine):
  ours vs fp32      the CUDA step against the fp32 oracle (torch autograd, TF32 off) — the bf16 storage noise of this graph
  autocast vs fp32  stock PyTorch autocast(bf16) (cuDNN) against the same oracle — the yardstick for that noise
  emu vs fp32       oracle/bf16_emulation.py (fp32 arithmetic + our storage points) against the oracle: the noise an EXACT
                    implementation of our storage scheme has
  ours vs emu64     the CUDA step against the emulation run in fp64
  emu32 vs emu64    two exact emulations that differ only in accumulation precision: the chaos floor of an end-to-end comparison
usage: python scripts/grad_error_table.py B S out.md
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unet_nested4tiny_objects_keypoints_b200 as pkg  # noqa: E402
from oracle import bf16_emulation as E  # noqa: E402
from oracle import unetpp_oracle as O  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def err_cos(g, r):
    g, r = g.double().cuda(), r.double().cuda()
    return float((g - r).abs().max()) / (float(r.abs().max()) + 1e-30), float((g * r).sum() / (g.norm() * r.norm() + 1e-300))


def main():
    B, S, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    sd = O.synth_state_dict(seed=42)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, 3, S, S, generator=g)
    target = torch.rand(B, 4, S, S, generator=g)
    m = pkg.UNet_Nested()
    m.load_state_dict(sd)
    m = m.cuda().train()
    m.drop_out.p = 0.0
    outs = m(x.cuda())
    (sum(F.mse_loss(o, target.cuda()) for o in outs) / 3).backward()
    ours = {k: p.grad.clone() for k, p in m.named_parameters()}
    del m, outs
    sdc = {k: v.cuda() for k, v in sd.items()}
    xc, tc = x.cuda(), target.cuda()
    _, _, ref, _ = O.train_step_grads(sdc, xc, tc)
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sdc.items() if v.dtype.is_floating_point and "running_" not in k}
    full = dict(sdc)
    full.update(params)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ao = O.forward(full, xc, training=True, dropout_masks=None)
    (sum(F.mse_loss(o.float(), tc) for o in ao) / 3).backward()
    auto = {k: p.grad for k, p in params.items()}
    del ao
    _, _, e32, _ = E.train_step_grads_bf16(sdc, xc, tc)
    sd64 = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in sdc.items()}
    _, _, e64, _ = E.train_step_grads_bf16(sd64, xc.double(), tc.double())
    rows = []
    for k in ours:
        if k.startswith("conv") and k.endswith(".0.bias"):
            continue  # bias in front of BatchNorm: analytically zero (exact zero here, rounding noise in autograd)
        rows.append((k, float(ref[k].abs().max()), err_cos(ours[k], ref[k]), err_cos(auto[k], ref[k]), err_cos(e32[k], ref[k]), err_cos(ours[k], e64[k]),
                     err_cos(e32[k], e64[k])))
    with open(out, "w") as f:
        f.write(f"# Parameter-gradient errors of one training step, B={B}, {S}x{S} (seed 42 weights, seed 7 data, MSE, dropout off)\n\n")
        f.write(__doc__.split("usage")[0].strip() + "\n\n")
        f.write("| parameter | max\\|g_ref\\| | ours vs fp32 err | cos | autocast vs fp32 err | emu vs fp32 err | ours vs emu64 err | cos | emu32 vs emu64 err |\n|---|---|---|---|---|---|---|---|---|\n")
        for k, s, a, b, c, d, e in rows:
            f.write(f"| {k} | {s:.2e} | {a[0]:.2e} | {a[1]:.5f} | {b[0]:.2e} | {c[0]:.2e} | {d[0]:.2e} | {d[1]:.5f} | {e[0]:.2e} |\n")

        def grp(k):
            return "encoder" if k.startswith("conv") else "deep decoder" if k.startswith(("up_concat11", "up_concat12", "up_concat21")) else "full-res decoder + heads"
        f.write("\n| group | worst ours vs fp32 err | min cos | worst autocast err | worst emu vs fp32 err | worst ours vs emu64 | worst emu32 vs emu64 |\n|---|---|---|---|---|---|---|\n")
        for gname in ("full-res decoder + heads", "deep decoder", "encoder"):
            rs = [r for r in rows if grp(r[0]) == gname]
            f.write(f"| {gname} | {max(r[2][0] for r in rs):.2e} | {min(r[2][1] for r in rs):.5f} | {max(r[3][0] for r in rs):.2e} | {max(r[4][0] for r in rs):.2e} | "
                    f"{max(r[5][0] for r in rs):.2e} | {max(r[6][0] for r in rs):.2e} |\n")
    print(open(out).read()[-1500:])


if __name__ == "__main__":
    main()
