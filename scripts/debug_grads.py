"""Debug helper: per-parameter gradient error table of the CUDA training path vs the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import unet_nested4tiny_objects_keypoints_b200 as pkg
from oracle import unetpp_oracle as O

B, H, W = (int(a) for a in (sys.argv[1:4] or (2, 64, 64)))
sd = O.synth_state_dict(seed=21)
g = torch.Generator().manual_seed(1)
x = torch.randn(B, 3, H, W, generator=g)
target = torch.rand(B, 4, H, W, generator=g)
m = pkg.UNet_Nested(); m.load_state_dict(sd); m = m.cuda().train(); m.drop_out.p = 0.0
outs = m(x.cuda())
loss = sum(F.mse_loss(o, target.cuda()) for o in outs) / 3
loss.backward(); torch.cuda.synchronize()
rl, routs, rg, _ = O.train_step_grads(sd, x, target, dropout_masks=None)
print("loss", float(loss), float(rl))
for o, r in zip(outs, routs):
    e = (o.detach().cpu() - r).abs(); print("heat err max %.3e mean %.3e" % (float(e.max()), float(e.mean())))
for k, p in m.named_parameters():
    r = rg[k].double(); gg = p.grad.cpu().double()
    sc = float(r.abs().max()); err = float((gg - r).abs().max())
    cos = float((gg * r).sum() / (gg.norm() * r.norm() + 1e-300))
    print("%-40s scale %.3e rel %.3e cos %.5f" % (k, sc, err / (sc + 1e-30), cos))

# calibration: the oracle itself under torch autocast(bf16) on the GPU (cuDNN bf16 kernels, fp32 master weights)
if len(sys.argv) > 4:
    from collections import OrderedDict
    params = OrderedDict((k, v.detach().clone().cuda().requires_grad_(True)) for k, v in sd.items() if v.dtype.is_floating_point and "running_" not in k)
    full = {k: v.cuda() for k, v in sd.items()}; full.update(params)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ao = O.forward(full, x.cuda(), training=True, dropout_masks=None)
    al = sum(F.mse_loss(o.float(), target.cuda()) for o in ao) / 3
    al.backward()
    print("---- calibration: torch autocast bf16 vs fp32 oracle")
    for k, p in params.items():
        r = rg[k].double(); gg = p.grad.cpu().double()
        sc = float(r.abs().max()); err = float((gg - r).abs().max())
        cos = float((gg * r).sum() / (gg.norm() * r.norm() + 1e-300))
        if "weight" in k and (".0." in k or "up.weight" in k): print("%-40s scale %.3e rel %.3e cos %.5f" % (k, sc, err / (sc + 1e-30), cos))
