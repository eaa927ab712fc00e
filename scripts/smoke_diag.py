import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet_nested4tiny_objects_keypoints_b200 as pkg
from oracle import unetpp_oracle as O
sd = O.synth_state_dict(seed=12)
model = pkg.UNet_Nested(); model.load_state_dict(sd); model = model.to("cuda:0").train(); model.drop_out.p = 0.0
g = torch.Generator().manual_seed(6)
x = torch.randn(2, 3, 32, 32, generator=g); target = torch.rand(2, 4, 32, 32, generator=g)
outs = model(x.cuda())
loss = sum(torch.nn.functional.mse_loss(o, target.cuda()) for o in outs) / 3
loss.backward(); torch.cuda.synchronize()
rl, _, rg, _ = O.train_step_grads(sd, x, target, dropout_masks=None)
print("loss", float(loss), float(rl))
rows = []
for k, p in model.named_parameters():
    ref = rg[k]
    err = float((p.grad.cpu() - ref).abs().max()) / (float(ref.abs().max()) + 1e-12)
    rows.append((err, k, float(ref.abs().max())))
for err, k, sc in sorted(rows, reverse=True)[:12]:
    print("%-44s err/max %.3e  scale %.3e" % (k, err, sc))
