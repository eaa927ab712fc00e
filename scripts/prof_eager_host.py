"""Host-side profile of the eager drop-in training step (model(x) -> loss -> backward -> optimizers.AdamW.step()): cProfile over 20 steps."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet_nested4tiny_objects_keypoints_b200 as pkg
from unet_nested4tiny_objects_keypoints_b200 import optimizers
torch.manual_seed(0)
m = pkg.UNet_Nested().cuda().train()
opt = optimizers.AdamW(m.parameters(), lr=3e-6, weight_decay=1e-4)
x = torch.randn(32, 3, 256, 256, device="cuda")
t = torch.rand(32, 4, 256, 256, device="cuda")
def one():
    opt.zero_grad()
    outs = m(x)
    loss = sum(torch.nn.functional.mse_loss(o, t) for o in outs) / 3
    loss.backward()
    opt.step()
for _ in range(3): one()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(20): one()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(22)
