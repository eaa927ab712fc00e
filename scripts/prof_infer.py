"""Two eager inference passes at B=128, 256x256 (+ arg-max) for ncu: every launch of a pass is visible."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet_nested4tiny_objects_keypoints_b200 as pkg
torch.manual_seed(0)
m = pkg.UNet_Nested().cuda().eval()
x = torch.randn(128, 3, 256, 256, device="cuda")
for _ in range(2):
    xy, val, heats = m.predict_keypoints(x, head=2)
torch.cuda.synchronize()
print("ok", xy.shape)
