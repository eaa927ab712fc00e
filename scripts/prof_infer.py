"""Two eager inference passes at B=128, 256x256 — the bench's default workload: three heads + one arg-max launch over the 12 planes
per image — for ncu: every launch of a pass is visible (the bench replays the same launches from a CUDA graph)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet_nested4tiny_objects_keypoints_b200 as pkg
from unet_nested4tiny_objects_keypoints_b200 import fused
torch.manual_seed(0)
m = pkg.UNet_Nested().cuda().eval()
B, S = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 256)
sess = fused.InferenceSession(m, B, S, S, head=(0, 1, 2), use_graph=False)  # the constructor runs the body twice (warm-up passes)
sess.xs[0].normal_()
with torch.no_grad():
    for _ in range(2):
        sess._body(0)
torch.cuda.synchronize()
print("ok", sess.out[0][0].shape)
