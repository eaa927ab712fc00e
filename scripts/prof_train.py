"""Two eager training steps at B=32, 256x256, MSE + AdamW (for ncu launch lists: every launch of a step is visible).
UNPP_WGRAD_STREAM=0 keeps every launch on one stream, so that `-s / -c` windows of ncu select whole steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet_nested4tiny_objects_keypoints_b200 as pkg
from unet_nested4tiny_objects_keypoints_b200 import fused
torch.manual_seed(0)
m = pkg.UNet_Nested().cuda().train()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
st = fused.FusedTrainStep(m, B, 256, 256, use_graph=False, loss="mse")
st.x.normal_(); st.target.uniform_()
for _ in range(2):
    st.step_device()
torch.cuda.synchronize()
print("loss", float(st.loss))
