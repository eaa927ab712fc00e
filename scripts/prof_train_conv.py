"""One training-variant conv_tc configuration (2x2-blocked K16, mask + stats) for ncu source-level profiling."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unet_nested4tiny_objects_keypoints_b200 import ops
dev = "cuda"
N, H = 32, 256
src = torch.randn(N, H, H, 16, device=dev).to(torch.bfloat16)
wf = torch.randn(16, 16, 3, 3, device=dev) * 0.05
w, nt = ops.pack_weights_b2(wf, False, 16), ops.NTile(16, b2=True)
out = torch.empty(N, H, H, 16, dtype=torch.bfloat16, device=dev)
mask = torch.randn(N, H, H, 16, device=dev).to(torch.bfloat16)
g = ops.conv_grid([16], N, H, H, 16, nt, 9)
part = torch.empty(g, 2, 16, device=dev)
for _ in range(3):
    ops.conv([src], N, H, H, w, 16, nt, 9, out=out, relu_mask_src=mask, stats_partial=part)
torch.cuda.synchronize()
print("ok")
