import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unet_nested4tiny_objects_keypoints_b200 import ops
dev = "cuda"
def run(cins, cout, H, N=32, reps=10):
    srcs = [torch.randn(N, H, H, c, device=dev).to(torch.bfloat16) for c in cins]
    dz = torch.randn(N, H, H, cout, device=dev).to(torch.bfloat16)
    cin = sum(cins)
    g = ops.wgrad_grid(cins, N, H, H, cout, 9)
    part = torch.empty(g * 9 * cin * cout, device=dev)
    for _ in range(2): ops.wgrad(srcs, N, H, H, dz, cout, 9, part)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): ops.wgrad(srcs, N, H, H, dz, cout, 9, part)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, g
print("UNPP_WGRAD_LEGACY =", os.environ.get("UNPP_WGRAD_LEGACY"))
for cins, cout, H in (([128], 128, 32), ([64], 128, 32), ([64, 64], 64, 64), ([64], 64, 64), ([32], 64, 64), ([16], 16, 256), ([16] * 2, 16, 256), ([16] * 3, 16, 256), ([16] * 5, 16, 256), ([16], 32, 128),
                      ([32], 32, 128), ([32] * 2, 32, 128), ([32] * 3, 32, 128)):
    ms, g = run(cins, cout, H)
    px = 32 * H * H
    print(cins, cout, H, "ms %.4f grid %d  algorithmic GB/s %.0f" % (ms, g, px * (sum(cins) + cout) * 2 / ms / 1e6), flush=True)
