import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unet_nested4tiny_objects_keypoints_b200 import ops
dev = "cuda"
N, H = 32, 256
heat = torch.rand(N, 4, H, H, device=dev); target = torch.rand(N, 4, H, H, device=dev)
x = torch.randn(N, H, H, 16, device=dev).relu().to(torch.bfloat16)
mask = torch.empty(N, H, H, dtype=torch.int16, device=dev); ops.dropout_mask(mask, 0.4, 7)
hw = torch.randn(4, 16, device=dev)
dx = torch.empty(N, H, H, 16, dtype=torch.bfloat16, device=dev)
g = ops.head_bwd_grid(N, H, H)
part = torch.empty(g, 85, device=dev)
def t(**kw):
    f = lambda: ops.head_bwd(heat, kw.get("dheat"), kw.get("target"), 1e-6, x, kw.get("mask"), 1 / 0.6, hw, dx, part, loss_kind=kw.get("loss_kind", 0))
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10 * 1e3
print("grid", g)
print("mse + mask   %.1f us" % t(target=target, mask=mask))
print("mse no mask  %.1f us" % t(target=target))
print("upstream     %.1f us" % t(dheat=target, mask=mask))
print("focal + mask %.1f us" % t(target=target, mask=mask, loss_kind=1))
