"""Which training-epilogue feature costs what (2x2-blocked K16 conv, B=128, 256x256)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unet_nested4tiny_objects_keypoints_b200 import ops
dev = "cuda"
N, H = 128, 256
src = torch.randn(N, H, H, 16, device=dev).to(torch.bfloat16)
wf = torch.randn(16, 16, 3, 3, device=dev) * 0.05
w, nt = ops.pack_weights_b2(wf, False, 16), ops.NTile(16, b2=True)
out = torch.empty(N, H, H, 16, dtype=torch.bfloat16, device=dev)
other = torch.randn(N, H, H, 16, device=dev).to(torch.bfloat16)
g = ops.conv_grid([16], N, H, H, 16, nt, 9)
part = torch.empty(g, 2, 16, device=dev)
mean, istd = torch.zeros(16, device=dev), torch.ones(16, device=dev)
def t(**kw):
    f = lambda: ops.conv([src], N, H, H, w, 16, nt, 9, out=out, **kw)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10 * 1e3
print("plain            %.1f us" % t())
print("stats only       %.1f us" % t(stats_partial=part))
print("mask only        %.1f us" % t(relu_mask_src=other))
print("addend only      %.1f us" % t(addend=other))
print("mask+stats       %.1f us" % t(relu_mask_src=other, stats_partial=part))
print("mask+stats+aux   %.1f us" % t(relu_mask_src=other, stats_partial=part, stats_aux=other, aux_mean=mean, aux_istd=istd))
print("all four         %.1f us" % t(addend=other, relu_mask_src=other, stats_partial=part, stats_aux=other, aux_mean=mean, aux_istd=istd))
