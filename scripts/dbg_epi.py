"""Epilogue-only experiments (no TMA, no MMA): where does the epilogue time go?  B=128, 256x256, K16."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unet_nested4tiny_objects_keypoints_b200 import ops
dev = "cuda"
N, H = 128, 256
src = torch.randn(N, H, H, 16, device=dev).to(torch.bfloat16)
wf = torch.randn(16, 16, 3, 3, device=dev) * 0.05
out = torch.empty(N, H, H, 16, dtype=torch.bfloat16, device=dev)
bias = torch.zeros(16, device=dev)
def t(b2, dbg):
    os.environ["UNPP_DBG"] = str(dbg)
    if b2: w, nt = ops.pack_weights_b2(wf, False, 16), ops.NTile(16, b2=True)
    else: w, nt = ops.pack_weights(wf, 0, 9, 16, 16, 16), 16
    f = lambda: ops.conv([src], N, H, H, w, 16, nt, 9, bias=bias, relu=True, out=out)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10 * 1e3
for name, dbg in (("all roles", 0), ("epilogue only", 6), ("epi, no stores", 6 + 8), ("epi, no tmem", 6 + 16), ("epi, neither", 6 + 24), ("nothing", 7)):
    print("%-16s b2 %.1f us   classic %.1f us" % (name, t(True, dbg), t(False, dbg)), flush=True)
