"""GPU-side fixed cost of one conv_tc launch: a CUDA graph of 40 identical launches at small batch sizes, with the role-disabling
UNPP_DBG bits (1 = epilogue skips its work, 2 = no MMAs, 4 = no TMA tile loads) set per process."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unet_nested4tiny_objects_keypoints_b200 import ops
dev = "cuda"
H = 256
wf = torch.randn(16, 16, 3, 3, device=dev) * 0.05
w, nt = ops.pack_weights_b2(wf, False, 16), ops.NTile(16, b2=True)
bias = torch.randn(16, device=dev)
res = []
for N in [int(v) for v in os.environ.get("FLOOR_B", "1,5,10,19,37,74").split(",")]:
    src = torch.randn(N, H, H, 16, device=dev).to(torch.bfloat16)
    out = torch.empty(N, H, H, 16, dtype=torch.bfloat16, device=dev)
    f = lambda: ops.conv([src], N, H, H, w, 16, nt, 9, bias=bias, relu=True, out=out)
    f(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        f()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        for _ in range(40):
            f()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    res.append("B=%d tiles/CTA=%.2f: %.2f us" % (N, N * 32 / 148, e0.elapsed_time(e1) / 200 * 1e3))
print("DBG=%s  " % os.environ.get("UNPP_DBG", "0") + " | ".join(res), flush=True)
