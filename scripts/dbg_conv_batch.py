"""conv_tc time vs batch (fixed per-launch cost vs per-tile cost), inference variant and training variant."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unet_nested4tiny_objects_keypoints_b200 import ops
dev = "cuda"
def run(nsrc, N, H=256, reps=20, b2=True, train=False, cout=16, csrc=16):
    srcs = [torch.randn(N, H, H, csrc, device=dev).to(torch.bfloat16) for _ in range(nsrc)]
    cin = csrc * nsrc
    wf = torch.randn(cout, cin, 3, 3, device=dev) * 0.05
    if b2:
        w, nt = ops.pack_weights_b2(wf, False, cin), ops.NTile(16, b2=True)
    else:
        nt = ops.pick_n_tile(cout, cin, 9); w = ops.pack_weights(wf, 0, 9, cout, nt, cin)
    out = torch.empty(N, H, H, cout, dtype=torch.bfloat16, device=dev)
    bias = torch.zeros(cout, device=dev)
    kw = {}
    if train:
        g = ops.conv_grid([csrc] * nsrc, N, H, H, cout, nt, 9)
        kw = dict(stats_partial=torch.empty(g, 2, cout, device=dev), relu_mask_src=torch.ones_like(out))
    f = lambda: ops.conv(srcs, N, H, H, w, cout, nt, 9, bias=bias, relu=True, out=out, **kw)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for N in (2, 8, 32, 128):
    print("N=%3d  b2 K16 infer %.1f us  train %.1f us | b2 K48 infer %.1f train %.1f | classic 32->32@128 infer %.1f train %.1f | 128->128@32 infer %.1f" % (
        N, run(1, N), run(1, N, train=True), run(3, N), run(3, N, train=True), run(1, N, H=128, b2=False, cout=32, csrc=32), run(1, N, H=128, b2=False, train=True, cout=32, csrc=32),
        run(1, N, H=32, b2=False, cout=128, csrc=128)), flush=True)
