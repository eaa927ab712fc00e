"""Experiment: time 2x2-blocked conv_tc launches (B=128, 256x256) with the UNPP_DBG role-disabling bits."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unet_nested4tiny_objects_keypoints_b200 import ops
dev = "cuda"
def run(nsrc, H=256, N=128, reps=5, b2=True):
    srcs = [torch.randn(N, H, H, 16, device=dev).to(torch.bfloat16) for _ in range(nsrc)]
    cin = 16 * nsrc
    wf = torch.randn(16, cin, 3, 3, device=dev) * 0.05
    w = ops.pack_weights_b2(wf, False, cin) if b2 else ops.pack_weights(wf, 0, 9, 16, 16, cin)
    out = torch.empty(N, H, H, 16, dtype=torch.bfloat16, device=dev)
    bias = torch.zeros(16, device=dev)
    for _ in range(2): ops.conv(srcs, N, H, H, w, 16, 16, 9, bias=bias, relu=True, out=out, b2=b2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): ops.conv(srcs, N, H, H, w, 16, 16, 9, bias=bias, relu=True, out=out, b2=b2)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for dbg in (0, 1, 2, 3, 4, 5, 6, 7):
    os.environ["UNPP_DBG"] = str(dbg)
    print("dbg", dbg, " b2: K16 %.3f K32 %.3f K64 %.3f   classic: K16 %.3f K64 %.3f" % (run(1), run(2), run(4), run(1, b2=False), run(4, b2=False)), flush=True)
