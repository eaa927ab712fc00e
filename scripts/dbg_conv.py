"""Experiment: time single conv_tc launches (B=128) with the UNPP_DBG role-disabling bits."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unet_nested4tiny_objects_keypoints_b200 import ops
dev = "cuda"
def run(cins, cout, H, N=128, reps=5):
    srcs = [torch.randn(N, H, H, c, device=dev).to(torch.bfloat16) for c in cins]
    cin = sum(cins)
    nt = ops.pick_n_tile(cout, cin, 9)
    w = ops.pack_weights(torch.randn(cout, cin, 3, 3, device=dev) * 0.05, 0, 9, cout, nt, cin)
    out = torch.empty(N, H, H, cout, dtype=torch.bfloat16, device=dev)
    bias = torch.zeros(cout, device=dev)
    for _ in range(2): ops.conv(srcs, N, H, H, w, cout, nt, 9, bias=bias, relu=True, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): ops.conv(srcs, N, H, H, w, cout, nt, 9, bias=bias, relu=True, out=out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for dbg in (0, 1, 2, 3, 4, 5, 6, 7):
    os.environ["UNPP_DBG"] = str(dbg)
    print("dbg", dbg, " K16N16@256 %.3f  K64N16@256 %.3f  K32N32@128 %.3f  K64N64@64 %.3f  K128N128@32 %.3f" % (
        run([16], 16, 256), run([16] * 4, 16, 256), run([32], 32, 128), run([64], 64, 64), run([128], 128, 32)), flush=True)
