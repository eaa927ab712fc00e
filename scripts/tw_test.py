import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unet_nested4tiny_objects_keypoints_b200 import ops
dev = "cuda"
def run_b2(nsrc, H=256, N=128, reps=5):
    srcs = [torch.randn(N, H, H, 16, device=dev).to(torch.bfloat16) for _ in range(nsrc)]
    cin = 16 * nsrc
    w = ops.pack_weights_b2(torch.randn(16, cin, 3, 3, device=dev) * 0.05, False, cin)
    out = torch.empty(N, H, H, 16, dtype=torch.bfloat16, device=dev); bias = torch.zeros(16, device=dev)
    f = lambda: ops.conv(srcs, N, H, H, w, 16, 16, 9, bias=bias, relu=True, out=out, b2=True)
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def run(cins, cout, H, N=128, reps=5):
    srcs = [torch.randn(N, H, H, c, device=dev).to(torch.bfloat16) for c in cins]
    cin = sum(cins); nt = ops.pick_n_tile(cout, cin, 9)
    w = ops.pack_weights(torch.randn(cout, cin, 3, 3, device=dev) * 0.05, 0, 9, cout, nt, cin)
    out = torch.empty(N, H, H, cout, dtype=torch.bfloat16, device=dev); bias = torch.zeros(cout, device=dev)
    f = lambda: ops.conv(srcs, N, H, H, w, cout, nt, 9, bias=bias, relu=True, out=out)
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for N in (128, 32):
    print("B=%d TW caps: b2=%s classic=%s | b2 K16 %.3f K32 %.3f K48 %.3f K64 %.3f | K16N32@128 %.3f K32N32@128 %.3f K64N32@128 %.3f K96N32@128 %.3f K64N64@64 %.3f K128N64@64 %.3f K128N128@32 %.3f" % (
        N, os.environ.get("UNPP_B2_TW"), os.environ.get("UNPP_TW"), run_b2(1, N=N), run_b2(2, N=N), run_b2(3, N=N), run_b2(4, N=N), run([16], 32, 128, N), run([32], 32, 128, N), run([32, 32], 32, 128, N),
        run([32] * 3, 32, 128, N), run([64], 64, 64, N), run([64, 64], 64, 64, N), run([128], 128, 32, N)), flush=True)
