"""Training-variant conv_tc configurations at B=32 (one warm-up + one launch each) for ncu source-level profiling,
and a quick event timing of each when run without a profiler."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unet_nested4tiny_objects_keypoints_b200 import ops
dev = "cuda"
N = int(os.environ.get("PROF_B", "32"))
reps = int(os.environ.get("PROF_REPS", "1"))

def timeit(name, f):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    print("%-40s %8.1f us" % (name, e0.elapsed_time(e1) / reps * 1e3), flush=True)

def level0():
    H = 256
    src = torch.randn(N, H, H, 16, device=dev).to(torch.bfloat16)
    wf = torch.randn(16, 16, 3, 3, device=dev) * 0.05
    w, nt = ops.pack_weights_b2(wf, False, 16), ops.NTile(16, b2=True)
    out = torch.empty(N, H, H, 16, dtype=torch.bfloat16, device=dev)
    mask = torch.randn(N, H, H, 16, device=dev).to(torch.bfloat16)
    bias = torch.randn(16, device=dev)
    g = ops.conv_grid([16], N, H, H, 16, nt, 9)
    part = torch.empty(g, 2, 16, device=dev)
    hw, hb = torch.randn(4, 16, device=dev), torch.randn(4, device=dev)
    heat = torch.empty(N, 4, H, H, device=dev)
    dmask = torch.empty(N, H, H, dtype=torch.int16, device=dev); ops.dropout_mask(dmask, 0.4, 7)
    timeit("b2 K16 plain", lambda: ops.conv([src], N, H, H, w, 16, nt, 9, bias=bias, relu=True, out=out))
    timeit("b2 K16 +st", lambda: ops.conv([src], N, H, H, w, 16, nt, 9, bias=bias, out=out, stats_partial=part))
    timeit("b2 K16 +st+mask", lambda: ops.conv([src], N, H, H, w, 16, nt, 9, out=out, relu_mask_src=mask, stats_partial=part))
    timeit("b2 K16 +head (infer variant)", lambda: ops.conv([src], N, H, H, w, 16, nt, 9, bias=bias, relu=True, out=out, head=(hw, hb, heat, None, None, 1.0)))
    timeit("b2 K16 +head+dropmask (train variant)", lambda: ops.conv([src], N, H, H, w, 16, nt, 9, bias=bias, relu=True, out=out, head=(hw, hb, heat, None, dmask, 1 / 0.6)))

def level1():
    H = 128
    src = torch.randn(N, H, H, 32, device=dev).to(torch.bfloat16)
    wf = torch.randn(32, 32, 3, 3, device=dev) * 0.05
    nt = ops.pick_n_tile(32, 32, 9)
    w = ops.pack_weights(wf, 0, 9, 32, nt, 32)
    out = torch.empty(N, H, H, 32, dtype=torch.bfloat16, device=dev)
    mask = torch.randn(N, H, H, 32, device=dev).to(torch.bfloat16)
    g = ops.conv_grid([32], N, H, H, 32, nt, 9)
    part = torch.empty(g, 2, 32, device=dev)
    timeit("K32 N32 128 plain", lambda: ops.conv([src], N, H, H, w, 32, nt, 9, relu=True, out=out))
    timeit("K32 N32 128 +st", lambda: ops.conv([src], N, H, H, w, 32, nt, 9, out=out, stats_partial=part))
    timeit("K32 N32 128 +st+mask", lambda: ops.conv([src], N, H, H, w, 32, nt, 9, out=out, relu_mask_src=mask, stats_partial=part))

level0()
level1()
