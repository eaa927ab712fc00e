"""Per-kernel CUDA-event breakdown of one eager training step (every libunpp launch), B x 256 x 256."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet_nested4tiny_objects_keypoints_b200 as pkg
from unet_nested4tiny_objects_keypoints_b200 import fused, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
torch.manual_seed(0)
model = pkg.UNet_Nested().cuda().train()
step = fused.FusedTrainStep(model, B, S, S, device=torch.device("cuda", 0), seed=0)
step.x.copy_(torch.randn(B, 3, S, S))
step.target.copy_(torch.rand(B, 4, S, S))
for _ in range(3):
    step.step_device()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step.step_device()
e1.record()
torch.cuda.synchronize()
print("graph step ms %.4f  (%.0f img/s)" % (e0.elapsed_time(e1) / 10, B * 10 / e0.elapsed_time(e1) * 1e3))
agg = {}
reps = 3
for _ in range(reps):
    ops.trace = []
    torch.cuda._sleep(int(80e6))  # GPU spins ~40 ms while the host enqueues the whole step: events then bracket pure GPU time
    step._fwd_bwd(); step._update()
    torch.cuda.synchronize()
    for label, a, b, nbytes, flops in ops.trace:
        v = agg.setdefault(label, [0.0, 0, nbytes])
        v[0] += a.elapsed_time(b); v[1] += 1
    ops.trace = None
tot = sum(v[0] for v in agg.values()) / reps
print("sum of traced kernel times: %.4f ms" % tot)
groups = {}
for k, v in agg.items():
    g = groups.setdefault(k.split(" ")[0], [0.0, 0]); g[0] += v[0] / reps; g[1] += v[1] // reps
for k, g in sorted(groups.items(), key=lambda kv: -kv[1][0]):
    print("%-24s %4d launches %8.1f us" % (k, g[1], g[0] * 1e3))
print()
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    ms = v[0] / v[1]
    print("%-46s x%-3d %8.1f us each %8.1f us total  %s" % (k, v[1] // reps, ms * 1e3, v[0] / reps * 1e3, ("%.0f GB/s" % (v[2] / ms / 1e6)) if v[2] else ""))
